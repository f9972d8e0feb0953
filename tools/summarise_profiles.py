"""Turns the raw artefacts a GPU run left in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarise_profiles.py r1 gpurun_out/prof_r1_final.ncu-rep gpurun_out/r1_launches_bench.csv
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag, rep, launches = sys.argv[1], Path(sys.argv[2]), Path(sys.argv[3])
prof = ROOT / "profiles"

# ---- launch list
rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
(prof / f"{tag}_launches_bench.csv").write_text(launches.read_text())
tot = {}
for r in rows[1:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    tot.setdefault(name, []).append(float(r[ix["Metric Value"]]) / 1e6)
allms = sum(sum(v) for v in tot.values())
launch_table = [(k, len(v), sum(v) / len(v), 100 * sum(v) / allms) for k, v in tot.items()]

# ---- full capture
raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct"]
stall = [x for x in h if "smsp__average_warps_issue_stalled" in x and "per_issue_active" in x]
with open(prof / f"{tag}_ncu_full_summary.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [r[h.index("Kernel Name")][:48] for r in rr[2:]])
    for k in keep + stall:
        if k in h:
            i = h.index(k)
            w.writerow([k, units[i]] + [r[i] for r in rr[2:]])


def nbytes(r, k):
    i = h.index(k)
    return float(r[i]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[i]]


traffic = {}
for r in rr[2:]:
    name = r[h.index("Kernel Name")]
    key = "pv_analyze" if "analyze" in name else "pv_scan" if "scan" in name else "pv_synth"
    traffic[key] = nbytes(r, "dram__bytes_read.sum") + nbytes(r, "dram__bytes_write.sum")
json.dump(dict(source=f"ncu --set full --clock-control none, bench.py default workload (64 tracks x 300 s, 2048/512), "
                      f"one launch each; profiles/{tag}_ncu_full_summary.csv",
               per_kernel_bytes_per_launch=traffic, path_bytes_per_step=sum(traffic.values()),
               algorithmic_bytes_per_step=4104 * 1800000), open(prof / "roofline_traffic.json", "w"), indent=1)
for k, n, avg, share in launch_table:
    print(f"| `{k}` | {n} | {avg:.2f} | {share:.1f} % |")
print(json.dumps(traffic))
