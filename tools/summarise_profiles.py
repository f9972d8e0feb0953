"""Turns the raw artefacts tools/gpu_final.sh left in gpurun_out/ into the tracked summaries under
profiles/ (bench lines, ncu launch list, per-kernel summary of the full captures, DRAM traffic).

    python tools/summarise_profiles.py r1
"""
import csv
import json
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
src = ROOT / "gpurun_out"
prof = ROOT / "profiles"

# ---- bench lines and secondary sweeps
for a, b in (("final_bench_n1.json", f"{tag}_bench_n1.json"), ("final_bench_ref.json", f"{tag}_bench_reference_arm.json"),
             ("final_extra.json", f"{tag}_extra.json"), ("final_seg_probe.log", f"{tag}_grain_segment_probe.txt"), ("final_picks_probe.log", f"{tag}_picks_probe.txt"), ("final_grain_probe.log", f"{tag}_grain_render_probe.txt"),
             ("final_pytest.log", f"{tag}_pytest_gpu.log"), ("final_smoke.log", f"{tag}_smoke.log")):
    if (src / a).exists():
        txt = (src / a).read_text()
        if a.endswith(".json") and not txt.lstrip().startswith("{"):  # multi-rank runs: keep the JSON line only
            txt = [ln for ln in txt.splitlines() if ln.startswith("{")][-1] + "\n"
        (prof / b).write_text(txt)

# ---- launch list
launches = src / "final_launches_bench.csv"
rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
ix = {h: i for i, h in enumerate(rows[0])}
(prof / f"{tag}_launches_bench.csv").write_text(launches.read_text())
tot = {}
for r in rows[1:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    tot.setdefault(name, []).append(float(r[ix["Metric Value"]]) / 1e6)
allms = sum(sum(v) for v in tot.values())
print("| kernel | launches | avg ms (ncu, serialised) | share |\n|---|---|---|---|")
for k, v in tot.items():
    print(f"| `{k}` | {len(v)} | {sum(v) / len(v):.3f} | {100 * sum(v) / allms:.1f} % |")

# ---- full captures (raw pages exported on the box)
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        # the L1 / shared-memory data pipe: ONE 128-byte wavefront per cycle and SM -- the resource that binds
        # the FFT kernels (K_S 87 %, K1r 82 %, K_A 72 % of its peak), not DRAM and not the issue slots
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct"]
cols, table, units_of = [], {}, {}
traffic = {}
l1pipe = {}
for cap in ("pv", "spec", "seg", "picks"):
    p = src / f"final_prof_{cap}.raw.csv"
    if not p.exists():
        continue
    rr = list(csv.reader(open(p)))
    h, units = rr[0], rr[1]
    stall = [x for x in h if "smsp__average_warps_issue_stalled" in x and "per_issue_active" in x and "not_issued" not in x]
    for r in rr[2:]:
        name = r[h.index("Kernel Name")]
        short = name.replace("void ", "").split("(")[0][:44]
        cols.append(short)
        for k in keep + stall:
            if k in h:
                val, unit = r[h.index(k)], units[h.index(k)]
                scale = {"us": ("ms", 1e-3), "ns": ("ms", 1e-6), "s": ("ms", 1e3), "Gbyte": ("Mbyte", 1e3),
                         "Kbyte": ("Mbyte", 1e-3), "byte": ("Mbyte", 1e-6)}
                if unit in scale and not k.startswith("launch__"):  # one unit per metric across the captures
                    unit, val = scale[unit][0], f"{float(val) * scale[unit][1]:.6f}"
                table.setdefault(k, {})[short] = val
                units_of[k] = unit

        def nbytes(k):
            i = h.index(k)
            return float(r[i]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[i]]
        if cap == "pv":
            key = "pv_analyze" if "analyze" in name else "pv_scan" if "scan" in name else "pv_synth"
            traffic[key] = nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")
            k1 = "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"
            if k1 in h:
                l1pipe[key] = float(r[h.index(k1)])
with open(prof / f"{tag}_ncu_full_summary.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + cols)
    for k, v in table.items():
        w.writerow([k, units_of[k]] + [v.get(c, "") for c in cols])
json.dump(dict(source=f"ncu --set full --clock-control none, bench.py default workload (64 tracks x 300 s, 2048/512), "
                      f"one launch each; profiles/{tag}_ncu_full_summary.csv",
               per_kernel_bytes_per_launch=traffic, path_bytes_per_step=sum(traffic.values()),
               l1_data_pipe_pct_of_peak=l1pipe,
               algorithmic_bytes_per_step=4104 * 1800000), open(prof / "roofline_traffic.json", "w"), indent=1)
print(json.dumps(traffic))
