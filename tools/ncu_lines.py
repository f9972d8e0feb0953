"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per kernel and CUDA
source line: share of executed warp instructions and of stall samples.

    python tools/ncu_lines.py gpurun_out/prof_pv.source.csv [kernel-substring] [top]
"""
import collections
import csv
import sys


def num(x):
    try:
        return int(x)
    except Exception:
        return 0


def main():
    rows = csv.reader(open(sys.argv[1]))
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    cur_file, cur_fun, hdr = None, None, None
    per = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0, ""]))
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) == 2 and r[0] == "Function Name":
            cur_fun = r[1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
            continue
        if hdr and len(r) > 8 and r[0].isdigit() and cur_fun and want in cur_fun:
            e = per[cur_fun][(cur_file, int(r[0]))]
            e[0] += num(r[i_s])
            e[1] += num(r[i_i])
            e[2] = r[1].strip()[:96]
    for fun, lines in per.items():
        ts = sum(v[0] for v in lines.values()) or 1
        ti = sum(v[1] for v in lines.values()) or 1
        print(f"== {fun[:100]}\n   warp instructions {ti}, samples {ts}")
        byfile = collections.Counter()
        for (f, _), v in lines.items():
            byfile[f] += v[1]
        print("   by file:", {f: f"{100 * c / ti:.1f}%" for f, c in byfile.most_common(6)})
        for (f, ln), v in sorted(lines.items(), key=lambda kv: -(kv[1][1] / ti + kv[1][0] / ts))[:top]:
            print(f"   {f:18s}:{ln:4d} inst {100 * v[1] / ti:5.2f}% samp {100 * v[0] / ts:5.2f}%  {v[2]}")


if __name__ == "__main__":
    main()
