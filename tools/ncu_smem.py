"""Shared-memory wavefronts per CUDA source line from `ncu --page source --csv` (L1 Wavefronts Shared,
ideal and excessive = bank conflicts):   python tools/ncu_smem.py X.source.csv [kernel-substring] [units]
`units` (e.g. the number of frames of the launch) turns the sums into wavefronts per unit."""
import collections
import csv
import sys


def main():
    rows = csv.reader(open(sys.argv[1]))
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    cur_file = cur_fun = hdr = None
    per = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0, 0, ""]))
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) == 2 and r[0] == "Function Name":
            cur_fun = r[1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            iw, ii, ie = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal"), hdr.index(
                "L1 Wavefronts Shared Excessive")
            continue
        if hdr and len(r) > ie and r[0].isdigit() and cur_fun and want in cur_fun:
            e = per[cur_fun][(cur_file, int(r[0]))]
            for j, i in enumerate((iw, ii, ie)):
                try:
                    e[j] += int(r[i])
                except ValueError:
                    pass
            e[3] = r[1].strip()[:90]
    for fun, lines in per.items():
        tw = sum(v[0] for v in lines.values())
        te = sum(v[2] for v in lines.values())
        if not tw:
            continue
        print(f"== {fun[:90]}\n   wavefronts {tw / units:.1f}, excessive {te / units:.1f} per unit")
        for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:24]:
            if v[0]:
                print(f"   {f:16s}:{ln:4d} wavefronts {v[0] / units:8.1f} ideal {v[1] / units:8.1f} excess {v[2] / units:7.1f}  {v[3]}")


if __name__ == "__main__":
    main()
