"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per CUDA source line."""
import csv
import sys


def num(x):
    try:
        return int(x)
    except Exception:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, out = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split('/')[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 8 and r[0].isdigit():
        out.append((cur, int(r[0]), r[1].strip()[:100], num(r[6]), num(r[7])))
ti = sum(o[4] for o in out) or 1
ts = sum(o[3] for o in out) or 1
print("total warp instr", ti, "samples", ts)
out.sort(key=lambda o: -(o[4] / ti + o[3] / ts))
for o in out[:top]:
    print(f"{o[0]:16s}:{o[1]:4d} inst {100 * o[4] / ti:5.2f}% samp {100 * o[3] / ts:5.2f}%  {o[2]}")
