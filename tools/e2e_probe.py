"""e2e probe: mlx_pv_process_host_fmt on the bench batch (64 x 300 s, 2048/512) from pinned host memory,
int16 or float32 on the wire, for several track-group sizes (MLX_PV_HOST_GROUP) -- plus the pipeline
timeline (MLX_TRACE) and a pure-copy floor (the same bytes both ways, no kernels).

    python tools/e2e_probe.py [i16|f32] [groups...]
"""
import os
import sys
import time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import melonix_b200 as m
fmt = sys.argv[1] if len(sys.argv) > 1 else "i16"
groups = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8, 16]
dt = torch.int16 if fmt == "i16" else torch.float32
nt, n = 64, 14_400_000
F = (n + 511) // 512
eng = m.Engine(0)
hx = torch.empty((nt, n), dtype=dt, pin_memory=True)
if dt == torch.int16:
    hx.random_(-3000, 3000)
else:
    hx.normal_(0, 0.1)
hy = torch.empty((nt, n), dtype=dt, pin_memory=True)
hp = torch.empty((nt, F), dtype=torch.int32, pin_memory=True)
hf = torch.empty((nt, F), dtype=torch.float32, pin_memory=True)
r = m.semitone_ratio(3.0)
ins = [hx[i] for i in range(nt)]
outs = ([hy[i] for i in range(nt)], [hp[i] for i in range(nt)], [hf[i] for i in range(nt)])
def go():
    eng.pv_process_host(ins, 2048, 512, r, *outs, wave_mib=-1)
# pure-copy floor: the same bytes in both directions at once, nothing else
d_in = torch.empty((nt, n), dtype=dt, device="cuda"); d_out = torch.empty((nt, n), dtype=dt, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s1):
        d_in.copy_(hx, non_blocking=True)
    with torch.cuda.stream(s2):
        hy.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); el = time.perf_counter() - t0
print(f"copy floor ({fmt}): {el * 1e3:.1f} ms for {hx.numel() * hx.element_size() / 1e9:.2f} GB each way "
      f"= {hx.numel() * hx.element_size() / el / 1e9:.1f} GB/s per direction")
del d_in, d_out
for g in groups:
    os.environ["MLX_PV_HOST_GROUP"] = str(g)
    go()
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter(); go(); best = min(best, time.perf_counter() - t0)
    print(f"group {g:2d}: {best * 1e3:.1f} ms per step = {nt * F / best / 1e6:.1f} M frames/s")
os.environ["MLX_PV_HOST_GROUP"] = "4"; os.environ["MLX_TRACE"] = "1"
go()
