import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import melonix_b200 as m
nt, n = 64, 14_400_000
F = (n + 511) // 512
eng = m.Engine(0)
hx = torch.empty((nt, n), dtype=torch.float32, pin_memory=True); hx.normal_(0, 0.1)
hy = torch.empty((nt, n), dtype=torch.float32, pin_memory=True)
hp = torch.empty((nt, F), dtype=torch.int32, pin_memory=True)
hf = torch.empty((nt, F), dtype=torch.float32, pin_memory=True)
r = m.semitone_ratio(3.0)
ins = [hx[i] for i in range(nt)]
def go(with_peak=True):
    eng.pv_process_host(ins, 2048, 512, r, [hy[i] for i in range(nt)], [hp[i] for i in range(nt)] if with_peak else None,
                        [hf[i] for i in range(nt)] if with_peak else None, wave_mib=-1)
go()
for rep in range(2):
    eng.profile_enable(True); eng.profile_read()
    t0 = time.perf_counter(); go(); el = time.perf_counter() - t0
    prof = eng.profile_read(); eng.profile_enable(False)
    print(f"e2e {el * 1e3:.1f} ms; kernel sums:", {k: (round(v[0], 1), v[1]) for k, v in prof.items() if v[1]})
t0 = time.perf_counter(); go(False); print(f"no peak/f0 copies: {(time.perf_counter() - t0) * 1e3:.1f} ms")
s = torch.cuda.Stream()
eng.set_stream(s.cuda_stream)
go(); torch.cuda.synchronize()
t0 = time.perf_counter(); go(); print(f"compute on a non-default stream: {(time.perf_counter() - t0) * 1e3:.1f} ms")
