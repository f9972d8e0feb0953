"""Times the device grain segmentation (K8) on the bench batch: 64 tracks x 300 s."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import melonix_b200 as m  # noqa: E402
from bench import gen_tracks_gpu  # noqa: E402

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = 48000 * 300
dev = torch.device("cuda", 0)
eng = m.Engine(0)
eng.use_torch_stream()
x = gen_tracks_gpu(torch, dev, nt, n, 0)
eng.upload_tracks_dev([x[i] for i in range(nt)])
cap = n // 751 + 1
gs = torch.zeros((nt, cap), dtype=torch.int32, device=dev)
gl = torch.zeros((nt, cap), dtype=torch.int32, device=dev)
cnt = torch.zeros(nt, dtype=torch.int32, device=dev)
for _ in range(2):
    eng.grain_segment_dev(gs, gl, cnt, cap)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 5
for _ in range(reps):
    eng.grain_segment_dev(gs, gl, cnt, cap)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
c = cnt.cpu().numpy()
print(f"grain segmentation: {nt} tracks x 300 s: {ms:.3f} ms per pass, {c.sum()} grains "
      f"({nt * n / ms / 1e6:.1f} G samples/s, {nt * n * 4.25 / ms / 1e6:.0f} GB/s algorithmic for the predicate pass)")
# host mirror on one track for scale
from melonix_b200 import hostlib as H  # noqa: E402
xh = x[0].cpu().numpy()
t0 = time.perf_counter()
hs, hl = H.grain_segment(xh)
t1 = time.perf_counter()
print(f"host C++ mirror, 1 track, 1 thread: {1e3 * (t1 - t0):.1f} ms ({hs.size} grains; device found {int(c[0])})")
