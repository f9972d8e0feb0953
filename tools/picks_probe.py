"""Times the device min/max pyramid (K9) on the bench batch: 64 tracks x 300 s."""
import sys
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
off = eng.picks_layout(n)
buf = torch.zeros((nt, int(off[-1]), 2), dtype=torch.float32, device=dev)
bufs = [buf[t] for t in range(nt)]
for _ in range(2):
    eng.picks_build_all_dev(bufs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    eng.picks_build_all_dev(bufs)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
algo = nt * (4 * n + 8 * int(off[-1]))
print(f"min/max pyramid: {nt} tracks x 300 s ({len(off) - 1} levels): {ms:.3f} ms per pass, "
      f"{algo / ms / 1e6:.0f} GB/s algorithmic (4 B read + 8 B written per sample)")
