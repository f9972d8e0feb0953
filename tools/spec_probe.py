"""Spec STFT probe for the ncu capture: `spec_probe.py N hop [all]` -- one 300 s track per launch, or with
`all` sixteen tracks in ONE batched launch (mlx_spec_frames_all_dev)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import melonix_b200 as m
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
hop = int(sys.argv[2]) if len(sys.argv) > 2 else N // 4
batched = len(sys.argv) > 3 and sys.argv[3] == "all"
nt = 16 if batched else 1
eng = m.Engine(0); eng.use_torch_stream()
n = 48000 * 300
xs = [(torch.randn(n, device="cuda") * 0.1).float() for _ in range(nt)]
eng.upload_tracks_dev(xs)
F = (n + hop - 1) // hop
bufs = [torch.empty((F, N // 2), dtype=torch.float32, device="cuda") for _ in range(nt)]
def run():
    if batched:
        eng.spec_frames_all_dev(N, hop, bufs)
    else:
        eng.spec_frames_dev(0, N, hop, 0, F, bufs[0])
for _ in range(3): run()
torch.cuda.synchronize()
eng.profile_enable(True); eng.profile_read()
for _ in range(5): run()
ms, ln = eng.profile_read()["spec"]
fps = 5 * nt * F / (ms * 1e-3)
print(f"spec N={N} hop={hop} tracks/launch={nt}: {fps / 1e6:.1f} M frames/s, {fps * (4 * hop + 2 * N) / 1e9:.0f} GB/s algorithmic, {ms / ln:.3f} ms/launch")
