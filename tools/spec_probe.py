import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import melonix_b200 as m
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
hop = int(sys.argv[2]) if len(sys.argv) > 2 else N // 4
eng = m.Engine(0); eng.use_torch_stream()
n = 48000 * 300
x = (torch.randn(n, device="cuda") * 0.1).float()
eng.upload_tracks_dev([x])
F = (n + hop - 1) // hop
buf = torch.empty((F, N // 2), dtype=torch.float32, device="cuda")
for _ in range(3): eng.spec_frames_dev(0, N, hop, 0, F, buf)
torch.cuda.synchronize()
eng.profile_enable(True); eng.profile_read()
for _ in range(5): eng.spec_frames_dev(0, N, hop, 0, F, buf)
ms, ln = eng.profile_read()["spec"]
fps = 5 * F / (ms * 1e-3)
print(f"spec N={N} hop={hop}: {fps / 1e6:.1f} M frames/s, {fps * (4 * hop + 2 * N) / 1e9:.0f} GB/s algorithmic, {ms / ln:.3f} ms/launch")
