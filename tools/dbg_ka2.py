"""Debug aid: where do K_A2 (MLX_PV_KA2=1) and the general analysis kernel differ?  Compares the staged
records (every bin of every frame) and the phase totals of both for one track."""
import os
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import melonix_b200 as m
import signals as S
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
semis = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
H = N // 4
xs = [S.vibrato_tone(4.0, seed=71), S.vibrato_tone(1.3, seed=72), np.zeros(700, np.float32), S.two_tone(2.0)]
r = m.semitone_ratio(semis)
eng = m.Engine(0); eng.use_torch_stream()
nb = N // 2 + 1
res = {}
for mode in ("general", "ka2"):
    os.environ.pop("MLX_PV_KA2", None); os.environ.pop("MLX_PV_NO_KA2", None)
    os.environ["MLX_PV_KA2" if mode == "ka2" else "MLX_PV_NO_KA2"] = "1"
    eng.upload_tracks(xs)
    tots = [torch.zeros(nb, dtype=torch.int32, device="cuda") for _ in xs]
    eng.pv_phase_totals_dev(N, H, r, tots, wave_mib=-1)
    torch.cuda.synchronize()
    t_a = [t.cpu().numpy().copy() for t in tots]
    eng.pv_analyze_dev(N, H, r, tots)
    out = []
    for t, x in enumerate(xs):
        F = (x.size + H - 1) // H
        sm = torch.empty((F, nb), dtype=torch.float32, device="cuda"); ph = torch.empty((F, nb), dtype=torch.int32, device="cuda")
        eng.pv_stage_export_dev(t, 0, F, sm, ph)
        torch.cuda.synchronize()
        out.append((sm.cpu().numpy(), ph.cpu().numpy()))
    res[mode] = (t_a, [t.cpu().numpy() for t in tots], out)
for t in range(len(xs)):
    ta0, tb0, (s0, p0) = res["general"][0][t], res["general"][1][t], res["general"][2][t]
    ta1, tb1, (s1, p1) = res["ka2"][0][t], res["ka2"][1][t], res["ka2"][2][t]
    ds = np.argwhere(s0.view(np.int32) != s1.view(np.int32)); dp = np.argwhere(p0 != p1)
    print("track", t, "frames", p0.shape[0], "| totals(phase_totals) differ at bins", np.nonzero(ta0 != ta1)[0][:10],
          "| totals(analyze) differ at", np.nonzero(tb0 != tb1)[0][:10], "| smag diffs", len(ds), "phase diffs", len(dp),
          "bins", np.unique(dp[:, 1])[:10] if len(dp) else [])
    print("   analyze totals == last frame phase:", np.array_equal(tb0, p0[-1]), np.array_equal(tb1, p1[-1]),
          "| phase_totals == analyze totals:", np.array_equal(ta0, tb0), np.array_equal(ta1, tb1))
    for f, j in dp[:3]:
        print("   frame", f, "bin", j, "general", p0[f, j], s0[f, j], "ka2", p1[f, j], s1[f, j])
