import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import melonix_b200 as m
from melonix_b200 import dist as D
import signals as S
eng = m.Engine(0)
N, H = 4096, 1024
x = S.vibrato_tone(20.0, seed=31)
r = m.semitone_ratio(3.0)
eng.upload_tracks([x])
full = eng.pv_run(N, H, r)[0]
xd = torch.from_numpy(x).cuda()
eng.use_torch_stream()
for world in (2,):
    shards = D.plan_time_shards(x.size, N, H, world)
    carry = torch.zeros(N // 2 + 1, dtype=torch.int64, device="cuda")
    for s in shards:
        win = xd[s.need_lo:s.need_hi].contiguous()
        eng.upload_tracks_dev([win])
        tot = torch.zeros(N // 2 + 1, dtype=torch.int32, device="cuda")
        eng.pv_phase_totals_dev(N, H, r, [tot], frame_begin=s.local_frame_begin, frame_end=s.local_frame_end, wave_mib=-1)
        c32 = torch.where(carry >= 2 ** 31, carry - 2 ** 32, carry).to(torch.int32)
        yd = torch.zeros_like(win)
        pk = torch.zeros(D.num_frames(win.numel(), H), dtype=torch.int32, device="cuda")
        eng.pv_run_dev(N, H, r, [yd], [pk], None, frame_begin=s.local_frame_begin, frame_end=s.local_frame_end, phase_in=[c32], wave_mib=-1)
        torch.cuda.synchronize()
        y = yd[s.left_halo:s.left_halo + (s.own_hi - s.own_lo)].cpu().numpy()
        ref = full["y"][s.own_lo:s.own_hi]
        d = np.nonzero(y != ref)[0]
        print(s, "ndiff", d.size, "first hop", (d[0] // H if d.size else -1), "last hop", (d[-1] // H if d.size else -1), "nhops", (s.own_hi - s.own_lo) // H, "maxabs", float(np.abs(y - ref).max()))
        pkr = pk[s.local_frame_begin:s.local_frame_end].cpu().numpy()
        print("   peak eq", np.array_equal(pkr, full["peak"][s.frame_begin:s.frame_end]))
        carry = (carry + (tot.to(torch.int64) & 0xFFFFFFFF)) & 0xFFFFFFFF
        if s.rank == 1:
            print("   peak mismatch idx", np.nonzero(pkr != full["peak"][s.frame_begin:s.frame_end])[0][:20], "of", pkr.size)
            # same window, whole-track run (frame_begin=0): tail behaviour without sharding params
            eng.upload_tracks_dev([win]); 
            yy = torch.zeros_like(win); pk2 = torch.zeros_like(pk)
            eng.pv_run_dev(N, H, r, [yy], [pk2], None, wave_mib=-1)
            torch.cuda.synchronize()
            print("   whole-window run: nan count", int(torch.isnan(yy).sum()), "peak tail", pk2[-6:].tolist(), "ref tail", full["peak"][-6:].tolist())
            yy2 = torch.zeros_like(win); pk3 = torch.zeros_like(pk)
            eng.pv_run_dev(N, H, r, [yy2], [pk3], None, frame_begin=4, frame_end=pk.numel(), wave_mib=-1)
            torch.cuda.synchronize()
            print("   fb=4 no phase_in: nan count", int(torch.isnan(yy2).sum()), "peak tail", pk3[-6:].tolist())
            nanidx = torch.nonzero(torch.isnan(yd)).flatten()
            print("   nan idx (hops)", (nanidx[:5] // H).tolist(), (nanidx[-5:] // H).tolist(), nanidx.numel())
