"""torchrun entry (one process per GPU): phase-vocodes one long mono track sharded by time range
across the ranks (NCCL seam exchange + phase-carry all-gather) and checks on rank 0 that the
gathered result is bit-identical to the unsharded single-GPU run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29541 tools/sharded_check.py [seconds] [fftN]
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import melonix_b200 as m  # noqa: E402
import signals as S  # noqa: E402
from melonix_b200 import dist as D  # noqa: E402


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    H = N // 4
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    x = S.vibrato_tone(seconds, seed=1234)           # every rank can rebuild the file; it only uses its slice
    rate = m.semitone_ratio(3.0)
    shard = D.plan_time_shards(x.size, N, H, world)[rank]
    own = torch.from_numpy(x[shard.own_lo:shard.own_hi]).cuda()
    eng = m.Engine(local)
    y_own, peak_own, f0_own = D.run_time_sharded(eng, own, x.size, N, H, rate)     # NCCL behind the C ABI
    y_t, peak_t, f0_t = D.run_time_sharded_torch(eng, own, x.size, N, H, rate)    # torch.distributed collectives
    torch.cuda.synchronize()
    same_paths = bool(torch.equal(y_own, y_t) and torch.equal(peak_own, peak_t) and torch.equal(f0_own, f0_t))
    # gather on rank 0
    sizes = [s.own_hi - s.own_lo for s in D.plan_time_shards(x.size, N, H, world)]
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=torch.float32, device="cuda")
    buf[:y_own.numel()] = y_own
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    ok = True
    if rank == 0:
        y = np.concatenate([g[:sz].cpu().numpy() for g, sz in zip(gathered, sizes)])
        eng.upload_tracks([x])
        full = eng.pv_run(N, H, rate)[0]
        ok = bool(np.array_equal(y, full["y"])) and same_paths
        print(f"sharded_check world={world} N={N} seconds={seconds}: bitwise_equal={ok} "
              f"max|diff|={float(np.abs(y - full['y']).max()):.3e} seam_payload_floats=({N},{3 * H})", flush=True)
    flag = torch.tensor([1 if (ok and same_paths) else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.broadcast(flag, 0)
    if rank == 0 and int(flag.item()) != 1:
        print("sharded_check: some rank disagrees (C-ABI NCCL path vs torch path, or vs unsharded)", flush=True)
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
