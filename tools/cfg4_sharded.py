"""BASELINE configs[3] shape: one long "stereo" (= two planar mono tracks) file, 4096-FFT / 1024-hop,
sharded by contiguous time range across the ranks, with the NCCL seam exchange and the phase-carry
all-gather inside the timed region.  Rank 0 also runs the unsharded job on its own GPU (it fits) and
checks bit-for-bit equality.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port 29551 tools/cfg4_sharded.py [seconds=7200]
"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import melonix_b200 as m  # noqa: E402
from bench import gen_tracks_gpu  # noqa: E402
from melonix_b200 import dist as D  # noqa: E402


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 7200.0
    N, H, FS = 4096, 1024, 48000
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    n = int(seconds * FS)
    rate = m.semitone_ratio(3.0)
    shard = D.plan_time_shards(n, N, H, world)[rank]
    # every rank synthesises only its own slice of the two channels (seeds 1234 / 1235, cfg-2 signal)
    full = gen_tracks_gpu(torch, dev, 2, n, 0) if world == 1 or rank == 0 else None
    if rank == 0:
        pieces = [[full[c, s.own_lo:s.own_hi].contiguous() for c in range(2)] for s in D.plan_time_shards(n, N, H, world)]
    owns = [torch.empty(shard.own_hi - shard.own_lo, dtype=torch.float32, device=dev) for _ in range(2)]
    for c in range(2):   # scatter the file (setup, outside the timed region)
        if rank == 0:
            owns[c].copy_(pieces[0][c])
            for r in range(1, world):
                dist.send(pieces[r][c], r)
        else:
            dist.recv(owns[c], 0)
    eng = m.Engine(local)
    res = D.run_time_sharded(eng, owns, n, N, H, rate)       # warm-up (allocations, tables)
    torch.cuda.synchronize()
    dist.barrier()
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        res = D.run_time_sharded(eng, owns, n, N, H, rate)
    torch.cuda.synchronize()
    dist.barrier()
    el = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(el, op=dist.ReduceOp.MAX)
    F = (n + H - 1) // H
    # gather for the bitwise check
    sizes = [s.own_hi - s.own_lo for s in D.plan_time_shards(n, N, H, world)]
    ok = True
    for c in range(2):
        buf = torch.zeros(max(sizes), dtype=torch.float32, device=dev)
        buf[:sizes[rank]] = res[c][0]
        gathered = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, gathered, 0)
        if rank == 0:
            y = torch.cat([g[:sz] for g, sz in zip(gathered, sizes)])
            eng.upload_tracks_dev([full[c]])
            ref = torch.zeros(n, dtype=torch.float32, device=dev)
            eng.pv_run_dev(N, H, rate, [ref], wave_mib=-1)
            torch.cuda.synchronize()
            ok = ok and bool(torch.equal(y, ref))
            del y, ref, gathered
    if rank == 0:
        print(json.dumps(dict(config="configs[3]: 2 planar tracks, 4096-FFT/1024-hop, time-range sharded", seconds=seconds,
                              n_gpus=world, frames=2 * F, ms=float(el.item()) * 1e3, frames_per_s=2 * F / float(el.item()),
                              bitwise_equal_to_unsharded=ok, seam_payload_floats_per_track=[N, 3 * H],
                              phase_carry_words_per_track=N // 2 + 1)), flush=True)
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
