"""torchrun entry (one process per GPU): the two sharding modes that need NO collective on the data path
 -- Spec job-list sharding (read-only sample halo uploaded with the shard) and
grain output-range sharding (replicated source track, disjoint output rows).  Every rank computes its
block on its own GPU; rank 0 gathers the blocks (check only) and compares with the unsharded run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29561 tools/shard_nocoll_check.py
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
from melonix_b200 import hostlib as H  # noqa: E402


def gather_np(a, rank, world):
    """variable-length gather of a numpy array to rank 0 (check path only; shipped as bytes: the NCCL
    process group does not take int16 tensors)."""
    a = np.ascontiguousarray(a)
    t = torch.from_numpy(a.reshape(-1).view(np.uint8)).cuda()
    shapes = [None] * world
    dist.all_gather_object(shapes, tuple(a.shape))
    out = []
    for r in range(world):
        if rank == 0:
            if r == 0:
                out.append(a)
            else:
                nbytes = int(np.prod(shapes[r])) * a.dtype.itemsize
                buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
                if nbytes:
                    dist.recv(buf, r)
                out.append(buf.cpu().numpy().view(a.dtype).reshape(shapes[r]))
        elif rank == r and t.numel():
            dist.send(t, 0)
    return out


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    eng = m.Engine(local)
    ok = True
    # --- Spec job list
    x = S.vibrato_tone(4.0, seed=5)
    for N, hop in ((1024, 256), (32768, 375)):
        jobs = S.regular_jobs(x.size, hop)
        own, rows = D.run_spec_sharded(eng, x, jobs, N, world, rank)
        parts = gather_np(rows, rank, world)
        if rank == 0:
            eng.upload_tracks([x])
            full = eng.spec_batch(0, N, jobs)
            ok = ok and bool(np.array_equal(np.concatenate(parts), full))
    # --- grain export
    y = S.two_tone(8.0)
    markers = [(10, 0, 0, 3.0), (y.size - 10, 0, 0, 3.0)]
    gs, gl = H.grain_segment(y)
    sch = H.export_schedule(y, 48000, markers, gs, gl)
    eng.upload_tracks([y])
    _, pcm, pcm16 = D.run_grain_sharded(eng, 0, sch, world, rank)
    parts = gather_np(pcm16, rank, world)
    if rank == 0:
        _, full16 = H.export_wav(eng, 0, y, 48000, markers)
        ok = ok and bool(np.array_equal(np.concatenate(parts), full16))
        print(f"shard_nocoll_check world={world}: bitwise_equal={ok}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
