"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in melonix_b200/dist.py."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from melonix_b200 import dist as D  # noqa: E402


def test_shard_tracks_partition():
    for n in (1, 7, 64, 65):
        for w in (1, 2, 3, 8):
            got = [i for r in range(w) for i in D.shard_tracks(n, w, r)]
            assert got == list(range(n))
            sizes = [len(D.shard_tracks(n, w, r)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1


def test_time_shards_cover_and_halo():
    n, N, H = 48000 * 20 + 123, 4096, 1024
    for w in (1, 2, 4, 8):
        sh = D.plan_time_shards(n, N, H, w)
        assert sh[0].frame_begin == 0 and sh[-1].frame_end == D.num_frames(n, H)
        for a, b in zip(sh[:-1], sh[1:]):
            assert a.frame_end == b.frame_begin and a.own_hi == b.own_lo
        for s in sh:
            assert s.need_lo <= s.own_lo <= s.own_hi <= s.need_hi <= n
            assert s.need_lo == max(0, (s.frame_begin - 4) * H)        # halo frame fb-1 starts at (fb-4)H
            assert s.need_hi == min(n, (s.frame_end + 3) * H)          # three OLA frames after the range
            assert s.need_lo % H == 0 and s.frame_offset * H == s.need_lo
        if w > 1:
            assert sh[1].left_halo == N and sh[0].right_halo == 3 * H  # the seam payloads


def test_c_abi_shard_rule_matches_closed_form():
    """mlx_shard_frames (pure host function of libmelonix_b200.so; loads without a GPU) against the rule
    restated here: ragged lengths, more ranks than frames, single rank."""
    for n in (0, 1, 511, 512, 513, 48000 * 3 + 77, 345_600_000):
        for N in (512, 2048, 4096):
            H = N // 4
            F = (n + H - 1) // H
            for w in (1, 2, 3, 8):
                for r in range(w):
                    s = D.shard_frames(n, N, H, w, r)
                    fb, fe = F * r // w, F * (r + 1) // w
                    off = max(fb - 4, 0)
                    assert (s.frame_begin, s.frame_end, s.frame_offset) == (fb, fe, off)
                    assert (s.own_lo, s.own_hi) == (min(n, fb * H), min(n, fe * H))
                    assert (s.need_lo, s.need_hi) == (off * H, min(n, (fe + 3) * H))


def test_spec_job_shards_cover_and_halo(oracle):
    """shard_spec_jobs: blocks partition the job list; evaluating a rank's LOCAL jobs on its uploaded window
    (here with the CPU oracle standing in for the kernel) gives exactly the rows of the unsharded run --
    regular hops, the reference geometry (32768 / 375), ragged ends, more ranks than jobs."""
    sys.path.insert(0, str(ROOT / "tests"))
    import signals as S
    x = S.vibrato_tone(1.5, seed=3)
    for N, hop in ((1024, 256), (4096, 375)):
        jobs = S.regular_jobs(x.size, hop)
        full = oracle.spec_batch(x, N, jobs, nthreads=2)
        for w in (1, 2, 3, 8):
            rows = []
            for r in range(w):
                own, lo, hi, local = D.shard_spec_jobs(jobs, N, x.size, w, r)
                assert 0 <= lo <= hi <= x.size and local.shape[0] == len(own)
                if len(own):
                    assert lo == max(0, int(jobs[own.start:own.stop, 1].min()) - N)
                    rows.append(oracle.spec_batch(x[lo:hi], N, local, nthreads=2))
            assert np.array_equal(np.concatenate(rows), full), (N, hop, w)
    own, lo, hi, local = D.shard_spec_jobs(jobs[:2], 1024, x.size, 8, 5)
    assert len(own) == 0 and local.shape[0] == 0


def test_grain_row_shards_partition_the_output(oracle):
    """shard_grain_rows: row ranges are contiguous, cover every row once and balance the output samples;
    concatenating per-shard renders (CPU oracle resampler standing in for the kernel) equals the whole."""
    from melonix_b200 import hostlib as H
    sys.path.insert(0, str(ROOT / "tests"))
    import signals as S
    x = S.two_tone(6.0)
    markers = [(10, 0, 0, 3.0), (x.size - 10, 0, 0, 3.0)]
    gs, gl = H.grain_segment(x)
    sch = H.export_schedule(x, 48000, markers, gs, gl)
    rows = sch["gstart"].size
    whole = oracle.grain_export(x, 48000, markers)["pcm"]
    for w in (1, 2, 3, 8, rows + 5):
        cuts = [D.shard_grain_rows(sch["out_off"], w, r) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == rows
        assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:])) and all(a <= b for a, b in cuts)
        sizes = [int(sch["out_off"][b] - sch["out_off"][a]) for a, b in cuts]
        assert sum(sizes) == int(sch["out_off"][-1])
        if w <= 8:
            assert max(sizes) - min(sizes) <= 2 * int(np.diff(sch["out_off"]).max())
    assert whole.size == int(sch["out_off"][-1]) + sch["tail_zeros"]


def _worker(rank, world, port, n, N, H, out):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        x = torch.arange(n, dtype=torch.float32)
        shards = D.plan_time_shards(n, N, H, world)
        s = shards[rank]
        win = D.exchange_seam_samples(x[s.own_lo:s.own_hi].clone(), s, world)
        ok_seam = torch.equal(win, x[s.need_lo:s.need_hi])
        # exact uint32 phase carry: values near 2^32 must wrap
        tot = torch.tensor([[4294967295, 5, 2 ** 31 + rank]], dtype=torch.int64) + rank
        pre = D.exclusive_phase_prefix(tot, world, rank)
        exp = torch.zeros_like(tot)
        for r in range(rank):
            exp = (exp + torch.tensor([[4294967295, 5, 2 ** 31 + r]], dtype=torch.int64) + r) & 0xFFFFFFFF
        out[rank] = bool(ok_seam) and torch.equal(pre, exp)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_seam_exchange_and_phase_prefix_gloo(world):
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, 48000 * 3 + 77, 2048, 512, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)) and len(out) == world
