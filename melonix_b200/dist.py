"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on B200s, gloo in the CPU tests).

Two sharding modes:

* independent tracks (BASELINE configs[2]): `shard_tracks` -- contiguous blocks of tracks per rank,
  no data-path collective at all.
* one long file sharded by contiguous time range (configs[3]): rank r owns frames
  [F*r/G, F*(r+1)/G).  The product path is `run_time_sharded`, a thin caller of the C ABI's
  mlx_pv_run_sharded_dev (NCCL inside the library, csrc/multi.cu); `run_time_sharded_torch` runs the same
  protocol with torch.distributed collectives.  Two small exchanges per track:
    1. seam samples: the analysis of a rank's first frame needs the `fftN` samples before its range
       (and one hop more for the halo frame), its last three overlap-add hops need `3*hop` samples
       after it -> one batched NCCL send/recv with each neighbour (`exchange_seam_samples`);
    2. phase carry: the exact uint32 per-bin phase totals of every rank's owned frames
       (fftN/2+1 values) are all-gathered and prefix-summed (`exclusive_phase_prefix`); integer
       addition is associative, so the sharded result is bit-identical to the unsharded one.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


def shard_tracks(ntracks: int, world: int, rank: int) -> range:
    """Contiguous block of track indices owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(ntracks, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def num_frames(n: int, hop: int) -> int:
    return (n + hop - 1) // hop


@dataclass(frozen=True)
class TimeShard:
    """Frame range owned by one rank and the sample window it must hold (global indices)."""
    rank: int
    frame_begin: int       # first owned frame (global)
    frame_end: int         # one past the last owned frame (global)
    own_lo: int            # owned samples [own_lo, own_hi) = output hops of the owned frames
    own_hi: int
    need_lo: int           # samples the rank must hold: [need_lo, need_hi), includes both halos
    need_hi: int
    frame_offset: int      # local frame index = global frame index - frame_offset

    @property
    def local_frame_begin(self) -> int:
        return self.frame_begin - self.frame_offset

    @property
    def local_frame_end(self) -> int:
        return self.frame_end - self.frame_offset

    @property
    def left_halo(self) -> int:
        return self.own_lo - self.need_lo

    @property
    def right_halo(self) -> int:
        return self.need_hi - self.own_hi


def shard_frames(n: int, fft_n: int, hop: int, world: int, rank: int) -> TimeShard:
    """Contiguous time-range shard of a track of n samples for `rank` of `world`.  The rule lives in the
    C ABI (mlx_shard_frames, csrc/multi.cu -- a pure host function, no GPU needed) so that the C++ and the
    Python drivers cannot disagree: frame f covers samples [(f-3)*hop, (f+1)*hop) (fft_n = 4*hop); needed
    frames are fb-1 (halo whose spectrum seeds the phase difference) .. fe+2 (their tails overlap-add
    into the owned hops)."""
    import ctypes as C

    from . import capi
    sh = capi.TimeShardC()
    capi.check(capi.lib().mlx_shard_frames(int(n), int(fft_n), int(hop), int(world), int(rank), C.byref(sh)))
    return TimeShard(rank, sh.frame_begin, sh.frame_end, sh.own_lo, sh.own_hi, sh.need_lo, sh.need_hi,
                     sh.frame_offset)


def exchange_seam_samples(own: torch.Tensor, shard: TimeShard, world: int, group=None) -> torch.Tensor:
    """Builds the rank's sample window [need_lo, need_hi) from its owned samples plus the seam samples
    of its two neighbours: ONE batched send/recv (left neighbour's tail, right neighbour's head).
    `own` holds samples [own_lo, own_hi) on the compute device.  Every rank must call this."""
    rank = shard.rank
    left, right = shard.left_halo, shard.right_halo
    buf = torch.zeros(left + own.numel() + right, dtype=own.dtype, device=own.device)
    buf[left:left + own.numel()] = own
    ops = []
    if rank > 0 and left > 0:
        ops.append(dist.P2POp(dist.irecv, buf[:left], rank - 1, group))
    if rank + 1 < world and right > 0:
        ops.append(dist.P2POp(dist.irecv, buf[left + own.numel():], rank + 1, group))
    # what the neighbours need from this rank is determined by THEIR shard; all ranks use the same rule
    send_bufs = []
    if rank + 1 < world:
        nxt = _neighbour_need(shard, +1)
        if nxt > 0:
            send_bufs.append(own[own.numel() - nxt:].contiguous())
            ops.append(dist.P2POp(dist.isend, send_bufs[-1], rank + 1, group))
    if rank > 0:
        prv = _neighbour_need(shard, -1)
        if prv > 0:
            send_bufs.append(own[:prv].contiguous())
            ops.append(dist.P2POp(dist.isend, send_bufs[-1], rank - 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return buf


def _neighbour_need(shard: TimeShard, direction: int) -> int:
    """Samples of this rank's owned range the neighbour needs (carried in shard by the caller via
    `with_neighbours`).  Stored as attributes to keep TimeShard a plain value type."""
    return getattr(shard, "_to_right" if direction > 0 else "_to_left")


def plan_time_shards(n: int, fft_n: int, hop: int, world: int) -> list[TimeShard]:
    """All ranks' shards, each annotated with how many of its samples its neighbours need."""
    shards = [shard_frames(n, fft_n, hop, world, r) for r in range(world)]
    for r, s in enumerate(shards):
        to_right = shards[r + 1].left_halo if r + 1 < world else 0
        to_left = shards[r - 1].right_halo if r > 0 else 0
        if to_right > s.own_hi - s.own_lo or to_left > s.own_hi - s.own_lo:
            raise ValueError("time shards are shorter than the seam overlap: use fewer ranks for this file")
        object.__setattr__(s, "_to_right", to_right)
        object.__setattr__(s, "_to_left", to_left)
    return shards


def exclusive_phase_prefix(totals: torch.Tensor, world: int, rank: int, group=None) -> torch.Tensor:
    """totals: [ntracks, nbins] int64 holding uint32 values (this rank's phase totals).  Returns the
    sum over lower ranks modulo 2^32 (the phase carried into this rank's first frame)."""
    gathered = [torch.empty_like(totals) for _ in range(world)]
    dist.all_gather(gathered, totals.contiguous(), group=group)
    acc = torch.zeros_like(totals)
    for r in range(rank):
        acc = (acc + gathered[r]) & 0xFFFFFFFF
    return acc


def make_comm(engine, group=None):
    """One NCCL communicator behind the C ABI per rank (mlx_comm): rank 0 draws the unique id, the bytes
    travel over the existing torch.distributed group (any backend), every rank joins."""
    from .engine import Comm
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    box = [Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return Comm(engine, box[0], world, rank)


def run_time_sharded(engine, owns, n_total: int, fft_n: int, hop: int, rate: float,
                     sample_rate: float = 48000.0, group=None, comm=None, outs=None):
    """Phase-vocodes long mono track(s) of n_total samples sharded by time range across the ranks
    (BASELINE configs[3]: "stereo" = two planar mono tracks).  Thin caller of mlx_pv_run_sharded_dev:
    seam send/recv, the single analysis pass, the phase-total all-gather and the synthesis all happen
    behind the C ABI (csrc/multi.cu).

    owns: this rank's owned samples, one CUDA float32 tensor per track (global samples
    [own_lo, own_hi) of plan_time_shards(...)[rank]); a single tensor is accepted too.
    comm: a Comm from make_comm (created and cached on the engine when omitted).
    outs: optional preallocated (ys, peaks, f0s) lists to write into.
    Returns per track (y_own, peak_own, f0_own) on the device.  Bit-identical to the unsharded run."""
    single = isinstance(owns, torch.Tensor)
    owns = [owns] if single else list(owns)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if comm is None:
        comm = getattr(engine, "_comm", None)
        if comm is None or comm.world != world or comm.rank != rank:
            comm = engine._comm = make_comm(engine, group)
    shard = shard_frames(n_total, fft_n, hop, world, rank)
    dev = owns[0].device
    nf = shard.frame_end - shard.frame_begin
    if outs is None:
        ys = [torch.empty_like(o) for o in owns]
        peaks = [torch.empty(nf, dtype=torch.int32, device=dev) for _ in owns]
        f0s = [torch.empty(nf, dtype=torch.float32, device=dev) for _ in owns]
    else:
        ys, peaks, f0s = outs
    engine.use_torch_stream()
    engine.pv_run_sharded_dev(comm, owns, n_total, fft_n, hop, rate, ys, peaks, f0s, sample_rate=sample_rate)
    out = list(zip(ys, peaks, f0s))
    return out[0] if single else out


def run_time_sharded_torch(engine, owns, n_total: int, fft_n: int, hop: int, rate: float,
                           sample_rate: float = 48000.0, group=None):
    """The same protocol with torch.distributed collectives around the split C-ABI calls
    (mlx_pv_analyze_dev -> all_gather -> mlx_pv_synth_dev): for process groups whose transport is not
    the library's own NCCL communicator.  One analysis pass as well; bit-identical results."""
    single = isinstance(owns, torch.Tensor)
    owns = [owns] if single else list(owns)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    shard = plan_time_shards(n_total, fft_n, hop, world)[rank]
    windows = [exchange_seam_samples(o, shard, world, group) for o in owns]     # seam exchange (1)
    dev = owns[0].device
    engine.use_torch_stream()
    engine.upload_tracks_dev(windows)
    nb = fft_n // 2 + 1
    nt = len(owns)
    lb, le = shard.local_frame_begin, shard.local_frame_end
    F_local = num_frames(windows[0].numel(), hop)
    totals32 = torch.zeros((nt, nb), dtype=torch.int32, device=dev)
    peak = torch.zeros((nt, F_local), dtype=torch.int32, device=dev)
    f0 = torch.zeros((nt, F_local), dtype=torch.float32, device=dev)
    engine.pv_analyze_dev(fft_n, hop, rate, [totals32[t] for t in range(nt)], [peak[t] for t in range(nt)],
                          [f0[t] for t in range(nt)], sample_rate=sample_rate, frame_begin=lb, frame_end=le)
    totals = totals32.to(torch.int64) & 0xFFFFFFFF
    carry = exclusive_phase_prefix(totals, world, rank, group)                  # phase carry (2)
    carry32 = torch.where(carry >= 2 ** 31, carry - 2 ** 32, carry).to(torch.int32).contiguous()
    ys = [torch.zeros_like(w) for w in windows]
    engine.pv_synth_dev(fft_n, hop, rate, ys, sample_rate=sample_rate, frame_begin=lb, frame_end=le,
                        phase_in=[carry32[t] for t in range(nt)])
    lo = shard.left_halo
    out = [(ys[t][lo:lo + owns[t].numel()], peak[t, lb:le], f0[t, lb:le]) for t in range(nt)]
    return out[0] if single else out


# ------------------------------------------------------------------------------------------------
# The two sharding modes that need no collective at all

def shard_spec_jobs(jobs, fft_n: int, n: int, world: int, rank: int):
    """Spec STFT jobs (reference Spec::getSpec ranges (start, end), spec.cpp:18,44-66) are independent: the
    job list is cut into contiguous blocks and each rank needs only the samples its windows
    [end - fft_n, end) touch -- a READ-ONLY halo that is uploaded with the shard (the host has the whole
    file; no exchange).  Returns (slice of `jobs` owned by `rank`, need_lo, need_hi, local_jobs) where
    local_jobs are the owned jobs in the coordinates of the uploaded window wav[need_lo:need_hi]."""
    import numpy as np
    jobs = np.asarray(jobs, np.int64).reshape(-1, 2)
    own = shard_tracks(jobs.shape[0], world, rank)          # same balanced contiguous split as for tracks
    mine = jobs[own.start:own.stop]
    if mine.shape[0] == 0:
        return own, 0, 0, mine.astype(np.int32)
    need_lo = int(max(0, mine[:, 1].min() - fft_n))          # indices < 0 read as zero (spec.cpp:50-54)
    need_hi = int(min(n, max(mine[:, 1].max(), need_lo)))
    return own, need_lo, need_hi, (mine - need_lo).astype(np.int32)


def run_spec_sharded(engine, wav, jobs, fft_n: int, world: int, rank: int):
    """This rank's block of Spec columns: uploads wav[need_lo:need_hi] only, returns (job slice,
    [count, fft_n/2] float32).  Rows of different ranks concatenate to the unsharded result, bit for bit."""
    import numpy as np
    own, lo, hi, local = shard_spec_jobs(jobs, fft_n, len(wav), world, rank)
    if local.shape[0] == 0:
        return own, np.zeros((0, fft_n // 2), np.float32)
    engine.upload_tracks([np.ascontiguousarray(wav[lo:hi], np.float32)])
    return own, engine.spec_batch(0, fft_n, local)


def shard_grain_rows(out_off, world: int, rank: int):
    """Grain path (reference App::process / exportWav, app.cpp:294-345, 1194-1215): the cursor recurrence
    that produces the render schedule is serial but O(#grains) and stays on the host; given the schedule,
    every output sample depends only on its own row, so the rows are cut where the OUTPUT sample count
    balances.  Returns the half-open row range [r0, r1) of `rank` (replicas of the source track, disjoint
    outputs, no collective)."""
    import numpy as np
    off = np.asarray(out_off, np.int64)
    rows = off.size - 1
    total = int(off[-1]) if rows >= 0 else 0
    cut = [int(np.searchsorted(off, total * r // world, side="left")) for r in range(world + 1)]
    cut[0], cut[-1] = 0, rows
    cut = [min(max(c, 0), rows) for c in cut]
    return cut[rank], max(cut[rank], cut[rank + 1])


def run_grain_sharded(engine, track: int, sched: dict, world: int, rank: int):
    """Renders this rank's rows of an export schedule (melonix_b200.hostlib.export_schedule) with
    mlx_grain_render: (first output sample, float32 pcm, int16 pcm).  The reference's trailing zeros
    (app.cpp:303-309) belong to the last rank."""
    r0, r1 = shard_grain_rows(sched["out_off"], world, rank)
    off = sched["out_off"][r0:r1 + 1] - sched["out_off"][r0]
    tail = sched["tail_zeros"] if rank == world - 1 else 0
    pcm, pcm16 = engine.grain_render(track, sched["gstart"][r0:r1], sched["glen"][r0:r1], sched["rate"][r0:r1], off,
                                     sched["next"][r0:r1], tail_zeros=tail)
    return int(sched["out_off"][r0]), pcm, pcm16


def bind_to_gpu_numa_node(device: int) -> dict:
    """Pins the calling process to the CPUs of the NUMA node the GPU hangs off, so that pinned host
    buffers allocated afterwards (first touch) and the copy-issuing threads are local to the GPU's
    PCIe root.  With one process per GPU this keeps the H2D/D2H streams of different ranks off the
    inter-socket link.  Returns what it did; never raises (no sysfs / no NVML -> no-op)."""
    import os
    info = dict(device=device, node=None, cpus=0)
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(node=node, cpus=len(cpus))
    except Exception as e:  # noqa: BLE001
        info["error"] = str(e)
    return info
