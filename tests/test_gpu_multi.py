"""-m gpu, needs >= 2 GPUs (skipped on a single-GPU box): time-range sharding over NCCL equals the
unsharded run bit for bit."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_time_sharded_nccl_equals_unsharded(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29540 + world),
                        str(ROOT / "tools" / "sharded_check.py"), "40", "4096"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "bitwise_equal=True" in r.stdout


def _host_test(*args):
    exe = ROOT / "melonix_b200" / "host_test"
    assert exe.exists(), "run __graft_entry__.build()"
    return subprocess.run([str(exe), *map(str, args)], capture_output=True, text=True, timeout=600)


@pytest.mark.parametrize("ngpu", [1, 2, 4])
def test_cpp_host_shards_one_file_without_python(tmp_path, ngpu):
    """A plain C++ host (melonix_b200/host/host_test.cpp `shard`): one thread per GPU, each with its own
    mlx_ctx and its rank of the NCCL communicator behind the C ABI (mlx_comm_create), hands in its owned
    samples through mlx_pv_run_sharded and gets its owned output -- equal, bit for bit, to the unsharded
    run.  ngpu = 1 exercises the same entry points on a single-GPU box."""
    import numpy as np
    sys.path.insert(0, str(ROOT / "tests"))
    import signals as S
    import melonix_b200 as m
    if _ngpu() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs")
    x = S.vibrato_tone(30.0, seed=77)
    x.tofile(tmp_path / "wav.f32")
    r = _host_test("shard", tmp_path / "wav.f32", 4096, "%.9g" % float(m.semitone_ratio(3.0)), ngpu,
                   tmp_path / "out.f32")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "failed_ranks=0" in r.stdout
    got = np.fromfile(tmp_path / "out.f32", np.float32)
    eng = m.Engine(0)
    try:
        eng.upload_tracks([x])
        full = eng.pv_run(4096, 1024, m.semitone_ratio(3.0))[0]
    finally:
        eng.close()
    assert np.array_equal(got, full["y"])


@pytest.mark.parametrize("world", [2, 4])
def test_collective_free_sharding_over_nccl_ranks(world):
    """Spec job-list sharding and grain output-range sharding (no data-path collective) on `world` GPUs."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29560 + world),
                        str(ROOT / "tools" / "shard_nocoll_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "bitwise_equal=True" in r.stdout


def test_collective_free_sharding_logical_shards_on_one_gpu():
    """The same two modes with the shards run one after the other on a single GPU (the driver's 1-GPU box):
    the concatenated blocks equal the unsharded results bit for bit."""
    import numpy as np
    sys.path.insert(0, str(ROOT / "tests"))
    import signals as S
    import melonix_b200 as m
    from melonix_b200 import dist as D
    from melonix_b200 import hostlib as H
    eng = m.Engine(0)
    try:
        x = S.vibrato_tone(3.0, seed=6)
        for N, hop in ((2048, 512), (32768, 375)):
            jobs = S.regular_jobs(x.size, hop)
            eng.upload_tracks([x])
            full = eng.spec_batch(0, N, jobs)
            for world in (3, 8):
                rows = [D.run_spec_sharded(eng, x, jobs, N, world, r)[1] for r in range(world)]
                assert np.array_equal(np.concatenate(rows), full), (N, hop, world)
        y = S.two_tone(6.0)
        markers = [(10, 0, 0, -2.0), (y.size - 10, 0, 0, 4.0)]
        gs, gl = H.grain_segment(y)
        sch = H.export_schedule(y, 48000, markers, gs, gl)
        eng.upload_tracks([y])
        pcm, pcm16 = H.export_wav(eng, 0, y, 48000, markers)
        for world in (2, 5):
            parts = [D.run_grain_sharded(eng, 0, sch, world, r) for r in range(world)]
            assert [p[0] for p in parts] == [int(sch["out_off"][D.shard_grain_rows(sch["out_off"], world, r)[0]])
                                             for r in range(world)]
            assert np.array_equal(np.concatenate([p[1] for p in parts]), pcm)
            assert np.array_equal(np.concatenate([p[2] for p in parts]), pcm16)
    finally:
        eng.close()
