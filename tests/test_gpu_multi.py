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
