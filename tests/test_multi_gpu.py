"""-m gpu, needs >= 2 GPUs: the one-process-per-GPU path (NCCL ray exchange over NVLink, framebuffer
reduce) against the oracle, launched the way the driver launches bench.py (torchrun, 127.0.0.1)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    from galaxy_b200 import gpu
    return gpu.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_one_process_per_gpu_matches_oracle(world):
    if _n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(ROOT, "tools", "mp_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sampler_one_process_per_gpu_matches_oracle(world):
    """gxy_sample with a communicator: every partition's set of samples and the summed ray statistics equal the oracle's"""
    if _n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + world), os.path.join(ROOT, "tools", "mp_sampler_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
