"""N > 1 on real GPUs: launches tests/multi_gpu_check.py under torchrun (one rank per GPU).  Skipped on a one-GPU box; the
host-side partition / exchange logic is covered on CPU by tests/test_partition_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    from tahoe_b200 import capi
    try:
        return capi.device_count()
    except capi.Tb2Error:
        return 0


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_partitioned_explicit_matches_single_gpu(world, exchange):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(29530 + world + (10 if exchange == "nccl" else 0)),
                        os.path.join(HERE, "multi_gpu_check.py"), exchange],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-3000:]
