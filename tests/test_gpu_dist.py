"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): tests/dist_worker.py under torch.distributed.run."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


TRANSPORTS = {
    "peer": {},                                     # NVLink peer memory, halo wait + partial publish fused into the SpMV kernel
    "peer-helpers": {"HB_PEER_FUSED_SPMV": "0"},    # peer memory, halo wait / publish as 1-warp kernels around the SpMV
    "nccl": {"HB_DIST_PEER": "0"},                  # ncclSend/Recv + ncclAllReduce
}


@pytest.mark.gpu
@pytest.mark.parametrize("transport", list(TRANSPORTS))
@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_partitioned_spmv_and_cg(world, transport):
    import hala_b200 as hb
    if hb.gpu_device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + world + 16 * list(TRANSPORTS).index(transport)
    env = dict(os.environ, **TRANSPORTS[transport], HB_EXPECT_TRANSPORT=transport.split("-")[0])
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py")], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert r.stdout.count("dist ok") == 4, r.stdout[-2000:]
