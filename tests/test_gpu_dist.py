"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): tests/dist_worker.py under torch.distributed.run."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_partitioned_spmv_and_cg(world):
    import hala_b200 as hb
    if hb.gpu_device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + world
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert r.stdout.count("dist ok") == 4, r.stdout[-2000:]
