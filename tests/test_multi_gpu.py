"""Multi-GPU slab parity (needs >= 2 GPUs on the box; skipped on the single-GPU test box).  Launches
tests/mgpu_check.py under torchrun: distributed operators / V-cycles / preconditioner / GMRES vs the
single-GPU plan on the same global vectors."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,dist_min_n,extra", [(256, 16, ["nokrylov"]), (1024, 0, [])])
def test_slab_parity_under_torchrun(n, dist_min_n, extra):
    """(256, 16): every level down to 16^2 distributed -- operators, V-cycles, preconditioner apply (bit-identical /
    1e-11 / 1e-10); (1024, 0): the default hierarchy, the same plus the Krylov histories under tests/conftest.hist_check
    (MGPU_ALL_PASS at 2, 4 and 8 ranks: profiles/r2_mgpu_parity_*gpu.log)."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if ngpu < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_check.py"),
           str(n), str(dist_min_n)] + extra
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert "MGPU_ALL_PASS" in res.stdout, res.stdout[-3000:] + res.stderr[-2000:]
