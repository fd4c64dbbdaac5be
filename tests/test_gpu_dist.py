"""y-slab pressure projection on two GPUs against the ORACLE (tests/dist_parity.py: config 3 in miniature at a stop rule both
meet, iteration delta of block-MIC(0), distance at the stock cap) and against the single-GPU projection (needs >= 2 devices;
the same oracle check also runs inside `bench.py --gpus N`, whose JSON line carries it as config.parity_checked)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_two_gpu_slab_projection_matches_oracle_and_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dist_check.py"), "192", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "rank 0 OK" in r.stdout and "rank 1 OK" in r.stdout
