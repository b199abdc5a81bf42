"""Multi-GPU data-parallel test (needs >= 2 B200 on the box; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu


def test_two_rank_nccl_allreduce_matches_oracle_and_ranks_agree():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "scripts", "dp_check.py"), "1", "4099", "10"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DP_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_two_rank_peer_memory_collective_matches_oracle_and_ranks_agree():
    """Opt-in fused reduce + peer-memory all-reduce + Adam kernel (pinn_comm_p2p_export/open) instead of NCCL."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29612", os.path.join(ROOT, "scripts", "dp_check.py"), "1", "4099", "10"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, PINN_P2P="1"))
    assert r.returncode == 0 and "DP_CHECK OK" in r.stdout and "collective=p2p-fused" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_two_rank_mma_path_matches_oracle_and_ranks_agree():
    """The width-20 mma.sync kernel (PINN_PREC_FP16X3) under data parallelism: per-warp rows -> reduce -> NCCL -> Adam + mirrors."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29613", os.path.join(ROOT, "scripts", "dp_check.py"), "1", "4099", "10"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, PINN_DP_PRECISION="4"))
    assert r.returncode == 0 and "DP_CHECK OK" in r.stdout and "precision=4" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
