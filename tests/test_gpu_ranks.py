"""Two ranks on two B200s (torchrun, NCCL + the P2P exchange kernel over NVLink): the partitioned run against the
single-GPU run of the same case, executed by scripts/dist_check.py on every rank.  Skipped on a one-GPU box; the
host-side logic of the same paths runs on emulated ranks in tests/test_emu_ranks.py and on gloo in tests/test_partition.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(n, *args, timeout=900):
    import torch
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", "dist_check.py"), *args]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0 and "DIST CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_two_ranks_block_jacobi_cg():
    _torchrun(2, "--levels", "1", "--pc", "jacobi")


def test_two_ranks_multigrid_bench_settings_nested_partition():
    """PC mg + extrapolated guess + lagged setup (what bench.py times) on cavern_regular x8^2 (918k cells): levels 1 and 2
    partitioned by the cells' ancestors, level 0 replicated; fields equal to the single-GPU run of the same settings."""
    out = _torchrun(2, "--levels", "2", "--pc", "mg", "--min-cells-per-rank", "20000", "--rtol", "1e-10")
    assert "distributed from level 1" in out


def test_two_ranks_multigrid_fused_operator_exchange():
    """The same with the opt-in fused launch (k_mg_ebe_pc_x: interface tiles first, communication CTAs overlap the interior
    tiles): same fields as the single-GPU run, and the fused launch was really used."""
    out = _torchrun(2, "--levels", "2", "--pc", "mg", "--min-cells-per-rank", "20000", "--rtol", "1e-10", "--fused-exchange")
    # the V-cycle's operator + halo exchange ran as one launch (k_mg_ebe_pc_x), not as a silent fallback to two
    import re
    counts = [int(m) for m in re.findall(r"fused operator\+exchange launches (\d+)", out)]
    assert counts and min(counts) > 0, out[-2000:]


def test_two_ranks_thermomechanical_steps():
    """Simulator_TM with HeatDiffusion on the partitioned grid (BASELINE config 4's physics on cavern_regular x8)."""
    _torchrun(2, "--levels", "1", "--pc", "jacobi", "--heat", "--rtol", "1e-12")
