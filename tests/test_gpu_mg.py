"""B200 parity tests of the geometric-multigrid preconditioned CG (csrc/mg.cu) through the C ABI, against
oracle/mg.py (assembled matrices, vector-for-vector V-cycle) and the sparse direct solve of oracle/fem.py.
The same checks run on the host through tests/hostemu in tests/test_emu_mg.py."""
import pytest

from tests import mg_checks as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sf():
    import safeincave_b200 as sf
    return sf


def test_setup_vcycle_solve_cube(sf):
    assert C.check_setup_vcycle_solve(sf, "cube_coarse", levels=2) <= 30


def test_setup_vcycle_solve_cube_nonsymmetric_tangent(sf):
    C.check_setup_vcycle_solve(sf, "cube_coarse", levels=2, nonsym=0.01)


def test_setup_vcycle_solve_cavern_regular(sf):
    """115k cells, graded cavern mesh: ~28 MG-CG iterations where block-Jacobi CG needs ~660."""
    assert C.check_setup_vcycle_solve(sf, "cavern_regular", levels=1, nonsym=0.01, full=False) <= 40


def test_time_steps_triaxial_cube_mg(sf):
    from safeincave_b200 import cases
    C.check_time_steps(sf, "cube_coarse", 2, cases.triaxial_case, 3, ksp_override="cg")


def test_time_step_cavern_regular_mg_equals_block_jacobi(sf):
    """The bench configuration (cavern physics, CG warm-started from the previous Newton iterate) on 115k cells:
    PC mg and the block-Jacobi CG give the same Newton history and fields."""
    from safeincave_b200 import cases
    its_mg, its_bj = C.check_mg_equals_block_jacobi(sf, "cavern_regular", 1, cases.cavern_case)
    assert its_mg <= 40 and its_bj > 10 * its_mg


def test_lagged_multigrid_setup(sf):
    C.check_lagged_setup(sf)


@pytest.mark.xfail(reason="opt-in path (SIC_MG_FUSED_COARSE=1) whose first run on a GPU is this test", strict=False)
def test_fused_coarse_level_sweep_matches_the_oracle():
    """k_mg_coarse_fused: the coarsest level's Chebyshev sweep as ONE cooperative launch.  Run in a child process (the
    switch is read once per process; a hang would be cut off by the timeout without taking the suite along): the
    V-cycle must still agree vector for vector with the oracle's, and the solves with the direct solve."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import safeincave_b200 as sf; from tests import mg_checks as C; "
            "its = C.check_setup_vcycle_solve(sf, 'cube_coarse', levels=2); "
            "its2 = C.check_setup_vcycle_solve(sf, 'cavern_regular', levels=1, nonsym=0.01, full=False); "
            "from safeincave_b200 import _lib; n = _lib.load().sic_mg_fused_coarse_launches(); "
            "assert n > 0, 'the cooperative launch was refused: nothing was tested'; print('FUSED_OK', its, its2, n)")
    env = dict(os.environ, SIC_MG_FUSED_COARSE="1")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "FUSED_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
