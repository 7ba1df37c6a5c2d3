"""The parity tests of tests/test_gpu_fem.py (operator, RHS, preconditioner blocks, Krylov solves, whole
time steps against oracle/fem.py), executed on the HOST through tests/hostemu: the product's own kernels
and drivers (fem.cu, solver.cu, constitutive.cu), compiled by g++ against a CUDA emulation layer.
What this covers that `-m "not gpu"` otherwise could not: indexing, scatter plans, reductions, the
device-resident Krylov recurrences and the Simulator_M loop.  What it cannot cover: nvcc code generation
and real concurrency (atomics, memory ordering) -- that is what `-m gpu` is for."""
import pytest

from tests import test_gpu_fem as G


@pytest.fixture(scope="module")
def sf():
    import safeincave_b200 as sf
    from tests.hostemu import EmuEngine
    old = sf.LinearMomentum.engine_cls
    sf.LinearMomentum.engine_cls = EmuEngine
    yield sf
    sf.LinearMomentum.engine_cls = old


test_operator_rhs_blocks_strain = G.test_operator_rhs_blocks_strain
test_operator_rhs_blocks_strain_configs2_grid = G.test_operator_rhs_blocks_strain_configs2_grid
test_neumann_and_body_force = G.test_neumann_and_body_force
test_time_steps_triaxial_cube = G.test_time_steps_triaxial_cube
test_time_steps_triaxial_cube_with_desai = G.test_time_steps_triaxial_cube_with_desai
test_krylov_solve_matches_direct = G.test_krylov_solve_matches_direct
test_time_steps_triaxial_cube_munson_dawson = G.test_time_steps_triaxial_cube_munson_dawson
test_time_steps_triaxial_cube_mohr_coulomb_and_matsuoka_nakai = G.test_time_steps_triaxial_cube_mohr_coulomb_and_matsuoka_nakai
test_pq_output_fields = G.test_pq_output_fields
test_dt_retry_and_restore_follow_the_reference = G.test_dt_retry_and_restore_follow_the_reference
test_guess_extrapolation_kernel = G.test_guess_extrapolation_kernel


def test_time_step_cavern_regular_extrapolated_guess(sf):
    """One step (17 Newton iterations) of the B200 test test_time_steps_cavern_regular_extrapolated_guess.  The plain
    warm start needs 2782 Krylov iterations for this step (measured with this emulation; not re-run here to keep the
    CPU suite short), the extrapolated guess 1992."""
    G.check_extrapolated_guess(sf, 1, plain_its=2782)


test_staged_triaxial_cube_with_desai = G.test_staged_triaxial_cube_with_desai
test_nan_step_is_restored_under_a_warm_started_solver = G.test_nan_step_is_restored_under_a_warm_started_solver
