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


def test_fused_coarse_level_sweep_matches_the_oracle(sf):
    """k_mg_coarse_fused (the default): the coarsest level's Chebyshev sweep as ONE cooperative launch.  The V-cycle must
    agree vector for vector with the oracle's, the solves with the direct solve -- and the cooperative launch must really
    have been used (a refused launch falls back to the launch-per-step sweep silently)."""
    from safeincave_b200 import _lib
    lib = _lib.load()
    n0 = lib.sic_mg_fused_coarse_launches()
    C.check_setup_vcycle_solve(sf, "cube_coarse", levels=2)
    C.check_setup_vcycle_solve(sf, "cavern_regular", levels=1, nonsym=0.01, full=False)
    assert lib.sic_mg_fused_coarse_launches() > n0, "the cooperative launch was refused: nothing was tested"


def test_launch_per_step_coarse_sweep_still_matches(sf):
    """fused_coarse=False: the 2 * coarse_its - 1 launches the fused kernel replaces (the path of the host emulation)."""
    C.check_setup_vcycle_solve(sf, "cube_coarse", levels=2, mg_kwargs=dict(fused_coarse=False))


def test_graph_replayed_iterations_equal_launched_ones(sf):
    """sic_ksp_t.use_graph: the Krylov iterations of a solve replayed from one captured CUDA graph give the same
    iteration counts and fields as the same kernels launched one by one, over two time steps with several set-ups
    (each re-captures: the Chebyshev coefficients are baked into the nodes)."""
    from safeincave_b200 import _lib, cases
    lib = _lib.load()
    out = {}
    for graph in (False, True):
        h, grid, case, eq, sim = C.make(sf, "cavern_regular", 1, cases.cavern_case, n_steps=2, ksp_type="cg", rtol=1e-10)
        eq.mg_options = dict(use_graph=graph)
        eq.solver.initial_guess_nonzero = True
        sim.verbose = False
        c0 = lib.sic_mg_graph_captures()
        hist = sim.run()
        out[graph] = (eq.X.clone(), [k[0] for k in eq.ksp_log], [r["iterations"] for r in hist], eq.mg.graph_launches,
                      lib.sic_mg_graph_captures() - c0, eq.mg.setups)
    (x0, its0, newton0, g0, cap0, _), (x1, its1, newton1, g1, cap1, setups) = out[False], out[True]
    assert g0 == 0 and cap0 == 0
    assert g1 > 0 and 2 <= cap1 <= 2 * (setups + 1), (g1, cap1, setups)      # two graphs per set-up: first cycle, iteration
    assert newton0 == newton1 and all(abs(a - b) <= 1 for a, b in zip(its0, its1)), (its0, its1)
    assert float((x0 - x1).abs().max() / x0.abs().max()) < 1e-9


def test_compressed_preconditioner_operator(sf):
    """The product's default: inside the V-cycle the operator reads float(sym(C_T)) and a float copy of the geometry
    (152 B per cell instead of 408).  Symmetric tangent: the cycle agrees with the exact-operator oracle to float
    precision and the solve takes the same iterations; 1 % non-symmetric tangent (10^4 x the reference's FD noise): the
    preconditioner is the symmetrised one, the solve still matches the direct solve at 1e-9 in about as many iterations."""
    C.check_setup_vcycle_solve(sf, "cube_coarse", levels=2, mg_kwargs=dict(compressed=True))
    # normal-shear couplings with the major symmetry of a creep tangent (W C_T symmetric, C_T itself not): the compressed
    # operator must be the SAME operator up to float rounding -- a plain sym(C_T) is wrong by O(coupling) here
    its_c = C.check_setup_vcycle_solve(sf, "cube_coarse", levels=2, coupled=0.05)
    its_cc = C.check_setup_vcycle_solve(sf, "cube_coarse", levels=2, coupled=0.05, mg_kwargs=dict(compressed=True))
    assert abs(its_cc - its_c) <= 1, (its_cc, its_c)
    its_exact = C.check_setup_vcycle_solve(sf, "cube_coarse", levels=2, nonsym=0.01, full=False)
    its = C.check_setup_vcycle_solve(sf, "cube_coarse", levels=2, nonsym=0.01, full=False, mg_kwargs=dict(compressed=True),
                                     cycle_tol=0.1)
    assert its <= its_exact + 3, (its, its_exact)


def test_bench_solver_settings_against_the_oracle_lu(sf):
    """what bench.py times (PC mg + extrapolated guess + lagged set-up) against the oracle's sparse LU on
    114 768 cells, two time steps, 1e-8 on u / sigma / eps, same Newton history."""
    C.check_bench_settings_against_lu_golden(sf)
