"""Parity checks of the multigrid path, shared by the host-emulation tests (tests/test_emu_mg.py, run
everywhere) and the B200 tests (tests/test_gpu_mg.py): the `sf` argument is the package with
``LinearMomentum.engine_cls`` pointing at the back end under test."""
import os

import numpy as np
import torch

from oracle import constitutive as oc
from oracle import fem
from oracle.mg import OracleMG
from tests.case_oracle import oracle_simulator

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def make(sf, name, levels, case_fn, **kw):
    from safeincave_b200 import cases
    from safeincave_b200.mesh import TetMesh
    from safeincave_b200.multigrid import refine_hierarchy
    h = refine_hierarchy(TetMesh.load_npz(os.path.join(GOLD, f"mesh_{name}.npz")), levels)
    grid = sf.GridHandlerGMSH.from_hierarchy(h)
    case = case_fn(grid, **kw)
    eq, sim = cases.build(case, grid)
    eq.solver.setType("cg")
    eq.solver.getPC().setType("mg")
    return h, grid, case, eq, sim


def random_tangent(n, seed, nonsym=0.0, coupled=0.0):
    """coupled: relative size of random normal-shear (and all other) couplings with the MAJOR SYMMETRY of a tangent of
    tensorial strains: W C_T symmetric, W = diag(1,1,1,2,2,2) -- so C_T[normal][shear] = 2 C_T[shear][normal], the structure
    of the creep tangents (doubled shear columns of G)."""
    rng = np.random.default_rng(seed)
    CT = oc.iso_matrix(102e9 * (1 + 0.5 * rng.random(n)), 0.3 * np.ones(n))
    if coupled:
        W = np.array([1.0, 1.0, 1.0, 2.0, 2.0, 2.0])
        S = W[None, :, None] * CT
        E = rng.standard_normal((n, 6, 6))
        S = S + coupled * S[:, :1, :1] * 0.5 * (E + E.transpose(0, 2, 1))
        CT = S / W[None, :, None]
    if nonsym:
        CT = CT * (1.0 + nonsym * rng.standard_normal((n, 6, 6)))
    return CT


def check_setup_vcycle_solve(sf, name="cube_coarse", levels=2, nonsym=0.0, full=True, mg_kwargs=None, cycle_tol=None,
                             coupled=0.0):
    """sic_mg_setup (Galerkin C_T, masks, blocks, lambda_max), one V-cycle and the full solve against the
    assembled-matrix oracle and the sparse direct solve."""
    from safeincave_b200 import cases
    from safeincave_b200.multigrid import Multigrid
    h, grid, case, eq, sim = make(sf, name, levels, cases.triaxial_case if name == "cube_coarse" else cases.cavern_case)
    eng = eq.engine
    N, M = eng.N, eng.M
    CT = random_tangent(N, 7, nonsym, coupled)
    eng.put_CT(CT)
    eps_rhs = 1e-5 * np.random.default_rng(3).standard_normal((N, 6))
    eng.put6(eng.eps_rhs, eps_rhs)
    eq.bc.update_dirichlet(0.0)
    eq.bc.update_neumann(0.0)
    # compressed=False: the preconditioner applies the exact FP64 operator, so the cycle can be compared vector for vector
    # at round-off level; with the float copy of sym(C_T) (the default of the product) the cycle agrees to ~1e-6
    mg_kwargs = {"compressed": False, **(mg_kwargs or {})}
    cycle_tol = cycle_tol if cycle_tol is not None else (1e-5 if mg_kwargs["compressed"] else 1e-10)
    eq.mg_options = dict(mg_kwargs)
    mg = Multigrid(eng, h, **mg_kwargs)
    mg.setup(eq.fixed, eq.dinv)
    fixed = eq.fixed.cpu().numpy().astype(bool)
    om = OracleMG(h.meshes, h.transfers, CT, fixed)
    # coarse tangents = mean of the children, masks injected, blocks inverted
    for l in range(h.n_levels):
        assert relerr(mg.engines[l].get_CT(), om.CT[l]) < 1e-14
        assert (mg.fixed[l].cpu().numpy().astype(bool) == om.fixed[l]).all()
        blocks = mg.dinv[l].cpu().numpy().reshape(-1, 3, 3)
        ref = np.stack([om.Dinv[l][3 * n:3 * n + 3, 3 * n:3 * n + 3].toarray() for n in range(0, h.meshes[l].n_nodes, 7)])
        assert relerr(blocks[::7], ref) < 1e-10
    # lambda_max: the power iteration approaches the largest eigenvalue (D^-1 K is not normal: not strictly from below); times `safety` it bounds it
    lam_dev = mg.lambda_max()
    for l in range(h.n_levels if full else 1):
        lam_ref = om.power_lambda(l, its=300)
        assert 0.88 * lam_ref < lam_dev[l] / mg.opts.safety <= 1.03 * lam_ref, (l, lam_dev[l], lam_ref)
        assert lam_dev[l] >= lam_ref, (l, lam_dev[l], lam_ref)
    if full:   # a second setup after the tangent changed restarts the power iteration from the kept vector (4 passes)
        CT2 = random_tangent(N, 11, nonsym, coupled) * (1.0 + 2.0 * (np.arange(N) % 3 == 0))[:, None, None]
        eng.put_CT(CT2)
        mg.setup(eq.fixed, eq.dinv)
        om2 = OracleMG(h.meshes, h.transfers, CT2, fixed)
        for l in range(h.n_levels):
            lam_ref = om2.power_lambda(l, its=300)
            assert lam_ref <= mg.lambda_max()[l] <= 1.2 * lam_ref, (l, mg.lambda_max()[l], lam_ref)
        eng.put_CT(CT)
        mg.setup(eq.fixed, eq.dinv)
        lam_dev = mg.lambda_max()
    # one V-cycle, vector for vector (same lambda_max on both sides)
    om.lam = lam_dev
    rng = np.random.default_rng(5)
    r = rng.standard_normal(3 * M) * (~fixed)
    z = mg.vcycle(torch.as_tensor(r).to(eng.device)).cpu().numpy()
    z_ref = om.vcycle(h.n_levels - 1, r)
    assert relerr(z, z_ref) < cycle_tol
    # a second cycle on the same data gives the same answer (no state leaks between cycles)
    z2 = mg.vcycle(torch.as_tensor(r).to(eng.device)).cpu().numpy()
    assert relerr(z2, z) < 1e-14
    # full solve against the sparse direct solve; iteration count as the oracle's PCG
    eq.solver.setTolerances(rtol=1e-12)
    res = eq._linear_solve()
    assert res.reason > 0, f"not converged: {res.reason} after {res.iterations}"
    osim = oracle_simulator(case, h.finest)
    u_ref = osim._solve(CT, eps_rhs, 0.0)
    assert relerr(eq.X.reshape(-1).cpu().numpy(), u_ref) < 1e-9
    if not full:          # large mesh: skip the oracle's own PCG (minutes of scipy on the host)
        return res.iterations
    dofs, vals, b_ext = osim._bc(0.0)
    b = (b_ext + fem.rhs_eps(h.finest.coords, h.finest.cells, CT, eps_rhs))
    u0 = np.zeros(3 * M)
    u0[dofs] = vals
    Kfull = fem.assemble_K(h.finest.coords, h.finest.cells, CT)
    rhs = (b - Kfull @ u0) * (~fixed)
    _, it_ref = om.pcg(rhs, rtol=1e-12)
    assert abs(res.iterations - it_ref) <= 2, (res.iterations, it_ref)
    return res.iterations


def check_time_steps(sf, name, levels, case_fn, n_steps, tol=1e-8, **kw):
    """Whole time steps with PC mg against the oracle's direct solves."""
    h, grid, case, eq, sim = make(sf, name, levels, case_fn, n_steps=n_steps, **kw)
    hist = sim.run()
    osim = oracle_simulator(case, h.finest)
    ohist = osim.run(0.0, [case["dt"]] * n_steps)
    assert [x["iterations"] for x in hist] == [x["iters"] for x in ohist[1:]]
    eng = eq.engine
    last = ohist[-1]
    assert relerr(eq.X.reshape(-1).cpu().numpy(), last["u"]) < tol
    assert relerr(eng.get6(eng.sig), last["sig"]) < tol
    for e_gpu, e_or in zip(eng.elems, osim.mat.elems):
        assert relerr(eng.get6(e_gpu.eps_old), e_or.eps_old) < tol
    return hist


def check_mg_equals_block_jacobi(sf, name, levels, case_fn, n_steps=1, tol=1e-8, warm_start=True, **kw):
    """Time steps with PC mg and with the block-Jacobi CG of solver.cu (itself checked against the oracle's direct
    solves in test_gpu_fem.py): same Newton iteration counts, same fields."""
    out = {}
    for pc in ("asm", "mg"):
        h, grid, case, eq, sim = make(sf, name, levels, case_fn, n_steps=n_steps, **kw)
        eq.solver.getPC().setType(pc)
        eq.solver.setInitialGuessNonzero(warm_start)
        hist = sim.run()
        out[pc] = (eq, hist)
    (ej, hj), (em, hm) = out["asm"], out["mg"]
    assert [x["iterations"] for x in hm] == [x["iterations"] for x in hj]
    assert all(k[1] > 0 for k in em.ksp_log)
    assert relerr(em.X.reshape(-1).cpu().numpy(), ej.X.reshape(-1).cpu().numpy()) < tol
    assert relerr(em.engine.get6(em.engine.sig), ej.engine.get6(ej.engine.sig)) < tol
    return max(k[0] for k in em.ksp_log), max(k[0] for k in ej.ksp_log)


def recurrence_sequence(n_nodes, lam=(0.7, 0.25), n_iter=4, seed=5, device="cpu"):
    """Iterates u_k = u* + lam1^k v1 + lam2^k v2 (a two-mode linear recurrence): the next one is predicted exactly."""
    import torch
    g = torch.Generator().manual_seed(seed)
    ustar, v1, v2 = (torch.randn(n_nodes, 3, generator=g, dtype=torch.float64) for _ in range(3))
    seq = [ustar + lam[0] ** k * v1 + lam[1] ** k * v2 for k in range(n_iter + 1)]
    return [u.to(device) for u in seq]       # seq[-1] is the iterate to predict from seq[:-1]


def check_guess_extrapolation(engine, seq, local_nodes=None):
    """sic_guess_extrapolate on the last four iterates of `seq[:-1]` must return seq[-1] (two-term model), fall back to the
    one-term model on a one-mode sequence and to the plain warm start on a sequence that is no recurrence at all."""
    import torch
    pick = (lambda u: u[local_nodes].contiguous().reshape(-1)) if local_nodes is not None else (lambda u: u.contiguous().reshape(-1))
    its = [pick(u) for u in seq[:-1]][::-1][:4]           # newest first
    x = torch.empty_like(its[0])
    a, b, used, f1, f2 = engine.guess_extrapolate(its, x, want_coef=True)
    assert used == 2 and f2 < 1e-20, (a, b, used, f1, f2)
    assert abs(a - 0.95) < 1e-9 and abs(b + 0.175) < 1e-9, (a, b)      # lam1 + lam2, -lam1 lam2
    assert relerr(x.cpu().numpy(), pick(seq[-1]).cpu().numpy()) < 1e-12
    # no validation possible with three iterates: plain warm start
    a3 = engine.guess_extrapolate(its[:3], x, want_coef=True)
    assert a3[2] == 0 and torch.equal(x, its[0])
    # a sequence that stops dead (the uniform triaxial cube): d1 is not explained by the earlier increments
    dead = [its[1] + 1e-3 * torch.roll(its[3], 7), its[1], its[2], its[3]]
    ad = engine.guess_extrapolate(dead, x, want_coef=True)
    assert ad[2] == 0 and torch.equal(x, dead[0]), ad
    return a, b


def check_lagged_setup(sf, name="cube_coarse", levels=2, n_steps=3, gravity=5e3):
    """KSP.mg_setup_first = 2: the multigrid preconditioner is rebuilt on the first two Newton iterations of a step and
    then kept while the Newton error is below 1e-3.  Same Newton history and fields as rebuilding for every tangent,
    (nearly) the same Krylov iterations, far fewer setups.  A body force makes the cube's stress non-uniform, so that a
    step takes 5-6 Newton iterations."""
    from safeincave_b200 import cases
    out = {}
    for lag in (0, 2):
        h, grid, case, eq, sim = make(sf, name, levels, cases.triaxial_case, n_steps=n_steps, ksp_override="cg")
        case["g"] = [0.0, 0.0, -gravity]
        eq, sim = cases.build(case, grid)
        eq.solver.getPC().setType("mg")
        eq.solver.setGuessExtrapolation(True)
        eq.solver.mg_setup_first = lag
        sim.verbose = False
        hist = sim.run()
        out[lag] = (eq, hist)
    (e0, h0), (e2, h2) = out[0], out[2]
    assert [x["iterations"] for x in h2] == [x["iterations"] for x in h0] and max(x["iterations"] for x in h0) >= 5
    assert relerr(e2.X.reshape(-1).cpu().numpy(), e0.X.reshape(-1).cpu().numpy()) < 1e-9
    assert relerr(e2.engine.get6(e2.engine.sig), e0.engine.get6(e0.engine.sig)) < 1e-9
    k0, k2 = sum(k[0] for k in e0.ksp_log), sum(k[0] for k in e2.ksp_log)
    assert k2 <= 1.05 * k0 + 2, (k0, k2)
    assert e0.mg.setups == len(e0.ksp_log) and e2.mg.setups <= 1 + 3 * n_steps, (e0.mg.setups, e2.mg.setups)
    return e0.mg.setups, e2.mg.setups, k0, k2


def check_bench_settings_against_lu_golden(sf, tol=1e-8):
    """The solver configuration bench.py times -- PC mg (compressed preconditioner operator, fused coarsest sweep, graph
    replay), Krylov guess extrapolated from the Newton iterates, multigrid set-up lagged after the second Newton
    iteration -- on cavern_regular x8 (114 768 cells), two time steps of BASELINE config 2, against the ORACLE's sparse-LU
    run of the same steps (tests/golden/cfg2_cavern_regular_L1.npz, oracle/gen_staged_golden.py cfg2)."""
    from safeincave_b200 import cases
    g = np.load(os.path.join(GOLD, "cfg2_cavern_regular_L1.npz"))
    h, grid, case, eq, sim = make(sf, "cavern_regular", int(g["levels"]), cases.cavern_case, n_steps=int(g["n_steps"]),
                                  ksp_type="cg", rtol=1e-12)
    assert h.finest.n_cells == int(g["n_cells"]) and abs(float(np.abs(h.finest.coords).sum()) / float(g["coords_checksum"]) - 1) < 1e-12
    eq.solver.initial_guess_nonzero = True
    eq.solver.guess_extrapolation = True
    eq.solver.mg_setup_first = 2
    sim.verbose = False
    hist = sim.run()
    assert all(r["converged"] and r["dt_used"] == r["dt"] for r in hist)
    assert all(k[1] > 0 for k in eq.ksp_log), "a Krylov solve did not reach its tolerance"
    assert [r["iterations"] for r in hist] == list(g["iters"])
    eng = eq.engine
    assert relerr(eq.X.reshape(-1).cpu().numpy(), g["u"]) < tol
    sel = g["cell_sel"]
    sig, eps = eng.get6(eng.sig), eng.get6(eng.eps)
    assert np.abs(sig[sel] - g["sig_sel"]).max() / float(g["sig_absmax"]) < tol
    assert np.abs(eps[sel] - g["eps_sel"]).max() / float(g["eps_absmax"]) < tol
    assert abs(np.linalg.norm(sig) / float(g["sig_norm"]) - 1) < tol
    assert eq.mg.compressed and eq.mg.setups < len(eq.ksp_log)       # the lag really skipped set-ups
    return hist
