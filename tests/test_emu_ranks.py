"""The multi-GPU code paths of the library (partition, halo plans, owner-weighted dot products, the `multi` branches
of solver.cu) executed by several emulated RANKS -- one host thread each -- that meet in the emulated sic_exchange
(tests/hostemu/ranks.py).  The partitioned run must reproduce the single-rank run: fields to 1e-8 and the same
Newton history, as scripts/dist_check.py checks on real GPUs."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def sf():
    import safeincave_b200 as sf
    from tests.hostemu import EmuEngine
    old = sf.LinearMomentum.engine_cls
    sf.LinearMomentum.engine_cls = EmuEngine
    yield sf
    sf.LinearMomentum.engine_cls = old


def partitioned_vs_single(sf, world, ksp="cg", levels=1, n_steps=2, setup=None, multigrid=False, nested=False,
                          min_cells_per_rank=200_000):
    from safeincave_b200 import cases, distributed
    from safeincave_b200.mesh import TetMesh
    from safeincave_b200.multigrid import refine_hierarchy
    from tests.hostemu.ranks import run_ranks
    h = refine_hierarchy(TetMesh.load_npz(os.path.join(GOLD, "mesh_cube_coarse.npz")), levels, nested=nested)
    tm = h.finest
    gg = sf.GridHandlerGMSH.from_hierarchy(h)
    case = cases.triaxial_case(gg, n_steps=n_steps, ksp_override=ksp)
    if multigrid:
        setup = lambda eq, grid, part: eq.solver.getPC().setType("mg")
    eq1, sim1 = cases.build(case, gg)
    if setup:
        setup(eq1, gg, None)
    sim1.verbose = False
    hist1 = sim1.run()

    def body(ctx):
        grid, part = distributed.partition_grid(ctx, tm, hierarchy=h if multigrid else None,
                                                min_cells_per_rank=min_cells_per_rank)
        eq, sim = cases.build(case, grid, part=part, ctx=ctx)
        if setup:
            setup(eq, grid, part)
        sim.verbose = False
        hist = sim.run()
        ln = part.local_nodes
        c0, c1 = part.cell_range
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
        return dict(e_u=rel(eq.X, eq1.X[ln]), e_s=rel(eq.engine.sig[:, :eq.engine.N], eq1.engine.sig[:, c0:c1]),
                    e_c=rel(eq.engine.elems[1].eps_old[:, :eq.engine.N], eq1.engine.elems[1].eps_old[:, c0:c1]),
                    newton=[h["iterations"] for h in hist], ksp=[h["ksp_iterations"] for h in hist],
                    peers=part.peers, n=eq.engine.N, lc=eq.mg.lc if eq.mg is not None else None,
                    level_cells=[e.N for e in eq.mg.engines] if eq.mg is not None else None)

    res = run_ranks(world, body)
    for r in res:
        r["ksp1"] = [x["ksp_iterations"] for x in hist1]
    assert sum(r["n"] for r in res) == tm.n_cells
    for r in res:
        assert r["newton"] == [h["iterations"] for h in hist1]
        assert r["e_u"] < 1e-8 and r["e_s"] < 1e-8 and r["e_c"] < 1e-8, r
    return res, hist1


@pytest.mark.parametrize("world,ksp", [(2, "cg"), (3, "cg"), (2, "bicg")])
def test_partitioned_block_jacobi_krylov_matches_single_rank(sf, world, ksp):
    res, hist1 = partitioned_vs_single(sf, world, ksp)
    assert all(len(r["peers"]) >= 1 for r in res)


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_multigrid_matches_single_rank(sf, world):
    """PC mg with the finest level distributed over the ranks and the coarser levels replicated (csrc/mg.cu): same
    fields, same Newton history and the SAME Krylov iteration counts as the single-rank multigrid run (the algorithm is
    identical; only the order of floating-point sums differs)."""
    res, hist1 = partitioned_vs_single(sf, world, "cg", levels=2, n_steps=2, multigrid=True)
    for r in res:
        assert all(abs(a - b) <= 2 for a, b in zip(r["ksp"], r["ksp1"])), (r["ksp"], r["ksp1"])
        assert max(r["ksp"]) < 250


@pytest.mark.parametrize("world", [2, 3])
def test_nested_partition_distributes_the_coarse_levels(sf, world):
    """A nested hierarchy (children listed with their parent) partitioned by the ancestors on level 1: levels 1 and 2 are
    distributed (rank-local transfers and C_T coarsening, restriction completed by a halo sum on the coarse level),
    level 0 is replicated.  Same fields, Newton history and Krylov counts as the single-rank run."""
    res, hist1 = partitioned_vs_single(sf, world, "cg", levels=2, n_steps=2, multigrid=True, nested=True,
                                       min_cells_per_rank=100)
    for r in res:
        assert r["lc"] == 1 and r["level_cells"][0] == 48 and r["level_cells"][2] == 8 * r["level_cells"][1]
        assert all(abs(a - b) <= 2 for a, b in zip(r["ksp"], r["ksp1"])), (r["ksp"], r["ksp1"])
    assert sum(r["level_cells"][1] for r in res) == 384


def test_bench_multigrid_probe_on_emulated_ranks(sf):
    """bench.py decides between multigrid and block-Jacobi for N > 1 with a collective in-process probe; run it here on
    three emulated ranks (cube instead of the cavern so that it takes seconds), and check that a rank that raises makes
    EVERY rank report failure instead of hanging."""
    import bench
    from safeincave_b200 import cases
    from tests.hostemu.ranks import run_ranks
    case_fn = lambda g: cases.triaxial_case(g, n_steps=1, ksp_override="cg")
    res = run_ranks(3, lambda ctx: bench.probe_mg_ranks(ctx, levels=2, mesh="cube_coarse", case_fn=case_fn))
    assert all(ok for ok, _ in res) and all(msg.startswith("MG_PROBE_OK") for _, msg in res), res

    def broken(g):
        raise RuntimeError("simulated failure")
    from tests import hostemu
    hostemu.load().sic_emu_set_timeout(3)          # the surviving rank gives up in the emulated exchange after 3 s
    try:
        res = run_ranks(2, lambda ctx: bench.probe_mg_ranks(ctx, levels=2, mesh="cube_coarse", case_fn=broken if ctx.rank == 1 else case_fn))
    finally:
        hostemu.load().sic_emu_set_timeout(120)
    assert not any(ok for ok, _ in res), res


@pytest.mark.parametrize("world", [1, 3])
def test_guess_extrapolation_same_coefficients_on_every_rank(sf, world):
    """sic_guess_extrapolate (solver.cu): exact on a two-mode recurrence; on several ranks the inner products use owner
    weights and are summed in the exchange, so every rank applies the single-rank coefficients."""
    from safeincave_b200 import cases, distributed
    from safeincave_b200.mesh import TetMesh
    from safeincave_b200.multigrid import refine_hierarchy
    from tests.hostemu.ranks import run_ranks
    from tests import mg_checks as C
    h = refine_hierarchy(TetMesh.load_npz(os.path.join(GOLD, "mesh_cube_coarse.npz")), 1)
    tm = h.finest
    gg = sf.GridHandlerGMSH.from_hierarchy(h)
    case = cases.triaxial_case(gg, n_steps=1, ksp_override="cg")
    seq = C.recurrence_sequence(tm.n_nodes)

    def body(ctx):
        if ctx.world == 1:
            eq, _ = cases.build(case, gg)
            return C.check_guess_extrapolation(eq.engine, seq)
        grid, part = distributed.partition_grid(ctx, tm)
        eq, _ = cases.build(case, grid, part=part, ctx=ctx)
        return C.check_guess_extrapolation(eq.engine, seq, part.local_nodes)

    res = run_ranks(world, body)
    assert all(r == res[0] for r in res)


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_heat_and_thermomechanics_match_single_rank(sf, world):
    """HeatDiffusion on a partitioned grid (the reference's heat solve is MPI-parallel, HeatEquation.py:344-364): halo sums
    of the operator / right-hand side / diagonal, owner-weighted dot products, Dirichlet node sets from the global mesh;
    then Simulator_TM on the same partition.  Temperatures and the coupled displacement equal the single-rank run."""
    from safeincave_b200 import cases, distributed
    from safeincave_b200.mesh import TetMesh, red_refine
    from tests import heat_checks as H
    from tests.hostemu import EmuEngine
    from tests.hostemu.ranks import run_ranks
    import torch as to
    old = sf.HeatDiffusion.engine_cls
    sf.HeatDiffusion.engine_cls = EmuEngine
    try:
        tm = red_refine(TetMesh.load_npz(os.path.join(GOLD, "mesh_cube_coarse.npz")))
        gg = sf.GridHandlerGMSH.from_mesh(tm)
        tm = gg.tetmesh
        hc = H.heat_case(gg, "cube_coarse")
        n_steps, dt = 3, 0.5 * H.DAY

        def run(grid, part=None, ctx=None):
            hcl = dict(hc, T0=hc["T0"][part.local_nodes.cpu().numpy()] if part is not None else hc["T0"])
            heat = H.build_heat(sf, grid, hcl)
            case = cases.triaxial_case(gg, n_steps=n_steps)
            case["dt"], case["t_final_run"], case["thermo_alpha"] = dt, n_steps * dt, 44e-6
            eq, sim_m = cases.build(case, grid, part=part, ctx=ctx)
            eq.mat.add_to_thermoelastic(sf.Thermoelastic(case["thermo_alpha"] * to.ones(grid.n_elems, dtype=to.float64)))
            eq.set_material(eq.mat)
            sim = sf.Simulator_TM(eq, heat, sim_m.t_control, [], compute_elastic_response=True, verbose=False)
            hist = sim.run()
            return heat, eq, hist

        heat1, eq1, hist1 = run(gg)

        def body(ctx):
            grid, part = distributed.partition_grid(ctx, tm)
            heat, eq, hist = run(grid, part, ctx)
            ln = part.local_nodes
            rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
            return dict(e_T=rel(heat.T_dev[:heat.n_nodes], heat1.T_dev[ln]),
                        e_dT=rel(heat.T_dev[:heat.n_nodes] - to.as_tensor(hc["T0"])[ln], heat1.T_dev[ln] - to.as_tensor(hc["T0"])[ln]),
                        e_u=rel(eq.X, eq1.X[ln]), newton=[h["iterations"] for h in hist],
                        heat_its=[k[0] for k in heat.ksp_log])

        res = run_ranks(world, body)
        for r in res:
            assert r["e_T"] < 1e-11 and r["e_dT"] < 1e-8 and r["e_u"] < 1e-8, r
            assert r["newton"] == [h["iterations"] for h in hist1]
            assert all(abs(a - b) <= 1 for a, b in zip(r["heat_its"], [k[0] for k in heat1.ksp_log]))
    finally:
        sf.HeatDiffusion.engine_cls = old
