"""Pin oracle/heat.py by properties that need no reference run (the reference's heat path lives in DOLFINx/PETSc)."""
import os

import numpy as np

from oracle import heat as oh
from safeincave_b200.mesh import TetMesh, red_refine

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cube(levels=0):
    tm = TetMesh.load_npz(os.path.join(GOLD, "mesh_cube_coarse.npz"))
    for _ in range(levels):
        tm = red_refine(tm)
    return tm


def tag(tm, name):
    up = {n.upper(): t for n, t in tm.names[2].items()}
    return up[name]


def const(v, t_final=1e9):
    return dict(values=[v, v], time_values=[0.0, t_final])


def test_mass_and_stiffness_basics():
    tm = cube(1)
    N = tm.n_cells
    M, K = oh.mass_stiffness(tm.coords, tm.cells, 3.0 * np.ones(N), 2.0 * np.ones(N))
    vol = (tm.coords[:, 0].max() - tm.coords[:, 0].min()) * (tm.coords[:, 1].max() - tm.coords[:, 1].min()) * \
        (tm.coords[:, 2].max() - tm.coords[:, 2].min())
    one = np.ones(tm.n_nodes)
    assert abs(one @ (M @ one) - 3.0 * vol) < 1e-12 * vol            # int rho cp dx
    assert np.abs(K @ one).max() < 1e-12 * abs(K).max()               # constants are in the kernel of K
    z = tm.coords[:, 2]
    assert abs(z @ (K @ z) - 2.0 * vol) < 1e-12 * vol                 # int k |grad z|^2 dx
    assert abs(M - M.T).max() < 1e-14 * abs(M).max() and abs(K - K.T).max() < 1e-14 * abs(K).max()


def test_steady_linear_profile_between_dirichlet_planes_is_exact():
    tm = cube(1)
    N = tm.n_cells
    z0, z1 = tm.coords[:, 2].min(), tm.coords[:, 2].max()
    h = oh.OracleHeat(tm.coords, tm.cells, tm.tris, tm.tri_tags, np.ones(N), np.ones(N), 5.0 * np.ones(N),
                      [dict(tag=tag(tm, "BOTTOM"), **const(300.0)), dict(tag=tag(tm, "TOP"), **const(350.0))], [], [])
    h.set_initial_T(np.full(tm.n_nodes, 300.0))
    for _ in range(3):
        h.step(1.0, 1e12)                                             # dt -> infinity: steady state
    exact = 300.0 + 50.0 * (tm.coords[:, 2] - z0) / (z1 - z0)
    assert np.abs(h.T - exact).max() < 1e-9


def test_energy_balance_with_neumann_flux():
    tm = cube(1)
    N = tm.n_cells
    rho, cp = 2200.0, 850.0
    q_in = 10.0                                                        # W/m2 into the domain through TOP
    h = oh.OracleHeat(tm.coords, tm.cells, tm.tris, tm.tri_tags, rho * np.ones(N), cp * np.ones(N), 7.0 * np.ones(N),
                      [], [dict(tag=tag(tm, "TOP"), **const(q_in))], [])
    h.set_initial_T(np.full(tm.n_nodes, 300.0))
    dt, steps = 3600.0, 5
    for i in range(steps):
        h.step((i + 1) * dt, dt)
    one = np.ones(tm.n_nodes)
    area_top = oh.tri_areas(tm.coords, tm.tris[tm.tri_tags == tag(tm, "TOP")]).sum()
    stored = one @ (h.M @ (h.T - 300.0))
    assert abs(stored - q_in * area_top * dt * steps) < 1e-9 * stored


def test_robin_relaxes_to_T_inf():
    tm = cube(0)
    N = tm.n_cells
    robin = [dict(tag=t, h=50.0, **const(280.0)) for t in np.unique(tm.tri_tags)]
    h = oh.OracleHeat(tm.coords, tm.cells, tm.tris, tm.tri_tags, np.ones(N), np.ones(N), 100.0 * np.ones(N), [], [], robin)
    h.set_initial_T(np.full(tm.n_nodes, 300.0))
    prev = 20.0
    for i in range(40):
        h.step(i + 1.0, 0.05)
        dev = np.abs(h.T - 280.0).max()
        assert dev <= prev + 1e-12
        prev = dev
    assert prev < 1e-3


def test_transient_1d_conduction_converges_to_the_series_solution():
    """Slab 0 < z < L, T(0) = T(L) = 0, T(z,0) = sin(pi z/L): T = exp(-kappa pi^2 t / L^2) sin(pi z/L)."""
    errs = []
    for lv in (2, 3):
        tm = cube(lv)
        N = tm.n_cells
        z0, z1 = tm.coords[:, 2].min(), tm.coords[:, 2].max()
        L = z1 - z0
        kappa = 1e-2
        h = oh.OracleHeat(tm.coords, tm.cells, tm.tris, tm.tri_tags, np.ones(N), np.ones(N), kappa * np.ones(N),
                          [dict(tag=tag(tm, "BOTTOM"), **const(0.0)), dict(tag=tag(tm, "TOP"), **const(0.0))], [], [])
        s = np.sin(np.pi * (tm.coords[:, 2] - z0) / L)
        h.set_initial_T(s)
        t_end, n = 0.2 * L * L / (kappa * np.pi ** 2), 100
        for i in range(n):
            h.step((i + 1) * t_end / n, t_end / n)
        exact = np.exp(-kappa * np.pi ** 2 * t_end / L ** 2) * s
        errs.append(np.abs(h.T - exact).max())
    # 4..16 cells across the slab: still pre-asymptotic on the boundary nodes (ratios 0.40, 0.36 -> 0.25)
    assert errs[1] < 0.45 * errs[0] and errs[1] < 0.03
