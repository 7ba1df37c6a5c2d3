"""The threaded CPU arm of bench.py (oracle/cpu_step.py) against the plain oracle (sparse LU, one chunk)."""
import os

import numpy as np

from safeincave_b200.mesh import TetMesh, red_refine

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_threaded_cpu_step_matches_the_oracle():
    import safeincave_b200 as sf
    from oracle.cpu_step import threaded_simulator
    from safeincave_b200 import cases
    from tests.case_oracle import oracle_simulator
    tm = red_refine(TetMesh.load_npz(os.path.join(GOLD, "mesh_cube_coarse.npz")))
    grid = sf.GridHandlerGMSH.from_mesh(tm)
    case = cases.triaxial_case(grid, n_steps=3)
    ref = oracle_simulator(case, tm).run(0.0, [case["dt"]] * 3)
    sim = threaded_simulator(case, tm, n_threads=3, rtol=1e-12)
    out = sim.run(0.0, [case["dt"]] * 3)
    assert [h["iters"] for h in out] == [h["iters"] for h in ref]
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    assert rel(out[-1]["u"], ref[-1]["u"]) < 1e-9 and rel(out[-1]["sig"], ref[-1]["sig"]) < 1e-9
    assert len(sim.krylov_iterations) == 1 + sum(h["iters"] for h in out) and max(sim.krylov_iterations) < 500
