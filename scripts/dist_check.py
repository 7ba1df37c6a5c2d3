"""torchrun --nproc-per-node N scripts/dist_check.py [--levels L] [--ksp cg] [--pc jacobi|mg] [--min-cells-per-rank K]
Runs the cavern case partitioned over N GPUs and (on every rank, on its own GPU) the same case on the whole mesh;
compares displacement / stress / creep strain on the rank's local nodes and cells.

--pc mg: multigrid CG with the bench's solver settings (extrapolated guess, lagged setup) on a NESTED hierarchy: every
level with at least --min-cells-per-rank cells per rank is partitioned (csrc/mg.cu), the ones below are replicated."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import safeincave_b200 as sf  # noqa: E402
from safeincave_b200 import cases, distributed  # noqa: E402
from safeincave_b200.mesh import TetMesh  # noqa: E402
from safeincave_b200.multigrid import refine_hierarchy  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--levels", type=int, default=1)
ap.add_argument("--ksp", default="cg")
ap.add_argument("--pc", default="jacobi", choices=["jacobi", "mg"])
ap.add_argument("--min-cells-per-rank", type=int, default=200_000)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--rtol", type=float, default=1e-12)
ap.add_argument("--tol", type=float, default=1e-8)
a = ap.parse_args()

ctx = distributed.init()
dev = ctx.device
h = refine_hierarchy(TetMesh.load_npz(os.path.join(ROOT, "tests/golden/mesh_cavern_regular.npz")), a.levels, device=dev,
                     nested=True)
tm = h.finest
gg = sf.GridHandlerGMSH.from_hierarchy(h)
case = cases.cavern_case(gg, n_steps=a.steps, ksp_type=a.ksp, rtol=a.rtol)


def settings(eq):
    eq.solver.single_reduction = os.environ.get("SIC_CGCG", "0") == "1"
    if a.pc == "mg":
        eq.solver.getPC().setType("mg")
        eq.solver.initial_guess_nonzero = True
        eq.solver.guess_extrapolation = True
        eq.solver.mg_setup_first = 2


grid, part = distributed.partition_grid(ctx, tm, hierarchy=h if a.pc == "mg" else None,
                                        min_cells_per_rank=a.min_cells_per_rank)
eq, sim = cases.build(case, grid, device=dev, part=part, ctx=ctx)
settings(eq)
sim.verbose = False
hist = sim.run()
lc = eq.mg.lc if eq.mg is not None else None
level_cells = [e.N for e in eq.mg.engines] if eq.mg is not None else None
# single-domain run of the same case on this rank's GPU
eq1, sim1 = cases.build(case, gg, device=dev)
settings(eq1)
sim1.verbose = False
hist1 = sim1.run()
ln = part.local_nodes.to(dev)
c0, c1 = part.cell_range
rel = lambda x, y: float((x - y).abs().max() / y.abs().max())
e_u = rel(eq.X, eq1.X[ln])
e_s = rel(eq.engine.sig[:, :eq.engine.N], eq1.engine.sig[:, c0:c1])
e_c = rel(eq.engine.elems[0].eps_old[:, :eq.engine.N], eq1.engine.elems[0].eps_old[:, c0:c1])
its = [r["iterations"] for r in hist], [r["iterations"] for r in hist1]
print(f"rank {ctx.rank}/{ctx.world}: cells {eq.engine.N} nodes {eq.engine.M} peers {part.peers} interface {part.n_interface} "
      f"| pc {a.pc} distributed from level {lc} level cells {level_cells} "
      f"| u err {e_u:.2e} sig err {e_s:.2e} eps_cr err {e_c:.2e} | newton {its} ksp {[r['ksp_iterations'] for r in hist]} "
      f"vs {[r['ksp_iterations'] for r in hist1]}", flush=True)
assert e_u < a.tol and e_s < a.tol and e_c < a.tol and its[0] == its[1]
ctx.barrier()
if ctx.rank == 0:
    print("DIST CHECK OK")
