"""torchrun --nproc-per-node N scripts/dist_check.py [levels]
Runs the cavern case partitioned over N GPUs and (on every rank, on its own GPU) the same case on the
whole mesh; compares displacement / stress / creep strain on the rank's local nodes and cells."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import safeincave_b200 as sf
from safeincave_b200 import cases, distributed
from safeincave_b200.mesh import TetMesh, red_refine, morton_order

levels = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ksp = sys.argv[2] if len(sys.argv) > 2 else "cg"
ctx = distributed.init()
dev = ctx.device
tm = TetMesh.load_npz(os.path.join(ROOT, "tests/golden/mesh_cavern_regular.npz"))
for _ in range(levels):
    tm = red_refine(tm, device=dev)
tm = morton_order(tm, device=dev)
gg = sf.GridHandlerGMSH.from_mesh(tm, reorder=False)
case = cases.cavern_case(gg, n_steps=2, ksp_type=ksp, rtol=1e-12)
grid, part = distributed.partition_grid(ctx, tm)
eq, sim = cases.build(case, grid, device=dev, part=part, ctx=ctx)
eq.solver.single_reduction = os.environ.get("SIC_CGCG", "0") == "1"
sim.verbose = False
hist = sim.run()
# single-domain run of the same case on this rank's GPU
eq1, sim1 = cases.build(case, gg, device=dev)
eq1.solver.single_reduction = eq.solver.single_reduction
sim1.verbose = False
hist1 = sim1.run()
ln = part.local_nodes.to(dev)
c0, c1 = part.cell_range
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
e_u = rel(eq.X, eq1.X[ln])
e_s = rel(eq.engine.sig[:, :eq.engine.N], eq1.engine.sig[:, c0:c1])
e_c = rel(eq.engine.elems[0].eps_old[:, :eq.engine.N], eq1.engine.elems[0].eps_old[:, c0:c1])
its = [h["iterations"] for h in hist], [h["iterations"] for h in hist1]
print(f"rank {ctx.rank}/{ctx.world}: cells {eq.engine.N} nodes {eq.engine.M} peers {part.peers} interface {part.n_interface} "
      f"| u err {e_u:.2e} sig err {e_s:.2e} eps_cr err {e_c:.2e} | newton {its} ksp {[h['ksp_iterations'] for h in hist]} vs {[h['ksp_iterations'] for h in hist1]}",
      flush=True)
assert e_u < 1e-8 and e_s < 1e-8 and e_c < 1e-8 and its[0] == its[1]
ctx.barrier()
if ctx.rank == 0:
    print("DIST CHECK OK")
