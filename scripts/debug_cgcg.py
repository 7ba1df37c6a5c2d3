import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import safeincave_b200 as sf
from safeincave_b200 import cases
from safeincave_b200.mesh import TetMesh, red_refine, morton_order
levels = int(sys.argv[1]); warm = int(sys.argv[2])
tm = TetMesh.load_npz(os.path.join(ROOT, "tests/golden/mesh_cavern_regular.npz"))
for _ in range(levels):
    tm = red_refine(tm, device="cuda")
tm = morton_order(tm, device="cuda")
grid = sf.GridHandlerGMSH.from_mesh(tm, reorder=False)
case = cases.cavern_case(grid, n_steps=1, ksp_type="cg", rtol=1e-10)
eq, sim = cases.build(case, grid)
eq.solver.single_reduction = True
eq.solver.initial_guess_nonzero = bool(warm)
eq.solver.respect_max_it, eq.solver.max_it = True, 20000
sim.verbose = False
t0 = time.time(); sim.initialize(); torch.cuda.synchronize()
print("init", eq.ksp_log[-1], f"{time.time()-t0:.2f}s", flush=True)
sim.maxiter = 4
t0 = time.time(); rec = sim.step(); torch.cuda.synchronize()
print("step", rec["iterations"], rec["error"], eq.ksp_log[-4:], f"{time.time()-t0:.2f}s", flush=True)
