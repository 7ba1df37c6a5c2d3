"""Scratch: iteration counts and time per step of the cavern case vs refinement level / KSP type."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import safeincave_b200 as sf
from safeincave_b200 import cases
from safeincave_b200.mesh import TetMesh, red_refine, morton_order

levels = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 1, 2]
ksps = sys.argv[2].split(",") if len(sys.argv) > 2 else ["cg", "bicg"]
rtol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-10
tm0 = TetMesh.load_npz(os.path.join(ROOT, "tests/golden/mesh_cavern_regular.npz"))
for L in levels:
    tm = tm0
    t0 = time.time()
    for _ in range(L):
        tm = red_refine(tm, device="cuda")
    tm = morton_order(tm, device="cuda")
    grid = sf.GridHandlerGMSH.from_mesh(tm, reorder=False)
    print(f"level {L}: {tm.n_cells} cells {tm.n_nodes} nodes, mesh build {time.time()-t0:.1f}s", flush=True)
    for k in ksps:
        case = cases.cavern_case(grid, n_steps=3, ksp_type=k, rtol=rtol)
        eq, sim = cases.build(case, grid)
        eq.engine.time_operator = True
        torch.cuda.synchronize(); t0 = time.time()
        sim.initialize()
        torch.cuda.synchronize(); t1 = time.time()
        print(f"  {k}: elastic solve {eq.ksp_log[-1]} in {t1-t0:.2f}s", flush=True)
        for s in range(3):
            torch.cuda.synchronize(); t0 = time.time()
            rec = sim.step()
            torch.cuda.synchronize(); t1 = time.time()
            e = eq.engine
            print(f"  {k}: step {s} newton {rec['iterations']} err {rec['error']:.2e} ksp its {rec['ksp_iterations']} "
                  f"{t1-t0:.2f}s  op {e.op_ms/max(e.op_samples,1):.3f} ms/apply  reasons {[l[1] for l in eq.ksp_log[-rec['iterations']:]]}", flush=True)
        del eq, sim
        torch.cuda.empty_cache()
