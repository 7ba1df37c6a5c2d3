"""torchrun --nproc-per-node N scripts/dist_smoke.py [levels]: the partitioned cavern case, 1 time step, with a
checksum that must be identical for every N (sum over OWNED nodes of |u|^2, global energy-like number)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import safeincave_b200 as sf
from safeincave_b200 import cases, distributed
from safeincave_b200.mesh import TetMesh, red_refine, morton_order
levels = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx = distributed.init()
dev = ctx.device
tm = TetMesh.load_npz(os.path.join(ROOT, "tests/golden/mesh_cavern_regular.npz"))
for _ in range(levels):
    tm = red_refine(tm, device=dev)
tm = morton_order(tm, device=dev)
gg = sf.GridHandlerGMSH.from_mesh(tm, reorder=False)
case = cases.cavern_case(gg, n_steps=1, ksp_type="cg", rtol=1e-12)
t0 = time.time()
if ctx.world > 1:
    grid, part = distributed.partition_grid(ctx, tm)
    eq, sim = cases.build(case, grid, device=dev, part=part, ctx=ctx)
    w = eq.engine.owner_w
else:
    eq, sim = cases.build(case, gg, device=dev)
    part, w = None, torch.ones(eq.engine.M, dtype=torch.float64, device=dev)
sim.verbose = False
hist = sim.run()
chk = torch.stack([(w[:, None] * eq.X ** 2).sum(), (eq.engine.sig[:, :eq.engine.N] ** 2).sum()])
ctx.all_reduce_sum(chk)
torch.cuda.synchronize()
if ctx.rank == 0:
    print(f"N={ctx.world} cells {tm.n_cells} peers {part.peers if part else []} newton {[h['iterations'] for h in hist]} "
          f"ksp {[h['ksp_iterations'] for h in hist]} checksum |u|^2 {chk[0].item():.12e} |sig|^2 {chk[1].item():.12e} "
          f"wall {time.time()-t0:.1f}s", flush=True)
if ctx.world > 1:
    torch.distributed.destroy_process_group()
