"""Scratch: time sic_apply (memset + operator kernel + mask) on a refined cavern mesh."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import safeincave_b200 as sf
from safeincave_b200.mesh import TetMesh, red_refine, morton_order
levels = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tm = TetMesh.load_npz(os.path.join(ROOT, "tests/golden/mesh_cavern_regular.npz"))
for _ in range(levels):
    tm = red_refine(tm, device="cuda")
tm = morton_order(tm, device="cuda")
grid = sf.GridHandlerGMSH.from_mesh(tm, reorder=False)
eq = sf.LinearMomentum(grid, theta=0.0)
eng = eq.engine
one = torch.ones(eng.N, dtype=torch.float64)
mat = sf.Material(eng.N); mat.add_to_elastic(sf.Spring(102e9 * one, 0.3 * one)); eq.set_material(mat)
eng.elastic_tangent()
x = torch.randn(eng.M, 3, dtype=torch.float64, device="cuda")
y = torch.zeros_like(x)
for _ in range(5):
    eng.apply(x, y, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 200
e0.record()
for _ in range(n):
    eng.apply(x, y, None)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
byt = eng.N * 408 + eng.M * 72
print(f"{'libsafeincave_cuda.so':60s} cells {eng.N} apply {ms*1e3:.1f} us  {byt/ms/1e6:.0f} GB/s")
