"""Convert the reference's Gmsh grids into compact .npz mesh fixtures.

TEST INFRASTRUCTURE.  Run in the build container (needs /root/reference):

    python oracle/gen_mesh_fixtures.py

/root/reference does not exist on the GPU box, so the grids the tests and bench.py need are read
here with this repository's own MSH reader (safeincave_b200/mesh.py) and stored as arrays
(coordinates, tetrahedra, region ids, outward-oriented tagged boundary triangles, physical names):
    tests/files/cube_coarse/geom.msh          -> tests/golden/mesh_cube_coarse.npz      (23 nodes / 48 tets, MSH 2.2)
    grids/cavern_regular/geom.msh             -> tests/golden/mesh_cavern_regular.npz   (3577 / 14346, MSH 4.1)
    grids/cavern_overburden_coarse/geom.msh   -> tests/golden/mesh_cavern_overburden_coarse.npz (5916 / 25608, MSH 2.2)
    grids/cavern_irregular_finemesh/geom.msh  -> tests/golden/mesh_cavern_irregular_finemesh.npz (BASELINE configs[2])
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from safeincave_b200.mesh import read_msh  # noqa: E402

REF = os.environ.get("SAFEINCAVE_REFERENCE", "/root/reference")
JOBS = [
    ("tests/files/cube_coarse/geom.msh", "mesh_cube_coarse.npz"),
    ("grids/cavern_regular/geom.msh", "mesh_cavern_regular.npz"),
    ("grids/cavern_overburden_coarse/geom.msh", "mesh_cavern_overburden_coarse.npz"),
    ("grids/cavern_irregular_finemesh/geom.msh", "mesh_cavern_irregular_finemesh.npz"),
]

if __name__ == "__main__":
    for src, dst in JOBS:
        m = read_msh(os.path.join(REF, src))
        out = os.path.join(ROOT, "tests", "golden", dst)
        m.save_npz(out)
        print(f"{src}: {m.n_nodes} nodes, {m.n_cells} tets, {m.tris.shape[0]} tagged triangles, names {m.names} -> "
              f"{dst} ({os.path.getsize(out)/1024:.0f} KiB)")
