"""Host-side multigrid plumbing (safeincave_b200/multigrid.py): the transfer tables of the nested
red-refinement hierarchy, and the identity the whole design rests on -- the Galerkin coarse operator
P^T K P equals the operator assembled on the coarse mesh with the MEAN of the children's C_T."""
import os

import numpy as np
import scipy.sparse as sp

from oracle import fem
from oracle import constitutive as oc
from safeincave_b200 import multigrid as mg
from safeincave_b200.mesh import TetMesh

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _vol(m):
    x = m.coords[m.cells]
    e = x[:, 1:] - x[:, :1]
    return np.abs(np.einsum("ni,ni->n", e[:, 0], np.cross(e[:, 1], e[:, 2]))) / 6


def test_transfer_tables_cube():
    h = mg.refine_hierarchy(TetMesh.load_npz(os.path.join(GOLD, "mesh_cube_coarse.npz")), 2)
    assert [m.n_cells for m in h.meshes] == [48, 384, 3072]
    for l in (1, 2):
        t, c, f = h.transfers[l], h.meshes[l - 1], h.meshes[l]
        P = mg.prolongation_matrix(t, c.n_nodes)
        lin = lambda x: 1 + 2 * x[:, 0] - 3 * x[:, 1] + 0.5 * x[:, 2]
        assert np.abs(P @ lin(c.coords) - lin(f.coords)).max() < 1e-14          # P1 interpolation is exact on linears
        assert np.abs(f.coords[t.inject] - c.coords).max() == 0.0                 # coarse nodes survive
        vc, vf = _vol(c), _vol(f)
        assert np.abs(vf[t.children] * 8 - vc[None, :]).max() < 1e-15 * vc.max() * 8   # 8 equal-volume children
        r = np.random.default_rng(0).standard_normal(f.n_nodes)
        out = np.array([0.5 * r[t.rst_idx[t.rst_ptr[k]:t.rst_ptr[k + 1]]].sum() for k in range(c.n_nodes)])
        assert np.abs(out - P.T @ r).max() < 1e-13                                 # restriction CSR = P^T
        # boundary tags are inherited: tagged triangle area per tag is conserved
        for tag in np.unique(c.tri_tags):
            area = lambda m: np.linalg.norm(np.cross(m.coords[m.tris[m.tri_tags == tag]][:, 1] - m.coords[m.tris[m.tri_tags == tag]][:, 0],
                                                     m.coords[m.tris[m.tri_tags == tag]][:, 2] - m.coords[m.tris[m.tri_tags == tag]][:, 0]), axis=1).sum()
            assert abs(area(c) - area(f)) < 1e-12 * area(c)


def test_galerkin_identity_with_mean_tangent():
    h = mg.refine_hierarchy(TetMesh.load_npz(os.path.join(GOLD, "mesh_cube_coarse.npz")), 1)
    c, f, t = h.meshes[0], h.meshes[1], h.transfers[1]
    rng = np.random.default_rng(4)
    CTf = oc.iso_matrix(1e9 * (1 + rng.random(f.n_cells)), 0.3 * np.ones(f.n_cells)) * (1 + 0.05 * rng.standard_normal((f.n_cells, 6, 6)))
    CTc = CTf[t.children.astype(np.int64)].mean(axis=0)
    Kf = fem.assemble_K(f.coords, f.cells, CTf)
    Kc = fem.assemble_K(c.coords, c.cells, CTc)
    P = sp.kron(mg.prolongation_matrix(t, c.n_nodes), sp.identity(3), format="csr")
    G = (P.T @ Kf @ P).tocsr()
    assert abs(G - Kc).max() < 1e-13 * abs(Kc).max()
