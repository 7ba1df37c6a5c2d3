"""Grid services through the reference's GridHandlerGMSH surface, mirroring the reference's tests/test_grid.py:17-64
(names, physical tags, MeshTags-like views) on its own cube grid (tests/files/cube_coarse, same physical names and
tags as grids/cube_regions), plus get_parameter (Grid.py:538-579) and the region maps (Grid.py:496-536)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def grid():
    import safeincave_b200 as sf
    from safeincave_b200.mesh import TetMesh
    return sf.GridHandlerGMSH.from_mesh(TetMesh.load_npz(os.path.join(GOLD, "mesh_cube_coarse.npz")))


def test_boundaries(grid):
    names = grid.get_boundary_names()
    assert names == ["NORTH", "SOUTH", "WEST", "EAST", "BOTTOM", "TOP"]                 # test_grid.py:20, 33-39
    assert [grid.get_boundary_tag(n) for n in names] == [21, 22, 23, 24, 25, 26]       # :41-46
    b = grid.get_boundaries()
    assert b.dim == 2 and b.values.shape == b.indices.shape == (grid.tetmesh.tris.shape[0],)
    assert set(b.values.tolist()) == {21, 22, 23, 24, 25, 26}
    for n in names:     # every tagged facet lies in the plane its name says
        tag = grid.get_boundary_tag(n)
        assert sorted(grid.get_boundary_tags(n)) == sorted(b.find(tag).tolist())
        x = grid.tetmesh.coords[grid.tetmesh.tris[b.find(tag)]]
        axis, side = {"WEST": (0, 0), "EAST": (0, 1), "SOUTH": (1, 0), "NORTH": (1, 1), "BOTTOM": (2, 0), "TOP": (2, 1)}[n]
        target = grid.tetmesh.coords[:, axis].min() if side == 0 else grid.tetmesh.coords[:, axis].max()
        assert np.allclose(x[..., axis], target)
    assert grid.get_boundary_tag(None) is None and grid.get_boundary_tags(None) is None


def test_subdomains(grid):
    names = grid.get_subdomain_names()
    assert names == ["OMEGA_A", "OMEGA_B"]                                              # test_grid.py:23, 57-59
    assert [grid.get_subdomain_tag(n) for n in names] == [27, 28]                       # :60-61
    s = grid.get_subdomains()
    assert s.dim == 3 and s.values.shape == (grid.n_elems,) and set(s.values.tolist()) == {27, 28}
    # region maps index cells in this package's own order and partition them
    idx = grid.region_indices
    assert sorted(idx["OMEGA_A"] + idx["OMEGA_B"]) == list(range(grid.n_elems))
    assert np.all(s.values[idx["OMEGA_A"]] == 27) and np.all(s.values[idx["OMEGA_B"]] == 28)
    assert grid.n_regions == 2 and grid.tags_dict == {27: "OMEGA_A", 28: "OMEGA_B"}


def test_get_parameter(grid):
    n = grid.n_elems
    p = grid.get_parameter(3.5)
    assert p.shape == (n,) and p.dtype == torch.float32 and bool((p == 3.5).all())      # float32 as in the reference (T1)
    q = grid.get_parameter([1.0, 2.0])
    assert bool((q[grid.region_indices["OMEGA_A"]] == 1.0).all()) and bool((q[grid.region_indices["OMEGA_B"]] == 2.0).all())
    r = grid.get_parameter(list(range(n)))
    assert r.tolist() == list(range(n))
    with pytest.raises(Exception):
        grid.get_parameter([1.0, 2.0, 3.0])


def test_mesh_views(grid):
    """What user scripts read (nobian/Simulation/Run.py:1370): geometry, connectivity, extents, comm."""
    assert grid.mesh.geometry.x.shape == (grid.n_nodes, 3)
    assert grid.mesh.topology.connectivity(3, 0).array.shape == (4 * grid.n_elems,)
    assert grid.mesh.comm.rank == 0 and grid.mesh.comm.size == 1
    assert (grid.Lx, grid.Ly, grid.Lz) == pytest.approx((1.0, 1.0, 1.0))
    assert grid.volumes.sum() == pytest.approx(grid.Lx * grid.Ly * grid.Lz)


def test_refine_builds_a_multigrid_hierarchy_on_a_loaded_mesh():
    """GridHandlerGMSH.refine: the loaded gmsh mesh becomes the coarsest level of a nested hierarchy (PC mg for any .msh);
    tags are inherited, volumes add up, the children of a cell are listed together."""
    import os
    import numpy as np
    import safeincave_b200 as sf
    from safeincave_b200.mesh import TetMesh
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g0 = sf.GridHandlerGMSH.from_mesh(TetMesh.load_npz(os.path.join(gold, "mesh_cube_coarse.npz")))
    g1 = g0.refine(1)
    assert g1.hierarchy is not None and g1.hierarchy.n_levels == 2 and g1.n_elems == 8 * g0.n_elems
    assert g1.get_boundary_names() == g0.get_boundary_names() and g1.get_subdomain_names() == g0.get_subdomain_names()
    assert np.isclose(np.asarray(g1.volumes).sum(), np.asarray(g0.volumes).sum(), rtol=1e-12)
    ch = g1.hierarchy.transfers[1].children           # (8, n_coarse): child j of coarse cell c is fine cell 8 c + j
    assert (ch == 8 * np.arange(g0.n_elems)[None, :] + np.arange(8)[:, None]).all()
    assert (g1.tetmesh.cell_tags.reshape(-1, 8) == g1.hierarchy.meshes[0].cell_tags[:, None]).all()
