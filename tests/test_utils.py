"""Host utilities with the reference's names (safeincave_b200/Utils.py), mirroring the reference's tests/test_utils.py.
The UFL helpers tested there (dotdot_ufl, epsilon, tensor2voigt, voigt2tensor: test_utils.py:117-143) have no
counterpart -- what those forms compute is done by the CUDA kernels and checked in test_gpu_fem.py.  The golden of
create_field_nodes / create_field_elems (test_utils.py:154-168) belongs to grids/cube_regions, which is not in the
reference checkout; the samplers are checked against direct evaluation instead."""
import os

import numpy as np
import pytest
import torch as to

import safeincave_b200.Utils as ut

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def iso(a, b, c):
    C = np.zeros((6, 6))
    C[:3, :3] = b
    C[np.arange(3), np.arange(3)] = a
    C[np.arange(3, 6), np.arange(3, 6)] = c
    return C


def test_dotdot_torch_golden():
    """test_utils.py:40-115, 122-126: two distinct cells (the reference repeats them three times), rtol 1e-4."""
    eps = to.tensor([[[1., 4., 5.], [4., 2., 6.], [5., 6., 3.]], [[6., 1., 2.], [1., 5., 3.], [2., 3., 4.]]] * 3, dtype=to.float64)
    C = to.tensor(np.stack([iso(1.1111e+09, 2.7778e+08, 8.3333e+08), iso(2.6923e+09, 1.1538e+09, 1.5385e+09)] * 3))
    expected = to.tensor([[[2.5000e+09, 3.3333e+09, 4.1666e+09], [3.3333e+09, 3.3333e+09, 5.0000e+09], [4.1666e+09, 5.0000e+09, 4.1666e+09]],
                          [[2.6538e+10, 1.5385e+09, 3.0770e+09], [1.5385e+09, 2.5000e+10, 4.6155e+09], [3.0770e+09, 4.6155e+09, 2.3461e+10]]] * 3,
                         dtype=to.float64)
    sigma = ut.dotdot_torch(C, eps)
    assert isinstance(sigma, to.Tensor) and sigma.shape == (6, 3, 3)
    to.testing.assert_close(sigma, expected, rtol=1e-4, atol=1e-9)


def test_numpy2torch():
    """test_utils.py:145-151."""
    a = np.array([[1, 2, 3], [4, 5, 6]], dtype=np.float32)
    t = ut.numpy2torch(a)
    assert isinstance(t, to.Tensor) and t.dtype == to.float64 and t.shape == (2, 3)
    np.testing.assert_allclose(t.numpy(), a.astype(np.float64))


def test_fields_and_units(tmp_path):
    import safeincave_b200 as sf
    from safeincave_b200.mesh import TetMesh
    grid = sf.GridHandlerGMSH.from_mesh(TetMesh.load_npz(os.path.join(GOLD, "mesh_cube_coarse.npz")))
    fun = lambda x, y, z: x**2 + y**2 + z**2
    fn, fe = ut.create_field_nodes(grid, fun), ut.create_field_elems(grid, fun)
    x = grid.mesh.geometry.x
    assert fn.dtype == to.float64 and fn.shape == (grid.n_nodes,) and fe.shape == (grid.n_elems,)
    np.testing.assert_allclose(fn.numpy(), (x**2).sum(axis=1))
    c = x[grid.tetmesh.cells].mean(axis=1)
    np.testing.assert_allclose(fe.numpy(), (c**2).sum(axis=1))
    # scalar-only callables (the reference calls fun point by point, Utils.py:285-342)
    import math
    g = lambda x, y, z: math.exp(-z)
    np.testing.assert_allclose(ut.create_field_nodes(grid, g).numpy(), np.exp(-x[:, 2]))
    assert (ut.GPa, ut.MPa, ut.kPa, ut.minute, ut.hour, ut.day, ut.year) == (1e9, 1e6, 1e3, 60, 3600, 86400, 365 * 86400)
    p = tmp_path / "a.json"
    ut.save_json({"k": [1, 2.5]}, str(p))
    assert ut.read_json(str(p)) == {"k": [1, 2.5]}


def test_case_parameters_per_region():
    """cases.per_cell: a number is uniform, a {region: value} dict is resolved through the mesh's cell tags (the salt /
    overburden materials of BASELINE config 4, cases.overburden_tm_case)."""
    import os
    import numpy as np
    from safeincave_b200 import cases
    from safeincave_b200.mesh import TetMesh
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    tm = TetMesh.load_npz(os.path.join(gold, "mesh_cavern_overburden_coarse.npz"))
    assert (cases.per_cell(2.5, tm) == 2.5).all()
    rho = cases.per_cell({"Salt": 2200.0, "Overburden": 2800.0}, tm)
    salt = tm.cell_tags == tm.names[3]["Salt"]
    assert (rho[salt] == 2200.0).all() and (rho[~salt] == 2800.0).all() and 0 < salt.sum() < tm.n_cells
    import safeincave_b200 as sf
    case = cases.overburden_tm_case(sf.GridHandlerGMSH.from_mesh(tm), n_steps=2)
    assert len(case["dirichlet"]) == 9 and {d["boundary"] for d in case["dirichlet"]} >= {"West_salt", "East_ovb", "Bottom"}
    assert case["thermal"]["robin"][0]["boundary"] == "Cavern" and case["thermo_alpha"]["Overburden"] == 0.0
