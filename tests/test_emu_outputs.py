"""SaveFields (OutputHandler) driven by Simulator_M on the host-emulated back end: files, XDMF structure, values."""
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest


@pytest.fixture()
def sf():
    import safeincave_b200 as sf
    from tests.hostemu import EmuEngine
    old = sf.LinearMomentum.engine_cls
    sf.LinearMomentum.engine_cls = EmuEngine
    yield sf
    sf.LinearMomentum.engine_cls = old


def test_save_fields_time_series(sf, tmp_path):
    from safeincave_b200 import cases
    from tests.test_gpu_fem import load_grid
    grid = load_grid(sf, "cube_coarse")
    case = cases.triaxial_case(grid, n_steps=2)
    eq, _ = cases.build(case, grid)
    out = sf.SaveFields(eq)
    out.set_output_folder(str(tmp_path))
    for name, label in (("u", "Displacement (m)"), ("sig", "Stress (Pa)"), ("eps_tot", "Total strain (-)"),
                        ("p_elems", "Mean stress (Pa)"), ("q_nodes", "Von Mises stress (Pa)")):
        out.add_output_field(name, label)
    _, sim = cases.build(case, grid)
    sim = sf.Simulator_M(eq, sim.t_control, [out], True, verbose=False)
    sim.run()
    N, M = grid.n_elems, grid.n_nodes
    shapes = {"u": (M, 3, "Node", "Vector"), "sig": (N, 9, "Cell", "Tensor"), "eps_tot": (N, 9, "Cell", "Tensor"),
              "p_elems": (N, 1, "Cell", "Scalar"), "q_nodes": (M, 1, "Node", "Scalar")}
    for name, (rows, comp, center, kind) in shapes.items():
        folder = tmp_path / name
        root = ET.parse(folder / f"{name}.xdmf").getroot()
        grids = root.findall("./Domain/Grid/Grid")
        assert len(grids) == 3                                   # t = 0 and two steps
        times = [float(g.find("Time").get("Value")) for g in grids]
        assert times == [0.0, case["dt"], 2 * case["dt"]]
        att = grids[-1].find("Attribute")
        assert att.get("Center") == center and att.get("AttributeType") == kind
        data = np.fromfile(folder / att.find("DataItem").text, dtype="<f8").reshape(rows, comp)
        assert np.isfinite(data).all() and np.abs(data).max() > 0
    u = np.fromfile(tmp_path / "u" / "u_000002.bin", dtype="<f8")
    assert np.array_equal(u, eq.X.reshape(-1).cpu().numpy())     # last saved state is the final state
    sig = np.fromfile(tmp_path / "sig" / "sig_000002.bin", dtype="<f8").reshape(N, 3, 3)
    assert np.array_equal(sig, eq.sig.to_tensor().numpy()) and np.array_equal(sig, sig.transpose(0, 2, 1))
    topo = np.fromfile(tmp_path / "u" / "mesh_topology.bin", dtype="<i4").reshape(N, 4)
    assert np.array_equal(topo, grid.tetmesh.cells)
    assert os.path.isfile(tmp_path / "mesh" / "mesh.npz")
