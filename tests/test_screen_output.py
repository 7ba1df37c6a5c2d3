"""ScreenPrinter (safeincave_b200/ScreenOutput.py): the report and log.txt keep the layout that the reference's
post-processing scripts parse (ScreenOutput.py:38-571 of the reference; parsers cited in the module docstring)."""
import os

import pytest


@pytest.fixture(scope="module")
def sf():
    import safeincave_b200 as sf
    from tests.hostemu import EmuEngine
    old = sf.LinearMomentum.engine_cls
    sf.LinearMomentum.engine_cls = EmuEngine
    yield sf
    sf.LinearMomentum.engine_cls = old


def parse_newton_iters(log_path):
    """What examples/mechanics/nobian/Simulation/Run_sensitivity.py:353-373 does with log.txt."""
    out = []
    with open(log_path) as fh:
        for line in fh:
            s = line.strip()
            if not s or not s.startswith("|"):
                continue
            parts = [p.strip() for p in s.strip("|").split("|")]
            if len(parts) < 4:
                continue
            try:
                int(parts[0])
                out.append(int(parts[3]))
            except ValueError:
                continue
    return out


def test_run_writes_the_reference_log_layout(sf, tmp_path, capsys):
    from safeincave_b200 import cases
    from safeincave_b200.mesh import TetMesh
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    grid = sf.GridHandlerGMSH.from_mesh(TetMesh.load_npz(os.path.join(gold, "mesh_cube_coarse.npz")))
    case = cases.triaxial_case(grid, n_steps=3)
    eq, sim = cases.build(case, grid, device="cpu")
    out = sf.SaveFields(eq)
    out.set_output_folder(str(tmp_path / "case"))
    out.add_output_field("u", "Displacement (m)")
    sim.outputs, sim.verbose = [out], True
    hist = sim.run()
    log_path = tmp_path / "case" / "log.txt"
    log = log_path.read_text().splitlines()
    assert capsys.readouterr().out.splitlines() == log[1:]            # the log starts with a newline, as the reference's
    # every line is framed to the banner width (cells longer than their column -- here the output folder -- are not
    # truncated, as in the reference)
    assert all(len(ln) == 97 for ln in log[1:] if str(tmp_path) not in ln)
    # the step table: Newton iterations in the 4th cell of integer-led rows
    rows = parse_newton_iters(str(log_path))
    assert rows[-3:] == [h["iterations"] for h in hist]
    # examples/mechanics/4_cavern/plot_results.py:139-150: sizes on the 4th line after "| Mesh info:"
    i = next(k for k, ln in enumerate(log) if "| Mesh info:" in ln)
    cells = log[i + 4].split("|")
    assert int(cells[1]) == 48 and int(cells[2]) == 23
    assert log[-1].strip().startswith("Total time: ") and log[-1].rstrip().endswith("seconds)")
    for title in (" Partition(s) info:", " Solver info:", " Constitutive model:", " Output info:"):
        assert any(ln.startswith("|" + title) for ln in log)
    assert any("kelvin, creep" in ln for ln in log) and any("bicg" in ln and "asm" in ln for ln in log)


def test_table_rendering_rules():
    """Narrow tables are closed with ' |' and padded; a table as wide as the frame is not (ScreenOutput.py:455-506)."""
    from safeincave_b200.ScreenOutput import _Table, FRAME_WIDTH
    t = _Table(["a" * 10, "b" * 5], "center")
    assert t.divider() == "+" + "-" * 12 + "+" + "-" * 7 + "+" + "-" * (FRAME_WIDTH - 23) + "+"
    line = t.line([3, "x"], ["center", "left"], ["%i", "%s"])
    assert line == "|     3      | x     |" + " " * (FRAME_WIDTH - 23) + "|"
    wide = _Table(["c" * 45, "d" * 45], "left")          # no room for the closing ' |': one pad and the frame bar
    assert wide.line(["1", "2"], ["left", "left"]) == "| 1" + " " * 44 + " | 2" + " " * 44 + " |"
    assert wide.divider() == "+" + "-" * 47 + "+" + "-" * 47 + "+"
    assert _Table.cell(1.23456, 8, "right", "%.2f") == "    1.23"
