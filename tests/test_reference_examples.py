"""Drop-in check at the level a user sees: the reference's OWN example script
examples/mechanics/1_triaxial/main.py is executed UNMODIFIED against this package (safeincave_b200.compat registers
stand-ins for `safeincave`, `petsc4py`, `dolfinx`, `mpi4py`), on the host-emulated back end, and its results are
compared with the CPU oracle.  Runs only where /root/reference is mounted (the build container)."""
import os
import runpy
import sys

import numpy as np
import pytest

REF = os.environ.get("SAFEINCAVE_REFERENCE", "/root/reference")
EXAMPLE = os.path.join(REF, "examples", "mechanics", "1_triaxial", "main.py")
CUBE = os.path.join(REF, "tests", "files", "cube_coarse")

pytestmark = pytest.mark.skipif(not (os.path.isfile(EXAMPLE) and os.path.isdir(CUBE)),
                                reason="reference checkout not available")


def test_reference_triaxial_example_runs_unmodified(tmp_path, monkeypatch):
    import safeincave_b200 as sf
    from safeincave_b200 import compat
    from tests.hostemu import EmuEngine
    monkeypatch.setattr(sf.LinearMomentum, "engine_cls", EmuEngine)
    registered = compat.install()
    try:
        # the script reads ../../../grids/cube relative to its own folder (that grid is not in the checkout: the
        # reference's tests use tests/files/cube_coarse) and writes output/case_0 next to itself
        work = tmp_path / "examples" / "mechanics" / "1_triaxial"
        work.mkdir(parents=True)
        (tmp_path / "grids").mkdir()
        os.symlink(CUBE, tmp_path / "grids" / "cube")
        monkeypatch.chdir(work)
        # 6 of the 48 half-hour steps are enough here: cap the final time the script asks for
        real_tc = sf.TimeController
        captured = {}

        class ShortTimeController(real_tc):
            def __init__(self, dt, initial_time, final_time, time_unit="second"):
                super().__init__(dt=dt, initial_time=initial_time, final_time=min(final_time, 3.0), time_unit=time_unit)

        real_sim = sf.Simulator_M

        class RecordingSimulator(real_sim):
            def __init__(self, *a, **k):
                super().__init__(*a, **k)
                self.verbose = False
                captured["sim"] = self

        monkeypatch.setattr(sf, "TimeController", ShortTimeController)
        monkeypatch.setattr(sf, "Simulator_M", RecordingSimulator)
        ns = runpy.run_path(EXAMPLE, run_name="not_main")
        ns["main"]()
    finally:
        for name in registered:
            for key in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
                del sys.modules[key]
    sim = captured["sim"]
    eq = sim.eq_mom
    assert len(sim.history) == 6 and all(h["converged"] for h in sim.history)
    # the script's own hook ran: its Functions hold the element strains / yield function of the final state
    assert np.array_equal(eq.eps_cr.x.array, eq.mat.elems_ne[1].eps_ne_k.numpy().reshape(-1))
    assert np.abs(eq.eps_cr.x.array).max() > 0
    assert eq.Fvp.x.array.shape == (eq.n_elems,)
    # and its SaveFields wrote the series it asked for
    for field in ("u", "eps_tot", "eps_ve", "eps_cr", "eps_vp", "Fvp"):
        assert os.path.isfile(work / "output" / "case_0" / field / f"{field}.xdmf")
    # same physics through the oracle (float32 user tensors as in the script: T1 deviation <= 1e-6)
    from safeincave_b200 import cases
    from tests.case_oracle import oracle_simulator
    case = cases.triaxial_case(eq.grid, elements=("kelvin", "dislocation", "desai"))
    osim = oracle_simulator(case, eq.grid.tetmesh)
    ohist = osim.run(0.0, [case["dt"]] * 6)
    rel = lambda a, b: float(np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max())
    assert [h["iterations"] for h in sim.history] == [h["iters"] for h in ohist[1:]]
    assert rel(eq.X.reshape(-1).numpy(), ohist[-1]["u"]) < 5e-6
    assert rel(eq.engine.get6(eq.engine.sig), ohist[-1]["sig"]) < 5e-6


@pytest.mark.skipif(not os.environ.get("SIC_SLOW"), reason="~13 min under emulation (25 608 cells, BiCGStab); set SIC_SLOW=1")
def test_reference_thermomechanics_cavern_example_runs_unmodified(tmp_path, monkeypatch):
    """BASELINE config 4's own script, examples/thermomechanics/2_cavern/main.py, on its own grid
    (grids/cavern_overburden_coarse: salt + overburden, region-wise parameters): an equilibrium stage with Simulator_M and
    a parabolic time controller, then the operation stage with HeatDiffusion + Simulator_TM (Dirichlet / Neumann / Robin
    heat BCs, thermal strain, Kelvin + dislocation + pressure-solution creep).  Executed unmodified; two steps per stage."""
    import safeincave_b200 as sf
    from safeincave_b200 import compat
    from tests.hostemu import EmuEngine
    example = os.path.join(REF, "examples", "thermomechanics", "2_cavern", "main.py")
    if not os.path.isfile(example):
        pytest.skip("example not in the reference checkout")
    monkeypatch.setattr(sf.LinearMomentum, "engine_cls", EmuEngine)
    monkeypatch.setattr(sf.HeatDiffusion, "engine_cls", EmuEngine)
    registered = compat.install()
    sims = []
    try:
        work = tmp_path / "examples" / "thermomechanics" / "2_cavern"
        work.mkdir(parents=True)
        os.symlink(os.path.join(REF, "grids"), tmp_path / "grids")
        monkeypatch.chdir(work)
        real_tc, real_tcp = sf.TimeController, sf.TimeControllerParabolic

        class TC(real_tc):
            def __init__(self, dt, initial_time, final_time, time_unit="second"):
                super().__init__(dt=dt, initial_time=initial_time, final_time=min(final_time, 2 * dt), time_unit=time_unit)

        class TCP(real_tcp):
            def keep_looping(self):
                return self.step_counter < 2 and super().keep_looping()

        monkeypatch.setattr(sf, "TimeController", TC)
        monkeypatch.setattr(sf, "TimeControllerParabolic", TCP)
        for name in ("Simulator_M", "Simulator_TM"):
            base = getattr(sf, name)

            class Rec(base):
                def __init__(self, *a, **k):
                    super().__init__(*a, **k)
                    self.verbose = False
                    sims.append(self)
            monkeypatch.setattr(sf, name, Rec)
        runpy.run_path(example, run_name="not_main")["main"]()
    finally:
        for name in registered:
            for key in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
                del sys.modules[key]
    sim_m, sim_tm = sims
    assert len(sim_m.history) == 2 and all(h["converged"] for h in sim_m.history)
    assert len(sim_tm.history) == 2 and all(h["error"] <= 1e-6 and h["heat_iterations"] > 0 for h in sim_tm.history)
    eq = sim_tm.eq_mom
    assert np.isfinite(eq.X.numpy()).all() and np.abs(eq.X.numpy()).max() > 0
    T = sim_tm.eq_heat.T.x.array
    assert 280.0 < T.min() and T.max() < 340.0          # geothermal profile, gas at 293 K on the cavern wall
    for stage, field in (("equilibrium", "u"), ("operation", "u"), ("operation", "T"), ("operation", "q_elems")):
        assert os.path.isfile(work / "output" / "case_1" / stage / field / f"{field}.xdmf")


def test_reference_thermal_cavern_example_runs_unmodified(tmp_path, monkeypatch, capsys):
    """examples/thermal/2_cavern/main.py (heat diffusion around a cavern on grids/cavern_regular: geothermal initial field,
    Dirichlet / Neumann / Robin conditions, parabolic time controller, Simulator_T) executed unmodified, two steps."""
    import safeincave_b200 as sf
    from safeincave_b200 import compat
    from tests.hostemu import EmuEngine
    example = os.path.join(REF, "examples", "thermal", "2_cavern", "main.py")
    if not os.path.isfile(example):
        pytest.skip("example not in the reference checkout")
    monkeypatch.setattr(sf.HeatDiffusion, "engine_cls", EmuEngine)
    registered = compat.install()
    sims = []
    try:
        work = tmp_path / "examples" / "thermal" / "2_cavern"
        work.mkdir(parents=True)
        os.symlink(os.path.join(REF, "grids"), tmp_path / "grids")
        monkeypatch.chdir(work)
        real_tc, real_tcp = sf.TimeController, sf.TimeControllerParabolic

        class TC(real_tc):
            def __init__(self, dt, initial_time, final_time, time_unit="second"):
                super().__init__(dt=dt, initial_time=initial_time, final_time=min(final_time, 2 * dt), time_unit=time_unit)

        class TCP(real_tcp):
            def keep_looping(self):
                return self.step_counter < 2 and super().keep_looping()

        class Rec(sf.Simulator_T):
            def __init__(self, *a, **k):
                super().__init__(*a, **k)
                sims.append(self)

        monkeypatch.setattr(sf, "TimeController", TC)
        monkeypatch.setattr(sf, "TimeControllerParabolic", TCP)
        monkeypatch.setattr(sf, "Simulator_T", Rec)
        runpy.run_path(example, run_name="not_main")["main"]()
    finally:
        for name in registered:
            for key in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
                del sys.modules[key]
    (sim,) = sims
    assert len(sim.history) == 2 and all(h["heat_iterations"] > 0 for h in sim.history)
    T = sim.eq_heat.T.x.array
    assert np.isfinite(T).all() and 270.0 < T.min() and T.max() < 340.0
    assert os.path.isfile(work / "output" / "case_0" / "T" / "T.xdmf") and os.path.isfile(work / "output" / "case_0" / "log.txt")
    assert "Total time:" in capsys.readouterr().out
