"""Heat equation (csrc/heat.cu) and Simulator_TM on the HOST through tests/hostemu, against oracle/heat.py."""
import pytest

from tests import heat_checks as C


@pytest.fixture(scope="module")
def sf():
    import safeincave_b200 as sf
    from tests.hostemu import EmuEngine
    old = (sf.LinearMomentum.engine_cls, sf.HeatDiffusion.engine_cls)
    sf.LinearMomentum.engine_cls = EmuEngine
    sf.HeatDiffusion.engine_cls = EmuEngine
    yield sf
    sf.LinearMomentum.engine_cls, sf.HeatDiffusion.engine_cls = old


def test_heat_steps_cube(sf):
    assert C.check_heat_steps(sf, "cube_coarse", 1, 6, 0.5 * C.DAY) < 100


def test_heat_steps_cavern_regular(sf):
    C.check_heat_steps(sf, "cavern_regular", 0, 3, 20 * C.DAY)


def test_thermomechanical_steps_cube(sf):
    C.check_thermomechanical_steps(sf)


@pytest.mark.skipif(not __import__("os").environ.get("SIC_SLOW"), reason="~13 min under emulation (25 608 cells); set SIC_SLOW=1")
def test_thermomechanical_steps_config4_cavern_overburden_coarse(sf):
    """BASELINE configs[3] on its own two-region grid (tests/test_gpu_heat.py runs it on the B200)."""
    C.check_thermomechanical_steps(sf, "cavern_overburden_coarse", 0, 2, 0.5 * C.DAY, tol_T=2e-9)
