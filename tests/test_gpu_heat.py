"""B200 parity tests of the heat equation (csrc/heat.cu) and of Simulator_TM, against oracle/heat.py."""
import pytest

from tests import heat_checks as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sf():
    import safeincave_b200 as sf
    return sf


def test_heat_steps_cube(sf):
    assert C.check_heat_steps(sf, "cube_coarse", 1, 6, 0.5 * C.DAY) < 100


def test_heat_steps_cavern_regular(sf):
    C.check_heat_steps(sf, "cavern_regular", 0, 3, 20 * C.DAY)


def test_thermomechanical_steps_cube(sf):
    C.check_thermomechanical_steps(sf)


def test_thermomechanical_step_cavern_regular(sf):
    C.check_thermomechanical_steps(sf, "cavern_regular", 0, 1, 2 * C.DAY)
