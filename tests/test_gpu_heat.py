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


def test_thermomechanical_steps_config4_cavern_overburden_coarse(sf):
    """BASELINE configs[3] on the grid it names (25 608 cells, salt + overburden with their own density, stiffness,
    viscosity, creep and thermal expansion; set-up of examples/thermomechanics/2_cavern/main.py, cases.overburden_tm_case):
    two Simulator_TM steps against OracleSimulatorTM -- same Newton history, u / sigma / creep strains <= 1e-8."""
    C.check_thermomechanical_steps(sf, "cavern_overburden_coarse", 0, 2, 0.5 * C.DAY, tol_T=2e-9)
