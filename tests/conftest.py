import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    """The oracle's C shim must exist (built by `make -C oracle` / __graft_entry__.build())."""
    so = os.path.join(ROOT, "oracle", "_build", "libsicmath.so")
    if not os.path.isfile(so):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    yield


# GPU tests of code that has so far only been through the host emulation (tests/hostemu) -- the round's GPU budget
# was spent before it was written.  They run LAST under `-m gpu -x`, so that a surprise in them cannot hide the
# B200-verified parity tests behind an early stop.  Remove a name once it has passed on a B200.
_FIRST_TIME_ON_GPU = (
    "test_cuda_vs_oracle_all_quantities_extended_elements", "test_cuda_vs_reference_goldens_extended_elements",
    "test_pq_output_fields", "test_time_steps_triaxial_cube_munson_dawson",
    "test_time_steps_triaxial_cube_mohr_coulomb_and_matsuoka_nakai", "test_dt_retry_and_restore_follow_the_reference",
    "test_heat_steps_cube", "test_heat_steps_cavern_regular", "test_thermomechanical_steps_cube",
    "test_thermomechanical_step_cavern_regular", "test_time_steps_cavern_regular_extrapolated_guess",
    "test_guess_extrapolation_kernel", "test_lagged_multigrid_setup",
    "test_operator_rhs_blocks_strain_configs2_grid", "test_fused_coarse_level_sweep_matches_the_oracle",
)


def pytest_collection_modifyitems(config, items):
    def late(item):
        return any(item.name == n or item.name.startswith(n + "[") for n in _FIRST_TIME_ON_GPU)
    items[:] = [i for i in items if not late(i)] + [i for i in items if late(i)]
