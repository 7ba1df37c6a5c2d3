"""tests/test_gpu_constitutive.py executed on the HOST through tests/hostemu: the product's constitutive
kernels (constitutive.cu, compiled by g++ against the CUDA emulation layer) against the reference's
golden vectors and against the oracle, phase by phase, to the same tolerances as on the B200."""
import pytest

from tests import test_gpu_constitutive as G


@pytest.fixture(scope="module")
def adapter():
    from tests.gpu_adapter import GpuMaterial
    from tests.hostemu import EmuEngine

    class EmuMaterial(GpuMaterial):
        engine_cls = EmuEngine
    return EmuMaterial


test_cuda_vs_reference_goldens = G.test_cuda_vs_reference_goldens
test_cuda_vs_oracle_all_quantities = G.test_cuda_vs_oracle_all_quantities
test_float32_params_match_reference_stiffness = G.test_float32_params_match_reference_stiffness
test_singular_tangent_falls_back_to_elastic = G.test_singular_tangent_falls_back_to_elastic
test_cuda_vs_oracle_all_quantities_extended_elements = G.test_cuda_vs_oracle_all_quantities_extended_elements
test_cuda_vs_reference_goldens_extended_elements = G.test_cuda_vs_reference_goldens_extended_elements
