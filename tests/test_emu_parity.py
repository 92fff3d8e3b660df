"""Container-side (no GPU) check of the PRODUCT's host logic + kernel bodies: the SIMT kernels are compiled for the
host over tests/emu/cuda_emu.h and the whole NNmodel forward/backward is compared with the reference's golden vectors.
The real parity tests are tests/test_gpu_parity.py (-m gpu, through libfvgn_b200.so on a B200)."""
import pytest

from tests import golden_util as GU
from tests import product_util as PU


@pytest.fixture(autouse=True)
def _emu():
    PU.use_emulated_kernels()
    yield
    PU.use_real_kernels()


@pytest.mark.parametrize("name", ["synth_ns_batch2_v1", "synth_ns_batch2_v2", "synth_ns_batch2_v1_nc"])
def test_emulated_product_matches_reference(name):
    model, out, loss, z = PU.run_product(name, "cpu")
    rep = {}
    PU.compare_with_golden(model, out, loss, z, "f64", tol=2e-4, gtol=2e-3, report=rep)
    print(rep)
