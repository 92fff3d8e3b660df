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


@pytest.mark.parametrize("name", ["synth_ns_batch2_v1", "synth_ns_batch2_v2", "synth_ns_batch2_v1_nc", "synth_ns_batch2_v1_explicit",
                                  "synth_ns_batch2_v1_implicit"])
def test_emulated_product_matches_reference(name):
    model, out, loss, z = PU.run_product(name, "cpu")
    rep = {}
    PU.compare_with_golden(model, out, loss, z, "f64", tol=2e-4, gtol=2e-3, report=rep)
    print(rep)


@pytest.mark.parametrize("mode_name,k1", [("FVGN_MLP_ENC_NODE", 12), ("FVGN_MLP_ENC_EDGE", 15)])
def test_encoder_layer1_never_reads_past_a_weight_row(mode_name, k1):
    """Regression: the encoders' first layer runs on an input tile zero-padded from 12 / 15 to 16 columns; the weight tile
    loader must not read the pad columns from memory (for the last output row that is past the end of the tensor, and
    0 x NaN-bits = NaN).  The weight is placed at the end of a NaN-filled buffer."""
    import torch
    from gen_fvgn_steady_b200 import _lib, ops
    g = torch.Generator().manual_seed(0)
    n = 70
    buf = torch.full((128 * k1 + 64,), float("nan"))
    buf[:128 * k1] = torch.randn(128 * k1, generator=g) / k1 ** 0.5
    w1 = buf[:128 * k1].view(128, k1)
    params = [w1, 0.1 * torch.randn(128, generator=g), torch.randn(128, 128, generator=g) / 128 ** 0.5,
              0.1 * torch.randn(128, generator=g), torch.randn(128, 128, generator=g) / 128 ** 0.5,
              0.1 * torch.randn(128, generator=g), torch.ones(128), torch.zeros(128)]
    x = torch.randn(n, 12, generator=g)
    mode = getattr(_lib, mode_name)
    if k1 == 12:
        out, _ = ops.mlp_forward(mode, "fp32", n, params, x)
    else:
        pos = torch.randn(n, 2, generator=g)
        e = 150
        s = torch.randint(0, n, (e,), generator=g).to(torch.int32)
        r = torch.randint(0, n, (e,), generator=g).to(torch.int32)
        out, _ = ops.mlp_forward(mode, "fp32", e, params, x, pos, s, r)
    assert not bool(torch.isnan(out).any())
