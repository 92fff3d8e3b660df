"""-m gpu: fvgn_gemm_tf32 (csrc/gemm_tf32.cu: tcgen05 kind::tf32, the Transolver block's dense projections and their autograd)
against fp64 torch, for every shape the block uses, ragged row counts, bias / addend epilogues, and determinism of the
weight-gradient reduction."""
import pytest
import torch

from tests import product_util as PU

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128), (256, 128), (128, 256)]   # (O, I) of in_project (cat), to_out, linear_pre, linear_post


def _rel(a, b):
    return float((a.double() - b).norm() / b.norm())


@pytest.mark.parametrize("rows", [1, 127, 128, 129, 1000, 148 * 128 + 5])
@pytest.mark.parametrize("O,I", SHAPES)
def test_linear_forward_dgrad_wgrad(O, I, rows):
    from gen_fvgn_steady_b200 import ops
    PU.use_real_kernels()
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(O + I + rows)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    x, w, b = rn(rows, I), rn(O, I) / I ** 0.5, rn(O)
    addend_y, addend_x, dy = rn(rows, O), rn(rows, I), rn(rows, O)
    # forward: y = x w^T + b (+ addend)
    y = ops.linear_fwd(x, w, b, tc=True)
    assert _rel(y, x.double() @ w.double().t() + b.double()) < 2e-3
    y2 = ops.linear_fwd(x, w, None, addend_y, tc=True)
    assert _rel(y2, x.double() @ w.double().t() + addend_y.double()) < 2e-3
    # dgrad: dx = dy w (+ addend)
    dx = ops.linear_dgrad(dy, w, tc=True)
    assert _rel(dx, dy.double() @ w.double()) < 2e-3
    dx2 = ops.linear_dgrad(dy, w, addend_x, tc=True)
    assert _rel(dx2, dy.double() @ w.double() + addend_x.double()) < 2e-3
    # wgrad: dW = dy^T x, deterministic
    dw = ops.linear_wgrad(dy, x, tc=True)
    assert tuple(dw.shape) == (O, I)
    assert _rel(dw, dy.double().t() @ x.double()) < 2e-3
    assert torch.equal(dw, ops.linear_wgrad(dy, x, tc=True))


def test_transolver_block_uses_the_tf32_gemms_in_tensor_core_modes():
    """In f16 / bf16 mode no library GEMM is left in the Transolver block's forward + backward except the [B,8,32,16]-sized
    token attention: the four projection GEMMs (in_project_fx | in_project_x as one, to_out, linear_pre, linear_post) and
    their eight gradient GEMMs are fvgn_gemm_tf32 calls."""
    from gen_fvgn_steady_b200 import _lib
    from gen_fvgn_steady_b200.FVMmodel.importer import NNmodel
    from gen_fvgn_steady_b200.utils.get_param import params as default_params
    from gen_fvgn_steady_b200.mesh import synthetic as S
    from tests.case_inputs import product_graphs
    PU.use_real_kernels()
    mesh, uvp = S.make_case(12, kind="mixed", bc="channel", seed=2)
    p = default_params(net="TransFVGN_v1", message_passing_num=1, dataset_size=1, precision="f16")
    torch.manual_seed(0)
    model = NNmodel(p).cuda()
    calls = []
    orig = _lib.call

    def spy(name, *a):
        calls.append(name)
        return orig(name, *a)
    _lib.call = spy
    try:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            out = model(*product_graphs([mesh], [uvp], "cuda"), is_training=True)
            PU.script_loss(out, p).backward()
            torch.cuda.synchronize()
    finally:
        _lib.call = orig
    assert calls.count("fvgn_gemm_tf32") == 4 + 8, calls.count("fvgn_gemm_tf32")
    names = [e.key for e in prof.key_averages()]
    assert any("gemm_tf32_kernel" in k for k in names)
    big = [k for k in names if ("cutlass" in k or "gemm" in k.lower() or "cublas" in k.lower()) and "gemm_tf32_kernel" not in k
           and "gemm_partial_reduce" not in k]
    # what is left are the token-attention products on [1,8,32,16] tensors (bmm / small gemm kernels), never an [N, .] GEMM:
    # every library GEMM kernel of the step runs for less than the smallest of ours
    ours = min(e.device_time_total / max(e.count, 1) for e in prof.key_averages() if "gemm_tf32_kernel" in e.key)
    for e in prof.key_averages():
        if e.key in big:
            assert e.device_time_total / max(e.count, 1) <= max(ours, 20.0), (e.key, e.device_time_total, ours)
