"""Transolver_block: the product's fused block (csrc/transolver.cu + library GEMMs) against the
oracle's restatement of the reference (oracle/fvgn_oracle.py: transolver_block, GraphTransolver.py:48-169) on ragged
batches -- forward, input gradient and every parameter gradient.  CPU: through the SIMT emulator (test infrastructure);
-m gpu: through libfvgn_b200.so.  Whole-model parity against the reference's golden vectors is in
tests/test_emu_parity.py / tests/test_gpu_parity.py (TransFVGN_v1 / _v2 cases)."""
import pytest
import torch

from oracle import fvgn_oracle as O
from tests import product_util as PU


def _block_and_inputs(sizes, device, seed=0, dtype=torch.float32):
    from gen_fvgn_steady_b200.FVMmodel.Models.GraphTransolver.GraphTransolver import Transolver_block
    g = torch.Generator().manual_seed(seed)
    blk = Transolver_block(num_heads=8, hidden_dim=128, dropout=0, act="gelu", mlp_ratio=2, slice_num=32)
    with torch.no_grad():
        for name, p in blk.named_parameters():
            if name.endswith("graph_temperature"):
                p.copy_(0.3 + 0.5 * torch.rand(p.shape, generator=g))
            elif p.dim() >= 2:
                p.copy_(torch.randn(p.shape, generator=g) / (p.shape[-1] ** 0.5))
            elif "ln_" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
    n = sum(sizes)
    x = torch.randn(n, 128, generator=g)
    batch = torch.cat([torch.full((c,), b, dtype=torch.int64) for b, c in enumerate(sizes)])
    cot = torch.randn(n, 128, generator=g)
    return blk.to(device), x.to(device), batch.to(device), cot.to(device)


def _oracle(blk, x, batch, cot):
    sd = {k: v.detach().double().cpu().requires_grad_(True) for k, v in blk.state_dict().items()}
    xd = x.detach().double().cpu().requires_grad_(True)
    out = O.transolver_block(sd, "", xd, batch.cpu())
    out.backward(cot.double().cpu())
    return out.detach(), xd.grad, {k: v.grad for k, v in sd.items()}


def _check(blk, x, batch, cot, tol):
    xr = x.clone().requires_grad_(True)
    out = blk(xr, batch)
    out.backward(cot)
    ref_out, ref_dx, ref_g = _oracle(blk, x, batch, cot)

    def rel(a, b):
        return float((a.detach().double().cpu() - b).norm() / b.norm().clamp_min(1e-30))

    errs = {"out": rel(out, ref_out), "d_x": rel(xr.grad, ref_dx)}
    for k, p in blk.named_parameters():
        if ref_g[k] is None:             # Attn.temperature and ln_1 are not on the in_layernorm=False path
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        errs[k] = rel(p.grad, ref_g[k])
    bad = {k: v for k, v in errs.items() if v > tol}
    assert not bad, bad
    return errs


@pytest.mark.parametrize("sizes", [(70,), (33, 1, 95), (64, 32)])
def test_transolver_block_emulated(sizes):
    PU.use_emulated_kernels()
    try:
        blk, x, batch, cot = _block_and_inputs(sizes, "cpu")
        _check(blk, x, batch, cot, tol=2e-5)
    finally:
        PU.use_real_kernels()


def test_public_graph_forward_and_mlp_emulated():
    """The stand-alone public methods keep the reference's meaning (graph_forward includes the to_out bias)."""
    import torch.nn.functional as F
    PU.use_emulated_kernels()
    try:
        blk, x, batch, _ = _block_and_inputs((40, 25), "cpu", seed=1)
        got = blk.Attn.graph_forward(x, batch)
        sd = {k: v.detach().double() for k, v in blk.state_dict().items()}
        # oracle block minus its second half: rebuild the attention output from the block identity
        full = O.transolver_block(sd, "", x.double(), batch)
        y = got.double() + x.double()
        h = F.layer_norm(y, (128,), sd["ln_2.weight"], sd["ln_2.bias"], 1e-5)
        h = O.gelu(F.linear(h, sd["mlp.linear_pre.0.weight"], sd["mlp.linear_pre.0.bias"]))
        want_full = F.linear(h, sd["mlp.linear_post.weight"], sd["mlp.linear_post.bias"]) + y
        assert float((want_full - full).norm() / full.norm()) < 2e-5
        z = torch.randn(17, 128)
        m = blk.mlp(z).double()
        mref = F.linear(O.gelu(F.linear(z.double(), sd["mlp.linear_pre.0.weight"], sd["mlp.linear_pre.0.bias"])),
                        sd["mlp.linear_post.weight"], sd["mlp.linear_post.bias"])
        assert float((m - mref).norm() / mref.norm()) < 2e-5
    finally:
        PU.use_real_kernels()


def _inlayernorm_wiring(device):
    """in_layernorm=True (original-Transolver variant) and the `embedding` argument against a composition of the
    separately verified public pieces."""
    import torch.nn.functional as F
    blk, x, batch, cot = _block_and_inputs((45, 30), device, seed=2)
    emb = torch.randn_like(x)
    xr, er = x.clone().requires_grad_(True), emb.clone().requires_grad_(True)
    out = blk(xr, batch, in_layernorm=True, embedding=er)
    out.backward(cot)
    got = [out.detach(), xr.grad.clone(), er.grad.clone()] + [p.grad.clone() for p in blk.parameters() if p.grad is not None]
    for p in blk.parameters():
        p.grad = None
    x2, e2 = x.clone().requires_grad_(True), emb.clone().requires_grad_(True)
    fx = x2 + e2
    y = blk.Attn.graph_forward(blk.ln_1(fx), batch) + fx
    ref = blk.mlp(F.layer_norm(y, (128,), blk.ln_2.weight, blk.ln_2.bias, 1e-5)) + y
    ref.backward(cot)
    want = [ref.detach(), x2.grad, e2.grad] + [p.grad for p in blk.parameters() if p.grad is not None]
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert float((a - b).norm() / b.norm().clamp_min(1e-30)) < 2e-5
    # embedding on the fused (in_layernorm=False) path == adding it outside
    for p in blk.parameters():
        p.grad = None
    x3, e3 = x.clone().requires_grad_(True), emb.clone().requires_grad_(True)
    o1 = blk(x3, batch, embedding=e3)
    o1.backward(cot)
    o2 = blk((x + emb).clone(), batch)
    assert float((o1 - o2).norm() / o2.norm()) < 1e-6
    assert torch.equal(x3.grad, e3.grad)


def test_inlayernorm_and_embedding_emulated():
    PU.use_emulated_kernels()
    try:
        _inlayernorm_wiring("cpu")
    finally:
        PU.use_real_kernels()


@pytest.mark.gpu
def test_inlayernorm_and_embedding_gpu():
    _inlayernorm_wiring("cuda")


@pytest.mark.gpu
@pytest.mark.parametrize("sizes", [(70,), (33, 1, 95), (5000, 12345, 777)])
def test_transolver_block_gpu(sizes):
    blk, x, batch, cot = _block_and_inputs(sizes, "cuda")
    errs = _check(blk, x, batch, cot, tol=2e-5)
    print(errs)


@pytest.mark.gpu
def test_transolver_block_gpu_deterministic_and_shadow():
    blk, x, batch, cot = _block_and_inputs((3000, 2000), "cuda", seed=3)
    blk.precision = "bf16"
    outs = []
    for _ in range(2):
        xr = x.clone().requires_grad_(True)
        for p in blk.parameters():
            p.grad = None
        out = blk(xr, batch)
        out.backward(cot)
        outs.append((out.detach().clone(), xr.grad.clone(), [p.grad.clone() for p in blk.parameters() if p.grad is not None]))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert all(torch.equal(a, b) for a, b in zip(outs[0][2], outs[1][2]))
    master, sh = blk.last_shadow
    assert master is out and torch.equal(sh, out.detach().to(torch.bfloat16))


@pytest.mark.gpu
def test_transolver_properties_at_scale():
    """Size-independent properties at BASELINE's mesh scale (1 M node rows, 3 ragged graphs): slice weights are
    distributions, the deterministic token sums and the de-slice equal their dense fp64 definitions
    (GraphTransolver.py:62-90), and the whole backward is exactly linear in the cotangent (x2 is exact in fp32)."""
    from gen_fvgn_steady_b200 import ops
    sizes = (400_000, 250_001, 349_999)
    blk, x, batch, cot = _block_and_inputs(sizes, "cuda", seed=5)
    A = blk.Attn
    tsp = ops.TsPlan.of(batch, None)
    with torch.no_grad():
        a, saved = ops._attn_forward(x, A.in_project_fx.weight, A.in_project_fx.bias, A.in_project_x.weight, A.in_project_x.bias,
                                     A.in_project_slice.weight, A.in_project_slice.bias, A.graph_temperature, A.to_q.weight,
                                     A.to_k.weight, A.to_v.weight, A.to_out[0].weight, A.scale, tsp, None)
        _, P, sw, rec, tok_out, out_x = saved[:6]
        n = x.shape[0]
        s3 = sw.view(n, 8, 32)
        assert float((s3.sum(-1) - 1).abs().max()) < 1e-5 and float(s3.min()) >= 0.0
        fx = P[:, :128].view(n, 8, 16).double()
        for b, (lo, hi) in enumerate(zip([0, sizes[0], sizes[0] + sizes[1]], [sizes[0], sizes[0] + sizes[1], n])):
            sb = s3[lo:hi].double()
            num = torch.einsum("nhg,nhd->hgd", sb, fx[lo:hi]).reshape(-1)
            nrm = sb.sum(0).reshape(-1)
            got = rec[b].double()
            assert float((got[:4096] - num).norm() / num.norm()) < 1e-5
            assert float((got[4096:] - nrm).norm() / nrm.norm()) < 1e-5
            want = torch.einsum("nhg,hgd->nhd", sb, tok_out[b].view(8, 32, 16).double()).reshape(hi - lo, 128)
            assert float((out_x[lo:hi].double() - want).norm() / want.norm()) < 1e-5
    grads = []
    for scale in (1.0, 2.0):
        xr = x.clone().requires_grad_(True)
        for p in blk.parameters():
            p.grad = None
        blk(xr, batch).backward(cot * scale)
        grads.append([xr.grad.clone()] + [p.grad.clone() for p in blk.parameters() if p.grad is not None])
    for g1, g2 in zip(*grads):
        assert torch.equal(g1 * 2.0, g2)


@pytest.mark.parametrize("sizes", [(1,), (31, 33, 1), (5000, 0, 123), (300_000, 17)])
def test_ts_plan_chunk_table(sizes):
    """Host logic of the chunk table: chunks tile the rows in order, never straddle a graph (empty graphs get no chunk),
    chunk_ptr groups them per graph; the batch vector must be sorted.  The table has n_chunks >= the real chunk count slots
    (an upper bound: no device round trip at plan time), the surplus slots are empty (their CTAs write zero partials)."""
    from gen_fvgn_steady_b200 import ops
    batch = torch.cat([torch.full((c,), b, dtype=torch.int64) for b, c in enumerate(sizes)])
    tsp = ops.TsPlan(batch)
    ptr = tsp.chunk_ptr.tolist()
    real = ptr[-1]   # n_chunks is an upper bound computed without a device round trip; the slots past `real` are empty
    assert tsp.nseg == len(sizes) and tsp.nb == len(sizes) and len(ptr) == len(sizes) + 1 and real <= tsp.n_chunks
    assert tsp.chunks[real:tsp.n_chunks].abs().sum() == 0
    ch = tsp.chunks[:real].tolist()
    pos = 0
    for seg, r0, r1 in ch:
        assert r0 == pos and r1 > r0 and (r1 - r0) <= 4096
        assert int(batch[r0]) == seg and int(batch[r1 - 1]) == seg
        pos = r1
    assert pos == sum(sizes)
    for b, c in enumerate(sizes):
        rows = sum(r1 - r0 for seg, r0, r1 in ch[ptr[b]:ptr[b + 1]])
        assert rows == c and all(seg == b for seg, _, _ in ch[ptr[b]:ptr[b + 1]])
    assert tsp.all_ptr.tolist() == [0, tsp.n_chunks]
    if len(sizes) > 1 and sizes[0] > 0 and sizes[1] > 0:
        with pytest.raises(RuntimeError):
            ops.TsPlan(torch.flip(batch, [0]))


@pytest.mark.parametrize("nb", [1, 3])
def test_token_attention_kernels_match_autograd(nb):
    """fvgn_ts_token_attention_forward / _backward (GraphTransolver.py:72-81: tok = num / (norm + 1e-5), q k v projections
    shared by the heads, softmax(q k^T scale) v) through the CPU emulator against the same expression differentiated by
    torch autograd in fp64."""
    from tests import product_util as PU
    from gen_fvgn_steady_b200 import ops
    PU.use_emulated_kernels()
    try:
        g = torch.Generator().manual_seed(3)
        num = torch.randn(nb, 4096, generator=g)
        norm = torch.rand(nb, 256, generator=g) * 5 + 0.1
        rec = torch.cat([num, norm], 1)
        ws = [torch.randn(16, 16, generator=g) / 4 for _ in range(3)]
        cot = torch.randn(nb, 4096, generator=g)
        scale = 16 ** -0.5
        out = ops._token_attention(rec, *ws, scale)
        d_rec, d_wq, d_wk, d_wv = ops._token_attention_backward(rec, *ws, scale, cot)
        leaves = [t.double().requires_grad_() for t in (rec, *ws)]
        r, wq, wk, wv = leaves
        tok = r[:, :4096].reshape(nb, 8, 32, 16) / (r[:, 4096:].reshape(nb, 8, 32, 1) + 1e-5)
        q, k, v = tok @ wq.t(), tok @ wk.t(), tok @ wv.t()
        ref = (torch.softmax(q @ k.transpose(-1, -2) * scale, -1) @ v).reshape(nb, 4096)
        grads = torch.autograd.grad(ref, leaves, cot.double())
    finally:
        PU.use_real_kernels()
    rel = lambda a, b: float((a.double() - b).norm() / b.norm())
    assert rel(out, ref) < 1e-6
    for a, b in zip((d_rec, d_wq, d_wk, d_wv), grads):
        assert rel(a, b) < 1e-5, rel(a, b)
