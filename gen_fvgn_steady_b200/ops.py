"""torch.autograd.Function wrappers over the C-ABI kernels, each with a hand-written backward.

No torch_scatter / PyG / Triton / torch.compile anywhere; PyTorch only owns memory, streams and the
autograd graph between these ops.
"""
import ctypes
import os
import weakref

import torch

from . import _lib
from ._lib import fptr, iptr

PREC = {"fp32": _lib.FVGN_PREC_FP32, "bf16": _lib.FVGN_PREC_BF16, "f16": _lib.FVGN_PREC_F16}
# tensor-core (tcgen05) modes and the torch dtype of their 16-bit operand streams:
#   bf16: bfloat16 operands (8-bit significand);
#   f16 : IEEE-half operands -- the 11-bit significand of TF32, the arithmetic the reference's GPU path runs its Linear
#         layers in (src/pre_train_Adam.py:29) -- with the backward's gradient operands pre-scaled by a power of two
#         (GradScaleFn below) so that they stay inside half's exponent range.
HDTYPE = {"bf16": torch.bfloat16, "f16": torch.float16}


def is_tc(precision):
    return precision in HDTYPE


# tensor-core modes: the agg[senders] | agg[receivers] columns of the edge MLP's first layer are differentiated per NODE
# (fvgn_mlp_desc.d_aggh, csrc/mlp_tc_bwd_node.cu) instead of per edge: no [E,256] gradient stream and no incidence
# reduction of it.  FVGN_NODE_LEVEL_LAYER1 = 2 (default): incidence sums of dZ1 by a many-CTA kernel into operand tile images
# + a node GEMM kernel fed by bulk copies; 1: one fused kernel (its gather runs one CTA per SM: slower); 0: edge-level path.
# Measured at 4 M cells on a B200: 145.8 ms per step (2) against 148.4-149.7 (0) and 150.0 (1); DESIGN.md section 4.2.
NODE_LEVEL_LAYER1 = int(os.environ.get("FVGN_NODE_LEVEL_LAYER1", "2") or 0)


# tensor-core modes: 16-bit latent streams between the GnBlocks of a model (GnBlockFn, GN_LATENTS16); FVGN_LATENTS16=0 keeps
# the fp32 residual streams x / e in HBM (rounded to 16 bit only as MLP operands).
LATENTS16 = os.environ.get("FVGN_LATENTS16", "1") != "0"


def default_precision():
    return os.environ.get("FVGN_PRECISION", "fp32")


def _empty(shape, like):
    return torch.empty(shape, dtype=torch.float32, device=like.device)


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------ raw kernel calls
BF16 = torch.bfloat16
_TCODE = {torch.float32: _lib.FVGN_T_F32, torch.bfloat16: _lib.FVGN_T_BF16, torch.float16: _lib.FVGN_T_F16}


def hptr(t, allow_none=False):
    if t is not None and t.dtype not in (torch.bfloat16, torch.float16):
        raise RuntimeError(f"fvgn_b200: expected a 16-bit operand stream, got {t.dtype}")
    return _lib.ptr(t, None, allow_none)


def _tcode(t):
    return _TCODE[t.dtype]


def adj_reduce(src, plan, width, flags=0, out=None, out_dtype=torch.float32):
    """out[i] = sum_{j in Adj(i)} src[j]  (blocks.py:92-99 / :44-51).  src / out may be fp32 or bf16 (fp32 accumulation)."""
    if out is None:
        out = torch.empty((plan.N, width), dtype=out_dtype, device=src.device)
    if src.dtype == torch.float32 and out.dtype == torch.float32:
        _lib.call("fvgn_adj_reduce", fptr(src), iptr(plan.inc_ptr), iptr(plan.inc_nbr), fptr(out), plan.N, width, flags,
                  _lib.stream_ptr(src.device))
    else:
        _lib.call("fvgn_adj_reduce_t", _lib.ptr(src), _tcode(src), iptr(plan.inc_ptr), iptr(plan.inc_nbr), _lib.ptr(out),
                  _tcode(out), plan.N, width, flags, _lib.stream_ptr(src.device))
    return out


def inc_reduce(src, plan, width, out_dtype=torch.float32):
    """out[i] = sum over incident (edge, role) of src[edge, role*width:(role+1)*width]  (blocks.py:24-42)."""
    out = torch.empty((plan.N, width), dtype=out_dtype, device=src.device)
    if src.dtype == torch.float32 and out.dtype == torch.float32:
        _lib.call("fvgn_inc_reduce", fptr(src), iptr(plan.inc_ptr), iptr(plan.inc_code), fptr(out), plan.N, width,
                  _lib.stream_ptr(src.device))
    else:
        _lib.call("fvgn_inc_reduce_t", _lib.ptr(src), _tcode(src), iptr(plan.inc_ptr), iptr(plan.inc_code), _lib.ptr(out),
                  _tcode(out), plan.N, width, _lib.stream_ptr(src.device))
    return out


def shadow(t, cached=None, dtype=BF16):
    """16-bit row-major shadow of an fp32 latent; `cached` = (master, shadow) as attached by the producing kernel."""
    if cached is not None and cached[0] is t and cached[1] is not None and cached[1].dtype == dtype:
        return cached[1]
    if dtype == torch.float16:
        return t.detach().clamp(-65504.0, 65504.0).to(dtype).contiguous()   # saturate like the kernels' conversions
    return t.detach().to(dtype).contiguous()


_MLP_K1 = {_lib.FVGN_MLP_EDGE: 384, _lib.FVGN_MLP_NODE: 192, _lib.FVGN_MLP_ENC_NODE: 12, _lib.FVGN_MLP_ENC_EDGE: 15,
           _lib.FVGN_MLP_DEC: 128}


class PackedWeights:
    """16-bit UMMA operand images of the MLP weights, rebuilt by EVERY forward (one ~3 us kernel per fused MLP; its
    backward reuses the forward's image through ctx.pk).  Nothing is cached across forwards on purpose: Tensor._version
    does not move under the in-place updates that matter most -- torch.optim.Adam(fused=True) leaves it untouched, and so
    do CUDA-graph replays of an optimizer step or `p.data` surgery -- so a cache keyed on it served stale weights."""

    packs = 0   # pack launches so far (tests)

    @classmethod
    def get(cls, mode, params, precision="bf16"):
        w1, w2, w3 = params[0], params[2], params[4]
        nbytes = int(_lib.load().fvgn_mlp_packed_bytes(mode))
        # always a fresh buffer: an autograd node of an earlier forward may still hold the previous image for its backward
        buf = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=w1.device)
        _lib.call("fvgn_mlp_pack_weights", mode, PREC[precision], fptr(_c(w1.detach())), fptr(_c(w2.detach())),
                  fptr(_c(w3.detach())), _lib.ptr(buf), _lib.stream_ptr(w1.device))
        cls.packs += 1
        return buf


def _img_buffer(nbytes, device):
    """1024-B aligned device scratch for bf16 tile images -> (owner tensor, aligned pointer)."""
    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    return buf, (buf.data_ptr() + 1023) // 1024 * 1024


class Z1Image:
    """16-bit tile images of the first pre-activation (written by the tensor-core forward, consumed by its backward)."""

    made = 0   # instances so far (tests: a rollout forward makes none, a training step exactly one per fused MLP)

    def __init__(self, mode, rows, device):
        Z1Image.made += 1
        nbytes = int(_lib.load().fvgn_mlp_bwd_workspace_bytes(mode, PREC["bf16"], rows))
        self.buf, self.ptr = _img_buffer(nbytes, device)


_SHADOW_MODES = (_lib.FVGN_MLP_EDGE, _lib.FVGN_MLP_NODE, _lib.FVGN_MLP_DEC)


def _mlp_desc(mode, precision, rows, params, in0, in1=None, idx_s=None, idx_r=None, flags=0, packed=None, in0h=None,
              in1h=None):
    d = _lib.MlpDesc()
    d.mode, d.precision, d.rows, d.flags = mode, PREC[precision], rows, flags
    d.in0, d.in1 = fptr(in0, True), fptr(in1, True)
    d.idx_s, d.idx_r = iptr(idx_s, True), iptr(idx_r, True)
    ps = [_c(p.detach()) for p in params]
    d._keep = ps  # keep contiguous copies alive for the duration of the call
    d.w1, d.b1, d.w2, d.b2, d.w3, d.b3 = (fptr(p) for p in ps[:6])
    if len(ps) == 8:
        d.ln_g, d.ln_b = fptr(ps[6]), fptr(ps[7])
    if is_tc(precision):
        d._packed = packed if packed is not None else PackedWeights.get(mode, params, precision)
        d.w_bf16 = _lib.ptr(d._packed)
        if mode in _SHADOW_MODES:
            # layer-1 operands are read from 16-bit shadows; make them here when the caller has none (stand-alone use)
            if in0h is None:
                in0h = shadow(in0, dtype=HDTYPE[precision])
            if in1h is None and in1 is not None:
                in1h = shadow(in1, dtype=HDTYPE[precision])
            d._h = (in0h, in1h)
            d.in0h, d.in1h = hptr(in0h), hptr(in1h, True)
    return d


def mlp_forward(mode, precision, rows, params, in0, in1=None, idx_s=None, idx_r=None, want_out=True, want_res=False,
                flags=0, packed=None, z1=None, in0h=None, in1h=None, want_outh=False, want_resh=False):
    """-> (out, out_res) [, outh, out_resh when requested: the bf16 shadows written by the bf16 kernels].
    z1: a Z1Image to fill (bf16 mode, needed by mlp_backward) or None (inference)."""
    d = _mlp_desc(mode, precision, rows, params, in0, in1, idx_s, idx_r, flags, packed, in0h, in1h)
    like = in0 if in0 is not None else in0h
    nout = 3 if mode == _lib.FVGN_MLP_DEC else 128
    out = _empty((rows, nout), like) if want_out else None
    res = _empty((rows, 128), like) if want_res else None
    d.out, d.out_res = fptr(out, True), fptr(res, True)
    outh = resh = None
    if is_tc(precision):
        if want_outh:
            outh = torch.empty((rows, 128), dtype=HDTYPE[precision], device=like.device)
            d.outh = hptr(outh)
        if want_resh:
            resh = torch.empty((rows, 128), dtype=HDTYPE[precision], device=like.device)
            d.out_resh = hptr(resh)
        if z1 is not None:
            d.z1_img = z1.ptr
    _lib.call("fvgn_mlp_forward", ctypes.byref(d), _lib.stream_ptr(like.device))
    if want_outh or want_resh:
        return out, res, outh, resh
    return out, res


def new_z1(mode, precision, rows, like):
    return Z1Image(mode, rows, like.device) if is_tc(precision) else None


_CALLER_GRAD_MODE = [True]


def apply(fn, *args):
    """fn.apply(*args) with the CALLER's grad mode recorded for _z1_for: inside Function.forward grad mode is always off
    and ctx.needs_input_grad mirrors requires_grad of the inputs even under torch.no_grad(), so neither tells a training
    forward from a rollout forward.  Direct fn.apply() calls keep the safe default (Z1 image written)."""
    prev = _CALLER_GRAD_MODE[0]
    _CALLER_GRAD_MODE[0] = torch.is_grad_enabled()
    try:
        return fn.apply(*args)
    finally:
        _CALLER_GRAD_MODE[0] = prev


def _z1_for(ctx, mode, precision, rows, like):
    """The Z1 image is only needed by a backward pass: under torch.no_grad() / with nothing requiring a gradient (the
    rollout regime of solve_without_grad_GPU.py) the forward kernel skips that store (256 B per row) altogether."""
    return new_z1(mode, precision, rows, like) if (_CALLER_GRAD_MODE[0] and any(ctx.needs_input_grad)) else None


def mlp_backward(mode, precision, rows, params, in0, in1, idx_s, idx_r, d_out, d_gather=None, d_in0=None, d_in1=None,
                 flags=0, packed=None, z1=None, in0h=None, in1h=None, d_in0h=None, d_gatherh=None, d_in0_row_ptr=None,
                 node_path=None, d_outh=None, d_in1h=None):
    """Runs the fused backward; returns the list of parameter gradients (views of one flat buffer).
    packed: the bf16 weight image used by the matching forward (bf16 mode); repacked from `params` when None.
    z1: the Z1Image the matching bf16 forward filled; when None (stand-alone use) the forward is re-run to make it.
    d_in0h: EDGE, bf16 mode: [E,256] bf16 destination of d(agg[s])|d(agg[r]) (instead of the fp32 d_in0)."""
    # d_out None: EDGE block whose outputs only feed the node block, or the upstream gradient arrives as 16-bit rows (d_outh)
    like = next(t for t in (d_out, d_outh, d_in1, d_in1h) if t is not None)
    if is_tc(precision) and z1 is None and rows > 0:
        z1 = new_z1(mode, precision, rows, like)
        mlp_forward(mode, precision, rows, params, in0, in1, idx_s, idx_r, want_out=True, want_res=False, flags=flags,
                    packed=packed, z1=z1, in0h=in0h, in1h=in1h)
    d = _mlp_desc(mode, precision, rows, params, in0, in1, idx_s, idx_r, flags, packed, in0h, in1h)
    lib = _lib.load()
    pc = int(lib.fvgn_mlp_param_count(mode))
    npart = int(lib.fvgn_mlp_bwd_partials(mode, PREC[precision], rows))
    partials = _empty((npart, pc), like)
    flat = _empty((pc,), like)
    d.d_out, d.d_gather = fptr(d_out, True), fptr(d_gather, True)
    d.d_in0, d.d_in1 = fptr(d_in0, True), fptr(d_in1, True)
    if d_outh is not None:      # 16-bit gradient streams (f16 mode: they carry the power-of-two pre-scaling)
        d.d_outh = hptr(d_outh)
    if d_in1h is not None:
        d.d_in1h = hptr(d_in1h)
    if d_in0h is not None:
        d.d_in0h = hptr(d_in0h)
    if d_gatherh is not None:
        d.d_gatherh = hptr(d_gatherh)
    if d_in0_row_ptr is not None:   # NODE, tensor-core modes: d_a2 rows leave divided by their node degree
        d.d_in0_row_ptr = iptr(d_in0_row_ptr)
    if node_path is not None:       # EDGE, tensor-core modes: node-level layer-1 backward -> (plan, d_aggh [N,128] 16-bit)
        nplan, d_aggh = node_path
        nnp = int(lib.fvgn_mlp_bwd_node_partials(nplan.N))
        d._node_partials = _empty((nnp, 128 * 256), like)
        d.inc_ptr, d.inc_code, d.n_nodes = iptr(nplan.inc_ptr), iptr(nplan.inc_code), nplan.N
        d.d_aggh, d.node_partials, d.n_node_partials = hptr(d_aggh), fptr(d._node_partials), nnp
        if NODE_LEVEL_LAYER1 == 2:
            d._node_ws, d.node_ws = _img_buffer(int(lib.fvgn_mlp_bwd_node_workspace_bytes(nplan.N)), like.device)
    d.partials, d.n_partials, d.d_params = fptr(partials), npart, fptr(flat)
    if precision == "f16":
        d.grad_unscale = grad_scale(like.device)[1:2].data_ptr()
    ws_bytes = int(lib.fvgn_mlp_bwd_workspace_bytes(mode, PREC[precision], rows))
    if ws_bytes > 0 and rows > 0:
        d._ws, d.workspace = _img_buffer(ws_bytes, like.device)
        d._z1, d.z1_img = z1, z1.ptr
    _lib.call("fvgn_mlp_backward", ctypes.byref(d), _lib.stream_ptr(like.device))
    k1 = _MLP_K1[mode]
    nout = 3 if mode == _lib.FVGN_MLP_DEC else 128
    sizes = [(128, k1), (128,), (128, 128), (128,), (nout, 128), (nout,)]
    if len(params) == 8:
        sizes += [(128,), (128,)]
    grads, off = [], 0
    for shp in sizes:
        n = 1
        for s in shp:
            n *= s
        grads.append(flat[off:off + n].view(shp))
        off += n
    return grads


GN_KEEP_E, GN_LATENTS16, GN_X_FP32 = 1, 2, 4   # GnBlockFn option bits


def placeholder(rows, like):
    """Stand-in for an fp32 latent stream that is not materialised (16-bit latent streams): a [rows,128] stride-0 view of one
    NaN -- it carries the autograd edge (its gradient is a real fp32 tensor), never data; reading it is loudly wrong."""
    return torch.full((1, 1), float("nan"), dtype=torch.float32, device=like.device).expand(rows, 128)


def is_placeholder(t):
    return t is not None and t.dim() == 2 and t.shape[0] > 1 and t.stride() == (0, 0)


# f16 mode with 16-bit latent streams: the GRADIENTS of those streams also travel as 16-bit rows between the blocks of a
# model (they carry the power-of-two pre-scaling of GradScaleFn, like every other half-precision gradient tensor).  autograd
# still sees fp32 placeholders; the rows go through a GradChannel shared by the producer and the (single) consumer of a
# placeholder latent.  FVGN_GRAD16=0 keeps fp32 gradient streams.
GRAD16 = os.environ.get("FVGN_GRAD16", "1") != "0"


class GradChannel:
    """Side channel for the 16-bit gradient rows of one (x, e) placeholder pair: the consumer's backward put()s them, the
    producer's backward take()s them when the gradient autograd hands it is a placeholder.  The placeholders are held by weak
    reference: the producing Function's ctx owns the channel and its output tensor owns the ctx, so a strong reference would
    close a cycle that only the garbage collector could free (measured: the small-mesh eager step slowed down by 2x)."""

    def __init__(self):
        self._x = self._e = None   # weak references to the placeholder tensors this channel belongs to
        self._g = {}

    @property
    def x_ref(self):
        return None if self._x is None else self._x()

    @x_ref.setter
    def x_ref(self, t):
        self._x = None if t is None else weakref.ref(t)

    @property
    def e_ref(self):
        return None if self._e is None else self._e()

    @e_ref.setter
    def e_ref(self, t):
        self._e = None if t is None else weakref.ref(t)

    def serves(self, key, tensor):
        return tensor is not None and (self.x_ref if key == "x" else self.e_ref) is tensor

    def put(self, key, rows):
        self._g[key] = rows

    def take(self, key):
        rows = self._g.pop(key, None)
        if rows is None:
            raise RuntimeError("fvgn_b200: a placeholder gradient arrived without its 16-bit rows")
        return rows


def _packed(mode, precision, params):
    return PackedWeights.get(mode, params, precision) if is_tc(precision) else None


# ------------------------------------------------------------------ gradient pre-scaling of the f16 mode
_GRAD_SCALE = {}
GRAD_SCALE_TARGET = 16.0   # the largest |d raw| of a backward pass is brought into [8, 16)


def grad_scale(device):
    """Device record [S, 1/S] of the current backward pass (S = 1 until a GradScaleFn.backward has run)."""
    key = str(device)
    if key not in _GRAD_SCALE:
        _GRAD_SCALE[key] = torch.ones(2, dtype=torch.float32, device=device)
    return _GRAD_SCALE[key]


class GradScaleFn(torch.autograd.Function):
    """Identity in forward.  Backward (f16 mode only): the gradient entering the network (d loss / d decoder output) is
    multiplied by S = 2^k, k chosen on the device so that its largest magnitude lands in [8, 16): half-precision gradient
    operands of the tensor-core backward (11-bit significand, 5-bit exponent) then sit in the middle of their range.
    The backward pass is linear in that gradient, so every fp32 gradient stream upstream carries S times its value and
    every PARAMETER gradient is multiplied by 1/S where it is emitted (fvgn_mlp_desc.grad_unscale for the fused MLPs,
    _unscale() for the Transolver block): both factors are powers of two, i.e. exact.  No host synchronisation."""

    @staticmethod
    def forward(ctx, raw, sync_ranks=False):
        ctx.sync_ranks = bool(sync_ranks)
        return raw.view_as(raw)

    @staticmethod
    def backward(ctx, g):
        rec = grad_scale(g.device)
        amax = g.detach().abs().max()
        if ctx.sync_ranks:
            # cell-partition mode: ghost-row gradients travel between ranks (HaloExchangeFn.backward) while they still carry
            # the factor S, so every rank must use the same S
            import torch.distributed as dist
            dist.all_reduce(amax, op=dist.ReduceOp.MAX)
        k = torch.floor(torch.log2(GRAD_SCALE_TARGET / amax.clamp(min=1e-30))).clamp(-60.0, 60.0)
        k = torch.where(torch.isfinite(amax) & (amax > 0), k, torch.zeros_like(k))
        rec[0] = torch.exp2(k)
        rec[1] = torch.exp2(-k)
        return g * rec[0], None


def _unscale(precision, grads, device):
    """f16 mode: parameter gradients computed by PyTorch / the ts_* kernels from pre-scaled streams -> true scale."""
    if precision != "f16":
        return grads
    inv = grad_scale(device)[1]
    return tuple(None if g is None else g * inv for g in grads)


# ------------------------------------------------------------------ Encoder
class EncoderFn(torch.autograd.Function):
    """Encoder.forward (EPD.py:116-153) with the relative edge features of importer.py:54-78 fused in.
    -> (node, edge, node_h, edge_h); the last two are the bf16 shadows (None in fp32 mode)."""

    @staticmethod
    def forward(ctx, xn, pos, plan, precision, opts, chan_out, *params):
        nb, eb = params[:8], params[8:]
        ctx.chan_out = chan_out
        ctx.set_materialize_grads(False)  # no zero tensors for the (non-differentiable) bf16 shadow outputs
        ctx.pk = (_packed(_lib.FVGN_MLP_ENC_NODE, precision, nb), _packed(_lib.FVGN_MLP_ENC_EDGE, precision, eb))
        ctx.z1 = (_z1_for(ctx, _lib.FVGN_MLP_ENC_NODE, precision, plan.N, xn), _z1_for(ctx, _lib.FVGN_MLP_ENC_EDGE, precision, plan.E, xn))
        bf = is_tc(precision)
        # 16-bit latent streams (GN_LATENTS16, see GnBlockFn): the fp32 latents are not written -- placeholders carry the
        # gradients -- except the node latent when GN_X_FP32 (TransFVGN adds it as embedding before its Transolver blocks)
        lat16 = bf and bool(int(opts) & GN_LATENTS16)
        x32 = not lat16 or bool(int(opts) & GN_X_FP32)
        rn = mlp_forward(_lib.FVGN_MLP_ENC_NODE, precision, plan.N, nb, xn, want_out=x32, packed=ctx.pk[0], z1=ctx.z1[0],
                         want_outh=bf)
        re = mlp_forward(_lib.FVGN_MLP_ENC_EDGE, precision, plan.E, eb, xn, pos, plan.edge_s, plan.edge_r, want_out=not lat16,
                         packed=ctx.pk[1], z1=ctx.z1[1], want_outh=bf)
        ctx.plan, ctx.precision = plan, precision
        ctx.save_for_backward(xn, pos, *params)
        node = rn[0] if x32 else placeholder(plan.N, xn)
        edge = re[0] if not lat16 else placeholder(plan.E, xn)
        if chan_out is not None:
            chan_out.x_ref, chan_out.e_ref = (None if x32 else node), (edge if lat16 else None)
        nodeh, edgeh = (rn[2], re[2]) if bf else (None, None)
        if bf:
            ctx.mark_non_differentiable(nodeh, edgeh)
        return node, edge, nodeh, edgeh

    @staticmethod
    def backward(ctx, d_node, d_edge, _dnh=None, _deh=None):
        xn, pos, *params = ctx.saved_tensors
        plan, precision = ctx.plan, ctx.precision
        d_nodeh = d_edgeh = None
        if is_placeholder(d_node):      # the gradient rows came through the channel (16-bit gradient streams)
            d_nodeh, d_node = ctx.chan_out.take("x"), None
        elif d_node is None:
            d_node = torch.zeros((plan.N, 128), device=xn.device)
        if is_placeholder(d_edge):
            d_edgeh, d_edge = ctx.chan_out.take("e"), None
        elif d_edge is None:
            d_edge = torch.zeros((plan.E, 128), device=xn.device)
        gn = mlp_backward(_lib.FVGN_MLP_ENC_NODE, precision, plan.N, params[:8], xn, None, None, None,
                          None if d_node is None else _c(d_node), packed=ctx.pk[0], z1=ctx.z1[0], d_outh=d_nodeh)
        ge = mlp_backward(_lib.FVGN_MLP_ENC_EDGE, precision, plan.E, params[8:], xn, pos, plan.edge_s, plan.edge_r,
                          None if d_edge is None else _c(d_edge), packed=ctx.pk[1], z1=ctx.z1[1], d_outh=d_edgeh)
        ctx.z1 = None
        return (None, None, None, None, None, None, *gn, *ge)


# ------------------------------------------------------------------ GnBlock
class GnBlockFn(torch.autograd.Function):
    """GnBlock.forward (EPD.py:177-195) = EdgeBlock (blocks.py:71-120) -> NodeBlock (blocks.py:13-63) -> residuals.

    forward : agg = Adj x ; e' = MLP_e([agg[s]|agg[r]|e]) ; a1 = incidence-sum(e' halves) ; a2 = D^-1 Adj a1 ;
              x' = MLP_n([a2|x]) ; returns (x + x', e + e', shadows)
    backward: the transposes of the three reductions are the same CSR kernels.
    fp32 mode: the MLPs recompute their hidden activations from the saved block inputs (x, agg, a2, e).
    tensor-core modes: fp32 is kept for the GRADIENTS of the residual streams x / e (and, without GN_LATENTS16, for the
              streams themselves); agg, e', a2 and the gathered gradients d(agg[s])|d(agg[r]) live as 16-bit rows (what the
              tensor cores consume anyway); saved for backward are
              the bf16 shadows xh, eh, aggh, a2h and the two Z1 tile images."""

    @staticmethod
    def forward(ctx, x, e, xh, eh, plan, precision, opts, chan_in, chan_out, *params):
        eb, nb = params[:8], params[8:]
        opts = int(opts) if is_tc(precision) else GN_KEEP_E
        # 16-bit gradient streams (f16 mode): for an input that is a placeholder served by the producer's channel, the gradient
        # leaves through that channel as 16-bit rows; chan_out is where the consumer of this block's placeholders puts theirs
        g16 = GRAD16 and precision == "f16" and chan_in is not None
        ctx.chan_in, ctx.chan_out = chan_in, chan_out
        ctx.g16_x = g16 and is_placeholder(x) and chan_in.serves("x", x)
        ctx.g16_e = g16 and is_placeholder(e) and chan_in.serves("e", e)
        ctx.keep_e = keep_e = bool(opts & GN_KEEP_E)
        # 16-bit latent streams (tensor-core modes, set by the models for their inner blocks): the residual rows are read from
        # the 16-bit shadows the MLPs consume anyway, e + e' / x + x' leave as 16-bit rows only -- the fp32 rows of both
        # streams (1 KB per edge and per node, read + written) never touch HBM; x / e and the fp32 outputs are placeholders
        # that only carry the (fp32) gradients.  GN_X_FP32: x + x' is also written in fp32 (a Transolver block reads it).
        lat16 = bool(opts & GN_LATENTS16)
        if lat16:
            if (xh is None and is_placeholder(x)) or (eh is None and is_placeholder(e)):
                raise RuntimeError("fvgn_b200: a 16-bit latent stream arrived without its shadow")
        else:
            x, e = _c(x), _c(e)
        ctx.set_materialize_grads(False)  # no zero tensors for the (non-differentiable) bf16 shadow outputs
        ctx.pk = (_packed(_lib.FVGN_MLP_EDGE, precision, eb), _packed(_lib.FVGN_MLP_NODE, precision, nb))
        ctx.plan, ctx.precision = plan, precision
        if is_tc(precision):
            BF16 = HDTYPE[precision]
            xh = xh if (xh is not None and xh.dtype == BF16) else shadow(_c(x), dtype=BF16)
            eh = eh if (eh is not None and eh.dtype == BF16) else shadow(_c(e), dtype=BF16)
            ctx.z1 = (_z1_for(ctx, _lib.FVGN_MLP_EDGE, precision, plan.E, xh), _z1_for(ctx, _lib.FVGN_MLP_NODE, precision, plan.N, xh))
            aggh = adj_reduce(xh, plan, 128, out_dtype=BF16)
            # keep_e False (last GnBlock of a model: the decoder / Transolver block read x only): the residual stream
            # e + e' and its shadow are never written, and the backward gets no upstream edge gradient (d_out = NULL)
            fl = _lib.FVGN_MLP_RESIDUAL_FROM_SHADOW if lat16 else 0
            _, e_out, e_newh, e_outh = mlp_forward(_lib.FVGN_MLP_EDGE, precision, plan.E, eb, None,
                                                   e if (keep_e and not lat16) else None, plan.edge_s, plan.edge_r,
                                                   want_out=False, want_res=keep_e and not lat16, flags=fl, packed=ctx.pk[0],
                                                   z1=ctx.z1[0], in0h=aggh, in1h=eh, want_outh=True, want_resh=keep_e)
            a1h = inc_reduce(e_newh, plan, 64, out_dtype=BF16)
            del e_newh
            a2h = adj_reduce(a1h, plan, 64, _lib.FVGN_ADJ_DIV_DST_BY_DEG, out_dtype=BF16)
            del a1h
            _, x_out, _, x_outh = mlp_forward(_lib.FVGN_MLP_NODE, precision, plan.N, nb, None, None if lat16 else x,
                                              want_out=False, want_res=(not lat16) or bool(opts & GN_X_FP32), flags=fl,
                                              packed=ctx.pk[1], z1=ctx.z1[1], in0h=a2h, in1h=xh, want_resh=True)
            if lat16:
                x_ph = x_out is None
                if x_ph:
                    x_out = placeholder(plan.N, xh)
                if keep_e:
                    e_out = placeholder(plan.E, xh)
                if chan_out is not None:
                    chan_out.x_ref, chan_out.e_ref = (x_out if x_ph else None), e_out
            ctx.save_for_backward(xh, eh, aggh, a2h, *params)
            ctx.mark_non_differentiable(*(t for t in (x_outh, e_outh) if t is not None))
            return x_out, e_out, x_outh, e_outh
        agg = adj_reduce(x, plan, 128)
        e_new, e_out = mlp_forward(_lib.FVGN_MLP_EDGE, precision, plan.E, eb, agg, e, plan.edge_s, plan.edge_r,
                                   want_out=True, want_res=True)
        a1 = inc_reduce(e_new, plan, 64)
        del e_new
        a2 = adj_reduce(a1, plan, 64, _lib.FVGN_ADJ_DIV_DST_BY_DEG)
        del a1
        _, x_out = mlp_forward(_lib.FVGN_MLP_NODE, precision, plan.N, nb, a2, x, want_out=False, want_res=True)
        ctx.save_for_backward(x, e, agg, a2, *params)
        return x_out, e_out, None, None

    @staticmethod
    def backward(ctx, d_x_out, d_e_out, _dxh=None, _deh=None):
        x, e, agg, a2, *params = ctx.saved_tensors
        plan, precision = ctx.plan, ctx.precision
        eb, nb = params[:8], params[8:]
        dev = x.device
        d_x_outh = d_e_outh = None   # upstream gradients that came through the channel as 16-bit rows
        if is_placeholder(d_x_out):
            d_x_outh, d_x_out = ctx.chan_out.take("x"), None
        else:
            d_x_out = _c(d_x_out) if d_x_out is not None else torch.zeros((plan.N, 128), device=dev)
        if is_placeholder(d_e_out):
            d_e_outh, d_e_out = ctx.chan_out.take("e"), None
        elif d_e_out is not None:
            d_e_out = _c(d_e_out)
        elif ctx.keep_e:
            d_e_out = torch.zeros((plan.E, 128), device=dev)
        d_a2 = _empty((plan.N, 64), x) if not is_tc(precision) else None
        # gradients of the inputs: fp32, or 16-bit rows handed to the producer through its channel (+ a placeholder for autograd)
        d_x = None if ctx.g16_x else _empty((plan.N, 128), x)
        d_e = None if ctx.g16_e else _empty((plan.E, 128), x)
        d_xh = torch.empty((plan.N, 128), dtype=x.dtype, device=dev) if ctx.g16_x else None
        d_eh = torch.empty((plan.E, 128), dtype=x.dtype, device=dev) if ctx.g16_e else None
        if is_tc(precision):
            BF16 = HDTYPE[precision]
            xh, eh, aggh, a2h = x, e, agg, a2
            d_a2h = torch.empty((plan.N, 64), dtype=BF16, device=dev)
            # transposed scatter_mean: d_a1 = Adj (D^-1 d_a2); the D^-1 is applied by the producing kernel's epilogue
            g_nb = mlp_backward(_lib.FVGN_MLP_NODE, precision, plan.N, nb, None, None, None, None, d_x_out, None, None, d_x,
                                packed=ctx.pk[1], z1=ctx.z1[1], in0h=a2h, in1h=xh, d_in0h=d_a2h, d_in0_row_ptr=plan.inc_ptr,
                                d_outh=d_x_outh, d_in1h=d_xh)
            d_a1h = adj_reduce(d_a2h, plan, 64, out_dtype=BF16)
            del d_a2h
            if NODE_LEVEL_LAYER1:
                # agg[s] | agg[r] columns of the edge MLP's first layer differentiated per NODE (csrc/mlp_tc_bwd_node.cu):
                # no [E,256] gradient stream, no incidence reduction of it
                d_agg = torch.empty((plan.N, 128), dtype=BF16, device=dev)
                g_eb = mlp_backward(_lib.FVGN_MLP_EDGE, precision, plan.E, eb, None, None, plan.edge_s, plan.edge_r, d_e_out,
                                    None, None, d_e, packed=ctx.pk[0], z1=ctx.z1[0], in0h=aggh, in1h=eh, d_gatherh=d_a1h,
                                    node_path=(plan, d_agg), d_outh=d_e_outh, d_in1h=d_eh)
                ctx.z1 = None
            else:
                d_srh = torch.empty((plan.E, 256), dtype=BF16, device=dev)
                g_eb = mlp_backward(_lib.FVGN_MLP_EDGE, precision, plan.E, eb, None, None, plan.edge_s, plan.edge_r, d_e_out,
                                    None, None, d_e, packed=ctx.pk[0], z1=ctx.z1[0], in0h=aggh, in1h=eh, d_in0h=d_srh,
                                    d_gatherh=d_a1h, d_outh=d_e_outh, d_in1h=d_eh)
                ctx.z1 = None
                d_agg = inc_reduce(d_srh, plan, 128, out_dtype=BF16)
                del d_srh
        else:
            g_nb = mlp_backward(_lib.FVGN_MLP_NODE, precision, plan.N, nb, a2, x, None, None, d_x_out, None, d_a2, d_x)
            d_a1 = adj_reduce(d_a2, plan, 64, _lib.FVGN_ADJ_DIV_SRC_BY_DEG)
            del d_a2
            d_sr = _empty((plan.E, 256), x)
            g_eb = mlp_backward(_lib.FVGN_MLP_EDGE, precision, plan.E, eb, agg, e, plan.edge_s, plan.edge_r, d_e_out, d_a1,
                                d_sr, d_e)
            d_agg = inc_reduce(d_sr, plan, 128)
            del d_sr
        adj_reduce(d_agg, plan, 128, _lib.FVGN_ADJ_ACCUMULATE, out=d_xh if ctx.g16_x else d_x)
        if ctx.g16_x:
            ctx.chan_in.put("x", d_xh)
            d_x = placeholder(plan.N, x)
        if ctx.g16_e:
            ctx.chan_in.put("e", d_eh)
            d_e = placeholder(plan.E, x)
        return (d_x, d_e, None, None, None, None, None, None, None, *g_eb, *g_nb)


class EdgeBlockFn(torch.autograd.Function):
    """Stand-alone EdgeBlock.forward (blocks.py:71-120): e' = MLP_e([agg[s]|agg[r]|e]), no residual."""

    @staticmethod
    def forward(ctx, x, e, plan, precision, *params):
        x, e = _c(x), _c(e)
        agg = adj_reduce(x, plan, 128)
        ctx.pk = _packed(_lib.FVGN_MLP_EDGE, precision, params)
        ctx.z1 = _z1_for(ctx, _lib.FVGN_MLP_EDGE, precision, plan.E, x)
        e_new, _ = mlp_forward(_lib.FVGN_MLP_EDGE, precision, plan.E, params, agg, e, plan.edge_s, plan.edge_r,
                               flags=_lib.FVGN_MLP_NO_RESIDUAL, packed=ctx.pk, z1=ctx.z1)
        ctx.plan, ctx.precision = plan, precision
        ctx.save_for_backward(agg, e, *params)
        return e_new

    @staticmethod
    def backward(ctx, d_e_new):
        agg, e, *params = ctx.saved_tensors
        plan, precision = ctx.plan, ctx.precision
        d_sr = _empty((plan.E, 256), e)
        d_e = _empty((plan.E, 128), e)
        g = mlp_backward(_lib.FVGN_MLP_EDGE, precision, plan.E, params, agg, e, plan.edge_s, plan.edge_r, _c(d_e_new), None,
                         d_sr, d_e, flags=_lib.FVGN_MLP_NO_RESIDUAL, packed=ctx.pk, z1=ctx.z1)
        d_agg = inc_reduce(d_sr, plan, 128)
        d_x = adj_reduce(d_agg, plan, 128)
        return (d_x, d_e, None, None, *g)


class NodeBlockFn(torch.autograd.Function):
    """Stand-alone NodeBlock.forward (blocks.py:13-63): x' = MLP_n([a2|x]) from the edge latents, no residual."""

    @staticmethod
    def forward(ctx, x, e, plan, precision, *params):
        x, e = _c(x), _c(e)
        a1 = inc_reduce(e, plan, 64)
        a2 = adj_reduce(a1, plan, 64, _lib.FVGN_ADJ_DIV_DST_BY_DEG)
        ctx.pk = _packed(_lib.FVGN_MLP_NODE, precision, params)
        ctx.z1 = _z1_for(ctx, _lib.FVGN_MLP_NODE, precision, plan.N, x)
        x_new, _ = mlp_forward(_lib.FVGN_MLP_NODE, precision, plan.N, params, a2, x, flags=_lib.FVGN_MLP_NO_RESIDUAL,
                               packed=ctx.pk, z1=ctx.z1)
        ctx.plan, ctx.precision = plan, precision
        ctx.save_for_backward(a2, x, *params)
        return x_new

    @staticmethod
    def backward(ctx, d_x_new):
        a2, x, *params = ctx.saved_tensors
        plan, precision = ctx.plan, ctx.precision
        d_a2 = _empty((plan.N, 64), x)
        d_x = _empty((plan.N, 128), x)
        g = mlp_backward(_lib.FVGN_MLP_NODE, precision, plan.N, params, a2, x, None, None, _c(d_x_new), None, d_a2, d_x,
                         flags=_lib.FVGN_MLP_NO_RESIDUAL, packed=ctx.pk, z1=ctx.z1)
        d_a1 = adj_reduce(d_a2, plan, 64, _lib.FVGN_ADJ_DIV_SRC_BY_DEG)
        # transpose of the incidence sum: d_e[f] = [d_a1[s_f] | d_a1[r_f]]
        d_e = torch.cat([d_a1[plan.edge_s.long()], d_a1[plan.edge_r.long()]], 1)
        return (d_x, d_e, None, None, *g)


# ------------------------------------------------------------------ Decoder + head
class DecoderFn(torch.autograd.Function):
    """Decoder.forward (EPD.py:215-219): Linear-GELU-Linear-GELU-Linear(128->3), no LayerNorm."""

    @staticmethod
    def forward(ctx, x, xh, precision, chan_in, *params):
        if not (is_tc(precision) and xh is not None and xh.dtype == HDTYPE[precision]):
            x = _c(x)   # (a 16-bit latent stream arrives as placeholder + shadow: only the shadow is read)
        # f16 mode: the gradient of a placeholder latent goes back to its producer as 16-bit rows (GradChannel)
        ctx.chan_in = chan_in if (GRAD16 and precision == "f16" and chan_in is not None and is_placeholder(x)
                                  and chan_in.serves("x", x)) else None
        ctx.pk = _packed(_lib.FVGN_MLP_DEC, precision, params)
        ctx.z1 = _z1_for(ctx, _lib.FVGN_MLP_DEC, precision, x.shape[0], x)
        ctx.precision = precision
        if is_tc(precision):
            xh = xh if (xh is not None and xh.dtype == HDTYPE[precision]) else shadow(x, dtype=HDTYPE[precision])
            out, _ = mlp_forward(_lib.FVGN_MLP_DEC, precision, x.shape[0], params, None, packed=ctx.pk, z1=ctx.z1, in0h=xh)
            ctx.save_for_backward(xh, *params)
        else:
            out, _ = mlp_forward(_lib.FVGN_MLP_DEC, precision, x.shape[0], params, x)
            ctx.save_for_backward(x, *params)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, *params = ctx.saved_tensors
        d_out = _c(d_out)
        n = x.shape[0]
        if ctx.chan_in is not None:
            d_xh = torch.empty((n, 128), dtype=x.dtype, device=x.device)
            g = mlp_backward(_lib.FVGN_MLP_DEC, ctx.precision, n, params, None, None, None, None, d_out, None, None,
                             packed=ctx.pk, z1=ctx.z1, in0h=x, d_in0h=d_xh)
            ctx.z1 = None
            ctx.chan_in.put("x", d_xh)
            return (placeholder(n, x), None, None, None, *g)
        d_x = _empty((n, 128), d_out)
        if is_tc(ctx.precision):
            g = mlp_backward(_lib.FVGN_MLP_DEC, ctx.precision, n, params, None, None, None, None, d_out, None, d_x,
                             packed=ctx.pk, z1=ctx.z1, in0h=x)
            ctx.z1 = None
        else:
            g = mlp_backward(_lib.FVGN_MLP_DEC, ctx.precision, n, params, x, None, None, None, d_out, None, d_x)
        return (d_x, None, None, None, *g)


INTEGRATORS = {"explicit": 0, "implicit": 1, "imex": 2}


class HeadFn(torch.autograd.Function):
    """importer.py:187-201: uvp = 10 tanh(raw/10), Dirichlet BC, uv_hat; returns phi[N,7] = [uvp | uv_hat | uv_old]."""

    @staticmethod
    def forward(ctx, raw, uv_old, y, node_type, integrator):
        raw = _c(raw)
        n = raw.shape[0]
        phi = _empty((n, 7), raw)
        _lib.call("fvgn_head_forward", fptr(raw), fptr(uv_old), fptr(y), iptr(node_type), integrator, fptr(phi), n,
                  _lib.stream_ptr(raw.device))
        ctx.integrator = integrator
        ctx.save_for_backward(raw, node_type)
        return phi

    @staticmethod
    def backward(ctx, d_phi):
        raw, node_type = ctx.saved_tensors
        d_raw = torch.empty_like(raw)
        _lib.call("fvgn_head_backward", fptr(raw), iptr(node_type), ctx.integrator, fptr(_c(d_phi)), fptr(d_raw),
                  raw.shape[0], _lib.stream_ptr(raw.device))
        return d_raw, None, None, None, None


# ------------------------------------------------------------------ per-graph column sums
def segment_colsum(x, width, ld, chunks, chunk_ptr, nchunks, nseg, center=None, power=1, col_offset=0):
    """[nseg, width] = per-graph sums of (x[:, off:off+width] - center[seg])**power, deterministic two-level tree."""
    part = _empty((max(nchunks, 1), width), x)
    out = _empty((nseg, width), x)
    st = _lib.stream_ptr(x.device)
    _lib.call("fvgn_chunk_colsum", fptr(x) + 4 * col_offset, width, ld, fptr(center, True), 0 if center is None else center.shape[1], power,
              iptr(chunks), nchunks, fptr(part), st)
    _lib.call("fvgn_chunk_combine", fptr(part), width, iptr(chunk_ptr), nseg, fptr(out), st)
    return out


# ------------------------------------------------------------------ generic deterministic segment sums
class SegmentPlan:
    """Stable CSR (ptr, perm) of an index vector: row i lists the entries k with index[k] == i in storage order, the order
    a sequential index_add_ / torch_scatter visits them.  Cached per index tensor (weak reference)."""
    _cache = {}

    def __init__(self, index, n):
        from .plan import csr_stable
        self.n = int(n)
        self.ptr, perm = csr_stable(index.reshape(-1), self.n)
        self.perm = perm.to(torch.int32).contiguous()

    @classmethod
    def of(cls, index, n):
        import weakref
        key = (index.data_ptr(), tuple(index.shape), index.dtype, str(index.device), int(n))
        hit = cls._cache.get(key)
        if hit is not None and hit[0]() is index:
            return hit[1]
        plan = cls(index, n)
        if len(cls._cache) > 64:
            cls._cache = {k: v for k, v in cls._cache.items() if v[0]() is not None}
        cls._cache[key] = (weakref.ref(index), plan)
        return plan


class SegmentSumFn(torch.autograd.Function):
    """out[i] = reduce_{k: index[k] == i} values[k]  (reduce = 'sum' | 'mean'), any trailing shape: the deterministic
    replacement of scatter(src, index, reduce=...) / index_add_ (utils/utilities.py:16-61, FVInterpolation.py:36-109,
    218-265, FVgrad.py:183-232).  fp32 sums in storage order = the result of a sequential index_add_.  Backward is a gather."""

    @staticmethod
    def forward(ctx, values, index, n, reduce):
        if reduce not in ("sum", "add", "mean"):
            raise ValueError(f"reduce={reduce!r} is not supported by fvgn_b200 (sum / mean are what the hot path uses)")
        shape = tuple(values.shape[1:])
        width = 1
        for d in shape:
            width *= int(d)
        f64 = values.dtype == torch.float64
        cdt = torch.float64 if f64 else torch.float32
        v2 = _c(values.reshape(values.shape[0], width).to(cdt))
        sp = SegmentPlan.of(index, n)
        out = torch.empty((sp.n, width), dtype=cdt, device=values.device)
        mean = reduce == "mean"
        if width > 0 and sp.n > 0:
            _lib.call("fvgn_csr_weighted_sum_f64" if f64 else "fvgn_csr_weighted_sum", _lib.ptr(v2, cdt), width, width, iptr(sp.ptr),
                      iptr(sp.perm), None, 1 if mean else 0, _lib.ptr(out, cdt), sp.n, _lib.stream_ptr(values.device))
        ctx.mean, ctx.shape, ctx.dtype = mean, shape, values.dtype
        ctx.save_for_backward(index, sp.ptr)
        return out.reshape((sp.n,) + shape).to(values.dtype)

    @staticmethod
    def backward(ctx, g):
        index, ptr = ctx.saved_tensors
        idx = index.reshape(-1).long()
        g = g.reshape(g.shape[0], -1)
        if ctx.mean:
            cnt = (ptr[1:] - ptr[:-1]).clamp(min=1).to(g.dtype).view(-1, 1)
            g = g / cnt
        return g[idx].reshape((idx.shape[0],) + ctx.shape).to(ctx.dtype), None, None, None


def segment_sum(values, index, n, reduce="sum"):
    return SegmentSumFn.apply(values, index, int(n), reduce)


# ------------------------------------------------------------------ WLSQ
class WlsqFn(torch.autograd.Function):
    """node_based_WLSQ (FVgrad.py:235-367, precomputed-moments branch) -> [N, C, nq]."""

    @staticmethod
    def forward(ctx, phi, plan, nq):
        phi = _c(phi)
        nc = phi.shape[1]
        q, qsum, qt = plan.wlsq_weights(nq)
        grad = _empty((plan.N, nc, nq), phi)
        _lib.call("fvgn_wlsq_forward", fptr(phi), nc, iptr(plan.w_ptr), iptr(plan.w_col), fptr(q), nq, fptr(grad), plan.N,
                  _lib.stream_ptr(phi.device))
        ctx.plan, ctx.nq, ctx.nc = plan, nq, nc
        return grad

    @staticmethod
    def backward(ctx, g):
        plan, nq, nc = ctx.plan, ctx.nq, ctx.nc
        q, qsum, qt = plan.wlsq_weights(nq)
        g = _c(g)
        d_phi = _empty((plan.N, nc), g)
        _lib.call("fvgn_wlsq_backward", fptr(g), nc, iptr(plan.w_tptr), iptr(plan.w_trow), fptr(qt), fptr(qsum), nq,
                  fptr(d_phi), 0, plan.N, _lib.stream_ptr(g.device))
        return d_phi, None, None


# ------------------------------------------------------------------ fused FV loss
def _fv_desc(plan, phi, grad, theta, dt):
    d = _lib.FvDesc()
    d.n_nodes, d.n_faces, d.n_cells, d.n_slots, d.n_graphs = plan.N, plan.E, plan.C, plan.K, plan.B
    d.phi, d.grad = fptr(phi), fptr(grad)
    d.pos, d.y, d.node_type = fptr(plan.pos), fptr(plan.y), iptr(plan.node_type)
    d.edge_s, d.edge_r = iptr(plan.edge_s), iptr(plan.edge_r)
    d.face_pos, d.face_area, d.face_type = fptr(plan.face_pos), fptr(plan.face_area), iptr(plan.face_type)
    d.centroid, d.cells_area, d.batch_cell = fptr(plan.centroid), fptr(plan.cells_area), iptr(plan.batch_cell)
    d.cell_ptr, d.slot_node, d.slot_face = iptr(plan.cell_ptr), iptr(plan.slot_node), iptr(plan.slot_face)
    d.slot_unv, d.slot_cell = fptr(plan.slot_unv), iptr(plan.slot_cell)
    d.face_slot_ptr, d.face_slot = iptr(plan.face_slot_ptr), iptr(plan.face_slot)
    d.node_slot_ptr, d.node_slot = iptr(plan.node_slot_ptr), iptr(plan.node_slot)
    d.inc_ptr, d.inc_code = iptr(plan.inc_ptr), iptr(plan.inc_code)
    d.theta, d.dt = fptr(theta), fptr(dt)
    return d


class FVLossFn(torch.autograd.Function):
    """Intergrator.forward (FVscheme.py:618-724) with conserved_form (:50-274): phi[N,7] -> losses[B,4]
    = (continuity, momentum-x, momentum-y, pressure-outlet) + non-differentiable uvp_node[N,3], uvp_cell[C,3]."""

    @staticmethod
    def forward(ctx, phi, plan, theta, sigma, dt, out_scale, ncn_smooth, conserved_form=True):
        phi = _c(phi)
        st = _lib.stream_ptr(phi.device)
        q, qsum, qt = plan.wlsq_weights(2)
        grad = _empty((plan.N, 7, 2), phi)
        _lib.call("fvgn_wlsq_forward", fptr(phi), 7, iptr(plan.w_ptr), iptr(plan.w_col), fptr(q), 2, fptr(grad), plan.N, st)
        res = _empty((plan.C, 4), phi)
        phic = _empty((plan.C, 5), phi)
        theta, dt = _c(theta.float()), _c(dt.reshape(-1).float())
        d = _fv_desc(plan, phi, grad, theta, dt)
        d.res, d.phic = fptr(res), fptr(phic)
        aux = None
        if not conserved_form:   # non_conserved_form (FVscheme.py:276-511): per-cell u_hat and grad(u_hat) kept for backward
            aux = _empty((plan.C, 6), phi)
            d.form, d.cell_aux = 1, fptr(aux)
        _lib.call("fvgn_fv_forward", ctypes.byref(d), st)
        sq = segment_colsum(res, 3, 4, plan.cell_chunks, plan.cell_chunk_ptr, plan.n_cell_chunks, plan.B, power=2)
        ps = segment_colsum(res, 1, 4, plan.cell_chunks, plan.cell_chunk_ptr, plan.n_cell_chunks, plan.B, power=1,
                            col_offset=3)
        S = torch.cat([sq, ps], 1)                                           # [B,4]
        halo = getattr(plan, "halo", None)
        nb = plan.B
        if halo is not None:
            # cell-partition mode: rows [0, nb) are the real graphs (owned cells), the rest collect the halo cells;
            # the squared-residual sums are completed over the ranks before the square root
            from .parallel import allreduce_sum_
            nb = halo.num_graphs
            S = allreduce_sum_(S[:nb].contiguous())
        scale = torch.cat([theta[:nb, 1:2], sigma[:nb, 0:2].float(), torch.ones_like(theta[:nb, 0:1])], 1)
        root = torch.sqrt(S)
        losses = root * scale
        uvp_node = _empty((plan.N, 3), phi)
        uvp_cell = _empty((plan.C, 3), phi)
        _lib.call("fvgn_fv_outputs", ctypes.byref(d), iptr(plan.batch_node), fptr(_c(out_scale.float())), int(ncn_smooth),
                  fptr(uvp_node), fptr(uvp_cell), st)
        ctx.plan, ctx.aux = plan, aux
        ctx.save_for_backward(phi, grad, res, root, scale, theta, dt)
        ctx.mark_non_differentiable(uvp_node, uvp_cell, grad)
        return losses, uvp_node, uvp_cell, grad

    @staticmethod
    def backward(ctx, g_losses, _gn, _gc, _gg):
        phi, grad, res, root, scale, theta, dt = ctx.saved_tensors
        plan = ctx.plan
        st = _lib.stream_ptr(phi.device)
        safe = torch.where(root > 0, root, torch.ones_like(root))
        coef = torch.where(root > 0, g_losses * scale / safe, torch.zeros_like(root))
        coef = torch.cat([coef[:, 0:3], coef[:, 3:4] * 0.5], 1)
        if coef.shape[0] < plan.B:  # cell-partition mode: halo cells (graph ids >= nb) get no gradient
            coef = torch.cat([coef, torch.zeros((plan.B - coef.shape[0], 4), dtype=coef.dtype, device=coef.device)], 0)
        coef = coef.contiguous()
        d = _fv_desc(plan, phi, grad, theta, dt)
        d_face = _empty((plan.E, 13), phi)
        d_phi = _empty((plan.N, 7), phi)
        d_grad = _empty((plan.N, 7, 2), phi)
        d.res, d.coef = fptr(res), fptr(coef)
        d.d_face, d.d_phi, d.d_grad = fptr(d_face), fptr(d_phi), fptr(d_grad)
        if ctx.aux is not None:
            d.form, d.cell_aux = 1, fptr(ctx.aux)
        _lib.call("fvgn_fv_backward", ctypes.byref(d), st)
        q, qsum, qt = plan.wlsq_weights(2)
        _lib.call("fvgn_wlsq_backward", fptr(d_grad), 7, iptr(plan.w_tptr), iptr(plan.w_trow), fptr(qt), fptr(qsum), 2,
                  fptr(d_phi), 1, plan.N, st)
        return d_phi, None, None, None, None, None, None, None


# ------------------------------------------------------------------ dense projections (tcgen05 kind::tf32)
# Tensor-core precision modes: the Linear layers of the Transolver block and their autograd run on fvgn_gemm_tf32
# (csrc/gemm_tf32.cu) -- TF32 operands, fp32 accumulation, the arithmetic the reference's GPU scripts select for nn.Linear
# (src/pre_train_Adam.py:29).  The fp32 parity mode keeps exact fp32 library GEMMs.  FVGN_TC_GEMM=0 forces the library.
TC_GEMM = os.environ.get("FVGN_TC_GEMM", "1") != "0"


def _gemm_ok(*dims):
    return TC_GEMM and all(d in (128, 256) for d in dims)


def linear_fwd(x, w, b=None, addend=None, tc=False):
    """y = x w^T (+ b) (+ addend); x [M,I], w [O,I]."""
    if tc and x.is_cuda and _gemm_ok(w.shape[0], w.shape[1]):
        x, w = _c(x), _c(w.detach())
        y = _empty((x.shape[0], w.shape[0]), x)
        _lib.call("fvgn_gemm_tf32", _lib.FVGN_GEMM_NT, fptr(x), fptr(w), fptr(None if b is None else _c(b.detach()), True),
                  fptr(None if addend is None else _c(addend), True), fptr(y), x.shape[0], w.shape[0], w.shape[1], None, 0,
                  _lib.stream_ptr(x.device))
        return y
    y = x @ w.t() if b is None else torch.addmm(b, x, w.t())
    return y if addend is None else y + addend


def linear_dgrad(dy, w, addend=None, tc=False):
    """dx = dy w (+ addend); dy [M,O], w [O,I]."""
    if tc and dy.is_cuda and _gemm_ok(w.shape[0], w.shape[1]):
        dy, w = _c(dy), _c(w.detach())
        dx = _empty((dy.shape[0], w.shape[1]), dy)
        _lib.call("fvgn_gemm_tf32", _lib.FVGN_GEMM_NN, fptr(dy), fptr(w), None, fptr(None if addend is None else _c(addend), True),
                  fptr(dx), dy.shape[0], w.shape[1], w.shape[0], None, 0, _lib.stream_ptr(dy.device))
        return dx
    return dy @ w if addend is None else torch.addmm(addend, dy, w)


def linear_wgrad(dy, x, tc=False):
    """dW = dy^T x; dy [M,O], x [M,I] -> [O,I] (deterministic: static row split, fixed-order sum of the per-CTA partials)."""
    if tc and dy.is_cuda and _gemm_ok(dy.shape[1], x.shape[1]) and dy.shape[1] // 128 * x.shape[1] <= 512:
        dy, x = _c(dy), _c(x)
        O, I, M = dy.shape[1], x.shape[1], dy.shape[0]
        npart = int(_lib.load().fvgn_gemm_tf32_partials(M))
        part = _empty((npart, O * I), dy)
        dw = _empty((O, I), dy)
        _lib.call("fvgn_gemm_tf32", _lib.FVGN_GEMM_TN, fptr(dy), fptr(x), None, None, fptr(dw), M, I, O, fptr(part), npart,
                  _lib.stream_ptr(dy.device))
        return dw
    return dy.t() @ x


# ------------------------------------------------------------------ Transolver_block (GraphTransolver.py:25-169)
TS_HEADS, TS_DH, TS_G = 8, 16, 32
TS_TOK = TS_HEADS * TS_G * TS_DH


class TsPlan:
    """Row chunks of the (graph-sorted) node rows for the slice kernels: every chunk lies inside one graph, one CTA per
    chunk; chunk_ptr groups the chunks per graph for the deterministic combine.  nb = number of graphs whose tokens are
    formed (cell-partition mode: graphs [nb, 2 nb) are the ghost rows, de-sliced with the tokens of graph id - nb)."""

    def __init__(self, batch, halo=None, num_graphs=None):
        from .plan import _chunks
        batch = batch.reshape(-1)
        n = int(batch.shape[0])
        dev = batch.device
        if num_graphs is None:   # a loader that does not say how many graphs it batched: one device round trip
            if n > 1 and bool((batch[1:] < batch[:-1]).any()):
                raise RuntimeError("fvgn_b200: Transolver kernels need the batch vector sorted by graph (Load_mesh batches are)")
            num_graphs = int(batch.max().item()) + 1 if n > 0 else 1
        nseg = int(num_graphs)
        if halo is not None:
            nseg = max(nseg, int(halo.num_graphs))
        rows_per_chunk = max(32, min(4096, -(-n // (296 * 32)) * 32))
        # chunk table built on the device (plan._chunks): U is an upper bound of the chunk count, slots past the last real
        # chunk are empty (0, 0, 0) -- their CTAs write zero partials, which every combine may read
        self.chunks, self.chunk_ptr, self.n_chunks = _chunks(batch, nseg, rows_per_chunk)
        self.n, self.nseg = n, nseg
        self.nb = nseg if halo is None else halo.num_graphs
        self.all_ptr = torch.tensor([0, self.n_chunks], dtype=torch.int32, device=dev)

    _cache = {}

    @classmethod
    def of(cls, batch, halo=None, num_graphs=None):
        key = (batch.data_ptr(), tuple(batch.shape), batch.device, None if halo is None else id(halo))
        hit = cls._cache.get(key)
        if hit is not None and hit[0]() is batch:
            return hit[1]
        import weakref
        plan = cls(batch, halo, num_graphs)
        if len(cls._cache) > 64:
            cls._cache = {k: v for k, v in cls._cache.items() if v[0]() is not None}
        cls._cache[key] = (weakref.ref(batch), plan)
        return plan


def _combine(partial, width, ptr, nseg):
    out = _empty((nseg, width), partial)
    _lib.call("fvgn_chunk_combine", fptr(partial), width, iptr(ptr), nseg, fptr(out), _lib.stream_ptr(partial.device))
    return out


def _row_partials(n):
    return int(_lib.load().fvgn_ts_row_partials(n))


_PTR01 = {}


def _ptr01(npart, device):
    """chunk_ptr [0, npart] of a single-segment combine (cached: no host-to-device copy inside a captured step)."""
    key = (npart, str(device))
    if key not in _PTR01:
        _PTR01[key] = torch.tensor([0, npart], dtype=torch.int32, device=device)
    return _PTR01[key]


def _token_attention(rec, wq, wk, wv, scale):
    """GraphTransolver.py:72-81 on the [nb, 4352] token record (numerators | norms) -> attended tokens [nb, 4096]
    (csrc/transolver.cu: ts_token_attention_fwd_kernel, one CTA per (graph, head))."""
    rec = _c(rec)
    nb = rec.shape[0]
    out = _empty((nb, TS_TOK), rec)
    _lib.call("fvgn_ts_token_attention_forward", fptr(rec), fptr(_c(wq.detach())), fptr(_c(wk.detach())), fptr(_c(wv.detach())),
              float(scale), nb, fptr(out), _lib.stream_ptr(rec.device))
    return out


def _token_attention_backward(rec, wq, wk, wv, scale, d_tok_out):
    """-> (d_rec [nb,4352], d_wq, d_wk, d_wv): the forward is recomputed from the token record; the weight gradients are
    per-(graph, head) partials summed in fixed order."""
    rec, d_tok_out = _c(rec), _c(d_tok_out)
    nb = rec.shape[0]
    d_rec = _empty((nb, _lib.FVGN_TS_TOKW), rec)
    part = _empty((max(nb * TS_HEADS, 1), 3 * TS_DH * TS_DH), rec)
    _lib.call("fvgn_ts_token_attention_backward", fptr(rec), fptr(_c(wq.detach())), fptr(_c(wk.detach())), fptr(_c(wv.detach())),
              float(scale), nb, fptr(d_tok_out), fptr(d_rec), fptr(part), _lib.stream_ptr(rec.device))
    if nb > 0:
        g = _combine(part, 3 * TS_DH * TS_DH, _ptr01(nb * TS_HEADS, rec.device), 1).reshape(3, TS_DH, TS_DH)
    else:
        g = torch.zeros((3, TS_DH, TS_DH), device=rec.device)
    return d_rec, g[0], g[1], g[2]


def _attn_forward(x, wfx, bfx, wx, bx, ws, bs, temp, wq, wk, wv, wo, scale, tsp, halo, tc=False):
    """-> (a = to_out.weight @ deslice(attention(slice(x))), tensors to keep for _attn_backward)."""
    n = x.shape[0]
    st = _lib.stream_ptr(x.device)
    wcat = torch.cat([wfx, wx], 0)
    P = linear_fwd(x, wcat, torch.cat([bfx, bx], 0), tc=tc)               # [N,256] = fx_mid | x_mid
    ws_c, bs_c, temp_c = _c(ws.detach()), _c(bs.detach()), _c(temp.detach().reshape(-1))
    sw = _empty((n, 256), x)
    part = _empty((max(tsp.n_chunks, 1), _lib.FVGN_TS_TOKW), x)
    _lib.call("fvgn_ts_slice_forward", fptr(P), fptr(ws_c), fptr(bs_c), fptr(temp_c), iptr(tsp.chunks), tsp.n_chunks,
              fptr(sw), fptr(part), st)
    rec = _combine(part, _lib.FVGN_TS_TOKW, tsp.chunk_ptr, tsp.nseg)[:tsp.nb].contiguous()
    if halo is not None:   # cell-partition mode: tokens are sums over the rows every rank owns
        from .parallel import allreduce_sum_
        rec = allreduce_sum_(rec)
    tok_out = _c(_token_attention(rec, wq, wk, wv, scale))
    out_x = _empty((n, 128), x)
    _lib.call("fvgn_ts_deslice", fptr(sw), fptr(tok_out), TS_TOK, tsp.nb, iptr(tsp.chunks), tsp.n_chunks, fptr(out_x), st)
    return linear_fwd(out_x, wo, tc=tc), (x, P, sw, rec, tok_out, out_x, wcat, ws_c, bs_c, temp_c, wq, wk, wv, wo)


def _attn_backward(saved, d_a, scale, tsp, halo, d_res=None, tc=False):
    """-> gradients of (x, wfx, bfx, wx, bx, ws, bs, temp, wq, wk, wv, wo); d_res (optional) is added to d_x inside the
    last GEMM (the block's residual gradient)."""
    x, P, sw, rec, tok_out, out_x, wcat, ws_c, bs_c, temp_c, wq, wk, wv, wo = saved
    n = x.shape[0]
    st = _lib.stream_ptr(x.device)
    d_a = _c(d_a)
    d_wo = linear_wgrad(d_a, out_x, tc=tc)
    d_ox = linear_dgrad(d_a, wo, tc=tc)                                     # [N,128]
    part = _empty((max(tsp.n_chunks, 1), _lib.FVGN_TS_TOKW), x)
    _lib.call("fvgn_ts_accumulate", fptr(sw), fptr(d_ox), iptr(tsp.chunks), tsp.n_chunks, fptr(part), st)
    acc = _combine(part, _lib.FVGN_TS_TOKW, tsp.chunk_ptr, tsp.nseg)[:, :TS_TOK]
    d_tok_out = acc[:tsp.nb]
    if tsp.nseg > tsp.nb:   # ghost rows de-slice with the tokens of graph id - nb
        extra = acc[tsp.nb:]
        d_tok_out = d_tok_out.clone()
        d_tok_out[:extra.shape[0]] += extra
    d_rec, d_wq, d_wk, d_wv = _token_attention_backward(rec, wq, wk, wv, scale, d_tok_out)
    if halo is not None:
        from .parallel import allreduce_sum_
        d_rec = allreduce_sum_(d_rec.contiguous())
    d_rec = _c(d_rec)
    dP = _empty((n, 256), x)
    ppart = _empty((max(tsp.n_chunks, 1), _lib.FVGN_TS_PARAMW), x)
    _lib.call("fvgn_ts_slice_backward", fptr(P), fptr(sw), fptr(d_ox), fptr(tok_out), fptr(d_rec), tsp.nb, fptr(ws_c),
              fptr(bs_c), fptr(temp_c), iptr(tsp.chunks), tsp.n_chunks, fptr(dP), fptr(ppart), st)
    if tsp.n_chunks > 0:
        pg = _combine(ppart, _lib.FVGN_TS_PARAMW, tsp.all_ptr, 1).reshape(-1)
    else:
        pg = torch.zeros(_lib.FVGN_TS_PARAMW, device=x.device)
    d_ws, d_bs = pg[:512].view(TS_G, TS_DH), pg[512:544]
    d_temp, d_bcat = pg[544:552].view(1, TS_HEADS, 1), pg[552:808]
    d_x = linear_dgrad(dP, wcat, d_res, tc=tc)
    d_wcat = linear_wgrad(dP, x, tc=tc)
    return (d_x, d_wcat[:128], d_bcat[:128], d_wcat[128:], d_bcat[128:], d_ws, d_bs, d_temp, d_wq, d_wk, d_wv, d_wo)


def _tail_forward(a, bo, res, gamma, beta, w1, b1, w2, b2, want_shadow, tc=False):
    """y = a + bo + res ; z = ln_2(y) ; h = GELU(z W1^T + b1) ; out = h W2^T + b2 + y -> (out, shadow, kept tensors)."""
    n = a.shape[0]
    st = _lib.stream_ptr(a.device)
    y, z, stats = _empty((n, 128), a), _empty((n, 128), a), _empty((n, 2), a)
    b1c, gc = _c(b1.detach()), _c(gamma.detach())
    _lib.call("fvgn_ts_residual_ln_forward", fptr(a), fptr(_c(bo.detach())), fptr(res), fptr(gc), fptr(_c(beta.detach())),
              fptr(y), fptr(z), fptr(stats), n, st)
    hpre = linear_fwd(z, w1, tc=tc)
    h = _empty((n, 256), a)
    _lib.call("fvgn_ts_bias_gelu_forward", fptr(hpre), fptr(b1c), fptr(h), n, st)
    o = linear_fwd(h, w2, tc=tc)
    out = _empty((n, 128), a)
    # want_shadow: None / False, or the 16-bit dtype of the shadow the next GnBlock / decoder reads
    sdt = BF16 if want_shadow is True else (want_shadow or None)
    outh = torch.empty((n, 128), dtype=sdt, device=a.device) if sdt is not None else None
    _lib.call("fvgn_ts_bias_residual", fptr(o), fptr(_c(b2.detach())), fptr(y), fptr(out), hptr(outh, True),
              _TCODE[sdt] if sdt is not None else 0, n, st)
    return out, outh, (y, stats, z, hpre, h, gc, w1, b1c, w2)


def _tail_backward(saved, d_out, tc=False):
    """-> (d_y = gradient of both a and res, d_bo, d_gamma, d_beta, d_w1, d_b1, d_w2, d_b2)."""
    y, stats, z, hpre, h, gc, w1, b1c, w2 = saved
    n = y.shape[0]
    st = _lib.stream_ptr(y.device)
    d_out = _c(d_out)
    npart = _row_partials(n)
    ptr = _ptr01(npart, y.device)
    d_w2 = linear_wgrad(d_out, h, tc=tc)
    d_h = linear_dgrad(d_out, w2, tc=tc)
    part = _empty((npart, 256), y)
    _lib.call("fvgn_ts_bias_gelu_backward", fptr(d_h), fptr(hpre), fptr(b1c), fptr(d_h), fptr(part), n, st)  # in place
    d_b1 = _combine(part, 256, ptr, 1).reshape(-1) if n > 0 else torch.zeros(256, device=y.device)
    d_w1 = linear_wgrad(d_h, z, tc=tc)
    d_z = linear_dgrad(d_h, w1, tc=tc)
    d_y = _empty((n, 128), y)
    part = _empty((npart, 512), y)
    _lib.call("fvgn_ts_residual_ln_backward", fptr(d_z), fptr(y), fptr(stats), fptr(gc), fptr(d_out), fptr(d_y), fptr(part),
              n, st)
    s = _combine(part, 512, ptr, 1).reshape(-1) if n > 0 else torch.zeros(512, device=y.device)
    return d_y, s[256:384], s[0:128], s[128:256], d_w1, d_b1, d_w2, s[384:512]


class SliceAttentionFn(torch.autograd.Function):
    """Graph_Physics_Attention_1D.graph_forward (GraphTransolver.py:48-95) without the to_out bias:
    x[N,128] -> to_out.weight @ deslice(attention(slice(x))).  Projections: fvgn_gemm_tf32 (tensor-core modes) or fp32 library
    GEMMs (parity mode); slice softmax, token sums, token attention, de-slice and their autograd are the ts_* kernels."""

    @staticmethod
    def forward(ctx, x, wfx, bfx, wx, bx, ws, bs, temp, wq, wk, wv, wo, scale, tsp, halo, precision=None):
        a, saved = _attn_forward(_c(x), wfx, bfx, wx, bx, ws, bs, temp, wq, wk, wv, wo, scale, tsp, halo, tc=is_tc(precision))
        ctx.tsp, ctx.halo, ctx.scale, ctx.precision = tsp, halo, scale, precision
        ctx.save_for_backward(*saved)
        return a

    @staticmethod
    def backward(ctx, d_a):
        g = _attn_backward(ctx.saved_tensors, d_a, ctx.scale, ctx.tsp, ctx.halo, tc=is_tc(ctx.precision))
        return (g[0], *_unscale(ctx.precision, g[1:], d_a.device), None, None, None, None)


class TransolverBlockFn(torch.autograd.Function):
    """Transolver_block.forward with in_layernorm=False (GraphTransolver.py:163-169) as ONE autograd node:
        fx = xa (+ xb) ; y = graph_forward(fx) + fx ; out = mlp(ln_2(y)) + y      (+ bf16 shadow of out)
    xb is the node embedding the TransFVGN processors add before the block (TransFVGN_v1.py:70, TransFVGN_v2.py:49).
    fvgn_gemm_tf32 (tensor-core modes) / fp32 library GEMMs (parity mode) for the five dense projections; everything else are
    the ts_* kernels; the residual gradient enters the
    last backward GEMM as its accumulator, so no separate gradient-accumulation passes remain."""

    @staticmethod
    def forward(ctx, xa, xb, wfx, bfx, wx, bx, ws, bs, temp, wq, wk, wv, wo, bo, gamma, beta, w1, b1, w2, b2, scale, tsp,
                halo, want_shadow):
        x = _c(xa) if xb is None else xa + xb
        tc = want_shadow is not None and want_shadow is not False
        a, s1 = _attn_forward(x, wfx, bfx, wx, bx, ws, bs, temp, wq, wk, wv, wo, scale, tsp, halo, tc=tc)
        out, outh, s2 = _tail_forward(a, bo, x, gamma, beta, w1, b1, w2, b2, want_shadow, tc=tc)
        ctx.tc = tc
        ctx.tsp, ctx.halo, ctx.scale, ctx.n1, ctx.has_b = tsp, halo, scale, len(s1), xb is not None
        ctx.precision = "f16" if want_shadow == torch.float16 else None
        ctx.set_materialize_grads(False)
        if outh is not None:
            ctx.mark_non_differentiable(outh)
        ctx.save_for_backward(*s1, *s2)
        return out, outh

    @staticmethod
    def backward(ctx, d_out, _dh=None):
        s1, s2 = ctx.saved_tensors[:ctx.n1], ctx.saved_tensors[ctx.n1:]
        d_y, d_bo, d_gamma, d_beta, d_w1, d_b1, d_w2, d_b2 = _tail_backward(s2, d_out, tc=ctx.tc)
        g = _attn_backward(s1, d_y, ctx.scale, ctx.tsp, ctx.halo, d_res=d_y, tc=ctx.tc)   # d fx = dP Wcat + d_y in one GEMM
        pg = _unscale(ctx.precision, (*g[1:], d_bo, d_gamma, d_beta, d_w1, d_b1, d_w2, d_b2), d_out.device)
        return (g[0], g[0] if ctx.has_b else None, *pg, None, None, None, None)


class BlockTailFn(torch.autograd.Function):
    """Second half of Transolver_block.forward (GraphTransolver.py:163-169) as one autograd node (used when the first half
    runs on ln_1(fx), in_layernorm=True): y = a + to_out.bias + fx ; out = mlp(ln_2(y)) + y."""

    @staticmethod
    def forward(ctx, a, bo, res, gamma, beta, w1, b1, w2, b2, want_shadow):
        tc = want_shadow is not None and want_shadow is not False
        out, outh, saved = _tail_forward(_c(a), bo, _c(res), gamma, beta, w1, b1, w2, b2, want_shadow, tc=tc)
        ctx.tc = tc
        ctx.precision = "f16" if want_shadow == torch.float16 else None
        ctx.set_materialize_grads(False)
        if outh is not None:
            ctx.mark_non_differentiable(outh)
        ctx.save_for_backward(*saved)
        return out, outh

    @staticmethod
    def backward(ctx, d_out, _dh=None):
        d_y, d_bo, d_gamma, d_beta, d_w1, d_b1, d_w2, d_b2 = _tail_backward(ctx.saved_tensors, d_out, tc=ctx.tc)
        d_bo, d_gamma, d_beta, d_w1, d_b1, d_w2, d_b2 = _unscale(ctx.precision, (d_bo, d_gamma, d_beta, d_w1, d_b1, d_w2, d_b2),
                                                                 d_out.device)
        return d_y, d_bo, d_y, d_gamma, d_beta, d_w1, d_b1, d_w2, d_b2, None


class BiasGeluFn(torch.autograd.Function):
    """h = GELU(hpre + bias) on [N,256] rows (MLP.linear_pre, GraphTransolver.py:105,124)."""

    @staticmethod
    def forward(ctx, hpre, bias):
        hpre = _c(hpre)
        n = hpre.shape[0]
        h = _empty((n, 256), hpre)
        _lib.call("fvgn_ts_bias_gelu_forward", fptr(hpre), fptr(_c(bias.detach())), fptr(h), n, _lib.stream_ptr(hpre.device))
        ctx.save_for_backward(hpre, bias)
        return h

    @staticmethod
    def backward(ctx, d_h):
        hpre, bias = ctx.saved_tensors
        n = hpre.shape[0]
        npart = _row_partials(n)
        d_hpre = _empty((n, 256), hpre)
        part = _empty((npart, 256), hpre)
        _lib.call("fvgn_ts_bias_gelu_backward", fptr(_c(d_h)), fptr(hpre), fptr(_c(bias.detach())), fptr(d_hpre), fptr(part), n,
                  _lib.stream_ptr(hpre.device))
        if n > 0:
            d_b = _combine(part, 256, _ptr01(npart, hpre.device), 1).reshape(-1)
        else:
            d_b = torch.zeros(256, device=hpre.device)
        return d_hpre, d_b
