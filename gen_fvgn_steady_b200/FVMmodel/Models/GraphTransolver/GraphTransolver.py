"""Transolver_block (src/FVMmodel/Models/GraphTransolver/GraphTransolver.py:25-169) -- SURVEY.md section 8(f) row f1.

Same classes, constructor arguments and state_dict keys as the reference.  The forward is the fused path of ops.py:
the dense projections are fvgn_gemm_tf32 calls (tensor-core modes; exact fp32 library GEMMs in the parity mode), everything
else (slice softmax, deterministic per-graph token sums, token attention, de-slice, bias + residual + LayerNorm, bias + GELU,
and all their backward passes) are the sm_100a kernels of csrc/transolver.cu.  No CPU path (the kernels raise on host
tensors); the reference's [N,8,32,16] broadcast temporary and its two torch_scatter calls do not exist here."""
import torch
from torch import nn

from .... import ops


class Graph_Physics_Attention_1D(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0, slice_num=64):
        super().__init__()
        if dim != 128 or heads != ops.TS_HEADS or dim_head != ops.TS_DH or slice_num != ops.TS_G:
            raise NotImplementedError("fvgn_b200 Transolver kernels are built for dim=128, heads=8, dim_head=16, slice_num=32 "
                                      "(TransFVGN_v1.py:24 / TransFVGN_v2.py:28)")
        if dropout:
            raise NotImplementedError("dropout is 0 on the live path")
        inner_dim = dim_head * heads
        self.dim_head, self.heads, self.scale = dim_head, heads, dim_head ** -0.5
        self.temperature = nn.Parameter(torch.ones([1, heads, 1, 1]) * 0.5)        # unused by graph_forward (as in the reference)
        self.graph_temperature = nn.Parameter(torch.ones([1, heads, 1]) * 0.5)
        self.in_project_x = nn.Linear(dim, inner_dim)
        self.in_project_fx = nn.Linear(dim, inner_dim)
        self.in_project_slice = nn.Linear(dim_head, slice_num)
        torch.nn.init.orthogonal_(self.in_project_slice.weight)
        self.to_q = nn.Linear(dim_head, dim_head, bias=False)
        self.to_k = nn.Linear(dim_head, dim_head, bias=False)
        self.to_v = nn.Linear(dim_head, dim_head, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout))

    def attend(self, x, batch, halo=None):
        """graph_forward without the to_out bias (the block fuses that bias into its residual + LayerNorm kernel)."""
        tsp = ops.TsPlan.of(batch, halo)
        return ops.SliceAttentionFn.apply(x, self.in_project_fx.weight, self.in_project_fx.bias, self.in_project_x.weight,
                                          self.in_project_x.bias, self.in_project_slice.weight, self.in_project_slice.bias,
                                          self.graph_temperature, self.to_q.weight, self.to_k.weight, self.to_v.weight,
                                          self.to_out[0].weight, self.scale, tsp, halo,
                                          getattr(self, "precision", None) or ops.default_precision())

    def graph_forward(self, x, batch, graph_ptr=None, halo=None):
        """GraphTransolver.py:48-95: x[N,128], batch[N] (sorted by graph) -> [N,128].
        Cell-partition mode (halo given, partition.mark_partition): graphs [0, nb) hold the rows this rank owns,
        [nb, 2 nb) its ghost rows; the slice tokens of graph b are summed over the OWNED rows of every rank (all-reduce)
        and de-slice both the owned and the ghost rows."""
        return self.attend(x, batch, halo) + self.to_out[0].bias


class MLP(nn.Module):
    def __init__(self, n_input, hidden_size, n_output, n_layers=1, act="gelu", res=True):
        super().__init__()
        if act != "gelu":
            raise NotImplementedError(act)
        if n_layers != 0 or n_input != 128 or hidden_size != 256 or n_output != 128:
            raise NotImplementedError("fvgn_b200 Transolver kernels are built for the block MLP 128 -> 256 -> 128, n_layers=0 "
                                      "(TransFVGN_v1.py:24 / TransFVGN_v2.py:28: mlp_ratio=2)")
        self.n_layers, self.res = n_layers, res
        self.linear_pre = nn.Sequential(nn.Linear(n_input, hidden_size), nn.GELU())
        self.linear_post = nn.Linear(hidden_size, n_output)
        self.linears = nn.ModuleList([])

    def hidden(self, z):
        return ops.BiasGeluFn.apply(z @ self.linear_pre[0].weight.t(), self.linear_pre[0].bias)

    def forward(self, x):
        return torch.addmm(self.linear_post.bias, self.hidden(x), self.linear_post.weight.t())


class Transolver_block(nn.Module):
    def __init__(self, num_heads, hidden_dim, dropout, act="gelu", mlp_ratio=4, slice_num=32):
        super().__init__()
        self.ln_1 = nn.LayerNorm(hidden_dim)
        self.Attn = Graph_Physics_Attention_1D(hidden_dim, heads=num_heads, dim_head=hidden_dim // num_heads,
                                               dropout=dropout, slice_num=slice_num)
        self.ln_2 = nn.LayerNorm(hidden_dim)
        self.mlp = MLP(hidden_dim, hidden_dim * mlp_ratio, hidden_dim, n_layers=0, res=False, act=act)
        self.last_shadow = None

    def forward(self, fx, batch, in_layernorm=False, graph_ptr=None, halo=None, embedding=None, num_graphs=None):
        """GraphTransolver.py:163-169.  `embedding` (optional) is added to fx first (the TransFVGN processors pass
        latent.x and the node embedding separately so that the sum and its gradient stay inside the fused op).
        In bf16 mode the last kernel also emits the bf16 shadow of the result (`self.last_shadow = (out, shadow)`) for
        the GnBlock / decoder that consumes it."""
        precision = getattr(self, "precision", None) or ops.default_precision()
        want_shadow = ops.HDTYPE.get(precision)   # None (fp32 mode) or the 16-bit dtype of the tensor-core mode
        A, m = self.Attn, self.mlp
        tail = (A.to_out[0].bias, self.ln_2.weight, self.ln_2.bias, m.linear_pre[0].weight, m.linear_pre[0].bias,
                m.linear_post.weight, m.linear_post.bias)
        if in_layernorm:   # original-Transolver variant: attention on ln_1(fx); not used by TransFVGN_v1/v2
            if embedding is not None:
                fx = fx + embedding
            a = A.attend(self.ln_1(fx), batch, halo)
            out, outh = ops.BlockTailFn.apply(a, tail[0], fx, *tail[1:], want_shadow)
        else:
            out, outh = ops.TransolverBlockFn.apply(
                fx, embedding, A.in_project_fx.weight, A.in_project_fx.bias, A.in_project_x.weight, A.in_project_x.bias,
                A.in_project_slice.weight, A.in_project_slice.bias, A.graph_temperature, A.to_q.weight, A.to_k.weight,
                A.to_v.weight, A.to_out[0].weight, *tail, A.scale, ops.TsPlan.of(batch, halo, num_graphs), halo, want_shadow)
        self.last_shadow = (out, outh) if want_shadow is not None else None
        return out
