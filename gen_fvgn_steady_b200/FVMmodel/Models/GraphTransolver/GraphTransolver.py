"""Transolver_block (src/FVMmodel/Models/GraphTransolver/GraphTransolver.py:25-169) -- SURVEY.md section 8(f) row f1
("next"): kept in PyTorch for now, but restated without the reference's [N,heads,slices,dim_head] temporary
(per-graph batched contractions over the sorted batch vector instead of a broadcast product + scatter_add), so it
runs at multi-million-node scale.  Parameter names/shapes match the reference state_dict."""
import torch
from torch import nn


class Graph_Physics_Attention_1D(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0, slice_num=64):
        super().__init__()
        inner_dim = dim_head * heads
        self.dim_head, self.heads, self.scale = dim_head, heads, dim_head ** -0.5
        self.temperature = nn.Parameter(torch.ones([1, heads, 1, 1]) * 0.5)
        self.graph_temperature = nn.Parameter(torch.ones([1, heads, 1]) * 0.5)
        self.in_project_x = nn.Linear(dim, inner_dim)
        self.in_project_fx = nn.Linear(dim, inner_dim)
        self.in_project_slice = nn.Linear(dim_head, slice_num)
        torch.nn.init.orthogonal_(self.in_project_slice.weight)
        self.to_q = nn.Linear(dim_head, dim_head, bias=False)
        self.to_k = nn.Linear(dim_head, dim_head, bias=False)
        self.to_v = nn.Linear(dim_head, dim_head, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout))

    def graph_forward(self, x, batch, graph_ptr=None, halo=None):
        """GraphTransolver.py:48-95.  The slice / de-slice contractions are written as ONE dense GEMM each against the
        head-block-diagonal operand ([n, H*G]^T @ [n, H*D] and [n, H*G] @ blockdiag[H*G, H*D]) instead of H batched
        GEMMs with K = n: same numbers, an order of magnitude faster at million-node scale (the off-diagonal head
        blocks of the [H*G, H*D] product are never read)."""
        n = x.size(0)
        H, D = self.heads, self.dim_head
        if graph_ptr is None:
            counts = torch.bincount(batch.reshape(-1).long())
            graph_ptr = [0] + torch.cumsum(counts, 0).cpu().tolist()
        # cell-partition mode (partition.mark_partition): graphs [0, nb) hold the rows this rank owns, [nb, 2 nb) its
        # ghost rows.  The slice tokens of graph b are summed over the OWNED rows of every rank (all-reduce) and then
        # used to de-slice both the owned and the ghost rows.
        nb = (len(graph_ptr) - 1) if halo is None else halo.num_graphs
        if halo is not None and len(graph_ptr) < 2 * nb + 1:   # a rank without ghost rows of the last graph(s)
            graph_ptr = list(graph_ptr) + [graph_ptr[-1]] * (2 * nb + 1 - len(graph_ptr))
        fx_mid = self.in_project_fx(x)                                          # [n, H*D]
        x_mid = self.in_project_x(x).view(n, H, D)
        sw = torch.softmax(self.in_project_slice(x_mid) / self.graph_temperature, dim=-1)  # [n,H,G]
        G = sw.shape[-1]
        swf = sw.reshape(n, H * G)
        outs, ghost_outs = [], []
        for b in range(nb):
            lo, hi = graph_ptr[b], graph_ptr[b + 1]
            swb, fxb = swf[lo:hi], fx_mid[lo:hi]
            norm = swb.sum(0).view(H, G)                                       # [H,G]
            full = (swb.t() @ fxb).view(H, G, H, D)                            # all head pairs; the diagonal is wanted
            num = torch.stack([full[h, :, h, :] for h in range(H)], 0)         # [H,G,D]
            if halo is not None:
                from ....parallel import AllReduceSumFn
                both = AllReduceSumFn.apply(torch.cat([num.reshape(-1), norm.reshape(-1)]), None)
                num, norm = both[:num.numel()].view(H, G, D), both[num.numel():].view(H, G)
            tok = num / (norm.unsqueeze(-1) + 1e-5)
            q, k, v = self.to_q(tok), self.to_k(tok), self.to_v(tok)
            attn = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * self.scale, dim=-1)
            out_tok = torch.matmul(attn, v)                                    # [H,G,D]
            Wd = torch.block_diag(*out_tok.unbind(0))
            outs.append(swb @ Wd)                                              # [n_b, H*D]
            if halo is not None:
                glo, ghi = graph_ptr[nb + b], graph_ptr[nb + b + 1]
                ghost_outs.append(swf[glo:ghi] @ Wd)
        outs = outs + ghost_outs
        out_x = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
        return self.to_out(out_x)


class MLP(nn.Module):
    def __init__(self, n_input, hidden_size, n_output, n_layers=1, act="gelu", res=True):
        super().__init__()
        if act != "gelu":
            raise NotImplementedError(act)
        self.n_layers, self.res = n_layers, res
        self.linear_pre = nn.Sequential(nn.Linear(n_input, hidden_size), nn.GELU())
        self.linear_post = nn.Linear(hidden_size, n_output)
        self.linears = nn.ModuleList([nn.Sequential(nn.Linear(hidden_size, hidden_size), nn.GELU()) for _ in range(n_layers)])

    def forward(self, x):
        x = self.linear_pre(x)
        for i in range(self.n_layers):
            x = self.linears[i](x) + x if self.res else self.linears[i](x)
        return self.linear_post(x)


class Transolver_block(nn.Module):
    def __init__(self, num_heads, hidden_dim, dropout, act="gelu", mlp_ratio=4, slice_num=32):
        super().__init__()
        self.ln_1 = nn.LayerNorm(hidden_dim)
        self.Attn = Graph_Physics_Attention_1D(hidden_dim, heads=num_heads, dim_head=hidden_dim // num_heads,
                                               dropout=dropout, slice_num=slice_num)
        self.ln_2 = nn.LayerNorm(hidden_dim)
        self.mlp = MLP(hidden_dim, hidden_dim * mlp_ratio, hidden_dim, n_layers=0, res=False, act=act)

    def forward(self, fx, batch, in_layernorm=False, graph_ptr=None, halo=None):
        if in_layernorm:
            fx = self.Attn.graph_forward(self.ln_1(fx), batch, graph_ptr, halo) + fx
        else:
            fx = self.Attn.graph_forward(fx, batch, graph_ptr, halo) + fx
        return self.mlp(self.ln_2(fx)) + fx
