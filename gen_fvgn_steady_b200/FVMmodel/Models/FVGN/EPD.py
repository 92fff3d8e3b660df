"""Encoder / GnBlock / Decoder / EncoderProcesserDecoder with the reference's module API and state_dict
keys (src/FVMmodel/Models/FVGN/EPD.py:10-270).  The nn.Linear / nn.LayerNorm children only hold the
parameters; every forward is a fused sm_100a kernel behind a torch.autograd.Function (ops.py)."""
import torch
from torch import nn

from ....data import Data
from .... import ops
from ....plan import GraphPlan
from .blocks import EdgeBlock, NodeBlock, mlp_params, _precision


def build_mlp(in_size, hidden_size, out_size, drop_out=True, lay_norm=True, dropout_prob=0.2):
    """Linear-GELU-Linear-GELU-Linear [+LayerNorm] parameter container (EPD.py:10-33)."""
    if drop_out:
        raise NotImplementedError("drop_out=True is never used on the live path (EPD.py:98-103,163-175)")
    if hidden_size != 128 or out_size != 128:
        raise NotImplementedError("fvgn_b200 kernels are built for hidden_size=128")
    module = nn.Sequential(nn.Linear(in_size, hidden_size), nn.GELU(), nn.Linear(hidden_size, hidden_size), nn.GELU(),
                           nn.Linear(hidden_size, out_size))
    if lay_norm:
        return nn.Sequential(module, nn.LayerNorm(normalized_shape=out_size))
    return module


def build_mlp_from_num_layer(in_size, hidden_size, out_size, drop_out=False, lay_norm=True, dropout_prob=0.2, num_layer=2):
    """EPD.py:36-63; the live path uses num_layer=2, lay_norm=False (the decoder)."""
    if drop_out or num_layer != 2 or lay_norm:
        raise NotImplementedError("only the decoder configuration (num_layer=2, no LayerNorm, no dropout) is on the path")
    return nn.Sequential(nn.Linear(in_size, hidden_size), nn.GELU(), nn.Linear(hidden_size, hidden_size), nn.GELU(),
                         nn.Linear(hidden_size, out_size))


def _carry(graph, **updates):
    out = Data(x=graph.x, edge_attr=getattr(graph, "edge_attr", None), edge_index=graph.edge_index,
               face=getattr(graph, "face", None), num_graphs=getattr(graph, "num_graphs", None),
               batch=getattr(graph, "batch", None))
    for k in ("pos", "_fvgn_plan", "_fvgn_halo"):
        if hasattr(graph, k):
            setattr(out, k, getattr(graph, k))
    for k, v in updates.items():
        setattr(out, k, v)
    return out


def _shadow_of(graph, key, master):
    """The bf16 shadow attached by the producing kernel, if it still belongs to `master` (the latent may have been
    replaced, e.g. by the Transolver block); None otherwise (the op then makes its own)."""
    c = getattr(graph, key, None)
    return c[1] if (c is not None and c[0] is master) else None


class Encoder(nn.Module):
    def __init__(self, node_input_size=128, edge_input_size=128, hidden_size=128):
        super().__init__()
        if node_input_size != 12 or edge_input_size != 15:
            raise NotImplementedError("encoder kernels are built for node_input_size=12, edge_input_size=15 (importer.py:22-30)")
        self.eb_encoder = build_mlp(edge_input_size, hidden_size, int(hidden_size), drop_out=False)
        self.nb_encoder = build_mlp(node_input_size, hidden_size, int(hidden_size), drop_out=False)

    def forward(self, graph_node, graph_cell=None, latents_16bit=False, x_fp32=True):
        """graph_node.x is the normalised [N,12] feature; the [E,15] relative edge feature of
        importer.py:54-78 is computed inside the edge-encoder kernel from x and pos.
        latents_16bit / x_fp32: as GnBlock.forward (passed by the models; tensor-core modes only)."""
        plan = GraphPlan.of(graph_node)
        opts = (ops.GN_LATENTS16 if latents_16bit and ops.LATENTS16 else 0) | (ops.GN_X_FP32 if x_fp32 else 0)
        ch = ops.GradChannel()   # 16-bit gradient rows of the placeholder latents come back through it (f16 mode)
        node_, edge_, nh, eh = ops.apply(ops.EncoderFn, graph_node.x.contiguous(), graph_node.pos.float().contiguous(), plan,
                                         _precision(self), opts, ch, *mlp_params(self.nb_encoder), *mlp_params(self.eb_encoder))
        # bf16 mode: the kernels also emit bf16 shadows of the latents; they travel with the graph as (master, shadow)
        return _carry(graph_node, x=node_, edge_attr=edge_, _fvgn_plan=plan, _xh=(node_, nh), _eh=(edge_, eh), _gch=ch), node_


class GnBlock(nn.Module):
    def __init__(self, hidden_size=128, drop_out=False):
        super().__init__()
        eb_input_dim = int(3 * hidden_size)
        nb_input_dim = int(hidden_size + (hidden_size // 2.0))
        self.nb_module = NodeBlock(hidden_size, custom_func=build_mlp(nb_input_dim, hidden_size, int(hidden_size), drop_out=drop_out))
        self.eb_module = EdgeBlock(input_size=hidden_size, custom_func=build_mlp(eb_input_dim, hidden_size, int(hidden_size), drop_out=drop_out))

    def forward(self, graph_node, keep_edge_latent=True, latents_16bit=False, x_fp32=True):
        """keep_edge_latent=False (passed by the models for their last GnBlock, whose edge latent e + e' nothing reads:
        EPD.py:262-270 decodes graph.x only): tensor-core modes skip that residual stream and its gradient; the returned
        graph then carries edge_attr = None.
        latents_16bit=True (passed by the models for blocks whose outputs feed another GnBlock / the decoder; tensor-core modes
        only): the latent streams live as 16-bit rows (graph._xh / graph._eh); graph.x / graph.edge_attr of the returned graph
        are placeholders that carry the gradients, unless x_fp32 (a Transolver block reads x next)."""
        plan = GraphPlan.of(graph_node)
        xh, eh = _shadow_of(graph_node, "_xh", graph_node.x), _shadow_of(graph_node, "_eh", graph_node.edge_attr)
        opts = (ops.GN_KEEP_E if keep_edge_latent else 0) | (ops.GN_LATENTS16 if latents_16bit and ops.LATENTS16 else 0) | \
               (ops.GN_X_FP32 if x_fp32 else 0)
        ch_in, ch_out = getattr(graph_node, "_gch", None), ops.GradChannel()
        x, e, xh, eh = ops.apply(ops.GnBlockFn, graph_node.x, graph_node.edge_attr, xh, eh, plan, _precision(self),
                                 opts, ch_in, ch_out, *mlp_params(self.eb_module.net), *mlp_params(self.nb_module.net))
        return _carry(graph_node, x=x, edge_attr=e, _fvgn_plan=plan, _xh=(x, xh), _eh=(e, eh), _gch=ch_out)


class Decoder(nn.Module):
    def __init__(self, hidden_sze=128, node_output_size=3):
        super().__init__()
        if node_output_size != 3:
            raise NotImplementedError("decoder kernel is built for node_output_size=3")
        self.node_decode_module = build_mlp_from_num_layer(hidden_sze, hidden_sze, node_output_size, drop_out=False,
                                                           lay_norm=False, num_layer=2)

    def forward(self, latent_graph_node=None):
        xh = _shadow_of(latent_graph_node, "_xh", latent_graph_node.x)
        return ops.apply(ops.DecoderFn, latent_graph_node.x, xh, _precision(self), getattr(latent_graph_node, "_gch", None),
                         *mlp_params(self.node_decode_module))


class EncoderProcesserDecoder(nn.Module):
    """Pure GN composition (EPD.py:222-270)."""

    def __init__(self, message_passing_num, edge_input_size, node_input_size, node_output_size, drop_out=False,
                 hidden_size=128, params=None):
        super().__init__()
        self.encoder = Encoder(node_input_size=node_input_size, edge_input_size=edge_input_size, hidden_size=hidden_size)
        self.GN_block_list = nn.ModuleList([GnBlock(hidden_size=hidden_size, drop_out=drop_out)
                                            for _ in range(message_passing_num)])
        self.decoder = Decoder(hidden_sze=hidden_size, node_output_size=node_output_size)

    def forward(self, graph_node=None, graph_edge=None, graph_cell=None):
        from ....parallel import halo_refresh, no_ghost_refresh
        nblk = len(self.GN_block_list)
        whole = no_ghost_refresh(graph_node, nblk)   # (a ghost refresh exchanges the fp32 rows of x and e)
        latent, _ = self.encoder(graph_node, latents_16bit=whole, x_fp32=False)  # point-wise: exact on the ghost rows too
        for i, model in enumerate(self.GN_block_list):
            # no ghost refresh: 16-bit latent streams between the blocks (the decoder reads the shadow of x)
            latent = model(latent, keep_edge_latent=not (whole and i == nblk - 1), latents_16bit=whole, x_fp32=False)
            latent = halo_refresh(latent, i, nblk)  # cell-partition mode only: ghost rows <- owners (no-op otherwise)
        return self.decoder(latent)
