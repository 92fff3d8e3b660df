"""NNmodel facade with the reference's constructor / forward / checkpoint API (src/FVMmodel/importer.py:11-313).

forward(graph_node, graph_node_x, graph_edge, graph_cell, graph_Index, is_training=True)
  -> (loss_cont[B,1], loss_mom_x[B,1], loss_mom_y[B,1], loss_press[B,1], uvp_node[N,3], uvp_cell[C,3])
Every device-side step is a hand-written kernel behind ops.py; torch only glues [B,k]-sized tensors."""
import os

import torch
from torch import nn

from .. import ops, _lib
from ..data import Data
from ..plan import GraphPlan
from ..utils.normalization import Normalizer
from .FVdiscretization.FVscheme import Intergrator


class NNmodel(nn.Module):
    def __init__(self, params) -> None:
        super().__init__()
        self.params = params
        net = params.net
        if net == "TransFVGN_v1":
            from .Models.TransFVGN.TransFVGN_v1 import Simulator
        elif net in ("TransFVGN_v2", "TransFVGN"):
            from .Models.TransFVGN.TransFVGN_v2 import Simulator
        elif net in ("EPD", "FVGN"):
            # the reference's --net FVGN does not import (GenFVGN.py:6); the pure GN composition it was meant to
            # select is EncoderProcesserDecoder (EPD.py:222-270)
            from .Models.FVGN.EPD import EncoderProcesserDecoder as Simulator
        else:
            raise ValueError(f"unknown net {net}")
        self.simulator = Simulator(
            message_passing_num=params.message_passing_num, node_input_size=params.node_input_size,
            edge_input_size=params.node_input_size + 3, node_output_size=params.node_output_size, drop_out=False,
            hidden_size=params.hidden_size, params=params)
        self.node_norm = Normalizer(size=params.node_input_size - params.node_phi_size, max_accumulations=params.dataset_size)
        self.integrator = Intergrator()
        self.node_phi_size = params.node_phi_size
        if self.node_phi_size != 3 or params.node_input_size != 12:
            raise NotImplementedError("kernels are built for node_phi_size=3, node_input_size=12 (get_param.py:69-70)")
        self.set_precision(getattr(params, "precision", None))
        self.dp_group = None
        self.initialize_weights()

    def enable_cell_partition(self, group=True):
        """Cell-partition mode: this rank holds one sub-mesh made by
        gen_fvgn_steady_b200.partition; per-graph statistics, Normalizer increments and residual norms are completed
        over the ranks, the latents' ghost rows are refreshed after the GnBlocks (as the halo depth requires) and the
        Transolver slice tokens of TransFVGN_v1/v2 are summed over the ranks' owned rows."""
        self.dp_group = group

    def enable_data_parallel(self, group=True):
        """Data-parallel mode: this rank holds a shard of the batch's graphs; the Normalizer
        increments are summed over ranks so that every rank keeps the statistics of the global batch.  The gradient
        all-reduce itself is done by the caller (gen_fvgn_steady_b200.parallel)."""
        self.dp_group = group

    def set_precision(self, precision):
        """'fp32' (SIMT FMA), 'bf16' or 'f16' (tcgen05; f16 = IEEE-half operands, the 11-bit significand of the TF32
        arithmetic the reference's GPU path uses, with power-of-two gradient pre-scaling).  None -> $FVGN_PRECISION or fp32."""
        precision = precision or ops.default_precision()
        if precision not in ops.PREC:
            raise ValueError(precision)
        self.precision = precision
        for m in self.modules():
            m.precision = precision

    def initialize_weights(self):
        self.apply(self._init_weights)

    def _init_weights(self, m):  # importer.py:42-52
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, (nn.LayerNorm, nn.BatchNorm1d)):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # ------------------------------------------------------------------ importer.py:80-130
    def update_x_attr(self, graph_node, graph_Index, plan):
        """Per-graph z-score of x[:, :3] and Normalizer on x[:, 3:12]; returns (xn[N,12], uv_old[N,2])."""
        if not getattr(graph_node, "norm_uvp", True):
            raise ValueError(" src/FVMmodel/importer.py The graph node features have already been normalized, "
                             "please check the graph.norm_uvp")
        x = graph_node.x.float().contiguous()
        N, B = plan.N, plan.B
        halo = getattr(plan, "halo", None)   # cell-partition mode: graphs [0, nb) = owned rows, [nb, 2 nb) = ghost rows
        nb = B if halo is None else halo.num_graphs
        sums = ops.segment_colsum(x, 12, 12, plan.node_chunks, plan.node_chunk_ptr, plan.n_node_chunks, B)
        counts = (plan.node_chunk_ptr_counts if hasattr(plan, "node_chunk_ptr_counts") else None)
        if counts is None:
            counts = torch.bincount(plan.batch_node.long(), minlength=B).to(torch.float32).view(-1, 1)
            if halo is not None:
                from ..parallel import allreduce_sum_
                counts = allreduce_sum_(counts[:nb].contiguous()).repeat(B // nb, 1)
            counts = counts.clamp(min=1)
            plan.node_chunk_ptr_counts = counts
        if halo is not None:
            from ..parallel import allreduce_sum_
            sums = allreduce_sum_(sums[:nb].contiguous()).repeat(B // nb, 1)   # ghost rows use their graph's statistics
        gmean = (sums[:, 0:3] / counts).contiguous()
        var = ops.segment_colsum(x, 3, 12, plan.node_chunks, plan.node_chunk_ptr, plan.n_node_chunks, B, center=gmean,
                                 power=2)
        if halo is not None:
            var = allreduce_sum_(var[:nb].contiguous()).repeat(B // nb, 1)
        var = var / counts
        gstd = torch.sqrt(var).contiguous()
        nmean = nstd = None
        if getattr(graph_node, "norm_global", True):
            if self.node_norm.wants_accumulation():
                sq = ops.segment_colsum(x, 9, 12, plan.node_chunks, plan.node_chunk_ptr, plan.n_node_chunks, B, power=2,
                                        col_offset=3)
                s1, s2, cnt = sums[:, 3:12].sum(0), sq.sum(0), N
                if halo is not None:  # owned rows only; `sums` is already global, the squares and the count are local
                    s1 = sums[:nb, 3:12].sum(0) / halo.world
                    s2, cnt = sq[:nb].sum(0), halo.rows["node"]["n_owned"]
                if self.dp_group is not None:
                    from ..parallel import allreduce_normalizer
                    s1, s2, cnt = allreduce_normalizer(self.node_norm, (s1, s2, cnt),
                                                       None if self.dp_group is True else self.dp_group)
                self.node_norm.accumulate(s1, s2, cnt)
            nmean, nstd = self.node_norm.mean().float().contiguous(), self.node_norm.std().float().contiguous()
        xn = torch.empty_like(x)
        uv_old = torch.empty((N, 2), dtype=torch.float32, device=x.device)
        _lib.call("fvgn_prologue", _lib.fptr(x), _lib.iptr(plan.batch_node), _lib.fptr(graph_Index.uvp_dim.float().contiguous()),
                  _lib.fptr(gmean), _lib.fptr(gstd), _lib.fptr(nmean, True), _lib.fptr(nstd, True), 1, _lib.fptr(xn),
                  _lib.fptr(uv_old), N, _lib.stream_ptr(x.device))
        return xn, uv_old

    def forward(self, graph_node, graph_node_x, graph_edge, graph_cell, graph_Index, is_training=True):
        if not is_training:
            raise NotImplementedError("the reference's is_training=False branch is dead code (importer.py:241-257 calls "
                                      "update_x_attr / simulator with wrong arities); use is_training=True under no_grad")
        params = self.params
        plan = GraphPlan.of(graph_node, graph_node_x, graph_edge, graph_cell, getattr(params, "order", "2nd"))
        xn, uv_old = self.update_x_attr(graph_node, graph_Index, plan)
        # The reference normalises graph_node.x IN PLACE (importer.py:121,127): a caller that keeps an alias of the tensor
        # -- solve_with_grad_GPU.py:138-145 restores `graph_node.x = uvp_pde_theta_backup` every inner iteration, the very
        # tensor the previous forward normalised -- sees the normalised values.  Same observable behaviour here.
        x_in = graph_node.x
        if (x_in.dtype == torch.float32 and x_in.is_contiguous() and x_in.shape == xn.shape and not x_in.requires_grad
                and x_in.data_ptr() != xn.data_ptr()):
            x_in.copy_(xn)
            xn = x_in
        graph_node.x = xn
        graph_node.norm_uvp = False
        graph_node.norm_global = False
        raw = self.simulator(graph_node, graph_edge, graph_cell)
        if self.precision == "f16" and torch.is_grad_enabled():
            # backward: gradient operands of the half-precision MMAs are kept in range (one S for all ranks of a partition)
            raw = ops.GradScaleFn.apply(raw, getattr(plan, "halo", None) is not None)
        phi = ops.HeadFn.apply(raw, uv_old, plan.y, plan.node_type, ops.INTEGRATORS[params.integrator])
        out_scale = (graph_Index.uvp_dim * graph_Index.sigma).float()
        losses, uvp_node, uvp_cell, grad_phi = ops.FVLossFn.apply(
            phi, plan, graph_Index.theta_PDE, graph_Index.sigma, graph_Index.dt_graph, out_scale,
            bool(getattr(params, "ncn_smooth", True)), bool(getattr(params, "conserved_form", True)))
        self._last = dict(decoder_out=raw, phi=phi, grad_phi=grad_phi)
        return losses[:, 0:1], losses[:, 1:2], losses[:, 2:3], losses[:, 3:4], uvp_node, uvp_cell

    # ------------------------------------------------------------------ importer.py:259-313
    def load_checkpoint(self, optimizer=None, scheduler=None, ckpdir=None, device=None):
        if ckpdir is None:
            raise ValueError("ckpdir is required")
        dicts = torch.load(ckpdir, map_location=device)
        self.load_state_dict(dicts["model"])
        for prefix, objs in (("optimizer", optimizer), ("scheduler", scheduler)):
            if objs is None:
                continue
            objs = objs if isinstance(objs, (list, tuple)) else [objs]
            for i, o in enumerate(objs):
                key = f"{prefix}{i}"
                if key in dicts:
                    o.load_state_dict(dicts[key])
        print("Simulator model loaded checkpoint %s" % ckpdir)

    def save_checkpoint(self, path=None, optimizer=None, scheduler=None):
        if path is None:
            raise ValueError("path is required")
        to_save = {"model": self.state_dict()}
        for prefix, objs in (("optimizer", optimizer), ("scheduler", scheduler)):
            if objs is None:
                continue
            objs = objs if isinstance(objs, (list, tuple)) else [objs]
            for i, o in enumerate(objs):
                to_save[f"{prefix}{i}"] = o.state_dict()
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        torch.save(to_save, path)
