"""Intergrator with the reference's forward signature (src/FVMmodel/FVdiscretization/FVscheme.py:618-724).

conserved_form (:50-274) or non_conserved_form (:276-511, params.conserved_form=False) + node_to_cell / node_to_face
interpolation + _fix_face_flux_BC + the per-graph norms are one fused forward kernel and a hand-written backward
(ops.FVLossFn)."""
import torch
from torch import nn

from ... import ops
from ...plan import GraphPlan
from .FVflux import FV_flux


class Intergrator(FV_flux):
    def __init__(self):
        super().__init__()
        self.epoch = 0

    def forward(self, uvp_new_node=None, uv_hat_node=None, uv_old_node=None, graph_node=None, graph_node_x=None,
                graph_edge=None, graph_cell=None, graph_Index=None, params=None, phi_node=None):
        """-> (loss_cont[B,1], loss_mom_x[B,1], loss_mom_y[B,1], loss_press[B,1], uvp_node[N,3], uvp_cell[C,3]).
        uvp_node / uvp_cell are dimensionless here (the caller re-dimensionalises, importer.py:223-231)."""
        order = getattr(params, "order", "2nd")
        plan = GraphPlan.of(graph_node, graph_node_x, graph_edge, graph_cell, order)
        if phi_node is None:
            phi_node = torch.cat([uvp_new_node[:, 0:3], uv_hat_node[:, 0:2], uv_old_node[:, 0:2]], dim=-1)  # :643-646
        ones = torch.ones((plan.B, 3), dtype=torch.float32, device=phi_node.device)
        losses, uvp_node, uvp_cell, _ = ops.FVLossFn.apply(phi_node, plan, graph_Index.theta_PDE, graph_Index.sigma,
                                                           graph_Index.dt_graph, ones, bool(getattr(params, "ncn_smooth", True)),
                                                           bool(getattr(params, "conserved_form", True)))
        return losses[:, 0:1], losses[:, 1:2], losses[:, 2:3], losses[:, 3:4], uvp_node, uvp_cell
