"""Interplot (src/FVMmodel/FVdiscretization/FVInterpolation.py:36-265): the three interpolation calls the live path uses.
Inside Intergrator.forward they are fused into the flux kernel (csrc/fv.cu).  The stand-alone methods below serve callers
that use Interplot directly: gathers and per-slot arithmetic are torch elementwise ops, every scatter of the reference is
the deterministic CSR segment sum ops.segment_sum (fvgn_csr_weighted_sum over a stable CSR built by fvgn_csr_build): no
atomics, fp32 sums in the reference's scatter order, differentiable."""
import torch
from torch import nn

from ... import ops


class Interplot(nn.Module):
    def __init__(self, mesh_pos=None, centroid=None, cells_node=None, cells_index=None):
        super().__init__()

    @staticmethod
    def _segment_mean(values, index, n):
        return ops.segment_sum(values, index, n, "mean")

    def node_to_cell_2nd_order(self, node_phi=None, node_grad=None, node_hessian=None, graph_node=None, graph_cell=None,
                               cells_node=None, cells_index=None, mesh_pos=None, centroid=None):
        """mean over the cell's vertices of (phi_n + grad_n . (x_c - x_n))  (FVInterpolation.py:36-109)."""
        if node_hessian is not None:
            raise NotImplementedError("Hessian correction is dropped on the live path (FVscheme.py:668)")
        cells_node = (graph_node.face if cells_node is None else cells_node).reshape(-1).long()
        cells_index = (graph_cell.face if cells_index is None else cells_index).reshape(-1).long()
        mesh_pos = graph_node.pos if mesh_pos is None else mesh_pos
        centroid = graph_cell.pos if centroid is None else centroid
        val = node_phi[cells_node]
        if node_grad is not None:
            r = (centroid[cells_index] - mesh_pos[cells_node]).to(node_phi.dtype)
            val = val + (node_grad[cells_node] * r.unsqueeze(1)).sum(-1)
        return self._segment_mean(val, cells_index, centroid.shape[0])

    def node_to_face_2nd_order(self, node_phi=None, node_grad=None, node_hessian=None, graph_node=None, graph_edge=None):
        """half-sum over the two end nodes of (phi_n + grad_n . (x_f - x_n)); plain average when node_grad is None
        (FVInterpolation.py:111-185)."""
        if node_hessian is not None:
            raise NotImplementedError("Hessian correction is dropped on the live path (FVscheme.py:668)")
        s, r = graph_node.edge_index[0].long(), graph_node.edge_index[1].long()
        vs, vr = node_phi[s], node_phi[r]
        if node_grad is not None:
            fp, pos = graph_edge.pos.to(node_phi.dtype), graph_node.pos.to(node_phi.dtype)
            vs = vs + (node_grad[s] * (fp - pos[s]).unsqueeze(1)).sum(-1)
            vr = vr + (node_grad[r] * (fp - pos[r]).unsqueeze(1)).sum(-1)
        return (vs + vr) / 2.0

    def cell_to_node_2nd_order(self, cell_phi=None, cell_grad=None, cells_node=None, cells_index=None, centroid=None,
                               mesh_pos=None):
        """inverse-distance weighted average of the surrounding cell values (FVInterpolation.py:218-265)."""
        if cell_grad is not None:
            raise NotImplementedError("cell_grad is None on the live path (FVscheme.py:253-261)")
        cells_node, cells_index = cells_node.reshape(-1).long(), cells_index.reshape(-1).long()
        w = 1.0 / torch.norm(mesh_pos[cells_node] - centroid[cells_index], dim=-1, keepdim=True).to(cell_phi.dtype)
        n = mesh_pos.shape[0]
        num = ops.segment_sum(cell_phi[cells_index] * w, cells_node, n, "sum")
        den = ops.segment_sum(w, cells_node, n, "sum")
        return num / den
