"""WLSQ gradient reconstruction with the reference's function API (src/FVMmodel/FVdiscretization/FVgrad.py:183-367).

node_based_WLSQ keeps the signature of FVgrad.py:235-244.  Instead of assembling the [N,5,7] right-hand side through
atomics and running a batched LU per call (FVgrad.py:314-359), the fp64 inverse of each node's moment matrix is folded
once into per-stencil-entry weights (plan time) and the call is a single deterministic CSR pass."""
import torch

from ...data import Data
from ... import ops
from ...plan import csr_stable, GraphPlan
from ... import _lib


def moments_order(order, displacement):
    """Taylor moment vectors and 1/|d| weights (FVorder.py:7-86) for orders '1st' and '2nd'."""
    d = displacement
    w = 1.0 / torch.norm(d, dim=1, keepdim=True)
    if order == "1st":
        return d, w
    if order == "2nd":
        return torch.cat([d, 0.5 * d ** 2, d[:, 0:1] * d[:, 1:2]], dim=-1), w
    raise ValueError(f"order {order} is not supported by fvgn_b200 (3rd/4th are ill-conditioned in the reference as well)")


def compute_normal_matrix(order="1st", mesh_pos=None, edge_index=None, extra_edge_index=None, periodic_idx=None):
    """(A[N,m,m], B_twoway[2X,m,1], B_extra[Xs,m,1]) exactly as FVgrad.py:183-232 (setup-time).  The per-entry moments are
    torch elementwise ops; the scatter into A is the deterministic CSR segment sum (ops.segment_sum), i.e. the order of a
    sequential index_add_, in the dtype of mesh_pos (fp32 or fp64)."""
    if periodic_idx is not None:
        raise NotImplementedError("periodic_idx is not used on the live path")
    two = torch.cat([edge_index, edge_index.flip(0)], dim=1)
    comp = torch.cat([two, extra_edge_index], dim=1) if extra_edge_index is not None else two
    out_i, in_i = comp[0].long(), comp[1].long()
    m, w = moments_order(order, mesh_pos[out_i] - mesh_pos[in_i])
    left = (m * w).unsqueeze(2) * m.unsqueeze(1)
    if left.is_cuda or _lib._allow_host_tensors:
        A = ops.segment_sum(left, in_i, mesh_pos.shape[0], "sum")
    else:   # host tensors (the CPU oracle comparisons of the test-suite build their inputs on the host)
        A = torch.zeros((mesh_pos.shape[0],) + tuple(left.shape[1:]), dtype=left.dtype, device=left.device).index_add_(0, in_i, left)
    Bm = (w * m).unsqueeze(2)
    split = two.shape[1]
    return A, Bm[:split], Bm[split:]


class _WlsqPlan:
    """Stand-alone stencil plan for node_based_WLSQ calls outside NNmodel (grad_rec_{acc,speed}_test.py)."""
    _cache = {}

    def __init__(self, edge_index, extra_edge_index, A, B1, Bx, N):
        dev = edge_index.device
        fx, se = edge_index.long(), extra_edge_index.long()
        out_i = torch.cat([fx[0], fx[1], se[0]])
        in_i = torch.cat([fx[1], fx[0], se[1]])
        self.N, self.device = N, dev
        self.w_ptr, perm = csr_stable(in_i, N)
        self.w_col = out_i[perm].to(torch.int32).contiguous()
        nm = int(A.shape[-1])
        B1 = B1.reshape(-1, nm).float()
        flip = B1.clone()
        flip[:, 0:2] *= -1
        self.w_moments = torch.cat([B1, flip, Bx.reshape(-1, nm).float()], 0)[perm].contiguous()
        self.w_A, self.w_nm = A.float().contiguous(), nm
        self.w_row = in_i[perm].to(torch.int32).contiguous()
        self.w_tptr, self.w_tperm = csr_stable(self.w_col, N)
        self.w_trow = self.w_row[self.w_tperm].contiguous()
        self._wq = {}

    wlsq_weights = GraphPlan.wlsq_weights

    @classmethod
    def get(cls, edge_index, extra_edge_index, A, B1, Bx, N):
        key = (edge_index.data_ptr(), extra_edge_index.data_ptr(), A.data_ptr(), tuple(edge_index.shape))
        if key not in cls._cache:
            if len(cls._cache) > 8:
                cls._cache.clear()
            cls._cache[key] = cls(edge_index, extra_edge_index, A, B1, Bx, N)
        return cls._cache[key]


def node_based_WLSQ(phi_node=None, edge_index=None, extra_edge_index=None, mesh_pos=None, order=None,
                    precompute_Moments: list = None, periodic_idx=None, rt_cond=False):
    """[N, C, n_moments] (FVgrad.py:235-367).  precompute_Moments = [A, B_single, B_extra] or None (computed on the fly)."""
    if order not in ("1st", "2nd"):
        raise ValueError(f"order must be '1st' or '2nd' (got {order}); higher orders diverge in the reference too")
    if periodic_idx is not None or rt_cond:
        raise NotImplementedError("periodic_idx / rt_cond are diagnostics of the reference, not on the live path")
    if phi_node.dim() == 1:
        phi_node = phi_node.unsqueeze(1)
    if precompute_Moments is None:
        A, B2, Bx = compute_normal_matrix(order, mesh_pos.double(), edge_index, extra_edge_index)
        A, B1, Bx = A.float(), B2[: edge_index.shape[1]].float(), Bx.float()
    else:
        A, B1, Bx = precompute_Moments
    plan = _WlsqPlan.get(edge_index, extra_edge_index, A, B1, Bx, phi_node.shape[0])
    nm = plan.w_nm
    outs = []
    for c0 in range(0, phi_node.shape[1], 8):  # kernel handles up to 8 channels per pass
        outs.append(ops.WlsqFn.apply(phi_node[:, c0:c0 + 8].float().contiguous(), plan, nm))
    return outs[0] if len(outs) == 1 else torch.cat(outs, 1)
