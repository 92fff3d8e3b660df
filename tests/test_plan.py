"""GraphPlan index structures are bit-exact against the oracle's numpy plans (stable CSR = the reference's scatter order)."""
import numpy as np
import pytest
import torch

from oracle import fvgn_oracle as O
from tests import product_util as PU
from tests.case_inputs import case_meshes, product_graphs


@pytest.fixture(autouse=True)
def _emu():
    PU.use_emulated_kernels()  # GraphPlan calls fvgn_wlsq_weights at build time
    yield
    PU.use_real_kernels()


@pytest.mark.parametrize("name", ["synth_ns_batch2_v2", "poisson_quad_tri_v2"])
def test_plan_matches_oracle_bitwise(name):
    from gen_fvgn_steady_b200.plan import GraphPlan
    meshes, uvps, _ = case_meshes(name)
    gn, gx, ge, gc, gi = product_graphs(meshes, uvps, "cpu")
    plan = GraphPlan.build(gn, gx, ge, gc)
    N = plan.N
    rowptr, edge, role, nbr = O.node_incidence_plan(gn.edge_index.numpy(), N)
    assert np.array_equal(plan.inc_ptr.numpy(), rowptr)
    assert np.array_equal(plan.inc_code.numpy(), edge * 2 + role)
    assert np.array_equal(plan.inc_nbr.numpy(), nbr)
    wptr, wperm, wcol = O.wlsq_entry_plan(gx.face_node_x.numpy(), gx.support_edge.numpy(), N)
    assert np.array_equal(plan.w_ptr.numpy(), wptr)
    assert np.array_equal(plan.w_col.numpy(), wcol)
    cptr, cperm = O.csr_stable(gc.face.numpy(), plan.C)
    assert np.array_equal(plan.cell_ptr.numpy(), cptr)
    assert np.array_equal(plan.slot_node.numpy(), gn.face.numpy()[cperm])
    assert np.array_equal(plan.slot_face.numpy(), ge.face.numpy()[cperm])
    # transposes are permutations of the same entry sets
    assert sorted(plan.w_trow.tolist()) == sorted(plan.w_row.tolist())
    assert plan.w_tptr[-1] == plan.w_ptr[-1]
    nptr, nperm = O.csr_stable(plan.slot_node.numpy(), N)
    assert np.array_equal(plan.node_slot_ptr.numpy(), nptr) and np.array_equal(plan.node_slot.numpy(), nperm)


def test_segmented_sums_equal_sequential_fp32_sum():
    """adj_reduce sums each row in CSR (= reference scatter) order: bit-equal to a sequential index_add_ on CPU."""
    from gen_fvgn_steady_b200 import ops
    from gen_fvgn_steady_b200.plan import GraphPlan
    meshes, uvps, _ = case_meshes("synth_ns_batch2_v1")
    gn, gx, ge, gc, gi = product_graphs(meshes, uvps, "cpu")
    plan = GraphPlan.build(gn)
    x = torch.randn(plan.N, 128, generator=torch.Generator().manual_seed(0))
    got = ops.adj_reduce(x, plan, 128)
    s, r = gn.edge_index[0], gn.edge_index[1]
    ref = O.scatter_add(x[torch.cat([r, s])], torch.cat([s, r]), plan.N)
    assert torch.equal(got, ref)
    e = torch.randn(plan.E, 128, generator=torch.Generator().manual_seed(1))
    a1 = ops.inc_reduce(e, plan, 64)
    ref1 = O.scatter_add(torch.cat(torch.chunk(e, 2, dim=-1), 0), torch.cat([s, r]), plan.N)
    assert torch.equal(a1, ref1)
