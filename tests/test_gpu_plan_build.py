"""-m gpu: plan-time kernels (fvgn_csr_build, fvgn_hash_words, fvgn_csr_weighted_sum), the content-keyed plan cache, the
device pool and the stand-alone FV API that runs on them -- against torch / the CPU oracle."""
import numpy as np
import pytest
import torch

from tests import product_util as PU

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.int64, torch.int32])
@pytest.mark.parametrize("case", ["mesh", "ragged", "empty_rows", "one_long_row", "no_entries"])
def test_csr_build_is_the_stable_grouping(case, dtype):
    """ptr / perm of fvgn_csr_build are bit-identical to a stable sort by destination (the reference's scatter order)."""
    from gen_fvgn_steady_b200.plan import csr_stable
    PU.use_real_kernels()
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(3)
    if case == "mesh":
        n = 250_000
        dest = torch.randint(0, n, (1_000_003,), device=dev, generator=g)
    elif case == "ragged":
        n = 5000
        dest = (torch.rand(200_000, device=dev, generator=g) ** 3 * n).long().clamp(max=n - 1)
    elif case == "empty_rows":
        n = 40_000
        dest = torch.randint(0, n // 4, (30_000,), device=dev, generator=g) * 4
    elif case == "one_long_row":
        n = 100
        dest = torch.full((70_000,), 37, device=dev, dtype=torch.int64)
        dest[::9] = 5
    else:
        n, dest = 17, torch.zeros(0, dtype=torch.int64, device=dev)
    dest = dest.to(dtype)
    ptr, perm = csr_stable(dest, n)
    ref_perm = torch.sort(dest.long(), stable=True).indices
    ref_ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    ref_ptr[1:] = torch.cumsum(torch.bincount(dest.long(), minlength=n), 0)
    assert torch.equal(ptr.long(), ref_ptr)
    assert torch.equal(perm, ref_perm)


def test_plan_on_device_matches_oracle_bitwise():
    """GraphPlan built with the device kernels is bit-exact against the oracle's numpy plans (the reference's scatter
    order), like tests/test_plan.py checks for the host-side build."""
    from oracle import fvgn_oracle as O
    from gen_fvgn_steady_b200.mesh import synthetic as S
    from gen_fvgn_steady_b200.plan import GraphPlan
    from tests.case_inputs import product_graphs
    PU.use_real_kernels()
    meshes, uvps = zip(*[S.make_case(14, kind="mixed", bc="channel", seed=1), S.make_case(9, kind="tri", bc="cavity", seed=2)])
    gn, gx, ge, gc, gi = product_graphs(list(meshes), list(uvps), "cuda")
    plan = GraphPlan.build(gn, gx, ge, gc)
    N = plan.N
    npy = lambda t: t.cpu().numpy()
    rowptr, edge, role, nbr = O.node_incidence_plan(npy(gn.edge_index), N)
    assert np.array_equal(npy(plan.inc_ptr), rowptr)
    assert np.array_equal(npy(plan.inc_code), edge * 2 + role)
    assert np.array_equal(npy(plan.inc_nbr), nbr)
    wptr, wperm, wcol = O.wlsq_entry_plan(npy(gx.face_node_x), npy(gx.support_edge), N)
    assert np.array_equal(npy(plan.w_ptr), wptr) and np.array_equal(npy(plan.w_col), wcol)
    cptr, cperm = O.csr_stable(npy(gc.face), plan.C)
    assert np.array_equal(npy(plan.cell_ptr), cptr)
    assert np.array_equal(npy(plan.slot_node), npy(gn.face)[cperm]) and np.array_equal(npy(plan.slot_face), npy(ge.face)[cperm])
    nptr, nperm = O.csr_stable(npy(plan.slot_node), N)
    assert np.array_equal(npy(plan.node_slot_ptr), nptr) and np.array_equal(npy(plan.node_slot), nperm)
    fptr_, fperm = O.csr_stable(npy(plan.slot_face), plan.E)
    assert np.array_equal(npy(plan.face_slot_ptr), fptr_) and np.array_equal(npy(plan.face_slot), fperm)
    tptr, tperm = O.csr_stable(npy(plan.w_col), N)
    assert np.array_equal(npy(plan.w_tptr), tptr) and np.array_equal(npy(plan.w_trow), npy(plan.w_row)[tperm])
    # chunk tables: every graph's rows covered once, in order, by chunks of <= 4096 rows
    assert plan.B == 2
    for chunks, cptr_, batch in ((plan.node_chunks, plan.node_chunk_ptr, plan.batch_node), (plan.cell_chunks, plan.cell_chunk_ptr, plan.batch_cell)):
        ch, cp, bt = npy(chunks), npy(cptr_), npy(batch)
        row = 0
        for gph in range(plan.B):
            for k in range(cp[gph], cp[gph + 1]):
                assert ch[k, 0] == gph and ch[k, 1] == row and 0 < ch[k, 2] - ch[k, 1] <= 4096
                assert (bt[ch[k, 1]:ch[k, 2]] == gph).all()
                row = ch[k, 2]
        assert row == len(bt)


def test_plan_cache_recognises_a_fresh_batch_by_content():
    """A loader hands new batch objects (new tensors, same content) every step: the plan is found through the content hash;
    any change of connectivity, geometry or boundary targets misses."""
    from gen_fvgn_steady_b200.mesh import synthetic as S
    from gen_fvgn_steady_b200.plan import GraphPlan
    from tests.case_inputs import product_graphs
    PU.use_real_kernels()
    mesh, uvp = S.make_case(16, kind="mixed", bc="channel", seed=5)
    g1 = product_graphs([mesh], [uvp], "cuda")
    g2 = product_graphs([mesh], [uvp], "cuda")
    p1 = GraphPlan.of(*g1[:4])
    assert GraphPlan.of(*g1[:4]) is p1                      # same object: pointer key
    assert GraphPlan.of(*g2[:4]) is p1                      # fresh object, same content
    g3 = product_graphs([mesh], [uvp], "cuda")
    g3[0].y = g3[0].y + 1e-3                                # other Dirichlet targets
    assert GraphPlan.of(*g3[:4]) is not p1
    g4 = product_graphs([mesh], [uvp], "cuda")
    ei = g4[0].edge_index.clone()
    ei[:, [0, 1]] = ei[:, [1, 0]]                           # same edge set, other storage order: another scatter order
    g4[0].edge_index = ei
    assert GraphPlan.of(*g4[:4]) is not p1


@pytest.mark.parametrize("reduce", ["sum", "mean"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_segment_sum_matches_sequential_index_add(reduce, dtype):
    from gen_fvgn_steady_b200 import ops
    PU.use_real_kernels()
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(11)
    n, m = 3001, 20_011
    index = torch.randint(0, n - 5, (m,), device=dev, generator=g)      # the last rows stay empty
    vals = torch.randn((m, 5, 5), device=dev, generator=g, dtype=dtype).requires_grad_(True)
    out = ops.segment_sum(vals, index, n, reduce)
    ref = torch.zeros((n, 5, 5), dtype=dtype).index_add_(0, index.cpu(), vals.detach().cpu())   # sequential on the host
    if reduce == "mean":
        ref = ref / torch.bincount(index.cpu(), minlength=n).clamp(min=1).to(dtype).view(-1, 1, 1)
    assert out.shape == ref.shape and out.dtype == dtype
    if reduce == "sum":
        assert torch.equal(out.detach().cpu(), ref)        # same order, same roundings
    else:
        assert float((out.detach().cpu() - ref).abs().max()) < (1e-6 if dtype == torch.float32 else 1e-14)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    gref = w[index]
    if reduce == "mean":
        gref = gref / torch.bincount(index, minlength=n).clamp(min=1).to(dtype)[index].view(-1, 1, 1)
    assert float((vals.grad - gref).abs().max()) < 1e-6


def test_standalone_fv_api_against_oracle():
    """utils.utilities.calc_*_centered_with_*_attr, Interplot.* and compute_normal_matrix on the device (segment sums on the
    CSR kernels) against the oracle's restatement of the reference on the host."""
    from oracle import fvgn_oracle as O
    from gen_fvgn_steady_b200.mesh import synthetic as S
    from gen_fvgn_steady_b200.utils import utilities as U
    from gen_fvgn_steady_b200.FVMmodel.FVdiscretization.FVInterpolation import Interplot
    from gen_fvgn_steady_b200.FVMmodel.FVdiscretization.FVgrad import compute_normal_matrix
    from tests.case_inputs import product_graphs
    PU.use_real_kernels()
    mesh, uvp = S.make_case(18, kind="mixed", bc="channel", seed=9)
    gn, gx, ge, gc, gi = product_graphs([mesh], [uvp], "cuda")
    N, C = gn.pos.shape[0], gc.pos.shape[0]
    g = torch.Generator(device="cuda").manual_seed(1)
    phi = torch.randn((N, 3), device="cuda", generator=g)
    cells_node, cells_index = gn.face.reshape(-1), gc.face.reshape(-1)
    # node -> cell mean / sum, cell -> node mean
    for reduce in ("mean", "sum"):
        got = U.calc_cell_centered_with_node_attr(phi, cells_node, cells_index, reduce=reduce)
        ref = torch.zeros((C, 3)).index_add_(0, cells_index.cpu(), phi.cpu()[cells_node.cpu()])
        if reduce == "mean":
            ref = ref / torch.bincount(cells_index.cpu(), minlength=C).clamp(min=1).view(-1, 1)
        assert float((got.cpu() - ref).abs().max()) < 1e-6
    cphi = torch.randn((C, 3), device="cuda", generator=g)
    got = U.calc_node_centered_with_cell_attr(cphi, cells_node, cells_index, reduce="mean")
    ref = torch.zeros((N, 3)).index_add_(0, cells_node.cpu(), cphi.cpu()[cells_index.cpu()])
    ref = ref / torch.bincount(cells_node.cpu(), minlength=N).clamp(min=1).view(-1, 1)
    assert float((got.cpu() - ref).abs().max()) < 1e-6
    # Interplot against the oracle (fp64 on the host)
    og = O.graphs_from_meshes([mesh], [uvp], torch.float64)
    grad = torch.randn((N, 3, 2), device="cuda", generator=g)
    ip = Interplot()
    got = ip.node_to_cell_2nd_order(node_phi=phi, node_grad=grad, graph_node=gn, graph_cell=gc)
    ref = O.node_to_cell(phi.double().cpu(), grad.double().cpu(), og["cells_node"], og["cells_index"], og["pos"], og["centroid"])
    assert float((got.double().cpu() - ref).abs().max()) < 1e-5
    got = ip.cell_to_node_2nd_order(cell_phi=cphi, cells_node=cells_node, cells_index=cells_index, centroid=gc.pos, mesh_pos=gn.pos)
    ref = O.cell_to_node(cphi.double().cpu(), og["cells_node"], og["cells_index"], og["centroid"], og["pos"])
    got = ip.node_to_face_2nd_order(node_phi=phi, node_grad=grad, graph_node=gn, graph_edge=ge)
    ref = O.node_to_face(phi.double().cpu(), grad.double().cpu(), og["edge_index"], og["pos"], og["face_pos"])
    assert float((got.double().cpu() - ref).abs().max()) < 1e-5
    got = ip.cell_to_node_2nd_order(cell_phi=cphi, cells_node=cells_node, cells_index=cells_index, centroid=gc.pos, mesh_pos=gn.pos)
    ref = O.cell_to_node(cphi.double().cpu(), og["cells_node"], og["cells_index"], og["centroid"], og["pos"])
    assert float((got.double().cpu() - ref).abs().max()) < 1e-5
    # WLSQ moment matrices: device (fp64 segment sums) vs the loader's stored A
    A, B2, Bx = compute_normal_matrix("2nd", gn.pos.double(), gx.face_node_x, gx.support_edge)
    assert float((A.float() - gx.A_node_to_node).abs().max() / gx.A_node_to_node.abs().max()) < 1e-6


def test_device_pool_sample_and_payback():
    """DevicePool.sample hands fresh batch objects equal to a host-collated batch, finds the plan without rebuilding it, and
    payback updates the pool in place (Graph_loader.py:370-383)."""
    from gen_fvgn_steady_b200.mesh import synthetic as S
    from gen_fvgn_steady_b200.plan import GraphPlan
    from gen_fvgn_steady_b200.pool import DevicePool
    from tests.case_inputs import product_graphs
    PU.use_real_kernels()
    cases = [S.make_case(10 + i, kind="mixed", bc="channel", seed=i) for i in range(3)]
    meshes, uvps = [c[0] for c in cases], [c[1] for c in cases]
    pool = DevicePool(meshes, uvps, "cuda")
    ids = [2, 0]
    graphs, gidx = pool.sample(ids)
    ref = product_graphs([meshes[i] for i in ids], [uvps[i] for i in ids], "cuda")
    for a, b in zip(graphs, ref):
        for k in b.keys():
            va, vb = getattr(a, k), getattr(b, k)
            if torch.is_tensor(vb):
                assert torch.equal(va, vb), k
    plan = GraphPlan.of(*graphs[:4])
    graphs2, gidx2 = pool.sample(ids)
    assert graphs2[0] is not graphs[0] and GraphPlan.of(*graphs2[:4]) is plan
    new = torch.randn((graphs[0].x.shape[0], 3), device="cuda")
    pool.payback(new, gidx)
    graphs3, _ = pool.sample(ids)
    assert torch.equal(graphs3[0].x[:, :3], new)
    other, _ = pool.sample([1])
    assert torch.equal(other[0].x[:, :3], torch.as_tensor(uvps[1], dtype=torch.float32).cuda())   # untouched graph
