"""-m gpu: size-independent properties of the hot path at a BASELINE.json size (1 M-cell synthetic mesh, config 5), where
the CPU oracle would take minutes: exactness of the WLSQ reconstruction on a quadratic field (grad_rec_acc_test.py's
check), geometric closure of the cells, linearity / symmetry of the CSR reductions, the adjoint identity of the fused
forward/backward kernels, and bit-reproducibility of a whole training step."""
import pytest
import torch

pytestmark = pytest.mark.gpu

N_SIDE = 1000  # 1 M cells, 1.002 M nodes, 2.002 M faces


@pytest.fixture(scope="module")
def big():
    from gen_fvgn_steady_b200.mesh import synthetic_torch as ST
    from gen_fvgn_steady_b200.mesh.batching import graphs_from_meshes
    from gen_fvgn_steady_b200.plan import GraphPlan
    from tests import product_util as PU
    PU.use_real_kernels()
    dev = torch.device("cuda")
    mesh, uvp = ST.make_case(N_SIDE, kind="quad", bc="cavity", seed=0, device=dev)
    graphs = graphs_from_meshes([mesh], [uvp], dev)
    plan = GraphPlan.of(graphs[0], graphs[1], graphs[2], graphs[3], "2nd")
    return mesh, graphs, plan


def test_wlsq_reproduces_quadratic_field_exactly(big):
    """A 2nd-order WLSQ stencil differentiates a quadratic exactly (src/grad_rec_acc_test.py:87-98,168-181 checks the
    same thing on an analytic field); 1 M nodes, fp32: relative error of the gradient below 2e-4."""
    from gen_fvgn_steady_b200 import ops
    mesh, graphs, plan = big
    pos = plan.pos.double()
    x, y = pos[:, 0], pos[:, 1]
    coef = torch.tensor([[0.3, -1.2, 0.7, 2.0, -0.5, 1.1], [1.0, 0.4, -0.9, -1.5, 0.8, 0.2]], dtype=torch.float64, device=pos.device)
    phi = torch.stack([c[0] + c[1] * x + c[2] * y + c[3] * x * x + c[4] * y * y + c[5] * x * y for c in coef], 1)
    gx = torch.stack([c[1] + 2 * c[3] * x + c[5] * y for c in coef], 1)
    gy = torch.stack([c[2] + 2 * c[4] * y + c[5] * x for c in coef], 1)
    grad = ops.WlsqFn.apply(phi.float().contiguous(), plan, 2)           # [N, 2, 2]
    ref = torch.stack([gx, gy], 2)
    err = float((grad.double() - ref).abs().max() / ref.abs().max())
    assert err < 2e-4, err


def test_cells_are_closed_surfaces(big):
    """sum over the faces of a cell of (outward unit normal * face length) = 0 (parse_to_h5.py:430-472 asserts the same)."""
    mesh, graphs, plan = big
    S = plan.slot_unv.double() * plan.face_area.double()[plan.slot_face.long()].unsqueeze(1)
    tot = torch.zeros((plan.C, 2), dtype=torch.float64, device=S.device).index_add_(0, plan.slot_cell.long(), S)
    scale = float(plan.face_area.double().mean())
    assert float(tot.abs().max()) < 1e-5 * scale


def test_reductions_linear_and_self_adjoint(big):
    """Adj is symmetric: <Adj x, y> = <x, Adj y>; the incidence reduction and its gather transpose are adjoint; all at 1 M rows."""
    from gen_fvgn_steady_b200 import ops
    mesh, graphs, plan = big
    dev = plan.pos.device
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn((plan.N, 128), device=dev, generator=g)
    y = torch.randn((plan.N, 128), device=dev, generator=g)
    ax, ay = ops.adj_reduce(x, plan, 128), ops.adj_reduce(y, plan, 128)
    lhs, rhs = float((ax.double() * y.double()).sum()), float((x.double() * ay.double()).sum())
    assert abs(lhs - rhs) < 1e-6 * max(abs(lhs), 1.0) + 1e-3
    a2 = ops.adj_reduce(2.0 * x + y, plan, 128)
    assert float((a2 - (2.0 * ax + ay)).abs().max()) < 1e-4 * float(a2.abs().max())
    e = torch.randn((plan.E, 128), device=dev, generator=g)
    z = torch.randn((plan.N, 64), device=dev, generator=g)
    ie = ops.inc_reduce(e, plan, 64)                                     # [N,64]
    gath = torch.cat([z[plan.edge_s.long()], z[plan.edge_r.long()]], 1)  # transpose: d_e[f] = [z[s_f] | z[r_f]]
    lhs, rhs = float((ie.double() * z.double()).sum()), float((e.double() * gath.double()).sum())
    assert abs(lhs - rhs) < 1e-6 * max(abs(lhs), 1.0) + 1e-3


@pytest.mark.parametrize("precision", ["bf16"])
def test_gnblock_adjoint_identity(big, precision):
    """<J dx, w> = <dx, J^T w> for one GnBlock at 1 M cells: the hand-written backward is the transpose of the forward
    (finite difference along a random direction vs the backward's directional derivative; bf16-mode tolerance)."""
    from gen_fvgn_steady_b200.FVMmodel.Models.FVGN.EPD import GnBlock
    from gen_fvgn_steady_b200.data import Data
    mesh, graphs, plan = big
    dev = plan.pos.device
    torch.manual_seed(0)
    blk = GnBlock().to(dev)
    for m in blk.modules():
        m.precision = precision
    with torch.no_grad():
        for q in blk.parameters():
            if q.dim() == 2:
                q.copy_(torch.randn_like(q) / q.shape[1] ** 0.5)
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn((plan.N, 128), device=dev, generator=g, requires_grad=True)
    e = torch.randn((plan.E, 128), device=dev, generator=g, requires_grad=True)
    dx = torch.randn((plan.N, 128), device=dev, generator=g)
    w = torch.randn((plan.N, 128), device=dev, generator=g)

    def run(xx):
        gr = Data(x=xx, edge_attr=e, edge_index=graphs[0].edge_index, pos=graphs[0].pos, _fvgn_plan=plan)
        return blk(gr).x

    out = run(x)
    (gx,) = torch.autograd.grad((out * w).sum(), x)
    analytic = float((gx.double() * dx.double()).sum())
    eps = 1e-2
    with torch.no_grad():
        fd = float((((run(x + eps * dx) - run(x - eps * dx)).double() / (2 * eps)) * w.double()).sum())
    assert abs(fd - analytic) < 3e-2 * max(abs(fd), abs(analytic)), (fd, analytic)


def test_training_step_is_bit_reproducible(big):
    """Two runs of fwd + bwd on the 1 M-cell mesh from the same state give identical losses and identical gradients
    (deterministic CSR reductions, static tile -> CTA maps, ordered partial reductions: no atomics anywhere)."""
    from gen_fvgn_steady_b200.FVMmodel.importer import NNmodel
    from gen_fvgn_steady_b200.utils.get_param import params as default_params
    from tests import product_util as PU
    mesh, graphs, plan = big
    dev = plan.pos.device
    p = default_params(net="EPD", message_passing_num=2, dataset_size=1, precision="bf16")
    torch.manual_seed(0)
    model = NNmodel(p).to(dev)
    x0 = graphs[0].x.clone()
    res = []
    for _ in range(2):
        graphs[0].x, graphs[0].norm_uvp, graphs[0].norm_global = x0.clone(), True, True
        model.zero_grad(set_to_none=True)
        model.node_norm.load_state_dict(NNmodel(p).node_norm.state_dict())
        model.node_norm.to(dev)
        out = model(*graphs, is_training=True)
        loss = PU.script_loss(out, p)
        loss.backward()
        res.append((float(loss.detach()), torch.cat([q.grad.reshape(-1) for q in model.parameters()]).clone()))
    assert res[0][0] == res[1][0]
    assert torch.equal(res[0][1], res[1][1])
