"""Cell-partition mode on CPU: world_size 2, gloo.  One synthetic mixed-element mesh is
split by gen_fvgn_steady_b200.partition; every rank runs the product's forward/backward on its sub-mesh (kernels through
the CPU SIMT emulator) with the per-GnBlock halo exchange; the losses, the fields on the owned rows and the summed
parameter gradients must equal the single-process run on the whole mesh."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

MESH = dict(n=10, kind="mixed", bc="channel", seed=3)
MP_NUM = 2
_CACHE = {}


def _model(net="EPD"):
    from gen_fvgn_steady_b200.FVMmodel.importer import NNmodel
    from gen_fvgn_steady_b200.utils.get_param import params as default_params
    p = default_params(net=net, message_passing_num=MP_NUM, dataset_size=1, precision="fp32")
    torch.manual_seed(0)
    model = NNmodel(p)
    # unit-scale weights (the 0.02 init makes every rank's output nearly input independent)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for q in model.parameters():
            if q.dim() == 2:
                q.copy_(torch.randn(q.shape, generator=g) / q.shape[1] ** 0.5)
            else:
                q.add_(0.1 * torch.randn(q.shape, generator=g))
    return model, p


def _run(mesh, uvp, halo, net="EPD"):
    from tests import product_util as PU
    from tests.case_inputs import product_graphs
    from gen_fvgn_steady_b200 import parallel, partition
    graphs = product_graphs([mesh], [uvp], "cpu")
    model, p = _model(net)
    if halo is not None:
        partition.mark_partition(graphs, halo)
        model.enable_cell_partition(True)
    flat = parallel.flatten_gradients(model)
    out = model(*graphs, is_training=True)
    loss = PU.script_loss(out, p)
    loss.backward()
    if halo is not None:
        parallel.sum_gradients(flat)
    return [o.detach().clone() for o in out], float(loss), flat.clone()


def _global_case():
    from gen_fvgn_steady_b200.mesh import synthetic as S
    mesh, uvp = S.make_case(**MESH)
    rng = np.random.default_rng(5)
    uvp = (uvp + 0.3 * rng.standard_normal(uvp.shape)).astype(np.float32)  # a non-trivial field
    return mesh, uvp


def _worker(rank, world, port, ret, halo_layers=3, net="EPD"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import product_util as PU
    from gen_fvgn_steady_b200 import partition
    PU.use_emulated_kernels()
    torch.set_num_threads(2)
    mesh, uvp = _global_case()
    lmesh, luvp, halo = partition.build(mesh, uvp, world, rank, halo_layers=halo_layers)
    out, loss, flat = _run(lmesh, luvp, halo, net)
    n_own = halo.rows["node"]["n_owned"]
    ret[rank] = dict(losses=[o for o in out[:4]], loss=loss, flat=flat, node_gid=halo.rows["node"]["gid"][:n_own].clone(),
                     uvp_node=out[4][:n_own].clone(), cell_gid=halo.cell_gid[:halo.n_owned_cells].clone(),
                     uvp_cell=out[5][:halo.n_owned_cells].clone(), n_local=halo.rows["node"]["n_local"], n_own=n_own)
    dist.destroy_process_group()


def test_partition_structure():
    """Every node / face / cell is owned exactly once; the send and receive lists of the two ranks pair up."""
    from gen_fvgn_steady_b200 import partition
    mesh, uvp = _global_case()
    N, E = mesh["node|pos"].shape[0], mesh["face|face_node"].shape[1]
    C = mesh["cell|centroid"].shape[0]
    halos = [partition.build(mesh, uvp, 2, r)[2] for r in range(2)]
    for kind, total in (("node", N), ("edge", E)):
        owned = torch.cat([h.rows[kind]["gid"][:h.rows[kind]["n_owned"]] for h in halos])
        assert owned.numel() == total and torch.unique(owned).numel() == total
        for r in range(2):
            q = 1 - r
            send = halos[r].rows[kind]["gid"][halos[r].rows[kind]["send"][q]]
            st, cnt = halos[q].rows[kind]["recv"][r]
            recv = halos[q].rows[kind]["gid"][st:st + cnt]
            assert torch.equal(send, recv)
    cells = torch.cat([h.cell_gid[:h.n_owned_cells] for h in halos])
    assert cells.numel() == C and torch.unique(cells).numel() == C
    # balanced bisection
    assert abs(halos[0].n_owned_cells - halos[1].n_owned_cells) <= 1


@pytest.mark.parametrize("halo_layers,net", [(3, "EPD"), (3 * MP_NUM + 2, "EPD"), (3, "TransFVGN_v1")])
def test_cell_partition_matches_single_process(halo_layers, net):
    """halo_layers = 3: ghost refresh after every GnBlock; 3 G + 2: no latent exchange at all (redundant halo compute);
    TransFVGN_v1: the Transolver slice tokens are summed over the ranks' owned rows."""
    from tests import product_util as PU
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret, halo_layers, net), nprocs=2, join=True)
    if net not in _CACHE:  # the single-process run of the whole mesh (shared by the parametrised cases)
        PU.use_emulated_kernels()
        try:
            mesh, uvp = _global_case()
            _CACHE[net] = _run(mesh, uvp, None, net)
        finally:
            PU.use_real_kernels()
    out, loss, flat = _CACHE[net]
    assert ret[0]["n_local"] > ret[0]["n_own"]  # there really is a halo
    from gen_fvgn_steady_b200.partition import HaloPlan
    hp = HaloPlan(0, 2)
    hp.layers = halo_layers
    assert any(hp.wants_exchange(i, MP_NUM) for i in range(MP_NUM)) == (halo_layers == 3)
    for r in range(2):
        for a, b in zip(ret[r]["losses"], out[:4]):
            assert torch.allclose(a, b, rtol=2e-5, atol=1e-7), (r, a, b)
        assert abs(ret[r]["loss"] - loss) < 2e-5 * max(1.0, abs(loss))
        un = out[4][ret[r]["node_gid"]]
        assert float((ret[r]["uvp_node"] - un).abs().max()) < 2e-5 * float(un.abs().max())
        uc = out[5][ret[r]["cell_gid"]]
        assert float((ret[r]["uvp_cell"] - uc).abs().max()) < 2e-5 * float(uc.abs().max())
    rel = float((ret[0]["flat"] - flat).norm() / flat.norm())
    assert rel < 5e-5, rel
    assert torch.equal(ret[0]["flat"], ret[1]["flat"])
