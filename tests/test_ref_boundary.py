"""Container-only (needs /root/reference): the drop-in boundary of SURVEY.md section 8(b), proven with the reference's OWN
loader classes and step loop.

1. The five batch objects made by the reference's Graph_loader dataset classes + the PyG collate rule (driven by the
   reference's CustomGraphData.__inc__ / __cat_dim__, B >= 2, graphs sampled out of order) carry exactly the tensors the
   product's own batching (mesh/batching.py, used on the GPU box where PyG does not exist) produces.
2. With sys.modules["FVMmodel"] aliased to the product's mirror, the reference's training loop
   (solve_with_grad_GPU.py:133-181: backup of graph_node.x, inner iterations, script loss, backward, Adam) runs unchanged on
   those reference-made batch objects through the product (kernels emulated on the CPU) and tracks the UNMODIFIED reference
   model stepping on the same objects."""
import copy
import importlib
import sys

import numpy as np
import pytest
import torch

from oracle import ref_shims
from tests import golden_util as GU
from tests import product_util as PU

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="reference sources not present (GPU box)")


def _cases():
    from oracle.make_golden import _to_t
    from gen_fvgn_steady_b200.mesh import synthetic as S
    cs = [S.make_case(6, kind="quad", bc="channel", seed=5, physics=dict(mean_u=1.2, mu=0.015, dt=0.3)),
          S.make_case(5, kind="mixed", bc="cavity", seed=6),
          S.make_case(4, kind="tri", bc="cavity", seed=2, physics=dict(mean_u=0.8, source=0.3))]
    return cs, [_to_t(c[0]) for c in cs], [torch.from_numpy(GU.perturbed_field(c[1], i)) for i, c in enumerate(cs)]


def test_reference_loader_batch_equals_product_batching():
    from oracle import ref_harness as H
    from gen_fvgn_steady_b200.mesh.batching import graphs_from_meshes
    cs, meshes, uvps = _cases()
    ids = [2, 0, 1]
    ref = H.ref_loader_graphs(meshes, uvps, ids)
    mine = graphs_from_meshes([cs[i][0] for i in ids], [uvps[i] for i in ids], "cpu")
    skip = {("cell", "x"),     # torch.empty in the reference (Graph_loader.py:715): never read
            ("index", "x")}    # the sampled dataset indices in the reference, arange(B) here: never read by the model
    for a, b, name in zip(ref, mine, ("node", "node_x", "edge", "cell", "index")):
        assert a.num_graphs == b.num_graphs == 3
        for k in b.keys():
            vb = getattr(b, k)
            if torch.is_tensor(vb) and (name, k) not in skip:
                va = getattr(a, k)
                assert va.shape == vb.shape and torch.equal(va.to(vb.dtype), vb), (name, k)


@pytest.fixture
def product_as_FVMmodel():
    """sys.modules['FVMmodel*'] -> the product's mirror package, as a maintainer dropping it in would arrange by path."""
    import gen_fvgn_steady_b200.FVMmodel.importer  # noqa: F401
    import gen_fvgn_steady_b200.FVMmodel.Models.TransFVGN.TransFVGN_v1  # noqa: F401
    import gen_fvgn_steady_b200.FVMmodel.Models.TransFVGN.TransFVGN_v2  # noqa: F401
    saved = {k: v for k, v in sys.modules.items() if k == "FVMmodel" or k.startswith("FVMmodel.")}
    for k in saved:
        del sys.modules[k]
    prefix = "gen_fvgn_steady_b200.FVMmodel"
    for k, v in list(sys.modules.items()):
        if k == prefix or k.startswith(prefix + "."):
            sys.modules["FVMmodel" + k[len(prefix):]] = v
    PU.use_emulated_kernels()
    try:
        yield
    finally:
        PU.use_real_kernels()
        for k in [k for k in sys.modules if k == "FVMmodel" or k.startswith("FVMmodel.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def _solve_loop(model, graphs, params, inner_steps, lr):
    """solve_with_grad_GPU.py:133-181 (one epoch), verbatim structure: the loop does not know which NNmodel it drives."""
    graph_node, graph_node_x, graph_edge, graph_cell, graph_Index = graphs
    optimizer = torch.optim.Adam(model.parameters(), lr=lr)
    model.train()
    uvp_pde_theta_backup = graph_node.x.clone()
    losses = []
    for i_iter in range(inner_steps):
        graph_node.x = uvp_pde_theta_backup
        graph_node.norm_uvp = params.norm_uvp
        graph_node.norm_global = params.norm_global
        optimizer.zero_grad()
        loss_cont, loss_mom_x, loss_mom_y, loss_press, uvp_node_new, uvp_cell_new = model(
            graph_node=graph_node, graph_node_x=graph_node_x, graph_edge=graph_edge, graph_cell=graph_cell,
            graph_Index=graph_Index, is_training=True)
        loss_batch = (params.loss_press * loss_press + params.loss_cont * loss_cont + params.loss_mom * loss_mom_x
                      + params.loss_mom * loss_mom_y)
        loss = torch.mean(torch.log(loss_batch))
        loss.backward()
        optimizer.step()
        losses.append([float(loss.detach())] + [v.detach().reshape(-1).tolist() for v in (loss_cont, loss_mom_x, loss_mom_y, loss_press)])
    return losses, uvp_node_new.detach().clone()


def test_reference_step_loop_runs_on_the_product_unchanged(product_as_FVMmodel):
    from oracle import ref_harness as H
    cs, meshes, uvps = _cases()
    ids = [1, 0]
    params = ref_shims.ref_params(net="TransFVGN_v1", message_passing_num=1, dataset_size=1)
    # --- the unmodified reference on its own loader objects
    for k in [k for k in sys.modules if k == "FVMmodel" or k.startswith("FVMmodel.")]:
        del sys.modules[k]                           # the fixture's aliases: the reference model must be the real one
    H.seed_all(0)
    import FVMmodel.importer as ref_importer          # resolves to /root/reference/src (first on sys.path)
    assert ref_importer.__file__.startswith(ref_shims.REF_SRC)
    ref_model = ref_importer.NNmodel(params)
    state = copy.deepcopy(ref_model.state_dict())
    ref_losses, ref_uvp = _solve_loop(ref_model, H.ref_loader_graphs(meshes, uvps, ids), params, 3, 1e-3)
    # --- the same loop, the same reference-made objects, FVMmodel = the product
    for k in [k for k in sys.modules if k == "FVMmodel" or k.startswith("FVMmodel.")]:
        del sys.modules[k]
    prefix = "gen_fvgn_steady_b200.FVMmodel"
    for k, v in list(sys.modules.items()):
        if k == prefix or k.startswith(prefix + "."):
            sys.modules["FVMmodel" + k[len(prefix):]] = v
    NNmodel = importlib.import_module("FVMmodel.importer").NNmodel
    assert NNmodel.__module__.startswith("gen_fvgn_steady_b200")
    params.precision = "fp32"
    model = NNmodel(params)
    missing = model.load_state_dict(state, strict=True)   # the reference's checkpoint loads as is
    assert not missing.missing_keys and not missing.unexpected_keys
    losses, uvp = _solve_loop(model, H.ref_loader_graphs(meshes, uvps, ids), params, 3, 1e-3)
    for it, (a, b) in enumerate(zip(losses, ref_losses)):
        assert abs(a[0] - b[0]) <= 2e-4 * max(abs(b[0]), 1.0), (it, a[0], b[0])
        for x, y in zip(a[1:], b[1:]):
            assert np.allclose(x, y, rtol=2e-3, atol=1e-9), (it, x, y)
    assert GU.rel_err(uvp, ref_uvp) < 2e-3
    # the loop restores `graph_node.x = uvp_pde_theta_backup`, the tensor the previous forward normalised IN PLACE
    # (importer.py:121,127): from the second inner iteration on both models see the normalised features -- the product
    # reproduces that observable behaviour of the reference, which is why the two trajectories above agree at all
    assert abs(ref_losses[1][0] - ref_losses[0][0]) > 1e-3
