"""Data-parallel path on CPU: world_size 2, gloo.  Each rank owns one graph of a 2-graph batch, runs the product's
forward/backward (kernels through the CPU SIMT emulator), all-reduces the flat gradient and the Normalizer increments;
the result must equal the single-process global-batch step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import golden_util as GU

NAME = "synth_ns_batch2_v1"


def _step(meshes, uvps, dp_group):
    from tests import product_util as PU
    from tests.case_inputs import case_state_dict, product_graphs
    from gen_fvgn_steady_b200 import parallel
    from gen_fvgn_steady_b200.FVMmodel.importer import NNmodel
    from gen_fvgn_steady_b200.utils.get_param import params as default_params
    case = GU.CASES[NAME]
    z = GU.load_case(NAME)
    graphs = product_graphs(meshes, uvps, "cpu")
    p = default_params(net=case["net"], dataset_size=case["dataset_size"], precision="fp32")
    model = NNmodel(p)
    model.load_state_dict(case_state_dict(z), strict=True)
    if dp_group:
        model.enable_data_parallel(True)
    flat = parallel.flatten_gradients(model)
    out = model(*graphs, is_training=True)
    loss = PU.script_loss(out, p)
    loss.backward()
    if dp_group:
        parallel.allreduce_gradients(flat, dist.get_world_size())
    return flat.clone(), {k: v.clone() for k, v in model.state_dict().items() if k.startswith("node_norm.")}


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import product_util as PU
    from tests.case_inputs import case_meshes
    from gen_fvgn_steady_b200.parallel import shard_graphs
    PU.use_emulated_kernels()
    torch.set_num_threads(2)
    meshes, uvps, _ = case_meshes(NAME)
    mine = shard_graphs(len(meshes), rank, world)
    flat, norm = _step([meshes[i] for i in mine], [uvps[i] for i in mine], True)
    if rank == 0:
        ret["flat"] = flat
        ret["norm"] = norm
    dist.destroy_process_group()


def test_data_parallel_matches_global_batch():
    from tests import product_util as PU
    from tests.case_inputs import case_meshes
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    PU.use_emulated_kernels()
    try:
        meshes, uvps, _ = case_meshes(NAME)
        ref_flat, ref_norm = _step(meshes, uvps, False)
    finally:
        PU.use_real_kernels()
    got = ret["flat"]
    rel = float((got - ref_flat).norm() / ref_flat.norm())
    assert rel < 2e-5, rel
    for k, v in ref_norm.items():
        assert torch.allclose(ret["norm"][k], v, rtol=1e-6, atol=1e-6), k
