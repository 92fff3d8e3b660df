"""torchrun --nproc-per-node 2 tools/check_partition_gpu.py : cell-partition mode on real GPUs (NCCL) against the
single-GPU run of the whole mesh, per precision mode and halo depth: loss, summed parameter gradients."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gen_fvgn_steady_b200 import parallel, partition  # noqa: E402
from gen_fvgn_steady_b200.FVMmodel.importer import NNmodel  # noqa: E402
from gen_fvgn_steady_b200.mesh import synthetic as S  # noqa: E402
from gen_fvgn_steady_b200.mesh.batching import graphs_from_meshes  # noqa: E402
from gen_fvgn_steady_b200.utils.get_param import params as default_params  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
MP = 2
mesh, uvp = S.make_case(40, kind="mixed", bc="channel", seed=3)
uvp = (uvp + 0.3 * np.random.default_rng(5).standard_normal(uvp.shape)).astype(np.float32)


def model_for(prec):
    p = default_params(net="EPD", message_passing_num=MP, dataset_size=1, precision=prec)
    torch.manual_seed(0)
    m = NNmodel(p)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for q in m.parameters():
            if q.dim() == 2:
                q.copy_(torch.randn(q.shape, generator=g) / q.shape[1] ** 0.5)
            else:
                q.add_(0.1 * torch.randn(q.shape, generator=g))
    return m.to(dev), p


def run(prec, halo_layers):
    m, p = model_for(prec)
    if halo_layers is None:
        graphs = graphs_from_meshes([mesh], [uvp], dev)
    else:
        lm, lu, halo = partition.build(mesh, uvp, world, rank, halo_layers=halo_layers, device=dev)
        graphs = graphs_from_meshes([lm], [lu], dev)
        partition.mark_partition(graphs, halo)
        m.enable_cell_partition(True)
    flat = parallel.flatten_gradients(m)
    out = m(*graphs, is_training=True)
    loss = torch.mean(torch.log(p.loss_press * out[3] + p.loss_cont * out[0] + p.loss_mom * out[1] + p.loss_mom * out[2]))
    loss.backward()
    if halo_layers is not None:
        parallel.sum_gradients(flat)
    return float(loss), flat.clone()


for prec in ("fp32", "bf16", "f16"):
    l0, g0 = run(prec, None)
    for hl in (3, 3 * MP + 2):
        l1, g1 = run(prec, hl)
        if rank == 0:
            print(f"{prec:5s} halo {hl:2d}: loss single {l0:.6f} partitioned {l1:.6f}  |dgrad|/|grad| = {float((g1 - g0).norm() / g0.norm()):.3e}",
                  flush=True)
dist.destroy_process_group()
