"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL/NVLink).

Data-parallel mode: graphs of a batch are block-diagonal and independent, the script-level
loss is a mean over graphs (pre_train_Adam.py:184), so rank r owns the graphs {b : b mod R = r}; the only exchange is
ONE all-reduce of the flat fp32 gradient per step (1.18 M parameters = 4.7 MB for TransFVGN_v2) and, while the
Normalizer is still accumulating, an all-reduce of its three accumulators (normalization.py:55-66)."""
import torch
import torch.distributed as dist


def flatten_gradients(model):
    """Make every parameter's .grad a view into one flat fp32 buffer (autograd then accumulates in place)."""
    params = [p for p in model.parameters() if p.requires_grad]
    total = sum(p.numel() for p in params)
    flat = torch.zeros(total, dtype=torch.float32, device=params[0].device)
    off = 0
    for p in params:
        n = p.numel()
        p.grad = flat[off:off + n].view_as(p)
        off += n
    return flat


def allreduce_gradients(flat, world_size, group=None):
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world_size)


def shard_graphs(num_graphs, rank, world_size):
    """Graph ids owned by `rank` (round-robin, same rule on every rank)."""
    return list(range(rank, num_graphs, world_size))


def allreduce_normalizer(normalizer, pending, group=None):
    """Sum the Normalizer increments of this step over ranks so every rank holds the global-batch statistics.
    pending = (data_sum, squared_sum, count) as produced by NNmodel.update_x_attr on the local shard."""
    s, ss, cnt = pending
    buf = torch.cat([s.reshape(-1), ss.reshape(-1), torch.tensor([float(cnt)], device=s.device)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    n = s.numel()
    return buf[:n], buf[n:2 * n], float(buf[-1])
