"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL/NVLink).

Data-parallel mode: graphs of a batch are block-diagonal and independent, the script-level
loss is a mean over graphs (pre_train_Adam.py:184), so rank r owns the graphs {b : b mod R = r}; the only exchange is
ONE all-reduce of the flat fp32 gradient per step (1.18 M parameters = 4.7 MB for TransFVGN_v2) and, while the
Normalizer is still accumulating, an all-reduce of its three accumulators (normalization.py:55-66).

Cell-partition mode (section 8(e).2, gen_fvgn_steady_b200.partition): one mesh split over the ranks; after every GnBlock
the ghost rows of the node and edge latents are refreshed from their owners (HaloExchangeFn; its backward returns the
ghost rows' gradients to the owners and adds them there), the per-graph sums of the hot path are all-reduced, and the
parameter gradients are SUMMED (every rank back-propagates the same global loss through its own sub-mesh)."""
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


# ----------------------------------------------------------------------------------------------- halo exchange
def _exchange(pairs, group, reverse):
    """pairs = [(tensor, rows), ...] exchanged in ONE grouped NCCL call.
    forward : ghost rows <- the owners' rows (contiguous receives, index-gather sends), in place.
    reverse : ghost rows' values are sent back to the owners and ADDED to the rows they mirror; ghost rows <- 0."""
    ops_, keep = [], []
    for x, rows in pairs:
        peers = sorted(set(rows["send"].keys()) | set(rows["recv"].keys()))
        for q in peers:
            if not reverse:
                if q in rows["send"]:
                    buf = x.index_select(0, rows["send"][q])
                    keep.append(buf)
                    ops_.append(dist.P2POp(dist.isend, buf, q, group=group))
                if q in rows["recv"]:
                    st, cnt = rows["recv"][q]
                    ops_.append(dist.P2POp(dist.irecv, x[st:st + cnt], q, group=group))
            else:
                if q in rows["recv"]:
                    st, cnt = rows["recv"][q]
                    ops_.append(dist.P2POp(dist.isend, x[st:st + cnt], q, group=group))
                if q in rows["send"]:
                    buf = torch.empty((rows["send"][q].numel(),) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
                    keep.append((x, rows, q, buf))
                    ops_.append(dist.P2POp(dist.irecv, buf, q, group=group))
    if ops_:
        for w in dist.batch_isend_irecv(ops_):
            w.wait()
    if reverse:
        for x, rows, q, buf in keep:  # ascending peer order per tensor: deterministic
            x.index_add_(0, rows["send"][q], buf)
        for x, rows in pairs:
            if rows["n_local"] > rows["n_owned"]:
                x[rows["n_owned"]:].zero_()


class HaloExchangeFn(torch.autograd.Function):
    """(x, e)[ghost rows] <- owner values (in place, one grouped exchange for both latents).  backward: the gradients of
    the owned rows receive (+=) the ghosts' gradients from every rank that mirrors them, the ghost rows' gradients are 0
    (their pre-exchange values are dead)."""

    @staticmethod
    def forward(ctx, x, e, rows_x, rows_e, group):
        ctx.rows, ctx.group = (rows_x, rows_e), group
        ctx.set_materialize_grads(False)
        _exchange([(x, rows_x), (e, rows_e)], group, reverse=False)
        ctx.mark_dirty(x, e)
        return x, e

    @staticmethod
    def backward(ctx, gx, ge):
        rows_x, rows_e = ctx.rows
        pairs = []
        if gx is not None:
            gx = gx.contiguous().clone()
            pairs.append((gx, rows_x))
        if ge is not None:
            ge = ge.contiguous().clone()
            pairs.append((ge, rows_e))
        # both ranks of a pair must post matching operations: a missing gradient is an all-zero one
        if gx is None:
            gx = torch.zeros((rows_x["n_local"], 128), device=ge.device)
            pairs.insert(0, (gx, rows_x))
        if ge is None:
            ge = torch.zeros((rows_e["n_local"], 128), device=gx.device)
            pairs.append((ge, rows_e))
        _exchange(pairs, ctx.group, reverse=True)
        return gx, ge, None, None, None


def no_ghost_refresh(graph, n_blocks):
    """True for a whole mesh and for a partitioned sub-mesh whose halo is deep enough (3 n_blocks + 2 layers) that no GnBlock
    is followed by a ghost refresh: the models then keep their latent streams in 16 bit and drop the last edge latent
    (halo_refresh exchanges the fp32 rows)."""
    halo = getattr(graph, "_fvgn_halo", None)
    return halo is None or halo.world == 1 or not any(halo.wants_exchange(i, n_blocks) for i in range(n_blocks))


def halo_refresh(graph, block_index=0, n_blocks=1, group=None):
    """Refresh the ghost rows of graph.x / graph.edge_attr (and of their bf16 shadows) after GnBlock `block_index` of
    `n_blocks`.  No-op when the graph is not a partitioned sub-mesh, or when the halo is deep enough for this block's
    result to be exact without it (partition.HaloPlan.wants_exchange: a halo of 3k layers needs an exchange only after
    every k-th block; 3 n_blocks + 2 layers need none at all)."""
    halo = getattr(graph, "_fvgn_halo", None)
    if halo is None or halo.world == 1 or not halo.wants_exchange(block_index, n_blocks):
        return graph
    x, e = graph.x, graph.edge_attr
    cx, ce = getattr(graph, "_xh", None), getattr(graph, "_eh", None)
    x2, e2 = HaloExchangeFn.apply(x, e, halo.rows["node"], halo.rows["edge"], group)
    graph.x, graph.edge_attr = x2, e2
    for t, t2, cached, key, kind in ((x, x2, cx, "_xh", "node"), (e, e2, ce, "_eh", "edge")):
        rows = halo.rows[kind]
        sh = cached[1] if (cached is not None and cached[0] is t) else None
        if sh is not None and rows["n_local"] > rows["n_owned"]:
            sh[rows["n_owned"]:] = t2.detach()[rows["n_owned"]:].to(sh.dtype)
        setattr(graph, key, (t2, sh))
    return graph


def allreduce_sum_(t, group=None):
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def sum_gradients(flat, group=None):
    """Cell-partition mode: every rank differentiates the same global loss through its sub-mesh -> gradients ADD."""
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)


class AllReduceSumFn(torch.autograd.Function):
    """y = sum over ranks of x, differentiable: every rank's loss depends on the summed quantity, so the gradient of a
    rank's partial is the SUM of all ranks' gradients of y (used for the Transolver slice tokens in cell-partition mode)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        y = x.contiguous().clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None
