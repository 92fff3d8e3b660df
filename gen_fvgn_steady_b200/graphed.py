"""Whole-step CUDA graph for the resident-batch regime.

src/solve_with_grad_GPU.py:133-208 and the inner loop of src/pre_train_Adam.py:122-191 iterate on ONE batch whose
topology and device buffers do not change: NNmodel.forward + the script loss + backward + Adam are then ~600 kernel
launches with fixed arguments.  On the reference's own example meshes (10 k - 30 k cells) such a step is launch / Python
bound (5.5 ms eager on a B200, independent of the mesh size), so the step is captured once and replayed.

    gstep = GraphedTrainStep(model, optimizer, graphs, loss_fn)      # optimizer: Adam(..., fused=True, capturable=True)
    for it in range(n): loss = gstep.step()                          # optional: gstep.step(new_x) refreshes graph_node.x

Every kernel of the path enqueues on torch's current stream and never allocates or synchronises (include/fvgn_b200.h),
so the capture needs nothing special; torch's caching allocator serves the transient buffers from the graph's pool."""
import torch


class GraphedTrainStep:
    def __init__(self, model, optimizer, graphs, loss_fn, warmup=3, freeze_normalizer=True):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedTrainStep needs a CUDA device: this package has no CPU path")
        self.model, self.optimizer, self.graphs, self.loss_fn = model, optimizer, graphs, loss_fn
        gn = graphs[0]
        self.x_static = gn.x.detach().clone()          # raw [N,3] / [N,12] node features, read by the captured prologue
        for g in optimizer.param_groups:
            if not g.get("capturable", False):
                raise ValueError("GraphedTrainStep: build the optimizer with capturable=True (e.g. Adam(fused=True, capturable=True))")
        self._norm_flags = (getattr(gn, "norm_uvp", True), getattr(gn, "norm_global", True))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        norm = getattr(model, "node_norm", None)
        if norm is not None and norm.wants_accumulation():
            if not freeze_normalizer:
                raise RuntimeError("the Normalizer is still accumulating; capture after it froze or pass freeze_normalizer=True")
            norm.max_accumulations = float(norm.num_accumulations)  # statistics stay as accumulated so far
        if norm is not None:
            norm._n_acc_host = float(norm.num_accumulations)        # host mirror fixed now: no device sync inside the capture
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=False) if self._flat_grads() else optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.out, self.loss = self._body()
        self.replays = 0

    def _flat_grads(self):
        return all(p.grad is not None for g in self.optimizer.param_groups for p in g["params"] if p.requires_grad)

    def _prepare(self):
        gn = self.graphs[0]
        gn.x = self.x_static
        gn.norm_uvp, gn.norm_global = self._norm_flags

    def _body(self):
        self._prepare()
        for g in self.optimizer.param_groups:
            for p in g["params"]:
                if p.grad is not None:
                    p.grad.zero_()
        out = self.model(*self.graphs, is_training=True)
        loss = self.loss_fn(out)
        loss.backward()
        self.optimizer.step()
        return out, loss

    def _eager_step(self):
        return self._body()

    def step(self, x=None):
        """One captured train step; x (optional) replaces the node features first.  Returns the (static) loss tensor."""
        if x is not None:
            self.x_static.copy_(x, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return self.loss
