"""Whole-step CUDA graph for the resident-batch regime.

src/solve_with_grad_GPU.py:133-208 and the inner loop of src/pre_train_Adam.py:122-191 iterate on ONE batch whose
topology and device buffers do not change: NNmodel.forward + the script loss + backward + Adam are then ~600 kernel
launches with fixed arguments.  On the reference's own example meshes (10 k - 30 k cells) such a step is launch / Python
bound (5.5 ms eager on a B200, independent of the mesh size), so the step is captured once and replayed.

    gstep = GraphedTrainStep(model, optimizer, graphs, loss_fn)      # optimizer: Adam(..., fused=True, capturable=True)
    for it in range(n): loss = gstep.step()                          # optional: gstep.step(new_x) refreshes graph_node.x
    gstep.close()                                                    # restores the Normalizer's accumulation window

Construction does NOT advance training: the eager warm-up steps the capture needs run on a snapshot of the model /
optimizer state, which is restored before the capture.  A Normalizer that is still accumulating cannot be captured (its
host-side counter decides which kernels run): pass freeze_normalizer=True to freeze its statistics for the lifetime of
this object (a warning says so; close() re-opens the window) or capture after `dataset_size` accumulations.
The packed 16-bit weight images (ops.PackedWeights) are rebuilt by every forward, captured or eager, so replays and
eager forwards between / after them always see the current weights.

Every kernel of the path enqueues on torch's current stream and never allocates or synchronises (include/fvgn_b200.h),
so the capture needs nothing special; torch's caching allocator serves the transient buffers from the graph's pool."""
import copy
import warnings

import torch


class GraphedTrainStep:
    def __init__(self, model, optimizer, graphs, loss_fn, warmup=3, freeze_normalizer=False, post_backward=None):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedTrainStep needs a CUDA device: this package has no CPU path")
        self.model, self.optimizer, self.graphs, self.loss_fn = model, optimizer, graphs, loss_fn
        # multi-GPU: called between backward and the optimizer step inside the captured body (gradient all-reduce; NCCL
        # collectives are captured into the graph like kernels)
        self.post_backward = post_backward
        gn = graphs[0]
        self.x_static = gn.x.detach().clone()          # raw [N,3] / [N,12] node features, read by the captured prologue
        for g in optimizer.param_groups:
            if not g.get("capturable", False):
                raise ValueError("GraphedTrainStep: build the optimizer with capturable=True (e.g. Adam(fused=True, capturable=True))")
        self._norm_flags = (getattr(gn, "norm_uvp", True), getattr(gn, "norm_global", True))
        norm = getattr(model, "node_norm", None)
        self._norm, self._norm_max = norm, None
        # warm-up (allocator pools, lazy initialisation) on a snapshot: construction leaves model and optimizer untouched
        model_state = {k: v.detach().clone() for k, v in model.state_dict().items()}
        opt_state = {p: {k: (v.detach().clone() if torch.is_tensor(v) else copy.deepcopy(v)) for k, v in st.items()}
                     for p, st in optimizer.state.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        with torch.no_grad():
            for k, v in model.state_dict().items():
                v.copy_(model_state[k])
        # optimizer state: values back to the snapshot IN PLACE (a fresh optimizer: zeros), the tensors stay allocated --
        # lazily creating them inside the capture would record their zero-initialisation into every replay
        with torch.no_grad():
            for p, st in optimizer.state.items():
                old = opt_state.get(p)
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if old is not None and k in old:
                            v.copy_(old[k])
                        else:
                            v.zero_()
                    elif old is not None and k in old:
                        st[k] = old[k]
        if norm is not None:
            norm._n_acc_host = None
        if norm is not None and norm.wants_accumulation():
            if not freeze_normalizer:
                raise RuntimeError("GraphedTrainStep: the Normalizer is still accumulating (num_accumulations < dataset_size); "
                                   "capture after it froze or pass freeze_normalizer=True to freeze its statistics while "
                                   "this object is alive")
            warnings.warn("GraphedTrainStep(freeze_normalizer=True): the Normalizer stops accumulating at "
                          f"{float(norm.num_accumulations):.0f} accumulations until close() is called; a checkpoint saved "
                          "meanwhile holds these statistics")
            self._norm_max = norm.max_accumulations
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
        gn.x = self.x_static.clone()   # the forward normalises graph_node.x in place (as the reference does): keep the raw copy
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
        if self.post_backward is not None:
            self.post_backward()
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

    def close(self):
        """Drop the captured graph and give the Normalizer its accumulation window back."""
        if self._norm is not None and self._norm_max is not None:
            self._norm.max_accumulations = self._norm_max
            self._norm_max = None
        self.graph = None
