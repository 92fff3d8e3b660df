"""Device-resident data pool: the counterpart of the reference's Data_Pool + DataLoader
collate + payback cycle (src/Load_mesh/Graph_loader.py:130-152 datapreprocessing, :370-396 payback, :830-1006 loaders;
src/pre_train_Adam.py:129-156,193-198) without the per-step host collate, host -> device copy and plan rebuild.

    pool = DevicePool(meshes, uvps, device)            # converter dictionaries, once
    for step in ...:
        graphs, global_idx = pool.sample(ids)          # five NEW batch objects every step, like a loader
        out = model(*graphs)                            # topology plan found on the batch: no hash, no rebuild, no sync
        pool.payback(out[4], global_idx)                # uvp_node_pool[global_idx] = uvp_new, in place on the device

Every mesh lives on the device once.  The batched index / geometry tensors of an id tuple are assembled on first use
(mesh.batching.graphs_from_meshes: device-side concatenation with the PyG offset rules) and kept; later samples of the
same ids only gather the current node fields from the pool and attach them to fresh batch objects that share the static
tensors, so GraphPlan.of recognises them by their index tensors."""
import torch

from .data import Data
from .mesh.batching import graphs_from_meshes
from .plan import GraphPlan


class DevicePool:
    def __init__(self, meshes, uvps, device, order="2nd", max_cached_batches=64):
        self.device = torch.device(device)
        self.order = order
        self.meshes = list(meshes)
        counts = [int(torch.as_tensor(m["node|pos"]).shape[0]) for m in self.meshes]
        self.node_offset = [0]
        for c in counts:
            self.node_offset.append(self.node_offset[-1] + c)
        # Data_Pool.uvp_node_pool: the current node fields of every graph, one flat [sum N, 3] device tensor
        self.uvp_node_pool = torch.cat([torch.as_tensor(u, dtype=torch.float32) for u in uvps], 0).to(self.device).contiguous()
        self._batches = {}
        self._max = max_cached_batches

    def __len__(self):
        return len(self.meshes)

    def _assemble(self, ids):
        uv = [self.uvp_node_pool[self.node_offset[i]:self.node_offset[i + 1]] for i in ids]
        graphs = graphs_from_meshes([self.meshes[i] for i in ids], uv, self.device)
        gidx = torch.cat([torch.arange(self.node_offset[i], self.node_offset[i + 1], device=self.device) for i in ids])
        theta_cols = graphs[0].x[:, 3:].contiguous()      # theta_PDE[batch], static per id tuple (datapreprocessing :148-150)
        plan = GraphPlan.of(graphs[0], graphs[1], graphs[2], graphs[3], self.order)
        return graphs, gidx, theta_cols, plan

    def sample(self, ids):
        """-> ((graph_node, graph_node_x, graph_edge, graph_cell, graph_Index), global_idx): new batch objects whose node
        features are the pool's CURRENT fields of graphs `ids`."""
        key = tuple(int(i) for i in ids)
        hit = self._batches.get(key)
        if hit is None:
            if len(self._batches) >= self._max:
                self._batches.pop(next(iter(self._batches)))
            hit = self._batches[key] = self._assemble(key)
        graphs, gidx, theta_cols, plan = hit
        fresh = tuple(Data(**{k: getattr(g, k) for k in g.keys()}) for g in graphs)
        gn = fresh[0]
        gn.x = torch.cat([self.uvp_node_pool[gidx], theta_cols], 1)
        gn.norm_uvp, gn.norm_global = True, True
        gn._fvgn_plan = plan
        return fresh, gidx

    def payback(self, uvp_new, global_idx):
        """Data_Pool.payback (Graph_loader.py:370-383): the pool takes the new node fields of the sampled graphs."""
        self.uvp_node_pool[global_idx] = uvp_new.detach().to(self.uvp_node_pool.dtype)
