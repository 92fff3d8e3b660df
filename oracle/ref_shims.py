"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Shim layer that lets the *unmodified* reference sources under /root/reference/src
import and run on CPU in this container (SURVEY.md section 8(c) recipe).  It is used by
`oracle/make_golden.py` to generate the committed golden vectors under tests/golden/ and
by the container-only tests that compare `oracle/fvgn_oracle.py` against the real
reference.  /root/reference does not exist on the GPU box, so nothing here is reachable
from `-m gpu` tests, smoke() or bench.py.

What is shimmed (third-party packages that are absent from this image; their sources are
not under /root/reference, so their semantics are restated here -- "parity unpinned at the
third-party boundary"):
  * torch_scatter.scatter / scatter_add / scatter_mean / scatter_max / scatter_min
    (only dim=0 with a 1-D index is used on the hot path: blocks.py:35-51,92-99,
    FVgrad.py:320-325, utilities.py:33, FVInterpolation.py:261-263)
  * torch_geometric.data.Data / Batch / InMemoryDataset, loader.DataLoader,
    nn.global_add_pool / global_mean_pool, utils.to_torch_coo_tensor
  * timm trunc_normal_, plotting / IO libs as MagicMock.
  * `Utils` -> `utils` alias (the snapshot only imports on case-insensitive filesystems).
"""
import importlib
import os
import sys
import types
from unittest.mock import MagicMock

import torch

REF_SRC = os.environ.get("FVGN_REFERENCE_SRC", "/root/reference/src")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "FVMmodel"))


# --------------------------------------------------------------------------- torch_scatter
def _dim_size(index, dim_size, out):
    if out is not None:
        return out.shape[0]
    if dim_size is not None:
        return int(dim_size)
    return int(index.max()) + 1 if index.numel() > 0 else 0


def _scatter_sum(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0 and index.dim() == 1, "shim supports dim=0 with 1-D index only"
    n = _dim_size(index, dim_size, out)
    if out is None:
        out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index, src)


def _scatter_mean(src, index, dim=0, out=None, dim_size=None):
    n = _dim_size(index, dim_size, out)
    s = _scatter_sum(src, index, dim, out, n)
    cnt = torch.bincount(index, minlength=n).clamp(min=1).to(src.dtype)
    return s / cnt.view((-1,) + (1,) * (src.dim() - 1))


def _scatter_minmax(src, index, dim, out, dim_size, kind):
    assert dim == 0 and index.dim() == 1
    n = _dim_size(index, dim_size, out)
    res = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    res = res.scatter_reduce(0, idx, src, reduce=kind, include_self=False)
    return res


def _scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    if reduce in ("sum", "add"):
        return _scatter_sum(src, index, dim, out, dim_size)
    if reduce == "mean":
        return _scatter_mean(src, index, dim, out, dim_size)
    if reduce == "max":
        return _scatter_minmax(src, index, dim, out, dim_size, "amax")
    if reduce == "min":
        return _scatter_minmax(src, index, dim, out, dim_size, "amin")
    raise NotImplementedError(reduce)


def _make_torch_scatter():
    m = types.ModuleType("torch_scatter")
    m.scatter = _scatter
    m.scatter_add = _scatter_sum
    m.scatter_sum = _scatter_sum
    m.scatter_mean = _scatter_mean
    m.scatter_max = lambda *a, **k: (_scatter(*a, reduce="max", **k), None)
    m.scatter_min = lambda *a, **k: (_scatter(*a, reduce="min", **k), None)
    m.scatter_softmax = MagicMock()
    m.scatter_mul = MagicMock()
    return m


# --------------------------------------------------------------------------- torch_geometric
class Data:
    """Minimal attribute bag standing in for torch_geometric.data.Data."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)      # num_nodes goes through the property setter below

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def to(self, device):
        for k in self.keys():
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self

    def cuda(self):
        return self.to("cuda")

    def cpu(self):
        return self.to("cpu")

    def clone(self):
        d = Data()
        for k in self.keys():
            v = getattr(self, k)
            setattr(d, k, v.clone() if torch.is_tensor(v) else v)
        return d

    def __inc__(self, key, value, *a, **k):
        # torch_geometric.data.Data.__inc__: 'batch' and '*index' / 'face' attributes are offset by num_nodes
        if "batch" in key and torch.is_tensor(value):
            return int(value.max()) + 1
        if "index" in key or key == "face":
            return self.num_nodes
        return 0

    def __cat_dim__(self, key, value, *a, **k):
        # torch_geometric.data.Data.__cat_dim__: '*index' / 'face' attributes concatenate along the last dimension
        if "index" in key or key == "face":
            return -1
        return 0

    _N_KEYS = ("x", "feat", "pos", "batch", "node_type", "n_id", "tf")

    @property
    def num_nodes(self):
        """torch_geometric NodeStorage.num_nodes: the explicit value, else the size of the first node-level attribute."""
        if "_num_nodes" in self.__dict__:
            return self.__dict__["_num_nodes"]
        for k, v in self.__dict__.items():
            if torch.is_tensor(v) and k in self._N_KEYS:
                return v.size(self.__cat_dim__(k, v))
        return None

    @num_nodes.setter
    def num_nodes(self, n):
        self.__dict__["_num_nodes"] = n


def collate(data_list):
    """torch_geometric.data.Batch.from_data_list restated for tensor attributes: every attribute is concatenated along
    data.__cat_dim__(key, value) after adding the running sum of data.__inc__(key, value) of the preceding graphs; the
    node -> graph vector `batch` and `num_graphs` are added (torch_geometric/data/collate.py).  The offsets and
    concatenation axes therefore come from the REFERENCE's CustomGraphData.__inc__ / __cat_dim__ (Graph_loader.py:405-480)."""
    first = data_list[0]
    out = first.__class__()
    for key in first.keys():
        vals = [getattr(d, key) for d in data_list]
        if not torch.is_tensor(vals[0]):
            setattr(out, key, vals)
            continue
        dim = first.__cat_dim__(key, vals[0])
        inc, shifted = 0, []
        for d, v in zip(data_list, vals):
            shifted.append(v + inc if (isinstance(inc, int) and inc != 0) or torch.is_tensor(inc) else v)
            step = d.__inc__(key, v)
            inc = inc + (step if step is not None else 0)
        setattr(out, key, torch.stack(shifted, 0) if dim is None else torch.cat(shifted, dim))
    nn = [d.num_nodes for d in data_list]
    if all(n is not None for n in nn):
        out.batch = torch.cat([torch.full((int(n),), i, dtype=torch.long) for i, n in enumerate(nn)])
        out.num_nodes = int(sum(nn))
    out.num_graphs = len(data_list)
    return out


def _global_add_pool(x, batch, size=None):
    n = int(size) if size is not None else int(batch.max()) + 1
    out = torch.zeros((n,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    return out.index_add_(0, batch, x)


def _global_mean_pool(x, batch, size=None):
    n = int(size) if size is not None else int(batch.max()) + 1
    s = _global_add_pool(x, batch, n)
    cnt = torch.bincount(batch, minlength=n).clamp(min=1).to(x.dtype)
    return s / cnt.view(-1, *([1] * (x.dim() - 1)))


def _to_torch_coo_tensor(edge_index, edge_attr=None, size=None):
    n = int(size) if size is not None else int(edge_index.max()) + 1
    if edge_attr is None:
        edge_attr = torch.ones(edge_index.shape[1])
    return torch.sparse_coo_tensor(edge_index, edge_attr, (n, n)).coalesce()


def _make_pyg():
    pyg = types.ModuleType("torch_geometric")
    data = types.ModuleType("torch_geometric.data")
    batch = types.ModuleType("torch_geometric.data.batch")
    loader = types.ModuleType("torch_geometric.loader")
    nn_ = types.ModuleType("torch_geometric.nn")
    utils = types.ModuleType("torch_geometric.utils")
    data.Data = Data
    data.Batch = MagicMock()
    data.InMemoryDataset = type("InMemoryDataset", (), {"__init__": lambda self, *a, **k: None})
    batch.Batch = MagicMock()
    loader.DataLoader = MagicMock()
    nn_.global_add_pool = _global_add_pool
    nn_.global_mean_pool = _global_mean_pool
    nn_.GCNConv = MagicMock()
    for name in ("knn_graph", "knn", "radius", "radius_graph", "knn_interpolate"):
        setattr(nn_, name, MagicMock())
    utils.to_torch_coo_tensor = _to_torch_coo_tensor
    utils.degree = MagicMock()
    pyg.data, pyg.loader, pyg.nn, pyg.utils = data, loader, nn_, utils
    data.batch = batch
    return {
        "torch_geometric": pyg,
        "torch_geometric.data": data,
        "torch_geometric.data.batch": batch,
        "torch_geometric.loader": loader,
        "torch_geometric.nn": nn_,
        "torch_geometric.utils": utils,
    }


_INSTALLED = False


def install():
    """Install all shims and put the reference on sys.path.  Idempotent."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not reference_available():
        raise RuntimeError(f"reference sources not found at {REF_SRC}")
    torch._dynamo.config.disable = True  # @torch.compile on Simulator.forward would trace the shim Data
    sys.modules["torch_scatter"] = _make_torch_scatter()
    sys.modules.update(_make_pyg())
    # timm
    timm = types.ModuleType("timm")
    timm_layers = types.ModuleType("timm.layers")
    timm_models = types.ModuleType("timm.models")
    timm_models_layers = types.ModuleType("timm.models.layers")
    timm_layers.trunc_normal_ = torch.nn.init.trunc_normal_
    timm_models_layers.trunc_normal_ = torch.nn.init.trunc_normal_
    timm.layers, timm.models, timm_models.layers = timm_layers, timm_models, timm_models_layers
    sys.modules.update({"timm": timm, "timm.layers": timm_layers, "timm.models": timm_models,
                        "timm.models.layers": timm_models_layers})
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.tri", "matplotlib.animation",
                 "vtk", "pyvista", "h5py", "natsort", "statsmodels", "statsmodels.api",
                 "statsmodels.nonparametric", "statsmodels.nonparametric.smoothers_lowess",
                 "trimesh", "trimesh.sample", "circle_fit", "tensorboard",
                 "torch.utils.tensorboard"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = MagicMock()
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import utils as ref_utils  # the reference's src/utils package
    if not getattr(ref_utils, "__file__", "").startswith(REF_SRC):
        raise RuntimeError("a different top-level 'utils' package shadows the reference's")
    sys.modules["Utils"] = ref_utils
    for sub in ("utilities", "normalization", "get_param"):
        mod = importlib.import_module(f"utils.{sub}")
        sys.modules[f"Utils.{sub}"] = mod
    _INSTALLED = True


def ref_params(**overrides):
    install()
    from utils import get_param
    p = get_param.params()
    for k, v in overrides.items():
        setattr(p, k, v)
    return p
