"""Minimal attribute bag with the interface of torch_geometric.data.Data that the hot path touches
(attribute access, keys(), to()/cuda()/cpu()).  PyG Data/Batch objects work unchanged as well: the
modules only read attributes."""
import torch


class Data:
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def to(self, device, non_blocking=False):
        for k in self.keys():
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device, non_blocking=non_blocking))
        return self

    def cuda(self, device=None):
        return self.to("cuda" if device is None else device)

    def cpu(self):
        return self.to("cpu")

    def clone(self):
        return Data(**{k: (getattr(self, k).clone() if torch.is_tensor(getattr(self, k)) else getattr(self, k))
                       for k in self.keys()})
