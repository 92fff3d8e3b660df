"""Online mean/std Normalizer with the reference's buffers (src/utils/normalization.py:4-85) so reference
checkpoints load.  The per-row normalisation itself runs inside the fused prologue kernel; this module
keeps the running statistics (tiny [size] tensors) and hands (mean, std) to it."""
import torch
from torch import nn


class Normalizer(nn.Module):
    def __init__(self, size, max_accumulations=10 ** 7, epsilon=1e-8, device=None):
        super().__init__()
        self.max_accumulations = max_accumulations
        self.epsilon = epsilon
        self.register_buffer("acc_count", torch.tensor(1.0, dtype=torch.float32, device=device))
        self.register_buffer("num_accumulations", torch.tensor(1.0, dtype=torch.float32, device=device))
        self.register_buffer("acc_sum", torch.zeros(size, dtype=torch.float32, device=device))
        self.register_buffer("acc_sum_squared", torch.zeros(size, dtype=torch.float32, device=device))
        self._n_acc_host = None  # host mirror of num_accumulations (avoids a device sync per step)

    def wants_accumulation(self):
        if self._n_acc_host is None:
            self._n_acc_host = float(self.num_accumulations)
        return self._n_acc_host < self.max_accumulations

    def accumulate(self, data_sum, squared_sum, count):
        """normalization.py:55-66 with the column sums computed by the caller."""
        self.acc_sum += data_sum.to(self.acc_sum.dtype)
        self.acc_sum_squared += squared_sum.to(self.acc_sum_squared.dtype)
        self.acc_count += float(count)
        self.num_accumulations += 1
        self._n_acc_host = (self._n_acc_host if self._n_acc_host is not None else 0.0) + 1.0

    def mean(self):
        safe = torch.clamp(self.acc_count, min=1.0)
        return self.acc_sum / safe

    def std(self):
        safe = torch.clamp(self.acc_count, min=1.0)
        # the reference takes sqrt(E[x^2] - mean^2) directly (normalization.py:80-83); for a constant feature fp32
        # rounding can push that a few ulp below zero and the reference then propagates NaN -- clamp at 0 instead.
        std = torch.sqrt(torch.clamp(self.acc_sum_squared / safe - self.mean() ** 2, min=0.0))
        return torch.where(std < self.epsilon, torch.ones_like(std), std)

    def _load_from_state_dict(self, *args, **kwargs):
        self._n_acc_host = None
        return super()._load_from_state_dict(*args, **kwargs)
