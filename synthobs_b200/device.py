"""Device-memory plumbing shared by the host-side API modules (torch only for memory/streams)."""

import numpy as np
import torch

from . import _cabi


def device():
    _cabi.require_device()
    return torch.device('cuda', torch.cuda.current_device())


def is_device_tensor(x):
    return isinstance(x, torch.Tensor) and x.is_cuda


def to_dev(x, dtype=torch.float64):
    """NumPy array / CPU tensor / CUDA tensor -> contiguous CUDA tensor of `dtype`.
    Host inputs are copied (H2D); pinned CPU tensors copy asynchronously."""
    dev = device()
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            return x.to(dtype).contiguous()
        return x.to(dev, dtype=dtype, non_blocking=True).contiguous()
    a = np.ascontiguousarray(x)
    t = torch.from_numpy(a)
    return t.to(dev, dtype=dtype, non_blocking=False)


def to_host(t, like_host=True):
    """CUDA tensor -> float64 NumPy array (the reference returns float64 everywhere)."""
    return t.detach().to(torch.float64).cpu().numpy()


def workspace(nbytes):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device())
