"""Julia-shaped arrays on top of torch tensors.

The reference works on Julia column-major arrays of shape `(rows, dims...)` (src/Data.jl:130-170): the first
(fastest) axis is the data / condition index, all trailing axes enumerate samples.  Here such an array is a torch
tensor with the SAME logical shape and column-major strides, so its memory is sample-contiguous -- exactly what
libdflow.so expects (`ptr[k + rows*b]`).  PyTorch is only used for device memory and streams.
"""
from __future__ import annotations

from typing import Sequence, Tuple, Union

import numpy as np
import torch

ArrayLike = Union[torch.Tensor, np.ndarray]


def _rev(nd: int):
    return tuple(range(nd - 1, -1, -1))


def is_colmajor(t: torch.Tensor) -> bool:
    return t.dim() == 0 or t.permute(*_rev(t.dim())).is_contiguous()


def jl_empty(shape: Sequence[int], device=None, dtype=torch.float32) -> torch.Tensor:
    """Uninitialised column-major tensor of Julia shape `shape`."""
    shape = tuple(int(s) for s in shape)
    base = torch.empty(tuple(reversed(shape)), device=device, dtype=dtype)
    return base.permute(*_rev(len(shape)))


def jl_zeros(shape, device=None, dtype=torch.float32) -> torch.Tensor:
    t = jl_empty(shape, device, dtype)
    t.zero_()
    return t


def jl_full(shape, value, device=None, dtype=torch.float32) -> torch.Tensor:
    t = jl_empty(shape, device, dtype)
    t.fill_(value)
    return t


def to_jl(x: ArrayLike, device=None, dtype=torch.float32) -> torch.Tensor:
    """Column-major tensor with the logical shape of `x` on `device` (copies only when needed)."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x.T)).permute(*_rev(x.ndim)) if x.ndim > 0 else torch.from_numpy(x)
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    if x.dtype != dtype:
        x = x.to(dtype)
    if device is not None and x.device != torch.device(device):
        x = x.to(device)
    if not is_colmajor(x):
        nd = x.dim()
        x = x.permute(*_rev(nd)).contiguous().permute(*_rev(nd))
    return x


def flat_view(x: torch.Tensor) -> torch.Tensor:
    """1-D view over the memory of a column-major tensor (for pointer hand-off)."""
    return x.permute(*_rev(x.dim())).reshape(-1)


def tail_shape(x: torch.Tensor) -> Tuple[int, ...]:
    return tuple(x.shape[1:])


def n_samples(x: torch.Tensor) -> int:
    n = 1
    for s in x.shape[1:]:
        n *= int(s)
    return n


def to_numpy(x: torch.Tensor) -> np.ndarray:
    """Logical-shape numpy copy (for tests against the oracle)."""
    return x.detach().cpu().numpy().copy()
