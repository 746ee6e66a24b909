"""Data containers and θ normalisation -- mirror of src/Data.jl, device resident.

DataArrays keeps x and θ on the GPU (sample-contiguous), the train/valid/test split is an int32 index vector on
the device and minibatches are gathered INSIDE the kernels through that index (no host gather per batch).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from .arrays import jl_empty, n_samples, to_jl
from .model import _gen, minmax_rows


def dflt_θ(*args) -> torch.Tensor:
    """dflt_θ([T,] dims...) / dflt_θ(x): a 0 x dims array (src/Data.jl:57-65)."""
    if len(args) == 1 and isinstance(args[0], (torch.Tensor, np.ndarray)):
        x = args[0]
        dev = x.device if isinstance(x, torch.Tensor) else None
        return jl_empty((0,) + tuple(x.shape[1:]), device=dev)
    if len(args) == 1 and isinstance(args[0], (tuple, list)):
        args = tuple(args[0])
    return jl_empty((0,) + tuple(int(a) for a in args))


dflt_theta = dflt_θ


class MetaData:
    """MetaData(hash, d, n, θ_min, θ_max), src/Data.jl:75-86."""

    def __init__(self, hash: str, d: int, n: int, θ_min, θ_max):
        self.hash, self.d, self.n = hash, int(d), int(n)
        self.θ_min = np.asarray(θ_min, np.float32).reshape(-1)
        self.θ_max = np.asarray(θ_max, np.float32).reshape(-1)


class DataPartition:
    """DataPartition(n, f_training=0.9, f_validation=0.1, rng): random permutation cut at round(n*f)
    (src/Data.jl:112-128; Julia's round is ties-to-even, like Python's).  Indices are 0-based int32 tensors."""

    def __init__(self, n_or_training, f_training=0.9, f_validation=0.1, rng: Optional[torch.Generator] = None, *,
                 validation=None, testing=None, device=None):
        if validation is not None:
            self.training, self.validation, self.testing = n_or_training, validation, testing
            return
        n = int(n_or_training)
        p = torch.randperm(n, generator=rng or _gen(), dtype=torch.int64).to(torch.int32)
        i1 = int(round(n * f_training))
        i2 = i1 + int(round(n * f_validation))
        if device is not None:
            p = p.to(device)
        self.training, self.validation, self.testing = p[:i1].contiguous(), p[i1:i2].contiguous(), p[i2:n].contiguous()


class DataArrays:
    """DataArrays(x, θ=dflt_θ(x); f_training=0.9, f_validation=0.1, rng), src/Data.jl:130-170.
    Partitioned along the second axis only (src/Data.jl:167)."""

    def __init__(self, x, θ=None, f_training=0.9, f_validation=0.1, rng: Optional[torch.Generator] = None, device=None):
        if device is None and isinstance(x, torch.Tensor):
            device = x.device
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        x = to_jl(x, device)
        θ = dflt_θ(x) if θ is None else to_jl(θ, device)
        assert x.dim() >= 2, "data must be an array of size (d, i1, ...) at least"  # src/Data.jl:164
        assert tuple(x.shape[1:]) == tuple(θ.shape[1:]), \
            "x and θ must have the same size -- except for the first dimension"  # src/Data.jl:165
        self.x, self.θ = x, θ
        self.partition = DataPartition(int(x.shape[1]), f_training, f_validation, rng, device=device)
        self._θ_range: Optional[Tuple[np.ndarray, np.ndarray]] = None

    theta = property(lambda self: self.θ)

    def summarize(self) -> str:
        n2 = int(self.x.shape[1])
        return (f"Data with size {tuple(self.x.shape)} and parameters / conditions with size {tuple(self.θ.shape)}.\n"
                f"-> f_training = {len(self.partition.training) / n2}, f_validation = {len(self.partition.validation) / n2}.")


def number_dimensions(data: DataArrays) -> int:
    return int(data.x.shape[0])


def number_conditions(data: DataArrays) -> int:
    return int(data.θ.shape[0])


def _θ_range(data: DataArrays):
    if data._θ_range is None:
        if number_conditions(data) == 0:
            data._θ_range = (np.zeros(0, np.float32), np.zeros(0, np.float32))
        else:
            data._θ_range = minmax_rows(data.θ)
    return data._θ_range


def minimum_θ(obj):
    """minimum_θ(data) = vec(minimum(data.θ, dims=2:N)) (src/Data.jl:182); minimum_θ(metadata) (src/Data.jl:89)."""
    return obj.θ_min if isinstance(obj, MetaData) else _θ_range(obj)[0]


def maximum_θ(obj):
    return obj.θ_max if isinstance(obj, MetaData) else _θ_range(obj)[1]


def _select2(a: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """selectdim(a, 2, idx) materialised column-major (src/Data.jl:185-187)."""
    return to_jl(a.index_select(1, idx.to(torch.int64)))


def training_data(data: DataArrays):
    return _select2(data.x, data.partition.training), _select2(data.θ, data.partition.training)


def validation_data(data: DataArrays):
    return _select2(data.x, data.partition.validation), _select2(data.θ, data.partition.validation)


def testing_data(data: DataArrays):
    return _select2(data.x, data.partition.testing), _select2(data.θ, data.partition.testing)


def normalize_input(x, x_min, x_max) -> torch.Tensor:
    """normalize_input(x, x_min, x_max) = (x - x_min) / (x_max - x_min), rows with zero range -> 0
    (src/Data.jl:213-218).  Host-side convenience (torch ops); the kernels fold this into their θ load."""
    x = to_jl(x)
    shp = (-1,) + (1,) * (x.dim() - 1)
    mn = torch.as_tensor(np.asarray(x_min, np.float32), device=x.device).reshape(shp)
    df = torch.as_tensor(np.asarray(x_max, np.float32) - np.asarray(x_min, np.float32), device=x.device).reshape(shp)
    y = (x - mn) / df
    y = torch.where(df == 0, torch.zeros_like(y), y)
    return to_jl(y)


def resize_output(y, x_min, x_max) -> torch.Tensor:
    """resize_output, src/Data.jl:231."""
    y = to_jl(y)
    shp = (-1,) + (1,) * (y.dim() - 1)
    mn = torch.as_tensor(np.asarray(x_min, np.float32), device=y.device).reshape(shp)
    mx = torch.as_tensor(np.asarray(x_max, np.float32), device=y.device).reshape(shp)
    return to_jl((mx - mn) * y + mn)


def normalized_training_data(data: DataArrays, metadata: MetaData):
    """src/Data.jl:189-193."""
    x, θ = training_data(data)
    return x, normalize_input(θ, metadata.θ_min, metadata.θ_max)


def normalized_validation_data(data: DataArrays, metadata: MetaData):
    """src/Data.jl:195-199."""
    x, θ = validation_data(data)
    return x, normalize_input(θ, metadata.θ_min, metadata.θ_max)
