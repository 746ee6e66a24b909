"""Data containers and θ normalisation -- mirror of src/Data.jl, device resident.

DataArrays keeps x and θ on the GPU (sample-contiguous), the train/valid/test split is an int32 index vector on
the device and minibatches are gathered INSIDE the kernels through that index (no host gather per batch).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from .arrays import jl_empty, n_samples, to_jl
from .model import _gen, minmax_rows


def device_permutation(n: int, seed: int, device, first: int = 0, count: Optional[int] = None,
                       base: Optional[torch.Tensor] = None) -> torch.Tensor:
    """int32 device tensor out[j] = base[perm(first + j)] (perm itself without `base`), perm = the seed's pseudo-random
    permutation of 0..n-1, drawn ON the device by dflow_shuffle_indices (stateless Feistel bijection, oracle/shuffle.py):
    replaces randperm (src/Data.jl:112-128) and the DataLoader shuffle (src/Flows.jl:394) without a host round trip."""
    device = torch.device(device)
    count = n - first if count is None else int(count)
    out = torch.empty(count, device=device, dtype=torch.int32)
    if count == 0:
        return out
    if base is not None:
        base = base.to(device=device, dtype=torch.int32).contiguous()
        assert int(base.numel()) >= n
    with torch.cuda.device(device):
        L.check(L.lib().dflow_shuffle_indices(int(seed) & 0xFFFFFFFFFFFFFFFF, int(n), int(first), count,
                                              None if base is None else base.data_ptr(), out.data_ptr(),
                                              torch.cuda.current_stream(device).cuda_stream))
    return out


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
        i1 = int(round(n * f_training))
        i2 = i1 + int(round(n * f_validation))
        if device is not None and torch.device(device).type == "cuda":
            # the permutation is drawn on the device (only its seed comes from the host generator)
            self.seed = int(torch.randint(0, 2**62, (1,), generator=rng or _gen()).item())
            p = device_permutation(n, self.seed, device)
        else:
            p = torch.randperm(n, generator=rng or _gen(), dtype=torch.int64).to(torch.int32)
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
        # data-parallel sharding (shard_): this rank's share of the GLOBAL training / validation sets
        self.is_shard = False
        self.n_training_global = int(self.partition.training.numel())
        self.n_validation_global = int(self.partition.validation.numel())

    def shard_(self, rank: int, world: int) -> "DataArrays":
        """Keep only this rank's contiguous share of the training / validation / testing index lists RESIDENT (SURVEY §8e:
        dataset sharded once; weights replicated).  Call it on every rank after the Flow / NormalizationLayer have been
        built from the full data (x_min / x_max / θ range are global statistics), with the same partition on every rank
        (train_ checks the global counts).  train_ then shuffles shard-locally: minibatch k is the union of every rank's
        k-th local slice -- sampling without replacement stratified by shard."""
        from .flows import shard_range

        _θ_range(self)  # global θ range before the columns go away
        parts = []
        for v in (self.partition.training, self.partition.validation, self.partition.testing):
            lo, hi = shard_range(int(v.numel()), rank, world)
            parts.append(v[lo:hi])
        self.n_training_global = int(self.partition.training.numel())
        self.n_validation_global = int(self.partition.validation.numel())
        cols = torch.cat(parts).to(torch.int64)
        self.x = to_jl(self.x.index_select(1, cols))
        self.θ = to_jl(self.θ.index_select(1, cols))
        o = 0
        new = []
        for v in parts:
            k = int(v.numel())
            new.append(torch.arange(o, o + k, device=self.x.device, dtype=torch.int32))
            o += k
        self.partition = DataPartition(new[0], validation=new[1], testing=new[2])
        self.is_shard = True
        self.shard_rank, self.shard_world = int(rank), int(world)
        return self

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
