"""Flow / train! / sample / logpdf -- mirror of src/Flows.jl, running on libdflow.so.

`train_` (= `train!`) keeps the dataset resident on the GPU, gathers each minibatch inside the adjoint kernel
through an index vector, runs the Adam update as a kernel and -- when torch.distributed is initialised (one process
per GPU) -- shards every minibatch across ranks and all-reduces the packed gradient (+ loss) with NCCL.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib as L
from .arrays import flat_view, jl_empty, n_samples, to_jl
from .data import (DataArrays, MetaData, device_permutation, maximum_θ, minimum_θ, number_conditions,
                   number_dimensions)
from .model import FlowChain, PackedChain, _gen

# ------------------------------------------------------------------------------------------------------------
# Optimisers.jl mirror (only what train! uses: Optimisers.setup(Optimisers.Adam(η), flow.model))
# ------------------------------------------------------------------------------------------------------------


class Adam:
    """Optimisers.Adam(η=1f-3, β=(0.9, 0.999), ϵ=1e-8)."""

    def __init__(self, η: float = 1e-3, β: Tuple[float, float] = (0.9, 0.999), ϵ: float = 1e-8):
        self.eta, self.beta, self.epsilon = float(η), (float(β[0]), float(β[1])), float(ϵ)


class OptimiserState:
    """What Optimisers.setup returns for a FlowChain: first/second moment buffers in the packed layout + step."""

    def __init__(self, rule: Adam, model: FlowChain):
        self.rule, self.model = rule, model
        self.m: Optional[torch.Tensor] = None
        self.v: Optional[torch.Tensor] = None
        self.t = 0

    def _ensure(self, pc: PackedChain):
        if self.m is None or self.m.device != pc.device or self.m.numel() != max(pc.P, 1):
            self.m = torch.zeros(max(pc.P, 1), device=pc.device, dtype=torch.float32)
            self.v = torch.zeros_like(self.m)


def setup(rule: Adam, model: FlowChain) -> OptimiserState:
    """Optimisers.setup(rule, flow.model) (test/runtests.jl:113)."""
    return OptimiserState(rule, model)


# ------------------------------------------------------------------------------------------------------------
# Flow
# ------------------------------------------------------------------------------------------------------------


class Flow:
    """Flow([base,] model, data), src/Flows.jl:37-122.  The base distribution is MvNormal(0, I_d)
    (src/Flows.jl:114); other bases are not on the fused path."""

    def __init__(self, model: FlowChain, data: DataArrays, train_loss: Optional[List[float]] = None,
                 valid_loss: Optional[List[float]] = None, base: Optional[str] = None):
        if base not in (None, "MvNormal"):
            raise NotImplementedError("only the default MvNormal(0, I) base distribution is fused")
        if not isinstance(model, FlowChain):
            raise TypeError("model must be a FlowChain")
        self.model = model
        self.base = "MvNormal"
        d, n = number_dimensions(data), number_conditions(data)
        self.metadata = MetaData("", d, n, minimum_θ(data), maximum_θ(data))
        self.train_loss: List[float] = [] if train_loss is None else train_loss
        self.valid_loss: List[float] = [] if valid_loss is None else valid_loss
        self.device = data.x.device
        self._packed: Optional[PackedChain] = None

    d = property(lambda self: self.metadata.d)
    n = property(lambda self: self.metadata.n)

    def packed(self) -> PackedChain:
        if self._packed is None:
            pc = PackedChain(self.model._leaves(), self.device if self.device.type == "cuda" else None,
                             self.metadata.θ_min if self.n > 0 else None, self.metadata.θ_max if self.n > 0 else None)
            if (pc.d, pc.n) != (self.d, self.n):
                raise ValueError(f"model is built for (d,n)=({pc.d},{pc.n}) but the data has ({self.d},{self.n})")
            self._packed = pc
            self.model._packed = pc
        self._packed.refresh()
        return self._packed

    def __call__(self, z, θ=None):  # @auto_functor Flow
        from .model import forward

        return forward(self, z, θ)

    def summarize(self) -> str:
        return "- model --------------------\n" + self.model.summarize() + "\n- base distribution --------\nMvNormal"


def training_loss(flow: Flow) -> List[float]:
    return flow.train_loss


def validation_loss(flow: Flow) -> List[float]:
    return flow.valid_loss


def predict(flow: Flow, z, θ=None):
    """predict(flow, z, θ) = forward(flow, z, θ)[1] (src/Flows.jl:126)."""
    from .model import forward

    return forward(flow, z, θ)[0]


# ------------------------------------------------------------------------------------------------------------
# sample -- src/Flows.jl:157-192
# ------------------------------------------------------------------------------------------------------------


def _dims_tuple(dims) -> Tuple[int, ...]:
    return (int(dims),) if isinstance(dims, (int, np.integer)) else tuple(int(v) for v in dims)


def sample(*args):
    """sample([rng,] flow, dims[, θ]).

    dims: Integer or tuple.  θ: array of size (n, dims...) (one condition per point) or an n-tuple (same condition
    for every point, src/Flows.jl:174-185).  The base draw r ~ N(0, I) happens inside the kernel (Philox4x32-10,
    counter = sample index), replacing rand(rng, flow.base, prod(dims)) + forward!; `rng` may be an int seed or a
    torch.Generator (a seed is drawn from it)."""
    args = list(args)
    rng = None
    if not isinstance(args[0], Flow):
        rng = args.pop(0)
    flow, dims = args[0], _dims_tuple(args[1])
    θ = args[2] if len(args) > 2 else None
    pc = flow.packed()
    if isinstance(rng, (int, np.integer)):
        seed_ = int(rng)
    else:
        seed_ = int(torch.randint(0, 2**62, (1,), generator=rng or _gen()).item())
    B = int(np.prod(dims)) if dims else 1
    out = jl_empty((flow.d,) + dims, pc.device)
    flags = L.THETA_NORMALIZE if flow.n > 0 else 0
    if isinstance(θ, tuple):
        if len(θ) != flow.n:
            raise ValueError(f"θ must be an NTuple of length n={flow.n}")
        θc = torch.tensor([float(v) for v in θ], device=pc.device, dtype=torch.float32) if flow.n > 0 else None
        pc.sample_rng(B, seed_, None, θc, flags, out=out)
    else:
        if θ is None:
            if flow.n != 0:
                raise ValueError("dimensions θ must match (n, dims...) with n number of trained parameters")
        else:
            θ = to_jl(θ, pc.device)
            # @assert K == M+1, src/Flows.jl:165
            assert θ.dim() == len(dims) + 1 and tuple(θ.shape[1:]) == dims and int(θ.shape[0]) == flow.n, \
                "dimensions θ must match (n, dims...) with n number of trained parameters"
        pc.sample_rng(B, seed_, θ if flow.n > 0 else None, None, flags, out=out)
    return out


def sample_with_rejection(*args):
    """sample_with_rejection([rng,] condition, flow, dims, θ::Tuple[, m=100]) -- src/Flows.jl:196-229.

    The reference draws ONE point per iteration and tests `condition(point, θ)`; here the same stream of points
    (Philox counter = draw index, so draw i is the same point whatever the batching) is generated in device-side
    batches, `condition(points (d, nb), θ)` returns a boolean mask over the batch, and the accepted points are compacted
    in draw order.  At most m * prod(dims) points are drawn; if that is not enough the reference's ArgumentError
    becomes a ValueError."""
    args = list(args)
    rng = None
    if not callable(args[0]):
        rng = args.pop(0)
    condition, flow, dims, θ = args[0], args[1], _dims_tuple(args[2]), tuple(args[3])
    m = int(args[4]) if len(args) > 4 else 100
    if len(θ) != flow.n:
        raise ValueError(f"θ must be an NTuple of length n={flow.n}")
    pc = flow.packed()
    if isinstance(rng, (int, np.integer)):
        seed_ = int(rng)
    else:
        seed_ = int(torch.randint(0, 2**62, (1,), generator=rng or _gen()).item())
    n_pts = int(np.prod(dims)) if dims else 1
    flags = L.THETA_NORMALIZE if flow.n > 0 else 0
    θc = torch.tensor([float(v) for v in θ], device=pc.device, dtype=torch.float32) if flow.n > 0 else None
    out = jl_empty((flow.d, n_pts), pc.device)
    have, drawn, budget = 0, 0, m * n_pts
    nb = max(1024, 2 * n_pts)
    while have < n_pts and drawn < budget:
        cur = min(nb, budget - drawn)
        pts = pc.sample_rng(cur, seed_, None, θc, flags, first_sample=drawn)
        mask = torch.as_tensor(condition(pts, θ), device=pc.device).reshape(-1).to(torch.bool)
        if mask.numel() != cur:
            raise ValueError("condition must return one boolean per point of the batch")
        acc = pts[:, mask]
        take = min(int(acc.shape[1]), n_pts - have)
        out[:, have:have + take] = acc[:, :take]
        have += take
        drawn += cur
        nb = min(2 * nb, 1 << 24)
    if have < n_pts:
        raise ValueError("Impossible to reach convergence of rejection sampling")  # src/Flows.jl:221-224
    res = jl_empty((flow.d,) + dims, pc.device)  # Julia reshape(r, (D, dims...)): same memory order
    flat_view(res).copy_(flat_view(out))
    return res


# ------------------------------------------------------------------------------------------------------------
# logpdf / pdf -- src/Flows.jl:272-349
# ------------------------------------------------------------------------------------------------------------


def _θ_broadcast(flow: Flow, θ: tuple, tail: Tuple[int, ...], device) -> Optional[torch.Tensor]:
    if flow.n == 0:
        return None
    col = torch.tensor([float(v) for v in θ], device=device, dtype=torch.float32)
    B = int(np.prod(tail)) if tail else 1
    return to_jl(col.reshape(-1, 1).expand(flow.n, B).reshape((flow.n,) + tuple(tail)), device)


def logpdf(flow: Flow, x, θ=None):
    """logpdf(flow, x[, θ]): x of size (d, dims...) with θ an array (n, dims...) or an n-tuple, or x a tuple of d
    vectors -> values on the tensor-product grid (src/Flows.jl:272-331)."""
    pc = flow.packed()
    flags = L.THETA_NORMALIZE if flow.n > 0 else 0
    if isinstance(x, (tuple, list)) and not isinstance(x, torch.Tensor):
        if len(x) != flow.d:
            raise ValueError(f"grid logpdf needs {flow.d} coordinate vectors")
        vs = [torch.as_tensor(np.asarray(v, np.float32) if not isinstance(v, torch.Tensor) else v, dtype=torch.float32,
                              device=pc.device).reshape(-1) for v in x]
        # Iterators.product(x...): first vector varies fastest == Julia column-major grid (src/Flows.jl:301).  The kernel
        # derives every point from its flat index (dflow_logpdf_grid): neither the (d, prod(lens)) point array nor the
        # (n, prod(lens)) broadcast of θ is materialised.
        θt = tuple(θ) if θ is not None else ()
        if len(θt) != flow.n:
            raise ValueError(f"θ must be an NTuple of length n={flow.n}")
        return pc.logpdf_grid(vs, θt, flags)
    x = to_jl(x, pc.device)
    if isinstance(θ, tuple):
        θ = _θ_broadcast(flow, θ, tuple(x.shape[1:]), pc.device)
    return pc.logpdf(x, θ, flags)


def pdf(flow: Flow, x, θ=None):
    """pdf = exp.(logpdf) (src/Flows.jl:345-349)."""
    return torch.exp(logpdf(flow, x, θ))


# ------------------------------------------------------------------------------------------------------------
# train! -- src/Flows.jl:380-445
# ------------------------------------------------------------------------------------------------------------


def _dist():
    import torch.distributed as dist

    return dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of an n-element index list owned by `rank` (SURVEY.md §8e partitioning)."""
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def allreduce_sum_(buf: torch.Tensor) -> torch.Tensor:
    """In-place all-reduce(sum) of the packed [grad | Σlogp | #non-finite] buffer (NCCL on GPUs, gloo in CPU tests);
    a no-op for a single process."""
    d = _dist()
    if d is not None:
        d.all_reduce(buf)
    return buf


class TrainStep:
    """One data-parallel minibatch step on resident data: adjoint kernel on this rank's slice of the index list
    (seed 1/B_global), NCCL all-reduce(sum) of [grad | Σlogp | #nonfinite], Adam kernel on every replica."""

    def __init__(self, pc: PackedChain, state: OptimiserState):
        self.pc, self.state = pc, state
        state._ensure(pc)
        self.buf = torch.zeros(max(pc.P, 1) + 2, device=pc.device, dtype=torch.float32)  # grad | loss2
        self.grad = self.buf[: max(pc.P, 1)]
        self.loss2 = self.buf[max(pc.P, 1):]

    def __call__(self, x, θ, idx: Optional[torch.Tensor], B_global: int, flags: int = 0) -> None:
        pc, st = self.pc, self.state
        self.buf.zero_()
        nb = n_samples(x) if idx is None else int(idx.numel())
        if nb > 0:
            pc.loss_grad(x, θ, self.grad, self.loss2, 1.0 / B_global, flags, idx)
        allreduce_sum_(self.buf)
        st.t += 1
        r = st.rule
        pc.adam_step(self.grad, st.m, st.v, st.t, r.eta, r.beta, r.epsilon)


class _DevView:
    """Float32 view of raw device memory for torch.as_tensor (no copy)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 3}


class PeerUnavailable(RuntimeError):
    """Raised on EVERY rank when some rank could not set up the NVLink peer all-reduce."""


class PeerTrainStep:
    """The same step with the collective fused into the optimiser kernel: every rank publishes its gradient buffer
    over CUDA IPC, and ONE kernel per rank waits for the peers' buffers of this step, sums them over NVLink peer
    loads in rank order and applies Adam (csrc/dflow_dp.cu).  No NCCL call on the step; replicas stay bit-identical.
    Construct it on every rank (it exchanges the IPC handles through torch.distributed)."""

    def __init__(self, pc: PackedChain, state: OptimiserState):
        import ctypes as C

        d = _dist()
        if d is None:
            raise RuntimeError("PeerTrainStep needs an initialised torch.distributed group with world_size > 1")
        self.pc, self.state = pc, state
        state._ensure(pc)
        self.rank, self.world = d.get_rank(), d.get_world_size()
        self.P = max(pc.P, 1)
        handle = C.create_string_buffer(64)
        self.dp = C.c_void_p()
        self._views = {}
        self.loss2 = torch.zeros(2, device=pc.device, dtype=torch.float32)

        # Every rank runs the SAME collective sequence whatever fails locally: phase result -> all_reduce(MIN) -> next
        # phase.  A failure on any rank makes every rank drop its partial context and raise together, so that
        # make_train_step can fall back to the NCCL step on all ranks without mismatched collectives.
        def agree(ok_local: bool, what: str) -> None:
            flag = torch.tensor([1.0 if ok_local else 0.0], device=pc.device)
            d.all_reduce(flag, op=d.ReduceOp.MIN)
            if flag.item() < 1:
                self._destroy()
                raise PeerUnavailable(f"peer all-reduce unavailable ({what} failed on some rank)")

        ok = True
        with torch.cuda.device(pc.device):
            try:
                L.check(L.lib().dflow_dp_create(self.rank, self.world, self.P, C.byref(self.dp), handle))
            except Exception:
                ok = False
            agree(ok, "dflow_dp_create")
            handles: list = [None] * self.world
            d.all_gather_object(handles, handle.raw)
            self._handles = b"".join(handles)
            try:
                L.check(L.lib().dflow_dp_connect(self.dp, self._handles))
            except Exception:
                ok = False
            agree(ok, "dflow_dp_connect")
        d.barrier()

    def _destroy(self) -> None:
        if getattr(self, "dp", None):
            L.lib().dflow_dp_destroy(self.dp)
            self.dp = None

    def _next_buffer(self) -> torch.Tensor:
        ptr = int(L.lib().dflow_dp_grad_buffer(self.dp))
        v = self._views.get(ptr)
        if v is None:
            v = torch.as_tensor(_DevView(ptr, self.P + 2), device=self.pc.device)
            self._views[ptr] = v
        return v

    def __call__(self, x, θ, idx: Optional[torch.Tensor], B_global: int, flags: int = 0) -> None:
        pc, st = self.pc, self.state
        buf = self._next_buffer()
        buf.zero_()
        nb = n_samples(x) if idx is None else int(idx.numel())
        if nb > 0:
            pc.loss_grad(x, θ, buf[: self.P], buf[self.P:], 1.0 / B_global, flags, idx)
        st.t += 1
        r = st.rule
        with torch.cuda.device(pc.device):
            L.check(L.lib().dflow_dp_allreduce_adam(self.dp, pc.W.data_ptr(), st.m.data_ptr(), st.v.data_ptr(), r.eta,
                                                    r.beta[0], r.beta[1], r.epsilon, st.t, self.loss2.data_ptr(),
                                                    pc._stream()))

    def check(self) -> None:
        """Raises if a peer failed to reach a step's barrier (synchronises the stream)."""
        with torch.cuda.device(self.pc.device):
            s = L.lib().dflow_dp_status(self.dp, self.pc._stream())
        if s != 0:
            raise RuntimeError("data-parallel step: a peer did not publish its gradient (timeout)")

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass


class LocalDataParallel:
    """Data-parallel training driven by ONE host process (the reference's train! is a single Julia process,
    src/Flows.jl:380-445): one replica per device -- libdflow handle, packed parameters, Adam moments, resident data
    shard -- and `dflow_dp_train_step` fanning a minibatch out over the devices' streams; the replicas meet in the
    fused NVLink peer all-reduce + Adam kernel (csrc/dflow_dp.cu) and stay bit-identical.

    `devices` may name the same GPU more than once (replicas sharing a device), which is how the single-GPU tests
    exercise the multi-rank protocol."""

    def __init__(self, chain: FlowChain, devices: Sequence[int], rule: Optional[Adam] = None, theta_min=None,
                 theta_max=None):
        import ctypes as C

        self.devices = [int(v) for v in devices]
        self.rule = rule or Adam()
        leaves = chain._leaves()
        dev0 = torch.device("cuda", self.devices[0])
        primary = chain._packed
        if primary is None or primary.device != dev0:
            primary = PackedChain(leaves, dev0, theta_min, theta_max)
            chain._packed = primary
        elif theta_min is not None:
            primary.set_theta_range(theta_min, theta_max)
        primary.refresh()
        self.replicas: List[PackedChain] = [primary]
        for dv in self.devices[1:]:
            self.replicas.append(PackedChain(leaves, torch.device("cuda", dv), theta_min, theta_max, replica_of=primary))
        self.P = max(primary.P, 1)
        self.m = [torch.zeros(self.P, device=r.device) for r in self.replicas]
        self.v = [torch.zeros(self.P, device=r.device) for r in self.replicas]
        self.loss2 = [torch.zeros(2, device=r.device) for r in self.replicas]
        self.t = 0
        n = len(self.devices)
        self._dps = (C.c_void_p * n)()
        devs = (C.c_int32 * n)(*self.devices)
        L.check(L.lib().dflow_dp_create_local(n, devs, self.P, self._dps))
        self._shards = (L.DpShard * n)()
        self.x: List[Optional[torch.Tensor]] = [None] * n
        self.θ: List[Optional[torch.Tensor]] = [None] * n
        self._keep: list = []

    def set_data(self, xs: Sequence, θs: Optional[Sequence] = None) -> None:
        """Resident shard of every rank: xs[r] of size (d, N_r), θs[r] of size (n, N_r) (moved to rank r's device)."""
        for r, rep in enumerate(self.replicas):
            self.x[r] = to_jl(xs[r], rep.device)
            self.θ[r] = to_jl(θs[r], rep.device) if (θs is not None and rep.n > 0) else None

    def step(self, idxs: Sequence[Optional[torch.Tensor]], B_global: int, flags: int = 0) -> None:
        """One minibatch: idxs[r] = int32 column indices into rank r's shard (None: all of its columns); the seed of
        every rank's adjoint is 1 / B_global so that the peer all-reduce is a pure sum."""
        self.t += 1
        self._keep = []
        for r, rep in enumerate(self.replicas):
            sh = self._shards[r]
            idx = idxs[r]
            if idx is not None:
                idx = idx.to(device=rep.device, dtype=torch.int32).contiguous()
                self._keep.append(idx)
            B = n_samples(self.x[r]) if idx is None else int(idx.numel())
            ws = rep._workspace(max(B, 1))
            sh.chain = rep.handle.value
            sh.W, sh.m, sh.v = rep.W.data_ptr(), self.m[r].data_ptr(), self.v[r].data_ptr()
            sh.x = flat_view(self.x[r]).data_ptr()
            sh.theta = flat_view(self.θ[r]).data_ptr() if self.θ[r] is not None else None
            sh.B = B
            sh.idx = idx.data_ptr() if idx is not None else None
            sh.ws, sh.ws_bytes = ws.data_ptr(), ws.numel()
            sh.loss2_out = self.loss2[r].data_ptr()
            sh.stream = None
        ru = self.rule
        L.check(L.lib().dflow_dp_train_step(self._dps, len(self.replicas), self._shards, 1.0 / B_global, flags, ru.eta,
                                            ru.beta[0], ru.beta[1], ru.epsilon, self.t))

    def sync(self) -> None:
        """Waits for every device; raises if a rank missed the peer barrier (its update was skipped)."""
        s = L.lib().dflow_dp_sync(self._dps, len(self.replicas), self._shards)
        if s > 0:
            raise RuntimeError("data-parallel step: a peer did not publish its gradient in time; the step was skipped")
        if s < 0:
            L.check(s)

    def __del__(self):
        try:
            for i in range(len(self.replicas)):
                if self._dps[i]:
                    L.lib().dflow_dp_destroy(self._dps[i])
                    self._dps[i] = None
        except Exception:
            pass


def make_train_step(pc: PackedChain, state: OptimiserState):
    """PeerTrainStep when several GPU ranks of one node train together (DFLOW_DP=nccl keeps the NCCL all-reduce),
    else TrainStep."""
    import os

    d = _dist()
    if d is None or pc.device.type != "cuda" or os.environ.get("DFLOW_DP", "peer") == "nccl":
        return TrainStep(pc, state)
    try:
        # PeerTrainStep agrees on success or failure across ranks internally (same collectives on every rank)
        return PeerTrainStep(pc, state)
    except PeerUnavailable:
        return TrainStep(pc, state)  # no P2P mapping on some rank: every rank uses the NCCL collective


def _full_loss(pc: PackedChain, x, θ, idx: torch.Tensor, n_global: int, flags: int, tmp: torch.Tensor) -> float:
    """loss = -mean(logpdf(base, z) + ldj) over a whole partition (src/Flows.jl:419-430), sharded over ranks."""
    tmp.zero_()
    if idx.numel() > 0:
        pc.logpdf_sum(x, θ, tmp, flags, idx)
    allreduce_sum_(tmp)
    s, bad = tmp.tolist()
    return float("nan") if bad > 0 else -s / max(n_global, 1)


def _shard_schedule(n_total: int, batchsize: int, world: int, n_local: Sequence[int]) -> List[List[int]]:
    """Per minibatch step, how many samples every rank contributes from its resident shard: the minibatch of nb samples is
    split like shard_range(nb, r, world), clipped to what the rank still has this epoch.  Pure arithmetic, identical on every
    rank, so the global minibatch size of a step (the 1 / B_global seed) needs no communication."""
    left = list(n_local)
    steps = []
    for b0 in range(0, n_total, batchsize):
        nb = min(batchsize, n_total - b0)
        row = []
        for r in range(world):
            lo, hi = shard_range(nb, r, world)
            k = min(hi - lo, left[r])
            left[r] -= k
            row.append(k)
        steps.append(row)
    # anything a rank has left (uneven shards) joins the last step so that every sample is visited once per epoch
    if steps:
        for r in range(world):
            steps[-1][r] += left[r]
    return steps


def train_(flow: Flow, data: DataArrays, optimiser_state: OptimiserState, epochs: int = 100, batchsize: int = 64,
           shuffle: bool = True, verbose: bool = True, debug: bool = False, rng: Optional[torch.Generator] = None):
    """train!(flow, data, state; epochs=100, batchsize=64, shuffle=true, verbose=true, debug=false),
    src/Flows.jl:380-445.  Per epoch: reshuffled minibatches (partial last batch kept, like Flux.DataLoader),
    gradient + Adam per batch, then the full-set training and validation losses are pushed to the flow.

    The epoch's order is drawn on the device (dflow_shuffle_indices).  Under torch.distributed every rank either holds the
    whole dataset and takes its slice of every (globally shuffled) minibatch, or -- after `data.shard_(rank, world)` --
    holds only its shard and shuffles shard-locally."""
    pc = flow.packed()
    flags = L.THETA_NORMALIZE if flow.n > 0 else 0  # θ is normalised in-kernel (src/Data.jl:189-199)
    x = to_jl(data.x, pc.device)
    θ = to_jl(data.θ, pc.device) if flow.n > 0 else None
    if x.dim() != 2:
        raise NotImplementedError("train! partitions along dim 2; only (d, N) arrays are supported (src/Data.jl:167)")
    tr = data.partition.training.to(pc.device).contiguous()
    va = data.partition.validation.to(pc.device).contiguous()
    d = _dist()
    rank, world = (d.get_rank(), d.get_world_size()) if d is not None else (0, 1)
    sharded = bool(getattr(data, "is_shard", False)) and world > 1
    gen = rng or _gen()
    seed0 = int(torch.randint(0, 2**62, (1,), generator=gen).item())
    if d is not None:
        # Ranks build their FlowChain / DataArrays independently (different RNG streams): make rank 0 authoritative for
        # the initial parameters, the optimiser state, the train / validation split and the shuffle stream, so that the
        # replicas really are replicas (they then stay bit-identical without any further broadcast).
        optimiser_state._ensure(pc)
        n_tr_g = data.n_training_global if sharded else int(tr.numel())
        n_va_g = data.n_validation_global if sharded else int(va.numel())
        sizes = torch.tensor([n_tr_g, n_va_g, optimiser_state.t, seed0], device=pc.device, dtype=torch.int64)
        mine = sizes.clone()
        d.broadcast(sizes, src=0)
        if mine[:2].tolist() != sizes[:2].tolist():
            raise ValueError("ranks disagree on the size of the training / validation partitions")
        optimiser_state.t, seed0 = int(sizes[2].item()), int(sizes[3].item())
        bufs = [pc.W, optimiser_state.m, optimiser_state.v] + ([] if sharded else [tr, va])
        for buf in bufs:
            d.broadcast(buf, src=0)
    step = make_train_step(pc, optimiser_state)
    n_tr = data.n_training_global if sharded else int(tr.numel())
    n_va = data.n_validation_global if sharded else int(va.numel())
    tmp = torch.zeros(2, device=pc.device, dtype=torch.float32)
    schedule = None
    if sharded:
        n_local = [shard_range(n_tr, r, world)[1] - shard_range(n_tr, r, world)[0] for r in range(world)]
        if n_local[rank] != int(tr.numel()):
            raise ValueError("this rank's resident training shard does not match shard_range(n_training_global, rank, world)")
        schedule = _shard_schedule(n_tr, batchsize, world, n_local)

    def shard(v: torch.Tensor) -> torch.Tensor:
        if world == 1 or sharded:
            return v
        lo, hi = shard_range(int(v.numel()), rank, world)
        return v[lo:hi]

    for ep in range(epochs):
        if shuffle:
            # drawn on the device; every rank evaluates the same bijection (global shuffle), or -- sharded -- its own
            # permutation of its resident shard
            eseed = seed0 + ep + (1000003 * (rank + 1) if sharded else 0)
            order = device_permutation(int(tr.numel()), eseed, pc.device, base=tr)
        else:
            order = tr
        if world == 1 and not debug and isinstance(step, TrainStep):
            # single GPU: the whole epoch is enqueued from C (dflow_train_epoch) -- with the reference's default
            # batchsize = 64 a step is launch-bound, and a Python round trip per minibatch would dominate it
            st_ = optimiser_state
            st_.t = pc.train_epoch(x, θ, order, batchsize, st_.m, st_.v, st_.t, st_.rule.eta, st_.rule.beta,
                                   st_.rule.epsilon, flags, step.buf)
        elif sharded:
            cur = 0
            for row in schedule:
                k = row[rank]
                step(x, θ, order[cur: cur + k], sum(row), flags)
                cur += k
        else:
            for b0 in range(0, int(order.numel()), batchsize):
                batch = order[b0: b0 + batchsize]
                step(x, θ, shard(batch), int(batch.numel()), flags)
                if debug:
                    s_, bad = step.loss2.tolist()
                    if bad > 0 or not math.isfinite(s_):
                        raise ValueError(f"non-finite minibatch loss (Σlogp={s_}, non-finite samples={bad})")  # Flows.jl:405-409
        if isinstance(step, PeerTrainStep):
            step.check()  # a rank that missed the peer barrier skipped its update: stop instead of training on
        train_loss = _full_loss(pc, x, θ, shard(tr), n_tr, flags, tmp)
        flow.train_loss.append(train_loss)
        if debug and not math.isfinite(train_loss):
            print(f"Problem with train loss {train_loss}")
            return pc.normalize(x.index_select(1, tr.long()), None if θ is None else θ.index_select(1, tr.long()), flags)
        valid_loss = _full_loss(pc, x, θ, shard(va), n_va, flags, tmp)
        flow.valid_loss.append(valid_loss)
        if debug and not math.isfinite(valid_loss):
            print(f"Problem with valid loss {valid_loss}")
            return pc.normalize(x.index_select(1, va.long()), None if θ is None else θ.index_select(1, va.long()), flags)
        if verbose and rank == 0:
            print(f"epoch: {len(flow.train_loss)} | train_loss = {train_loss}, valid_loss = {valid_loss}")
    if debug:
        return None, None
    return None
