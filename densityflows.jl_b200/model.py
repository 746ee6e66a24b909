"""Host-side mirror of the reference's flow-element types (same names, argument meaning and error behaviour):
CouplingAxes (src/Axes.jl), Dense/Chain conditioners (Flux, src/Layers.jl:33-50), RNVPCouplingLayer
(src/affine/RNVP.jl), NICECouplingLayer (src/affine/NICE.jl), CouplingLayer factory (src/Layers.jl:113-158),
CouplingBlock (src/Blocks.jl), NormalizationLayer (src/norm/Normalization.jl), FlowChain / concatenate
(src/Chains.jl) and the generic `backward` / `forward` / `forward_` (= `forward!`) functions.

All arithmetic on samples happens in libdflow.so (hand-written sm_100a kernels) through the C ABI; this module
only builds descriptors, owns the packed parameter buffer and hands device pointers across.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib as L
from .arrays import flat_view, is_colmajor, jl_empty, n_samples, tail_shape, to_jl

# ------------------------------------------------------------------------------------------------------------
# CouplingAxes -- src/Axes.jl (1-based integer vectors exactly as the reference builds them)
# ------------------------------------------------------------------------------------------------------------


class CouplingAxes:
    """CouplingAxes(d, mask; n) / CouplingAxes(d, j=d÷2; n, reverse) / CouplingAxes(data, ...), src/Axes.jl:79-119."""

    __slots__ = ("d", "n", "axis_id", "axis_af", "axis_nn")

    def __init__(self, d, mask_or_j=None, n: Optional[int] = None, reverse: bool = False, *, _raw=None):
        if _raw is not None:
            self.d, self.n, self.axis_id, self.axis_af, self.axis_nn = _raw
            return
        from .data import DataArrays, number_conditions, number_dimensions

        if isinstance(d, DataArrays):  # src/Axes.jl:117-119
            data = d
            d = number_dimensions(data)
            if n is not None:
                raise TypeError("n is taken from the data")
            n = number_conditions(data)
        n = 0 if n is None else int(n)
        d = int(d)
        if mask_or_j is None or isinstance(mask_or_j, (int, np.integer)):
            j = d // 2 if mask_or_j is None else int(mask_or_j)
            mask = list(range(j + 1, d + 1)) if not reverse else list(range(1, j + 1))  # src/Axes.jl:111
        else:
            if reverse:
                raise TypeError("reverse is only defined for the cut form")
            mask = [int(m) for m in mask_or_j]
        assert max(mask) <= d, "The mask cannot contain values higher than the dimension"  # src/Axes.jl:85
        self.d = d
        self.n = n
        self.axis_id = [k for k in range(1, d + 1) if k not in mask]  # src/Axes.jl:88
        self.axis_af = list(mask)  # src/Axes.jl:91 (caller order kept)
        self.axis_nn = list(range(1, n + 1)) + [k + n for k in self.axis_id]  # src/Axes.jl:98

    def __eq__(self, other):  # src/Axes.jl:46-56
        if not isinstance(other, CouplingAxes):
            return NotImplemented
        return (
            self.d == other.d
            and self.n == other.n
            and sorted(self.axis_id) == sorted(other.axis_id)
            and sorted(self.axis_af) == sorted(other.axis_af)
            and sorted(self.axis_nn) == sorted(other.axis_nn)
        )

    __hash__ = None

    def __repr__(self):
        return (f"(d,n)=({self.d},{self.n}); identity=({','.join(map(str, self.axis_id))}), "
                f"transformed=({','.join(map(str, self.axis_af))})")


def reverse(axes: CouplingAxes) -> CouplingAxes:
    """Base.reverse(axes), src/Axes.jl:129-135."""
    axis_nn = list(range(1, axes.n + 1)) + [k + axes.n for k in axes.axis_af]
    return CouplingAxes(None, _raw=(axes.d, axes.n, list(axes.axis_af), list(axes.axis_id), axis_nn))


def is_reverse(a1: CouplingAxes, a2: CouplingAxes) -> bool:
    """src/Axes.jl:137-139."""
    return list(a1.axis_af) == list(a2.axis_id) and list(a2.axis_af) == list(a1.axis_id) and a1.n == a2.n


# ------------------------------------------------------------------------------------------------------------
# Conditioner networks (Flux.Dense / Flux.Chain mirror)
# ------------------------------------------------------------------------------------------------------------

_ACTS = {"identity": L.ACT_IDENTITY, "relu": L.ACT_RELU, "tanh": L.ACT_TANH, "sigmoid": L.ACT_SIGMOID}
relu, tanh, sigmoid, identity = "relu", "tanh", "sigmoid", "identity"

_default_gen: Optional[torch.Generator] = None


def seed(s: int) -> None:
    """Seed the host RNG used for weight initialisation and data partitions (Random.seed! analogue)."""
    global _default_gen
    _default_gen = torch.Generator(device="cpu")
    _default_gen.manual_seed(int(s))


def _gen() -> torch.Generator:
    global _default_gen
    if _default_gen is None:
        _default_gen = torch.Generator(device="cpu")
        # torch.initial_seed() only READS the global generator's seed (torch.seed() would re-seed it, clobbering the
        # user's RNG stream and giving every torchrun rank a different stream)
        _default_gen.manual_seed(torch.initial_seed() & 0x7FFFFFFF)
    return _default_gen


class Dense:
    """Flux.Dense(in => out, σ; bias): weight is (out, in) column-major, glorot_uniform; bias zeros."""

    def __init__(self, in_dim: int, out_dim: int, act: Union[str, int] = identity, bias: bool = True,
                 weight: Optional[torch.Tensor] = None, bias_value: Optional[torch.Tensor] = None):
        self.act = _ACTS[act] if isinstance(act, str) else int(act)
        if weight is None:
            s = math.sqrt(6.0 / (in_dim + out_dim))  # Flux.glorot_uniform
            w = (torch.rand((in_dim, out_dim), generator=_gen(), dtype=torch.float32) * 2 - 1) * s
            weight = w.t()  # logical (out, in) with column-major strides
        else:
            weight = to_jl(weight)
            assert tuple(weight.shape) == (out_dim, in_dim)
        self.weight = weight
        if bias:
            self.bias = torch.zeros(out_dim, dtype=torch.float32) if bias_value is None else to_jl(bias_value)
        else:
            self.bias = None

    @property
    def in_dim(self) -> int:
        return int(self.weight.shape[1])

    @property
    def out_dim(self) -> int:
        return int(self.weight.shape[0])


class Chain:
    """Flux.Chain of Dense layers."""

    def __init__(self, *layers: Dense):
        self.layers = list(layers)

    def widths(self) -> List[int]:
        return [self.layers[0].in_dim] + [l.out_dim for l in self.layers]

    def n_params(self) -> int:
        return sum(l.weight.numel() + (l.bias.numel() if l.bias is not None else 0) for l in self.layers)


def _dflt_net(input_dim: int, output_dim: int, n: int, hidden_dim: int = 32, σ=relu, bias: bool = True) -> Chain:
    """_dflt_net, src/Layers.jl:33-50."""
    layers = [Dense(input_dim, hidden_dim, σ, bias=bias)]
    layers += [Dense(hidden_dim, hidden_dim, σ, bias=bias) for _ in range(n - 1)]
    layers += [Dense(hidden_dim, output_dim, identity, bias=bias)]
    return Chain(*layers)


# ------------------------------------------------------------------------------------------------------------
# Flow elements
# ------------------------------------------------------------------------------------------------------------


class FlowElement:
    """abstract type FlowElement, src/DensityFlows.jl:45."""

    _packed = None

    def __call__(self, z, θ=None):  # @auto_functor, src/Macros.jl:68-82
        return forward(self, z, θ)

    def _leaves(self) -> List["FlowElement"]:
        return [self]


class CouplingLayerBase(FlowElement):
    """abstract type CouplingLayer <: FlowElement, src/DensityFlows.jl:48."""

    axes: CouplingAxes


class RNVPCouplingLayer(CouplingLayerBase):
    """RNVPCouplingLayer(s_net, t_net, axes), src/affine/RNVP.jl:41-51 (trainable = (s_net, t_net))."""

    def __init__(self, s_net: Chain, t_net: Chain, axes: CouplingAxes):
        self.s_net, self.t_net, self.axes = s_net, t_net, axes

    def summarize(self) -> str:  # src/affine/RNVP.jl:59-69
        return (f"RNVPCouplingLayer | s_net > {self.s_net.widths()} ({self.s_net.n_params()} parameters)\n"
                f"                  | t_net > {self.t_net.widths()} ({self.t_net.n_params()} parameters)\n"
                f"                  | axes  > {self.axes!r}")


class NICECouplingLayer(CouplingLayerBase):
    """NICECouplingLayer(t_net, axes), src/affine/NICE.jl (additive coupling, ln_det_jac = 0)."""

    def __init__(self, t_net: Chain, axes: CouplingAxes):
        self.t_net, self.axes = t_net, axes


def CouplingLayer(*args, **kws) -> CouplingLayerBase:
    """CouplingLayer([T=RNVPCouplingLayer,] axes | d[, j|mask] | data[, j|mask]; n, reverse, n_sublayers_{s,t},
    hidden_dim_{s,t}, σ_{s,t}, bias) and CouplingLayer([s_net,] t_net, axes|data[, mask]); src/Layers.jl:110-158."""
    from .data import DataArrays

    args = list(args)
    T = RNVPCouplingLayer
    if args and isinstance(args[0], type) and issubclass(args[0], CouplingLayerBase):
        T = args.pop(0)
    nets = []
    while args and isinstance(args[0], Chain):
        nets.append(args.pop(0))
    # resolve the axes
    if args and isinstance(args[0], CouplingAxes):
        axes = args.pop(0)
    else:
        n = kws.pop("n", None)
        rev = kws.pop("reverse", False)
        first = args.pop(0)
        second = args.pop(0) if args else None
        if isinstance(first, DataArrays):
            axes = CouplingAxes(first, second, reverse=rev)
        else:
            axes = CouplingAxes(first, second, n=0 if n is None else n, reverse=rev)
    if args:
        raise TypeError(f"unexpected positional arguments {args}")
    if nets:  # src/Layers.jl:110-111
        if kws:
            raise TypeError(f"unexpected keyword arguments {sorted(kws)}")
        return NICECouplingLayer(nets[0], axes) if len(nets) == 1 else RNVPCouplingLayer(nets[0], nets[1], axes)
    n_sublayers_t = kws.pop("n_sublayers_t", 2)
    n_sublayers_s = kws.pop("n_sublayers_s", 2)
    hidden_dim_t = kws.pop("hidden_dim_t", 32)
    hidden_dim_s = kws.pop("hidden_dim_s", 32)
    σ_t = kws.pop("σ_t", kws.pop("act_t", relu))
    σ_s = kws.pop("σ_s", kws.pop("act_s", relu))
    bias = kws.pop("bias", True)
    if kws:
        raise TypeError(f"unexpected keyword arguments {sorted(kws)}")
    input_dim, output_dim = len(axes.axis_nn), len(axes.axis_af)  # src/Layers.jl:126-127
    t_net = _dflt_net(input_dim, output_dim, n_sublayers_t, hidden_dim=hidden_dim_t, σ=σ_t, bias=bias)
    if T is NICECouplingLayer:
        return NICECouplingLayer(t_net, axes)
    s_net = _dflt_net(input_dim, output_dim, n_sublayers_s, hidden_dim=hidden_dim_s, σ=σ_s, bias=bias)
    return T(s_net, t_net, axes)


class CouplingBlock(FlowElement):
    """CouplingBlock(layer_1, layer_2) / CouplingBlock([T,] axes | d[, j|mask] | data[, j|mask]; kws...),
    src/Blocks.jl:64-120: two layers with complementary axes."""

    def __init__(self, *args, **kws):
        if len(args) == 2 and all(isinstance(a, CouplingLayerBase) for a in args) and not kws:
            layer_1, layer_2 = args
        else:
            args = list(args)
            T = RNVPCouplingLayer
            if args and isinstance(args[0], type) and issubclass(args[0], CouplingLayerBase):
                T = args.pop(0)
            if args and isinstance(args[0], CouplingAxes):
                first_axes = args.pop(0)
            else:
                from .data import DataArrays

                n = kws.pop("n", None)
                rev = kws.pop("reverse", False)
                first = args.pop(0)
                second = args.pop(0) if args else None
                if isinstance(first, DataArrays):
                    first_axes = CouplingAxes(first, second, reverse=rev)
                else:
                    first_axes = CouplingAxes(first, second, n=0 if n is None else n, reverse=rev)
            second_axes = reverse(first_axes)  # src/Blocks.jl:96
            layer_1 = CouplingLayer(T, first_axes, **kws)
            layer_2 = CouplingLayer(T, second_axes, **kws)
        if not is_reverse(layer_1.axes, layer_2.axes):  # src/Blocks.jl:70-73
            raise ValueError("layer_1 and layer_2 need to have complementary axes")
        self.layer_1, self.layer_2 = layer_1, layer_2

    def __len__(self):
        return 2

    def _leaves(self):
        return [self.layer_1, self.layer_2]


class NormalizationElement(FlowElement):
    pass


class NormalizationLayer(NormalizationElement):
    """NormalizationLayer(x, α=0, β=1) / NormalizationLayer(x_min, x_max, α, β), src/norm/Normalization.jl:30-59.
    Not trainable.  The min/max over the data are computed on the device (dflow_minmax) when x lives there."""

    def __init__(self, x, *rest):
        from .data import DataArrays

        if isinstance(x, DataArrays):
            x = x.x
        if len(rest) == 3:  # raw constructor (x_min, x_max, α, β)
            x_min, (x_max, α, β) = x, rest
            self.x_min = np.asarray(torch.as_tensor(x_min).cpu(), np.float32).reshape(-1)
            self.x_max = np.asarray(torch.as_tensor(x_max).cpu(), np.float32).reshape(-1)
        else:
            α = rest[0] if len(rest) > 0 else 0.0
            β = rest[1] if len(rest) > 1 else 1.0
            self.x_min, self.x_max = minmax_rows(x)
        assert β > α, "Bounds of the normalisation need to be in the correct order, β > α."  # :55
        self.α, self.β = float(α), float(β)

    alpha = property(lambda self: self.α)
    beta = property(lambda self: self.β)


def minmax_rows(x) -> Tuple[np.ndarray, np.ndarray]:
    """vec(minimum(x, dims=2:N)), vec(maximum(x, dims=2:N)) (src/norm/Normalization.jl:52-53, src/Data.jl:182-183)."""
    if isinstance(x, torch.Tensor) and x.is_cuda:
        x = to_jl(x)
        rows, B = int(x.shape[0]), n_samples(x)
        out = torch.empty(2 * rows, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            st = torch.cuda.current_stream().cuda_stream
            L.check(L.lib().dflow_minmax(flat_view(x).data_ptr(), rows, B, out.data_ptr(),
                                         out.data_ptr() + 4 * rows, st))
        h = out.cpu().numpy()
        return h[:rows].copy(), h[rows:].copy()
    a = np.asarray(x.cpu() if isinstance(x, torch.Tensor) else x, np.float32)
    flat = a.reshape(a.shape[0], -1)
    return flat.min(axis=1), flat.max(axis=1)


class FlowChain(FlowElement):
    """FlowChain(elements...) / FlowChain(elements::Tuple) / FlowChain([T=CouplingBlock,] n, args...; kws...),
    src/Chains.jl:78-101."""

    def __init__(self, *xs, **kws):
        if xs and isinstance(xs[0], type) and issubclass(xs[0], FlowElement):
            T, n, args = xs[0], xs[1], xs[2:]
            layers = [T(*args, **kws) for _ in range(n)]
        elif xs and isinstance(xs[0], (int, np.integer)):
            layers = [CouplingBlock(*xs[1:], **kws) for _ in range(xs[0])]
        elif len(xs) == 1 and isinstance(xs[0], (tuple, list)):
            layers = list(xs[0])
        else:
            layers = list(xs)
        if kws and not (xs and (isinstance(xs[0], type) or isinstance(xs[0], (int, np.integer)))):
            raise TypeError(f"unexpected keyword arguments {sorted(kws)}")
        for l in layers:
            if not isinstance(l, FlowElement):
                raise TypeError(f"{l!r} is not a FlowElement")
        self.layers = layers

    # Base forwarding, src/Chains.jl:125-138
    def __len__(self):
        return len(self.layers)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return self.layers[i]
        return self.layers[i]

    def __iter__(self):
        return iter(self.layers)

    def _leaves(self):
        out = []
        for l in self.layers:
            out.extend(l._leaves())
        return out

    def summarize(self) -> str:
        return "\n".join(l.summarize() if hasattr(l, "summarize") else type(l).__name__ for l in self._leaves())

    # packed-parameter access (the device cache of the Flux structs; SURVEY.md §8b ownership)
    def packed(self, device=None) -> "PackedChain":
        return _packed_of(self, device)


def concatenate(*xs) -> FlowChain:
    """concatenate, src/Chains.jl:112-123."""
    if len(xs) == 1 and isinstance(xs[0], (tuple, list)):
        xs = tuple(xs[0])
    out: List[FlowElement] = []
    for x in xs:
        if isinstance(x, (tuple, list)):
            out.extend(x)
        elif isinstance(x, FlowChain):
            out.extend(x.layers)
        else:
            out.append(x)
    return FlowChain(*out)


# ------------------------------------------------------------------------------------------------------------
# PackedChain: descriptor + libdflow handle + the packed parameter buffer
# ------------------------------------------------------------------------------------------------------------


def _nets_of(e: FlowElement) -> List[Chain]:
    if isinstance(e, RNVPCouplingLayer):
        return [e.s_net, e.t_net]  # Flux.@layer ... trainable=(s_net, t_net)
    if isinstance(e, NICECouplingLayer):
        return [e.t_net]
    return []


def chain_dims(leaves: Sequence[FlowElement]) -> Tuple[int, int]:
    d = n = None
    for e in leaves:
        if isinstance(e, CouplingLayerBase):
            if d is not None and (e.axes.d != d or (n is not None and e.axes.n != n)):
                raise ValueError("all coupling layers of a chain must share (d, n)")
            d, n = e.axes.d, e.axes.n
        elif isinstance(e, NormalizationLayer):
            dd = len(e.x_min)
            if d is not None and dd != d:
                raise ValueError("NormalizationLayer dimension does not match the chain")
            d = dd
        else:
            raise DflowCustomElement(e)
    return int(d), int(0 if n is None else n)


class DflowCustomElement(NotImplementedError):
    def __init__(self, e):
        super().__init__(f"{type(e).__name__} is not one of the fused element kinds (RNVP, NICE, Normalization); "
                         "custom FlowElements (docs/src/documentation.md:170-197) have no CUDA kernel")


class ChainDescriptor:
    """Builds (and keeps alive) the ctypes dflow_chain_desc of a list of leaf elements."""

    def __init__(self, leaves: Sequence[FlowElement], theta_min=None, theta_max=None):
        self.leaves = list(leaves)
        self.d, self.n = chain_dims(self.leaves)
        self._keep = []
        elems = (L.ElemDesc * len(self.leaves))()
        for i, e in enumerate(self.leaves):
            ed = elems[i]
            if isinstance(e, NormalizationLayer):
                ed.kind = L.ELEM_NORM
                xm = (C.c_float * self.d)(*[float(v) for v in e.x_min])
                xM = (C.c_float * self.d)(*[float(v) for v in e.x_max])
                self._keep += [xm, xM]
                ed.x_min = C.cast(xm, L.c_f32p)
                ed.x_max = C.cast(xM, L.c_f32p)
                ed.alpha, ed.beta = e.α, e.β
                continue
            ed.kind = L.ELEM_RNVP if isinstance(e, RNVPCouplingLayer) else L.ELEM_NICE
            af0 = [k - 1 for k in e.axes.axis_af]  # 1-based -> 0-based at the C boundary
            arr = (C.c_int32 * len(af0))(*af0)
            self._keep.append(arr)
            ed.n_af = len(af0)
            ed.axis_af = C.cast(arr, L.c_i32p)
            id0 = [k - 1 for k in e.axes.axis_id]
            if len(id0) != self.d - len(af0) or set(id0) & set(af0):
                raise L.DflowInvalidArg(L.E_INVALID_ARG, "axis_id and axis_af must partition 1:d")
            if id0:
                arr_id = (C.c_int32 * len(id0))(*id0)
                self._keep.append(arr_id)
                ed.n_id = len(id0)
                ed.axis_id = C.cast(arr_id, L.c_i32p)
            if isinstance(e, RNVPCouplingLayer):
                self._fill_net(ed.s_net, e.s_net)
            self._fill_net(ed.t_net, e.t_net)
        self.elems = elems
        self.desc = L.ChainDesc()
        self.desc.d, self.desc.n, self.desc.n_elems = self.d, self.n, len(self.leaves)
        self.desc.elems = C.cast(elems, C.POINTER(L.ElemDesc))
        if theta_min is not None:
            tm = (C.c_float * max(self.n, 1))(*[float(v) for v in theta_min])
            tM = (C.c_float * max(self.n, 1))(*[float(v) for v in theta_max])
            self._keep += [tm, tM]
            self.desc.theta_min = C.cast(tm, L.c_f32p)
            self.desc.theta_max = C.cast(tM, L.c_f32p)

    def _fill_net(self, nd: "L.NetDesc", net: Chain) -> None:
        w = net.widths()
        has_bias = [l.bias is not None for l in net.layers]
        if any(has_bias) != all(has_bias):
            raise L.DflowUnsupported(L.E_UNSUPPORTED, "mixed bias/no-bias Dense layers in one conditioner")
        wa = (C.c_int32 * len(w))(*w)
        aa = (C.c_int32 * len(net.layers))(*[l.act for l in net.layers])
        self._keep += [wa, aa]
        nd.depth = len(net.layers)
        nd.widths = C.cast(wa, L.c_i32p)
        nd.acts = C.cast(aa, L.c_i32p)
        nd.has_bias = 1 if all(has_bias) else 0


_bind_epoch = 0


class PackedChain:
    """Device-side twin of a FlowChain: libdflow handle + packed Float32 parameter buffer `W`.

    After packing, every Dense.weight / Dense.bias of the chain is a VIEW into `W` (column-major vec(weight) then
    bias, chain order, s_net before t_net), so the Flux-like structs stay authoritative and `unpack!` is free."""

    def __init__(self, leaves: Sequence[FlowElement], device=None, theta_min=None, theta_max=None,
                 replica_of: Optional["PackedChain"] = None):
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("densityflows.jl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.cd = ChainDescriptor(leaves, theta_min, theta_max)
        self.d, self.n = self.cd.d, self.cd.n
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(L.lib().dflow_chain_create(C.byref(self.cd.desc), C.byref(self.handle)))
        self.P = int(L.lib().dflow_param_count(self.handle))
        self.has_theta_range = theta_min is not None
        self.W = torch.zeros(max(self.P, 1), device=self.device, dtype=torch.float32)
        self._ws: Optional[torch.Tensor] = None  # adjoint workspace (dflow_workspace_bytes)
        self._scratch_buf: Optional[torch.Tensor] = None  # forward-call scratch (dflow_scratch_bytes), caller owned
        if replica_of is not None:
            # a data-parallel replica on another device: same descriptor, its own copy of the parameters; the Dense
            # tensors of the chain stay views into the primary's buffer
            self.W.copy_(replica_of.W)
            self._epoch = None
        else:
            self._bind_views()

    def _bind_views(self) -> None:
        global _bind_epoch
        _bind_epoch += 1
        self._epoch = _bind_epoch
        off = 0
        for e in self.cd.leaves:
            for net in _nets_of(e):
                for dl in net.layers:
                    o, i = dl.out_dim, dl.in_dim
                    view = self.W[off:off + o * i].view(i, o).t()
                    view.copy_(dl.weight.to(self.device))
                    dl.weight = view
                    off += o * i
                    if dl.bias is not None:
                        bview = self.W[off:off + o]
                        bview.copy_(dl.bias.to(self.device))
                        dl.bias = bview
                        off += o
        assert off == self.P, (off, self.P)

    def refresh(self) -> None:
        """Re-adopt the Dense tensors if another PackedChain has re-bound them since (shared layers)."""
        if self._epoch is not None and self._epoch != _bind_epoch:
            self._bind_views()

    def set_theta_range(self, theta_min, theta_max) -> None:
        tm = (C.c_float * max(self.n, 1))(*[float(v) for v in theta_min])
        tM = (C.c_float * max(self.n, 1))(*[float(v) for v in theta_max])
        with torch.cuda.device(self.device):
            L.check(L.lib().dflow_chain_set_theta_range(self.handle, tm, tM))
        self.has_theta_range = True

    def axes_of(self, elem: int) -> Tuple[List[int], List[int]]:
        """0-based (axis_id, axis_nn) as derived inside libdflow (bit-exactness check of src/Axes.jl:88,98)."""
        aid = (C.c_int32 * self.d)()
        ann = (C.c_int32 * (self.d + self.n))()
        nid, nnn = C.c_int32(), C.c_int32()
        L.check(L.lib().dflow_chain_axes(self.handle, elem, aid, C.byref(nid), ann, C.byref(nnn)))
        return list(aid[: nid.value]), list(ann[: nnn.value])

    def tune(self, **kv) -> None:
        for k, v in kv.items():
            L.check(L.lib().dflow_set_tuning(self.handle, k.encode(), int(v)))

    def launch_count(self) -> int:
        return int(L.lib().dflow_launch_count(self.handle))

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                L.lib().dflow_chain_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass

    # ---- raw calls (pointers in, nothing allocated in the hot path except outputs) -------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _prep(self, x, θ):
        self.refresh()
        x = to_jl(x, self.device)
        if int(x.shape[0]) != self.d:
            raise ValueError(f"array has {x.shape[0]} rows but the flow has d={self.d}")
        if θ is None:
            if self.n != 0:
                raise ValueError(f"flow has n={self.n} conditions but no θ was given")
            return x, None
        θ = to_jl(θ, self.device)
        if int(θ.shape[0]) != self.n or tuple(θ.shape[1:]) != tuple(x.shape[1:]):
            raise ValueError(f"θ must have size ({self.n}, dims...) matching x: got {tuple(θ.shape)} vs {tuple(x.shape)}")
        return x, (θ if self.n > 0 else None)

    @staticmethod
    def _ptr(t: Optional[torch.Tensor]):
        return None if t is None else flat_view(t).data_ptr()

    def normalize(self, x, θ, flags: int = 0):
        x, θ = self._prep(x, θ)
        B = n_samples(x)
        z = jl_empty(x.shape, self.device)
        ldj = jl_empty(tail_shape(x) or (1,), self.device)
        self.ensure_scratch(B)
        with torch.cuda.device(self.device):
            L.check(L.lib().dflow_normalize(self.handle, self.W.data_ptr(), self._ptr(x), self._ptr(θ), B, flags,
                                            self._ptr(z), self._ptr(ldj), self._stream()))
        return z, ldj

    def forward_ldj(self, z, θ, flags: int = 0):
        z, θ = self._prep(z, θ)
        B = n_samples(z)
        x = jl_empty(z.shape, self.device)
        ldj = jl_empty(tail_shape(z) or (1,), self.device)
        self.ensure_scratch(B)
        with torch.cuda.device(self.device):
            L.check(L.lib().dflow_forward_ldj(self.handle, self.W.data_ptr(), self._ptr(z), self._ptr(θ), B, flags,
                                              self._ptr(x), self._ptr(ldj), self._stream()))
        return x, ldj

    def sample_inplace(self, z: torch.Tensor, θ=None, θ_const: Optional[torch.Tensor] = None, flags: int = 0) -> None:
        if not (isinstance(z, torch.Tensor) and z.device == self.device and z.dtype == torch.float32 and is_colmajor(z)):
            raise ValueError("forward! needs a Float32 column-major tensor on the flow's device (it is mutated in place)")
        self.refresh()
        if int(z.shape[0]) != self.d:
            raise ValueError(f"array has {z.shape[0]} rows but the flow has d={self.d}")
        if θ is not None and self.n > 0:
            θ = to_jl(θ, self.device)
            if int(θ.shape[0]) != self.n or tuple(θ.shape[1:]) != tuple(z.shape[1:]):
                raise ValueError("θ must have size (n, dims...) matching z")
        else:
            θ = None
        self.ensure_scratch(n_samples(z))
        with torch.cuda.device(self.device):
            L.check(L.lib().dflow_sample_inplace(self.handle, self.W.data_ptr(), self._ptr(z), self._ptr(θ),
                                                 None if θ_const is None else θ_const.data_ptr(), n_samples(z), flags,
                                                 self._stream()))

    def logpdf(self, x, θ, flags: int = 0, idx: Optional[torch.Tensor] = None, B: Optional[int] = None):
        x, θ = self._prep(x, θ)
        if idx is None:
            B = n_samples(x)
            out = jl_empty(tail_shape(x) or (1,), self.device)
        else:
            B = int(idx.numel()) if B is None else B
            out = torch.empty(B, device=self.device, dtype=torch.float32)
        self.ensure_scratch(B)
        with torch.cuda.device(self.device):
            L.check(L.lib().dflow_logpdf(self.handle, self.W.data_ptr(), self._ptr(x), self._ptr(θ), B,
                                         None if idx is None else idx.data_ptr(), flags, self._ptr(out), self._stream()))
        return out

    def logpdf_grid(self, vectors: Sequence[torch.Tensor], θ: Sequence[float] = (), flags: int = 0) -> torch.Tensor:
        """log-density on the tensor-product grid of d coordinate vectors with one fixed condition (src/Flows.jl:287-331);
        result of size (len_1, ..., len_d), column-major."""
        self.refresh()
        if len(vectors) != self.d:
            raise ValueError(f"grid logpdf needs {self.d} coordinate vectors")
        lens = [int(v.numel()) for v in vectors]
        B = int(np.prod(lens)) if lens else 0
        vals = torch.cat([v.to(device=self.device, dtype=torch.float32).reshape(-1) for v in vectors])
        θc = torch.tensor([float(v) for v in θ], device=self.device, dtype=torch.float32) if self.n > 0 else None
        out = jl_empty(tuple(lens), self.device)
        self.ensure_scratch(B)
        lens_c = (C.c_int64 * self.d)(*lens)
        with torch.cuda.device(self.device):
            L.check(L.lib().dflow_logpdf_grid(self.handle, self.W.data_ptr(), vals.data_ptr() if vals.numel() else None,
                                              lens_c, None if θc is None else θc.data_ptr(), flags, self._ptr(out),
                                              self._stream()))
        return out

    def logpdf_sum(self, x, θ, out2: torch.Tensor, flags: int = 0, idx: Optional[torch.Tensor] = None) -> int:
        x, θ = self._prep(x, θ)
        B = n_samples(x) if idx is None else int(idx.numel())
        self.ensure_scratch(B)
        with torch.cuda.device(self.device):
            L.check(L.lib().dflow_logpdf_sum(self.handle, self.W.data_ptr(), self._ptr(x), self._ptr(θ), B,
                                             None if idx is None else idx.data_ptr(), flags, out2.data_ptr(),
                                             self._stream()))
        return B

    def sample_rng(self, B: int, seed_: int, θ=None, θ_const: Optional[torch.Tensor] = None, flags: int = 0,
                   offset: int = 0, first_sample: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        self.refresh()
        if out is None:
            out = jl_empty((self.d, B), self.device)
        if θ is not None and self.n > 0:
            θ = to_jl(θ, self.device)
        else:
            θ = None
        self.ensure_scratch(B)
        with torch.cuda.device(self.device):
            L.check(L.lib().dflow_sample_rng(self.handle, self.W.data_ptr(), seed_, offset, first_sample, self._ptr(θ),
                                             None if θ_const is None else θ_const.data_ptr(), B, flags, self._ptr(out),
                                             self._stream()))
        return out

    def loss_grad(self, x, θ, grad: torch.Tensor, loss2: torch.Tensor, inv_btot: Optional[float] = None,
                  flags: int = 0, idx: Optional[torch.Tensor] = None) -> int:
        """grad (P floats) and loss2 (2 floats) are ACCUMULATED into; returns the batch size."""
        x, θ = self._prep(x, θ)
        B = n_samples(x) if idx is None else int(idx.numel())
        if B == 0:
            return 0  # nothing to accumulate (an empty shard of a data-parallel minibatch)
        ib = (1.0 / B) if inv_btot is None else float(inv_btot)
        need = int(L.lib().dflow_workspace_bytes(self.handle, B))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, device=self.device, dtype=torch.uint8)
        with torch.cuda.device(self.device):
            L.check(L.lib().dflow_loss_grad(self.handle, self.W.data_ptr(), self._ptr(x), self._ptr(θ), B,
                                            None if idx is None else idx.data_ptr(), ib, flags, loss2.data_ptr(),
                                            grad.data_ptr(), self._ws.data_ptr(), self._ws.numel(), self._stream()))
        return B

    def ensure_scratch(self, B: int) -> None:
        """Attach dflow_scratch_bytes(B) of caller-owned scratch to the handle (grow-only; a no-op for chains on the
        CUDA-core kernels).  Raw ABI users call this once for their largest batch; the library never allocates."""
        need = int(L.lib().dflow_scratch_bytes(self.handle, int(B)))
        if need and (self._scratch_buf is None or self._scratch_buf.numel() < need):
            torch.cuda.synchronize(self.device)  # earlier calls may still work in the old buffer
            self._scratch_buf = torch.empty(need, device=self.device, dtype=torch.uint8)
            L.check(L.lib().dflow_chain_set_scratch(self.handle, self._scratch_buf.data_ptr(), need))

    def _workspace(self, B: int) -> torch.Tensor:
        need = int(L.lib().dflow_workspace_bytes(self.handle, B))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, device=self.device, dtype=torch.uint8)
        return self._ws

    def vjp(self, x, θ, z̄, j̄=None, flags: int = 0, grad: Optional[torch.Tensor] = None, want_x̄: bool = True,
            want_θ̄: bool = False):
        """Pullback of `backward(chain, x, θ) -> (z, ln_det_jac)` for caller cotangents (z̄, j̄): what
        ChainRulesCore.rrule(::typeof(backward), chain, x, θ) returns (src/affine/RNVP.jl:99-147 through the chain).
        Returns (grad, x̄, θ̄): grad is the packed-layout parameter cotangent (accumulated into `grad` if given)."""
        x, θ = self._prep(x, θ)
        B = n_samples(x)
        z̄ = to_jl(z̄, self.device)
        if tuple(z̄.shape) != tuple(x.shape):
            raise ValueError(f"z̄ must have the size of x: {tuple(z̄.shape)} vs {tuple(x.shape)}")
        if j̄ is not None:
            j̄ = to_jl(j̄, self.device)
            if int(j̄.numel()) != B:
                raise ValueError("j̄ must have one entry per sample (size(x)[2:N])")
        if grad is None:
            grad = torch.zeros(max(self.P, 1), device=self.device, dtype=torch.float32)
        x̄ = jl_empty(x.shape, self.device) if want_x̄ else None
        θ̄ = jl_empty(θ.shape, self.device) if (want_θ̄ and θ is not None) else None
        if B == 0:
            return grad, x̄, θ̄
        ws = self._workspace(B)
        with torch.cuda.device(self.device):
            L.check(L.lib().dflow_vjp(self.handle, self.W.data_ptr(), self._ptr(x), self._ptr(θ), B, flags, self._ptr(z̄),
                                      self._ptr(j̄), grad.data_ptr(), self._ptr(x̄), self._ptr(θ̄), ws.data_ptr(),
                                      ws.numel(), self._stream()))
        return grad, x̄, θ̄

    def adam_step(self, grad: torch.Tensor, m: torch.Tensor, v: torch.Tensor, t: int, lr=1e-3, β=(0.9, 0.999),
                  ϵ=1e-8) -> None:
        with torch.cuda.device(self.device):
            L.check(L.lib().dflow_adam_step(self.W.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), self.P,
                                            lr, β[0], β[1], ϵ, t, self._stream()))


def _train_epoch(self, x, θ, order: torch.Tensor, batchsize: int, m: torch.Tensor, v: torch.Tensor, t: int, lr=1e-3,
                 β=(0.9, 0.999), ϵ=1e-8, flags: int = 0, scratch: Optional[torch.Tensor] = None,
                 loss2: Optional[torch.Tensor] = None) -> int:
    """One epoch of minibatch steps enqueued from C (dflow_train_epoch): `order` = the epoch's shuffled training indices
    (int32, device).  Returns the new Adam step count.  Same kernels and arithmetic as loss_grad + adam_step per batch."""
    x, θ = self._prep(x, θ)
    n = int(order.numel())
    if n == 0:
        return t
    need = int(L.lib().dflow_workspace_bytes(self.handle, min(int(batchsize), n)))
    if self._ws is None or self._ws.numel() < need:
        self._ws = torch.empty(need, device=self.device, dtype=torch.uint8)
    if scratch is None:
        scratch = torch.empty(max(self.P, 1) + 2, device=self.device, dtype=torch.float32)
    tio = C.c_int64(int(t))
    with torch.cuda.device(self.device):
        L.check(L.lib().dflow_train_epoch(self.handle, self.W.data_ptr(), m.data_ptr(), v.data_ptr(), self._ptr(x),
                                          self._ptr(θ), order.data_ptr(), n, int(batchsize), lr, β[0], β[1], ϵ,
                                          C.byref(tio), flags, scratch.data_ptr(),
                                          None if loss2 is None else loss2.data_ptr(), self._ws.data_ptr(),
                                          self._ws.numel(), self._stream()))
    return int(tio.value)


PackedChain.train_epoch = _train_epoch


def _packed_of(elem: FlowElement, device=None) -> PackedChain:
    p = elem._packed
    if p is None or (device is not None and torch.device(device) != p.device):
        p = PackedChain(elem._leaves(), device)
        elem._packed = p
    return p


# ------------------------------------------------------------------------------------------------------------
# generic functions -- src/Chains.jl:33-72 (+ θ-less methods of @unconditional_wrapper, src/Macros.jl:126-128)
# ------------------------------------------------------------------------------------------------------------


def _is_flow(obj) -> bool:
    from .flows import Flow

    return isinstance(obj, Flow)


def backward(f, x, θ=None):
    """backward(f, x[, θ]) -> (f⁻¹(x|θ), ln|det J|): the normalising direction x -> z (src/Chains.jl:33-44,149-164).
    For a `Flow`, θ is first normalised with the flow's θ range (src/Macros.jl:104-112)."""
    if _is_flow(f):
        return f.packed().normalize(x, θ, L.THETA_NORMALIZE if f.n > 0 else 0)
    return _packed_of(f).normalize(x, θ)


def forward(f, z, θ=None):
    """forward(f, z[, θ]) -> (f(z|θ), ln|det J|): the sampling direction z -> x (src/Chains.jl:47-58,167-183)."""
    if _is_flow(f):
        return f.packed().forward_ldj(z, θ, L.THETA_NORMALIZE if f.n > 0 else 0)
    return _packed_of(f).forward_ldj(z, θ)


def forward_(f, z, θ=None) -> None:
    """forward!(f, z[, θ]): replace z by f(z|θ) in place, no ln_det_jac (src/Chains.jl:61-72,187-197)."""
    if _is_flow(f):
        f.packed().sample_inplace(z, θ, None, L.THETA_NORMALIZE if f.n > 0 else 0)
    else:
        _packed_of(f).sample_inplace(z, θ)
    return None
