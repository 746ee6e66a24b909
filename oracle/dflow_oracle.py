"""CPU oracle for the DensityFlows.jl coupling-chain hot path.  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement (float32 = parity target, float64 = truth arbiter) of the
reference's algorithm for the path named in BASELINE.json `north_star`.  It is imported only by
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs, as
the *checker*.  The product (`densityflows.jl_b200`) never imports it and has no CPU fallback.

PARITY UNPINNED for numerical values: the reference is pure Julia (Flux/Zygote/Optimisers/
Distributions), Julia is not installed here or on the GPU box, and the reference's tests hold no
known-answer vectors -- only self-consistency properties (round trip, ldj antisymmetry, index-set
equality, normalisation range, output shape; test/runtests.jl:7-121).  Those properties and the
fixture test/datatest.jld2 are what this oracle is pinned against (tests/test_oracle_reference.py).

Third-party arithmetic that is NOT under /root/reference (no Manifest.toml; Project.toml:16-26
only gives [compat] lower bounds) is restated from the packages' published behaviour:
  * Flux 0.16 `Dense`: y = sigma.(W*x .+ b), W is (out,in), glorot_uniform init
    U(+-sqrt(6/(fan_in+fan_out))), zero bias; `Chain` = composition; relu subgradient at 0 is 0.
  * Optimisers 0.4 `Adam(eta, (0.9,0.999), 1e-8)`:
      m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ;
      W <- W - eta * (m/(1-b1^t)) / (sqrt(v/(1-b2^t)) + eps)
  * Distributions 0.25 `MvNormal(0, I_d)`: logpdf(z) = -(d*log(2pi))/2 - |z|^2/2.
  * MLUtils `DataLoader`: batches along the last dim, reshuffled each epoch, partial last batch kept.

Array convention: every array has the reference's Julia shape `(rows, B...)`; element (k, b) of a
Julia column-major array is `a[k, b]` here (memory order is irrelevant to the oracle).
Index vectors are kept 1-based exactly as the reference builds them (`Axes.af0` etc. give 0-based).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

LOG2PI = math.log(2.0 * math.pi)

# --------------------------------------------------------------------------------------------
# L1 integer bookkeeping -- src/Axes.jl
# --------------------------------------------------------------------------------------------


@dataclass
class Axes:
    """CouplingAxes struct, src/Axes.jl:28-37 (1-based integer vectors)."""

    d: int
    n: int
    axis_id: List[int]
    axis_af: List[int]
    axis_nn: List[int]

    @property
    def id0(self) -> np.ndarray:
        return np.asarray(self.axis_id, dtype=np.int64) - 1

    @property
    def af0(self) -> np.ndarray:
        return np.asarray(self.axis_af, dtype=np.int64) - 1

    @property
    def nn0(self) -> np.ndarray:
        return np.asarray(self.axis_nn, dtype=np.int64) - 1


def coupling_axes(d: int, mask: Sequence[int], n: int = 0) -> Axes:
    """CouplingAxes(d, mask; n), src/Axes.jl:79-102.

    axis_af = mask (caller order kept), axis_id = findall(x -> !(x in mask), 1:d) (ascending),
    axis_nn = vcat(1:n, axis_id .+ n).
    """
    mask = [int(m) for m in mask]
    assert max(mask) <= d, "The mask cannot contain values higher than the dimension"  # Axes.jl:85
    axis_id = [k for k in range(1, d + 1) if k not in mask]
    axis_af = list(mask)
    axis_nn = list(range(1, n + 1)) + [k + n for k in axis_id]
    return Axes(d, n, axis_id, axis_af, axis_nn)


def coupling_axes_cut(d: int, j: Optional[int] = None, n: int = 0, reverse: bool = False) -> Axes:
    """CouplingAxes(d, j=d÷2; n, reverse), src/Axes.jl:104-114."""
    if j is None:
        j = d // 2
    mask = list(range(j + 1, d + 1)) if not reverse else list(range(1, j + 1))
    return coupling_axes(d, mask, n=n)


def reverse_axes(a: Axes) -> Axes:
    """Base.reverse(axes), src/Axes.jl:129-135: swap id/af, rebuild axis_nn from the old axis_af."""
    axis_nn = list(range(1, a.n + 1)) + [k + a.n for k in a.axis_af]
    return Axes(a.d, a.n, list(a.axis_af), list(a.axis_id), axis_nn)


def is_reverse(a1: Axes, a2: Axes) -> bool:
    """src/Axes.jl:137-139 (element-wise comparison, so lengths must agree)."""
    if len(a1.axis_af) != len(a2.axis_id) or len(a2.axis_af) != len(a1.axis_id):
        return False
    return (list(a1.axis_af) == list(a2.axis_id)) and (list(a2.axis_af) == list(a1.axis_id)) and a1.n == a2.n


def axes_equal(x: Axes, y: Axes) -> bool:
    """==(::CouplingAxes, ::CouplingAxes), src/Axes.jl:46-56: compares sorted index sets."""
    return (
        x.d == y.d
        and x.n == y.n
        and sorted(x.axis_id) == sorted(y.axis_id)
        and sorted(x.axis_af) == sorted(y.axis_af)
        and sorted(x.axis_nn) == sorted(y.axis_nn)
    )


# --------------------------------------------------------------------------------------------
# Conditioner networks -- src/Layers.jl:33-50 (+ Flux Dense semantics)
# --------------------------------------------------------------------------------------------

ACT_IDENTITY, ACT_RELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3
_ACT_NAMES = {"identity": ACT_IDENTITY, "relu": ACT_RELU, "tanh": ACT_TANH, "sigmoid": ACT_SIGMOID}


def act_code(a: Union[int, str]) -> int:
    return _ACT_NAMES[a] if isinstance(a, str) else int(a)


def apply_act(code: int, v: np.ndarray) -> np.ndarray:
    if code == ACT_IDENTITY:
        return v
    if code == ACT_RELU:
        return np.maximum(v, v.dtype.type(0))
    if code == ACT_TANH:
        return np.tanh(v)
    if code == ACT_SIGMOID:
        one = v.dtype.type(1)
        return one / (one + np.exp(-v))
    raise ValueError(code)


def act_grad_from_output(code: int, y: np.ndarray) -> np.ndarray:
    """d act / d pre-activation expressed through the OUTPUT y (relu'(0) = 0, NNlib convention)."""
    one = y.dtype.type(1)
    if code == ACT_IDENTITY:
        return np.ones_like(y)
    if code == ACT_RELU:
        return (y > 0).astype(y.dtype)
    if code == ACT_TANH:
        return one - y * y
    if code == ACT_SIGMOID:
        return y * (one - y)
    raise ValueError(code)


@dataclass
class Dense:
    """Flux.Dense: weight (out,in), optional bias (out,), activation code."""

    W: np.ndarray
    b: Optional[np.ndarray]
    act: int


def glorot_uniform(rng: np.random.Generator, out_dim: int, in_dim: int) -> np.ndarray:
    """Flux.glorot_uniform: U(-s, s), s = sqrt(24/(fan_in+fan_out))/2 = sqrt(6/(in+out)); Float32."""
    s = math.sqrt(6.0 / (in_dim + out_dim))
    return ((rng.random((out_dim, in_dim)) * 2.0 - 1.0) * s).astype(np.float32)


def dflt_net(
    input_dim: int,
    output_dim: int,
    n: int,
    hidden_dim: int = 32,
    act: Union[int, str] = "relu",
    bias: bool = True,
    rng: Optional[np.random.Generator] = None,
    bias_scale: float = 0.0,
) -> List[Dense]:
    """_dflt_net, src/Layers.jl:33-50: Dense(in,h,act), (n-1) x Dense(h,h,act), Dense(h,out,identity).

    Flux initialises biases to zero; `bias_scale` > 0 draws U(-bias_scale, bias_scale) instead
    (synthetic test/bench inputs only, so that bias paths are exercised).
    """
    rng = rng or np.random.default_rng(0)
    code = act_code(act)
    dims = [(input_dim, hidden_dim, code)] + [(hidden_dim, hidden_dim, code)] * (n - 1) + [(hidden_dim, output_dim, ACT_IDENTITY)]
    net = []
    for fi, fo, c in dims:
        W = glorot_uniform(rng, fo, fi)
        if bias:
            b = ((rng.random(fo) * 2.0 - 1.0) * bias_scale).astype(np.float32) if bias_scale > 0 else np.zeros(fo, np.float32)
        else:
            b = None
        net.append(Dense(W, b, c))
    return net


def net_apply(net: List[Dense], inp: np.ndarray, dtype=np.float32, keep: bool = False):
    """Flux.Chain of Dense: y = act.(W*x .+ b).  `keep` returns every layer's output (for the adjoint)."""
    h = inp.astype(dtype, copy=False)
    outs = []
    for dl in net:
        h = dl.W.astype(dtype) @ h
        if dl.b is not None:
            h = h + dl.b.astype(dtype)[:, None]
        h = apply_act(dl.act, h)
        outs.append(h)
    return (h, outs) if keep else h


# --------------------------------------------------------------------------------------------
# Flow elements -- src/affine/RNVP.jl, src/affine/NICE.jl, src/norm/Normalization.jl
# --------------------------------------------------------------------------------------------


@dataclass
class RNVPLayer:
    """RNVPCouplingLayer, src/affine/RNVP.jl:41-48."""

    s_net: List[Dense]
    t_net: List[Dense]
    axes: Axes
    kind: str = "rnvp"


@dataclass
class NICELayer:
    """NICECouplingLayer (s == 0), src/affine/NICE.jl."""

    t_net: List[Dense]
    axes: Axes
    kind: str = "nice"


@dataclass
class NormLayer:
    """NormalizationLayer, src/norm/Normalization.jl:30-35."""

    x_min: np.ndarray
    x_max: np.ndarray
    alpha: float
    beta: float
    kind: str = "norm"


@dataclass
class Block:
    """CouplingBlock, src/Blocks.jl:64-75 (layer_2.axes must be reverse(layer_1.axes))."""

    layer_1: Union[RNVPLayer, NICELayer]
    layer_2: Union[RNVPLayer, NICELayer]
    kind: str = "block"

    def __post_init__(self):
        if not is_reverse(self.layer_1.axes, self.layer_2.axes):  # Blocks.jl:70-73
            raise ValueError("layer_1 and layer_2 need to have complementary axes")


def coupling_layer(
    axes: Axes,
    kind: str = "rnvp",
    n_sublayers_t: int = 2,
    n_sublayers_s: int = 2,
    hidden_dim_t: int = 32,
    hidden_dim_s: int = 32,
    act_t: Union[int, str] = "relu",
    act_s: Union[int, str] = "relu",
    bias: bool = True,
    rng: Optional[np.random.Generator] = None,
    bias_scale: float = 0.0,
    s_out_scale: float = 1.0,
):
    """CouplingLayer(T, axes; kws...), src/Layers.jl:113-136 (t_net is built first, then s_net)."""
    rng = rng or np.random.default_rng(0)
    input_dim = len(axes.axis_nn)
    output_dim = len(axes.axis_af)
    t_net = dflt_net(input_dim, output_dim, n_sublayers_t, hidden_dim_t, act_t, bias, rng, bias_scale)
    if kind == "nice":
        return NICELayer(t_net, axes)
    s_net = dflt_net(input_dim, output_dim, n_sublayers_s, hidden_dim_s, act_s, bias, rng, bias_scale)
    if s_out_scale != 1.0:  # synthetic-only: keeps |s| small over deep chains
        s_net[-1].W = (s_net[-1].W * np.float32(s_out_scale)).astype(np.float32)
        if s_net[-1].b is not None:
            s_net[-1].b = (s_net[-1].b * np.float32(s_out_scale)).astype(np.float32)
    return RNVPLayer(s_net, t_net, axes)


def coupling_block(first_axes: Axes, **kws) -> Block:
    """CouplingBlock(T, first_axes; kws...), src/Blocks.jl:88-103."""
    second_axes = reverse_axes(first_axes)
    l1 = coupling_layer(first_axes, **kws)
    l2 = coupling_layer(second_axes, **kws)
    return Block(l1, l2)


def norm_layer_from_data(x: np.ndarray, alpha: float = 0.0, beta: float = 1.0) -> NormLayer:
    """NormalizationLayer(x, α, β), src/norm/Normalization.jl:51-57: min/max over dims 2:N."""
    flat = x.reshape(x.shape[0], -1)
    assert beta > alpha, "Bounds of the normalisation need to be in the correct order"
    return NormLayer(flat.min(axis=1).astype(np.float32), flat.max(axis=1).astype(np.float32), float(alpha), float(beta))


def _nn_input(axes: Axes, x: np.ndarray, theta: np.ndarray) -> np.ndarray:
    """selectdim(vcat(θ, x), 1, axis_nn), src/affine/RNVP.jl:157."""
    return np.concatenate([theta, x], axis=0)[axes.nn0]


def _flat2(a: np.ndarray) -> Tuple[np.ndarray, Tuple[int, ...]]:
    return a.reshape(a.shape[0], int(np.prod(a.shape[1:], dtype=np.int64))), a.shape[1:]


def rnvp_backward(layer: RNVPLayer, x: np.ndarray, theta: np.ndarray, dtype=np.float32):
    """backward(::RNVPCouplingLayer) + RNVP_backward, src/affine/RNVP.jl:150-165, :77-96.

    z[id] = x[id]; z[af] = (x[af] - t) .* exp.(-s); ln_det_jac = -sum(s, dims=1).
    """
    x2, tail = _flat2(x.astype(dtype, copy=False))
    th2, _ = _flat2(theta.astype(dtype, copy=False))
    inp = _nn_input(layer.axes, x2, th2)
    s = net_apply(layer.s_net, inp, dtype)
    t = net_apply(layer.t_net, inp, dtype)
    ldj = -s.sum(axis=0, dtype=dtype)
    z = np.empty_like(x2)
    z[layer.axes.id0] = x2[layer.axes.id0]
    z[layer.axes.af0] = (x2[layer.axes.af0] - t) * np.exp(-s)
    return z.reshape(x.shape), ldj.reshape(tail)


def rnvp_forward(layer: RNVPLayer, z: np.ndarray, theta: np.ndarray, dtype=np.float32):
    """forward(::RNVPCouplingLayer), src/affine/RNVP.jl:168-187: x[af] = z[af].*exp.(s) .+ t, ldj=+sum(s)."""
    z2, tail = _flat2(z.astype(dtype, copy=False))
    th2, _ = _flat2(theta.astype(dtype, copy=False))
    inp = _nn_input(layer.axes, z2, th2)
    s = net_apply(layer.s_net, inp, dtype)
    t = net_apply(layer.t_net, inp, dtype)
    ldj = s.sum(axis=0, dtype=dtype)
    x = np.empty_like(z2)
    x[layer.axes.id0] = z2[layer.axes.id0]
    x[layer.axes.af0] = z2[layer.axes.af0] * np.exp(s) + t
    return x.reshape(z.shape), ldj.reshape(tail)


def nice_backward(layer: NICELayer, x, theta, dtype=np.float32):
    """src/affine/NICE.jl:63-81,118-133: z[af] = x[af] - t, ldj = 0."""
    x2, tail = _flat2(x.astype(dtype, copy=False))
    th2, _ = _flat2(theta.astype(dtype, copy=False))
    t = net_apply(layer.t_net, _nn_input(layer.axes, x2, th2), dtype)
    z = x2.copy()
    z[layer.axes.af0] = x2[layer.axes.af0] - t
    return z.reshape(x.shape), np.zeros(tail, dtype)


def nice_forward(layer: NICELayer, z, theta, dtype=np.float32):
    """src/affine/NICE.jl:136-155: x[af] = z[af] + t, ldj = 0."""
    z2, tail = _flat2(z.astype(dtype, copy=False))
    th2, _ = _flat2(theta.astype(dtype, copy=False))
    t = net_apply(layer.t_net, _nn_input(layer.axes, z2, th2), dtype)
    x = z2.copy()
    x[layer.axes.af0] = z2[layer.axes.af0] + t
    return x.reshape(z.shape), np.zeros(tail, dtype)


def norm_ldj_const(nl: NormLayer, dtype=np.float32):
    """sum(log.(x_diff ./ δ)), src/norm/Normalization.jl:73 (sign applied by the caller)."""
    x_diff = nl.x_max.astype(dtype) - nl.x_min.astype(dtype)
    delta = dtype(nl.beta) - dtype(nl.alpha)
    return np.sum(np.log(x_diff / delta), dtype=dtype)


def norm_backward(nl: NormLayer, x, theta=None, dtype=np.float32):
    """backward(::NormalizationLayer), src/norm/Normalization.jl:64-77."""
    x2, tail = _flat2(x.astype(dtype, copy=False))
    xmin = nl.x_min.astype(dtype)[:, None]
    xmax = nl.x_max.astype(dtype)[:, None]
    a, b = dtype(nl.alpha), dtype(nl.beta)
    z = (b * (x2 - xmin) + a * (xmax - x2)) / (xmax - xmin)
    ldj = -norm_ldj_const(nl, dtype) * np.ones(tail, dtype)
    return z.reshape(x.shape), ldj


def norm_forward(nl: NormLayer, z, theta=None, dtype=np.float32):
    """forward(::NormalizationLayer), src/norm/Normalization.jl:79-92 (forward! :95-103 is the same map)."""
    z2, tail = _flat2(z.astype(dtype, copy=False))
    xmin = nl.x_min.astype(dtype)[:, None]
    xmax = nl.x_max.astype(dtype)[:, None]
    a, b = dtype(nl.alpha), dtype(nl.beta)
    x = ((xmax - xmin) * z2 - a * xmax + b * xmin) / (b - a)
    ldj = norm_ldj_const(nl, dtype) * np.ones(tail, dtype)
    return x.reshape(z.shape), ldj


# --------------------------------------------------------------------------------------------
# Composition -- src/Chains.jl:149-197, src/Blocks.jl:127-161
# --------------------------------------------------------------------------------------------

Element = Union[RNVPLayer, NICELayer, NormLayer, Block, "Chain"]


@dataclass
class Chain:
    """FlowChain, src/Chains.jl:78-80."""

    layers: List[Element] = field(default_factory=list)
    kind: str = "chain"


def concatenate(*xs) -> Chain:
    """concatenate, src/Chains.jl:112-123 (chains are spliced, bare elements appended)."""
    out: List[Element] = []
    for x in xs:
        if isinstance(x, (tuple, list)):
            out.extend(concatenate(*x).layers)
        elif isinstance(x, Chain):
            out.extend(x.layers)
        else:
            out.append(x)
    return Chain(out)


def elem_backward(e: Element, x, theta, dtype=np.float32):
    if e.kind == "rnvp":
        return rnvp_backward(e, x, theta, dtype)
    if e.kind == "nice":
        return nice_backward(e, x, theta, dtype)
    if e.kind == "norm":
        return norm_backward(e, x, theta, dtype)
    if e.kind == "block":  # Blocks.jl:127-137: layer_2 first, then layer_1
        y, l2 = elem_backward(e.layer_2, x, theta, dtype)
        z, l1 = elem_backward(e.layer_1, y, theta, dtype)
        return z, l1 + l2
    if e.kind == "chain":
        return chain_backward(e, x, theta, dtype)
    raise TypeError(e)


def elem_forward(e: Element, z, theta, dtype=np.float32):
    if e.kind == "rnvp":
        return rnvp_forward(e, z, theta, dtype)
    if e.kind == "nice":
        return nice_forward(e, z, theta, dtype)
    if e.kind == "norm":
        return norm_forward(e, z, theta, dtype)
    if e.kind == "block":  # Blocks.jl:140-150: layer_1 first, then layer_2
        y, l1 = elem_forward(e.layer_1, z, theta, dtype)
        x, l2 = elem_forward(e.layer_2, y, theta, dtype)
        return x, l1 + l2
    if e.kind == "chain":
        return chain_forward(e, z, theta, dtype)
    raise TypeError(e)


def chain_backward(chain: Chain, x, theta, dtype=np.float32):
    """backward(::FlowChain), src/Chains.jl:149-164: elements LAST -> FIRST, ldj summed."""
    n = len(chain.layers)
    x_i, ldj = elem_backward(chain.layers[n - 1], x, theta, dtype)
    for i in range(2, n + 1):
        x_i, ldj_i = elem_backward(chain.layers[n - i], x_i, theta, dtype)
        ldj = ldj + ldj_i
    return x_i, ldj


def chain_forward(chain: Chain, z, theta, dtype=np.float32):
    """forward(::FlowChain), src/Chains.jl:167-183: elements FIRST -> LAST, ldj summed.

    forward! (src/Chains.jl:187-197) applies the same maps in place without the ldj.
    """
    z_i, ldj = elem_forward(chain.layers[0], z, theta, dtype)
    for e in chain.layers[1:]:
        z_i, ldj_i = elem_forward(e, z_i, theta, dtype)
        ldj = ldj + ldj_i
    return z_i, ldj


def flatten(chain: Union[Chain, Element]) -> List[Element]:
    """Chain order list of leaf elements (blocks -> layer_1, layer_2; nested chains spliced).

    Equivalent evaluation order: normalising = reversed(flatten), sampling = flatten.
    """
    out: List[Element] = []
    if chain.kind == "chain":
        for e in chain.layers:
            out.extend(flatten(e))
    elif chain.kind == "block":
        out.extend([chain.layer_1, chain.layer_2])
    else:
        out.append(chain)
    return out


# --------------------------------------------------------------------------------------------
# Data helpers -- src/Data.jl
# --------------------------------------------------------------------------------------------


def dflt_theta(x: np.ndarray) -> np.ndarray:
    """dflt_θ(x), src/Data.jl:57-65: a 0 x dims array."""
    return np.empty((0,) + tuple(x.shape[1:]), dtype=x.dtype)


def normalize_input(x: np.ndarray, x_min: np.ndarray, x_max: np.ndarray) -> np.ndarray:
    """normalize_input, src/Data.jl:213-218: (x - min)/(max - min), rows with zero range -> 0."""
    dt = x.dtype.type
    shp = (-1,) + (1,) * (x.ndim - 1)
    x_diff = (x_max - x_min).astype(x.dtype)
    with np.errstate(divide="ignore", invalid="ignore"):
        y = (x - x_min.astype(x.dtype).reshape(shp)) / x_diff.reshape(shp)
    y[x_diff == dt(0)] = dt(0)
    return y


def resize_output(y, x_min, x_max):
    """resize_output, src/Data.jl:231."""
    shp = (-1,) + (1,) * (y.ndim - 1)
    return (x_max - x_min).reshape(shp) * y + x_min.reshape(shp)


def data_partition(n: int, f_training: float = 0.9, f_validation: float = 0.1, rng: Optional[np.random.Generator] = None):
    """DataPartition, src/Data.jl:112-128: random permutation cut at round(n*f) (Julia round = ties-to-even).

    Returns 0-based index vectors.  (Julia's `randperm` stream itself cannot be reproduced.)
    """
    rng = rng or np.random.default_rng()
    p = rng.permutation(n)
    i1 = int(np.rint(n * f_training))
    i2 = i1 + int(np.rint(n * f_validation))
    return p[:i1], p[i1:i2], p[i2:n]


# --------------------------------------------------------------------------------------------
# Density / loss -- src/Flows.jl:114,272-281,352-359
# --------------------------------------------------------------------------------------------


def mvnormal_logpdf(z: np.ndarray, dtype=np.float32) -> np.ndarray:
    """Distributions.logpdf(MvNormal(zeros(d), I), z) per column: -(d*log2π)/2 - |z|²/2."""
    z2, tail = _flat2(z.astype(dtype, copy=False))
    d = z2.shape[0]
    c0 = dtype(-(d * dtype(LOG2PI)) / 2)
    return (c0 - np.sum(z2 * z2, axis=0, dtype=dtype) / dtype(2)).reshape(tail)


def logpdf(chain: Chain, x, theta, theta_min=None, theta_max=None, dtype=np.float32):
    """logpdf(flow, x, θ), src/Flows.jl:272-281 (θ normalised by the Flow wrapper, src/Macros.jl:104-112)."""
    th = theta if theta_min is None else normalize_input(theta.astype(dtype), theta_min, theta_max)
    z, ldj = chain_backward(chain, x, th, dtype)
    return mvnormal_logpdf(z, dtype) + ldj


def loss_value(z, ldj, dtype=np.float32):
    """loss, src/Flows.jl:352-359: -mean(logpdf(base, z) .+ ln_det_jac)."""
    return -np.mean(mvnormal_logpdf(z, dtype) + ldj.astype(dtype), dtype=dtype)


# --------------------------------------------------------------------------------------------
# Parameter packing (the C-ABI buffer layout; include/dflow.h)
# --------------------------------------------------------------------------------------------


def _trainable_nets(e) -> List[List[Dense]]:
    if e.kind == "rnvp":
        return [e.s_net, e.t_net]  # Flux.@layer trainable=(s_net, t_net), RNVP.jl:51
    if e.kind == "nice":
        return [e.t_net]
    return []


def pack_params(chain) -> np.ndarray:
    """Flat float32 vector: chain order; per layer s_net then t_net; per Dense vec(weight) (column-major) then bias."""
    parts = []
    for e in flatten(chain):
        for net in _trainable_nets(e):
            for dl in net:
                parts.append(np.asarray(dl.W, np.float32).reshape(-1, order="F"))
                if dl.b is not None:
                    parts.append(np.asarray(dl.b, np.float32))
    return np.concatenate(parts) if parts else np.zeros(0, np.float32)


def unpack_params(chain, flat: np.ndarray) -> None:
    """Inverse of pack_params (in place)."""
    off = 0
    for e in flatten(chain):
        for net in _trainable_nets(e):
            for dl in net:
                o, i = dl.W.shape
                dl.W = np.asarray(flat[off : off + o * i], np.float32).reshape((o, i), order="F").copy()
                off += o * i
                if dl.b is not None:
                    dl.b = np.asarray(flat[off : off + o], np.float32).copy()
                    off += o
    assert off == flat.size


# --------------------------------------------------------------------------------------------
# Adjoint through a chain -- rrule of RNVP_backward (src/affine/RNVP.jl:99-147) + Dense pullbacks
# --------------------------------------------------------------------------------------------


def _net_backward(net: List[Dense], inp: np.ndarray, outs: List[np.ndarray], gout: np.ndarray, dtype):
    """Pullback of a Dense chain: returns (input cotangent, [(dW, db), ...])."""
    grads = []
    delta = gout
    for li in range(len(net) - 1, -1, -1):
        dl = net[li]
        delta = delta * act_grad_from_output(dl.act, outs[li])
        hin = inp if li == 0 else outs[li - 1]
        dW = delta @ hin.T
        db = delta.sum(axis=1, dtype=dtype) if dl.b is not None else None
        grads.append((dW, db))
        delta = dl.W.astype(dtype).T @ delta
    grads.reverse()
    return delta, grads


def chain_loss_and_grad(chain, x, theta, dtype=np.float64, inv_btot: Optional[float] = None):
    """loss = -mean(logpdf(base,z) + ldj) and its gradient w.r.t. the packed parameters.

    Closed form: seeds z̄ = z/B, j̄ = -1/B (Flows.jl:352-359); per RNVP layer the pullback of
    src/affine/RNVP.jl:119-143 (s̄ = -z̄_af.*z_af - j̄, t̄ = -z̄_af.*exp(-s), ū_af = z̄_af.*exp(-s),
    ū_id = z̄_id) plus the conditioner-input cotangent (vcat/selectdim adjoint; θ rows discarded).
    Returns (loss, flat_grad, z, ldj).  `inv_btot` overrides 1/B (data-parallel shards).
    """
    elems = flatten(chain)
    x2, _ = _flat2(np.asarray(x, dtype))
    th2, _ = _flat2(np.asarray(theta, dtype))
    B = x2.shape[1]
    ib = dtype(1.0 / B if inv_btot is None else inv_btot)
    # forward (normalising) sweep, last element first; remember each layer's input
    u = x2
    ldj = np.zeros(B, dtype)
    tape = []
    for e in reversed(elems):
        if e.kind == "norm":
            z, l = norm_backward(e, u, None, dtype)
            tape.append((e, None))
        elif e.kind in ("rnvp", "nice"):
            inp = _nn_input(e.axes, u, th2)
            if e.kind == "rnvp":
                s, s_outs = net_apply(e.s_net, inp, dtype, keep=True)
            else:
                s, s_outs = np.zeros((len(e.axes.axis_af), B), dtype), None
            t, t_outs = net_apply(e.t_net, inp, dtype, keep=True)
            z = u.copy()
            z[e.axes.af0] = (u[e.axes.af0] - t) * np.exp(-s)
            l = -s.sum(axis=0, dtype=dtype)
            tape.append((e, (inp, s, s_outs, t, t_outs, z)))
        else:
            raise TypeError(e)
        ldj = ldj + l
        u = z
    z_final = u
    logp = mvnormal_logpdf(z_final, dtype) + ldj
    loss = -np.sum(logp, dtype=dtype) * ib
    # reverse sweep (chain order = reverse of the normalising order)
    zbar = z_final * ib
    jbar = -ib
    grads_per_elem = {}
    for e, saved in reversed(tape):
        if e.kind == "norm":
            xdiff = (e.x_max.astype(dtype) - e.x_min.astype(dtype))[:, None]
            zbar = zbar * (dtype(e.beta) - dtype(e.alpha)) / xdiff
            continue
        inp, s, s_outs, t, t_outs, z = saved
        af, idx = e.axes.af0, e.axes.id0
        em = np.exp(-s)
        zb_af = zbar[af]
        tbar = -zb_af * em
        ubar = np.zeros_like(zbar)
        ubar[af] = zb_af * em
        ubar[idx] = zbar[idx]
        nets_g = []
        n = e.axes.n
        if e.kind == "rnvp":
            sbar = -zb_af * z[af] - jbar
            in_s, g_s = _net_backward(e.s_net, inp, s_outs, sbar, dtype)
            ubar[idx] += in_s[n:]
            nets_g.append(g_s)
        in_t, g_t = _net_backward(e.t_net, inp, t_outs, tbar, dtype)
        ubar[idx] += in_t[n:]
        nets_g.append(g_t)
        grads_per_elem[id(e)] = nets_g
        zbar = ubar
    parts = []
    for e in elems:
        for net, g in zip(_trainable_nets(e), grads_per_elem.get(id(e), [])):
            for dl, (dW, db) in zip(net, g):
                parts.append(dW.reshape(-1, order="F"))
                if dl.b is not None:
                    parts.append(db)
    flat = np.concatenate(parts) if parts else np.zeros(0, dtype)
    return loss, flat.astype(dtype), z_final, ldj


# --------------------------------------------------------------------------------------------
# Optimiser -- Optimisers.jl Adam (call site src/Flows.jl:415)
# --------------------------------------------------------------------------------------------


def adam_step(w, g, m, v, t: int, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, dtype=np.float32):
    """One Optimisers.Adam update, step counter t >= 1.  Returns (w, m, v) new arrays."""
    w, g, m, v = (np.asarray(a, dtype) for a in (w, g, m, v))
    b1, b2, lr, eps = dtype(b1), dtype(b2), dtype(lr), dtype(eps)
    one = dtype(1)
    m = b1 * m + (one - b1) * g
    v = b2 * v + (one - b2) * (g * g)
    # Optimisers keeps βt as a running product in the parameter eltype: state βt .* β after every step
    b1t, b2t = dtype(1), dtype(1)
    for _ in range(t):
        b1t = dtype(b1t * b1)
        b2t = dtype(b2t * b2)
    step = m / (one - b1t) / (np.sqrt(v / (one - b2t)) + eps) * lr
    return w - step, m, v


def train(chain, x_train, th_train, x_valid, th_valid, epochs, batchsize=64, shuffle=True, lr=1e-3,
          rng: Optional[np.random.Generator] = None, dtype=np.float32, perms: Optional[list] = None):
    """train!, src/Flows.jl:380-445, on already-normalised θ: minibatch Adam + epoch-end full-set losses.

    `perms` (list of index arrays, one per epoch) pins the shuffle so that a device run can be compared.
    Returns (train_losses, valid_losses).
    """
    rng = rng or np.random.default_rng(0)
    w = pack_params(chain)
    m = np.zeros_like(w)
    v = np.zeros_like(w)
    t = 0
    n = x_train.shape[1]
    tl, vl = [], []
    for ep in range(epochs):
        order = perms[ep] if perms is not None else (rng.permutation(n) if shuffle else np.arange(n))
        for b0 in range(0, n, batchsize):
            idx = order[b0 : b0 + batchsize]
            _, g, _, _ = chain_loss_and_grad(chain, x_train[:, idx], th_train[:, idx], dtype)
            t += 1
            w, m, v = adam_step(w, g, m, v, t, lr=lr, dtype=dtype)
            unpack_params(chain, w)
        z, ldj = chain_backward(chain, x_train, th_train, dtype)
        tl.append(float(loss_value(z, ldj, dtype)))
        z, ldj = chain_backward(chain, x_valid, th_valid, dtype)
        vl.append(float(loss_value(z, ldj, dtype)))
    return tl, vl


# --------------------------------------------------------------------------------------------
# Builders for the named configurations (BASELINE.json / SURVEY.md §8)
# --------------------------------------------------------------------------------------------


def readme_chain(n: int = 2, x_for_norm: Optional[np.ndarray] = None, seed: int = 42, hidden: int = 16,
                 bias_scale: float = 0.1) -> Chain:
    """README / test/runtests.jl:104-109 chain: d=5, masks [1,2,3],[3,4,5],[5,1,2], h=16, Norm(x,-1,1)."""
    layers = []
    for li, mask in enumerate(([1, 2, 3], [3, 4, 5], [5, 1, 2])):
        rng = np.random.default_rng(seed + li)
        layers.append(coupling_layer(coupling_axes(5, mask, n=n), hidden_dim_s=hidden, hidden_dim_t=hidden, rng=rng,
                                     bias_scale=bias_scale))
    if x_for_norm is not None:
        layers.append(norm_layer_from_data(x_for_norm, -1.0, 1.0))
    return Chain(layers)


def block_chain(d: int, n: int, n_layers: int, hidden: int, x_for_norm: Optional[np.ndarray] = None, seed: int = 42,
                bias_scale: float = 0.1, s_out_scale: float = 0.1) -> Chain:
    """C3-C5 fill: CouplingBlock(d, d÷2; n) x L/2 + NormalizationLayer(x,-1,1)."""
    assert n_layers % 2 == 0
    layers = []
    for bi in range(n_layers // 2):
        rng = np.random.default_rng(seed + bi)
        layers.append(coupling_block(coupling_axes_cut(d, d // 2, n=n), hidden_dim_s=hidden, hidden_dim_t=hidden, rng=rng,
                                     bias_scale=bias_scale, s_out_scale=s_out_scale))
    if x_for_norm is not None:
        layers.append(norm_layer_from_data(x_for_norm, -1.0, 1.0))
    return Chain(layers)


def synthetic_data(d: int, n: int, B: int, seed: int = 1234):
    """SURVEY.md §8d synthetic inputs: x_k = mu_k + sigma_k N(0,1), mu_k=0.1k, sigma_k=1+0.05k; θ ~ U(-1,2)."""
    rng = np.random.default_rng(seed)
    k = np.arange(d, dtype=np.float64)[:, None]
    x = ((0.1 * k) + (1.0 + 0.05 * k) * rng.standard_normal((d, B))).astype(np.float32)
    th = (rng.random((n, B)) * 3.0 - 1.0).astype(np.float32)
    return x, th
