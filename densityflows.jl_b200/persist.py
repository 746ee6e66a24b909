"""Persistence of flows on this side of the boundary -- counterpart of src/Loading.jl (save_flow / load_flow).

The reference stores a directory tree of JLD2 files holding Julia-serialised structs (Flux.state named tuples, CouplingAxes,
MetaData; src/Loading.jl:78-109,324-377).  Reading those needs Julia's own type machinery, so the interchange with a model
trained by the reference happens on the Julia side of the C ABI: `load_flow(dir)` (reference code) -> `pack!`
(julia/DensityFlowsB200.jl) moves every Dense into the packed device buffer, `unpack!` + `save_flow` goes back.  THIS module is
the persistence of the Python mirror: one `.npz` with the chain structure (JSON) and the packed parameter vector in the
C ABI's layout (chain order, s_net then t_net, vec(weight) column-major then bias) -- the same vector `pack!` produces, so a
`W` written by either side loads on the other.
"""
from __future__ import annotations

import json
from typing import List

import numpy as np
import torch

from .data import MetaData
from .model import (Chain, CouplingAxes, CouplingBlock, Dense, FlowChain, NICECouplingLayer, NormalizationLayer,
                    RNVPCouplingLayer, _nets_of)

_ACT_NAMES = {0: "identity", 1: "relu", 2: "tanh", 3: "sigmoid"}
FORMAT = 1


def _net_struct(net: Chain) -> dict:
    return {"widths": net.widths(), "acts": [_ACT_NAMES[l.act] for l in net.layers], "bias": net.layers[0].bias is not None}


def _elem_struct(e) -> dict:
    if isinstance(e, NormalizationLayer):
        return {"kind": "norm", "x_min": [float(v) for v in e.x_min], "x_max": [float(v) for v in e.x_max], "alpha": e.α,
                "beta": e.β}
    if isinstance(e, CouplingBlock):
        return {"kind": "block", "layers": [_elem_struct(e.layer_1), _elem_struct(e.layer_2)]}
    if isinstance(e, FlowChain):
        return {"kind": "chain", "layers": [_elem_struct(l) for l in e.layers]}
    a = e.axes
    out = {"kind": "rnvp" if isinstance(e, RNVPCouplingLayer) else "nice",
           "axes": {"d": a.d, "n": a.n, "axis_id": list(a.axis_id), "axis_af": list(a.axis_af), "axis_nn": list(a.axis_nn)},
           "t_net": _net_struct(e.t_net)}
    if isinstance(e, RNVPCouplingLayer):
        out["s_net"] = _net_struct(e.s_net)
    return out


def packed_parameters(chain: FlowChain) -> np.ndarray:
    """The chain's parameters in the packed C-ABI layout, gathered from the Dense tensors (host copy)."""
    parts: List[np.ndarray] = []
    for e in chain._leaves():
        for net in _nets_of(e):
            for l in net.layers:
                w = l.weight.detach().cpu().numpy()
                parts.append(np.asarray(w, np.float32).reshape(-1, order="F"))  # vec(weight), Flux (out, in) column-major
                if l.bias is not None:
                    parts.append(np.asarray(l.bias.detach().cpu().numpy(), np.float32).reshape(-1))
    return np.concatenate(parts) if parts else np.zeros(0, np.float32)


def save_flow(path: str, flow) -> None:
    """save_flow(directory, flow) analogue: structure + packed parameters + metadata + loss histories in one .npz."""
    md = flow.metadata
    struct = {"format": FORMAT, "model": _elem_struct(flow.model),
              "metadata": {"d": md.d, "n": md.n, "theta_min": [float(v) for v in md.θ_min],
                           "theta_max": [float(v) for v in md.θ_max]}}
    np.savez(path, structure=np.frombuffer(json.dumps(struct).encode(), np.uint8), W=packed_parameters(flow.model),
             train_loss=np.asarray(flow.train_loss, np.float64), valid_loss=np.asarray(flow.valid_loss, np.float64))


def _build_net(st: dict, w: np.ndarray, off: List[int]) -> Chain:
    layers = []
    for j, act in enumerate(st["acts"]):
        i, o = st["widths"][j], st["widths"][j + 1]
        W = torch.from_numpy(w[off[0]: off[0] + o * i].reshape(o, i, order="F").copy())
        off[0] += o * i
        b = None
        if st["bias"]:
            b = torch.from_numpy(w[off[0]: off[0] + o].copy())
            off[0] += o
        layers.append(Dense(i, o, act, bias=st["bias"], weight=W, bias_value=b))
    return Chain(*layers)


def _build_elem(st: dict, w: np.ndarray, off: List[int]):
    k = st["kind"]
    if k == "norm":
        return NormalizationLayer(np.asarray(st["x_min"], np.float32), np.asarray(st["x_max"], np.float32), st["alpha"], st["beta"])
    if k == "block":
        return CouplingBlock(*[_build_elem(s, w, off) for s in st["layers"]])
    if k == "chain":
        return FlowChain(*[_build_elem(s, w, off) for s in st["layers"]])
    a = st["axes"]
    axes = CouplingAxes(None, _raw=(a["d"], a["n"], list(a["axis_id"]), list(a["axis_af"]), list(a["axis_nn"])))
    if k == "rnvp":
        s_net = _build_net(st["s_net"], w, off)  # pack order: s_net before t_net
        return RNVPCouplingLayer(s_net, _build_net(st["t_net"], w, off), axes)
    return NICECouplingLayer(_build_net(st["t_net"], w, off), axes)


class _LoadedData:
    """Just enough of a DataArrays for Flow(): dimensions and the θ range saved with the flow."""

    def __init__(self, md: MetaData, device):
        self.x = torch.empty((md.d, 0), device=device)
        self.θ = torch.empty((md.n, 0), device=device)
        self._θ_range = (md.θ_min, md.θ_max)


def load_flow(path: str, device=None):
    """load_flow(directory) analogue; `device` is where the flow will run (default: the current CUDA device, or the CPU
    for structure-only work -- compute always needs the GPU)."""
    from .flows import Flow

    z = np.load(path if path.endswith(".npz") else path + ".npz")
    struct = json.loads(bytes(z["structure"]).decode())
    if struct.get("format") != FORMAT:
        raise ValueError(f"unknown flow file format {struct.get('format')}")
    w, off = np.asarray(z["W"], np.float32), [0]
    model = _build_elem(struct["model"], w, off)
    if off[0] != w.size:
        raise ValueError("parameter vector does not match the stored structure")
    if not isinstance(model, FlowChain):
        model = FlowChain(model)
    m = struct["metadata"]
    md = MetaData("", m["d"], m["n"], m["theta_min"], m["theta_max"])
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    flow = Flow(model, _LoadedData(md, torch.device(device)), list(z["train_loss"]), list(z["valid_loss"]))
    return flow
