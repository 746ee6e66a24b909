"""Test helpers: build the SAME chain in the oracle and in the product with identical weights."""
import numpy as np
import torch

import densityflows.jl_b200 as df
from oracle import dflow_oracle as O


def _net_from_oracle(net):
    return df.Chain(*[df.Dense(dl.W.shape[1], dl.W.shape[0], dl.act, bias=dl.b is not None,
                               weight=torch.from_numpy(np.ascontiguousarray(dl.W)),
                               bias_value=None if dl.b is None else torch.from_numpy(dl.b.copy()))
                      for dl in net])


def _axes_from_oracle(a):
    return df.CouplingAxes(None, _raw=(a.d, a.n, list(a.axis_id), list(a.axis_af), list(a.axis_nn)))


def elem_from_oracle(e):
    if e.kind == "rnvp":
        return df.RNVPCouplingLayer(_net_from_oracle(e.s_net), _net_from_oracle(e.t_net), _axes_from_oracle(e.axes))
    if e.kind == "nice":
        return df.NICECouplingLayer(_net_from_oracle(e.t_net), _axes_from_oracle(e.axes))
    if e.kind == "norm":
        return df.NormalizationLayer(e.x_min, e.x_max, e.alpha, e.beta)
    if e.kind == "block":
        return df.CouplingBlock(elem_from_oracle(e.layer_1), elem_from_oracle(e.layer_2))
    if e.kind == "chain":
        return chain_from_oracle(e)
    raise TypeError(e)


def chain_from_oracle(ochain) -> "df.FlowChain":
    return df.FlowChain(*[elem_from_oracle(e) for e in ochain.layers])


def assert_close(got, want, rtol, atol, what=""):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    err = np.abs(got - want)
    tol = atol + rtol * np.abs(want)
    bad = err > tol
    assert not bad.any(), (f"{what}: {bad.sum()} / {bad.size} outside rtol={rtol} atol={atol}; "
                           f"max err {err.max():.3e} at want={want.flat[err.argmax()]:.6e}")
