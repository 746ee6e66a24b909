"""Flux-equivalent CPU restatement in torch (autograd) -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

Same op sequence as the reference, unfused, so that it can stand in for the Flux CPU path that
cannot run here (no Julia): per layer `vcat(θ,x)` -> row gather by axis_nn -> Dense chain for s_net
and again for t_net (GEMM + bias + activation each) -> exp, scatter into a fresh (d,B) array,
column sum (src/affine/RNVP.jl:77-96,150-205); chain loop (src/Chains.jl:149-197); Gaussian base
term and mean (src/Flows.jl:279,352-359); reverse-mode autograd for the train step (stands in for
Zygote + the rrule at src/affine/RNVP.jl:99-147); Adam with the Optimisers.jl formula.

Only `tests/`, `__graft_entry__.smoke()` and bench.py's cpu_baseline / `--impl reference` legs may
import this.  PARITY UNPINNED (see oracle/dflow_oracle.py header).
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch

from . import dflow_oracle as O

LOG2PI = math.log(2.0 * math.pi)


def _act(code: int, v: torch.Tensor) -> torch.Tensor:
    if code == O.ACT_IDENTITY:
        return v
    if code == O.ACT_RELU:
        return torch.relu(v)
    if code == O.ACT_TANH:
        return torch.tanh(v)
    if code == O.ACT_SIGMOID:
        return torch.sigmoid(v)
    raise ValueError(code)


class TorchChain:
    """Holds torch parameter tensors for the leaf elements of an oracle chain (chain order)."""

    def __init__(self, chain, dtype=torch.float32, requires_grad: bool = False):
        self.dtype = dtype
        self.elems = O.flatten(chain)
        self.params: List[torch.Tensor] = []  # flat list in pack order
        self.nets = []  # per element: list of nets; net = list of (W, b, act)
        for e in self.elems:
            enets = []
            for net in O._trainable_nets(e):
                tn = []
                for dl in net:
                    W = torch.tensor(np.asarray(dl.W), dtype=dtype, requires_grad=requires_grad)
                    self.params.append(W)
                    b = None
                    if dl.b is not None:
                        b = torch.tensor(np.asarray(dl.b), dtype=dtype, requires_grad=requires_grad)
                        self.params.append(b)
                    tn.append((W, b, dl.act))
                enets.append(tn)
            self.nets.append(enets)

    # -- parameter (un)packing, same layout as O.pack_params ---------------------------------
    def flat_grad(self) -> torch.Tensor:
        parts = []
        for p in self.params:
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            parts.append(g.t().reshape(-1) if g.dim() == 2 else g.reshape(-1))  # column-major vec(W)
        return torch.cat(parts) if parts else torch.zeros(0, dtype=self.dtype)

    def flat_params(self) -> torch.Tensor:
        parts = [(p.detach().t().reshape(-1) if p.dim() == 2 else p.detach().reshape(-1)) for p in self.params]
        return torch.cat(parts) if parts else torch.zeros(0, dtype=self.dtype)

    def set_flat_params(self, flat: torch.Tensor) -> None:
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                if p.dim() == 2:
                    p.copy_(flat[off : off + k].reshape(p.shape[1], p.shape[0]).t())
                else:
                    p.copy_(flat[off : off + k])
                off += k

    # -- element maps -----------------------------------------------------------------------
    @staticmethod
    def _net(tn, inp):
        h = inp
        for W, b, act in tn:
            h = W @ h
            if b is not None:
                h = h + b[:, None]
            h = _act(act, h)
        return h

    def backward(self, x: torch.Tensor, theta: torch.Tensor):
        """Normalising direction (reference `backward`): elements last -> first."""
        B = x.shape[1]
        ldj = torch.zeros(B, dtype=self.dtype)
        u = x
        for e, enets in zip(reversed(self.elems), reversed(self.nets)):
            if e.kind == "norm":
                xmin = torch.tensor(e.x_min, dtype=self.dtype)[:, None]
                xmax = torch.tensor(e.x_max, dtype=self.dtype)[:, None]
                xd = xmax - xmin
                u = (e.beta * (u - xmin) + e.alpha * (xmax - u)) / xd
                ldj = ldj - torch.sum(torch.log(xd / (e.beta - e.alpha))) * torch.ones(B, dtype=self.dtype)
                continue
            nn_idx = torch.as_tensor(e.axes.nn0)
            af = torch.as_tensor(e.axes.af0)
            idx = torch.as_tensor(e.axes.id0)
            inp = torch.cat([theta, u], dim=0)[nn_idx]
            if e.kind == "rnvp":
                s = self._net(enets[0], inp)
                t = self._net(enets[1], inp)
                z = torch.empty_like(u)
                z = z.index_copy(0, idx, u[idx]).index_copy(0, af, (u[af] - t) * torch.exp(-s))
                ldj = ldj - s.sum(dim=0)
            else:
                t = self._net(enets[0], inp)
                z = torch.empty_like(u)
                z = z.index_copy(0, idx, u[idx]).index_copy(0, af, u[af] - t)
            u = z
        return u, ldj

    def forward(self, z: torch.Tensor, theta: torch.Tensor):
        """Sampling direction (reference `forward` / `forward!`): elements first -> last."""
        B = z.shape[1]
        ldj = torch.zeros(B, dtype=self.dtype)
        u = z
        for e, enets in zip(self.elems, self.nets):
            if e.kind == "norm":
                xmin = torch.tensor(e.x_min, dtype=self.dtype)[:, None]
                xmax = torch.tensor(e.x_max, dtype=self.dtype)[:, None]
                xd = xmax - xmin
                u = (xd * u - e.alpha * xmax + e.beta * xmin) / (e.beta - e.alpha)
                ldj = ldj + torch.sum(torch.log(xd / (e.beta - e.alpha))) * torch.ones(B, dtype=self.dtype)
                continue
            nn_idx = torch.as_tensor(e.axes.nn0)
            af = torch.as_tensor(e.axes.af0)
            idx = torch.as_tensor(e.axes.id0)
            inp = torch.cat([theta, u], dim=0)[nn_idx]
            if e.kind == "rnvp":
                s = self._net(enets[0], inp)
                t = self._net(enets[1], inp)
                x = torch.empty_like(u)
                x = x.index_copy(0, idx, u[idx]).index_copy(0, af, u[af] * torch.exp(s) + t)
                ldj = ldj + s.sum(dim=0)
            else:
                t = self._net(enets[0], inp)
                x = torch.empty_like(u)
                x = x.index_copy(0, idx, u[idx]).index_copy(0, af, u[af] + t)
            u = x
        return u, ldj

    def logpdf(self, x, theta):
        z, ldj = self.backward(x, theta)
        d = z.shape[0]
        return -(d * LOG2PI) / 2 - 0.5 * (z * z).sum(dim=0) + ldj

    def loss(self, x, theta, inv_btot: Optional[float] = None):
        lp = self.logpdf(x, theta)
        return -lp.mean() if inv_btot is None else -lp.sum() * inv_btot

    def loss_and_grad(self, x, theta, inv_btot: Optional[float] = None):
        for p in self.params:
            p.requires_grad_(True)
            p.grad = None
        l = self.loss(x, theta, inv_btot)
        l.backward()
        return l.detach(), self.flat_grad()


def adam_step_(w: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, t: int, lr=1e-3, b1=0.9, b2=0.999,
               eps=1e-8) -> None:
    """Optimisers.Adam, in place on flat tensors."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    step = m / (1 - b1 ** t) / (torch.sqrt(v / (1 - b2 ** t)) + eps) * lr
    w.sub_(step)


def train_step(tc: TorchChain, x, theta, m, v, t: int, lr=1e-3):
    """One reference minibatch step: gradient (Flows.jl:400-413) + Adam update (Flows.jl:415)."""
    l, g = tc.loss_and_grad(x, theta)
    w = tc.flat_params()
    adam_step_(w, g, m, v, t, lr=lr)
    tc.set_flat_params(w)
    return float(l)
