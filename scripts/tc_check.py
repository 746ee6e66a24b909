"""Debug harness for the generation-2 tensor-core path: prints error magnitudes against the Float64 oracle
(forward, sampling, adjoint) instead of asserting, so that one GPU call gives the whole picture."""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from tests.helpers import chain_from_oracle

DEV = "cuda:0"
CASES = {
    # name: (d, n, layers, hidden, force tensor cores)
    "h64_d16": (16, 4, 2, 64, True),
    "h128_d8": (8, 2, 2, 128, False),
    "h96_d6_n0": (6, 0, 2, 96, False),
    "c4_like_h256": (32, 8, 2, 256, False),
    "c5_like_h512": (64, 16, 2, 512, False),
    "h32_d10": (10, 3, 4, 32, True),
}


def setup(name, B, seed=11):
    d, n, L, h, force = CASES[name]
    xn = O.synthetic_data(d, n, 1000, seed=99)[0]
    ochain = O.block_chain(d, n, L, h, xn, s_out_scale=0.3)
    x, th = O.synthetic_data(d, n, B, seed=seed)
    chain = chain_from_oracle(ochain)
    pc = chain.packed(DEV)
    if force:
        pc.tune(tc_mode=1)
    return ochain, chain, pc, x, th


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max()))


def check_fwd(name, B):
    ochain, chain, pc, x, th = setup(name, B)
    tharg = th if th.shape[0] else None
    zo, lo = O.chain_backward(ochain, x, th, np.float64)
    z, ldj = df.backward(chain, x, tharg)
    torch.cuda.synchronize()
    print(f"[fwd] {name} B={B}: z rel {rel(df.to_numpy(z), zo):.2e}  ldj abs {np.abs(df.to_numpy(ldj) - lo).max():.2e}", flush=True)
    zz = (np.random.default_rng(2).standard_normal(x.shape) * 0.7).astype(np.float32)
    xg, l2 = df.forward(chain, zz, tharg)
    xo, lo2 = O.chain_forward(ochain, zz, th, np.float64)
    torch.cuda.synchronize()
    print(f"[smp] {name} B={B}: x rel {rel(df.to_numpy(xg), xo):.2e}  ldj abs {np.abs(df.to_numpy(l2) - lo2).max():.2e}", flush=True)


def check_grad(name, B):
    ochain, chain, pc, x, th = setup(name, B, seed=21)
    tharg = th if th.shape[0] else None
    grad = torch.zeros(pc.P, device=DEV)
    loss2 = torch.zeros(2, device=DEV)
    pc.loss_grad(x, tharg, grad, loss2)
    torch.cuda.synchronize()
    lo, go, _, _ = O.chain_loss_and_grad(ochain, x, th, np.float64)
    g = grad.cpu().numpy()
    _, go32, _, _ = O.chain_loss_and_grad(ochain, x, th, np.float32)
    slack = np.abs(go32 - go).max()
    print(f"[grad] {name} B={B}: loss {-loss2[0].item() / B:.6f} vs {lo:.6f}  grad rel(max-norm) {rel(g, go):.2e}"
          f"  f32-oracle slack rel {slack / np.abs(go).max():.2e}", flush=True)
    off = 0
    for ei, e in enumerate(O.flatten(ochain)):
        for ni, net in enumerate(O._trainable_nets(e)):
            for di, dl in enumerate(net):
                for nm, k in (("W", dl.W.size), ("b", dl.b.size if dl.b is not None else 0)):
                    if k == 0:
                        continue
                    r = go[off:off + k]
                    err = np.abs(g[off:off + k] - r).max() / max(1e-30, np.abs(r).max())
                    if err > 1e-4:
                        print(f"    elem {ei} net {ni} dense {di} {nm}: rel err {err:.2e} (|ref| {np.abs(r).max():.2e})", flush=True)
                    off += k


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    names = sys.argv[2].split(",") if len(sys.argv) > 2 else list(CASES)
    for name in names:
        for B in (1, 130, 700):
            if what in ("all", "fwd"):
                try:
                    check_fwd(name, B)
                except Exception:
                    traceback.print_exc()
        if what in ("all", "grad") and CASES[name][3] <= 256:
            for B in (5, 300, 2049):
                try:
                    check_grad(name, B)
                except Exception:
                    traceback.print_exc()


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("done in %.1fs" % (time.time() - t0))
