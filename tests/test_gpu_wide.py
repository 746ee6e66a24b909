"""GPU parity of the wide-conditioner path (tcgen05 / TMEM, 3xTF32 split accumulation) against the Float64 oracle.

Tolerance: rtol 1e-5 on z / x / ldj / logp (+ the measured Float32-oracle distance from Float64), as for the narrow path.
"""
import numpy as np
import pytest
import torch

import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from oracle import philox as PH
from tests.helpers import assert_close, chain_from_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

CASES = {
    # name: (d, n, layers, hidden)
    "h128_d8": (8, 2, 2, 128),
    "h96_d6_n0": (6, 0, 2, 96),
    "c4_like_h256": (32, 8, 2, 256),
    "c5_like_h512": (64, 16, 2, 512),
}


def _setup(name, B, seed=11):
    d, n, L, h = CASES[name]
    xn = O.synthetic_data(d, n, 1000, seed=99)[0]
    ochain = O.block_chain(d, n, L, h, xn, s_out_scale=0.3)
    x, th = O.synthetic_data(d, n, B, seed=seed)
    return ochain, chain_from_oracle(ochain), x, th


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("B", [1, 130, 700])
def test_wide_normalize_logpdf(name, B):
    ochain, chain, x, th = _setup(name, B)
    tharg = th if th.shape[0] else None
    z, ldj = df.backward(chain, x, tharg)
    zo32, lo32 = O.chain_backward(ochain, x, th)
    zo, lo = O.chain_backward(ochain, x, th, np.float64)
    slack = np.abs(zo32 - zo).max() + np.abs(lo32 - lo).max()
    assert_close(df.to_numpy(z), zo, 1e-5, 1e-5 + slack, f"{name} z")
    assert_close(df.to_numpy(ldj), lo, 1e-5, 1e-5 + slack, f"{name} ldj")
    lp = chain.packed().logpdf(x, tharg)
    assert_close(df.to_numpy(lp), O.mvnormal_logpdf(zo, np.float64) + lo, 1e-5, 1e-4 + 10 * slack, f"{name} logpdf")


@pytest.mark.parametrize("name", list(CASES))
def test_wide_sampling_round_trip(name):
    ochain, chain, x, th = _setup(name, 300)
    d, n, _, _ = CASES[name]
    tharg = th if n else None
    z = (np.random.default_rng(2).standard_normal((d, 300)) * 0.7).astype(np.float32)
    xg, ldj = df.forward(chain, z, tharg)
    xo, lo = O.chain_forward(ochain, z, th, np.float64)
    scale = max(1.0, np.abs(xo).max())
    assert_close(df.to_numpy(xg), xo, 1e-5, 1e-5 * scale, f"{name} forward x")
    assert_close(df.to_numpy(ldj), lo, 1e-5, 1e-4, f"{name} forward ldj")
    zt = df.to_jl(z, DEV).clone()
    df.forward_(chain, zt, tharg)
    assert torch.equal(zt, xg)
    z2, ldj2 = df.backward(chain, xg, tharg)
    assert_close(df.to_numpy(z2), z, 1e-4, 1e-4, f"{name} round trip")
    assert_close(df.to_numpy(ldj2) + df.to_numpy(ldj), 0 * lo, 0, 1e-4, f"{name} ldj antisymmetry")


def test_wide_sample_rng_and_index_gather():
    ochain, chain, x, th = _setup("h128_d8", 400)
    pc = chain.packed()
    thc = torch.tensor([0.3, -0.2], device=DEV)
    xs = pc.sample_rng(500, 99, None, thc, first_sample=5)
    z = PH.normal_samples(8, 500, 99, 0, first_sample=5)
    thb = np.tile(np.array([[0.3], [-0.2]], np.float32), (1, 500))
    xo, _ = O.chain_forward(ochain, z, thb, np.float64)
    assert_close(df.to_numpy(xs), xo, 1e-5, 2e-5 * max(1.0, np.abs(xo).max()), "wide sample_rng")
    perm = torch.randperm(400, generator=torch.Generator().manual_seed(1)).to(torch.int32).to(DEV)
    xd, td = df.to_jl(x, DEV), df.to_jl(th, DEV)
    lp = pc.logpdf(xd, td)
    lpi = pc.logpdf(xd, td, 0, perm)
    assert torch.equal(lp[perm.long()], lpi)


def test_wide_adjoint_is_rejected_not_emulated():
    ochain, chain, x, th = _setup("h128_d8", 64)
    pc = chain.packed()
    g = torch.zeros(pc.P, device=DEV)
    l2 = torch.zeros(2, device=DEV)
    with pytest.raises(df.DflowUnsupported):
        pc.loss_grad(x, th, g, l2)
