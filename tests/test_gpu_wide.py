"""GPU parity of the tensor-core path (tcgen05 / TMEM, 3xTF32 split accumulation; dflow_tc.cu) against the Float64
oracle: both directions, log-density, sampling and the adjoint.

Tolerances: rtol 1e-5 on z / x / ldj / logp (+ the measured Float32-oracle distance from Float64), as for the narrow
path; gradients 1e-4 of the gradient's max-norm (2e-4 per Dense) plus a conditioning slack: a hidden unit whose
pre-activation sits within ~1e-6 of the ReLU kink flips its mask under ANY Float32 evaluation order and moves that
sample's contribution, so the slack is the larger of |grad_f32_oracle - grad_f64_oracle| and the change of the
Float64 gradient under a 1e-6 relative perturbation of the inputs.
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
    # 20 transformed coordinates (a16 = 32) and 28 conditioner inputs (K0p = 32): the column limits of the TMEM-A weight-gradient kernel
    "h128_d40": (40, 8, 2, 128),
    # 56 conditioner inputs and 32 transformed coordinates at hidden 256: 92 operand row-blocks per weight-gradient stage
    "h256_d64_n24": (64, 24, 2, 256),
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


def _grad_slack(ochain, x, th, go):
    _, go32, _, _ = O.chain_loss_and_grad(ochain, x, th, np.float32)
    s = np.abs(go32 - go)
    for eps in (1e-6, -1e-6):
        _, gp, _, _ = O.chain_loss_and_grad(ochain, x * (1.0 + eps), th, np.float64)
        s = np.maximum(s, np.abs(gp - go))
    return s


def _check_grad(name, ochain, pc, x, th, B, per_dense=2e-4):
    grad = torch.zeros(pc.P, device=DEV)
    loss2 = torch.zeros(2, device=DEV)
    pc.loss_grad(x, th if th.shape[0] else None, grad, loss2)
    lo, go, _, _ = O.chain_loss_and_grad(ochain, x, th, np.float64)
    assert pc.P == go.size
    assert loss2[1].item() == 0
    assert abs(-loss2[0].item() / B - lo) <= 1e-5 * abs(lo) + 1e-4
    g = grad.cpu().numpy()
    slack = _grad_slack(ochain, x, th, go)
    gmax = np.abs(go).max()
    assert np.abs(g - go).max() <= 1e-4 * gmax + slack.max(), (name, np.abs(g - go).max(), gmax, slack.max())
    off = 0
    for e in O.flatten(ochain):
        for net in O._trainable_nets(e):
            for dl in net:
                k = dl.W.size + (dl.b.size if dl.b is not None else 0)
                ref = go[off:off + k]
                err = np.abs(g[off:off + k] - ref).max()
                assert err <= per_dense * np.abs(ref).max() + 1e-7 * gmax + slack[off:off + k].max(), \
                    (name, off, err, np.abs(ref).max(), slack[off:off + k].max())
                off += k


@pytest.mark.parametrize("name", ["h128_d8", "h96_d6_n0", "c4_like_h256", "h128_d40", "h256_d64_n24"])
@pytest.mark.parametrize("B", [5, 300, 2049])
def test_wide_loss_grad(name, B):
    """Adjoint on tensor cores: stored activations, transposed-chain input gradients, K = samples weight-gradient GEMMs."""
    d, n, L, h = CASES[name]
    xn = O.synthetic_data(d, n, 1000, seed=99)[0]
    ochain = O.block_chain(d, n, L, h, xn, s_out_scale=0.3)
    x, th = O.synthetic_data(d, n, B, seed=21)
    chain = chain_from_oracle(ochain)
    _check_grad(name, ochain, chain.packed(), x, th, B)


def _mixed_wide_chain(x, hidden=128):
    """RNVP + NICE coupling layers with a NormalizationLayer in the middle and one at the end, hidden 128 (or 64)."""
    rng = np.random.default_rng(7)
    l1 = O.coupling_layer(O.coupling_axes(8, [1, 2, 3, 4], n=2), hidden_dim_s=hidden, hidden_dim_t=hidden, rng=rng,
                          bias_scale=0.1, s_out_scale=0.3)
    l2 = O.coupling_layer(O.coupling_axes(8, [8, 5, 6], n=2), kind="nice", hidden_dim_t=hidden, rng=rng, bias_scale=0.1)
    l3 = O.coupling_layer(O.coupling_axes(8, [7, 2, 4, 6, 1], n=2), hidden_dim_s=hidden, hidden_dim_t=hidden, rng=rng,
                          bias_scale=0.1, s_out_scale=0.3)
    return O.Chain([l1, O.norm_layer_from_data(x, -2.0, 3.0), l2, l3, O.norm_layer_from_data(x)])


@pytest.mark.parametrize("hidden", [128, 64])
def test_wide_nice_layer_and_inner_normalization(hidden):
    """NICE (no s-net) layers and a NormalizationLayer inside the chain on the tensor-core kernels, both directions
    and the adjoint (the cotangent is scaled through the inner NormalizationLayer).  hidden = 64: the narrow
    TMEM-sourced kernel, forced onto the tensor cores at this small batch."""
    xn = O.synthetic_data(8, 2, 1000, seed=99)[0]
    ochain = _mixed_wide_chain(xn, hidden)
    B = 700
    x, th = O.synthetic_data(8, 2, B, seed=13)
    chain = chain_from_oracle(ochain)
    if hidden <= 64:
        chain.packed().tune(tc_mode=1)
    z, ldj = df.backward(chain, x, th)
    zo, lo = O.chain_backward(ochain, x, th, np.float64)
    zo32, lo32 = O.chain_backward(ochain, x, th)
    slack = np.abs(zo32 - zo).max() + np.abs(lo32 - lo).max()
    assert_close(df.to_numpy(z), zo, 1e-5, 1e-5 + slack, "mixed z")
    assert_close(df.to_numpy(ldj), lo, 1e-5, 1e-5 + slack, "mixed ldj")
    x2, ldj2 = df.forward(chain, df.to_numpy(z), th)
    assert_close(df.to_numpy(x2), x, 1e-4, 1e-4, "mixed round trip")
    _check_grad(f"mixed_h{hidden}", ochain, chain.packed(), x, th, B)


def test_wide_adjoint_macro_batches():
    """A workspace cap forces the adjoint to run in several macro-batches; the gradient is the same sum."""
    ochain, chain, x, th = _setup("h128_d8", 3000, seed=5)
    pc = chain.packed()
    full = torch.zeros(pc.P, device=DEV)
    l_full = torch.zeros(2, device=DEV)
    pc.loss_grad(x, th, full, l_full)
    pc.tune(tc_ws_budget_mb=8)  # 8 MiB: a few hundred samples per macro-batch
    part = torch.zeros(pc.P, device=DEV)
    l_part = torch.zeros(2, device=DEV)
    n0 = pc.launch_count()
    pc.loss_grad(x, th, part, l_part)
    assert pc.launch_count() - n0 > 2 * 20  # several sweeps
    pc.tune(tc_ws_budget_mb=0)
    assert torch.allclose(part, full, rtol=1e-4, atol=2e-6 * full.abs().max().item())
    assert abs(l_part[0].item() - l_full[0].item()) <= 1e-5 * abs(l_full[0].item())


def test_wide_grad_idx_and_dp_seed():
    """Gather through an index + data-parallel seed: two shards with inv_btot = 1/B sum to the full gradient."""
    ochain, chain, x, th = _setup("h128_d8", 600, seed=4)
    pc = chain.packed()
    xd, td = df.to_jl(x, DEV), df.to_jl(th, DEV)
    perm = torch.randperm(600, generator=torch.Generator().manual_seed(1)).to(torch.int32).to(DEV)
    full = torch.zeros(pc.P, device=DEV)
    l_full = torch.zeros(2, device=DEV)
    pc.loss_grad(xd, td, full, l_full)
    acc = torch.zeros(pc.P, device=DEV)
    l_acc = torch.zeros(2, device=DEV)
    pc.loss_grad(xd, td, acc, l_acc, 1.0 / 600, 0, perm[:250].contiguous())
    pc.loss_grad(xd, td, acc, l_acc, 1.0 / 600, 0, perm[250:].contiguous())
    assert torch.allclose(acc, full, rtol=1e-4, atol=2e-6 * full.abs().max().item())
    assert abs(l_acc[0].item() - l_full[0].item()) <= 1e-4 * abs(l_full[0].item())


@pytest.mark.parametrize("B", [5, 300, 2049])
def test_wide_adjoint_h512(B):
    """Hidden 512 (C5's conditioner) trains since round 2: the input-gradient chain runs its two 256-column passes on the
    transposed matrices, and the weight-gradient kernel splits the 512 columns of dW2 over two CTAs per 128-row tile (TMEM
    holds 512 columns; the second CTA carries only dW2)."""
    d, n, L, h = CASES["c5_like_h512"]
    xn = O.synthetic_data(d, n, 1000, seed=99)[0]
    ochain = O.block_chain(d, n, L, h, xn, s_out_scale=0.3)
    x, th = O.synthetic_data(d, n, B, seed=21)
    chain = chain_from_oracle(ochain)
    # whole gradient: 1e-4 of its max-norm like every other case; per Dense the 512 x 512 blocks sit at 2.6e-4 of their own
    # max-norm at B = 2049 (3xTF32 products accumulated over 512-wide contractions and 2049 samples), hence 4e-4 here
    _check_grad("c5_like_h512", ochain, chain.packed(), x, th, B, per_dense=4e-4)


NARROW_ON_TC = {"h64_d16": (16, 4, 2, 64), "h32_d10": (10, 3, 4, 32)}


@pytest.mark.parametrize("tc_ts", [0, -1])
@pytest.mark.parametrize("name", list(NARROW_ON_TC))
def test_narrow_chain_on_tensor_cores(name, tc_ts):
    """tc_mode=1 routes an eligible hidden <= 64 chain through the tensor-core kernels: same results as the oracle and
    as the CUDA-core kernels.  tc_ts = 0 (default): the TMEM-sourced four-chain kernel (dflow_tcs.cuh); tc_ts = -1: the
    warp-specialised pipeline that serves the wider nets."""
    d, n, L, h = NARROW_ON_TC[name]
    xn = O.synthetic_data(d, n, 1000, seed=99)[0]
    ochain = O.block_chain(d, n, L, h, xn, s_out_scale=0.3)
    B = 700
    x, th = O.synthetic_data(d, n, B, seed=11)
    chain = chain_from_oracle(ochain)
    pc = chain.packed()
    lp_cuda = pc.logpdf(x, th).clone()
    n0 = pc.launch_count()
    pc.tune(tc_mode=1, tc_ts=tc_ts)
    lp_tc = pc.logpdf(x, th)
    assert pc.launch_count() - n0 > 2 * L  # one launch per conditioner, not the single fused narrow kernel
    zo, lo = O.chain_backward(ochain, x, th, np.float64)
    lpo = O.mvnormal_logpdf(zo, np.float64) + lo
    assert_close(df.to_numpy(lp_tc), lpo, 1e-5, 1e-4, f"{name} logpdf on tensor cores")
    assert_close(df.to_numpy(lp_tc), df.to_numpy(lp_cuda), 1e-5, 1e-4, f"{name} tensor vs CUDA cores")
    _check_grad(name, ochain, pc, x, th, B)
    # sampling direction through the same kernels: the round trip returns x
    z, _ = pc.normalize(df.to_jl(x, DEV), df.to_jl(th, DEV))
    x2, _ = pc.forward_ldj(z, df.to_jl(th, DEV))
    assert_close(df.to_numpy(x2), x, 1e-4, 1e-4, f"{name} round trip on tensor cores")
    pc.tune(tc_mode=0, tc_ts=0)


@pytest.mark.parametrize("name", ["c4_like_h256", "c5_like_h512", "h128_d8", "h64_forced"])
def test_wide_many_tiles_per_cta_equal_small_launches(name):
    """Size-independent property at a batch where every persistent CTA walks several tiles (ring slots and barrier
    phases wrap many times, last tile partial): per-sample results must not depend on how the batch is cut, so one
    big call equals the concatenation of small calls bit for bit; the gradient of the big batch equals the sum of the
    chunks' gradients."""
    if name == "h64_forced":
        d, n, L, h = 16, 4, 2, 64
    else:
        d, n, L, h = CASES[name]
    xn = O.synthetic_data(d, n, 1000, seed=99)[0]
    ochain = O.block_chain(d, n, L, h, xn, s_out_scale=0.3)
    chain = chain_from_oracle(ochain)
    pc = chain.packed()
    if h <= 64:
        pc.tune(tc_mode=1)
    B = 148 * 128 * 5 + 77
    g = torch.Generator(device=DEV).manual_seed(3)
    x = df.jl_empty((d, B), DEV)
    x.normal_(generator=g)
    th = df.jl_empty((n, B), DEV)
    th.uniform_(0, 1, generator=g)
    big = pc.logpdf(x, th).clone()
    step = 9000
    parts = [pc.logpdf(df.to_jl(x[:, i:i + step].clone(), DEV), df.to_jl(th[:, i:i + step].clone(), DEV))
             for i in range(0, B, step)]
    small = torch.cat([p_.reshape(-1) for p_ in parts])
    assert torch.equal(big.reshape(-1), small)
    xs = pc.sample_rng(B, 5, None, torch.full((n,), 0.4, device=DEV))
    xs2 = torch.cat([df.to_jl(pc.sample_rng(min(step, B - i), 5, None, torch.full((n,), 0.4, device=DEV), first_sample=i), DEV)
                     for i in range(0, B, step)], dim=1)
    assert torch.equal(df.to_jl(xs, DEV), xs2)
    if h <= 256:
        gb = torch.zeros(pc.P, device=DEV)
        lb = torch.zeros(2, device=DEV)
        pc.loss_grad(x, th, gb, lb)
        gs = torch.zeros(pc.P, device=DEV)
        ls = torch.zeros(2, device=DEV)
        for i in range(0, B, 3 * step):
            pc.loss_grad(df.to_jl(x[:, i:i + 3 * step].clone(), DEV), df.to_jl(th[:, i:i + 3 * step].clone(), DEV), gs, ls,
                         1.0 / B)
        assert torch.allclose(gb, gs, rtol=1e-4, atol=3e-6 * gb.abs().max().item())
        assert abs(lb[0].item() - ls[0].item()) <= 1e-5 * abs(lb[0].item())


def test_release_build_rejects_experiment_keys():
    """The timing-experiment switches ("tc_debug": results wrong by construction) and the CTA-pair / first-generation
    variants (measured slower, profiles/r01_tc_summary.md) are compiled out of the release library: the public tuning
    entry point must refuse their keys instead of silently accepting them."""
    ochain, chain, x, th = _setup("h128_d8", 64)
    pc = chain.packed()
    for key in ("tc_debug", "tc_cluster", "wide_gen", "tc_ns_max"):
        with pytest.raises(df.DflowInvalidArg):
            pc.tune(**{key: 1})


@pytest.mark.parametrize("name", ["h128_d8", "h96_d6_n0", "h64_d16", "h32_d10"])
def test_fused_conditioner_pair_matches_separate_conditioners(name):
    """tc_fuse: the s and t conditioners of a RealNVP layer (same input, same widths, 2h <= 256) run as one conditioner
    of width 2h with block-diagonal W2 / W3 (default in the train step; tc_fuse=2 also in forward-type calls).  The zero
    blocks contribute exact zeros, so forward results are bit-identical to the separate conditioners; gradients agree to
    summation order."""
    narrow = name in NARROW_ON_TC
    d, n, L, h = NARROW_ON_TC[name] if narrow else CASES[name]
    B = 1300
    x, th = O.synthetic_data(d, n, B, seed=21)
    ochain = O.block_chain(d, n, L, h, O.synthetic_data(d, n, 1000, seed=99)[0], s_out_scale=0.3)
    chain = chain_from_oracle(ochain)
    pc = chain.packed()
    if narrow:
        pc.tune(tc_mode=1, tc_ts=-1)  # hidden <= 64 pairs fuse on the warp-specialised pipeline only
    xj = df.to_jl(x, DEV)
    tj = df.to_jl(th, DEV) if n else None
    out = {}
    for fuse in (2, 0):
        pc.tune(tc_fuse=fuse)
        before = pc.launch_count()
        lp = pc.logpdf(xj, tj).clone()
        nl = pc.launch_count() - before
        z, ldj = pc.normalize(xj, tj)
        grad = torch.zeros(pc.P, device=DEV)
        l2 = torch.zeros(2, device=DEV)
        pc.loss_grad(xj, tj, grad, l2)
        out[fuse] = (lp, z.clone(), ldj.clone(), grad, l2, nl)
    pc.tune(tc_fuse=1)
    if narrow:
        pc.tune(tc_mode=0, tc_ts=0)
    assert out[2][5] < out[0][5], "fused pairs need fewer launches"
    for i in range(3):
        assert torch.equal(out[2][i], out[0][i])
    gmax = float(out[0][3].abs().max())
    assert float((out[2][3] - out[0][3]).abs().max()) <= 2e-5 * gmax
    assert abs(float(out[2][4][0] - out[0][4][0])) <= 1e-5 * abs(float(out[0][4][0]))


def _nondefault_wide_chain(x, bias, hidden=128):
    """hidden 128 (or 64) conditioners with tanh (s) / sigmoid (t) hidden activations, with or without bias; a second block
    with relu s-nets and identity-hidden t-nets.  Reference: CouplingLayer(...; σ_s, σ_t, bias) (src/Layers.jl:113-123)."""
    rng = np.random.default_rng(17)
    b1 = O.coupling_block(O.coupling_axes_cut(8, 4, n=2), hidden_dim_s=hidden, hidden_dim_t=hidden, act_s="tanh", act_t="sigmoid",
                          bias=bias, rng=rng, bias_scale=0.1, s_out_scale=0.3)
    b2 = O.coupling_block(O.coupling_axes_cut(8, 4, n=2), hidden_dim_s=hidden, hidden_dim_t=hidden, act_s="relu", act_t="identity",
                          bias=bias, rng=rng, bias_scale=0.1, s_out_scale=0.3)
    return O.Chain([b1, b2, O.norm_layer_from_data(x)])


@pytest.mark.parametrize("hidden", [128, 64])
@pytest.mark.parametrize("bias", [True, False])
def test_wide_nondefault_activations_and_bias(bias, hidden):
    """Round 2: the tensor-core kernels take any of relu / tanh / sigmoid / identity on the two hidden layers and
    bias = false (the non-relu derivative comes from the stored activations in the adjoint sweep).  hidden = 64 runs the
    narrow TMEM-sourced kernel (tc_mode = 1 forces the tensor cores at this small batch)."""
    xn = O.synthetic_data(8, 2, 1000, seed=99)[0]
    ochain = _nondefault_wide_chain(xn, bias, hidden)
    B = 700
    x, th = O.synthetic_data(8, 2, B, seed=13)
    chain = chain_from_oracle(ochain)
    if hidden <= 64:
        chain.packed().tune(tc_mode=1)
    z, ldj = df.backward(chain, x, th)
    zo, lo = O.chain_backward(ochain, x, th, np.float64)
    zo32, lo32 = O.chain_backward(ochain, x, th)
    slack = np.abs(zo32 - zo).max() + np.abs(lo32 - lo).max()
    assert_close(df.to_numpy(z), zo, 1e-5, 1e-5 + slack, "nondefault z")
    assert_close(df.to_numpy(ldj), lo, 1e-5, 1e-5 + slack, "nondefault ldj")
    x2, _ = df.forward(chain, df.to_numpy(z), th)
    assert_close(df.to_numpy(x2), x, 1e-4, 1e-4, "nondefault round trip")
    _check_grad(f"nondefault_h{hidden}", ochain, chain.packed(), x, th, B)


def _random_narrow_chain(seed):
    """A random hidden 32 / 64 chain the narrow tensor-core kernel is eligible for: random d, n, masks (unsorted, ragged),
    RealNVP / NICE layers, activations, bias on / off, a NormalizationLayer at a random position."""
    rng = np.random.default_rng(seed)
    h = int(rng.choice([32, 64]))
    n = int(rng.integers(0, 5))
    d = int(rng.integers(3, 14 if h == 32 else 20))
    acts = ["relu", "relu", "tanh", "sigmoid", "identity"]
    layers = []
    nl = int(rng.integers(2, 5))
    for li in range(nl):
        na = int(rng.integers(1, min(d - 1, 12) + 1))
        mask = [int(m) + 1 for m in rng.permutation(d)[:na]]
        kind = "nice" if rng.random() < 0.25 else "rnvp"
        layers.append(O.coupling_layer(O.coupling_axes(d, mask, n=n), kind=kind, hidden_dim_s=h, hidden_dim_t=h,
                                       act_s=str(rng.choice(acts)), act_t=str(rng.choice(acts)), bias=bool(rng.random() < 0.7),
                                       rng=rng, bias_scale=0.1, s_out_scale=0.3))
    xn = O.synthetic_data(d, n, 500, seed=seed + 1000)[0]
    layers.insert(int(rng.integers(0, len(layers) + 1)), O.norm_layer_from_data(xn, -1.0, 2.0))
    return O.Chain(layers), d, n, h


@pytest.mark.parametrize("seed", list(range(10)))
def test_narrow_tensor_core_kernel_random_chains(seed):
    """Randomised structure sweep: the TMEM-sourced kernel (tc_mode = 1) against the CUDA-core kernels (tc_mode = -1, exact
    Float32 FMA chains) on the same chain and batch -- log-density, the sampling direction and the gradient; batch sizes
    with ragged tiles, fewer tiles than chains per CTA, and several tiles per chain."""
    ochain, d, n, h = _random_narrow_chain(seed)
    B = [1, 77, 129, 600, 1500, 5000, 700, 333, 2048, 4097][seed]
    x, th = O.synthetic_data(d, n, B, seed=seed + 50)
    chain = chain_from_oracle(ochain)
    pc = chain.packed()
    xj = df.to_jl(x, DEV)
    tj = df.to_jl(th, DEV) if n else None
    out = {}
    launches = {}
    for mode in (-1, 1):
        pc.tune(tc_mode=mode)
        n0 = pc.launch_count()
        lp = pc.logpdf(xj, tj).clone()
        launches[mode] = pc.launch_count() - n0
        z, ldj = pc.normalize(xj, tj)
        xb, ldjb = pc.forward_ldj(z.clone(), tj)
        grad = torch.zeros(pc.P, device=DEV)
        l2 = torch.zeros(2, device=DEV)
        pc.loss_grad(xj, tj, grad, l2)
        out[mode] = (lp, z.clone(), ldj.clone(), xb.clone(), grad, l2.clone())
    pc.tune(tc_mode=0)
    assert launches[1] > launches[-1] + 2, "the chain did not route to the tensor-core kernels"
    ref, got = out[-1], out[1]
    assert_close(df.to_numpy(got[0]), df.to_numpy(ref[0]), 2e-5, 2e-4, f"seed {seed} logpdf")
    assert_close(df.to_numpy(got[1]), df.to_numpy(ref[1]), 2e-5, 2e-5, f"seed {seed} z")
    assert_close(df.to_numpy(got[2]), df.to_numpy(ref[2]), 2e-5, 2e-5, f"seed {seed} ldj")
    assert_close(df.to_numpy(got[3]), x, 2e-4, 2e-4, f"seed {seed} round trip")
    gmax = float(ref[4].abs().max())
    assert float((got[4] - ref[4]).abs().max()) <= 2e-4 * gmax + 1e-7, (seed, float((got[4] - ref[4]).abs().max()), gmax)
    assert abs(float(got[5][0] - ref[5][0])) <= 1e-5 * abs(float(ref[5][0])) + 1e-3


@pytest.mark.parametrize("name", ["h128_d8", "c4_like_h256", "h64_d16"])
def test_weight_gradient_kernels_agree(name):
    """The two weight-gradient kernels -- A operands through TMEM (default where the TMEM column budget allows) and all six
    operand segments staged in shared memory (tc_dw_ts = -1; the only one for hidden 512 with many conditioner inputs) --
    accumulate the same products: gradients agree to summation order, at a batch where every CTA walks several stages
    and the last tile is ragged."""
    narrow = name in NARROW_ON_TC
    d, n, L, h = NARROW_ON_TC[name] if narrow else CASES[name]
    B = 148 * 128 + 4 * 128 + 77
    ochain = O.block_chain(d, n, L, h, O.synthetic_data(d, n, 1000, seed=99)[0], s_out_scale=0.3)
    chain = chain_from_oracle(ochain)
    pc = chain.packed()
    if narrow:
        pc.tune(tc_mode=1)
    g = torch.Generator(device=DEV).manual_seed(9)
    x = df.jl_empty((d, B), DEV)
    x.normal_(generator=g)
    th = df.jl_empty((n, B), DEV)
    th.uniform_(0, 1, generator=g)
    grads = {}
    for ts in (0, -1):
        pc.tune(tc_dw_ts=ts)
        grad = torch.zeros(pc.P, device=DEV)
        l2 = torch.zeros(2, device=DEV)
        pc.loss_grad(x, th, grad, l2)
        grads[ts] = (grad, l2)
    pc.tune(tc_dw_ts=0)
    if narrow:
        pc.tune(tc_mode=0)
    gmax = float(grads[-1][0].abs().max())
    assert float((grads[0][0] - grads[-1][0]).abs().max()) <= 2e-5 * gmax
    # the loss does not depend on the weight-gradient kernel (its block sums are added atomically: summation order only)
    assert abs(float(grads[0][1][0]) - float(grads[-1][1][0])) <= 2e-6 * abs(float(grads[-1][1][0]))
