"""The oracle pinned against everything the reference's own tests hold for this path (test/runtests.jl:7-121):
five property testsets + the fixture, plus the committed golden vectors and independent gradient checks.
CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import dflow_oracle as O
from oracle import torch_ref as T
from tests.golden.cases import readme_n1_chain, ref_chain_d7

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SQRT_EPS32 = float(np.sqrt(np.finfo(np.float32).eps))  # Julia's default ≈ rtol for Float32


def fixture():
    fx = np.load(os.path.join(GOLD, "fixture_datatest.npz"))
    return fx["x"], fx["theta"]


# ---- testset "data" (test/runtests.jl:7-31) -------------------------------------------------------------------
def test_ref_data_testset():
    x = 0.2 * np.ones((7, 10), np.float32)
    th = 0.1 * np.ones((2, 10), np.float32)
    x[0, 1] = 0.3
    th[0, 1] = 0.4
    tmin, tmax = th.min(axis=1), th.max(axis=1)
    assert O.dflt_theta(x).shape[1:] == x.shape[1:] and O.dflt_theta(x).shape[0] == 0
    tr, va, te = O.data_partition(10, 0.9, 0.1, np.random.default_rng(0))
    assert len(tr) == 9 and len(va) == 1 and len(te) == 0
    th_t = O.normalize_input(th[:, tr], tmin, tmax)
    assert th_t.max() <= 1 and th_t.min() >= 0
    # the second θ row has zero range -> exactly 0 (src/Data.jl:216)
    assert np.all(th_t[1] == 0)
    assert x[:, tr].max() <= 1 and x[:, tr].min() >= 0


# ---- testset "axes" (test/runtests.jl:33-41) ------------------------------------------------------------------
def test_ref_axes_testset():
    assert O.axes_equal(O.coupling_axes(7, [4, 5, 6, 7], n=2), O.coupling_axes_cut(7, 3, n=2))
    assert O.axes_equal(O.coupling_axes_cut(7, None, n=2), O.coupling_axes_cut(7, 3, n=2))  # CouplingAxes(data)
    a = O.coupling_axes(5, [5, 1, 2], n=2)  # SURVEY §8 a1 worked example
    assert (a.axis_af, a.axis_id, a.axis_nn) == ([5, 1, 2], [3, 4], [1, 2, 5, 6])
    r = O.reverse_axes(O.coupling_axes(7, [4, 2, 5, 1], n=2))
    assert (r.axis_id, r.axis_af, r.axis_nn) == ([4, 2, 5, 1], [3, 6, 7], [1, 2, 6, 4, 7, 3])
    assert O.is_reverse(O.coupling_axes(7, [4, 2, 5, 1], n=2), r)
    with pytest.raises(AssertionError):
        O.coupling_axes(3, [4], n=0)  # src/Axes.jl:85


# ---- testset "real_NVP" (test/runtests.jl:43-64) --------------------------------------------------------------
@pytest.mark.parametrize("axes", [O.coupling_axes_cut(7, 3, n=2), O.coupling_axes(7, [1, 3, 5, 7], n=2)])
def test_ref_real_nvp_testset(axes):
    z1 = 0.2 * np.ones((7, 10), np.float32)
    th = 0.1 * np.ones((2, 10), np.float32)
    layer = O.coupling_layer(axes, rng=np.random.default_rng(1))
    x, l1 = O.rnvp_forward(layer, z1, th)
    z2, l2 = O.rnvp_backward(layer, x, th)
    np.testing.assert_allclose(z2, z1, rtol=SQRT_EPS32)
    np.testing.assert_allclose(l1 + l2, 0, atol=SQRT_EPS32)
    assert x.dtype == np.float32 and l1.shape == (10,)


# ---- testset "chain" (test/runtests.jl:66-95) -----------------------------------------------------------------
def test_ref_chain_testset():
    r = np.random.default_rng(0)
    l1 = O.coupling_layer(O.coupling_axes(7, [1, 3, 5, 7], n=2), rng=r)
    l2 = O.coupling_layer(O.coupling_axes(7, [4, 2, 5, 1, 6], n=2), rng=r)
    blk = O.coupling_block(O.coupling_axes(7, [4, 2, 5, 1], n=2), rng=r)
    small = O.Chain([l1, l2])
    assert len(O.concatenate(small, blk).layers) == 3 and len(O.concatenate(blk, small).layers) == 3
    x1 = 0.2 * np.ones((7, 10), np.float32)
    th = 0.1 * np.ones((2, 10), np.float32)
    x1[:, 1] = 0.4
    th[0, 1] = 0.4
    chain = O.concatenate((small, O.Chain([blk, O.norm_layer_from_data(x1)])))
    assert chain.layers[-1].kind == "norm" and len(chain.layers) == 4
    np.testing.assert_array_equal(chain.layers[-1].x_min, np.full(7, 0.2, np.float32))
    np.testing.assert_array_equal(chain.layers[-1].x_max, np.full(7, 0.4, np.float32))
    z, lb = O.chain_backward(chain, x1, th)
    x2, lf = O.chain_forward(chain, z, th)
    np.testing.assert_allclose(x2, x1, rtol=SQRT_EPS32)
    assert np.all(np.abs(lf + lb) <= 2e-6)  # atol = 2f-6, test/runtests.jl:93
    # order: the normalisation is applied FIRST in the normalising direction, and the block runs layer_2 then layer_1
    z_manual, _ = O.norm_backward(chain.layers[3], x1)
    z_manual, _ = O.rnvp_backward(blk.layer_2, z_manual, th)
    z_manual, _ = O.rnvp_backward(blk.layer_1, z_manual, th)
    z_manual, _ = O.rnvp_backward(l2, z_manual, th)
    z_manual, _ = O.rnvp_backward(l1, z_manual, th)
    np.testing.assert_array_equal(z_manual, z)


# ---- testset "flow" (test/runtests.jl:97-121) -----------------------------------------------------------------
def test_ref_flow_testset_on_fixture():
    x, th = fixture()
    assert x.shape == (5, 1000) and th.shape == (1, 1000) and x.dtype == np.float32
    chain = readme_n1_chain(x)
    assert O.pack_params(chain).size == 2322  # fixture chain, n = 1
    tr, va, _ = O.data_partition(1000, 0.9, 0.1, np.random.default_rng(0))
    assert len(tr) == 900 and len(va) == 100
    thn = O.normalize_input(th, th.min(axis=1), th.max(axis=1))
    tl, vl = O.train(chain, x[:, tr], thn[:, tr], x[:, va], thn[:, va], epochs=2, batchsize=64,
                     rng=np.random.default_rng(1))
    assert len(tl) == 2 and np.isfinite(tl).all() and np.isfinite(vl).all() and tl[1] < tl[0]
    # sample(flow, (2,5,7), (-1f0,)) has size (5,2,5,7): N(0,I) draw pushed through forward! with broadcast θ
    z = np.random.default_rng(2).standard_normal((5, 2, 5, 7)).astype(np.float32)
    thb = O.normalize_input(np.full((1, 2, 5, 7), -1.0, np.float32), th.min(axis=1), th.max(axis=1))
    xs, _ = O.chain_forward(chain, z, thb)
    assert xs.shape == (5, 2, 5, 7) and np.isfinite(xs).all()


# ---- golden vectors ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["golden_readme_n1", "golden_ref_chain_d7"])
def test_oracle_reproduces_golden(case):
    g = np.load(os.path.join(GOLD, case + ".npz"))
    if case == "golden_readme_n1":
        chain = readme_n1_chain(fixture()[0])
    else:
        chain = ref_chain_d7(O.synthetic_data(7, 2, 80, seed=5)[0])
    np.testing.assert_array_equal(O.pack_params(chain), g["W"])
    z, ldj = O.chain_backward(chain, g["x"], g["theta"], np.float64)
    np.testing.assert_allclose(z, g["z"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(ldj, g["ldj"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(O.mvnormal_logpdf(z, np.float64) + ldj, g["logp"], rtol=1e-12, atol=1e-12)
    xf, lf = O.chain_forward(chain, g["z_in"], g["theta"], np.float64)
    np.testing.assert_allclose(xf, g["x_fwd"], rtol=1e-12, atol=1e-12)
    loss, grad, _, _ = O.chain_loss_and_grad(chain, g["x"], g["theta"], np.float64)
    np.testing.assert_allclose(loss, g["loss"], rtol=1e-12)
    np.testing.assert_allclose(grad, g["grad"], rtol=1e-10, atol=1e-14)
    # the Float32 restatement stays within Float32 rounding of the Float64 one
    z32, ldj32 = O.chain_backward(chain, g["x"], g["theta"], np.float32)
    np.testing.assert_allclose(z32, g["z"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ldj32, g["ldj"], rtol=1e-5, atol=1e-5)


# ---- adjoint: closed form (rrule of src/affine/RNVP.jl:99-147) vs autograd vs finite differences ----------------
def _mixed_chain(x):
    r = np.random.default_rng(7)
    l1 = O.coupling_layer(O.coupling_axes(6, [2, 5], n=1), hidden_dim_s=10, hidden_dim_t=12, n_sublayers_s=1,
                          n_sublayers_t=3, act_s="tanh", act_t="sigmoid", rng=r, bias_scale=0.2)
    l2 = O.coupling_layer(O.coupling_axes(6, [1, 3, 4, 6], n=1), kind="nice", hidden_dim_t=8, rng=r, bias_scale=0.2)
    blk = O.coupling_block(O.coupling_axes(6, [6, 1, 2], n=1), hidden_dim_s=8, hidden_dim_t=8, bias=False, rng=r)
    return O.Chain([l1, O.norm_layer_from_data(x, -2.0, 3.0), l2, blk])


def test_hand_adjoint_equals_autograd_and_finite_differences():
    x, th = O.synthetic_data(6, 1, 40, seed=8)
    chain = _mixed_chain(x)
    loss, g, _, _ = O.chain_loss_and_grad(chain, x, th, np.float64)
    tc = T.TorchChain(chain, torch.float64)
    l, gt = tc.loss_and_grad(torch.tensor(x, dtype=torch.float64), torch.tensor(th, dtype=torch.float64))
    assert abs(float(l) - loss) < 1e-12
    np.testing.assert_allclose(g, gt.numpy(), rtol=1e-9, atol=1e-12)
    # central finite differences on a few random parameters
    w0 = O.pack_params(chain).astype(np.float64)
    rng = np.random.default_rng(0)
    for i in rng.choice(w0.size, 12, replace=False):
        eps = 1e-4
        vals = []
        for sgn in (+1, -1):
            w = w0.copy()
            w[i] += sgn * eps
            tc.set_flat_params(torch.tensor(w))
            vals.append(float(tc.loss(torch.tensor(x, dtype=torch.float64), torch.tensor(th, dtype=torch.float64))))
        fd = (vals[0] - vals[1]) / (2 * eps)
        assert abs(fd - g[i]) <= 1e-5 * max(1.0, abs(g[i])), (i, fd, g[i])


def test_dp_seed_makes_the_allreduce_a_pure_sum():
    """Σ over shards of grad(shard; inv_btot = 1/B) == grad(full batch)."""
    x, th = O.synthetic_data(5, 2, 96, seed=3)
    chain = O.readme_chain(2, x)
    _, g_full, _, _ = O.chain_loss_and_grad(chain, x, th, np.float64)
    acc = 0
    for lo, hi in ((0, 10), (10, 64), (64, 96)):
        _, g, _, _ = O.chain_loss_and_grad(chain, x[:, lo:hi], th[:, lo:hi], np.float64, inv_btot=1.0 / 96)
        acc = acc + g
    np.testing.assert_allclose(acc, g_full, rtol=1e-10, atol=1e-14)


def test_adam_matches_closed_form_first_step():
    g = np.array([0.5, -2.0, 1e-3], np.float32)
    w, m, v = O.adam_step(np.zeros(3, np.float32), g, np.zeros(3, np.float32), np.zeros(3, np.float32), 1, lr=1e-3)
    # first step of Adam moves every weight by ≈ lr * sign(g)
    np.testing.assert_allclose(w, -1e-3 * np.sign(g), rtol=1e-4)
    np.testing.assert_allclose(m, 0.1 * g, rtol=1e-6)
    np.testing.assert_allclose(v, 0.001 * g * g, rtol=1e-4)


def test_pack_unpack_roundtrip_and_layout():
    x, _ = O.synthetic_data(5, 2, 10)
    chain = O.readme_chain(2, x)
    w = O.pack_params(chain)
    assert w.size == 2418
    # first Dense of the first layer's s_net: vec(weight) column-major then bias
    d0 = chain.layers[0].s_net[0]
    np.testing.assert_array_equal(w[: d0.W.size], d0.W.reshape(-1, order="F"))
    np.testing.assert_array_equal(w[d0.W.size: d0.W.size + d0.b.size], d0.b)
    w2 = w + 1
    O.unpack_params(chain, w2)
    np.testing.assert_array_equal(O.pack_params(chain), w2)


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors from the Random123 distribution (kat_vectors)."""
    from oracle import philox as PH

    r = PH.philox4x32_10(np.uint32(0), np.uint32(0), np.uint32(0), np.uint32(0), 0, 0)
    assert [int(v) for v in r] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    r = PH.philox4x32_10(np.uint32(f), np.uint32(f), np.uint32(f), np.uint32(f), f, f)
    assert [int(v) for v in r] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    r = PH.philox4x32_10(np.uint32(0x243F6A88), np.uint32(0x85A308D3), np.uint32(0x13198A2E), np.uint32(0x03707344),
                         0xA4093822, 0x299F31D0)
    assert [int(v) for v in r] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
