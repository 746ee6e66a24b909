"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical weights and inputs.

Tolerances (BASELINE.json north_star): transforms / log-dets / log-densities rtol 1e-5 (+ atol 1e-5 on O(1)
quantities, the reference's own `≈`/atol=2f-6 style), gradients rtol 1e-4 of the gradient's max-norm.
"""
import numpy as np
import pytest
import torch

import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from oracle import philox as PH
from tests.helpers import assert_close, chain_from_oracle

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-5
DEV = "cuda:0"


def _chains():
    """name -> (oracle chain builder, d, n)"""
    out = {}

    def readme(n):
        def mk(x):
            return O.readme_chain(n, x)
        return mk

    out["readme_n2"] = (readme(2), 5, 2)
    out["readme_n1"] = (readme(1), 5, 1)

    def ref_chain_test(x):  # test/runtests.jl:66-95
        rng = np.random.default_rng(3)
        l1 = O.coupling_layer(O.coupling_axes(7, [1, 3, 5, 7], n=2), rng=rng, bias_scale=0.1)
        l2 = O.coupling_layer(O.coupling_axes(7, [4, 2, 5, 1, 6], n=2), rng=rng, bias_scale=0.1)
        blk = O.coupling_block(O.coupling_axes(7, [4, 2, 5, 1], n=2), rng=rng, bias_scale=0.1)
        return O.concatenate((O.Chain([l1, l2]), O.Chain([blk, O.norm_layer_from_data(x)])))

    out["ref_chain_d7"] = (ref_chain_test, 7, 2)

    def hetero(x):  # different widths / depths / activations per net, NICE layer, no bias, norm in the middle
        rng = np.random.default_rng(5)
        l1 = O.coupling_layer(O.coupling_axes(6, [2, 5], n=0), hidden_dim_s=10, hidden_dim_t=24, n_sublayers_s=1,
                              n_sublayers_t=3, act_s="tanh", act_t="sigmoid", rng=rng, bias_scale=0.2)
        l2 = O.coupling_layer(O.coupling_axes(6, [1, 3, 4, 6], n=0), kind="nice", hidden_dim_t=12, rng=rng, bias_scale=0.2)
        l3 = O.coupling_layer(O.coupling_axes(6, [6, 1, 2, 3, 4], n=0), hidden_dim_s=32, hidden_dim_t=8, bias=False, rng=rng)
        return O.Chain([l1, O.norm_layer_from_data(x, -2.0, 3.0), l2, l3])

    out["hetero_d6_n0"] = (hetero, 6, 0)

    def c3small(x):
        return O.block_chain(16, 4, 8, 64, x)

    out["c3_d16_h64"] = (c3small, 16, 4)

    def h32(x):
        return O.block_chain(10, 3, 4, 32, x, s_out_scale=0.5)

    out["blocks_d10_h32"] = (h32, 10, 3)
    return out


CHAINS = _chains()


def _setup(name, B, seed=11):
    mk, d, n = CHAINS[name]
    x, th = O.synthetic_data(d, n, B, seed=seed)
    ochain = mk(O.synthetic_data(d, n, 1000, seed=99)[0])  # NormalizationLayer constants from a separate draw
    chain = chain_from_oracle(ochain)
    return ochain, chain, x, th


@pytest.mark.parametrize("name", list(CHAINS))
@pytest.mark.parametrize("B", [1, 33, 1000])
def test_normalize_and_logpdf(name, B):
    ochain, chain, x, th = _setup(name, B)
    z, ldj = df.backward(chain, x, th if th.shape[0] else None)
    zo, lo = O.chain_backward(ochain, x, th)
    zo64, lo64 = O.chain_backward(ochain, x, th, np.float64)
    # the f32 oracle itself sits this far from the f64 truth; the kernel must not be further than that + tolerance
    slack = np.abs(zo - zo64).max() + np.abs(lo - lo64).max()
    assert_close(df.to_numpy(z), zo64, RTOL, ATOL + slack, f"{name} z")
    assert_close(df.to_numpy(ldj), lo64, RTOL, ATOL + slack, f"{name} ldj")
    lp = chain.packed().logpdf(x, th if th.shape[0] else None)
    lpo = O.mvnormal_logpdf(zo64, np.float64) + lo64
    assert_close(df.to_numpy(lp), lpo, RTOL, 1e-4 + 10 * slack, f"{name} logpdf")


@pytest.mark.parametrize("name", list(CHAINS))
def test_sampling_direction(name):
    ochain, chain, x, th = _setup(name, 777)
    mk, d, n = CHAINS[name]
    z = np.random.default_rng(2).standard_normal((d, 777)).astype(np.float32) * 0.7
    tharg = th if n else None
    xg, ldj = df.forward(chain, z, tharg)
    xo, lo = O.chain_forward(ochain, z, th, np.float64)
    scale = max(1.0, np.abs(xo).max())
    assert_close(df.to_numpy(xg), xo, RTOL, ATOL * scale, f"{name} forward x")
    assert_close(df.to_numpy(ldj), lo, RTOL, ATOL * 10, f"{name} forward ldj")
    # forward! (in place, no ldj) gives the same x
    zt = df.to_jl(z, DEV).clone()
    zt = df.to_jl(zt)
    df.forward_(chain, zt, tharg)
    assert torch.equal(zt, xg)
    # round trip + ldj antisymmetry (test/runtests.jl:89-93)
    z2, ldj2 = df.backward(chain, xg, tharg)
    assert_close(df.to_numpy(z2), z, 1e-4, 1e-4, f"{name} round trip")
    assert_close(df.to_numpy(ldj2) + df.to_numpy(ldj), 0 * lo, 0, 1e-4, f"{name} ldj antisymmetry")


@pytest.mark.parametrize("name", list(CHAINS))
@pytest.mark.parametrize("B", [5, 300, 2049])
def test_loss_grad(name, B):
    ochain, chain, x, th = _setup(name, B, seed=21)
    pc = chain.packed()
    grad = torch.zeros(max(pc.P, 1), device=DEV)
    loss2 = torch.zeros(2, device=DEV)
    pc.loss_grad(x, th if th.shape[0] else None, grad, loss2)
    lo, go, _, _ = O.chain_loss_and_grad(ochain, x, th, np.float64)
    assert pc.P == go.size
    assert loss2[1].item() == 0
    assert abs(-loss2[0].item() / B - lo) <= 1e-5 * abs(lo) + 1e-4
    g = grad.cpu().numpy()[: pc.P]
    # distance of the Float32 oracle (what the reference's Flux path computes in) from the Float64 truth
    _, go32, _, _ = O.chain_loss_and_grad(ochain, x, th, np.float32)
    slack = np.abs(go32 - go).max()
    assert np.abs(g - go).max() <= 1e-4 * np.abs(go).max() + slack, (np.abs(g - go).max(), np.abs(go).max(), slack)
    # per-Dense check so that a small block cannot hide behind a large one
    off = 0
    for e in O.flatten(ochain):
        for net in O._trainable_nets(e):
            for dl in net:
                k = dl.W.size + (dl.b.size if dl.b is not None else 0)
                ref = go[off:off + k]
                err = np.abs(g[off:off + k] - ref).max()
                assert err <= 2e-4 * np.abs(ref).max() + 1e-7 * np.abs(go).max() + slack, (name, off, err, np.abs(ref).max())
                off += k


def test_grad_idx_and_dp_seed():
    """Gather through an index + data-parallel seed: two half batches with inv_btot = 1/B sum to the full gradient."""
    ochain, chain, x, th = _setup("readme_n2", 512, seed=4)
    pc = chain.packed()
    xd, td = df.to_jl(x, DEV), df.to_jl(th, DEV)
    perm = torch.randperm(512, generator=torch.Generator().manual_seed(1)).to(torch.int32).to(DEV)
    full = torch.zeros(pc.P, device=DEV)
    l_full = torch.zeros(2, device=DEV)
    pc.loss_grad(xd, td, full, l_full)
    acc = torch.zeros(pc.P, device=DEV)
    l_acc = torch.zeros(2, device=DEV)
    pc.loss_grad(xd, td, acc, l_acc, 1.0 / 512, 0, perm[:200].contiguous())
    pc.loss_grad(xd, td, acc, l_acc, 1.0 / 512, 0, perm[200:].contiguous())
    assert torch.allclose(acc, full, rtol=1e-4, atol=1e-6 * full.abs().max().item())
    assert abs(l_acc[0].item() - l_full[0].item()) <= 1e-4 * abs(l_full[0].item())
    # logpdf through the same index
    lp = pc.logpdf(xd, td)
    lpi = pc.logpdf(xd, td, 0, perm)
    assert torch.equal(lp[perm.long()], lpi)


def test_theta_normalize_flag():
    """Flow-level calls normalise θ with (θ - θ_min)/(θ_max - θ_min), zero range -> 0 (src/Data.jl:213-218)."""
    x, th = O.synthetic_data(5, 2, 300, seed=9)
    th[1] = 0.75  # zero-range row
    ochain = O.readme_chain(2, x)
    chain = chain_from_oracle(ochain)
    data = df.DataArrays(x, th, device=DEV)
    flow = df.Flow(chain, data)
    lp = df.logpdf(flow, x, th)
    tmin, tmax = th.min(axis=1), th.max(axis=1)
    np.testing.assert_array_equal(flow.metadata.θ_min, tmin)
    np.testing.assert_array_equal(flow.metadata.θ_max, tmax)
    lpo = O.logpdf(ochain, x, th, tmin, tmax, np.float64)
    assert_close(df.to_numpy(lp), lpo, RTOL, 1e-4, "flow logpdf")
    # tuple θ == broadcast array θ
    lp_t = df.logpdf(flow, x, (0.3, 0.75))
    th_b = np.tile(np.array([[0.3], [0.75]], np.float32), (1, 300))
    assert torch.equal(lp_t, df.logpdf(flow, x, th_b))


def test_adam_matches_optimisers_formula():
    rng = np.random.default_rng(0)
    P = 5000
    w = rng.standard_normal(P).astype(np.float32)
    m = np.zeros(P, np.float32)
    v = np.zeros(P, np.float32)
    wt, mt, vt = (torch.tensor(a, device=DEV) for a in (w, m, v))
    import ctypes

    lib = df._lib.lib()
    for t in range(1, 6):
        g = rng.standard_normal(P).astype(np.float32) * 0.1
        gt = torch.tensor(g, device=DEV)
        df._lib.check(lib.dflow_adam_step(wt.data_ptr(), gt.data_ptr(), mt.data_ptr(), vt.data_ptr(), P, 1e-3, 0.9,
                                          0.999, 1e-8, t, torch.cuda.current_stream().cuda_stream))
        w, m, v = O.adam_step(w, g, m, v, t)
        np.testing.assert_allclose(wt.cpu().numpy(), w, rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(mt.cpu().numpy(), m, rtol=2e-6, atol=1e-9)
        np.testing.assert_allclose(vt.cpu().numpy(), v, rtol=2e-6, atol=1e-12)


def test_minmax_rows():
    rng = np.random.default_rng(1)
    for rows, B in [(1, 1), (5, 1000), (7, 123457), (64, 4099)]:
        a = (rng.standard_normal((rows, B)) * 3).astype(np.float32)
        a[0, 0] = -0.0
        mn, mx = df.minmax_rows(df.to_jl(a, DEV))
        np.testing.assert_array_equal(mn, a.min(axis=1))
        np.testing.assert_array_equal(mx, a.max(axis=1))
    a = np.abs(rng.standard_normal((3, 100))).astype(np.float32) + 1  # all positive
    mn, mx = df.minmax_rows(df.to_jl(a, DEV))
    np.testing.assert_array_equal(mn, a.min(axis=1))
    a = -a  # all negative
    mn, mx = df.minmax_rows(df.to_jl(a, DEV))
    np.testing.assert_array_equal(mx, a.max(axis=1))
    np.testing.assert_array_equal(mn, a.min(axis=1))


def test_sample_rng_matches_philox_spec():
    ochain, chain, x, th = _setup("readme_n2", 64)
    pc = chain.packed()
    B, seed = 1000, 0x1234_5678_9ABC
    thc = torch.tensor([0.3, -0.2], device=DEV)
    xs = pc.sample_rng(B, seed, None, thc, first_sample=17)
    z = PH.normal_samples(5, B, seed, 0, first_sample=17)
    thb = np.tile(np.array([[0.3], [-0.2]], np.float32), (1, B))
    xo, _ = O.chain_forward(ochain, z, thb, np.float64)
    assert_close(df.to_numpy(xs), xo, 1e-5, 2e-5 * max(1.0, np.abs(xo).max()), "sample_rng")
    # moments of the base draw itself (identity chain = NICE layer with zero weights is overkill; check z stats)
    z_big = PH.normal_samples(5, 200000, 7)
    assert abs(z_big.mean()) < 0.01 and abs(z_big.std() - 1) < 0.01


def test_reference_flow_testset_shapes():
    """test/runtests.jl:97-121 on synthetic fixture-shaped data: train! runs, sample(flow,(2,5,7),(-1,)) has size (5,2,5,7)."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 1000)).astype(np.float32)
    th = np.where(np.arange(1000) < 500, -1.0, 2.0).astype(np.float32)[None, :]
    df.seed(0)
    data = df.DataArrays(x, th, device=DEV)
    chain = df.FlowChain(
        df.CouplingLayer(data, [1, 2, 3], hidden_dim_s=16, hidden_dim_t=16),
        df.CouplingLayer(data, [3, 4, 5], hidden_dim_s=16, hidden_dim_t=16),
        df.CouplingLayer(data, [5, 1, 2], hidden_dim_s=16, hidden_dim_t=16),
        df.NormalizationLayer(data.x, -1.0, 1.0),
    )
    flow = df.Flow(chain, data)
    state = df.setup(df.Adam(1e-3), flow.model)
    df.train_(flow, data, state, epochs=5, verbose=False)
    assert len(flow.train_loss) == 5 and len(flow.valid_loss) == 5
    assert all(np.isfinite(flow.train_loss)) and flow.train_loss[-1] < flow.train_loss[0]
    x_new = df.sample(flow, (2, 5, 7), (-1.0,))
    assert tuple(x_new.shape) == (5, 2, 5, 7)
    assert torch.isfinite(x_new).all()


def test_train_matches_oracle_trajectory():
    """Three epochs of minibatch Adam with a pinned shuffle: loss histories and final weights follow the oracle."""
    x, th = O.synthetic_data(5, 1, 400, seed=3)
    ochain = O.readme_chain(1, x)
    chain = chain_from_oracle(ochain)
    data = df.DataArrays(x, th, device=DEV, rng=torch.Generator().manual_seed(5))
    flow = df.Flow(chain, data)
    state = df.setup(df.Adam(1e-3), flow.model)
    gen = torch.Generator().manual_seed(77)
    df.train_(flow, data, state, epochs=3, batchsize=64, verbose=False, rng=gen)
    # replay on the oracle with the same partition and permutations
    tr = data.partition.training.cpu().numpy()
    va = data.partition.validation.cpu().numpy()
    # train_ draws one seed from the generator and shuffles epoch e on the device with seed + e (oracle/shuffle.py)
    from oracle import shuffle as SH

    gen2 = torch.Generator().manual_seed(77)
    seed0 = int(torch.randint(0, 2**62, (1,), generator=gen2).item())
    perms = [SH.permutation(seed0 + e, len(tr)) for e in range(3)]
    tmin, tmax = th.min(axis=1), th.max(axis=1)
    thn = O.normalize_input(th, tmin, tmax)
    tl, vl = O.train(ochain, x[:, tr], thn[:, tr], x[:, va], thn[:, va], epochs=3, batchsize=64, perms=perms,
                     dtype=np.float64)
    np.testing.assert_allclose(flow.train_loss, tl, rtol=2e-4)
    np.testing.assert_allclose(flow.valid_loss, vl, rtol=2e-4)
    w = flow.packed().W.cpu().numpy()
    np.testing.assert_allclose(w, O.pack_params(ochain), rtol=0, atol=2e-4)


def test_smoke_entry():
    import __graft_entry__ as g

    g.smoke()


def test_fused_peer_allreduce_adam_single_rank_equals_adam_kernel():
    """dflow_dp_* with one rank: the fused all-reduce + Adam kernel (csrc/dflow_dp.cu) must reproduce dflow_adam_step
    bit for bit and pass the reduced loss terms through; two steps exercise the ping-pong halves."""
    import ctypes as C

    lib = df._lib.lib()
    P = 5000
    rng = np.random.default_rng(3)
    w0 = rng.standard_normal(P).astype(np.float32)
    wa, wb = torch.tensor(w0, device=DEV), torch.tensor(w0, device=DEV)
    ma, va, mb, vb = (torch.zeros(P, device=DEV) for _ in range(4))
    dp = C.c_void_p()
    handle = C.create_string_buffer(64)
    df._lib.check(lib.dflow_dp_create(0, 1, P, C.byref(dp), handle))
    df._lib.check(lib.dflow_dp_connect(dp, handle.raw))
    loss2 = torch.zeros(2, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    from densityflows.jl_b200.flows import _DevView

    for t in (1, 2, 3):
        g = torch.tensor(rng.standard_normal(P).astype(np.float32) * 0.1, device=DEV)
        buf = torch.as_tensor(_DevView(int(lib.dflow_dp_grad_buffer(dp)), P + 2), device=DEV)
        buf[:P] = g
        buf[P] = -123.5 * t
        buf[P + 1] = 0.0
        df._lib.check(lib.dflow_dp_allreduce_adam(dp, wa.data_ptr(), ma.data_ptr(), va.data_ptr(), 1e-3, 0.9, 0.999, 1e-8,
                                                  t, loss2.data_ptr(), st))
        df._lib.check(lib.dflow_adam_step(wb.data_ptr(), g.data_ptr(), mb.data_ptr(), vb.data_ptr(), P, 1e-3, 0.9, 0.999,
                                          1e-8, t, st))
        torch.cuda.synchronize()
        assert torch.equal(wa, wb) and torch.equal(ma, mb) and torch.equal(va, vb)
        assert loss2[0].item() == -123.5 * t and loss2[1].item() == 0.0
    assert lib.dflow_dp_status(dp, st) == 0
    df._lib.check(lib.dflow_dp_destroy(dp))


@pytest.mark.parametrize("name", ["readme_n2", "readme_n1", "ref_chain_d7", "blocks_d10_h32"])
@pytest.mark.parametrize("B", [3, 513, 70001])
def test_constant_bank_kernel_bitwise_equals_shared_memory_kernel(name, B):
    """Eligible relu chains (hidden <= 32) default to the constant-bank forward kernel (weights as uniform-datapath
    operands, FFMA2).  It performs the same fmas in the same order as the shared-memory-column kernel, so normalise,
    log-density, sampling and the in-kernel draw must agree bit for bit; ragged tails and several tiles per CTA included."""
    ochain, chain, x, th = _setup(name, B)
    n = th.shape[0]
    pc = chain.packed()
    xj = df.to_jl(x, DEV)
    tj = df.to_jl(th, DEV) if n else None
    thc = torch.full((max(n, 1),), 0.25, device=DEV)
    res = {}
    for mode in (0, -1):
        pc.tune(fwd_const=mode)
        before = pc.launch_count()
        lp = pc.logpdf(xj, tj)
        launches = pc.launch_count() - before
        z, ldj = pc.normalize(xj, tj)
        xs, l2 = pc.forward_ldj(z, tj)
        smp = pc.sample_rng(B, 1234, None, thc if n else None)
        res[mode] = (lp.clone(), z.clone(), ldj.clone(), xs.clone(), l2.clone(), smp.clone(), launches)
    pc.tune(fwd_const=0)
    assert res[0][6] == res[-1][6] == 2, "prepack + one chain kernel either way (bank uploads are memcpys, not kernels)"
    for a, b in zip(res[0][:6], res[-1][:6]):
        assert torch.equal(a, b)


def test_constant_bank_is_safe_across_chains_and_streams():
    """Two different chains share one constant bank per kernel instantiation.  Calls interleaved on two CUDA streams
    must still see their own weights: every (bank upload, kernel) pair is enqueued atomically and ordered after the
    previous user of the bank (dflow_inst.cu)."""
    _, chain_a, xa, tha = _setup("readme_n2", 200001, seed=3)
    _, chain_b, xb, thb = _setup("readme_n1", 150001, seed=4)
    pa, pb = chain_a.packed(), chain_b.packed()
    xa, tha, xb, thb = df.to_jl(xa, DEV), df.to_jl(tha, DEV), df.to_jl(xb, DEV), df.to_jl(thb, DEV)
    ref_a, ref_b = pa.logpdf(xa, tha).clone(), pb.logpdf(xb, thb).clone()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(device=DEV), torch.cuda.Stream(device=DEV)
    outs = []
    for _ in range(20):
        with torch.cuda.stream(s1):
            outs.append(("a", pa.logpdf(xa, tha)))
        with torch.cuda.stream(s2):
            outs.append(("b", pb.logpdf(xb, thb)))
    torch.cuda.synchronize()
    for tag, o in outs:
        assert torch.equal(o, ref_a if tag == "a" else ref_b)


def test_sample_with_rejection():
    """src/Flows.jl:196-229: keep drawing until prod(dims) points satisfy the condition; here in device-side batches
    (draw, mask, compact) instead of one point per iteration.  Error when m * n draws are not enough."""
    ochain, chain, x, th = _setup("readme_n2", 2000)
    data = df.DataArrays(x, th)
    flow = df.Flow(chain, data)
    θ = (0.3, 1.1)
    out = df.sample_with_rejection(7, lambda pts, θ_: pts[0] > 0.0, flow, (50, 4), θ)
    assert tuple(out.shape) == (5, 50, 4)
    assert bool((out[0] > 0).all())
    # accepted points are a prefix-ordered subset of the unconditioned stream of the same seed
    allpts = df.sample(7, flow, 50 * 4 * 8, θ)
    keep = allpts[:, allpts[0] > 0][:, :200]
    assert torch.equal(df.arrays.flat_view(out).reshape(-1, 5).T, keep)
    with pytest.raises(ValueError):
        df.sample_with_rejection(7, lambda pts, θ_: pts[0] > 1e9, flow, 10, θ, 3)


def test_empty_batch_and_unaligned_arrays():
    """B = 0 is a no-op on every entry point (the reference's zero-column arrays), and arrays that are only 4-byte
    aligned (a view into a larger buffer) take the scalar global-memory path of the kernels: same numbers."""
    ochain, chain, x, th = _setup("readme_n2", 1030)
    pc = chain.packed()
    d, n, B = 5, 2, 1030
    # empty
    e_x, e_t = df.jl_empty((d, 0), DEV), df.jl_empty((n, 0), DEV)
    assert pc.logpdf(e_x, e_t).numel() == 0
    z0, l0 = pc.normalize(e_x, e_t)
    assert z0.shape == (d, 0) and l0.numel() == 0
    g = torch.zeros(pc.P, device=DEV)
    l2 = torch.zeros(2, device=DEV)
    assert pc.loss_grad(e_x, e_t, g, l2) == 0 and float(g.abs().sum()) == 0.0
    # unaligned: column-major (d, B) views starting one float into their storage
    xa, ta = df.to_jl(x, DEV), df.to_jl(th, DEV)
    ref = pc.logpdf(xa, ta).clone()
    zr, lr = pc.normalize(xa, ta)
    bx = torch.empty(d * B + 1, device=DEV)
    bt = torch.empty(n * B + 1, device=DEV)
    bx[1:].copy_(df.arrays.flat_view(xa))
    bt[1:].copy_(df.arrays.flat_view(ta))
    xu = bx[1:].view(B, d).t()
    tu = bt[1:].view(B, n).t()
    assert xu.data_ptr() % 16 != 0 and df.arrays.is_colmajor(xu)
    for mode in (0, -1):  # constant-bank kernel and shared-memory kernel
        pc.tune(fwd_const=mode)
        assert torch.equal(pc.logpdf(xu, tu), ref)
        zu, lu = pc.normalize(xu, tu)
        assert torch.equal(zu, zr) and torch.equal(lu, lr)
    pc.tune(fwd_const=0)


def test_train_epoch_from_c_equals_step_by_step():
    """dflow_train_epoch enqueues a whole epoch of minibatch steps (reference default batchsize 64, partial last batch)
    from C; weights, Adam moments and step count must equal the per-minibatch loss_grad + adam_step path bit for bit."""
    ochain, chain_a, x, th = _setup("readme_n2", 1000, seed=5)
    _, chain_b, _, _ = _setup("readme_n2", 1000, seed=5)
    pa, pb = chain_a.packed(), chain_b.packed()
    pa.tune(epoch_kernel=-1)  # the per-minibatch launches (the persistent epoch kernel is compared below, to tolerance)
    assert torch.equal(pa.W, pb.W)
    xj, tj = df.to_jl(x, DEV), df.to_jl(th, DEV)
    order = torch.randperm(900, generator=torch.Generator().manual_seed(3)).to(torch.int32).to(DEV)
    bs, P = 64, pa.P
    ma, va = torch.zeros(P, device=DEV), torch.zeros(P, device=DEV)
    mb, vb = torch.zeros(P, device=DEV), torch.zeros(P, device=DEV)
    loss2 = torch.zeros(2, device=DEV)
    t = 0
    for _ in range(2):
        t = pa.train_epoch(xj, tj, order, bs, ma, va, t, 1e-3, (0.9, 0.999), 1e-8, 0, None, loss2)
    assert t == 2 * 15  # ceil(900 / 64) = 15 minibatches per epoch
    buf = torch.zeros(P + 2, device=DEV)
    tb, acc = 0, torch.zeros(2, device=DEV)
    for _ in range(2):
        for b0 in range(0, 900, bs):
            idx = order[b0:b0 + bs]
            buf.zero_()
            pb.loss_grad(xj, tj, buf[:P], buf[P:], 1.0 / int(idx.numel()), 0, idx)
            acc += buf[P:]
            tb += 1
            pb.adam_step(buf[:P], mb, vb, tb, 1e-3, (0.9, 0.999), 1e-8)
    assert torch.equal(pa.W, pb.W) and torch.equal(ma, mb) and torch.equal(va, vb)
    assert torch.equal(loss2, acc)
    assert float((pa.W - chain_from_oracle(ochain).packed().W).abs().max()) > 0  # the weights did move


@pytest.mark.parametrize("name,bs", [("readme_n2", 64), ("readme_n1", 100), ("hetero_d6_n0", 37), ("blocks_d10_h32", 64),
                                     ("ref_chain_d7", 512)])
def test_persistent_epoch_kernel_matches_step_by_step(name, bs):
    """dflow_train_epoch with minibatches <= 512 runs the whole epoch inside one persistent CTA (csrc/dflow_small.cu:
    parameters + Adam moments + the minibatch's activations in shared memory, a sample spread over 16 / 32 lanes, fixed-
    order weight-gradient reduction).  Same mathematics as loss_grad + adam_step per minibatch -- different summation
    order, so equal to Float32 tolerance, and bitwise reproducible run to run (no atomics)."""
    ochain, chain_a, x, th = _setup(name, 1000, seed=5)
    _, chain_b, _, _ = _setup(name, 1000, seed=5)
    _, chain_c, _, _ = _setup(name, 1000, seed=5)
    pa, pb, pcc = chain_a.packed(), chain_b.packed(), chain_c.packed()
    n = th.shape[0]
    xj = df.to_jl(x, DEV)
    tj = df.to_jl(th, DEV) if n else None
    N = 900
    order = torch.randperm(N, generator=torch.Generator().manual_seed(3)).to(torch.int32).to(DEV)
    P = pa.P
    out = {}
    for tag, pc in (("a", pa), ("c", pcc)):
        m_, v_ = torch.zeros(P, device=DEV), torch.zeros(P, device=DEV)
        loss2 = torch.zeros(2, device=DEV)
        before = pc.launch_count()
        t = 0
        for _ in range(2):
            t = pc.train_epoch(xj, tj, order, bs, m_, v_, t, 1e-3, (0.9, 0.999), 1e-8, 0, None, loss2)
        if name in ("readme_n2", "readme_n1", "hetero_d6_n0"):
            assert pc.launch_count() - before == 2, "one launch per epoch"
        # (chains whose parameters + Adam moments + tape exceed one SM's shared memory stay on the per-minibatch launches)
        assert t == 2 * ((N + bs - 1) // bs)
        out[tag] = (pc.W.clone(), m_, v_, loss2)
    if name in ("readme_n2", "readme_n1", "hetero_d6_n0"):  # the generic path reduces with float atomics
        for i in range(4):
            assert torch.equal(out["a"][i], out["c"][i]), "run-to-run reproducible"
    mb, vb = torch.zeros(P, device=DEV), torch.zeros(P, device=DEV)
    buf = torch.zeros(P + 2, device=DEV)
    tb, acc = 0, torch.zeros(2, device=DEV)
    for _ in range(2):
        for b0 in range(0, N, bs):
            idx = order[b0:b0 + bs]
            buf.zero_()
            pb.loss_grad(xj, tj, buf[:P], buf[P:], 1.0 / int(idx.numel()), 0, idx)
            acc += buf[P:]
            tb += 1
            pb.adam_step(buf[:P], mb, vb, tb, 1e-3, (0.9, 0.999), 1e-8)
    wa, ma, va, l2 = out["a"]
    assert float((wa - pb.W).abs().max()) <= 2e-5, float((wa - pb.W).abs().max())
    assert torch.allclose(ma, mb, rtol=1e-3, atol=1e-6 * float(mb.abs().max()) + 1e-9)
    assert torch.allclose(va, vb, rtol=2e-3, atol=1e-6 * float(vb.abs().max()) + 1e-12)
    assert abs(float(l2[0] - acc[0])) <= 1e-5 * abs(float(acc[0])) and float(l2[1]) == 0
    assert float((wa - chain_from_oracle(ochain).packed().W).abs().max()) > 1e-3  # the weights did move
