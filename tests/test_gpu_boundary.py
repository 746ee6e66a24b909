"""GPU tests of the round-2 boundary additions (include/dflow.h):

* dflow_vjp -- pullback of backward(chain, x, θ) for ARBITRARY cotangents (z̄, j̄), against torch autograd of the Float64
  oracle (the ChainRulesCore.rrule the reference defines at src/affine/RNVP.jl:99-147, composed through the chain);
* dflow_dp_create_local / dflow_dp_train_step / dflow_dp_sync -- one host process driving several replicas;
* the peer barrier's time-out poisons the step instead of reducing stale buffers;
* train_ under torch.distributed with chains built INDEPENDENTLY on every rank (two processes sharing cuda:0, gloo).
"""
import os
import socket

import numpy as np
import pytest
import torch

import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from oracle import torch_ref as T
from tests.helpers import chain_from_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _oracle_vjp(oc, x, th, zbar, jbar, dtype=torch.float64):
    tc = T.TorchChain(oc, dtype)
    for p in tc.params:
        p.requires_grad_(True)
    xt = torch.tensor(x, dtype=dtype, requires_grad=True)
    tt = torch.tensor(th, dtype=dtype, requires_grad=th.shape[0] > 0)
    z, ldj = tc.backward(xt, tt)
    obj = (z * torch.tensor(zbar, dtype=dtype)).sum() + (ldj * torch.tensor(jbar, dtype=dtype)).sum()
    obj.backward()
    tb = tt.grad.numpy() if th.shape[0] > 0 else np.zeros_like(th, dtype=np.float64)
    return tc.flat_grad().numpy().astype(np.float64), xt.grad.numpy(), tb


def _chains():
    def readme(x):
        return O.readme_chain(2, x), 5, 2

    def hetero(x):  # tanh / sigmoid, unequal depth, NICE, no bias, NormalizationLayer in the middle
        rng = np.random.default_rng(5)
        l1 = O.coupling_layer(O.coupling_axes(6, [2, 5], n=1), hidden_dim_s=10, hidden_dim_t=24, n_sublayers_s=1,
                              n_sublayers_t=3, act_s="tanh", act_t="sigmoid", rng=rng, bias_scale=0.2)
        l2 = O.coupling_layer(O.coupling_axes(6, [1, 3, 4, 6], n=1), kind="nice", hidden_dim_t=12, rng=rng, bias_scale=0.2)
        l3 = O.coupling_layer(O.coupling_axes(6, [6, 1, 2, 3, 4], n=1), hidden_dim_s=32, hidden_dim_t=8, bias=False, rng=rng)
        return O.Chain([l1, O.norm_layer_from_data(x, -2.0, 3.0), l2, l3]), 6, 1

    def c3(x):
        return O.block_chain(16, 4, 8, 64, x), 16, 4

    def h128(x):
        return O.block_chain(8, 2, 2, 128, x, s_out_scale=0.3), 8, 2

    def c4like(x):
        return O.block_chain(32, 8, 4, 256, x, s_out_scale=0.3), 32, 8

    return {"readme_n2": (readme, 5, 2), "hetero_d6_n1": (hetero, 6, 1), "c3_cuda_cores": (c3, 16, 4),
            "c3_tensor_cores": (c3, 16, 4), "h128_d8": (h128, 8, 2), "c4_like_h256_L4": (c4like, 32, 8)}


CH = _chains()


@pytest.mark.parametrize("name,B", [("readme_n2", 1), ("readme_n2", 777), ("hetero_d6_n1", 300), ("c3_cuda_cores", 1031),
                                    ("c3_tensor_cores", 32768 + 77), ("h128_d8", 517), ("c4_like_h256_L4", 300)])
@pytest.mark.parametrize("normalize_theta", [False, True])
def test_vjp_arbitrary_cotangents(name, B, normalize_theta):
    mk, d, n = CH[name]
    if normalize_theta and name not in ("readme_n2", "h128_d8"):
        pytest.skip("θ-normalisation chain rule is covered on one narrow and one tensor-core chain")
    xn = O.synthetic_data(d, n, 1000, seed=99)[0]
    oc, _, _ = mk(xn)
    chain = chain_from_oracle(oc)
    x, th = O.synthetic_data(d, n, B, seed=17)
    rng = np.random.default_rng(3)
    if B < 5000:
        zbar = rng.standard_normal((d, B)).astype(np.float32)
        jbar = rng.standard_normal(B).astype(np.float32)
    else:
        # Random-sign cotangents make the sum over 3e4 samples cancel to ~1/sqrt(B) of its terms: even the Float32 oracle is
        # then 1e-3 away from Float64.  Large batches use cotangents that are coherent across samples (a loss-like
        # functional of the outputs: z̄ = (0.6 z + 0.2 sin(3 x_1)) / B, j̄ = (-1 + 0.3 cos(x_2)) / B).
        z64, _ = O.chain_backward(oc, x, th, np.float64)
        zbar = ((0.6 * z64 + 0.2 * np.sin(3.0 * x[0:1])) / B).astype(np.float32)
        jbar = ((-1.0 + 0.3 * np.cos(x[1])) / B).astype(np.float32)
    tmin, tmax = th.min(axis=1), th.max(axis=1)
    pc = df.PackedChain(chain._leaves(), DEV, tmin if normalize_theta else None, tmax if normalize_theta else None)
    flags = df._lib.THETA_NORMALIZE if normalize_theta else 0
    before = pc.launch_count()
    g, xb, tb = pc.vjp(x, th, zbar, jbar, flags, want_θ̄=True)
    launches = pc.launch_count() - before
    if name == "c3_tensor_cores" or "h128" in name or "h256" in name:
        assert launches > 8, "expected the tensor-core adjoint"
    elif name == "c3_cuda_cores":
        assert launches == 2, "prepack + one fused adjoint kernel"
    th_or = O.normalize_input(th, tmin, tmax) if normalize_theta else th
    go, xo, to = _oracle_vjp(oc, x, th_or, zbar, jbar)
    go32, xo32, to32 = _oracle_vjp(oc, x, th_or, zbar, jbar, torch.float32)
    if normalize_theta:  # cotangent of the RAW θ
        rngk = (tmax - tmin).astype(np.float64)[:, None]
        to = np.where(rngk == 0, 0.0, to / np.where(rngk == 0, 1.0, rngk))
        to32 = np.where(rngk == 0, 0.0, to32 / np.where(rngk == 0, 1.0, rngk))
    gg = g.cpu().numpy()[: pc.P].astype(np.float64)
    sl = np.abs(go32 - go).max()
    assert np.abs(gg - go).max() <= 1e-4 * np.abs(go).max() + sl, (np.abs(gg - go).max(), np.abs(go).max(), sl)
    xg, tg = df.to_numpy(xb), df.to_numpy(tb)
    slx = np.abs(xo32 - xo).max()

    def per_sample_ok(got, want, want32, what):
        # x̄ / θ̄ are PER-SAMPLE quantities: a sample with a hidden unit within rounding distance of its ReLU kink gets a
        # different mask under any Float32 evaluation order and its cotangent jumps (the Float32 oracle shows the same
        # jumps).  Every sample must be within 1e-4 of the max-norm unless it is one of those: at most 0.1 % of them.
        tol = 1e-4 * np.abs(want).max() + np.abs(want32 - want).max() * (1.0 if B < 5000 else 0.0)
        bad = (np.abs(got - want) > tol + (1e-5 * np.abs(want).max() if B >= 5000 else 0.0)).any(axis=0)
        assert bad.sum() <= max(0, int(1e-3 * B)) * (1 if B >= 5000 else 0), (what, int(bad.sum()), B, np.abs(got - want).max(),
                                                                               np.abs(want).max())

    per_sample_ok(xg, xo, xo32, "xbar")
    per_sample_ok(tg, to, to32, "thetabar")
    # jbar = None is a zero cotangent of ln_det_jac; the parameter cotangent is accumulated into `grad`
    g2, xb2, _ = pc.vjp(x, th, zbar, None, flags, grad=g.clone())
    go0, xo0, _ = _oracle_vjp(oc, x, th_or, zbar, np.zeros_like(jbar))
    assert np.abs(g2.cpu().numpy()[: pc.P] - (gg + go0)).max() <= 2e-4 * np.abs(go).max() + 2 * sl
    if B < 5000:
        assert np.abs(df.to_numpy(xb2) - xo0).max() <= 1e-4 * np.abs(xo0).max() + slx


def test_vjp_with_loss_seeds_equals_loss_grad():
    """dflow_loss_grad is the special case z̄ = z / B, j̄ = -1 / B of dflow_vjp."""
    xn = O.synthetic_data(5, 2, 1000, seed=99)[0]
    oc = O.readme_chain(2, xn)
    chain = chain_from_oracle(oc)
    pc = chain.packed()
    B = 900
    x, th = O.synthetic_data(5, 2, B, seed=8)
    z, _ = pc.normalize(x, th)
    g_ref = torch.zeros(pc.P, device=DEV)
    l2 = torch.zeros(2, device=DEV)
    pc.loss_grad(x, th, g_ref, l2)
    g, _, _ = pc.vjp(x, th, z / B, torch.full((B,), -1.0 / B, device=DEV))
    assert torch.allclose(g, g_ref, rtol=1e-4, atol=1e-6 * g_ref.abs().max().item())


def _small_problem(N=4096):
    xn = O.synthetic_data(5, 2, 1000, seed=99)[0]
    oc = O.readme_chain(2, xn)
    x, th = O.synthetic_data(5, 2, N, seed=12)
    return oc, x, th


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0]])
def test_single_process_data_parallel_step(devices):
    """dflow_dp_create_local + dflow_dp_train_step: every replica ends bit-identical, and the trajectory equals the
    single-device step on the whole minibatch (up to the summation order of the gradient)."""
    oc, x, th = _small_problem()
    N, nr = x.shape[1], len(devices)
    chain = chain_from_oracle(oc)
    ldp = df.LocalDataParallel(chain, devices)
    per = N // nr
    ldp.set_data([x[:, r * per:(r + 1) * per] for r in range(nr)], [th[:, r * per:(r + 1) * per] for r in range(nr)])
    # reference: one device, whole minibatch through loss_grad + adam_step
    ref = chain_from_oracle(oc)
    pr = ref.packed()
    m, v = torch.zeros(pr.P, device=DEV), torch.zeros(pr.P, device=DEV)
    xj, tj = df.to_jl(x, DEV), df.to_jl(th, DEV)
    gen = torch.Generator().manual_seed(5)
    for t in range(1, 5):
        # minibatch = 600 columns of every shard (ragged: the last rank takes 77 fewer)
        idxs, glob = [], []
        for r in range(nr):
            k = 600 - (77 if r == nr - 1 and nr > 1 else 0)
            loc = torch.randperm(per, generator=gen)[:k].to(torch.int32)
            idxs.append(loc)
            glob.append(loc.to(torch.int64) + r * per)
        Bg = sum(int(i.numel()) for i in idxs)
        ldp.step(idxs, Bg)
        g = torch.zeros(pr.P, device=DEV)
        l2 = torch.zeros(2, device=DEV)
        pr.loss_grad(xj, tj, g, l2, 1.0 / Bg, 0, torch.cat(glob).to(torch.int32).to(DEV))
        pr.adam_step(g, m, v, t)
        ldp.sync()
        for r in range(1, nr):
            assert torch.equal(ldp.replicas[r].W, ldp.replicas[0].W), "replicas must stay bit-identical"
            assert torch.equal(ldp.m[r], ldp.m[0]) and torch.equal(ldp.v[r], ldp.v[0])
        assert torch.allclose(ldp.replicas[0].W, pr.W, rtol=0, atol=2e-6), float((ldp.replicas[0].W - pr.W).abs().max())
        assert abs(ldp.loss2[0][0].item() - l2[0].item()) <= 1e-5 * abs(l2[0].item())
        assert ldp.loss2[0][1].item() == 0
    assert float((pr.W - chain_from_oracle(oc).packed().W).abs().max()) > 1e-3  # the weights did move


def test_peer_barrier_timeout_skips_the_update():
    """A rank whose peer never shows up must NOT sum stale buffers or apply Adam: the step is skipped and the status
    flag raised (dflow_dp_status / LocalDataParallel.sync)."""
    oc, x, th = _small_problem(512)
    chain = chain_from_oracle(oc)
    ldp = df.LocalDataParallel(chain, [0, 0])
    lib = df._lib.lib()
    df._lib.check(lib.dflow_dp_set_timeout_ms(ldp._dps[0], 30))
    w0 = ldp.replicas[0].W.clone()
    st = torch.cuda.current_stream().cuda_stream
    # only rank 0 enters the step; rank 1 never publishes its flag
    buf = int(lib.dflow_dp_grad_buffer(ldp._dps[0]))
    assert buf != 0
    df._lib.check(lib.dflow_dp_allreduce_adam(ldp._dps[0], ldp.replicas[0].W.data_ptr(), ldp.m[0].data_ptr(),
                                              ldp.v[0].data_ptr(), 1e-3, 0.9, 0.999, 1e-8, 1, None, st))
    assert lib.dflow_dp_status(ldp._dps[0], st) == 1
    assert torch.equal(ldp.replicas[0].W, w0) and float(ldp.m[0].abs().sum()) == 0.0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _train_worker(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["DFLOW_DP"] = "nccl"  # gradient all-reduce through torch.distributed (the peer kernel is covered above)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # two ranks share cuda:0: NCCL would refuse that
    torch.cuda.set_device(0)
    torch.manual_seed(1000 + rank)  # every rank has its OWN RNG stream: different weights, partition and shuffles
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 3000)).astype(np.float32)
    th = rng.random((2, 3000)).astype(np.float32)
    data = df.DataArrays(x, th, device="cuda:0")
    chain = df.FlowChain(df.CouplingLayer(data, [1, 2, 3], hidden_dim_s=16, hidden_dim_t=16),
                         df.CouplingLayer(data, [3, 4, 5], hidden_dim_s=16, hidden_dim_t=16),
                         df.NormalizationLayer(data.x, -1.0, 1.0))
    flow = df.Flow(chain, data)
    w_init = flow.packed().W.clone()
    state = df.setup(df.Adam(1e-3), flow.model)
    df.train_(flow, data, state, epochs=2, batchsize=256, verbose=False)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), w=flow.packed().W.cpu().numpy(), w_init=w_init.cpu().numpy(),
             m=state.m.cpu().numpy(), v=state.v.cpu().numpy(), t=state.t, tl=np.array(flow.train_loss),
             vl=np.array(flow.valid_loss))
    dist.barrier()
    dist.destroy_process_group()


def test_train_with_independently_built_ranks(tmp_path):
    """ranks that construct their FlowChain / DataArrays independently used to train diverged
    replicas.  train_ now broadcasts rank 0's parameters, optimiser state, partition and shuffle seed."""
    import torch.multiprocessing as mp

    mp.spawn(_train_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert np.abs(r0["w_init"] - r1["w_init"]).max() > 1e-3, "the ranks really did start from different weights"
    np.testing.assert_array_equal(r0["w"], r1["w"])
    np.testing.assert_array_equal(r0["m"], r1["m"])
    np.testing.assert_array_equal(r0["v"], r1["v"])
    assert int(r0["t"]) == int(r1["t"]) == 2 * ((2700 + 255) // 256)
    np.testing.assert_array_equal(r0["tl"], r1["tl"])
    np.testing.assert_array_equal(r0["vl"], r1["vl"])
    assert np.isfinite(r0["tl"]).all() and r0["tl"][-1] < r0["tl"][0]


# ---- N1: permutations drawn on the device; N3: tensor-product grid generated in-kernel ------------------------------
@pytest.mark.parametrize("n", [1, 2, 5, 64, 1000, 90000, (1 << 20) + 3])
def test_device_shuffle_is_bit_exact_with_its_specification(n):
    """Index work must be bit-exact: dflow_shuffle_indices against the NumPy restatement of the same Feistel bijection,
    whole permutations and rank-style slices, with and without a base index list."""
    from oracle import shuffle as SH

    for seed in (0, 12345, 2**61 + 7):
        ref = SH.permutation(seed, n)
        assert np.array_equal(np.sort(ref), np.arange(n)), "the specification itself must be a bijection"
        got = df.device_permutation(n, seed, DEV).cpu().numpy()
        np.testing.assert_array_equal(got, ref)
        lo, hi = n // 3, n - n // 5
        np.testing.assert_array_equal(df.device_permutation(n, seed, DEV, first=lo, count=hi - lo).cpu().numpy(), ref[lo:hi])
        base = torch.arange(n, dtype=torch.int32).flip(0) * 2
        np.testing.assert_array_equal(df.device_permutation(n, seed, DEV, base=base).cpu().numpy(), base.numpy()[ref])


@pytest.mark.parametrize("name", ["readme_n2", "c3_tensor_cores"])
def test_grid_logpdf_in_kernel(name):
    """logpdf(flow, (v_1, ..., v_d), θ::Tuple) (src/Flows.jl:287-331): values on the tensor-product grid, first vector
    fastest, identical to the materialised (d, prod(lens)) evaluation and within tolerance of the oracle."""
    mk, d, n = CH[name]
    xn = O.synthetic_data(d, n, 1000, seed=99)[0]
    oc, _, _ = mk(xn)
    chain = chain_from_oracle(oc)
    x, th = O.synthetic_data(d, n, 64, seed=2)
    data = df.DataArrays(x, th, device=DEV)
    flow = df.Flow(chain, data)
    rng = np.random.default_rng(1)
    lens = [7, 3, 5, 2, 4] if d == 5 else [4, 3, 2, 5, 3, 2, 2, 3, 1, 2, 3, 2, 1, 2, 2, 3]  # C3: 622 080 points -> tcgen05
    vecs = [np.sort(rng.standard_normal(m)).astype(np.float32) for m in lens]
    θt = tuple(float(v) for v in np.linspace(-0.5, 1.5, n))
    lp = df.logpdf(flow, tuple(vecs), θt)
    assert tuple(lp.shape) == tuple(lens)
    # materialise the grid on the host the way the reference does: Iterators.product, first vector fastest
    B = int(np.prod(lens))
    grids = np.meshgrid(*vecs, indexing="ij")
    pts = np.stack([g.reshape(-1, order="F") for g in grids]).astype(np.float32)
    thb = np.tile(np.array(θt, np.float32)[:, None], (1, B))
    lp_mat = df.logpdf(flow, pts, thb)
    assert torch.equal(df.arrays.flat_view(lp), df.arrays.flat_view(lp_mat))
    cols = np.unique(np.linspace(0, B - 1, 200).astype(np.int64))
    lpo = O.logpdf(oc, pts[:, cols], thb[:, cols], th.min(axis=1), th.max(axis=1), np.float64)
    lpo32 = O.logpdf(oc, pts[:, cols], thb[:, cols], th.min(axis=1), th.max(axis=1), np.float32)
    got = df.arrays.flat_view(lp).cpu().numpy()[cols]
    assert np.abs(got - lpo).max() <= 1e-5 * np.abs(lpo).max() + 2e-5 + np.abs(lpo32 - lpo).max()


def _shard_worker(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["DFLOW_DP"] = "nccl"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 3001)).astype(np.float32)
    th = rng.random((2, 3001)).astype(np.float32)
    df.seed(7)  # same partition on every rank (the shards are slices of the same index lists)
    data = df.DataArrays(x, th, device="cuda:0")
    chain = df.FlowChain(df.CouplingLayer(data, [1, 2, 3], hidden_dim_s=16, hidden_dim_t=16),
                         df.CouplingLayer(data, [3, 4, 5], hidden_dim_s=16, hidden_dim_t=16),
                         df.NormalizationLayer(data.x, -1.0, 1.0))
    flow = df.Flow(chain, data)
    n_before = int(data.x.shape[1])
    data.shard_(rank, world)
    state = df.setup(df.Adam(1e-3), flow.model)
    df.train_(flow, data, state, epochs=2, batchsize=250, verbose=False)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), w=flow.packed().W.cpu().numpy(), t=state.t,
             tl=np.array(flow.train_loss), vl=np.array(flow.valid_loss), resident=int(data.x.shape[1]), total=n_before)
    dist.barrier()
    dist.destroy_process_group()


def test_train_on_sharded_resident_dataset(tmp_path):
    """N1: after data.shard_(rank, world) every rank keeps only its share of the columns resident and shuffles shard-
    locally; replicas stay identical, the epoch-end losses are the global ones, and every sample is visited once."""
    import torch.multiprocessing as mp

    mp.spawn(_shard_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert int(r0["resident"]) + int(r1["resident"]) == int(r0["total"]) == 3001
    assert abs(int(r0["resident"]) - int(r1["resident"])) <= 3
    np.testing.assert_array_equal(r0["w"], r1["w"])
    np.testing.assert_array_equal(r0["tl"], r1["tl"])
    np.testing.assert_array_equal(r0["vl"], r1["vl"])
    assert int(r0["t"]) == int(r1["t"]) == 2 * ((2701 + 249) // 250)
    assert np.isfinite(r0["tl"]).all() and r0["tl"][-1] < r0["tl"][0]
