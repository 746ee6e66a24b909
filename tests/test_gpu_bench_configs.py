"""GPU parity at the BASELINE shapes, depths and batch sizes that bench.py times, on the DEFAULT-ROUTED kernels:

  C3  d=16 n=4  L=8  h=64   -> tcgen05 3xTF32 forward from B >= 65536, tcgen05 adjoint from B >= 8192
  C4  d=32 n=8  L=12 h=256  -> tcgen05 forward + adjoint
  C5  d=64 n=16 L=16 h=512  -> tcgen05 forward / sampling (two 256-column passes)

Batches span several 128-sample tiles per CTA and end in a ragged tail.  The Float64 oracle is evaluated on a subset
of columns (head, strided middle, the whole ragged tail), which keeps it to seconds.  Gradients use the linearity of
the loss in the samples: the batch is a gather (idx) with repetitions of Nd distinct columns, so that the oracle's
count-weighted gradient over the Nd columns IS the gradient of the whole batch.  (Nd must not be small: a hidden unit
within the kernels' rounding distance of its ReLU kink flips its mask, and a repeated column carries that flip with
its whole weight count/B -- 320 distinct columns at C4 inflated the error to 3e-4 of the max-norm.)

Also here: the committed golden vectors (tests/golden/*.npz, fixture included) against the CUDA path, and the
host-buffer entry points dflow_logpdf_host / dflow_sample_host.

Measured maximum errors are written to gpurun_out/parity_r02.json (table in DESIGN.md).
"""
import json
import os

import numpy as np
import pytest
import torch

import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from oracle import philox as PH
from oracle import torch_ref as T
from tests.golden.cases import readme_n1_chain, ref_chain_d7
from tests.helpers import assert_close, chain_from_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

# name -> (d, n, L, h, columns of the NormalizationLayer draw: the same chains bench.py builds)
CFG = {"c3": (16, 4, 8, 64, 65536), "c4": (32, 8, 12, 256, 8192), "c5": (64, 16, 16, 512, 8192)}
_ERR = {}


def _record(key, **kv):
    _ERR.setdefault(key, {}).update({k: float(v) for k, v in kv.items()})
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, "parity_r02.json")
        old = json.load(open(path)) if os.path.exists(path) else {}
        old.setdefault(key, {}).update(_ERR[key])
        json.dump(old, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


_chain_cache = {}


def _bench_chain(name):
    if name not in _chain_cache:
        d, n, L, h, nx = CFG[name]
        xs, _ = O.synthetic_data(d, n, nx, seed=1234)
        oc = O.block_chain(d, n, L, h, xs)
        _chain_cache[name] = (oc, chain_from_oracle(oc))
    return _chain_cache[name]


def _subset(B, head=48, mid=96, tail=77 + 51):
    cols = np.unique(np.concatenate([np.arange(min(head, B)), np.linspace(0, B - 1, mid).astype(np.int64),
                                     np.arange(max(0, B - tail), B)]))
    return cols


def _relerr(got, want):
    return float(np.max(np.abs(got - want) / (np.abs(want) + 1.0)))


@pytest.mark.parametrize("name,B", [("c3", 131072 + 77), ("c3", 4 * 148 * 128 + 1), ("c4", 2 * 148 * 128 + 77),
                                    ("c5", 2 * 148 * 128 + 77)])
def test_normalize_logpdf_default_route_full_depth(name, B):
    d, n, L, h, _ = CFG[name]
    oc, chain = _bench_chain(name)
    x, th = O.synthetic_data(d, n, B, seed=31)
    pc = chain.packed()
    xj, tj = df.to_jl(x, DEV), df.to_jl(th, DEV)
    z, ldj = pc.normalize(xj, tj)
    lp = pc.logpdf(xj, tj)
    cols = _subset(B)
    zo, lo = O.chain_backward(oc, x[:, cols], th[:, cols], np.float64)
    zo32, lo32 = O.chain_backward(oc, x[:, cols], th[:, cols], np.float32)
    lpo = O.mvnormal_logpdf(zo, np.float64) + lo
    lpo32 = O.mvnormal_logpdf(zo32, np.float32) + lo32
    zg, lg, lpg = df.to_numpy(z)[:, cols], df.to_numpy(ldj)[cols], df.to_numpy(lp)[cols]
    slack_z = np.abs(zo32 - zo).max() + np.abs(lo32 - lo).max()
    slack_lp = np.abs(lpo32 - lpo).max()
    _record(f"{name}_normalize_B{B}", z_rel=_relerr(zg, zo), ldj_rel=_relerr(lg, lo), logp_rel=_relerr(lpg, lpo),
            f32_oracle_z_rel=_relerr(zo32, zo), f32_oracle_logp_rel=_relerr(lpo32, lpo))
    assert_close(zg, zo, 1e-5, 1e-5 + slack_z, f"{name} z")
    assert_close(lg, lo, 1e-5, 1e-5 + slack_z, f"{name} ldj")
    # rtol 1e-5 on the value + the Float32 oracle's own distance from Float64 (|logp| is O(10-100) here)
    assert_close(lpg, lpo, 1e-5, 1e-5 + slack_lp, f"{name} logpdf")
    # the batch really is multi-tile / ragged on the tensor-core route
    assert B % 128 != 0 and B > 148 * 128


@pytest.mark.parametrize("name,B", [("c3", 131072 + 77), ("c4", 2 * 148 * 128 + 77), ("c5", 2 * 148 * 128 + 77)])
def test_sampling_default_route_full_depth(name, B):
    d, n, L, h, _ = CFG[name]
    oc, chain = _bench_chain(name)
    _, th = O.synthetic_data(d, n, B, seed=32)
    z = (np.random.default_rng(2).standard_normal((d, B)) * 0.8).astype(np.float32)
    pc = chain.packed()
    zj, tj = df.to_jl(z, DEV), df.to_jl(th, DEV)
    xg, ldj = pc.forward_ldj(zj, tj)
    cols = _subset(B)
    xo, lo = O.chain_forward(oc, z[:, cols], th[:, cols], np.float64)
    xo32, lo32 = O.chain_forward(oc, z[:, cols], th[:, cols], np.float32)
    scale = max(1.0, float(np.abs(xo).max()))
    slack = np.abs(xo32 - xo).max() + np.abs(lo32 - lo).max()
    _record(f"{name}_sample_B{B}", x_rel=_relerr(df.to_numpy(xg)[:, cols], xo), ldj_rel=_relerr(df.to_numpy(ldj)[cols], lo),
            f32_oracle_x_rel=_relerr(xo32, xo))
    assert_close(df.to_numpy(xg)[:, cols], xo, 1e-5, 1e-5 * scale + slack, f"{name} forward x")
    assert_close(df.to_numpy(ldj)[cols], lo, 1e-5, 1e-5 + slack, f"{name} forward ldj")
    # forward! (in place, no ldj) == forward
    zt = zj.clone()
    pc.sample_inplace(zt, tj)
    assert torch.equal(zt, xg)
    # in-kernel Philox draw with a fixed condition (the bench's sample() call): same stream as the oracle's spec
    thc = np.linspace(-0.5, 1.5, n).astype(np.float32)
    xs = pc.sample_rng(B, 777, None, torch.tensor(thc, device=DEV), first_sample=5)
    zs = PH.normal_samples(d, B, 777, 0, first_sample=5)[:, cols]
    xso, _ = O.chain_forward(oc, zs, np.tile(thc[:, None], (1, len(cols))), np.float64)
    assert_close(df.to_numpy(xs)[:, cols], xso, 1e-5, 2e-5 * max(1.0, float(np.abs(xso).max())) + slack, f"{name} sample_rng")


def _weighted_oracle_grad(oc, x, th, counts, inv_btot, dtype):
    tc = T.TorchChain(oc, dtype)
    for p in tc.params:
        p.requires_grad_(True)
    lp = tc.logpdf(torch.tensor(x, dtype=dtype), torch.tensor(th, dtype=dtype))
    w = torch.tensor(counts, dtype=dtype)
    loss = -(w * lp).sum() * inv_btot
    loss.backward()
    return float(loss), tc.flat_grad().numpy().astype(np.float64), float((w * lp).sum())


@pytest.mark.parametrize("name,B,Nd", [("c3", 32768 + 4 * 148 * 128 + 77, 1536), ("c4", 148 * 128 + 2 * 128 + 77, 4096)])
def test_loss_grad_default_route_full_depth(name, B, Nd):
    d, n, L, h, _ = CFG[name]
    oc, chain = _bench_chain(name)
    x, th = O.synthetic_data(d, n, Nd, seed=33)
    rng = np.random.default_rng(9)
    idx = rng.integers(0, Nd, size=B).astype(np.int32)
    counts = np.bincount(idx, minlength=Nd).astype(np.float64)
    pc = chain.packed()
    grad = torch.zeros(pc.P, device=DEV)
    loss2 = torch.zeros(2, device=DEV)
    before = pc.launch_count()
    pc.loss_grad(df.to_jl(x, DEV), df.to_jl(th, DEV), grad, loss2, 1.0 / B, 0, torch.tensor(idx, device=DEV))
    assert pc.launch_count() - before > 3 * L, "expected the tensor-core adjoint (one launch per layer and sweep)"
    lo, go, slp = _weighted_oracle_grad(oc, x, th, counts, 1.0 / B, torch.float64)
    _, go32, _ = _weighted_oracle_grad(oc, x, th, counts, 1.0 / B, torch.float32)
    g = grad.cpu().numpy().astype(np.float64)
    assert loss2[1].item() == 0
    assert abs(loss2[0].item() - slp) <= 2e-5 * abs(slp) + 1e-3
    slack = np.abs(go32 - go).max()
    gerr = np.abs(g - go).max()
    _record(f"{name}_grad_B{B}", grad_maxnorm_rel=gerr / np.abs(go).max(), f32_oracle_rel=slack / np.abs(go).max(),
            loss_rel=abs(-loss2[0].item() / B - lo) / abs(lo))
    assert gerr <= 1e-4 * np.abs(go).max() + slack, (gerr, np.abs(go).max(), slack)
    off = 0
    for e in O.flatten(oc):
        for net in O._trainable_nets(e):
            for dl in net:
                k = dl.W.size + (dl.b.size if dl.b is not None else 0)
                ref = go[off:off + k]
                err = np.abs(g[off:off + k] - ref).max()
                # per Dense: 2e-4 of the block's own max-norm, on top of a floor of 2e-5 of the whole gradient's max-norm: the
                # Float32 accumulation over ~1e5 samples (in TMEM here, in the sgemm of the reference's Dense pullback alike)
                assert err <= 2e-4 * np.abs(ref).max() + 2e-5 * np.abs(go).max() + slack, (name, off, err, np.abs(ref).max())
                off += k


# ---- committed golden vectors against the CUDA path ---------------------------------------------------------------
@pytest.mark.parametrize("case", ["golden_readme_n1", "golden_ref_chain_d7"])
def test_cuda_path_reproduces_golden_vectors(case):
    g = np.load(os.path.join(GOLD, case + ".npz"))
    if case == "golden_readme_n1":
        fx = np.load(os.path.join(GOLD, "fixture_datatest.npz"))
        oc = readme_n1_chain(fx["x"])
        # the golden inputs ARE the first 96 columns of the reference's fixture, θ normalised (src/Data.jl:213-218)
        np.testing.assert_array_equal(g["x"], fx["x"][:, :96])
    else:
        oc = ref_chain_d7(g["x"])
    np.testing.assert_array_equal(O.pack_params(oc), g["W"])  # same seeded weights as the generator script
    chain = chain_from_oracle(oc)
    pc = chain.packed()
    np.testing.assert_array_equal(pc.W.cpu().numpy(), g["W"])
    x, th = g["x"], g["theta"]
    z, ldj = pc.normalize(x, th)
    assert_close(df.to_numpy(z), g["z"], 1e-5, 1e-5, case + " z")
    assert_close(df.to_numpy(ldj), g["ldj"], 1e-5, 1e-5, case + " ldj")
    assert_close(df.to_numpy(pc.logpdf(x, th)), g["logp"], 1e-5, 2e-5, case + " logp")
    xf, lf = pc.forward_ldj(g["z_in"], th)
    assert_close(df.to_numpy(xf), g["x_fwd"], 1e-5, 1e-5 * max(1.0, np.abs(g["x_fwd"]).max()), case + " x_fwd")
    assert_close(df.to_numpy(lf), g["ldj_fwd"], 1e-5, 1e-5, case + " ldj_fwd")
    grad = torch.zeros(pc.P, device=DEV)
    loss2 = torch.zeros(2, device=DEV)
    B = x.shape[1]
    pc.loss_grad(x, th, grad, loss2)
    assert abs(-loss2[0].item() / B - float(g["loss"])) <= 1e-5 * abs(float(g["loss"])) + 1e-5
    gg = grad.cpu().numpy()
    assert np.abs(gg - g["grad"]).max() <= 1e-4 * np.abs(g["grad"]).max()


def test_fixture_flow_logpdf_through_flow_api():
    """The reference's fixture (test/datatest.jld2, θ ∈ {-1, 2}) through the Flow-level API with in-kernel θ
    normalisation, against the oracle on every one of its 1000 columns."""
    fx = np.load(os.path.join(GOLD, "fixture_datatest.npz"))
    x, th = fx["x"], fx["theta"]
    oc = readme_n1_chain(x)
    chain = chain_from_oracle(oc)
    data = df.DataArrays(x, th, device=DEV)
    flow = df.Flow(chain, data)
    lp = df.logpdf(flow, x, th)
    lpo = O.logpdf(oc, x, th, th.min(axis=1), th.max(axis=1), np.float64)
    assert_close(df.to_numpy(lp), lpo, 1e-5, 2e-5, "fixture logpdf")


# ---- host-buffer entry points (the bench's e2e call) ----------------------------------------------------------------
@pytest.mark.parametrize("name,B,chunk", [("readme", 100003, 1 << 14), ("readme", 5, 0), ("c4", 3000, 1024)])
def test_logpdf_host_and_sample_host(name, B, chunk):
    import ctypes as C

    lib = df._lib.lib()
    if name == "readme":
        d, n = 5, 2
        xs, _ = O.synthetic_data(d, n, 4096, seed=1234)
        oc = O.readme_chain(n, xs)
        chain = chain_from_oracle(oc)
    else:
        d, n = CFG[name][:2]
        oc, chain = _bench_chain(name)
    x, th = O.synthetic_data(d, n, B, seed=44)
    tmin, tmax = th.min(axis=1), th.max(axis=1)
    pc = df.PackedChain(chain._leaves(), DEV, tmin, tmax)
    flags = df._lib.THETA_NORMALIZE
    # pageable and pinned host memory, sample-major (Julia column-major) flat buffers
    xh = torch.from_numpy(np.ascontiguousarray(x.T).reshape(-1)).pin_memory()
    thh = torch.from_numpy(np.ascontiguousarray(th.T).reshape(-1))
    oh = torch.empty(B, dtype=torch.float32).pin_memory()
    with torch.cuda.device(DEV):
        df._lib.check(lib.dflow_logpdf_host(pc.handle, pc.W.data_ptr(), xh.data_ptr(), thh.data_ptr(), B, flags,
                                            oh.data_ptr(), chunk))
    dev = pc.logpdf(df.to_jl(x, DEV), df.to_jl(th, DEV), flags)
    assert torch.equal(oh, dev.cpu()), "host pipeline and device entry point must agree bit for bit"
    cols = _subset(B, 32, 64, 40)
    lpo = O.logpdf(oc, x[:, cols], th[:, cols], tmin, tmax, np.float64)
    lpo32 = O.logpdf(oc, x[:, cols], th[:, cols], tmin, tmax, np.float32)
    assert_close(oh.numpy()[cols], lpo, 1e-5, 2e-5 + np.abs(lpo32 - lpo).max(), "dflow_logpdf_host")
    # dflow_sample_host: Philox stream continues across chunks (counter = sample index), fixed condition
    thc = np.linspace(-0.5, 1.5, n).astype(np.float32)
    thc_h = torch.from_numpy(thc)
    xo_h = torch.empty(B * d, dtype=torch.float32).pin_memory()
    with torch.cuda.device(DEV):
        df._lib.check(lib.dflow_sample_host(pc.handle, pc.W.data_ptr(), C.c_uint64(4242), thc_h.data_ptr(), B, flags,
                                            xo_h.data_ptr(), chunk))
    xdev = pc.sample_rng(B, 4242, None, torch.tensor(thc, device=DEV), flags)
    assert torch.equal(xo_h.view(B, d).t(), xdev.cpu())
    zs = PH.normal_samples(d, B, 4242, 0)[:, cols]
    thn = O.normalize_input(np.tile(thc[:, None], (1, len(cols))), tmin, tmax)
    xso, _ = O.chain_forward(oc, zs, thn, np.float64)
    xso32, _ = O.chain_forward(oc, zs, thn, np.float32)
    assert_close(xo_h.view(B, d).t().numpy()[:, cols], xso, 1e-5,
                 2e-5 * max(1.0, float(np.abs(xso).max())) + np.abs(xso32 - xso).max(), "dflow_sample_host")
