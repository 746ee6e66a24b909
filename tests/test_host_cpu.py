"""CPU tests of the host mirror and of the C-ABI library surface (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import densityflows.jl_b200 as df
from densityflows.jl_b200 import _lib as L
from densityflows.jl_b200.model import ChainDescriptor
from oracle import dflow_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dflow.h")).read()
    declared = set(re.findall(r"\b(dflow_[a-z_0-9]+)\s*\(", hdr))
    bound = {name for name, _, _ in L.SYMBOLS}
    assert declared == bound, (declared - bound, bound - declared)
    lib = L.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.dflow_version() == 100


def test_axes_are_bit_exact_with_the_oracle():
    for d, mask, n in [(7, [4, 5, 6, 7], 2), (5, [5, 1, 2], 2), (7, [4, 2, 5, 1, 6], 2), (3, [2], 0), (16, list(range(9, 17)), 4)]:
        a, o = df.CouplingAxes(d, mask, n=n), O.coupling_axes(d, mask, n=n)
        assert (a.axis_id, a.axis_af, a.axis_nn) == (o.axis_id, o.axis_af, o.axis_nn)
        r, ro = df.reverse(a), O.reverse_axes(o)
        assert (r.axis_id, r.axis_af, r.axis_nn) == (ro.axis_id, ro.axis_af, ro.axis_nn)
        assert df.is_reverse(a, r)
    # reference testset "axes" (test/runtests.jl:33-41)
    data = df.DataArrays(np.ones((7, 10), np.float32), np.ones((2, 10), np.float32), device="cpu")
    assert df.CouplingAxes(7, [4, 5, 6, 7], n=2) == df.CouplingAxes(7, 3, n=2)
    assert df.CouplingAxes(data) == df.CouplingAxes(7, 3, n=2)
    assert df.CouplingAxes(data, [4, 5, 6, 7]) == df.CouplingAxes(7, 3, n=2)
    assert df.CouplingAxes(data, 3) == df.CouplingAxes(7, 3, n=2)
    assert df.CouplingAxes(7, 3, n=2, reverse=True).axis_af == [1, 2, 3]
    with pytest.raises(AssertionError):
        df.CouplingAxes(3, [4])


def test_coupling_layer_factory_shapes():
    l = df.CouplingLayer(df.RNVPCouplingLayer, 7, [4, 2, 5, 1, 6], n=2)
    assert l.s_net.widths() == [4, 32, 32, 5] and l.t_net.widths() == [4, 32, 32, 5]  # defaults src/Layers.jl:113-123
    assert [dl.act for dl in l.s_net.layers] == [L.ACT_RELU, L.ACT_RELU, L.ACT_IDENTITY]
    # docstring example src/Layers.jl:93-98
    l = df.CouplingLayer(3, [1, 3], n=2, hidden_dim_s=10, hidden_dim_t=10, n_sublayers_s=1, σ_s=df.tanh, σ_t=df.tanh)
    assert l.s_net.widths() == [3, 10, 2] and l.s_net.n_params() == 62
    assert l.t_net.widths() == [3, 10, 10, 2] and l.t_net.n_params() == 172
    nice = df.CouplingLayer(df.NICECouplingLayer, 6, 2, n=1)
    assert isinstance(nice, df.NICECouplingLayer) and nice.t_net.widths() == [3, 32, 32, 4]
    w = l.s_net.layers[0].weight
    s = np.sqrt(6.0 / (3 + 10))
    assert w.shape == (10, 3) and w.stride() == (1, 10) and w.abs().max() <= s  # glorot_uniform, column-major
    assert torch.all(l.s_net.layers[0].bias == 0)
    nb = df.CouplingLayer(4, 2, bias=False)
    assert nb.t_net.layers[0].bias is None


def test_block_and_chain_structure():
    blk = df.CouplingBlock(df.RNVPCouplingLayer, 7, [4, 2, 5, 1], n=2)
    assert blk.layer_2.axes.axis_id == [4, 2, 5, 1] and blk.layer_2.axes.axis_af == [3, 6, 7]
    assert blk.layer_2.t_net.widths()[0] == 2 + 4 and len(blk) == 2
    l1 = df.CouplingLayer(7, [1, 3, 5, 7], n=2)
    l2 = df.CouplingLayer(7, [4, 2, 5, 1, 6], n=2)
    with pytest.raises(ValueError):  # src/Blocks.jl:70-73
        df.CouplingBlock(l1, l2)
    small = df.FlowChain(l1, l2)
    assert len(df.concatenate(small, blk)) == 3 and len(df.concatenate(blk, small)) == 3
    assert isinstance(small[0], df.CouplingLayerBase)
    x1 = np.full((7, 10), 0.2, np.float32)
    x1[:, 1] = 0.4
    chain = df.concatenate((small, df.FlowChain(blk, df.NormalizationLayer(x1))))
    assert isinstance(chain, df.FlowChain) and isinstance(chain[-1], df.NormalizationElement)
    assert [type(e).__name__ for e in chain._leaves()] == ["RNVPCouplingLayer"] * 4 + ["NormalizationLayer"]
    np.testing.assert_array_equal(chain[-1].x_min, np.full(7, 0.2, np.float32))
    np.testing.assert_array_equal(chain[-1].x_max, np.full(7, 0.4, np.float32))
    assert len(df.FlowChain(3, 6, 3, n=1)) == 3  # FlowChain(n, args...; kws...) of CouplingBlocks, src/Chains.jl:100-101
    with pytest.raises(AssertionError):
        df.NormalizationLayer(x1, 1.0, 0.0)


def test_data_containers():
    x = 0.2 * np.ones((7, 10), np.float32)
    th = 0.1 * np.ones((2, 10), np.float32)
    x[0, 1] = 0.3
    th[0, 1] = 0.4
    assert tuple(df.dflt_θ(x).shape) == (0, 10)
    df.seed(1)
    data = df.DataArrays(x, th, device="cpu")
    assert df.number_dimensions(data) == 7 and df.number_conditions(data) == 2
    p = data.partition
    assert len(p.training) == 9 and len(p.validation) == 1 and len(p.testing) == 0
    assert sorted(torch.cat([p.training, p.validation]).tolist()) == list(range(10))
    md = df.MetaData("", 7, 2, df.minimum_θ(data), df.maximum_θ(data))
    np.testing.assert_array_equal(md.θ_min, th.min(axis=1))
    x_t, th_t = df.normalized_training_data(data, md)
    assert th_t.max() <= 1 and th_t.min() >= 0 and torch.all(th_t[1] == 0)
    np.testing.assert_allclose(df.to_numpy(th_t), O.normalize_input(th[:, p.training.numpy()], md.θ_min, md.θ_max))
    # rounding of the split sizes is Julia's round (ties to even): 25 * 0.9 = 22.5 -> 22, then 2.5 -> 2
    p2 = df.DataPartition(25, 0.9, 0.1)
    assert (len(p2.training), len(p2.validation), len(p2.testing)) == (22, 2, 1)
    with pytest.raises(AssertionError):
        df.DataArrays(np.ones((3, 4), np.float32), np.ones((1, 5), np.float32), device="cpu")


def test_julia_shaped_arrays_are_sample_contiguous():
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    t = df.to_jl(a)
    assert tuple(t.shape) == (2, 3, 4) and t.stride() == (1, 2, 6)
    flat = df.arrays.flat_view(t).numpy()
    np.testing.assert_array_equal(flat, a.reshape(-1, order="F"))
    e = df.jl_empty((5, 2, 5, 7))
    assert tuple(e.shape) == (5, 2, 5, 7) and e.stride()[0] == 1 and df.arrays.n_samples(e) == 70


def _desc_for(leaves):
    return ChainDescriptor(leaves)


def test_descriptor_and_error_codes_without_a_device():
    lib = L.lib()
    h = C.c_void_p()
    # valid descriptor contents
    l = df.CouplingLayer(5, [5, 1, 2], n=2, hidden_dim_s=16, hidden_dim_t=16)
    cd = _desc_for([l])
    e = cd.elems[0]
    assert (cd.d, cd.n, e.kind, e.n_af) == (5, 2, L.ELEM_RNVP, 3)
    assert [e.axis_af[i] for i in range(3)] == [4, 0, 1] and [e.axis_id[i] for i in range(2)] == [2, 3]
    assert [e.s_net.widths[i] for i in range(4)] == [4, 16, 16, 3]
    # invalid arguments are rejected before any device work (reference: @assert / ArgumentError)
    assert lib.dflow_chain_create(None, C.byref(h)) == L.E_INVALID_ARG
    bad = _desc_for([l])
    bad.elems[0].axis_af[0] = 9
    assert lib.dflow_chain_create(C.byref(bad.desc), C.byref(h)) == L.E_INVALID_ARG
    assert b"mask cannot contain values higher than the dimension" in lib.dflow_last_error()
    bad = _desc_for([l])
    bad.elems[0].axis_af[1] = 4  # duplicate
    assert lib.dflow_chain_create(C.byref(bad.desc), C.byref(h)) == L.E_INVALID_ARG
    bad = _desc_for([l])
    bad.elems[0].s_net.widths[0] = 7  # conditioner input does not match length(axis_nn)
    assert lib.dflow_chain_create(C.byref(bad.desc), C.byref(h)) == L.E_INVALID_ARG
    # hidden > 64 runs on the tensor-core path, which needs a multiple of 32 (dflow_tc.cu): 100 is refused up front
    wide = df.CouplingLayer(5, [1, 2], n=0, hidden_dim_s=100, hidden_dim_t=100)
    assert lib.dflow_chain_create(C.byref(_desc_for([wide]).desc), C.byref(h)) == L.E_UNSUPPORTED
    assert b"multiple of 32" in lib.dflow_last_error()
    with pytest.raises(df.DflowUnsupported):
        L.check(L.E_UNSUPPORTED)
    act = _desc_for([l])
    act.elems[0].t_net.acts[0] = 17
    assert lib.dflow_chain_create(C.byref(act.desc), C.byref(h)) == L.E_UNSUPPORTED
    nrm = df.NormalizationLayer(np.zeros(5, np.float32), np.ones(5, np.float32), 0.0, 1.0)
    nd = _desc_for([l, nrm])
    nd.elems[1].beta = -1.0
    assert lib.dflow_chain_create(C.byref(nd.desc), C.byref(h)) == L.E_INVALID_ARG
    # null / negative arguments on compute entry points
    assert lib.dflow_adam_step(None, None, None, None, 10, 1e-3, 0.9, 0.999, 1e-8, 1, None) == L.E_INVALID_ARG
    assert lib.dflow_minmax(None, 0, 10, None, None, None) == L.E_INVALID_ARG
    assert lib.dflow_logpdf(None, None, None, None, 1, None, 0, None, None) == L.E_INVALID_ARG
    assert lib.dflow_param_count(None) == L.E_INVALID_ARG


def test_packed_chain_requires_cuda():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    l = df.CouplingLayer(5, [1, 2], n=0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        df.backward(l, np.zeros((5, 3), np.float32))


def test_custom_elements_are_rejected_not_emulated():
    class NewLayer(df.FlowElement):  # docs/src/documentation.md:170-197
        pass

    with pytest.raises(NotImplementedError):
        ChainDescriptor([NewLayer()])


def _split_top_level(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_julia_glue_matches_the_c_abi():
    """julia/DensityFlowsB200.jl cannot be executed here (no Julia), so at least keep it ABI-consistent: every ccall names
    an exported symbol, passes as many arguments as the ctypes binding (which the GPU tests exercise) declares, and uses
    pointer / scalar classes in the same positions.  Also the oracle's permutation is a bijection (device shuffle spec)."""
    src = open(os.path.join(ROOT, "julia", "DensityFlowsB200.jl")).read()
    assert "using Random" in src
    bound = {name: args for name, _, args in L.SYMBOLS}
    calls = list(re.finditer(r"ccall\(\(:(dflow_[a-z_0-9]+), libdflow\),\s*([A-Za-z{}]+),\s*\(", src))
    assert len(calls) >= 18
    seen = set()
    for m in calls:
        name = m.group(1)
        assert name in bound, f"{name} is not exported by libdflow.so"
        seen.add(name)
        i, depth = m.end(), 1
        while depth:  # the argument-type tuple
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        types = _split_top_level(src[m.end():i - 1])
        assert len(types) == len(bound[name]), (name, types, len(bound[name]))
        for jt, ct in zip(types, bound[name]):
            is_ptr_j = any(k in jt for k in ("Ptr", "Ref", "Cstring"))
            is_ptr_c = ct in (L.vp, L.C.c_char_p) or hasattr(ct, "_type_") and not isinstance(ct._type_, str)
            assert is_ptr_j == is_ptr_c, (name, jt, ct)
    for must in ("dflow_vjp", "dflow_train_epoch", "dflow_dp_create_local", "dflow_dp_train_step", "dflow_logpdf_grid",
                 "dflow_shuffle_indices", "dflow_chain_set_scratch", "dflow_sample_rng"):
        assert must in seen, must
    assert "ChainRulesCore.rrule(::typeof(backward)" in src
    from oracle import shuffle as SH

    for n in (1, 2, 3, 17, 4096, 12345):
        assert sorted(SH.permutation(99, n).tolist()) == list(range(n))
        np.testing.assert_array_equal(SH.permutation(99, n, 3 % n, max(0, n - 5)), SH.permutation(99, n)[3 % n: 3 % n + max(0, n - 5)])


def test_save_and_load_flow_round_trip(tmp_path):
    """Persistence of the mirror (counterpart of src/Loading.jl save_flow / load_flow): structure, axes (incl. un-sorted
    masks and a block), activations, bias flags, the packed parameter vector, metadata and loss histories survive."""
    df.seed(3)
    x = np.random.default_rng(0).standard_normal((7, 50)).astype(np.float32)
    th = np.random.default_rng(1).random((2, 50)).astype(np.float32)
    data = df.DataArrays(x, th, device="cpu")
    chain = df.FlowChain(df.CouplingLayer(data, [4, 2, 5, 1, 6], hidden_dim_s=8, hidden_dim_t=12, n_sublayers_t=3, σ_s="tanh"),
                         df.CouplingBlock(data, [1, 3, 5, 7], hidden_dim_s=16, hidden_dim_t=16, bias=False),
                         df.CouplingLayer(df.NICECouplingLayer, data, [2, 3], hidden_dim_t=8),
                         df.NormalizationLayer(data.x, -1.0, 1.0))
    flow = df.Flow(chain, data, train_loss=[3.0, 2.5], valid_loss=[3.1, 2.6])
    p = str(tmp_path / "flow.npz")
    df.save_flow(p, flow)
    back = df.load_flow(p, device="cpu")
    np.testing.assert_array_equal(df.packed_parameters(back.model), df.packed_parameters(flow.model))
    assert df.packed_parameters(flow.model).size == O.pack_params(O.Chain([])).size + sum(
        l.weight.numel() + (l.bias.numel() if l.bias is not None else 0)
        for e in chain._leaves() for net in df.model._nets_of(e) for l in net.layers)
    la, lb = flow.model._leaves(), back.model._leaves()
    assert [type(e).__name__ for e in la] == [type(e).__name__ for e in lb]
    for ea, eb in zip(la, lb):
        if hasattr(ea, "axes"):
            assert (ea.axes.axis_id, ea.axes.axis_af, ea.axes.axis_nn) == (eb.axes.axis_id, eb.axes.axis_af, eb.axes.axis_nn)
            for na, nb in zip(df.model._nets_of(ea), df.model._nets_of(eb)):
                assert na.widths() == nb.widths() and [l.act for l in na.layers] == [l.act for l in nb.layers]
                assert [l.bias is None for l in na.layers] == [l.bias is None for l in nb.layers]
        else:
            np.testing.assert_array_equal(ea.x_min, eb.x_min)
            assert (ea.α, ea.β) == (eb.α, eb.β)
    assert back.train_loss == [3.0, 2.5] and back.valid_loss == [3.1, 2.6]
    np.testing.assert_array_equal(back.metadata.θ_min, flow.metadata.θ_min)
    assert (back.d, back.n) == (7, 2)
