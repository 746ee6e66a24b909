"""Generates the committed golden fixtures (run in the build container, where /root/reference exists):

  python tests/golden/make_golden.py

* fixture_datatest.npz -- x (5,1000) and θ (1,1000) Float32 decoded from the reference's own test fixture
  /root/reference/test/datatest.jld2 (JLD2/HDF5 contiguous datasets: x at byte 672, θ at byte 20784).
* golden_<case>.npz -- seeded inputs, packed weights and the ORACLE's outputs (float64 restatement) for the
  normalising direction, log-density, sampling direction and loss gradient.  The reference itself cannot run here
  (no Julia), so these pin the oracle against drift and give the CUDA path a fixed target: "parity unpinned"
  with respect to the Julia implementation (oracle/dflow_oracle.py header).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dflow_oracle as O  # noqa: E402
from tests.golden.cases import readme_n1_chain, ref_chain_d7  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF_FIXTURE = "/root/reference/test/datatest.jld2"


def decode_fixture():
    raw = open(REF_FIXTURE, "rb").read()
    x = np.frombuffer(raw, dtype="<f4", count=5000, offset=672).reshape(1000, 5).T.copy()
    th = np.frombuffer(raw, dtype="<f4", count=1000, offset=20784).reshape(1000, 1).T.copy()
    assert set(np.unique(th)) == {-1.0, 2.0}, np.unique(th)
    return x, th


def case_outputs(chain, x, th, z_in):
    z, ldj = O.chain_backward(chain, x, th, np.float64)
    logp = O.mvnormal_logpdf(z, np.float64) + ldj
    xf, ldjf = O.chain_forward(chain, z_in, th, np.float64)
    loss, grad, _, _ = O.chain_loss_and_grad(chain, x, th, np.float64)
    return dict(W=O.pack_params(chain), x=x, theta=th, z=z, ldj=ldj, logp=logp, z_in=z_in, x_fwd=xf, ldj_fwd=ldjf,
                loss=np.float64(loss), grad=grad)


def main():
    if os.path.exists(REF_FIXTURE):
        x, th = decode_fixture()
        np.savez_compressed(os.path.join(HERE, "fixture_datatest.npz"), x=x, theta=th)
    fx = np.load(os.path.join(HERE, "fixture_datatest.npz"))
    x, th = fx["x"], fx["theta"]
    rng = np.random.default_rng(2024)

    # case 1: the reference's flow testset chain (test/runtests.jl:104-109) on the first 96 fixture samples, n = 1
    chain = readme_n1_chain(x)
    sel = slice(0, 96)
    thn = O.normalize_input(th, th.min(axis=1), th.max(axis=1))
    out = case_outputs(chain, x[:, sel], thn[:, sel], rng.standard_normal((5, 96)).astype(np.float32))
    out["x_min"], out["x_max"] = chain.layers[-1].x_min, chain.layers[-1].x_max
    np.savez_compressed(os.path.join(HERE, "golden_readme_n1.npz"), **out)

    # case 2: the reference's chain testset structure (test/runtests.jl:66-95): unsorted masks, block, normalisation
    xs, ths = O.synthetic_data(7, 2, 80, seed=5)
    chain2 = ref_chain_d7(xs)
    out2 = case_outputs(chain2, xs, ths, rng.standard_normal((7, 80)).astype(np.float32))
    out2["x_min"], out2["x_max"] = chain2.layers[-1].x_min, chain2.layers[-1].x_max
    np.savez_compressed(os.path.join(HERE, "golden_ref_chain_d7.npz"), **out2)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
