"""Random wide-conditioner shapes (hidden 96..512, d <= 64, n <= 32, ragged masks): chain creation, log-density, sampling
round trip and the train step must either work and agree with the Float64 oracle or be refused at chain creation -- never
fail later.  Exploration tool behind tests/test_gpu_wide.py's shape cases."""
import sys, os, json, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from tests.helpers import chain_from_oracle, assert_close

nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 24
bad = 0
for seed in range(nseeds):
    rng = np.random.default_rng(1000 + seed)
    h = int(rng.choice([96, 128, 160, 192, 224, 256, 512]))
    d = int(rng.integers(4, 65))
    n = int(rng.integers(0, 33))
    layers = []
    for li in range(2):
        amax = min(d - 1, 32)
        na = int(rng.integers(max(1, d - (64 - n)), amax + 1)) if d - (64 - n) <= amax else None
        if na is None:
            break
        mask = [int(m) + 1 for m in rng.permutation(d)[:na]]
        layers.append(O.coupling_layer(O.coupling_axes(d, mask, n=n), kind="nice" if rng.random() < 0.2 else "rnvp",
                                       hidden_dim_s=h, hidden_dim_t=h, bias=bool(rng.random() < 0.8), rng=rng, bias_scale=0.1,
                                       s_out_scale=0.2))
    if len(layers) < 2:
        continue
    xn = O.synthetic_data(d, n, 300, seed=seed)[0]
    layers.append(O.norm_layer_from_data(xn, -1.0, 1.0))
    ochain = O.Chain(layers)
    B = int(rng.choice([3, 130, 700]))
    x, th = O.synthetic_data(d, n, B, seed=seed + 7)
    tag = {"seed": seed, "h": h, "d": d, "n": n, "a": [len(l.axes.axis_af) for l in layers[:2]], "B": B}
    try:
        chain = chain_from_oracle(ochain)
        pc = chain.packed()
    except (df.DflowUnsupported, df.DflowInvalidArg) as e:
        print(json.dumps({**tag, "status": "refused at creation", "msg": str(e)[:80]}), flush=True)
        continue
    try:
        xj, tj = df.to_jl(x, "cuda:0"), (df.to_jl(th, "cuda:0") if n else None)
        lp = pc.logpdf(xj, tj)
        zo, lo = O.chain_backward(ochain, x, th, np.float64)
        lpo = O.mvnormal_logpdf(zo, np.float64) + lo
        assert_close(df.to_numpy(lp), lpo, 1e-4, 1e-3, "logpdf")
        grad = torch.zeros(pc.P, device="cuda:0"); l2 = torch.zeros(2, device="cuda:0")
        pc.loss_grad(xj, tj, grad, l2)
        _, go, _, _ = O.chain_loss_and_grad(ochain, x, th, np.float64)
        err = float(np.abs(grad.cpu().numpy() - go).max() / np.abs(go).max())
        assert err <= 5e-4, ("grad", err)
        print(json.dumps({**tag, "status": "ok", "grad_rel": err}), flush=True)
    except Exception as e:  # noqa
        bad += 1
        print(json.dumps({**tag, "status": "FAILED after creation", "msg": repr(e)[:160]}), flush=True)
print("fuzz done, failures:", bad)
