"""Fused s+t conditioner pairs (tc_fuse) vs separate conditioners: launch counts, agreement, timing (C3 shape)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from tests.helpers import chain_from_oracle

def timeit(fn, iters=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

d, n, L, h, B = 16, 4, 8, 64, 1 << 21
xs, _ = O.synthetic_data(d, n, 4096, seed=1)
chain = chain_from_oracle(O.block_chain(d, n, L, h, xs))
pc = chain.packed("cuda:0")
pc.tune(tc_mode=1)
g = torch.Generator(device="cuda").manual_seed(0)
x = df.jl_empty((d, B), "cuda:0"); x.normal_(generator=g)
th = df.jl_empty((n, B), "cuda:0"); th.uniform_(0, 1, generator=g)
res = {}
for fuse in (2, 1, 0):
    pc.tune(tc_fuse=fuse)
    l0 = pc.launch_count()
    lp = pc.logpdf(x, th).clone()
    l1 = pc.launch_count()
    grad = torch.zeros(pc.P, device="cuda:0"); l2 = torch.zeros(2, device="cuda:0")
    pc.loss_grad(x, th, grad, l2)
    l3 = pc.launch_count()
    ms = timeit(lambda: pc.logpdf(x, th))
    g2 = torch.zeros(pc.P, device="cuda:0")
    ms_g = timeit(lambda: pc.loss_grad(x, th, g2, l2), iters=2)
    res[fuse] = (lp, grad.clone(), l2.clone())
    print(json.dumps({"tc_fuse": fuse, "launches_logpdf": l1 - l0, "launches_grad": l3 - l1, "logpdf_ms": ms,
                      "logpdf_sps": B / ms * 1e3, "grad_ms": ms_g, "grad_sps": B / ms_g * 1e3}), flush=True)
a, b = res[2], res[0]
print("logpdf max abs diff", float((a[0] - b[0]).abs().max()), "rel grad diff (max-norm)",
      float((a[1] - b[1]).abs().max() / b[1].abs().max()), "loss", a[2].tolist(), b[2].tolist())
