"""Adjoint configurations for a hidden-32 chain (d=10, n=3, 4 blocks): samples/s at a large batch."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from tests.helpers import chain_from_oracle
d, n, B = 10, 3, 1 << 21
xs, _ = O.synthetic_data(d, n, 4096, seed=1)
chain = chain_from_oracle(O.block_chain(d, n, 4, 32, xs, s_out_scale=0.5))
pc = chain.packed("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
x = df.jl_empty((d, B), "cuda:0"); x.normal_(generator=g)
th = df.jl_empty((n, B), "cuda:0"); th.uniform_(0, 1, generator=g)
grad = torch.zeros(pc.P, device="cuda:0"); l2 = torch.zeros(2, device="cuda:0")
for tune in [dict(grad_spt=0, grad_threads=0), dict(grad_spt=2, grad_threads=0), dict(grad_spt=1, grad_threads=0), dict(grad_spt=1, grad_threads=128), dict(tc_mode=1)]:
    pc.tune(**tune)
    for _ in range(2): pc.loss_grad(x, th, grad, l2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): pc.loss_grad(x, th, grad, l2)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"tune": tune, "ms": ms, "sps": B / ms * 1e3}), flush=True)
pc.tune(tc_mode=0)
ms = None
for fc in (0, -1):
    pc.tune(fwd_const=fc)
    for _ in range(2): pc.logpdf(x, th)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): pc.logpdf(x, th)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"fwd_const": fc, "logpdf_ms": ms, "sps": B / ms * 1e3}), flush=True)
