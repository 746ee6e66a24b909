"""Latency of one minibatch gradient at the reference's default batch size (64) for different adjoint configurations."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from tests.helpers import chain_from_oracle
d, n = 5, 2
xs, ths = O.synthetic_data(d, n, 4096, seed=1)
chain = chain_from_oracle(O.readme_chain(2, xs))
pc = chain.packed("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
N = 100000
x = df.jl_empty((d, N), "cuda:0"); x.normal_(generator=g)
th = df.jl_empty((n, N), "cuda:0"); th.uniform_(0, 1, generator=g)
grad = torch.zeros(pc.P, device="cuda:0"); l2 = torch.zeros(2, device="cuda:0")
for B in (64, 256, 1024):
    idx = torch.randperm(N, device="cuda:0")[:B].to(torch.int32)
    for tune in [dict(grad_spt=0, grad_threads=0), dict(grad_spt=1, grad_threads=0), dict(grad_spt=1, grad_threads=64),
                 dict(grad_spt=1, grad_threads=32), dict(grad_spt=2, grad_threads=32), dict(grad_spt=2, grad_threads=64), dict(grad_spt=-1, grad_threads=32), dict(grad_spt=-1, grad_threads=64)]:
        pc.tune(**tune)
        for _ in range(5): pc.loss_grad(x, th, grad, l2, None, 0, idx)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): pc.loss_grad(x, th, grad, l2, None, 0, idx)
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"B": B, "tune": tune, "us_per_call": e0.elapsed_time(e1) / 50 * 1e3}), flush=True)
