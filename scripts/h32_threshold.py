import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from tests.helpers import chain_from_oracle
d, n = 10, 3
xs, _ = O.synthetic_data(d, n, 4096, seed=1)
chain = chain_from_oracle(O.block_chain(d, n, 4, 32, xs, s_out_scale=0.5))
pc = chain.packed("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
for B in (32768, 65536, 131072, 262144, 524288):
    x = df.jl_empty((d, B), "cuda:0"); x.normal_(generator=g)
    th = df.jl_empty((n, B), "cuda:0"); th.uniform_(0, 1, generator=g)
    grad = torch.zeros(pc.P, device="cuda:0"); l2 = torch.zeros(2, device="cuda:0")
    r = {}
    for mode in (-1, 1):
        pc.tune(tc_mode=mode)
        for _ in range(3): pc.loss_grad(x, th, grad, l2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): pc.loss_grad(x, th, grad, l2)
        e1.record(); torch.cuda.synchronize()
        r[mode] = e0.elapsed_time(e1) / 5
    print(json.dumps({"B": B, "cuda_ms": r[-1], "tc_ms": r[1]}), flush=True)
