"""C3 (hidden 64) log-density: CUDA-core chain kernel vs tensor-core kernels across batch sizes."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from tests.helpers import chain_from_oracle
d, n, L, h = 16, 4, 8, 64
xs, _ = O.synthetic_data(d, n, 4096, seed=1)
chain = chain_from_oracle(O.block_chain(d, n, L, h, xs))
pc = chain.packed("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
for B in (8192, 16384, 32768, 65536, 131072, 262144, 1048576):
    x = df.jl_empty((d, B), "cuda:0"); x.normal_(generator=g)
    th = df.jl_empty((n, B), "cuda:0"); th.uniform_(0, 1, generator=g)
    r = {}
    for mode in (-1, 1):
        pc.tune(tc_mode=mode)
        for _ in range(3): pc.logpdf(x, th)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): pc.logpdf(x, th)
        e1.record(); torch.cuda.synchronize()
        r[mode] = e0.elapsed_time(e1) / 5
    print(json.dumps({"B": B, "cuda_ms": r[-1], "tc_ms": r[1]}), flush=True)
