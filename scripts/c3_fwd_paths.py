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
g = torch.Generator(device="cuda").manual_seed(0)
x = df.jl_empty((d, B), "cuda:0"); x.normal_(generator=g)
th = df.jl_empty((n, B), "cuda:0"); th.uniform_(0, 1, generator=g)
for mode, ts in ((-1, 0), (1, -1), (1, 0)):
    pc.tune(tc_mode=mode, tc_ts=ts)
    ms = timeit(lambda: pc.logpdf(x, th))
    print(json.dumps({"tc_mode": mode, "tc_ts": ts, "logpdf_ms": ms, "sps": B / ms * 1e3}))
