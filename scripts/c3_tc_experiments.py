"""Timing experiments on the tensor-core forward at C3 (needs a -DDFLOW_TC_EXPERIMENTS build; results are wrong by construction)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from tests.helpers import chain_from_oracle
d, n, L, h, B = 16, 4, 8, 64, 1 << 21
xs, _ = O.synthetic_data(d, n, 4096, seed=1)
chain = chain_from_oracle(O.block_chain(d, n, L, h, xs))
pc = chain.packed("cuda:0")
pc.tune(tc_mode=1)
g = torch.Generator(device="cuda").manual_seed(0)
x = df.jl_empty((d, B), "cuda:0"); x.normal_(generator=g)
th = df.jl_empty((n, B), "cuda:0"); th.uniform_(0, 1, generator=g)
def t(fn, it=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
for name, dbg in [("full", 0), ("no D1", 2), ("no D2", 8), ("no D3", 4), ("no MMA", 14), ("no MMA, no A2 writes", 30),
                  ("no MMA/A2/global", 158), ("no global", 128), ("weights/8", 64), ("no MMA, weights/8", 78)]:
    pc.tune(tc_debug=dbg)
    print(json.dumps({"exp": name, "tc_debug": dbg, "logpdf_ms": round(t(lambda: pc.logpdf(x, th)), 3)}), flush=True)
