"""One loss_grad call on a named config (for ncu launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from tests.helpers import chain_from_oracle
CFG = {"c4": (32, 8, 12, 256, 1 << 18), "c3": (16, 4, 8, 64, 1 << 21), "c3w": (16, 4, 8, 128, 1 << 19)}
name = sys.argv[1]
d, n, L, h, B = CFG[name]
xs, _ = O.synthetic_data(d, n, 4096, seed=1)
chain = chain_from_oracle(O.block_chain(d, n, L, h, xs))
pc = chain.packed("cuda:0")
if h <= 64:
    pc.tune(tc_mode=int(os.environ.get("DFLOW_TC_MODE", "1")))
if "DFLOW_TC_TS" in os.environ:
    pc.tune(tc_ts=int(os.environ["DFLOW_TC_TS"]))
if "DFLOW_DW_GROUPS" in os.environ:
    pc.tune(tc_dw_groups=int(os.environ["DFLOW_DW_GROUPS"]))
if "DFLOW_DW_TS" in os.environ:
    pc.tune(tc_dw_ts=int(os.environ["DFLOW_DW_TS"]))
g = torch.Generator(device="cuda").manual_seed(0)
x = df.jl_empty((d, B), "cuda:0"); x.normal_(generator=g)
th = df.jl_empty((n, B), "cuda:0"); th.uniform_(0, 1, generator=g)
grad = torch.zeros(pc.P, device="cuda:0"); l2 = torch.zeros(2, device="cuda:0")
pc.loss_grad(x, th, grad, l2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for _ in range(nrep):
    pc.loss_grad(x, th, grad, l2)
e1.record()
torch.cuda.synchronize()
print("ms per loss_grad", e0.elapsed_time(e1) / nrep)
print("ok", name, l2.tolist())
