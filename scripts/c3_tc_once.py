"""One tensor-core log-density call (and optionally one train step) at C3 — short command line for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from tests.helpers import chain_from_oracle
d, n, L, h, B = 16, 4, 8, 64, 1 << 20
xs, _ = O.synthetic_data(d, n, 4096, seed=1)
chain = chain_from_oracle(O.block_chain(d, n, L, h, xs))
pc = chain.packed("cuda:0")
pc.tune(tc_mode=1)
g = torch.Generator(device="cuda").manual_seed(0)
x = df.jl_empty((d, B), "cuda:0"); x.normal_(generator=g)
th = df.jl_empty((n, B), "cuda:0"); th.uniform_(0, 1, generator=g)
lp = pc.logpdf(x, th)
if len(sys.argv) > 1 and sys.argv[1] == "train":
    grad = torch.zeros(pc.P, device="cuda:0"); l2 = torch.zeros(2, device="cuda:0")
    pc.loss_grad(x, th, grad, l2)
torch.cuda.synchronize()
print("ok", float(lp.sum()))
