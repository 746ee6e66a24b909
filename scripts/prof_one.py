"""Runs a few launches of one hot-path op on a small batch; used under `ncu -k regex:chain_ ...` (see profiles/)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from tests.helpers import chain_from_oracle

op = sys.argv[1] if len(sys.argv) > 1 else "logpdf"
cfg = sys.argv[2] if len(sys.argv) > 2 else "c2"
B = int(float(sys.argv[3])) if len(sys.argv) > 3 else 1 << 23
tune = dict(kv.split("=") for kv in sys.argv[4:])
if cfg == "c2":
    d, n, mk = 5, 2, lambda x: O.readme_chain(2, x)
else:
    d, n, mk = 16, 4, lambda x: O.block_chain(16, 4, 8, 64, x)
xs, _ = O.synthetic_data(d, n, 4096, seed=1)
chain = chain_from_oracle(mk(xs))
pc = chain.packed("cuda:0")
pc.tune(**{k: int(v) for k, v in tune.items()})
g = torch.Generator(device="cuda").manual_seed(0)
x = df.jl_empty((d, B), "cuda:0"); x.normal_(generator=g)
th = df.jl_empty((n, B), "cuda:0"); th.uniform_(0, 1, generator=g)
grad = torch.zeros(pc.P, device="cuda:0"); l2 = torch.zeros(2, device="cuda:0")
for _ in range(3):
    if op == "logpdf":
        out = pc.logpdf(x, th)
    elif op == "sample":
        out = pc.sample_rng(B, 7, None, torch.zeros(n, device="cuda:0"))
    else:
        pc.loss_grad(x, th, grad, l2)
torch.cuda.synchronize()
print("done", op, cfg, B, tune)
