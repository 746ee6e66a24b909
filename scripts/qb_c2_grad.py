"""C2 adjoint timing for tuning knobs (CUDA events)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from tests.helpers import chain_from_oracle
from scripts.quick_bench import timeit
d, n, B = 5, 2, 1 << 23
xs, ths = O.synthetic_data(d, n, 4096, seed=1)
chain = chain_from_oracle(O.readme_chain(2, xs))
pc = chain.packed("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
x = df.jl_empty((d, B), "cuda:0"); x.normal_(generator=g)
th = df.jl_empty((n, B), "cuda:0"); th.uniform_(0, 1, generator=g)
grads = {}
for tune in [dict(grad_smem=0), dict(grad_smem=-1)]:
    pc.tune(**tune)
    grad = torch.zeros(pc.P, device="cuda:0"); l2 = torch.zeros(2, device="cuda:0")
    pc.loss_grad(x, th, grad, l2)
    grads[str(tune)] = grad.clone()
    g2 = torch.zeros(pc.P, device="cuda:0")
    med, mn = timeit(lambda: pc.loss_grad(x, th, g2, l2), iters=3, warm=1)
    print(json.dumps({"tune": tune, "ms": med, "sps": B / med * 1e3}), flush=True)
ks = list(grads)
print("rel diff", float((grads[ks[0]] - grads[ks[1]]).abs().max() / grads[ks[0]].abs().max()))
