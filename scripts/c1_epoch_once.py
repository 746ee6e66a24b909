import sys, os
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
order = torch.randperm(6400, generator=torch.Generator().manual_seed(5)).to(torch.int32).to("cuda:0")
m = torch.zeros(pc.P, device="cuda:0"); v = torch.zeros(pc.P, device="cuda:0")
t = pc.train_epoch(x, th, order, 64, m, v, 0)
torch.cuda.synchronize()
print("done", t)
