"""Hidden 32 / 64 chains: CUDA-core kernels vs tensor-core kernels (log-density and train step) across batch sizes --
the measurements behind the automatic routing thresholds in dflow_internal.h (use_tc_fwd / use_tc_grad)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from tests.helpers import chain_from_oracle


def t(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


for (d, n, L, h) in ((16, 4, 8, 64), (10, 3, 4, 32)):
    xs, _ = O.synthetic_data(d, n, 4096, seed=1)
    chain = chain_from_oracle(O.block_chain(d, n, L, h, xs, s_out_scale=0.5))
    pc = chain.packed("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(0)
    for B in (8192, 16384, 32768, 65536, 131072, 262144, 1048576):
        x = df.jl_empty((d, B), "cuda:0"); x.normal_(generator=g)
        th = df.jl_empty((n, B), "cuda:0"); th.uniform_(0, 1, generator=g)
        grad = torch.zeros(pc.P, device="cuda:0"); l2 = torch.zeros(2, device="cuda:0")
        r = {"h": h, "B": B}
        for mode, name in ((-1, "cuda"), (1, "tc")):
            pc.tune(tc_mode=mode)
            r[f"logpdf_{name}_ms"] = round(t(lambda: pc.logpdf(x, th)), 4)
            r[f"grad_{name}_ms"] = round(t(lambda: pc.loss_grad(x, th, grad, l2)), 4)
        print(json.dumps(r), flush=True)
