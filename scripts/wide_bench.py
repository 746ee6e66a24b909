"""Device timing of the wide (tcgen05) path on the C4 / C5 shapes."""
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

CFG = {"c4": (32, 8, 12, 256, 1 << 18, 3637248), "c5": (64, 16, 16, 512, 1 << 17, 19398656), "c3w": (16, 4, 8, 128, 1 << 19, 0),
       "c3": (16, 4, 8, 64, 1 << 21, 172032)}
GEN = int(os.environ.get("DFLOW_WIDE_GEN", "2"))
TCMODE = int(os.environ.get("DFLOW_TC_MODE", "1"))
TRAIN = int(os.environ.get("DFLOW_TRAIN", "0"))
for name in (sys.argv[1:] or ["c4", "c5"]):
    d, n, L, h, B, flops = CFG[name]
    xs, _ = O.synthetic_data(d, n, 4096, seed=1)
    chain = chain_from_oracle(O.block_chain(d, n, L, h, xs))
    pc = chain.packed("cuda:0")
    pc.tune(wide_gen=GEN)
    pc.tune(tc_debug=int(os.environ.get("DFLOW_TC_DEBUG", "0")))
    pc.tune(tc_ns_max=int(os.environ.get("DFLOW_TC_NS_MAX", "0")))
    pc.tune(tc_cluster=int(os.environ.get("DFLOW_TC_CLUSTER", "0")))
    if h <= 64:
        pc.tune(tc_mode=TCMODE)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = df.jl_empty((d, B), "cuda:0"); x.normal_(generator=g)
    th = df.jl_empty((n, B), "cuda:0"); th.uniform_(0, 1, generator=g)
    ms = timeit(lambda: pc.logpdf(x, th))
    thc = torch.zeros(n, device="cuda:0")
    out = df.jl_empty((d, B), "cuda:0")
    ms2 = timeit(lambda: pc.sample_rng(B, 7, None, thc, out=out))
    fl = flops or 2 * 2 * L * ((n + d // 2) * h + h * h + h * (d // 2))
    tr = {}
    if TRAIN and h <= 256:
        grad = torch.zeros(pc.P, device="cuda:0"); l2 = torch.zeros(2, device="cuda:0")
        ms3 = timeit(lambda: pc.loss_grad(x, th, grad, l2), iters=2)
        tr = {"grad_ms": ms3, "grad_samples_per_s": B / ms3 * 1e3, "grad_tensor_tflops_3x": 9 * fl * B / ms3 * 1e3 / 1e12}
    print(json.dumps({"gen": GEN, "launches": pc.launch_count(), **tr, "cfg": name, "B": B, "P": pc.P, "logpdf_ms": ms, "logpdf_samples_per_s": B / ms * 1e3,
                      "sample_ms": ms2, "sample_samples_per_s": B / ms2 * 1e3,
                      "fp32_equiv_tflops": fl * B / ms * 1e3 / 1e12, "tensor_tflops_3x": 3 * fl * B / ms * 1e3 / 1e12}), flush=True)
