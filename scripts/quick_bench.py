"""Quick device timing of the chain kernels for tuning sweeps (not the contract bench; see bench.py)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from tests.helpers import chain_from_oracle

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    return float(np.median(ts)), float(np.min(ts))

def run(name, d, n, mk, B, Bg, tunes):
    xs, ths = O.synthetic_data(d, n, 4096, seed=1)
    chain = chain_from_oracle(mk(xs))
    pc = chain.packed("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(0)
    x = df.jl_empty((d, B), "cuda:0"); x.normal_(generator=g)
    th = df.jl_empty((n, B), "cuda:0"); th.uniform_(0, 1, generator=g)
    out = torch.empty(B, device="cuda:0")
    lib = df._lib.lib(); st = torch.cuda.current_stream().cuda_stream
    xp, tp = df.arrays.flat_view(x).data_ptr(), df.arrays.flat_view(th).data_ptr()
    grad = torch.zeros(pc.P, device="cuda:0"); l2 = torch.zeros(2, device="cuda:0")
    for tune in tunes:
        pc.tune(**tune)
        f = lambda: df._lib.check(lib.dflow_logpdf(pc.handle, pc.W.data_ptr(), xp, tp, B, None, 0, out.data_ptr(), st))
        med, mn = timeit(f)
        print(json.dumps({"cfg": name, "op": "logpdf", "B": B, "tune": tune, "ms": med, "ms_min": mn, "samples_per_s": B / med * 1e3}), flush=True)
    gtunes = [dict(grad_spt=s, grad_threads=t, ctas_per_sm=c) for s, t, c in
              ([(2, 0, 0), (2, 0, 1), (2, 0, 2), (2, 128, 2), (1, 0, 0), (1, 0, 2), (4, 0, 1), (4, 0, 2), (-1, 256, 0)]
               if name == "C2" else [(1, 0, 0), (1, 0, 2), (1, 64, 2), (2, 0, 0), (2, 0, 1), (-1, 128, 0)])]
    for tune in gtunes:
        pc.tune(**tune)
        wsb = int(lib.dflow_workspace_bytes(pc.handle, Bg))
        ws = torch.empty(wsb, device="cuda:0", dtype=torch.uint8)
        f = lambda: df._lib.check(lib.dflow_loss_grad(pc.handle, pc.W.data_ptr(), xp, tp, Bg, None, 1.0 / Bg, 0, l2.data_ptr(), grad.data_ptr(), ws.data_ptr(), wsb, st))
        med, mn = timeit(f, iters=3, warm=1)
        print(json.dumps({"cfg": name, "op": "loss_grad", "B": Bg, "tune": tune, "ms": med, "ms_min": mn, "samples_per_s": Bg / med * 1e3}), flush=True)
    thc = torch.zeros(max(n, 1), device="cuda:0")
    pc.tune(fwd_spt=0, fwd_threads=0, grad_threads=0, ctas_per_sm=0)
    f = lambda: df._lib.check(lib.dflow_sample_rng(pc.handle, pc.W.data_ptr(), 1, 0, 0, None, thc.data_ptr(), B, 0, xp, st))
    med, mn = timeit(f)
    print(json.dumps({"cfg": name, "op": "sample_rng", "B": B, "ms": med, "samples_per_s": B / med * 1e3}), flush=True)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "c2"
    if which in ("c2", "all"):
        tunes = [dict(fwd_spt=s, fwd_threads=t, grad_threads=gt, ctas_per_sm=c) for s, t, gt, c in
                 [(4, 0, 0, 0), (4, 0, 0, 2), (4, 0, 0, 3), (4, 64, 0, 4), (2, 0, 0, 0), (2, 0, 0, 2), (2, 128, 0, 4), (-2, 0, 0, 3)]]
        run("C2", 5, 2, lambda x: O.readme_chain(2, x), 1 << 25, 1 << 23, tunes)
    if which in ("c3", "all"):
        tunes = [dict(fwd_spt=1, fwd_threads=t, grad_threads=gt, ctas_per_sm=c) for t, gt, c in [(128, 128, 0), (256, 64, 0), (64, 128, 2)]]
        run("C3", 16, 4, lambda x: O.block_chain(16, 4, 8, 64, x), 1 << 21, 1 << 19, tunes)
