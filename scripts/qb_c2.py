"""C2 forward timing: constant-bank kernel vs shared-memory-column kernel (CUDA events)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from tests.helpers import chain_from_oracle
from scripts.quick_bench import timeit

d, n, B = 5, 2, 1 << 25
xs, ths = O.synthetic_data(d, n, 4096, seed=1)
chain = chain_from_oracle(O.readme_chain(2, xs))
pc = chain.packed("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
x = df.jl_empty((d, B), "cuda:0"); x.normal_(generator=g)
th = df.jl_empty((n, B), "cuda:0"); th.uniform_(0, 1, generator=g)
out = torch.empty(B, device="cuda:0")
lib = df._lib.lib(); st = torch.cuda.current_stream().cuda_stream
xp, tp = df.arrays.flat_view(x).data_ptr(), df.arrays.flat_view(th).data_ptr()
thc = torch.zeros(n, device="cuda:0")
res = {}
for tune in [dict(fwd_const=0, ctas_per_sm=0), dict(fwd_const=0, ctas_per_sm=2), dict(fwd_const=0, ctas_per_sm=4), dict(fwd_const=-1, ctas_per_sm=0)]:
    pc.tune(**tune)
    f = lambda: df._lib.check(lib.dflow_logpdf(pc.handle, pc.W.data_ptr(), xp, tp, B, None, 0, out.data_ptr(), st))
    med, mn = timeit(f)
    ref = out.clone()
    res[str(tune)] = ref[:4096].cpu()
    f2 = lambda: df._lib.check(lib.dflow_sample_rng(pc.handle, pc.W.data_ptr(), 1, 0, 0, None, thc.data_ptr(), B, 0, xp, st))
    med2, _ = timeit(f2)
    x.normal_(generator=g)
    print(json.dumps({"tune": tune, "logpdf_ms": med, "logpdf_sps": B / med * 1e3, "sample_ms": med2, "sample_sps": B / med2 * 1e3}), flush=True)
