"""C2 train-step timing of the shipped narrow adjoint (chain_grad2_kernel) across samples-per-thread settings.
(The register-resident v3 variant this script once compared against was measured and removed: profiles/r02_narrow_adjoint_v3.md.)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import densityflows.jl_b200 as df
from bench import packed, device_inputs
dev = torch.device("cuda:0")
B = 1 << 24
x, th = device_inputs(5, 2, B, dev, 1)
tmin, tmax = df.minmax_rows(th)
chain, pc = packed("c2", dev, tmin, tmax)
grads = {}
for tune in [dict(grad_spt=0), dict(grad_spt=1), dict(grad_spt=2)]:
    pc.tune(**tune)
    grad = torch.zeros(pc.P, device=dev); l2 = torch.zeros(2, device=dev)
    pc.loss_grad(x, th, grad, l2, None, 1)
    grads[str(tune)] = (grad.clone(), l2.clone())
    ts = []
    for _ in range(5):
        grad.zero_(); l2.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pc.loss_grad(x, th, grad, l2, None, 1); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    med = float(np.median(ts))
    print(json.dumps({"tune": tune, "ms": med, "sps": B / med * 1e3}), flush=True)
ks = list(grads)
a, b = grads[ks[0]], grads[ks[-1]]
print("max |dg| / max|g| between the first and the last setting =", float((a[0] - b[0]).abs().max() / b[0].abs().max()), "loss", float(a[1][0]), float(b[1][0]))
