"""Soak test of the tensor-core train step: repeated loss + gradient calls on C3- and C4-like chains at ragged batch sizes;
every repetition must reproduce the first one's loss exactly to summation order and its gradient to 1e-5 of the max-norm
(catches rare races / stale barrier phases that a single parity run can miss)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import densityflows.jl_b200 as df
from oracle import dflow_oracle as O
from tests.helpers import chain_from_oracle

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for (d, n, L, h, B) in ((16, 4, 8, 64, 148 * 128 * 3 + 77), (32, 8, 4, 256, 148 * 128 + 333), (10, 3, 4, 32, 70000 + 5)):
    xs, _ = O.synthetic_data(d, n, 4096, seed=1)
    chain = chain_from_oracle(O.block_chain(d, n, L, h, xs, s_out_scale=0.3))
    pc = chain.packed("cuda:0")
    pc.tune(tc_mode=1)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = df.jl_empty((d, B), "cuda:0"); x.normal_(generator=g)
    th = df.jl_empty((n, B), "cuda:0"); th.uniform_(0, 1, generator=g)
    ref = None
    worst = 0.0
    for r in range(reps):
        grad = torch.zeros(pc.P, device="cuda:0"); l2 = torch.zeros(2, device="cuda:0")
        pc.loss_grad(x, th, grad, l2)
        lp = pc.logpdf(x, th)
        if ref is None:
            ref = (grad.clone(), l2.clone(), lp.clone())
            continue
        assert torch.isfinite(grad).all() and float(l2[1]) == 0.0, (h, r)
        dg = float((grad - ref[0]).abs().max() / ref[0].abs().max())
        worst = max(worst, dg)
        assert dg <= 1e-5, (h, r, dg)
        assert abs(float(l2[0] - ref[1][0])) <= 2e-6 * abs(float(ref[1][0])), (h, r)
        assert torch.equal(lp, ref[2]), (h, r, "log-density not bit-reproducible")
    print(json.dumps({"hidden": h, "B": B, "reps": reps, "worst_grad_rel_diff": worst}), flush=True)
print("soak OK")
