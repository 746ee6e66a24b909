"""2+ rank check of the fused peer all-reduce + Adam step against the NCCL path (run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import densityflows.jl_b200 as df
from densityflows.jl_b200.flows import TrainStep, PeerTrainStep
from oracle import dflow_oracle as O
from tests.helpers import chain_from_oracle

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

def run(name, mk, d, n, B):
    xs, _ = O.synthetic_data(d, n, 4096, seed=1)
    ochain = mk(xs)
    x, th = O.synthetic_data(d, n, B * world, seed=5)
    lo, hi = rank * B, (rank + 1) * B
    xl, tl = x[:, lo:hi].copy(), th[:, lo:hi].copy()
    res = {}
    for kind in ("nccl", "peer"):
        chain = chain_from_oracle(ochain)
        pc = chain.packed(dev)
        state = df.setup(df.Adam(1e-3), chain)
        step = TrainStep(pc, state) if kind == "nccl" else PeerTrainStep(pc, state)
        xd, td = df.to_jl(xl, dev), df.to_jl(tl, dev)
        for _ in range(3):
            step(xd, td, None, B * world, 0)
        torch.cuda.synchronize()
        if kind == "peer":
            step.check()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            step(xd, td, None, B * world, 0)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        W = pc.W.clone()
        allW = [torch.empty_like(W) for _ in range(world)]
        dist.all_gather(allW, W)
        same = all(torch.equal(allW[0], w) for w in allW)
        res[kind] = (W.cpu().numpy(), dt, same, float(step.loss2[0].item()))
    err = np.abs(res["nccl"][0] - res["peer"][0]).max()
    if rank == 0:
        print(f"{name}: P={res['peer'][0].size} |W_nccl - W_peer| max {err:.3e}  replicas identical: nccl {res['nccl'][2]} peer {res['peer'][2]}  "
              f"step ms: nccl {res['nccl'][1]*1e3:.3f} peer {res['peer'][1]*1e3:.3f}  loss-sum nccl {res['nccl'][3]:.4f} peer {res['peer'][3]:.4f}", flush=True)

run("readme_c2", lambda xs: O.readme_chain(2, xs), 5, 2, 1 << 16)
run("c3_h64", lambda xs: O.block_chain(16, 4, 8, 64, xs), 16, 4, 1 << 16)
run("c4_h256", lambda xs: O.block_chain(32, 8, 12, 256, xs), 32, 8, 1 << 14)
dist.barrier()
dist.destroy_process_group()
