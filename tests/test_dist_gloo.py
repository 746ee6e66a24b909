"""world_size-2 gloo test of the data-parallel host logic (CPU): per-rank shards of a minibatch, the 1/B_global
seed, and the all-reduce(sum) of the packed [grad | Σlogp | #non-finite] buffer reproduce the single-process step.
The per-shard gradients come from the oracle (the CUDA kernels need a GPU; their DP behaviour is covered by
tests/test_gpu_parity.py::test_grad_idx_and_dp_seed)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dflow_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import densityflows.jl_b200 as df
    from densityflows.jl_b200.flows import allreduce_sum_, shard_range

    x, th = O.synthetic_data(5, 2, 101, seed=3)
    chain = O.readme_chain(2, x)
    P = O.pack_params(chain).size
    batch = np.random.default_rng(0).permutation(101)[:77]  # one (odd-sized) minibatch of the epoch's order
    lo, hi = shard_range(len(batch), rank, world)
    mine = batch[lo:hi]
    buf = torch.zeros(P + 2, dtype=torch.float64)
    if len(mine):
        loss, g, z, ldj = O.chain_loss_and_grad(chain, x[:, mine], th[:, mine], np.float64, inv_btot=1.0 / len(batch))
        buf[:P] = torch.from_numpy(g)
        buf[P] = float(np.sum(O.mvnormal_logpdf(z, np.float64) + ldj))
    allreduce_sum_(buf)
    # Adam on every replica from the identical reduced gradient -> identical weights without a broadcast
    w, m, v = O.adam_step(O.pack_params(chain), buf[:P].numpy().astype(np.float32), np.zeros(P, np.float32),
                          np.zeros(P, np.float32), 1)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), buf=buf.numpy(), w=w, lo=lo, hi=hi)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert (int(r0["lo"]), int(r0["hi"]), int(r1["lo"]), int(r1["hi"])) == (0, 39, 39, 77)
    np.testing.assert_array_equal(r0["buf"], r1["buf"])
    np.testing.assert_array_equal(r0["w"], r1["w"])
    x, th = O.synthetic_data(5, 2, 101, seed=3)
    chain = O.readme_chain(2, x)
    batch = np.random.default_rng(0).permutation(101)[:77]
    loss, g, z, ldj = O.chain_loss_and_grad(chain, x[:, batch], th[:, batch], np.float64)
    P = g.size
    np.testing.assert_allclose(r0["buf"][:P], g, rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(-r0["buf"][P] / 77, loss, rtol=1e-12)


def test_shard_range_covers_everything():
    from densityflows.jl_b200.flows import shard_range

    for n in (0, 1, 7, 64, 65, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
