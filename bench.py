#!/usr/bin/env python
"""bench.py -- samples/sec of the fused coupling-chain hot path on B200 (contract in the task description).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1], "C2"): README chain d=5, n=2, three RNVP coupling layers (masks [1,2,3],
[3,4,5],[5,1,2], conditioners 4->16->16->3) + NormalizationLayer(x,-1,1), Float32, synthetic inputs.
A step is one log-density pass over B = 1e8 samples per GPU (x 2.0 GB + θ 0.8 GB read, 0.4 GB written:
inputs >> the 126 MB L2, so no L2 flush is needed).  `value` = samples/s with inputs resident in HBM;
`e2e` = the same through the host-buffer C-ABI call (pinned host -> device -> host, copies inside the timed
region).  Also reported in `ops`: sample() (in-kernel Philox, fixed θ) and the train step (adjoint + gradient
all-reduce + Adam) on C2 and on C3 (d=16, n=4, 8 layers, hidden 64, global batch 4 Mi sharded over the ranks; on the
CUDA-core kernels and on the tensor-core kernels), and the wide configs on the tcgen05 path: C4 (hidden 256)
log-density + train step, C5 (hidden 512) sampling.

Under torchrun (N > 1) every rank owns one GPU and an equal shard (weak scaling for logpdf / sample: no
communication; the train step all-reduces the packed gradient over NCCL).  Time = max over ranks of the
CUDA-event time of the K steps, bracketed by barrier + synchronize.

`--impl reference` times the CPU restatement of the reference's Flux path (oracle/torch_ref.py; Julia cannot
run here) on this box's host cores over a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "samples/sec: log-density fwd+bwd train step and sample() at 1/2/4/8 B200"
D, N_COND = 5, 2
BYTES_PER_SAMPLE_LOGPDF = 4 * (D + N_COND) + 4  # 32 B
FLOP_PER_SAMPLE_FWD = 4416  # 2*MAC of the conditioners
FP32_FMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # derived, 74.4


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def tensor_peak_tf32():
    """Dense TF32 peak in TFLOP/s: half of the measured cuBLAS bf16 burst figure (same tensor pipe, half rate);
    falls back to half of the recipe's 1590 TFLOP/s."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["bf16_tflops"]) / 2.0
    except Exception:
        return 1590.0 / 2.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for nm, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        busy = [v for v in sm if v > 0.5 * mx] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def readme_oracle_chain(n=N_COND):
    from oracle import dflow_oracle as O

    xs, _ = O.synthetic_data(D, n, 65536, seed=1234)
    return O.readme_chain(n, xs), xs


def c3_oracle_chain():
    from oracle import dflow_oracle as O

    xs, _ = O.synthetic_data(16, 4, 65536, seed=1234)
    return O.block_chain(16, 4, 8, 64, xs), xs


def device_inputs(d, n, B, dev, seed):
    """x_k = 0.1k + (1+0.05k) N(0,1), θ ~ U(-1,2) generated on the device (SURVEY.md §8d shapes)."""
    import densityflows.jl_b200 as df

    g = torch.Generator(device=dev).manual_seed(seed)
    x = df.jl_empty((d, B), dev)
    x.normal_(generator=g)
    k = torch.arange(d, device=dev, dtype=torch.float32).reshape(d, 1)
    x.mul_(1.0 + 0.05 * k).add_(0.1 * k)
    th = df.jl_empty((n, B), dev)
    th.uniform_(-1.0, 2.0, generator=g)
    return x, th


def timed(fn, steps, warmup, dist):
    for _ in range(warmup):
        fn()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    return ms / steps


def run_ours(args):
    import densityflows.jl_b200 as df
    from densityflows.jl_b200.flows import PeerTrainStep, make_train_step
    from tests.helpers import chain_from_oracle

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist_mod

        dist_mod.init_process_group("nccl", device_id=dev)
        dist = dist_mod
    B = int(args.batch)
    lib = df._lib.lib()
    st = torch.cuda.current_stream().cuda_stream

    ochain, xs = readme_oracle_chain()
    chain = chain_from_oracle(ochain)
    x, th = device_inputs(D, N_COND, B, dev, 1234 + rank)
    tmin, tmax = df.minmax_rows(th)
    pc = df.PackedChain(chain._leaves(), dev, tmin, tmax)
    flags = df._lib.THETA_NORMALIZE
    out = torch.empty(B, device=dev)
    xp, tp = df.arrays.flat_view(x).data_ptr(), df.arrays.flat_view(th).data_ptr()

    def step_logpdf():
        df._lib.check(lib.dflow_logpdf(pc.handle, pc.W.data_ptr(), xp, tp, B, None, flags, out.data_ptr(), st))

    l0 = pc.launch_count()
    with ClockSampler(local) as cs:
        ms = timed(step_logpdf, args.steps, args.warmup, dist)
    launches = (pc.launch_count() - l0) * args.steps // (args.steps + args.warmup)
    value = world * B / (ms * 1e-3)
    assert torch.isfinite(out[:: max(1, B // 4096)]).all()

    ops = {}

    def coll(step):
        if world == 1:
            return "none (1 GPU)"
        return ("fused NVLink peer all-reduce + Adam kernel (dflow_dp.cu)" if isinstance(step, PeerTrainStep)
                else "NCCL all-reduce + Adam kernel")

    # ---- sample(): in-kernel Philox base draw + sampling direction, fixed θ (src/Flows.jl:174-185) ----
    thc = torch.tensor([0.5, 0.5], device=dev)

    xo = df.jl_empty((D, B), dev)
    xo_ptr = df.arrays.flat_view(xo).data_ptr()

    def step_sample():
        df._lib.check(lib.dflow_sample_rng(pc.handle, pc.W.data_ptr(), 777, 0, rank * B, None, thc.data_ptr(), B, flags,
                                           xo_ptr, st))

    ms_s = timed(step_sample, max(2, args.steps // 2), 3, dist)
    ops["sample_rng_c2"] = {"samples_per_s": world * B / (ms_s * 1e-3), "ms_per_step": ms_s, "B_per_gpu": B,
                            "hbm_frac": 4 * D * B / (ms_s * 1e-3) / 1e9 / peaks()[0]}
    del xo

    # ---- train step on C2 (per-GPU shard of a 2^24-sample minibatch, weak) ----
    Bt = 1 << 24
    state = df.setup(df.Adam(1e-3), chain)
    ts = make_train_step(pc, state)
    w_save = pc.W.clone()

    def step_train():
        ts(x[:, :Bt], th[:, :Bt], None, Bt * world, flags)

    ms_t = timed(step_train, 3, 3, dist)
    ops["train_step_c2"] = {"samples_per_s": world * Bt / (ms_t * 1e-3), "ms_per_step": ms_t, "B_per_gpu": Bt,
                            "scaling": "weak", "allreduce_bytes": 4 * (pc.P + 2), "collective": coll(ts)}
    pc.W.copy_(w_save)

    # ---- C1: the README example's own training configuration (1e5 samples, 90 % training split, batchsize 64: the
    # reference default, src/Flows.jl:380): one epoch = 1407 launch-bound minibatch steps enqueued from C ----
    if world == 1:
        n_tr, bs = 90000, 64
        order = torch.randperm(n_tr, generator=torch.Generator().manual_seed(5)).to(torch.int32).to(dev)
        m1 = torch.zeros(pc.P, device=dev)
        v1 = torch.zeros(pc.P, device=dev)
        tt = [0]

        def step_epoch():
            tt[0] = pc.train_epoch(x[:, :100000], th[:, :100000], order, bs, m1, v1, tt[0], 1e-3, (0.9, 0.999), 1e-8, flags)

        ms_e = timed(step_epoch, 3, 1, dist)
        nsteps = (n_tr + bs - 1) // bs
        ops["train_epoch_c1"] = {"samples_per_s": n_tr / (ms_e * 1e-3), "ms_per_epoch": ms_e, "minibatch_steps": nsteps,
                                 "us_per_step": ms_e * 1e3 / nsteps, "batchsize": bs,
                                 "api": "dflow_train_epoch (loss+gradient kernel and Adam kernel per minibatch, no host round trip)"}
        pc.W.copy_(w_save)

    # ---- e2e: host buffers through the C-ABI host entry point (pinned host -> HBM -> host every step) ----
    e2e = None
    if not args.no_e2e:
        Be = B
        xh = torch.empty(Be * D, dtype=torch.float32).pin_memory()
        thh = torch.empty(Be * N_COND, dtype=torch.float32).pin_memory()
        oh = torch.empty(Be, dtype=torch.float32).pin_memory()
        xh.copy_(df.arrays.flat_view(x)[: Be * D])
        thh.copy_(df.arrays.flat_view(th)[: Be * N_COND])
        torch.cuda.synchronize()

        def step_host():
            df._lib.check(lib.dflow_logpdf_host(pc.handle, pc.W.data_ptr(), xh.data_ptr(), thh.data_ptr(), Be, flags,
                                                oh.data_ptr(), 1 << 23))

        for _ in range(2):
            step_host()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        ke = max(2, args.steps // 4)
        for _ in range(ke):
            step_host()  # synchronous: returns when the results are in host memory
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert torch.allclose(oh[:4096], out[:4096].cpu(), rtol=1e-6, atol=1e-6)
        e2e = {"value": world * Be * ke / dt, "unit": "samples/s", "h2d_bytes_per_step": 4 * (D + N_COND) * Be,
               "d2h_bytes_per_step": 4 * Be, "steps": ke, "api": "dflow_logpdf_host (chunked 2-stream pipeline)"}
        del xh, thh, oh

    # ---- train step on C3 (global batch 4 Mi sharded over the ranks, strong) ----
    del x, th, out
    torch.cuda.empty_cache()
    if not args.no_c3:
        oc3, _ = c3_oracle_chain()
        c3 = chain_from_oracle(oc3)
        Bg = 1 << 22
        Bl = Bg // world
        x3, th3 = device_inputs(16, 4, Bl, dev, 99 + rank)
        t3min, t3max = np.full(4, -1.0, np.float32), np.full(4, 2.0, np.float32)
        pc3 = df.PackedChain(c3._leaves(), dev, t3min, t3max)
        ts3 = make_train_step(pc3, df.setup(df.Adam(1e-3), c3))

        def step_c3():
            ts3(x3, th3, None, Bg, flags)

        ms3 = timed(step_c3, 2, 1, dist)
        out3 = torch.empty(Bl, device=dev)
        x3p, t3p = df.arrays.flat_view(x3).data_ptr(), df.arrays.flat_view(th3).data_ptr()

        def step_c3_logpdf():
            df._lib.check(lib.dflow_logpdf(pc3.handle, pc3.W.data_ptr(), x3p, t3p, Bl, None, flags, out3.data_ptr(), st))

        ms3l = timed(step_c3_logpdf, 3, 2, dist)
        ops["logpdf_c3"] = {"samples_per_s": world * Bl / (ms3l * 1e-3), "ms_per_step": ms3l, "B_per_gpu": Bl,
                            "path": "tcgen05 3xTF32 (automatic for hidden 64, B >= 131072)",
                            "fp32_equiv_tflops_per_gpu": 172032.0 * Bl / (ms3l * 1e-3) / 1e12}
        # default routing: at hidden 64 and a batch this large the adjoint runs on the tensor cores (dflow_tc.cu)
        ops["train_step_c3"] = {"samples_per_s": Bg / (ms3 * 1e-3), "ms_per_step": ms3, "global_batch": Bg,
                                "scaling": "strong", "allreduce_bytes": 4 * (pc3.P + 2), "collective": coll(ts3),
                                "path": "tcgen05 3xTF32 adjoint (automatic for hidden 64, B >= 32768)"}
        # the same step forced onto the CUDA-core adjoint kernel (tc_mode=-1)
        pc3.tune(tc_mode=-1)
        ts3b = make_train_step(pc3, df.setup(df.Adam(1e-3), c3))
        ms3b = timed(lambda: ts3b(x3, th3, None, Bg, flags), 2, 1, dist)
        ops["train_step_c3_cuda_cores"] = {"samples_per_s": Bg / (ms3b * 1e-3), "ms_per_step": ms3b, "global_batch": Bg,
                                           "scaling": "strong", "path": "FP32 FFMA adjoint kernel (tc_mode=-1)"}
        del x3, th3, pc3, ts3, ts3b
        torch.cuda.empty_cache()

    # ---- wide conditioners on the tensor cores: C4 (h=256) log-density + train step, C5 (h=512) sampling ----
    if not args.no_wide:
        from oracle import dflow_oracle as O

        tf32_peak = tensor_peak_tf32()

        def wide_chain(d_, n_, L_, h_):
            xs_, _ = O.synthetic_data(d_, n_, 8192, seed=1234)
            return chain_from_oracle(O.block_chain(d_, n_, L_, h_, xs_))

        # C4: d=32, n=8, 12 coupling layers, hidden 256 (+ NormalizationLayer); 256 Ki samples per GPU and step
        c4 = wide_chain(32, 8, 12, 256)
        B4 = 1 << 18
        x4, th4 = device_inputs(32, 8, B4, dev, 7 + rank)
        pc4 = df.PackedChain(c4._leaves(), dev, np.full(8, -1.0, np.float32), np.full(8, 2.0, np.float32))
        out4 = torch.empty(B4, device=dev)
        x4p, t4p = df.arrays.flat_view(x4).data_ptr(), df.arrays.flat_view(th4).data_ptr()

        def step_c4_logpdf():
            df._lib.check(lib.dflow_logpdf(pc4.handle, pc4.W.data_ptr(), x4p, t4p, B4, None, flags, out4.data_ptr(), st))

        ms4 = timed(step_c4_logpdf, 3, 2, dist)
        f4 = 3637248.0  # 2*MAC per sample, forward
        ops["logpdf_c4"] = {"samples_per_s": world * B4 / (ms4 * 1e-3), "ms_per_step": ms4, "B_per_gpu": B4,
                            "fp32_equiv_tflops_per_gpu": f4 * B4 / (ms4 * 1e-3) / 1e12,
                            "tensor_tflops_3xtf32_per_gpu": 3 * f4 * B4 / (ms4 * 1e-3) / 1e12,
                            "tensor_frac_of_measured_tf32": 3 * f4 * B4 / (ms4 * 1e-3) / 1e12 / tf32_peak,
                            "tf32_peak_tflops": tf32_peak, "tf32_peak_source": "MEASURED_PEAKS.json bf16_tflops / 2"}
        ts4 = make_train_step(pc4, df.setup(df.Adam(1e-3), c4))

        def step_c4_train():
            ts4(x4, th4, None, B4 * world, flags)

        ms4t = timed(step_c4_train, 2, 1, dist)
        ops["train_step_c4"] = {"samples_per_s": world * B4 / (ms4t * 1e-3), "ms_per_step": ms4t, "B_per_gpu": B4,
                                "scaling": "weak", "allreduce_bytes": 4 * (pc4.P + 2), "collective": coll(ts4),
                                "tensor_tflops_3xtf32_per_gpu": 9 * f4 * B4 / (ms4t * 1e-3) / 1e12,
                                "tensor_frac_of_measured_tf32": 9 * f4 * B4 / (ms4t * 1e-3) / 1e12 / tf32_peak}
        del x4, th4, pc4, ts4, out4
        torch.cuda.empty_cache()
        # C5: d=64, n=16, 16 coupling layers, hidden 512; inverse sampling with a fixed condition, no communication
        c5 = wide_chain(64, 16, 16, 512)
        B5 = 1 << 18
        pc5 = df.PackedChain(c5._leaves(), dev, np.full(16, -1.0, np.float32), np.full(16, 2.0, np.float32))
        th5 = torch.full((16,), 0.5, device=dev)
        x5 = df.jl_empty((64, B5), dev)
        x5p = df.arrays.flat_view(x5).data_ptr()

        def step_c5_sample():
            df._lib.check(lib.dflow_sample_rng(pc5.handle, pc5.W.data_ptr(), 777, 0, rank * B5, None, th5.data_ptr(), B5,
                                               flags, x5p, st))

        ms5 = timed(step_c5_sample, 3, 2, dist)
        f5 = 19398656.0
        ops["sample_rng_c5"] = {"samples_per_s": world * B5 / (ms5 * 1e-3), "ms_per_step": ms5, "B_per_gpu": B5,
                                "seconds_for_1e9_samples": 1e9 / (world * B5 / (ms5 * 1e-3)),
                                "tensor_tflops_3xtf32_per_gpu": 3 * f5 * B5 / (ms5 * 1e-3) / 1e12,
                                "tensor_frac_of_measured_tf32": 3 * f5 * B5 / (ms5 * 1e-3) / 1e12 / tf32_peak}
        assert torch.isfinite(df.arrays.flat_view(x5)[:: max(1, 64 * B5 // 4096)]).all()
        del x5, pc5
        torch.cuda.empty_cache()

    hbm, how = peaks()
    achieved = BYTES_PER_SAMPLE_LOGPDF * B / (ms * 1e-3) / 1e9
    traffic = None
    tp_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp_file):
        try:
            # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture, scaled per launch
            traffic = float(json.load(open(tp_file))["chain_fwd_kernel_bytes_per_sample"]) * B
        except Exception:
            traffic = None
    tfl = FLOP_PER_SAMPLE_FWD * B / (ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2 logpdf: d=5 n=2 L=3 RNVP h=16 + NormalizationLayer, B=%d per GPU" % B,
                   "l2": "inputs (%.1f GB per step) larger than L2; no flush" % (BYTES_PER_SAMPLE_LOGPDF * B / 1e9),
                   "parallelism": "sample-sharded x%d, no data-path collective" % world},
        "clocks": cs.summary(),
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                     "traffic": traffic, "peak_source": how,
                     "note": "binding pipe is FP32 FMA (138 flop/B >> ridge 11): see fma_* keys",
                     "fma_achieved_tflops": tfl, "fma_peak_tflops": FP32_FMA_PEAK_TFLOPS,
                     "fma_frac": tfl / FP32_FMA_PEAK_TFLOPS},
        "ops": ops,
    }
    if rank == 0 and world == 1 and not args.no_cpu:  # reported on rank 0 at N=1 only (bounded sample)
        line["cpu_baseline"] = cpu_baseline(sample_s=12.0)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def cpu_logpdf_rate(nthreads: int, budget_s: float):
    """Flux-equivalent CPU restatement (oracle/torch_ref.py), logpdf on C2-shaped synthetic data."""
    from oracle import dflow_oracle as O
    from oracle import torch_ref as T

    torch.set_num_threads(nthreads)
    ochain, _ = readme_oracle_chain()
    tc = T.TorchChain(ochain, torch.float32)
    Bc = 1 << 18
    x, th = O.synthetic_data(D, N_COND, Bc, seed=5)
    thn = O.normalize_input(th, th.min(axis=1), th.max(axis=1))
    xt, tt = torch.from_numpy(x), torch.from_numpy(thn)
    with torch.no_grad():
        tc.logpdf(xt, tt)  # warm-up
        t0 = time.perf_counter()
        n = 0
        while True:
            tc.logpdf(xt, tt)
            n += 1
            dt = time.perf_counter() - t0
            if dt > budget_s or n >= 200:
                break
    return n * Bc / dt, n * Bc


def cpu_baseline(sample_s: float):
    cores = os.cpu_count() or 1
    rate, nsamp = cpu_logpdf_rate(cores, sample_s)
    return {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": "C2 logpdf, %d samples in batches of 2^18 (torch-CPU restatement of the Flux path, "
                      "oracle/torch_ref.py; Julia/Flux is not installable here)" % nsamp}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from oracle import dflow_oracle as O
    from oracle import torch_ref as T

    ochain, _ = readme_oracle_chain()
    tc = T.TorchChain(ochain, torch.float32)
    Bc = 1 << 20  # bounded sample of the C2 workload per step
    x, th = O.synthetic_data(D, N_COND, Bc, seed=5)
    thn = O.normalize_input(th, th.min(axis=1), th.max(axis=1))
    xt, tt = torch.from_numpy(x), torch.from_numpy(thn)
    with torch.no_grad():
        for _ in range(args.warmup):
            tc.logpdf(xt, tt)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            tc.logpdf(xt, tt)
        dt = time.perf_counter() - t0
    v = Bc * args.steps / dt
    sample = "C2 logpdf, %d samples per step (bounded sample of the 1e8-sample workload)" % Bc
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": int(args.gpus),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2 logpdf: d=5 n=2 L=3 RNVP h=16 + NormalizationLayer (CPU restatement of the "
                               "reference's Flux path; the Julia reference cannot be installed: no julia binary, "
                               "no network)"},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=float, default=1e8, help="samples per GPU per step")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-c3", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-wide", action="store_true", help="skip the tensor-core configs C4 / C5")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU baseline")
        run_ours(args)


if __name__ == "__main__":
    main()
