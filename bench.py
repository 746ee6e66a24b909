#!/usr/bin/env python
"""bench.py -- samples/sec of the fused coupling-chain hot path on B200 (contract in the task description).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1], "C2"): README chain d=5, n=2, three RNVP coupling layers (masks [1,2,3],
[3,4,5],[5,1,2], conditioners 4->16->16->3) + NormalizationLayer(x,-1,1), Float32, synthetic inputs.
A step is one log-density pass over B = 1e8 samples per GPU (x 2.0 GB + θ 0.8 GB read, 0.4 GB written:
inputs >> the 126 MB L2, so no L2 flush is needed).  `value` = samples/s with inputs resident in HBM;
`e2e` = the same through the host-buffer C-ABI call (pinned host -> device -> host, copies inside the timed
region).  `ops` reports the other legs of the metric, every one with both roofline fractions (bytes-based against the
measured HBM copy bandwidth, flop-based against the peak of the pipe the kernel actually uses):
  sample() (in-kernel Philox, fixed θ), the train step (adjoint + gradient reduction + Adam) on C2 and on C3 (d=16, n=4,
  8 layers, hidden 64, global batch 4 Mi sharded over the ranks), the C1 epoch (1e5 samples, batchsize 64: the README
  example's own training configuration), and the wide configs on the tcgen05 path: C4 (hidden 256) log-density + train
  step, C5 (hidden 512) sampling.

Chains are built with the product's own constructors (df.CouplingLayer / CouplingBlock / NormalizationLayer / FlowChain);
their parameters are then set to the seeded synthetic weights of SURVEY.md §8d (the oracle's builders produce them) so
that the CPU baseline evaluates the SAME model.

Under torchrun (N > 1) every rank owns one GPU and an equal shard (weak scaling for logpdf / sample: no
communication; the train step reduces the packed gradient over NVLink peer memory fused with Adam).  Time = max over
ranks of the CUDA-event time of the K steps, bracketed by barrier + synchronize.  At N > 1 the run also ASSERTS that the
replicas are bit-identical after the train steps and that the data-parallel gradient equals the single-rank full-batch
gradient, and splits the C3 step into adjoint / peer-barrier wait / reduce + Adam.

`--impl reference` times the CPU restatement of the reference's Flux path (oracle/torch_ref.py; Julia cannot
run here) on this box's host cores over a bounded sample of the same workload, with the same chunking procedure as the
`cpu_baseline` leg of the GPU arm.
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
FP32_FMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # derived, 74.4

# SURVEY.md §8d / BASELINE.md §3: algorithmic 2*MAC flops and compulsory HBM bytes per sample
WORK = {
    "c2": {"fwd_flop": 4416.0, "train_flop": 13248.0, "logpdf_B": 32.0, "train_B": 28.0, "sample_B": 20.0},
    "c3": {"fwd_flop": 172032.0, "train_flop": 516096.0, "logpdf_B": 84.0, "train_B": 80.0, "sample_B": 64.0},
    "c4": {"fwd_flop": 3637248.0, "train_flop": 10911744.0, "logpdf_B": 164.0, "train_B": 160.0, "sample_B": 128.0},
    "c5": {"fwd_flop": 19398656.0, "train_flop": 0.0, "logpdf_B": 324.0, "train_B": 320.0, "sample_B": 256.0},
}


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


def roofline_of(cfg: str, leg: str, samples_per_s_per_gpu: float, pipe: str):
    """Both fractions for one op: bytes-based vs measured HBM, flop-based vs the pipe in use (`fp32`: derived FFMA peak;
    `tf32x3`: three TF32 MMAs per algorithmic product against the measured bf16 / 2)."""
    w = WORK[cfg]
    hbm, how = peaks()
    by = {"logpdf": w["logpdf_B"], "train": w["train_B"], "sample": w["sample_B"]}[leg]
    fl = w["train_flop"] if leg == "train" else w["fwd_flop"]
    gbs = by * samples_per_s_per_gpu / 1e9
    tfl = fl * samples_per_s_per_gpu / 1e12
    if pipe == "fp32":
        peak, exec_tfl = FP32_FMA_PEAK_TFLOPS, tfl
    else:
        peak, exec_tfl = tensor_peak_tf32(), 3.0 * tfl
    return {"hbm_gbs": gbs, "hbm_frac": gbs / hbm, "hbm_peak_source": how, "algorithmic_tflops": tfl, "pipe": pipe,
            "pipe_tflops": exec_tfl, "pipe_peak_tflops": peak, "pipe_frac": exec_tfl / peak,
            "bound": "pipe" if exec_tfl / peak > gbs / hbm else "hbm"}


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


# ---- the benchmarked models --------------------------------------------------------------------------------------
def oracle_chain(cfg: str):
    """Seeded synthetic weights of SURVEY.md §8d (glorot weights, U(-0.1,0.1) biases, last s Dense x0.1 for L >= 8)."""
    from oracle import dflow_oracle as O

    if cfg in ("c1", "c2"):
        xs, _ = O.synthetic_data(D, N_COND, 65536, seed=1234)
        return O.readme_chain(N_COND, xs)
    d, n, L, h, nx = {"c3": (16, 4, 8, 64, 65536), "c4": (32, 8, 12, 256, 8192), "c5": (64, 16, 16, 512, 8192)}[cfg]
    xs, _ = O.synthetic_data(d, n, nx, seed=1234)
    return O.block_chain(d, n, L, h, xs)


def product_chain(cfg: str, ochain):
    """The same model through the product's public constructors (mirror of the reference API); parameters copied from the
    oracle chain's packed vector (identical layout) so that both arms evaluate one model."""
    import densityflows.jl_b200 as df
    from oracle import dflow_oracle as O

    norm = O.flatten(ochain)[-1]
    if cfg in ("c1", "c2"):
        layers = [df.CouplingLayer(D, m, n=N_COND, hidden_dim_s=16, hidden_dim_t=16) for m in ([1, 2, 3], [3, 4, 5], [5, 1, 2])]
    else:
        d, n, L, h = {"c3": (16, 4, 8, 64), "c4": (32, 8, 12, 256), "c5": (64, 16, 16, 512)}[cfg]
        layers = [df.CouplingBlock(d, d // 2, n=n, hidden_dim_s=h, hidden_dim_t=h) for _ in range(L // 2)]
    layers.append(df.NormalizationLayer(norm.x_min, norm.x_max, norm.alpha, norm.beta))
    return df.FlowChain(*layers)


def packed(cfg: str, dev, theta_min, theta_max):
    import densityflows.jl_b200 as df
    from oracle import dflow_oracle as O

    oc = oracle_chain(cfg)
    chain = product_chain(cfg, oc)
    pc = df.PackedChain(chain._leaves(), dev, theta_min, theta_max)
    w = O.pack_params(oc)
    assert w.size == pc.P, (w.size, pc.P)
    pc.W.copy_(torch.from_numpy(w))
    chain._packed = pc
    return chain, pc


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


def bind_to_gpu_numa_node(local: int):
    """Pin this rank (and therefore its pinned staging buffers, first touch) to the CPU cores of its GPU's NUMA node."""
    info = {"bound": False}
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        bus = bus[-12:] if len(bus) > 12 else bus  # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        info["numa_node"] = node
        if node >= 0:
            cpus = []
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus += list(range(int(lo), int(hi or lo) + 1))
            allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
            if allowed:
                os.sched_setaffinity(0, allowed)
                info.update({"bound": True, "cpus": f"{allowed[0]}-{allowed[-1]} ({len(allowed)})"})
    except Exception as e:  # not fatal: the number is then simply measured unbound
        info["error"] = str(e)[:80]
    return info


def host_copy_ceiling(dev, dist=None, nbytes=1 << 30):
    """What the host can feed this GPU WHILE THE OTHER RANKS DO THE SAME: concurrent pinned H2D + D2H copies of `nbytes` +
    `nbytes / 7` (the e2e call's in : out ratio), every repetition started behind a barrier; median GB/s (both directions
    summed) of this rank."""
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes // 7, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(nbytes // 7, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    rates = []
    for _ in range(5):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        rates.append((nbytes + nbytes // 7) / (time.perf_counter() - t0) / 1e9)
    return float(np.median(rates[1:]))


def run_ours(args):
    import densityflows.jl_b200 as df
    from densityflows.jl_b200.flows import PeerTrainStep, TrainStep, make_train_step

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa_node(local)
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
    hbm, how = peaks()
    checks = {}

    x, th = device_inputs(D, N_COND, B, dev, 1234 + rank)
    tmin, tmax = df.minmax_rows(th)
    chain, pc = packed("c2", dev, tmin, tmax)
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
                            "roofline": roofline_of("c2", "sample", B / (ms_s * 1e-3), "fp32")}
    del xo

    # ---- train step on C2 (per-GPU shard of a 2^24-sample minibatch, weak) ----
    Bt = min(1 << 24, B)
    state = df.setup(df.Adam(1e-3), chain)
    ts = make_train_step(pc, state)
    w_save = pc.W.clone()

    def step_train():
        ts(x[:, :Bt], th[:, :Bt], None, Bt * world, flags)

    ms_t = timed(step_train, 3, 3, dist)
    ops["train_step_c2"] = {"samples_per_s": world * Bt / (ms_t * 1e-3), "ms_per_step": ms_t, "B_per_gpu": Bt,
                            "scaling": "weak", "allreduce_bytes": 4 * (pc.P + 2), "collective": coll(ts),
                            "roofline": roofline_of("c2", "train", Bt / (ms_t * 1e-3), "fp32")}
    if world > 1:
        checks.update(dp_correctness(df, dist, pc, state, ts, x, th, flags, rank, world, dev))
    pc.W.copy_(w_save)

    # ---- C1: the README example's own training configuration (1e5 samples, 90 % training split, batchsize 64: the
    # reference default, src/Flows.jl:380): one epoch = 1407 minibatch steps inside ONE persistent kernel ----
    if world == 1:
        n_tr, bs = 90000, 64
        order = df.device_permutation(n_tr, 5, dev)
        m1 = torch.zeros(pc.P, device=dev)
        v1 = torch.zeros(pc.P, device=dev)
        tt = [0]
        nsteps = (n_tr + bs - 1) // bs
        modes = (("train_epoch_c1", 0),) if args.profile else (("train_epoch_c1", 0), ("train_epoch_c1_per_minibatch_launches", -1))
        for tag, mode in modes:
            pc.tune(epoch_kernel=mode)
            pc.W.copy_(w_save)
            m1.zero_()
            v1.zero_()
            tt[0] = 0

            def step_epoch():
                tt[0] = pc.train_epoch(x[:, :100000], th[:, :100000], order, bs, m1, v1, tt[0], 1e-3, (0.9, 0.999), 1e-8, flags)

            ms_e = timed(step_epoch, 3, 1, dist)
            ops[tag] = {"samples_per_s": n_tr / (ms_e * 1e-3), "ms_per_epoch": ms_e, "minibatch_steps": nsteps,
                        "us_per_step": ms_e * 1e3 / nsteps, "batchsize": bs,
                        "api": ("dflow_train_epoch: persistent one-CTA epoch kernel (csrc/dflow_small.cu), 1 launch per epoch"
                                if mode == 0 else "dflow_train_epoch: loss+gradient kernel and Adam kernel per minibatch")}
        pc.tune(epoch_kernel=0)
        pc.W.copy_(w_save)

    # ---- e2e: host buffers through the C-ABI host entry point (pinned host -> HBM -> host every step) ----
    e2e = None
    if not args.no_e2e:
        Be = B
        ceiling = host_copy_ceiling(dev, dist)  # every rank copies at the same time: the ceiling the ranks see TOGETHER
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
            t = torch.tensor([dt, -ceiling], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, ceiling = float(t[0].item()), -float(t[1].item())
        assert torch.allclose(oh[:4096], out[:4096].cpu(), rtol=1e-6, atol=1e-6)
        e2e = {"value": world * Be * ke / dt, "unit": "samples/s", "h2d_bytes_per_step": 4 * (D + N_COND) * Be,
               "d2h_bytes_per_step": 4 * Be, "steps": ke, "api": "dflow_logpdf_host (chunked 2-stream pipeline)",
               "achieved_host_copy_gbs_per_gpu": 32.0 * Be * ke / dt / 1e9,
               "host_copy_ceiling_gbs_per_gpu": ceiling,
               "host_copy_ceiling_note": "pinned H2D + D2H of 1 GiB + 1/7 GiB on every rank at the same time (barrier before each "
                                         "repetition), median per rank, min over ranks; measured in this run",
               "numa_binding": numa}
        del xh, thh, oh

    # ---- train step on C3 (global batch 4 Mi sharded over the ranks, strong) ----
    del x, th, out
    torch.cuda.empty_cache()
    if not args.no_c3:
        Bg = 1 << 22
        Bl = Bg // world
        x3, th3 = device_inputs(16, 4, Bl, dev, 99 + rank)
        t3min, t3max = np.full(4, -1.0, np.float32), np.full(4, 2.0, np.float32)
        c3, pc3 = packed("c3", dev, t3min, t3max)
        ts3 = make_train_step(pc3, df.setup(df.Adam(1e-3), c3))

        def step_c3():
            ts3(x3, th3, None, Bg, flags)

        ms3 = timed(step_c3, 2, 1, dist)
        out3 = torch.empty(Bl, device=dev)
        x3p, t3p = df.arrays.flat_view(x3).data_ptr(), df.arrays.flat_view(th3).data_ptr()
        pc3.ensure_scratch(Bl)

        def step_c3_logpdf():
            df._lib.check(lib.dflow_logpdf(pc3.handle, pc3.W.data_ptr(), x3p, t3p, Bl, None, flags, out3.data_ptr(), st))

        ms3l = timed(step_c3_logpdf, 3, 2, dist)
        ops["logpdf_c3"] = {"samples_per_s": world * Bl / (ms3l * 1e-3), "ms_per_step": ms3l, "B_per_gpu": Bl,
                            "path": "tcgen05 3xTF32 (automatic for hidden 64, B >= 65536: TMEM-sourced four-chain kernel)",
                            "roofline": roofline_of("c3", "logpdf", Bl / (ms3l * 1e-3), "tf32x3")}
        # default routing: at hidden 64 and a batch this large the adjoint runs on the tensor cores (dflow_tc.cu)
        ops["train_step_c3"] = {"samples_per_s": Bg / (ms3 * 1e-3), "ms_per_step": ms3, "global_batch": Bg,
                                "scaling": "strong", "allreduce_bytes": 4 * (pc3.P + 2), "collective": coll(ts3),
                                "path": "tcgen05 3xTF32 adjoint (automatic for hidden 64, B >= 8192)",
                                "roofline": roofline_of("c3", "train", Bl / (ms3 * 1e-3), "tf32x3")}
        if world > 1 and isinstance(ts3, PeerTrainStep):
            ops["train_step_c3"]["split"] = dp_split(df, dist, lib, pc3, ts3, x3, th3, Bg, flags, ms3)
        # the same step forced onto the CUDA-core adjoint kernel (tc_mode=-1)
        pc3.tune(tc_mode=-1)
        ts3b = make_train_step(pc3, df.setup(df.Adam(1e-3), c3))
        ms3b = timed(lambda: ts3b(x3, th3, None, Bg, flags), 2, 1, dist)
        ops["train_step_c3_cuda_cores"] = {"samples_per_s": Bg / (ms3b * 1e-3), "ms_per_step": ms3b, "global_batch": Bg,
                                           "scaling": "strong", "path": "FP32 FFMA adjoint kernel (tc_mode=-1)",
                                           "roofline": roofline_of("c3", "train", Bl / (ms3b * 1e-3), "fp32")}
        del x3, th3, pc3, ts3, ts3b
        torch.cuda.empty_cache()

    # ---- wide conditioners on the tensor cores: C4 (h=256) log-density + train step, C5 (h=512) sampling ----
    if not args.no_wide:
        # C4: d=32, n=8, 12 coupling layers, hidden 256 (+ NormalizationLayer); 256 Ki samples per GPU and step
        B4 = 1 << 18
        x4, th4 = device_inputs(32, 8, B4, dev, 7 + rank)
        c4, pc4 = packed("c4", dev, np.full(8, -1.0, np.float32), np.full(8, 2.0, np.float32))
        out4 = torch.empty(B4, device=dev)
        x4p, t4p = df.arrays.flat_view(x4).data_ptr(), df.arrays.flat_view(th4).data_ptr()
        pc4.ensure_scratch(B4)

        def step_c4_logpdf():
            df._lib.check(lib.dflow_logpdf(pc4.handle, pc4.W.data_ptr(), x4p, t4p, B4, None, flags, out4.data_ptr(), st))

        ms4 = timed(step_c4_logpdf, 3, 2, dist)
        ops["logpdf_c4"] = {"samples_per_s": world * B4 / (ms4 * 1e-3), "ms_per_step": ms4, "B_per_gpu": B4,
                            "roofline": roofline_of("c4", "logpdf", B4 / (ms4 * 1e-3), "tf32x3")}
        ts4 = make_train_step(pc4, df.setup(df.Adam(1e-3), c4))

        def step_c4_train():
            ts4(x4, th4, None, B4 * world, flags)

        ms4t = timed(step_c4_train, 2, 1, dist)
        ops["train_step_c4"] = {"samples_per_s": world * B4 / (ms4t * 1e-3), "ms_per_step": ms4t, "B_per_gpu": B4,
                                "scaling": "weak", "allreduce_bytes": 4 * (pc4.P + 2), "collective": coll(ts4),
                                "roofline": roofline_of("c4", "train", B4 / (ms4t * 1e-3), "tf32x3")}
        del x4, th4, pc4, ts4, out4
        torch.cuda.empty_cache()
        # C5: d=64, n=16, 16 coupling layers, hidden 512; inverse sampling with a fixed condition, no communication
        B5 = 1 << 18
        c5, pc5 = packed("c5", dev, np.full(16, -1.0, np.float32), np.full(16, 2.0, np.float32))
        th5 = torch.full((16,), 0.5, device=dev)
        x5 = df.jl_empty((64, B5), dev)
        x5p = df.arrays.flat_view(x5).data_ptr()
        pc5.ensure_scratch(B5)

        def step_c5_sample():
            df._lib.check(lib.dflow_sample_rng(pc5.handle, pc5.W.data_ptr(), 777, 0, rank * B5, None, th5.data_ptr(), B5,
                                               flags, x5p, st))

        ms5 = timed(step_c5_sample, 3, 2, dist)
        ops["sample_rng_c5"] = {"samples_per_s": world * B5 / (ms5 * 1e-3), "ms_per_step": ms5, "B_per_gpu": B5,
                                "seconds_for_1e9_samples": 1e9 / (world * B5 / (ms5 * 1e-3)),
                                "roofline": roofline_of("c5", "sample", B5 / (ms5 * 1e-3), "tf32x3")}
        assert torch.isfinite(df.arrays.flat_view(x5)[:: max(1, 64 * B5 // 4096)]).all()
        del x5, pc5
        torch.cuda.empty_cache()

    head = roofline_of("c2", "logpdf", B / (ms * 1e-3), "fp32")
    traffic, traffic_src = None, None
    tp_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp_file):
        try:
            # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel, scaled per launch
            tj = json.load(open(tp_file))
            traffic = float(tj["chain_fwd_kernel_bytes_per_sample"]) * B
            traffic_src = tj.get("source", "profiles/traffic.json")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2 logpdf: d=5 n=2 L=3 RNVP h=16 + NormalizationLayer, B=%d per GPU" % B,
                   "l2": "inputs (%.1f GB per step) larger than L2; no flush" % (32.0 * B / 1e9),
                   "parallelism": "sample-sharded x%d, no data-path collective" % world},
        "clocks": cs.summary(),
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": head["hbm_gbs"], "peak": hbm, "unit": "GB/s", "frac": head["hbm_frac"],
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": how,
                     "note": "binding pipe is FP32 FMA (138 flop/B >> ridge 11): see fma_* keys",
                     "fma_achieved_tflops": head["pipe_tflops"], "fma_peak_tflops": FP32_FMA_PEAK_TFLOPS,
                     "fma_frac": head["pipe_frac"]},
        "ops": ops,
    }
    if checks:
        line["multi_gpu_checks"] = checks
    if rank == 0 and world == 1 and not args.no_cpu:  # reported on rank 0 at N=1 only (bounded sample)
        line["cpu_baseline"] = cpu_baseline()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def dp_correctness(df, dist, pc, state, ts, x, th, flags, rank, world, dev):
    """Asserted in every multi-GPU run: (1) the replicas are bit-identical after the timed train steps; (2) the data-
    parallel gradient (per-rank shards, seed 1/B_global, summed over ranks) equals the single-rank gradient of the gathered
    batch to 1e-4 of its max-norm; (3) one step of the step under test from a common state leaves the same parameters as
    the NCCL all-reduce + Adam step (to summation-order rounding)."""
    from densityflows.jl_b200.flows import TrainStep

    out = {}
    # (1) replicas
    wsum = torch.stack([pc.W.double().sum(), (pc.W.double() * torch.arange(pc.P, device=dev)).sum(),
                        state.m.double().sum(), state.v.double().sum()])
    lo, hi = wsum.clone(), wsum.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "replicas diverged after the data-parallel train steps"
    out["replicas_bit_identical_after_train_steps"] = True
    # (2) gradient
    Bs = 1 << 15
    xs, ths = df.to_jl(x[:, :Bs]), df.to_jl(th[:, :Bs])
    g = torch.zeros(pc.P + 2, device=dev)
    pc.loss_grad(xs, ths, g[: pc.P], g[pc.P:], 1.0 / (Bs * world), flags)
    dist.all_reduce(g)
    gx = [torch.empty_like(df.arrays.flat_view(xs)) for _ in range(world)]
    gt = [torch.empty_like(df.arrays.flat_view(ths)) for _ in range(world)]
    dist.all_gather(gx, df.arrays.flat_view(xs).contiguous())
    dist.all_gather(gt, df.arrays.flat_view(ths).contiguous())
    xa = torch.cat(gx).view(Bs * world, x.shape[0]).t()
    ta = torch.cat(gt).view(Bs * world, th.shape[0]).t()
    gf = torch.zeros(pc.P + 2, device=dev)
    pc.loss_grad(xa, ta, gf[: pc.P], gf[pc.P:], 1.0 / (Bs * world), flags)
    err = float((g[: pc.P] - gf[: pc.P]).abs().max() / gf[: pc.P].abs().max())
    assert err <= 1e-4, f"data-parallel gradient differs from the full-batch gradient: {err}"
    assert abs(float(g[pc.P] - gf[pc.P])) <= 1e-5 * abs(float(gf[pc.P]))
    out["dp_gradient_vs_full_batch_rel_maxnorm"] = err
    # (3) the fused peer step against the NCCL step, one step from a common state
    w0, m0, v0, t0 = pc.W.clone(), state.m.clone(), state.v.clone(), state.t
    ts(xs, ths, None, Bs * world, flags)
    w_a = pc.W.clone()
    pc.W.copy_(w0)
    state.m.copy_(m0)
    state.v.copy_(v0)
    state.t = t0
    TrainStep(pc, state)(xs, ths, None, Bs * world, flags)
    dw = float((pc.W - w_a).abs().max())
    assert dw <= 2e-6, f"fused peer all-reduce + Adam differs from NCCL all-reduce + Adam: {dw}"
    out["peer_step_vs_nccl_step_max_abs_dW"] = dw
    pc.W.copy_(w0)
    state.m.copy_(m0)
    state.v.copy_(v0)
    state.t = t0
    dist.barrier()
    return out


def dp_split(df, dist, lib, pc3, ts3, x3, th3, Bg, flags, ms_step):
    """Where a data-parallel C3 step spends its time on this rank: adjoint kernels, waiting in the peer barrier (skew between
    the ranks), and the reduce + Adam part of the fused kernel (CUDA events + the kernel's own wait counter)."""
    import ctypes as C

    st = torch.cuda.current_stream().cuda_stream
    w0, s0 = C.c_int64(), C.c_int64()
    df._lib.check(lib.dflow_dp_wait_stats(ts3.dp, st, C.byref(w0), C.byref(s0)))
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    k = 3
    t_adj = t_dp = 0.0
    for _ in range(k):
        dist.barrier()
        torch.cuda.synchronize()
        buf = ts3._next_buffer()
        buf.zero_()
        e[0].record()
        pc3.loss_grad(x3, th3, buf[: ts3.P], buf[ts3.P:], 1.0 / Bg, flags, None)
        e[1].record()
        ts3.state.t += 1
        r = ts3.state.rule
        df._lib.check(lib.dflow_dp_allreduce_adam(ts3.dp, pc3.W.data_ptr(), ts3.state.m.data_ptr(), ts3.state.v.data_ptr(), r.eta,
                                                  r.beta[0], r.beta[1], r.epsilon, ts3.state.t, ts3.loss2.data_ptr(), st))
        e[2].record()
        torch.cuda.synchronize()
        t_adj += e[0].elapsed_time(e[1])
        t_dp += e[1].elapsed_time(e[2])
    w1, s1 = C.c_int64(), C.c_int64()
    df._lib.check(lib.dflow_dp_wait_stats(ts3.dp, st, C.byref(w1), C.byref(s1)))
    clk_mhz = 1965.0
    wait_ms = (w1.value - w0.value) / max(1, s1.value - s0.value) / (clk_mhz * 1e3)
    v = torch.tensor([t_adj / k, t_dp / k, wait_ms], device="cuda")
    vmax, vmin = v.clone(), v.clone()
    dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(vmin, op=dist.ReduceOp.MIN)
    return {"adjoint_ms_max_over_ranks": float(vmax[0]), "adjoint_ms_min_over_ranks": float(vmin[0]),
            "fused_reduce_adam_kernel_ms_max": float(vmax[1]), "of_which_peer_barrier_wait_ms_max": float(vmax[2]),
            "of_which_peer_barrier_wait_ms_min": float(vmin[2]), "step_ms": ms_step,
            "note": "barrier wait = skew between the ranks' adjoint sweeps (SM clocks of the waiting CTA at 1965 MHz); "
                    "reduce + Adam proper = kernel time - wait"}


# ---- CPU legs (Flux-equivalent torch-CPU restatement, oracle/torch_ref.py) ----------------------------------------
def _cpu_chain(cfg="c2"):
    from oracle import torch_ref as T

    return T.TorchChain(oracle_chain(cfg), torch.float32)


def cpu_logpdf_rate(nthreads: int, chunk: int, total: int, budget_s: float):
    """samples/s of logpdf over `total` C2-shaped samples evaluated in chunks of `chunk` (None: one call)."""
    from oracle import dflow_oracle as O

    torch.set_num_threads(nthreads)
    tc = _cpu_chain()
    x, th = O.synthetic_data(D, N_COND, total, seed=5)
    thn = O.normalize_input(th, th.min(axis=1), th.max(axis=1))
    xt, tt = torch.from_numpy(x), torch.from_numpy(thn)
    chunk = total if chunk is None else chunk
    with torch.no_grad():
        tc.logpdf(xt[:, :chunk], tt[:, :chunk])  # warm-up
        t0 = time.perf_counter()
        n = 0
        while True:
            for c0 in range(0, total, chunk):
                tc.logpdf(xt[:, c0:c0 + chunk], tt[:, c0:c0 + chunk])
            n += total
            dt = time.perf_counter() - t0
            if dt > budget_s:
                break
    return n / dt


def cpu_best_logpdf(nthreads: int, budget_s: float):
    """Same procedure for `cpu_baseline` and `--impl reference`: the Flux path evaluated on 2^20 samples in chunks of 2^16 /
    2^18 / 2^20 (a chunk's h x B temporaries either fit the caches or not), best one reported, all three listed."""
    total = 1 << 20
    rates = {}
    for ch in (1 << 16, 1 << 18, 1 << 20):
        rates[str(ch)] = cpu_logpdf_rate(nthreads, ch, total, budget_s / 3)
    best = max(rates, key=lambda k: rates[k])
    return rates[best], int(best), rates


def cpu_train_rate(nthreads: int, B: int, budget_s: float):
    """C2 train step on the CPU: autograd through the chain + Adam (Flux.gradient + Optimisers.update!)."""
    from oracle import dflow_oracle as O
    from oracle import torch_ref as T

    torch.set_num_threads(nthreads)
    tc = _cpu_chain()
    x, th = O.synthetic_data(D, N_COND, B, seed=6)
    thn = O.normalize_input(th, th.min(axis=1), th.max(axis=1))
    xt, tt = torch.from_numpy(x), torch.from_numpy(thn)
    P = tc.flat_params().numel()
    m, v = torch.zeros(P), torch.zeros(P)
    T.train_step(tc, xt, tt, m, v, 1)
    t0 = time.perf_counter()
    n = 0
    while True:
        T.train_step(tc, xt, tt, m, v, n + 2)
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s:
            break
    return n * B / dt, dt / n


def cpu_sample_rate(nthreads: int, B: int, budget_s: float):
    """sample(flow, B, θ::Tuple) on the CPU: randn + the sampling direction (src/Flows.jl:174-185)."""
    torch.set_num_threads(nthreads)
    tc = _cpu_chain()
    tt = torch.full((N_COND, B), 0.5)
    with torch.no_grad():
        tc.forward(torch.randn(D, B), tt)
        t0 = time.perf_counter()
        n = 0
        while True:
            tc.forward(torch.randn(D, B), tt)
            n += 1
            dt = time.perf_counter() - t0
            if dt > budget_s:
                break
    return n * B / dt


def cpu_c1_step(nthreads: int, budget_s: float):
    """README configuration: minibatch steps of 64 samples (gather + gradient + Adam), seconds per step."""
    from oracle import dflow_oracle as O
    from oracle import torch_ref as T

    torch.set_num_threads(nthreads)
    tc = _cpu_chain()
    x, th = O.synthetic_data(D, N_COND, 100000, seed=7)
    thn = O.normalize_input(th, th.min(axis=1), th.max(axis=1))
    xt, tt = torch.from_numpy(x), torch.from_numpy(thn)
    P = tc.flat_params().numel()
    m, v = torch.zeros(P), torch.zeros(P)
    perm = torch.randperm(90000)
    n = 0
    T.train_step(tc, xt[:, perm[:64]], tt[:, perm[:64]], m, v, 1)
    t0 = time.perf_counter()
    while True:
        idx = perm[(n * 64) % 89000:(n * 64) % 89000 + 64]
        T.train_step(tc, xt[:, idx], tt[:, idx], m, v, n + 2)
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s:
            break
    return dt / n


def cpu_baseline():
    cores = os.cpu_count() or 1
    rate, chunk, rates = cpu_best_logpdf(cores, 9.0)
    rate1 = cpu_logpdf_rate(1, 1 << 16, 1 << 18, 3.0)
    tr_rate, tr_s = cpu_train_rate(cores, 1 << 18, 4.0)
    sm_rate = cpu_sample_rate(cores, 1 << 18, 3.0)
    c1_s = cpu_c1_step(cores, 3.0)
    c1_s1 = cpu_c1_step(1, 2.0)
    return {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": "C2 logpdf on 2^20 samples in chunks of %d (best of 2^16 / 2^18 / 2^20; torch-CPU restatement of the "
                      "Flux path, oracle/torch_ref.py; Julia/Flux is not installable here)" % chunk,
            "logpdf_by_chunk": rates, "logpdf_1_thread": rate1,
            "train_step_c2": {"samples_per_s": tr_rate, "s_per_step": tr_s, "B": 1 << 18, "what": "autograd + Adam"},
            "sample_c2": {"samples_per_s": sm_rate, "B": 1 << 18, "what": "randn + sampling direction"},
            "train_step_c1_batch64": {"us_per_step": c1_s * 1e6, "us_per_step_1_thread": c1_s1 * 1e6,
                                      "samples_per_s": 64 / c1_s, "what": "gather + autograd + Adam, batchsize 64"}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # same procedure as cpu_baseline: calibrate the chunk size during warm-up, then time K steps of 2^20 samples each
    _, chunk, rates = cpu_best_logpdf(cores, 3.0 * max(1, args.warmup))
    from oracle import dflow_oracle as O

    torch.set_num_threads(cores)
    tc = _cpu_chain()
    Bc = 1 << 20  # bounded sample of the C2 workload per step
    x, th = O.synthetic_data(D, N_COND, Bc, seed=5)
    thn = O.normalize_input(th, th.min(axis=1), th.max(axis=1))
    xt, tt = torch.from_numpy(x), torch.from_numpy(thn)
    with torch.no_grad():
        t0 = time.perf_counter()
        for _ in range(args.steps):
            for c0 in range(0, Bc, chunk):
                tc.logpdf(xt[:, c0:c0 + chunk], tt[:, c0:c0 + chunk])
        dt = time.perf_counter() - t0
    v = Bc * args.steps / dt
    sample = ("C2 logpdf, %d samples per step evaluated in chunks of %d (best of 2^16 / 2^18 / 2^20, calibrated in warm-up); "
              "bounded sample of the 1e8-sample workload" % (Bc, chunk))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": int(args.gpus),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2 logpdf: d=5 n=2 L=3 RNVP h=16 + NormalizationLayer (CPU restatement of the "
                               "reference's Flux path; the Julia reference cannot be installed: no julia binary, "
                               "no network)"},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample,
                         "logpdf_by_chunk": rates},
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
    ap.add_argument("--profile", action="store_true",
                    help="for ncu launch lists: skip the 5 600-launch comparison leg (C1 epoch on per-minibatch launches)")
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
