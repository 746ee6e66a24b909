"""Builds densityflows.jl_b200/lib/libdflow.so for sm_100a with nvcc (cross-compiles without a GPU).

One object per (HP, S) instantiation of the chain kernels, compiled in parallel, then one shared link.
Objects are rebuilt only when a source is newer (or `force=True`).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libdflow.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]

# (HP, samples per thread, register-resident activations)
FWD_INST = [(16, 1, 0), (16, 2, 0), (16, 4, 0), (32, 1, 0), (32, 2, 0), (32, 4, 0), (64, 1, 0), (64, 2, 0)]
CFWD_INST = [(16, 4), (32, 2)]  # constant-bank forward kernels (weights as uniform-datapath operands)
GRAD2_INST = [(16, 1), (16, 2), (16, 4), (32, 1), (32, 2), (64, 1), (64, 2)]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libdflow.so cannot be built (there is no CPU fallback)")


def _units():
    units = [("dflow_api.o", "dflow_api.cu", []), ("dflow_kernels.o", "dflow_kernels.cu", [])]
    for hp, s, reg in FWD_INST:
        units.append((f"inst_fwd_{hp}_{s}_{reg}.o", "dflow_inst.cu",
                      ["-DDFLOW_INST_FWD", f"-DDFLOW_HP={hp}", f"-DDFLOW_S={s}", f"-DDFLOW_REG={reg}"]))
    for hp, s in CFWD_INST:
        units.append((f"inst_cfwd_{hp}_{s}.o", "dflow_inst.cu",
                      ["-DDFLOW_INST_CFWD", "-DDFLOW_CBANK", f"-DDFLOW_HP={hp}", f"-DDFLOW_S={s}"]))
    for hp, s in GRAD2_INST:
        units.append((f"inst_grad2_{hp}_{s}.o", "dflow_inst.cu", ["-DDFLOW_INST_GRAD2", f"-DDFLOW_HP={hp}", f"-DDFLOW_S={s}"]))
    if os.path.exists(os.path.join(CSRC, "dflow_tc.cu")):
        units.append(("dflow_tc.o", "dflow_tc.cu", []))
    units.append(("dflow_small.o", "dflow_small.cu", []))
    if os.path.exists(os.path.join(CSRC, "dflow_dp.cu")):
        units.append(("dflow_dp.o", "dflow_dp.cu", []))
    return units


def _newest_source() -> float:
    t = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    return max(t, os.path.getmtime(os.path.abspath(__file__)))


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    src_t = _newest_source()
    todo = []
    objs = []
    for obj, src, defs in _units():
        op = os.path.join(OBJ, obj)
        objs.append(op)
        if force or not os.path.exists(op) or os.path.getmtime(op) < src_t:
            cmd = [nvcc, *ARCH, *COMMON, *defs, "-c", os.path.join(CSRC, src), "-o", op]
            if ptxas_info:
                cmd[1:1] = ["-Xptxas", "-v"]
            todo.append((obj, cmd))

    def run(item):
        obj, cmd = item
        r = subprocess.run(cmd, capture_output=True, text=True)
        return obj, r.returncode, r.stdout + r.stderr

    if todo:
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as ex:
            for obj, rc, out in ex.map(run, todo):
                if verbose or rc != 0 or ptxas_info:
                    sys.stderr.write(f"[build] {obj}: rc={rc}\n{out}\n")
                if rc != 0:
                    raise RuntimeError(f"nvcc failed for {obj}:\n{out}")
    if todo or not os.path.exists(LIB):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, ptxas_info="--ptxas" in sys.argv))
