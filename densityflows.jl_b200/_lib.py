"""ctypes binding of libdflow.so (include/dflow.h).  No CPU fallback: a missing library is a hard error."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libdflow.so")

OK, E_INVALID_ARG, E_UNSUPPORTED, E_CUDA, E_NCCL, E_NOMEM = 0, -1, -2, -3, -4, -5
ELEM_RNVP, ELEM_NICE, ELEM_NORM = 0, 1, 2
ACT_IDENTITY, ACT_RELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3
THETA_NORMALIZE = 1

c_i32p = C.POINTER(C.c_int32)
c_f32p = C.POINTER(C.c_float)


class NetDesc(C.Structure):
    _fields_ = [("depth", C.c_int32), ("widths", c_i32p), ("acts", c_i32p), ("has_bias", C.c_int32)]


class ElemDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("n_af", C.c_int32),
        ("axis_af", c_i32p),
        ("n_id", C.c_int32),
        ("axis_id", c_i32p),
        ("s_net", NetDesc),
        ("t_net", NetDesc),
        ("x_min", c_f32p),
        ("x_max", c_f32p),
        ("alpha", C.c_float),
        ("beta", C.c_float),
    ]


class ChainDesc(C.Structure):
    _fields_ = [
        ("d", C.c_int32),
        ("n", C.c_int32),
        ("n_elems", C.c_int32),
        ("elems", C.POINTER(ElemDesc)),
        ("theta_min", c_f32p),
        ("theta_max", c_f32p),
    ]


class DpShard(C.Structure):
    """dflow_dp_shard (include/dflow.h): one rank's share of a minibatch for dflow_dp_train_step."""
    _fields_ = [("chain", C.c_void_p), ("W", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("x", C.c_void_p),
                ("theta", C.c_void_p), ("B", C.c_int64), ("idx", C.c_void_p), ("ws", C.c_void_p), ("ws_bytes", C.c_size_t),
                ("loss2_out", C.c_void_p), ("stream", C.c_void_p)]


class DflowError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libdflow error {code}: {msg}")
        self.code = code


class DflowUnsupported(DflowError, NotImplementedError):
    pass


class DflowInvalidArg(DflowError, ValueError):
    """Maps the reference's @assert / ArgumentError (src/Axes.jl:85, src/Blocks.jl:71, ...)."""


# every symbol include/dflow.h declares: (name, restype, argtypes)
vp = C.c_void_p
SYMBOLS = [
    ("dflow_version", C.c_int, []),
    ("dflow_last_error", C.c_char_p, []),
    ("dflow_chain_create", C.c_int, [C.POINTER(ChainDesc), C.POINTER(vp)]),
    ("dflow_chain_destroy", C.c_int, [vp]),
    ("dflow_param_count", C.c_int64, [vp]),
    ("dflow_chain_axes", C.c_int, [vp, C.c_int32, c_i32p, c_i32p, c_i32p, c_i32p]),
    ("dflow_param_offset", C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("dflow_chain_set_theta_range", C.c_int, [vp, c_f32p, c_f32p]),
    ("dflow_scratch_bytes", C.c_size_t, [vp, C.c_int64]),
    ("dflow_chain_set_scratch", C.c_int, [vp, vp, C.c_size_t]),
    ("dflow_normalize", C.c_int, [vp, vp, vp, vp, C.c_int64, C.c_int32, vp, vp, vp]),
    ("dflow_logpdf", C.c_int, [vp, vp, vp, vp, C.c_int64, vp, C.c_int32, vp, vp]),
    ("dflow_logpdf_grid", C.c_int, [vp, vp, vp, C.POINTER(C.c_int64), vp, C.c_int32, vp, vp]),
    ("dflow_logpdf_sum", C.c_int, [vp, vp, vp, vp, C.c_int64, vp, C.c_int32, vp, vp]),
    ("dflow_sample_inplace", C.c_int, [vp, vp, vp, vp, vp, C.c_int64, C.c_int32, vp]),
    ("dflow_forward_ldj", C.c_int, [vp, vp, vp, vp, C.c_int64, C.c_int32, vp, vp, vp]),
    ("dflow_sample_rng", C.c_int, [vp, vp, C.c_uint64, C.c_uint32, C.c_uint64, vp, vp, C.c_int64, C.c_int32, vp, vp]),
    ("dflow_workspace_bytes", C.c_size_t, [vp, C.c_int64]),
    ("dflow_loss_grad", C.c_int, [vp, vp, vp, vp, C.c_int64, vp, C.c_float, C.c_int32, vp, vp, vp, C.c_size_t, vp]),
    ("dflow_vjp", C.c_int, [vp, vp, vp, vp, C.c_int64, C.c_int32, vp, vp, vp, vp, vp, vp, C.c_size_t, vp]),
    ("dflow_adam_step", C.c_int, [vp, vp, vp, vp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int64, vp]),
    ("dflow_train_epoch", C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_float,
                                    C.c_float, C.POINTER(C.c_int64), C.c_int32, vp, vp, vp, C.c_size_t, vp]),
    ("dflow_minmax", C.c_int, [vp, C.c_int32, C.c_int64, vp, vp, vp]),
    ("dflow_shuffle_indices", C.c_int, [C.c_uint64, C.c_int64, C.c_int64, C.c_int64, vp, vp, vp]),
    ("dflow_logpdf_host", C.c_int, [vp, vp, vp, vp, C.c_int64, C.c_int32, vp, C.c_int64]),
    ("dflow_sample_host", C.c_int, [vp, vp, C.c_uint64, vp, C.c_int64, C.c_int32, vp, C.c_int64]),
    ("dflow_dp_create", C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.POINTER(vp), vp]),
    ("dflow_dp_connect", C.c_int, [vp, vp]),
    ("dflow_dp_grad_buffer", vp, [vp]),
    ("dflow_dp_allreduce_adam", C.c_int, [vp, vp, vp, vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int64, vp, vp]),
    ("dflow_dp_status", C.c_int, [vp, vp]),
    ("dflow_dp_set_timeout_ms", C.c_int, [vp, C.c_int64]),
    ("dflow_dp_wait_stats", C.c_int, [vp, vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("dflow_dp_destroy", C.c_int, [vp]),
    ("dflow_dp_create_local", C.c_int, [C.c_int32, c_i32p, C.c_int64, C.POINTER(vp)]),
    ("dflow_dp_train_step", C.c_int, [C.POINTER(vp), C.c_int32, vp, C.c_float, C.c_int32, C.c_float, C.c_float, C.c_float,
                                      C.c_float, C.c_int64]),
    ("dflow_dp_sync", C.c_int, [C.POINTER(vp), C.c_int32, vp]),
    ("dflow_set_tuning", C.c_int, [vp, C.c_char_p, C.c_int32]),
    ("dflow_launch_count", C.c_int64, [vp]),
]

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load libdflow.so (once).  Raises if it has not been built -- there is deliberately no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(densityflows.jl_b200 has no CPU or eager fallback)"
            )
        l = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc == OK:
        return
    msg = lib().dflow_last_error().decode("utf-8", "replace")
    if rc == E_UNSUPPORTED:
        raise DflowUnsupported(rc, msg)
    if rc == E_INVALID_ARG:
        raise DflowInvalidArg(rc, msg)
    raise DflowError(rc, msg)
