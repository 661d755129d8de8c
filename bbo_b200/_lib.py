"""ctypes binding of libbbo_b200.so (the C ABI declared in include/bbo_b200.h).

There is no CPU fallback: if the shared library is missing, or a compute entry
point is called without a CUDA device, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libbbo_b200.so")

BBO_OK = 0
STATUS_NAMES = {0: "BBO_OK", 1: "BBO_ERR_BAD_SHAPE", 2: "BBO_ERR_BAD_ARG", 3: "BBO_ERR_WORKSPACE",
                4: "BBO_ERR_CUDA", 5: "BBO_ERR_NOT_PD", 6: "BBO_ERR_NAN"}
KERNEL_RBF, KERNEL_MATERN52 = 0, 1
ACQ_EI, ACQ_UCB, ACQ_PI = 0, 1, 2
PREC_FP64, PREC_TF32, PREC_TF32_REFINE = 0, 1, 2
PRECISIONS = {"fp64": PREC_FP64, "tf32": PREC_TF32, "tf32+refine": PREC_TF32_REFINE}
REFINE_KMAX = 32
ROW_ALIGN = 128
TOPK_MAX = 128
APPEND_MAX = 32
HOST_STAGE_CHUNKS = 8


class GpHyper(C.Structure):
    _fields_ = [("kind", C.c_int32), ("d", C.c_int32), ("pairwise_eps", C.c_double), ("noise", C.c_double)]


class GpState(C.Structure):
    _fields_ = [("h", GpHyper), ("n", C.c_int64), ("X", C.c_void_p), ("hyp", C.c_void_p),
                ("linv", C.c_void_p), ("alpha", C.c_void_p), ("linv_tf32", C.c_void_p)]


class Acq(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("best_f", C.c_double), ("eps", C.c_double),
                ("beta", C.c_double)]


class BboError(RuntimeError):
    pass


_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
_SZ = C.c_size_t
_D = C.c_double

# name -> (restype, argtypes); must list every symbol of include/bbo_b200.h
SIGNATURES = {
    "bbo_version": (C.c_int, []),
    "bbo_last_cuda_error": (C.c_char_p, []),
    "bbo_launch_count": (_I64, []),
    "bbo_profile_enable": (None, [_I32]),
    "bbo_profile_read": (C.c_int, [_I32, C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "bbo_padded_n": (_I64, [_I64]),
    "bbo_tf32_rows": (_I64, [_I64]),
    "bbo_gp_fit_workspace_bytes": (_SZ, [_I64, _I64, _I64]),
    "bbo_gp_score_workspace_bytes": (_SZ, [_I64, _I64, _I64, _I32]),
    "bbo_gp_predict_workspace_bytes": (_SZ, [_I64, _I64, _I64, _I32]),
    "bbo_cov_matrix": (C.c_int, [C.POINTER(GpHyper), _P, _P, _I64, _P, _I64, _P, _I64, _P]),
    "bbo_gp_fit_value_grad": (C.c_int, [C.POINTER(GpHyper), _I64, _I64, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "bbo_gp_fit_value_grad_inputs": (C.c_int, [C.POINTER(GpHyper), _I64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "bbo_gp_fit_adam": (C.c_int, [C.POINTER(GpHyper), _I64, _I64, _I32, _I32, _P, _P, _P, _P, _P, _I64, _I64,
                                  _D, _D, _P, _P, _P, _SZ, _P]),
    "bbo_gp_theta_to_hyp": (C.c_int, [_I64, _I64, _I32, _I32, _P, _P, _P]),
    "bbo_gp_factorize": (C.c_int, [C.POINTER(GpHyper), _I64, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "bbo_gp_append_workspace_bytes": (_SZ, [_I64, _I64]),
    "bbo_gp_append": (C.c_int, [C.POINTER(GpHyper), _I64, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "bbo_gp_make_tf32": (C.c_int, [_I64, _P, _P, _P]),
    "bbo_gp_predict": (C.c_int, [C.POINTER(GpState), _P, _I64, _P, _P, _P, _P, _SZ, _I64, _P]),
    "bbo_acq_eval": (C.c_int, [C.POINTER(Acq), _P, _P, _P, _I64, _I32, _P]),
    "bbo_gp_score_pool": (C.c_int, [C.POINTER(GpState), C.POINTER(Acq), _I32, _P, _I32, _I64, _I64, _I32, _P, _P,
                                    _P, _P, _P, _P, _P, _SZ, _I64, _I32, _P]),
    "bbo_gp_score_pool_host": (C.c_int, [C.POINTER(GpState), C.POINTER(Acq), _I32, _P, _I32, _I64, _I64, _I32,
                                         _P, _P, _P, _P, _P, _P, _SZ, _I64, _P]),
    "bbo_refine_shortlist": (_I32, [_I32]),
    "bbo_gp_refine_info": (C.c_int, [_P, _SZ, _I64, _I64, _I64, C.POINTER(C.c_double), _P]),
    "bbo_gp_score_batched_workspace_bytes": (_SZ, [_I64, _I64, _I64, _I64]),
    "bbo_gp_score_pool_batched": (C.c_int, [C.POINTER(GpHyper), _I64, _I64, _P, _P, _P, _P, _P, _P, _I64, _I32, _P, _P,
                                            _P, _SZ, _P]),
    "bbo_topk_merge": (C.c_int, [_P, _P, _I64, _I32, _P, _P, _P]),
    "bbo_topk_pack": (C.c_int, [_P, _P, _I32, _P, _P]),
    "bbo_topk_merge_packed": (C.c_int, [_P, _I64, _I32, _P, _P, _P]),
    "bbo_warp_kumaraswamy": (C.c_int, [_P, _I64, _D, _D, _D, _P, _I32, _P]),
    "bbo_pool_uniform": (C.c_int, [C.c_uint64, _I64, _I64, _I64, _P, _I32, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BboError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C bbo_b200/csrc`.  bbo_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != BBO_OK:
        msg = STATUS_NAMES.get(rc, str(rc))
        if rc == 4:
            msg += ": " + (load().bbo_last_cuda_error() or b"").decode()
        raise BboError(f"{what or 'libbbo_b200'} failed: {msg}")


def require_cuda(device=None) -> torch.device:
    """The product path is CUDA only; fail loudly otherwise."""
    if not torch.cuda.is_available():
        raise BboError("bbo_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.device("cuda" if device is None else device)
    if dev.type != "cuda":
        raise BboError(f"bbo_b200 needs a CUDA device, got {dev}; there is no CPU fallback")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


PROF_TAGS = {"vnorm": 0, "kstar": 1, "select": 2, "chol": 3, "trtri": 4, "grad": 5, "tf32": 6}


def profile_read(tag: str):
    """(regions, total_ms) of a profiled library region since the last read."""
    n, ms = C.c_int64(0), C.c_double(0.0)
    check(load().bbo_profile_read(PROF_TAGS[tag], C.byref(n), C.byref(ms)), "bbo_profile_read")
    return n.value, ms.value


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
