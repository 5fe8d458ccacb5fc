"""ctypes binding of the C ABI declared in include/skb.h (libskb.so, built from csrc/ for sm_100a).

There is deliberately no fallback: if the shared library is missing or no B200 is visible the calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libskb.so")

SKB_OK = 0
SKB_ERR_NOT_POSDEF = -6
KERNEL_RBF, KERNEL_MATERN12, KERNEL_MATERN32, KERNEL_MATERN52, KERNEL_HAMMING = 0, 1, 2, 3, 4
ACQ_EI, ACQ_PI, ACQ_LCB = 0, 1, 2
ACQ_INDEX = {"EI": ACQ_EI, "PI": ACQ_PI, "LCB": ACQ_LCB}
PRECISION = {"fp64": 0, "split_int8": 1, "split_int8_p7": 2}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


class StateDesc(C.Structure):
    _fields_ = [("n_train", C.c_int32), ("n_dims", C.c_int32), ("kernel", C.c_int32), ("reserved", C.c_int32),
                ("amplitude", C.c_double), ("bias", C.c_double), ("white", C.c_double),
                ("y_mean", C.c_double), ("y_std", C.c_double),
                ("length_scale", C.c_void_p), ("X_train", C.c_void_p), ("alpha", C.c_void_p), ("L_inv", C.c_void_p), ("K_inv", C.c_void_p)]


class AcqParams(C.Structure):
    _fields_ = [("y_opt", C.c_double), ("xi", C.c_double), ("kappa", C.c_double), ("kappa_inf", C.c_int32),
                ("acq_mask", C.c_int32), ("top_k", C.c_int32), ("precision", C.c_int32)]


class ScoreOut(C.Structure):
    _fields_ = [("mean", C.c_void_p), ("std", C.c_void_p), ("acq", C.c_void_p * 3),
                ("topk_val", C.c_void_p * 3), ("topk_idx", C.c_void_p * 3), ("first_nan", C.c_void_p * 3),
                ("n_var_clamped", C.c_void_p)]


class DimDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("low", C.c_double), ("high", C.c_double)]


class FitDesc(C.Structure):
    _fields_ = [("n_train", C.c_int32), ("n_dims", C.c_int32), ("kernel", C.c_int32), ("reserved", C.c_int32),
                ("alpha", C.c_double), ("X_train", C.c_void_p), ("y_train", C.c_void_p)]


class PeerGroup(C.Structure):
    _fields_ = [("buffers", C.c_void_p * 8), ("rank", C.c_int32), ("world", C.c_int32)]


class GradOut(C.Structure):
    _fields_ = [("mean", C.c_void_p), ("std", C.c_void_p), ("mean_grad", C.c_void_p), ("std_grad", C.c_void_p),
                ("acq", C.c_void_p * 3), ("acq_grad", C.c_void_p * 3)]


# every symbol include/skb.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("skb_abi_version", C.c_int, []),
    ("skb_last_error", C.c_char_p, []),
    ("skb_device_count", C.c_int, []),
    ("skb_state_create", C.c_int, [C.POINTER(StateDesc), C.c_int, C.POINTER(C.c_void_p)]),
    ("skb_state_destroy", None, [C.c_void_p]),
    ("skb_state_attach_kinv", C.c_int, [C.c_void_p, C.c_void_p]),
    ("skb_state_split_info", C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
    ("skb_state_split_bounds", C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("skb_state_decide_split", C.c_int, [C.c_void_p]),
    ("skb_state_stream", C.c_void_p, [C.c_void_p]),
    ("skb_state_sync", C.c_int, [C.c_void_p]),
    ("skb_state_set_scratch_limit", C.c_int, [C.c_void_p, C.c_uint64]),
    ("skb_state_set_timing", C.c_int, [C.c_void_p, C.c_int]),
    ("skb_state_get_timing", C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int]),
    ("skb_score_dev", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(AcqParams),
                                C.POINTER(ScoreOut), C.c_void_p]),
    ("skb_score_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(AcqParams),
                                 C.POINTER(ScoreOut)]),
    ("skb_predict_mean_dev", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    ("skb_predict_mean_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    ("skb_score_grad_dev", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(AcqParams), C.POINTER(GradOut),
                                     C.c_void_p]),
    ("skb_score_grad_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(AcqParams), C.POINTER(GradOut)]),
    ("skb_acq_from_posterior_host", C.c_int, [C.c_int, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.POINTER(AcqParams), C.c_void_p * 3, C.c_void_p * 3]),
    ("skb_ps_combine_host", C.c_int, [C.c_int, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("skb_topk_host", C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                C.c_void_p]),
    ("skb_predict_cov_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    ("skb_topk_dev", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p]),
    ("skb_topk_merge_records_dev", C.c_int, [C.c_int, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                             C.c_void_p]),
    ("skb_peer_buffer_bytes", C.c_int64, []),
    ("skb_topk_exchange_peer_dev", C.c_int, [C.c_int, C.POINTER(PeerGroup), C.c_uint64, C.c_void_p, C.c_int32, C.c_int32,
                                             C.c_void_p, C.c_int32, C.c_void_p]),
    ("skb_mt19937_uniform_dev", C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_int32), C.c_int64, C.c_void_p, C.c_void_p]),
    ("skb_candidates_generate_dev", C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_int32), C.c_int64, C.c_int32,
                                              C.POINTER(DimDesc), C.c_void_p, C.c_void_p]),
    ("skb_candidates_generate_jump_dev", C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                                   C.POINTER(DimDesc), C.c_void_p, C.c_void_p, C.c_void_p]),
    ("skb_gather_rows_dev", C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    ("skb_fit_create", C.c_int, [C.POINTER(FitDesc), C.c_int, C.POINTER(C.c_void_p)]),
    ("skb_fit_destroy", None, [C.c_void_p]),
    ("skb_fit_lml_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    ("skb_fit_lml_batch_host", C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    ("skb_fit_finalize_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("skb_state_create_from_fit", C.c_int, [C.c_void_p, C.POINTER(StateDesc), C.POINTER(C.c_void_p)]),
    ("skb_measure_peaks", C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("skb_debug_k2o", C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("skb_launch_count", C.c_int64, []),
]

_lib = None


class SkbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libskb error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Load libskb.so once.  Raises if it has not been built (python __graft_entry__.py / make -C csrc)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("CUDA extension %s is missing: build it with `make -C %s` "
                               "(there is no CPU fallback)" % (LIB_PATH, os.path.dirname(LIB_PATH)))
        handle = C.CDLL(LIB_PATH)
        for name, restype, argtypes in SYMBOLS:
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(rc):
    if rc != SKB_OK:
        raise SkbError(rc, lib().skb_last_error().decode("utf-8", "replace"))
