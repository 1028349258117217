"""ctypes binding of libkvmatch_gpu.so (include/kvmatch_gpu.h).

This is the same C ABI a JNI / Panama shim binds from the reference's Java classes.
There is no CPU fallback: a missing library or a missing B200 raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KVM_LIB") or os.path.join(_HERE, "libkvmatch_gpu.so")  # KVM_LIB: developer A/B builds

KVM_OK = 0
KVM_E_NODEVICE = -1
KVM_E_ARG = -2
KVM_E_OOM = -3
KVM_E_CUDA = -4
KVM_E_IO = -5
KVM_E_STATE = -6
KVM_E_RANGE = -7
KVM_E_NCCL = -8
KVM_OPT_CNSM_PATH, KVM_OPT_STREAM_FLAG_ALL, KVM_OPT_PLAN_CACHE = 1, 2, 3
KVM_CNSM_STREAM, KVM_CNSM_RELAY = 0, 1

ERROR_NAMES = {
    KVM_E_NODEVICE: "KVM_E_NODEVICE", KVM_E_ARG: "KVM_E_ARG", KVM_E_OOM: "KVM_E_OOM", KVM_E_CUDA: "KVM_E_CUDA",
    KVM_E_IO: "KVM_E_IO", KVM_E_STATE: "KVM_E_STATE", KVM_E_RANGE: "KVM_E_RANGE", KVM_E_NCCL: "KVM_E_NCCL",
}

# every symbol include/kvmatch_gpu.h declares
EXPORTS = [
    "kvm_abi_version", "kvm_create", "kvm_destroy", "kvm_last_error", "kvm_set_option", "kvm_load_series_host",
    "kvm_load_series_file",
    "kvm_verify_ed", "kvm_verify_cnsm_ed", "kvm_verify_dtw", "kvm_verify_cnsm_dtw", "kvm_verify_cnsm_ed_batch", "kvm_scan_ucr_dtw", "kvm_scan_ucr_ed",
    "kvm_envelope", "kvm_window_mean_runs", "kvm_window_mean_runs_all", "kvm_build_index_file", "kvm_index_image_from_runs", "kvm_image_free", "kvm_index_row_positions",
    "kvm_result_free", "kvm_runs_free",
    "kvm_multi_create", "kvm_multi_destroy", "kvm_multi_last_error", "kvm_multi_devices", "kvm_multi_load_series_host",
    "kvm_multi_verify", "kvm_comm_unique_id", "kvm_comm_init", "kvm_gather_result", "kvm_comm_ipc_handle", "kvm_comm_ipc_attach", "kvm_intervals_sort_merge", "kvm_intervals_intersect", "kvm_intervals_first_segment",
    "kvm_norm_intervals_sort_merge", "kvm_norm_intervals_intersect", "kvm_norm_intervals_first_segment",
]
KVM_ENGINE_ED, KVM_ENGINE_CNSM_ED, KVM_ENGINE_DTW, KVM_ENGINE_CNSM_DTW = 0, 1, 2, 3


class KvmError(RuntimeError):
    """Non-zero return code of the C ABI.  The Java shim maps this to IOException (engines' query() throws it)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


class KvmResult(C.Structure):
    _fields_ = [
        ("count", C.c_int64),
        ("offsets", C.c_void_p),     # const int32_t*
        ("distances", C.c_void_p),   # const double*
        ("cnt_candidate", C.c_int64),
        ("n_verified", C.c_int64),
        ("s_total", C.c_int64),
        ("n_gate_pass", C.c_int64),
        ("n_lb_pass", C.c_int64),
        ("n_exact", C.c_int64),
        ("kernel_ms", C.c_double),
        ("stage_ms", C.c_double * 4),
        ("n_launches", C.c_int32),
        ("h2d_bytes", C.c_int32),
        ("n_rewalked", C.c_int64),
        ("n_chains_rewalked", C.c_int64),
        ("n_dtw_cells", C.c_int64),
    ]


class KvmRuns(C.Structure):
    _fields_ = [
        ("count", C.c_int64),
        ("keys", C.POINTER(C.c_double)),
        ("first", C.POINTER(C.c_int32)),
        ("last", C.POINTER(C.c_int32)),
        ("kernel_ms", C.c_double),
        ("n_launches", C.c_int32),
        ("reserved", C.c_int32),
    ]


class KvmIndexInfo(C.Structure):
    _fields_ = [
        ("file_bytes", C.c_int64), ("n_runs", C.c_int64), ("n_intervals", C.c_int64), ("n_offsets", C.c_int64),
        ("n_rows_step1", C.c_int32), ("n_rows", C.c_int32), ("kernel_ms", C.c_double), ("host_ms", C.c_double),
    ]


_lib = None
_dp = C.c_void_p  # const double* / const int32_t* arguments are passed as plain addresses (cheapest ctypes path)
_ip = C.c_void_p


def load():
    """dlopen the library and declare the prototypes.  Does not touch the GPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: build it with `make -C kvmatch_b200/csrc` (or __graft_entry__.build()). "
            "kvmatch_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    R = C.POINTER(KvmResult)
    L.kvm_abi_version.restype = C.c_int
    L.kvm_create.argtypes = [C.POINTER(vp), C.c_int]
    L.kvm_destroy.argtypes = [vp]
    L.kvm_destroy.restype = None
    L.kvm_last_error.argtypes = [vp]
    L.kvm_last_error.restype = C.c_char_p
    L.kvm_set_option.argtypes = [vp, C.c_int32, C.c_int64]
    L.kvm_load_series_host.argtypes = [vp, _dp, C.c_int64, C.c_int64, C.c_int64]
    L.kvm_load_series_file.argtypes = [vp, C.c_char_p, C.c_int64, C.c_int64, C.c_int64]
    L.kvm_verify_ed.argtypes = [vp, _dp, C.c_int32, C.c_double, _ip, C.c_int32, C.c_int32, R]
    L.kvm_verify_cnsm_ed.argtypes = [vp, _dp, C.c_int32, C.c_double, C.c_double, C.c_double, _ip, C.c_int32, C.c_int32,
                                     R]
    L.kvm_verify_dtw.argtypes = [vp, _dp, C.c_int32, C.c_double, C.c_int32, _ip, C.c_int32, C.c_int32, R]
    L.kvm_verify_cnsm_dtw.argtypes = [vp, _dp, C.c_int32, C.c_double, C.c_int32, C.c_double, C.c_double, _ip,
                                      C.c_int32, C.c_int32, R]
    L.kvm_verify_cnsm_ed_batch.argtypes = [vp, _dp, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double, _ip, C.c_int32,
                                           C.c_int32, R]
    L.kvm_scan_ucr_dtw.argtypes = [vp, _dp, C.c_int32, C.c_double, C.c_int32, C.c_double, C.c_double, R]
    L.kvm_scan_ucr_ed.argtypes = [vp, _dp, C.c_int32, C.c_double, C.c_double, C.c_double, R]
    L.kvm_comm_unique_id.argtypes = [vp]
    L.kvm_comm_init.argtypes = [vp, vp, C.c_int32, C.c_int32]
    L.kvm_comm_ipc_handle.argtypes = [vp, vp]
    L.kvm_comm_ipc_attach.argtypes = [vp, vp]
    L.kvm_gather_result.argtypes = [vp, R, R, C.POINTER(C.c_double), C.POINTER(C.c_int32)]
    L.kvm_window_mean_runs.argtypes = [vp, C.c_int32, C.POINTER(KvmRuns)]
    L.kvm_envelope.argtypes = [vp, C.c_int32, C.c_int64, C.c_int32, vp, vp]
    L.kvm_window_mean_runs_all.argtypes = [vp, vp, C.c_int32, C.POINTER(KvmRuns)]
    L.kvm_build_index_file.argtypes = [vp, C.c_int32, C.c_char_p, C.POINTER(KvmIndexInfo)]
    L.kvm_index_image_from_runs.argtypes = [vp, vp, vp, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(KvmIndexInfo)]
    L.kvm_image_free.argtypes = [vp]
    L.kvm_image_free.restype = None
    L.kvm_index_row_positions.argtypes = [vp, C.c_int64, vp, C.c_int64, C.POINTER(C.c_int64)]
    L.kvm_result_free.argtypes = [vp, R]
    L.kvm_result_free.restype = None
    L.kvm_runs_free.argtypes = [vp, C.POINTER(KvmRuns)]
    L.kvm_runs_free.restype = None
    i64p, f64p = C.POINTER(C.c_int64), C.POINTER(C.c_double)
    L.kvm_intervals_sort_merge.argtypes = [vp, vp, C.c_int64, C.c_int32, vp, vp, C.c_int64, i64p, i64p, i64p]
    L.kvm_intervals_intersect.argtypes = [vp, vp, C.c_int64, vp, vp, C.c_int64, C.c_double, C.c_int32, vp, vp, C.c_int64, i64p, f64p]
    L.kvm_intervals_first_segment.argtypes = [vp, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, vp,
                                              C.c_int64, i64p, f64p]
    L.kvm_norm_intervals_sort_merge.argtypes = [vp, C.c_int64, C.c_int32, vp, C.c_int64, i64p, i64p, i64p]
    L.kvm_norm_intervals_intersect.argtypes = [vp, C.c_int64, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double,
                                               C.c_double, C.c_double, C.c_int32, C.c_int32, vp, C.c_int64, i64p]
    L.kvm_norm_intervals_first_segment.argtypes = [vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, C.c_int64,
                                                   i64p]
    L.kvm_multi_create.argtypes = [C.POINTER(vp), vp, C.c_int32]
    L.kvm_multi_destroy.argtypes = [vp]
    L.kvm_multi_destroy.restype = None
    L.kvm_multi_last_error.argtypes = [vp]
    L.kvm_multi_last_error.restype = C.c_char_p
    L.kvm_multi_devices.argtypes = [vp]
    L.kvm_multi_load_series_host.argtypes = [vp, _dp, C.c_int64, C.c_int64, C.c_int64]
    L.kvm_multi_verify.argtypes = [vp, C.c_int32, _dp, C.c_int32, C.c_double, C.c_int32, C.c_double, C.c_double, _ip,
                                   C.c_int32, C.c_int32, R]
    _lib = L
    return L


def as_f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data


def as_intervals(intervals):
    lr = np.ascontiguousarray(np.asarray(intervals, dtype=np.int32).reshape(-1, 2))
    return lr, lr.ctypes.data, int(lr.shape[0])


def copy_out(addr, count: int, dtype):
    """A numpy copy of `count` elements at a library-owned host address."""
    if not count:
        return np.zeros(0, dtype)
    return np.frombuffer(C.string_at(addr, count * np.dtype(dtype).itemsize), dtype=dtype)


def index_image_from_runs(keys, first, last):
    """kvm_index_image_from_runs: (file image bytes, KvmIndexInfo).  Host-only (no GPU needed)."""
    L = load()
    k = np.ascontiguousarray(keys, dtype=np.float64)
    f = np.ascontiguousarray(first, dtype=np.int32)
    l = np.ascontiguousarray(last, dtype=np.int32)
    img = C.c_void_p()
    info = KvmIndexInfo()
    rc = L.kvm_index_image_from_runs(k.ctypes.data, f.ctypes.data, l.ctypes.data, len(k), C.byref(img), C.byref(info))
    if rc != 0:
        raise KvmError(rc, "kvm_index_image_from_runs failed")
    data = C.string_at(img.value, info.file_bytes)
    L.kvm_image_free(img)
    return data, info
