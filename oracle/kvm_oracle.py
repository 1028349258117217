"""ctypes binding of the CPU oracle (oracle/libkvm_oracle.so).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs — never from the product package kvmatch_b200.

PARITY UNPINNED (see the header of kvm_oracle.cpp): the reference ships no golden vectors and
cannot be executed in this image (Java 8, no JVM).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkvm_oracle.so")


def build(force: bool = False) -> str:
    """Compile oracle/kvm_oracle.cpp (g++, -ffp-contract=off).  Building the checker is not using it."""
    src = os.path.join(_HERE, "kvm_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libkvm_oracle.so"])
    return _SO


class _Result(C.Structure):
    _fields_ = [
        ("count", C.c_int64),
        ("offsets", C.POINTER(C.c_int32)),
        ("distances", C.POINTER(C.c_double)),
        ("cnt_candidate", C.c_int64),
        ("n_verified", C.c_int64),
        ("s_total", C.c_int64),
        ("n_gate_pass", C.c_int64),
        ("n_kim_pass", C.c_int64),
        ("n_keogh_pass", C.c_int64),
        ("n_dtw", C.c_int64),
        ("terms", C.c_int64),
        ("dtw_cells", C.c_int64),
    ]


class _Runs(C.Structure):
    _fields_ = [
        ("count", C.c_int64),
        ("keys", C.POINTER(C.c_double)),
        ("first", C.POINTER(C.c_int32)),
        ("last", C.POINTER(C.c_int32)),
    ]


@dataclass
class OracleResult:
    offsets: np.ndarray        # int32
    distances: np.ndarray      # float64, sqrt(dist^2)
    cnt_candidate: int
    n_verified: int
    s_total: int
    n_gate_pass: int
    n_kim_pass: int
    n_keogh_pass: int
    n_dtw: int
    terms: int
    dtw_cells: int

    @property
    def count(self) -> int:
        return int(len(self.offsets))


class ReferenceThrows(Exception):
    """The reference would throw (IllegalArgumentException / ArrayIndexOutOfBounds) on this input."""


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.kvo_to_round.restype = C.c_double
        L.kvo_to_round.argtypes = [C.c_double]
        L.kvo_lower_upper_lemire.argtypes = [_dp, C.c_int, C.c_int, _dp, _dp]
        L.kvo_lb_kim.restype = C.c_double
        L.kvo_lb_kim.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.kvo_lb_keogh.restype = C.c_double
        L.kvo_lb_keogh.argtypes = [_ip, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.kvo_lb_keogh_data.restype = C.c_double
        L.kvo_lb_keogh_data.argtypes = [_ip, _dp, _dp, C.c_int, _dp, _dp, C.c_int, C.c_double, C.c_double, C.c_double]
        L.kvo_dtw.restype = C.c_double
        L.kvo_dtw.argtypes = [_dp, _dp, _dp, C.c_int, C.c_int, C.c_double]
        L.kvo_query_stats.argtypes = [_dp, C.c_int, _dp, _dp]
        L.kvo_sorted_query.argtypes = [_dp, C.c_int, _dp, _ip]
        R = C.POINTER(_Result)
        L.kvo_verify_ed.argtypes = [_dp, C.c_int64, _dp, C.c_int, C.c_double, _ip, C.c_int, C.c_int, R]
        L.kvo_verify_cnsm_ed.argtypes = [_dp, C.c_int64, _dp, C.c_int, C.c_double, C.c_double, C.c_double, _ip,
                                         C.c_int, C.c_int, R]
        L.kvo_verify_dtw.argtypes = [_dp, C.c_int64, _dp, C.c_int, C.c_double, C.c_int, _ip, C.c_int, C.c_int, R]
        L.kvo_verify_cnsm_dtw.argtypes = [_dp, C.c_int64, _dp, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double,
                                          _ip, C.c_int, C.c_int, R]
        L.kvo_window_mean_runs.argtypes = [_dp, C.c_int64, C.c_int64, C.c_int, C.POINTER(_Runs)]
        L.kvo_ucr_ed.argtypes = [_dp, C.c_int64, C.c_int64, _dp, C.c_int, C.c_double, C.c_double, C.c_double, R]
        L.kvo_ucr_dtw.argtypes = [_dp, C.c_int64, C.c_int64, _dp, C.c_int, C.c_double, C.c_int, C.c_double,
                                  C.c_double, R]
        L.kvo_result_free.argtypes = [R]
        L.kvo_runs_free.argtypes = [C.POINTER(_Runs)]
        L.kvo_write_series_be.argtypes = [C.c_char_p, _dp, C.c_int64]
        L.kvo_read_series_be.argtypes = [C.c_char_p, _dp, C.c_int64]
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)


def _check(rc):
    if rc == -2:
        raise ReferenceThrows()
    if rc != 0:
        raise ValueError(f"oracle error {rc}")


def _take(res: _Result) -> OracleResult:
    n = res.count
    off = np.ctypeslib.as_array(res.offsets, shape=(max(n, 1),))[:n].copy()
    dist = np.ctypeslib.as_array(res.distances, shape=(max(n, 1),))[:n].copy()
    out = OracleResult(off, dist, res.cnt_candidate, res.n_verified, res.s_total, res.n_gate_pass, res.n_kim_pass,
                       res.n_keogh_pass, res.n_dtw, res.terms, res.dtw_cells)
    lib().kvo_result_free(C.byref(res))
    return out


def _intervals(intervals):
    lr = np.ascontiguousarray(np.asarray(intervals, dtype=np.int32).reshape(-1, 2))
    return lr, lr.ctypes.data_as(_ip), lr.shape[0]


def to_round(v: float) -> float:
    return lib().kvo_to_round(float(v))


def lower_upper_lemire(t, r: int):
    t, tp = _d(t)
    l = np.zeros(len(t))
    u = np.zeros(len(t))
    _check(lib().kvo_lower_upper_lemire(tp, len(t), r, l.ctypes.data_as(_dp), u.ctypes.data_as(_dp)))
    return l, u


def lb_kim(t, q, j, length, mean, std, bsf):
    t, tp = _d(t)
    q, qp = _d(q)
    return lib().kvo_lb_kim(tp, qp, j, length, mean, std, bsf)


def lb_keogh(order, t, uo, lo, j, length, mean, std, bsf):
    order, op = _i(order)
    t, tp = _d(t)
    uo, up = _d(uo)
    lo, lp = _d(lo)
    cb = np.zeros(length)
    v = lib().kvo_lb_keogh(op, tp, up, lp, cb.ctypes.data_as(_dp), j, length, mean, std, bsf)
    return v, cb


def lb_keogh_data(order, qo, I, l, u, length, mean, std, bsf):
    order, op = _i(order)
    qo, qp = _d(qo)
    l, lp = _d(l)
    u, up = _d(u)
    cb = np.zeros(length)
    v = lib().kvo_lb_keogh_data(op, qp, cb.ctypes.data_as(_dp), I, lp, up, length, mean, std, bsf)
    return v, cb


def dtw(A, B, cb, r: int, bsf: float) -> float:
    A, ap = _d(A)
    B, bp = _d(B)
    cb, cp = _d(cb)
    return lib().kvo_dtw(ap, bp, cp, len(A), r, bsf)


def query_stats(q):
    q, qp = _d(q)
    a = C.c_double()
    b = C.c_double()
    lib().kvo_query_stats(qp, len(q), C.byref(a), C.byref(b))
    return a.value, b.value


def sorted_query(q):
    q, qp = _d(q)
    z = np.zeros(len(q))
    o = np.zeros(len(q), dtype=np.int32)
    lib().kvo_sorted_query(qp, len(q), z.ctypes.data_as(_dp), o.ctypes.data_as(_ip))
    return z, o


def verify_ed(series, q, epsilon, intervals, shift=0) -> OracleResult:
    s, sp = _d(series)
    q, qp = _d(q)
    lr, lp, K = _intervals(intervals)
    r = _Result()
    _check(lib().kvo_verify_ed(sp, len(s), qp, len(q), epsilon, lp, K, shift, C.byref(r)))
    return _take(r)


def verify_cnsm_ed(series, q, epsilon, alpha, beta, intervals, shift=0) -> OracleResult:
    s, sp = _d(series)
    q, qp = _d(q)
    lr, lp, K = _intervals(intervals)
    r = _Result()
    _check(lib().kvo_verify_cnsm_ed(sp, len(s), qp, len(q), epsilon, alpha, beta, lp, K, shift, C.byref(r)))
    return _take(r)


def verify_dtw(series, q, epsilon, rho, intervals, shift=0) -> OracleResult:
    s, sp = _d(series)
    q, qp = _d(q)
    lr, lp, K = _intervals(intervals)
    r = _Result()
    _check(lib().kvo_verify_dtw(sp, len(s), qp, len(q), epsilon, rho, lp, K, shift, C.byref(r)))
    return _take(r)


def verify_cnsm_dtw(series, q, epsilon, rho, alpha, beta, intervals, shift=0) -> OracleResult:
    s, sp = _d(series)
    q, qp = _d(q)
    lr, lp, K = _intervals(intervals)
    r = _Result()
    _check(lib().kvo_verify_cnsm_dtw(sp, len(s), qp, len(q), epsilon, rho, alpha, beta, lp, K, shift, C.byref(r)))
    return _take(r)


def window_mean_runs(series, w: int, n: int | None = None):
    """IndexBuilder step 1 for one window width: returns (keys f64, first i32, last i32)."""
    s, sp = _d(series)
    n = len(s) if n is None else n
    r = _Runs()
    _check(lib().kvo_window_mean_runs(sp, len(s), n, w, C.byref(r)))
    c = r.count
    keys = np.ctypeslib.as_array(r.keys, shape=(max(c, 1),))[:c].copy()
    first = np.ctypeslib.as_array(r.first, shape=(max(c, 1),))[:c].copy()
    last = np.ctypeslib.as_array(r.last, shape=(max(c, 1),))[:c].copy()
    lib().kvo_runs_free(C.byref(r))
    return keys, first, last


class _Bytes(C.Structure):
    _fields_ = [("count", C.c_int64), ("data", C.POINTER(C.c_ubyte)), ("n_rows_step1", C.c_int32), ("n_rows", C.c_int32)]


def index_file_image(series, w: int, n: int | None = None):
    """IndexBuilder steps 1+2 and IndexFileOperator.writeAll for one window width: (file bytes, rows before the
    step-2 merge, rows after)."""
    s, sp = _d(series)
    n = len(s) if n is None else n
    b = _Bytes()
    L = lib()
    L.kvo_index_file_image.restype = C.c_int
    _check(L.kvo_index_file_image(sp, C.c_int64(len(s)), C.c_int64(n), C.c_int(w), C.byref(b)))
    data = C.string_at(b.data, b.count)
    rows1, rows = b.n_rows_step1, b.n_rows
    L.kvo_bytes_free(C.byref(b))
    return data, rows1, rows


def ucr_ed(series, q, epsilon, alpha, beta, N=None) -> OracleResult:
    s, sp = _d(series)
    q, qp = _d(q)
    r = _Result()
    _check(lib().kvo_ucr_ed(sp, len(s), len(s) if N is None else N, qp, len(q), epsilon, alpha, beta, C.byref(r)))
    return _take(r)


def ucr_dtw(series, q, epsilon, rho, alpha, beta, N=None) -> OracleResult:
    s, sp = _d(series)
    q, qp = _d(q)
    r = _Result()
    _check(lib().kvo_ucr_dtw(sp, len(s), len(s) if N is None else N, qp, len(q), epsilon, rho, alpha, beta,
                             C.byref(r)))
    return _take(r)


def write_series_be(path: str, series) -> None:
    s, sp = _d(series)
    _check(lib().kvo_write_series_be(path.encode(), sp, len(s)))


def read_series_be(path: str, n: int) -> np.ndarray:
    out = np.zeros(n)
    _check(lib().kvo_read_series_be(path.encode(), out.ctypes.data_as(_dp), n))
    return out
