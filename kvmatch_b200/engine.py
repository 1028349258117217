"""Host-side mirror of the reference's engine classes for the phase-2 path, over the C ABI.

The reference is Java and no JVM exists in this image, so the host side the parity tests and the
bench drive is this Python mirror (same class names, argument meaning, statistics slots, output
lines and error behaviour as K/QueryEngine.java, K/NormQueryEngine.java, K/QueryEngineDtw.java,
K/NormQueryEngineDtw.java and K/IndexBuilder.java for the part of `query()` / `run()` that follows
`sortAndMergeIntervals`).  Phase 0/1 (query planning, index probing) stay in the Java classes and are
out of scope: `query()` here takes their output (`valid_positions`, `last_segment`), or scans the whole
series index-free when none is given.  The Java-side binding is shown in INTEGRATION.md / java/.
"""
from __future__ import annotations

import ctypes as C
import logging
import time
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .datagen import chain_intervals

logger = logging.getLogger("kvmatch_b200")

WU_LIST = (25, 50, 100, 200, 400)  # K/QueryEngine.java:51
EPOCH = 100000                     # K/IndexBuilder.java:136


@dataclass
class VerifyResult:
    offsets: np.ndarray      # int32, 1-based, ascending (= reference scan order)
    distances: np.ndarray    # float64
    cnt_candidate: int
    n_verified: int
    s_total: int
    n_gate_pass: int
    n_lb_pass: int
    n_exact: int
    kernel_ms: float
    n_launches: int
    stage_ms: tuple = (0.0, 0.0, 0.0, 0.0)
    h2d_bytes: int = 0
    n_rewalked: int = 0
    n_chains_rewalked: int = 0
    n_dtw_cells: int = 0

    @property
    def count(self) -> int:
        return int(len(self.offsets))


class StatisticInfo:
    """K/statistic/StatisticInfo.java — the six per-query slots the engines append to."""

    def __init__(self):
        self.values = []

    def append(self, v):
        self.values.append(float(v))

    def get_average(self):
        return sum(self.values) / len(self.values) if self.values else 0.0


class GpuSeries:
    """One kvm_ctx: a device-resident offset range [first, first+count-1] of a length-n series.
    Takes the place of the TimeSeriesOperator the engines read phase-2 data through
    (K/operator/TimeSeriesOperator.java:38)."""

    def __init__(self, device: int = 0):
        self._L = _lib.load()
        h = C.c_void_p()
        rc = self._L.kvm_create(C.byref(h), device)
        if rc != 0:
            raise _lib.KvmError(rc, self._L.kvm_last_error(None).decode())
        self._h = h
        self.n = self.first = self.count = 0

    def close(self):
        if getattr(self, "_h", None):
            self._L.kvm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise _lib.KvmError(rc, self._L.kvm_last_error(self._h).decode())

    def set_option(self, option: int, value: int):
        """kvm_set_option: _lib.KVM_OPT_* (cNSM path, flag-all self-check, plan cache).  Never changes a result."""
        self._check(self._L.kvm_set_option(self._h, option, value))
        return self

    def load(self, samples, n: int | None = None, first: int = 1):
        a, p = _lib.as_f64(samples)
        n = len(a) if n is None else n
        self._check(self._L.kvm_load_series_host(self._h, p, n, first, len(a)))
        self.n, self.first, self.count = n, first, len(a)
        return self

    def load_file(self, path: str, n: int, first: int = 1, count: int | None = None):
        count = n - first + 1 if count is None else count
        self._check(self._L.kvm_load_series_file(self._h, path.encode(), n, first, count))
        self.n, self.first, self.count = n, first, count
        return self

    def _take(self, r: _lib.KvmResult) -> VerifyResult:
        c = r.count
        off = _lib.copy_out(r.offsets, c, np.int32)
        dist = _lib.copy_out(r.distances, c, np.float64)
        out = VerifyResult(off, dist, r.cnt_candidate, r.n_verified, r.s_total, r.n_gate_pass, r.n_lb_pass, r.n_exact,
                           r.kernel_ms, r.n_launches, tuple(r.stage_ms), int(r.h2d_bytes), int(r.n_rewalked),
                           int(r.n_chains_rewalked), int(r.n_dtw_cells))
        self._L.kvm_result_free(self._h, C.byref(r))
        return out

    # ---- one process per GPU: the multi-GPU tail inside the library (kvm_comm_*, kvm_gather_result) ----
    def comm_init(self, rank: int | None = None, world: int | None = None, p2p: bool = True):
        """Join the library's NCCL communicator.  The 128-byte id is obtained on rank 0 and broadcast through
        torch.distributed (any backend); rank / world default to the process group's."""
        import torch.distributed as dist
        rank = dist.get_rank() if rank is None else rank
        world = dist.get_world_size() if world is None else world
        ident = (C.c_ubyte * 128)()
        if rank == 0:
            self._check(self._L.kvm_comm_unique_id(ident))
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=0)
        ident = (C.c_ubyte * 128).from_buffer_copy(box[0])
        self._check(self._L.kvm_comm_init(self._h, ident, rank, world))
        self.comm_world = world
        if p2p and 1 < world <= 8:
            # the peer-memory fast path of the exchange: every rank's buffer mapped into every other rank (CUDA IPC)
            mine = (C.c_ubyte * 64)()
            self._check(self._L.kvm_comm_ipc_handle(self._h, mine))
            every = [None] * world
            dist.all_gather_object(every, bytes(mine))
            blob = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(every))
            self._check(self._L.kvm_comm_ipc_attach(self._h, blob))
        return self

    def gather(self, local: VerifyResult):
        """COLLECTIVE: merge this rank's result with every other rank's (one packed ncclAllGather on the library's
        stream).  Returns (merged VerifyResult, best) with best = (distance, offset) of the reference's `Best:` line or
        None when nobody has an answer."""
        lo = np.ascontiguousarray(local.offsets, dtype=np.int32)
        ld = np.ascontiguousarray(local.distances, dtype=np.float64)
        a = _lib.KvmResult()
        a.count = len(lo)
        a.offsets = lo.ctypes.data
        a.distances = ld.ctypes.data
        a.cnt_candidate, a.n_verified, a.s_total = local.cnt_candidate, local.n_verified, local.s_total
        a.n_gate_pass, a.n_lb_pass, a.n_exact = local.n_gate_pass, local.n_lb_pass, local.n_exact
        a.kernel_ms, a.n_launches = local.kernel_ms, local.n_launches
        a.n_rewalked, a.n_chains_rewalked, a.n_dtw_cells = local.n_rewalked, local.n_chains_rewalked, local.n_dtw_cells
        m = _lib.KvmResult()
        bd, bo = C.c_double(), C.c_int32()
        self._check(self._L.kvm_gather_result(self._h, C.byref(a), C.byref(m), C.byref(bd), C.byref(bo)))
        c = m.count
        out = VerifyResult(_lib.copy_out(m.offsets, c, np.int32), _lib.copy_out(m.distances, c, np.float64), m.cnt_candidate,
                           m.n_verified, m.s_total, m.n_gate_pass, m.n_lb_pass, m.n_exact, m.kernel_ms, m.n_launches,
                           tuple(m.stage_ms), int(local.h2d_bytes), int(m.n_rewalked), int(m.n_chains_rewalked),
                           int(m.n_dtw_cells))
        return out, ((bd.value, bo.value) if c > 0 else None)

    def verify_ed(self, q, epsilon, intervals, shift=0) -> VerifyResult:
        q, qp = _lib.as_f64(q)
        lr, lp, K = _lib.as_intervals(intervals)
        r = _lib.KvmResult()
        self._check(self._L.kvm_verify_ed(self._h, qp, len(q), epsilon, lp, K, shift, C.byref(r)))
        return self._take(r)

    def verify_cnsm_ed_batch(self, queries, epsilon, alpha, beta, intervals, shift=0) -> list:
        """A set of equal-length queries over one interval list: one statistics pass, per-query results
        (kvm_verify_cnsm_ed_batch).  `queries`: 2-D array (Q, m)."""
        qs = np.ascontiguousarray(queries, dtype=np.float64)
        if qs.ndim != 2:
            raise ValueError("queries must be a (Q, m) array")
        lr, lp, K = _lib.as_intervals(intervals)
        res = (_lib.KvmResult * qs.shape[0])()
        self._check(self._L.kvm_verify_cnsm_ed_batch(self._h, qs.ctypes.data, qs.shape[0], qs.shape[1], epsilon, alpha, beta,
                                                     lp, K, shift, res))
        out = []
        for r in res:
            c = r.count
            out.append(VerifyResult(_lib.copy_out(r.offsets, c, np.int32), _lib.copy_out(r.distances, c, np.float64),
                                    r.cnt_candidate, r.n_verified, r.s_total, r.n_gate_pass, r.n_lb_pass, r.n_exact,
                                    r.kernel_ms, r.n_launches, tuple(r.stage_ms), int(r.h2d_bytes), int(r.n_rewalked),
                           int(r.n_chains_rewalked), int(r.n_dtw_cells)))
        return out

    def scan_ucr_dtw(self, q, epsilon, rho, alpha, beta) -> VerifyResult:
        """K/experiments/ucr/UcrDtwQueryExecutor.java:84-314: index-free cNSM-DTW scan, 0-based offsets."""
        q, qp = _lib.as_f64(q)
        r = _lib.KvmResult()
        self._check(self._L.kvm_scan_ucr_dtw(self._h, qp, len(q), epsilon, rho, alpha, beta, C.byref(r)))
        return self._take(r)

    def scan_ucr_ed(self, q, epsilon, alpha, beta) -> VerifyResult:
        """K/experiments/ucr/UcrEdQueryExecutor.java:101-183: index-free cNSM-ED scan on ONE never-reset statistics
        chain, 1-based offsets."""
        q, qp = _lib.as_f64(q)
        r = _lib.KvmResult()
        self._check(self._L.kvm_scan_ucr_ed(self._h, qp, len(q), epsilon, alpha, beta, C.byref(r)))
        return self._take(r)

    def build_index_file(self, w: int, path: str | None):
        """K/IndexBuilder.java:186-347 for one window width: returns the kvm_index_info fields."""
        info = _lib.KvmIndexInfo()
        self._check(self._L.kvm_build_index_file(self._h, w, path.encode() if path else None, C.byref(info)))
        return info

    def verify_cnsm_ed(self, q, epsilon, alpha, beta, intervals, shift=0) -> VerifyResult:
        q, qp = _lib.as_f64(q)
        lr, lp, K = _lib.as_intervals(intervals)
        r = _lib.KvmResult()
        self._check(self._L.kvm_verify_cnsm_ed(self._h, qp, len(q), epsilon, alpha, beta, lp, K, shift, C.byref(r)))
        return self._take(r)

    def verify_dtw(self, q, epsilon, rho, intervals, shift=0) -> VerifyResult:
        q, qp = _lib.as_f64(q)
        lr, lp, K = _lib.as_intervals(intervals)
        r = _lib.KvmResult()
        self._check(self._L.kvm_verify_dtw(self._h, qp, len(q), epsilon, rho, lp, K, shift, C.byref(r)))
        return self._take(r)

    def verify_cnsm_dtw(self, q, epsilon, rho, alpha, beta, intervals, shift=0) -> VerifyResult:
        q, qp = _lib.as_f64(q)
        lr, lp, K = _lib.as_intervals(intervals)
        r = _lib.KvmResult()
        self._check(self._L.kvm_verify_cnsm_dtw(self._h, qp, len(q), epsilon, rho, alpha, beta, lp, K, shift,
                                                C.byref(r)))
        return self._take(r)

    def window_mean_runs(self, w: int):
        r = _lib.KvmRuns()
        self._check(self._L.kvm_window_mean_runs(self._h, w, C.byref(r)))
        c = r.count
        keys = np.ctypeslib.as_array(r.keys, shape=(c,)).copy() if c else np.zeros(0)
        first = np.ctypeslib.as_array(r.first, shape=(c,)).copy() if c else np.zeros(0, np.int32)
        last = np.ctypeslib.as_array(r.last, shape=(c,)).copy() if c else np.zeros(0, np.int32)
        ms, nl = r.kernel_ms, r.n_launches
        self._L.kvm_runs_free(self._h, C.byref(r))
        return keys, first, last, ms, nl

    def envelope(self, r: int, first: int, length: int):
        """kvm_envelope: (lower, upper) = DtwUtils.lowerUpperLemire over samples [first, first+length-1] (1-based)."""
        lo = np.empty(length)
        up = np.empty(length)
        self._check(self._L.kvm_envelope(self._h, r, first, length, lo.ctypes.data, up.ctypes.data))
        return lo, up

    def window_mean_runs_all(self, widths=WU_LIST, copy=True):
        """kvm_window_mean_runs_all: every width of an index build in ONE pass over the series.  Returns WindowMeanRuns
        with per-width (keys, first, last) tuples, the device time of the whole pass and the run / re-walk counts.
        copy=False hands out views of the library's buffers (valid until the next window-mean call on this series)."""
        ws = np.ascontiguousarray(widths, dtype=np.int32)
        arr = (_lib.KvmRuns * len(ws))()
        self._check(self._L.kvm_window_mean_runs_all(self._h, ws.ctypes.data, len(ws), arr))
        take = (lambda a: a.copy()) if copy else (lambda a: a)
        per = []
        for r in arr:
            c = r.count
            per.append((take(np.ctypeslib.as_array(r.keys, shape=(c,))) if c else np.zeros(0),
                        take(np.ctypeslib.as_array(r.first, shape=(c,))) if c else np.zeros(0, np.int32),
                        take(np.ctypeslib.as_array(r.last, shape=(c,))) if c else np.zeros(0, np.int32)))
        return WindowMeanRuns([int(w) for w in ws], per, float(arr[0].kernel_ms), int(arr[0].n_launches),
                              int(sum(r.count for r in arr)), int(sum(r.reserved for r in arr)))


class MultiGpuSeries:
    """kvm_multi: one process driving several GPUs; the series sharded by offset range with a halo (see
    include/kvmatch_gpu.h).  verify(engine, ...) returns what the single-device call returns for the whole series."""

    def __init__(self, device_ids):
        self._L = _lib.load()
        ids = np.ascontiguousarray(device_ids, dtype=np.int32)
        h = C.c_void_p()
        rc = self._L.kvm_multi_create(C.byref(h), ids.ctypes.data, len(ids))
        if rc != 0:
            raise _lib.KvmError(rc, self._L.kvm_last_error(None).decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.kvm_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise _lib.KvmError(rc, self._L.kvm_multi_last_error(self._h).decode())

    def load(self, samples, halo: int, grid: int = 1):
        a, p = _lib.as_f64(samples)
        self._check(self._L.kvm_multi_load_series_host(self._h, p, len(a), halo, grid))
        return self

    def verify(self, engine: int, q, epsilon, intervals, shift=0, rho=0, alpha=1.0, beta=0.0) -> VerifyResult:
        q, qp = _lib.as_f64(q)
        lr, lp, K = _lib.as_intervals(intervals)
        r = _lib.KvmResult()
        self._check(self._L.kvm_multi_verify(self._h, engine, qp, len(q), epsilon, rho, alpha, beta, lp, K, shift, C.byref(r)))
        c = r.count
        return VerifyResult(_lib.copy_out(r.offsets, c, np.int32), _lib.copy_out(r.distances, c, np.float64), r.cnt_candidate,
                            r.n_verified, r.s_total, r.n_gate_pass, r.n_lb_pass, r.n_exact, r.kernel_ms, r.n_launches,
                            tuple(r.stage_ms), int(r.h2d_bytes), int(r.n_rewalked), int(r.n_chains_rewalked),
                            int(r.n_dtw_cells))


@dataclass
class WindowMeanRuns:
    widths: list
    runs: list             # per width: (keys f64, first i32, last i32)
    kernel_ms: float
    n_launches: int
    n_runs: int
    n_chains_rewalked: int


def stable_sort_by_distance(offsets, distances):
    """answers.sort(Comparator.comparing(Pair::getSecond)) — stable, K/QueryEngine.java:373."""
    order = np.argsort(distances, kind="stable")
    return offsets[order], distances[order]


@dataclass
class _EngineBase:
    series: GpuSeries
    answers: list = field(default_factory=list)   # [(offset, distance)] sorted by distance, as the reference leaves them
    last: VerifyResult | None = None

    def _positions(self, valid_positions, m, chunk):
        if valid_positions is None:  # index-free: every window start, cut into statistic chains
            s = self.series
            hi = min(s.n - m + 1, s.first + s.count - m)
            return chain_intervals(s.n, m, chunk or (EPOCH - m + 1), lo=s.first, hi=hi)
        return np.asarray(valid_positions, dtype=np.int32).reshape(-1, 2)

    def _phase1(self, fn, query_data, *args):
        """Phases 0 / 1 over index file images (kvmatch_b200/phase1.py): sets valid_positions / last_segment / phase1_ms."""
        from . import phase1
        t0 = time.perf_counter()
        indexes = [phase1.open_index(self._images[w]) for w in WU_LIST]
        valid, last_segment, _ = getattr(phase1, fn)(query_data, *args, self.series.n, indexes)
        self.phase1_ms = 1e3 * (time.perf_counter() - t0)
        self.valid_positions, self.last_segment = valid, last_segment
        if not valid:
            self.answers = []
        return valid, last_segment

    def _finish(self, statistics, res: VerifyResult, t0, t1_ms=0.0):
        t2_ms = (time.perf_counter() - t0) * 1e3
        off, dist = stable_sort_by_distance(res.offsets, res.distances)
        self.answers = list(zip(off.tolist(), dist.tolist()))
        self.last = res
        if statistics is not None:  # K/QueryEngine.java:366-371
            for slot, v in enumerate([t1_ms + t2_ms, t1_ms, t2_ms, res.cnt_candidate, res.count, 0]):
                statistics[slot].append(v)
        if self.answers:
            logger.info("Best: %s, distance: %s", self.answers[0][0], self.answers[0][1])  # :376
        logger.info("T: %d ms, T_1: %d ms, T_2: %d ms, #candidates: %d, #answers: %d", round(t1_ms + t2_ms),
                    round(t1_ms), round(t2_ms), res.cnt_candidate, res.count)  # :378
        return bool(self.answers)


class QueryEngine(_EngineBase):
    """RSM-ED — phase 2 of K/QueryEngine.java:162 (lines 341-363); query_with_index() adds phases 0 and 1
    (kvmatch_b200/phase1.py) over index file images, i.e. the reference's whole query()."""

    def query_with_index(self, statistics, query_data, epsilon, index_images):
        from . import phase1
        t0 = time.perf_counter()
        indexes = [phase1.open_index(index_images[w]) for w in WU_LIST]
        valid, last_segment, _ = phase1.phase1(query_data, epsilon, self.series.n, indexes)
        self.phase1_ms = 1e3 * (time.perf_counter() - t0)
        self.valid_positions, self.last_segment = valid, last_segment
        if not valid:
            self.answers = []
            return False
        return self.query(statistics, query_data, epsilon, valid, last_segment)

    def query(self, statistics, query_data, epsilon, valid_positions=None, last_segment=1, chunk=None):
        t0 = time.perf_counter()
        iv = self._positions(valid_positions, len(query_data), chunk)
        res = self.series.verify_ed(query_data, epsilon, iv, (last_segment - 1) * WU_LIST[0])
        return self._finish(statistics, res, t0)


class NormQueryEngine(_EngineBase):
    """cNSM-ED — phase 2 of K/NormQueryEngine.java:177 (lines 432-528); query_with_index() adds phases 0 and 1
    (phase1.phase1_norm, :190-430) over index file images."""

    def query_with_index(self, statistics, query_data, epsilon, alpha, beta, index_images):
        self._images = index_images
        valid, last_segment = self._phase1("phase1_norm", query_data, epsilon, alpha, beta)
        return bool(valid) and self.query(statistics, query_data, epsilon, alpha, beta, valid, last_segment)

    def query(self, statistics, query_data, epsilon, alpha, beta, valid_positions=None, last_segment=1, chunk=None):
        t0 = time.perf_counter()
        iv = self._positions(valid_positions, len(query_data), chunk)
        res = self.series.verify_cnsm_ed(query_data, epsilon, alpha, beta, iv, (last_segment - 1) * WU_LIST[0])
        return self._finish(statistics, res, t0)


class QueryEngineDtw(_EngineBase):
    """RSM-DTW — phase 2 of K/QueryEngineDtw.java:172 (lines 349-452); query_with_index() adds phases 0 and 1
    (phase1.phase1_dtw, :185-347)."""

    def query_with_index(self, statistics, query_data, epsilon, rho, index_images):
        self._images = index_images
        valid, last_segment = self._phase1("phase1_dtw", query_data, epsilon, rho)
        return bool(valid) and self.query(statistics, query_data, epsilon, rho, valid, last_segment)

    def query(self, statistics, query_data, epsilon, rho, valid_positions=None, last_segment=1, chunk=None):
        t0 = time.perf_counter()
        iv = self._positions(valid_positions, len(query_data), chunk)
        res = self.series.verify_dtw(query_data, epsilon, rho, iv, (last_segment - 1) * WU_LIST[0])
        return self._finish(statistics, res, t0)


class NormQueryEngineDtw(_EngineBase):
    """cNSM-DTW — phase 2 of K/NormQueryEngineDtw.java:190 (lines 457-603); query_with_index() adds phases 0 and 1
    (phase1.phase1_norm_dtw, :203-455)."""

    def query_with_index(self, statistics, query_data, epsilon, rho, alpha, beta, index_images):
        self._images = index_images
        valid, last_segment = self._phase1("phase1_norm_dtw", query_data, epsilon, rho, alpha, beta)
        return bool(valid) and self.query(statistics, query_data, epsilon, rho, alpha, beta, valid, last_segment)

    def query(self, statistics, query_data, epsilon, rho, alpha, beta, valid_positions=None, last_segment=1,
              chunk=None):
        t0 = time.perf_counter()
        iv = self._positions(valid_positions, len(query_data), chunk)
        res = self.series.verify_cnsm_dtw(query_data, epsilon, rho, alpha, beta, iv,
                                          (last_segment - 1) * WU_LIST[0])
        return self._finish(statistics, res, t0)


def rho_from_prompt(r: float, length: int) -> int:
    """The `Rho (|Q|%) = ` prompt: r <= 1 is a fraction of the query length (K/QueryEngineDtw.java:134-140)."""
    return int(np.floor(r * length)) if r <= 1 else int(np.floor(r))


class IndexBuilder:
    """Step 1 of K/IndexBuilder.java SingleIndexBuilder.run (lines 194-301) for each w in WuList."""

    def __init__(self, series: GpuSeries):
        self.series = series

    def window_mean_runs(self, w: int):
        keys, first, last, _, _ = self.series.window_mean_runs(w)
        return keys, first, last

    def build_all(self, widths=WU_LIST, shards: int = 1):
        """The whole index build for every width (K/IndexBuilder.java:98-120): ONE window-mean pass on the GPU
        (kvm_window_mean_runs_all), then step 2 and the file image per width on the host.  Returns {w: file bytes}, or with
        shards > 1 {w: [file bytes per shard]}: the per-shard layout (phase1.ShardedIndexFile), each file holding the
        window starts of one contiguous range with the single-file index's keys and positions."""
        from concurrent.futures import ThreadPoolExecutor   # widths and shards are independent; the library call drops the GIL
        from . import phase1
        res = self.series.window_mean_runs_all(widths, copy=False)   # consumed right here
        jobs = []
        for w, (k, f, l) in zip(res.widths, res.runs):
            if shards <= 1:
                jobs.append((w, (k, f, l)))
            else:
                n_windows = int(l[-1]) if len(l) else 0
                jobs.extend((w, piece) for piece in phase1.split_runs(k, f, l, phase1.shard_ranges(n_windows, shards)))
        with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), 16))) as pool:
            images = list(pool.map(lambda job: _lib.index_image_from_runs(*job[1])[0], jobs))
        if shards <= 1:
            return {w: img for (w, _), img in zip(jobs, images)}
        out = {}
        for (w, _), img in zip(jobs, images):
            out.setdefault(w, []).append(img)
        return out

    def build_rows(self, w: int):
        """{key: [(first, last), ...]} — what the reference's indexNodeMap holds after step 1."""
        keys, first, last = self.window_mean_runs(w)
        rows = {}
        for k, f, l in zip(keys.tolist(), first.tolist(), last.tolist()):
            rows.setdefault(k, []).append((f, l))
        return rows
