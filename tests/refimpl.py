"""A SECOND, independent CPU restatement used only to cross-check oracle/ (small inputs, pure Python).

It indexes the series directly (no circular buffers, no read grouping) and brute-forces DTW on the
full band matrix with no lower-bound cascade and no early abandoning, so agreement with the oracle
checks both the oracle's bookkeeping and the claim that the answer set does not depend on pruning.
Python floats are IEEE binary64 and CPython never contracts a*b+c, so the arithmetic order below
is the reference's (K/ = /root/reference/src/main/java/cn/edu/fudan/dsm/kvmatch/).
"""
import math

INF = 1e20  # K/utils/DtwUtils.java:24


def to_round(v):
    # K/utils/MeanIntervalUtils.java:44-61
    v = v * 10.0
    f = math.floor(v)
    r = f + (0.5 if (v - f) >= 0.5 else 0.0)
    return r * 0.1


def envelope(t, r):
    n = len(t)
    lo = [min(t[max(0, i - r):min(n, i + r + 1)]) for i in range(n)]
    up = [max(t[max(0, i - r):min(n, i + r + 1)]) for i in range(n)]
    return lo, up


def query_stats(q):
    ex = 0.0
    ex2 = 0.0
    for v in q:
        ex += v
        ex2 += v * v
    m = len(q)
    mean = ex / m
    return mean, math.sqrt(ex2 / m - mean * mean)


def interval_span(left, right, shift, m, n):
    begin = max(1, left - shift)
    end = min(n, right - shift + m - 1)
    return begin, end


def chain_stats(series, begin, end, m):
    """Yield (start_1based, mean, std) for each window of the chain that starts at `begin`."""
    ex = 0.0
    ex2 = 0.0
    for pos in range(begin, end + 1):
        d = series[pos - 1]
        ex += d
        ex2 += d * d
        if pos - begin >= m - 1:
            mean = ex / m
            var = ex2 / m - mean * mean
            std = math.sqrt(var) if var >= 0 else float("nan")
            yield pos - m + 1, mean, std
            o = series[pos - m]
            ex -= o
            ex2 -= o * o


def verify_ed(series, q, eps, intervals, shift=0):
    n, m = len(series), len(q)
    eps2 = eps * eps
    out = []
    for left, right in intervals:
        begin, end = interval_span(left, right, shift, m, n)
        for s in range(begin, end - m + 2):
            dist = 0.0
            for j in range(m):
                if not dist <= eps2:
                    break
                dd = series[s - 1 + j] - q[j]
                dist += dd * dd
            if dist <= eps2:
                out.append((s, math.sqrt(dist)))
    return out


def gate(mean, std, meanQ, stdQ, alpha, beta):
    if std != std or stdQ != stdQ:
        return False
    try:
        ratio = std / stdQ
    except ZeroDivisionError:
        return False
    return abs(mean - meanQ) <= beta and ratio <= alpha and ratio >= 1.0 / alpha


def verify_cnsm_ed(series, q, eps, alpha, beta, intervals, shift=0):
    n, m = len(series), len(q)
    eps2 = eps * eps
    meanQ, stdQ = query_stats(q)
    z = [(v - meanQ) / stdQ for v in q]
    order = sorted(range(m), key=lambda i: -abs(z[i]))  # stable, |z| descending
    out = []
    for left, right in intervals:
        begin, end = interval_span(left, right, shift, m, n)
        for s, mean, std in chain_stats(series, begin, end, m):
            if not gate(mean, std, meanQ, stdQ, alpha, beta):
                continue
            dist = 0.0
            for k in order:
                if not dist <= eps2:
                    break
                x = (series[s - 1 + k] - mean) / std
                dist += (x - z[k]) * (x - z[k])
            if dist <= eps2:
                out.append((s, math.sqrt(dist)))
    return out


def dtw_full(a, b, r):
    """Sakoe-Chiba banded DTW, squared-difference cost, out-of-band = 1e20 (no abandoning)."""
    m = len(a)
    prev = {}
    for i in range(m):
        cur = {}
        for j in range(max(0, i - r), min(m - 1, i + r) + 1):
            d = (a[i] - b[j]) * (a[i] - b[j])
            if i == 0 and j == 0:
                cur[j] = d
                continue
            y = cur.get(j - 1, INF)
            x = prev.get(j, INF)
            zz = prev.get(j - 1, INF)
            mn = x if x < y else y
            mn = mn if mn < zz else zz
            cur[j] = mn + d
        prev = cur
    return prev[m - 1]


def verify_dtw(series, q, eps, rho, intervals, shift=0):
    n, m = len(series), len(q)
    eps2 = eps * eps
    out = []
    for left, right in intervals:
        begin, end = interval_span(left, right, shift, m, n)
        for s in range(begin, end - m + 2):
            d = dtw_full(series[s - 1:s - 1 + m], q, rho)
            if d <= eps2:
                out.append((s, math.sqrt(d)))
    return out


def verify_cnsm_dtw(series, q, eps, rho, alpha, beta, intervals, shift=0):
    n, m = len(series), len(q)
    eps2 = eps * eps
    meanQ, stdQ = query_stats(q)
    z = [(v - meanQ) / stdQ for v in q]
    out = []
    for left, right in intervals:
        begin, end = interval_span(left, right, shift, m, n)
        for s, mean, std in chain_stats(series, begin, end, m):
            if not gate(mean, std, meanQ, stdQ, alpha, beta):
                continue
            zt = [(series[s - 1 + k] - mean) / std for k in range(m)]
            d = dtw_full(zt, z, rho)
            if d <= eps2:
                out.append((s, math.sqrt(d)))
    return out


def window_mean_runs(series, w, n=None, epoch=100000, max_diff=256):
    """IndexBuilder step 1 (K/IndexBuilder.java:194-301) on a series whose length is a multiple of 125."""
    n = len(series) if n is None else n
    runs = []
    stride = epoch - w + 1
    it = 0
    last_key = None
    while it * stride + w - 1 < len(series):
        start = it * stride
        stop = min(len(series), start + epoch)
        ex = 0.0
        for g in range(start, stop):
            ex += series[g]
            if g - start >= w - 1:
                loc = g - w + 2
                if loc > n:
                    break
                key = to_round(ex / w)
                if last_key is None or key != last_key or math.copysign(1, key) != math.copysign(1, last_key) \
                        or loc - runs[-1][1] == max_diff - 1:
                    runs.append([key, loc, loc])
                    last_key = key
                else:
                    runs[-1][2] = loc
                ex -= series[g - w + 1]
        it += 1
    return runs
