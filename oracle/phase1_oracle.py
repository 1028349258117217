"""Pure-Python restatement of the reference's phase-1 interval algebra (K/QueryEngine.java:264-308, 593-693).

TEST INFRASTRUCTURE ONLY (the checker of kvm_intervals_*, csrc/phase1.hpp).  PARITY UNPINNED like the rest of oracle/:
the reference ships no vectors for these functions and cannot run here.  Intervals are (left, right, epsilon) tuples."""


def sort_but_not_merge(ivs, count=False):  # :593-622 / :624-662
    if len(ivs) <= 1:
        out = list(ivs)
        return (out, len(ivs), (ivs[0][1] - ivs[0][0] + 1) if ivs else 0) if count else out
    ivs = sorted(ivs, key=lambda t: t[0])  # List.sort: stable
    start, end, eps = ivs[0]
    out, disjoint, offsets = [], len(ivs), 0
    for left, right, e in ivs[1:]:
        if left - 1 <= end:
            disjoint -= 1
        if left - 1 < end or (left - 1 == end and abs(e - eps) < 1):
            end = max(right, end)
            eps = min(e, eps)
        else:
            out.append((start, end, eps))
            offsets += end - start + 1
            start, end, eps = left, right, e
    out.append((start, end, eps))
    offsets += end - start + 1
    return (out, disjoint, offsets) if count else out


def sort_and_merge(ivs):  # :664-693
    if len(ivs) <= 1:
        return list(ivs)
    ivs = sorted(ivs, key=lambda t: t[0])
    start, end, eps = ivs[0]
    out = []
    for left, right, e in ivs[1:]:
        if left - 1 <= end:
            end = max(right, end)
            eps = min(e, eps)
        else:
            out.append((start, end, eps))
            start, end, eps = left, right, e
    out.append((start, end, eps))
    return out


def intersect(cs, csi, eps2, delta_w):  # :282-308
    out, min_eps = [], 1.7976931348623157e308
    i1 = i2 = 0
    while i1 < len(cs) and i2 < len(csi):
        a, b = cs[i1], csi[i2]
        if a[1] < b[0]:
            i1 += 1
        elif b[1] < a[0]:
            i2 += 1
        else:
            s = a[2] + b[2]
            if a[1] < b[1]:
                if s <= eps2:
                    out.append((max(a[0], b[0]) + delta_w, a[1] + delta_w, s))
                    min_eps = min(min_eps, s)
                i1 += 1
            else:
                if s <= eps2:
                    out.append((max(a[0], b[0]) + delta_w, b[1] + delta_w, s))
                    min_eps = min(min_eps, s)
                i2 += 1
    return out, min_eps


def first_segment(pos, order, w0, length, n, delta_w):  # :264-280
    out, min_eps = [], 1.7976931348623157e308
    sh = (order - 1) * w0
    for left, right, e in pos:
        if right - sh + length - 1 > n:
            if left - sh + length - 1 <= n:
                out.append((left + delta_w, n - length + 1 + sh + delta_w, e))
        elif left - sh < 1:
            if right - sh >= 1:
                out.append((1 + sh + delta_w, right + delta_w, e))
        else:
            out.append((left + delta_w, right + delta_w, e))
        min_eps = min(min_eps, e)
    return out, min_eps


# ---------------------------------------------------------------- cNSM variants (K/NormQueryEngine.java:313-397, 788-896;
# K/NormQueryEngineDtw.java:326-425, 926-1046).  Intervals are (left, right, ex_lower, ex2_lower, ex_upper, ex2_upper,
# beta_partitions) tuples; the ED engine leaves the upper sums at zero.
import struct as _struct


def _same_double(a, b):  # Double.compare(a, b) == 0
    return _struct.pack(">d", a) == _struct.pack(">d", b) or (a != a and b != b)


def _java_min(a, b):  # Math.min: NaN wins, then -0.0 < +0.0
    if a != a or b != b:
        return float("nan")
    if a == 0.0 and b == 0.0:
        return a if _struct.pack(">d", a)[0] & 0x80 else b
    return a if a <= b else b


def norm_sort_but_not_merge(ivs, count=False):  # :788-823 / :825-869
    if len(ivs) <= 1:
        out = list(ivs)
        return (out, len(ivs), (ivs[0][1] - ivs[0][0] + 1) if ivs else 0) if count else out
    ivs = sorted(ivs, key=lambda t: t[0])
    start, end, ex, ex2, exu, ex2u, bp = ivs[0]
    out, disjoint, offsets = [], len(ivs), 0
    for left, right, cex, cex2, cexu, cex2u, cbp in ivs[1:]:
        if left - 1 <= end:
            disjoint -= 1
        if left - 1 < end or (left - 1 == end and _same_double(cex, ex) and _same_double(cex2, ex2)):
            end = max(right, end)
            ex, ex2 = _java_min(cex, ex), _java_min(cex2, ex2)
            exu, ex2u = _java_min(cexu, exu), _java_min(cex2u, ex2u)   # Math.min for the upper sums too (Dtw :949-950)
            bp |= cbp
        else:
            out.append((start, end, ex, ex2, exu, ex2u, bp))
            offsets += end - start + 1
            start, end, ex, ex2, exu, ex2u, bp = left, right, cex, cex2, cexu, cex2u, cbp
    out.append((start, end, ex, ex2, exu, ex2u, bp))
    offsets += end - start + 1
    return (out, disjoint, offsets) if count else out


def norm_sort_and_merge(ivs):  # :871-896 (new NormInterval(start, end): sums and partitions are zero)
    if len(ivs) <= 1:
        return list(ivs)
    ivs = sorted(ivs, key=lambda t: t[0])
    start, end = ivs[0][0], ivs[0][1]
    out = []
    for left, right, *_ in ivs[1:]:
        if left - 1 <= end:
            end = max(right, end)
        else:
            out.append((start, end, 0.0, 0.0, 0.0, 0.0, 0))
            start, end = left, right
    out.append((start, end, 0.0, 0.0, 0.0, 0.0, 0))
    return out


def norm_intersect(cs, csi, pre_length, w0, query_length, mean_q, std_q, alpha, beta, delta_w, dtw=False):  # :333-397 / Dtw :349-425
    out = []
    i1 = i2 = 0
    while i1 < len(cs) and i2 < len(csi):
        a, b = cs[i1], csi[i2]
        if a[1] < b[0]:
            i1 += 1
            continue
        if b[1] < a[0]:
            i2 += 1
            continue
        common = a[6] & b[6]
        if common == 0:
            if a[1] < b[1]:
                i1 += 1
            else:
                i2 += 1
            continue
        sum_ex, sum_ex2 = a[2] + b[2], a[3] + b[3]
        sum_exu = sum_ex2u = 0.0
        mean = sum_ex / pre_length
        std2 = 0.0
        if mean > mean_q + beta:
            new_value = mean_q + beta - (mean - mean_q - beta) * pre_length * w0 / (query_length - pre_length * 1.0 * w0)
            mean = mean_q + beta
            std2 = (sum_ex2 * w0 + (query_length - pre_length * w0) * new_value * new_value) / query_length - mean * mean
        if dtw:
            sum_exu, sum_ex2u = a[4] + b[4], a[5] + b[5]
            mean_u = sum_exu / pre_length
            if mean_u < mean_q - beta:
                new_value = mean_q - beta - (mean_q - beta - mean_u) * pre_length * w0 / (query_length - pre_length * 1.0 * w0)
                mean_u = mean_q - beta
                std2 = (sum_ex2u * w0 + (query_length - pre_length * w0) * new_value * new_value) / query_length - mean_u * mean_u
        limit = alpha * alpha * std_q * std_q
        keep = (limit != limit) or (std2 == std2 and std2 <= limit)   # Double.compare(std2, limit) <= 0 (no signed-zero case arises: limit >= +0.0)
        right = a[1] if a[1] < b[1] else b[1]
        if keep:
            out.append((max(a[0], b[0]) + delta_w, right + delta_w, sum_ex, sum_ex2, sum_exu, sum_ex2u, common))
        if a[1] < b[1]:
            i1 += 1
        else:
            i2 += 1
    return out


def norm_first_segment(pos, order, w0, length, n, delta_w):  # :313-332 / Dtw :326-348
    out = []
    sh = (order - 1) * w0
    for left, right, *rest in pos:
        if right - sh + length - 1 > n:
            if left - sh + length - 1 <= n:
                out.append((left + delta_w, n - length + 1 + sh + delta_w, *rest))
        elif left - sh < 1:
            if right - sh >= 1:
                out.append((1 + sh + delta_w, right + delta_w, *rest))
        else:
            out.append((left + delta_w, right + delta_w, *rest))
    return out
