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
