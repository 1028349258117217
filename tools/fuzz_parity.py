"""Randomised differential test: every entry point against the oracle on random series, lengths, thresholds and interval
structures (chain grids, pruned sets, shifts).  usage: fuzz_parity.py [iterations] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, kvmatch_b200
from kvmatch_b200 import datagen
from oracle import kvm_oracle as o

bad = 0


exempt = 0


def same(a, e, what, ctx, eps=None):
    """Bit-exact offsets and distances, except (north_star's exemption) answers whose REFERENCE distance is within 1e-9
    relative of epsilon: DtwUtils.dtw's early abandon returns min_cost + cb[...], which can equal bsf = eps^2 exactly (on
    quantised data it does) and is then accepted with distance eps although the true band DTW is larger
    (K/utils/DtwUtils.java:324-326, K/QueryEngineDtw.java:444).  Those are listed, not counted."""
    global bad, exempt
    ao, ad, eo, ed = a.offsets.tolist(), a.distances.tolist(), e.offsets.tolist(), e.distances.tolist()
    if eps is not None and (ao != eo or ad != ed):
        band = {off for off, dd in zip(eo, ed) if abs(dd - eps) <= 1e-9 * eps}
        if band:
            exempt += len(band)
            print("EXEMPT (reference distance == eps)", what, ctx, sorted(band), flush=True)
            keep_e = [(x, y) for x, y in zip(eo, ed) if x not in band]
            keep_a = [(x, y) for x, y in zip(ao, ad) if x not in band]
            eo, ed = [x for x, _ in keep_e], [y for _, y in keep_e]
            ao, ad = [x for x, _ in keep_a], [y for _, y in keep_a]
    ok = ao == eo and ad == ed and a.n_verified == e.n_verified
    if not ok:
        bad += 1
        print("MISMATCH", what, ctx, a.count, e.count, flush=True)


def run(iters=40, seed=2026, verbose=True):
    """Returns (#mismatches, #answers in the eps-tie exemption)."""
    global bad, exempt
    bad = exempt = 0
    rng = np.random.default_rng(seed)
    g = kvmatch_b200.GpuSeries(0)
    t0 = time.time()
    for it in range(iters):
        n = int(rng.integers(60_000, 400_000))
        s = datagen.generate(n, seed=int(rng.integers(1, 1 << 30)))
        if rng.random() < 0.2:
            s = np.round(s, 1)          # ties and exact repeats
        g.load(s)
        m = int(rng.choice([16, 25, 50, 100, 127, 128, 256, 300, 512, 1000, 1024, 2048]))
        m = min(m, n // 8)
        shift = int(rng.choice([0, 0, 25, 50, 75]))
        kind = rng.integers(0, 3)
        if kind == 0:
            iv = datagen.chain_intervals(n, m, int(rng.integers(max(64, m // 2), 20_000)))
            shift = 0
        elif kind == 1:
            K = int(rng.integers(1, 3000))
            stride = max(2, (n - m - 400) // K)
            lefts = np.sort(rng.choice(np.arange(1 + shift, n - m - 300, stride), size=min(K, (n - m - 301 - shift) // stride), replace=False))
            iv = np.stack([lefts, lefts + rng.integers(0, min(stride - 1, 250), len(lefts))], axis=1).astype(np.int32)
        else:
            lo = int(rng.integers(1 + shift, n // 2))
            hi = int(rng.integers(lo, n - m + 1))
            iv = datagen.chain_intervals(n, m, int(rng.integers(500, 50_000)), lo, hi)
        off = int(rng.integers(0, n - m))
        q = s[off:off + m].copy()
        if rng.random() < 0.5:
            q = q + rng.normal(0, 0.05 * (np.std(q) + 1e-3), m)
        eps = float(rng.choice([0.5, 2.0, 5.0, 20.0]))
        alpha, beta = float(rng.choice([1.1, 1.5, 3.0])), float(rng.choice([0.5, 5.0, 100.0]))
        rho = int(max(1, min(100, rng.choice([0.02, 0.05, 0.1]) * m)))
        ctx = f"it={it} n={n} m={m} K={len(iv)} kind={kind} shift={shift} eps={eps} a={alpha} b={beta} rho={rho}"
        got_c, exp_c = g.verify_cnsm_ed(q, eps, alpha, beta, iv, shift), o.verify_cnsm_ed(s, q, eps, alpha, beta, iv, shift)
        same(got_c, exp_c, "cnsm-ed", ctx)
        if got_c.n_gate_pass != exp_c.n_gate_pass:
            bad += 1
            print("MISMATCH gate count", ctx, got_c.n_gate_pass, exp_c.n_gate_pass, flush=True)
        if it % 3 == 0:  # the relay walker (every chain walked exactly) must agree with the stream
            from kvmatch_b200 import _lib
            g.set_option(_lib.KVM_OPT_CNSM_PATH, _lib.KVM_CNSM_RELAY)
            same(g.verify_cnsm_ed(q, eps, alpha, beta, iv, shift), exp_c, "cnsm-ed (relay)", ctx)
            g.set_option(_lib.KVM_OPT_CNSM_PATH, _lib.KVM_CNSM_STREAM)
        sc = float(np.sqrt(m))
        same(g.verify_ed(q, eps * sc, iv, shift), o.verify_ed(s, q, eps * sc, iv, shift), "ed", ctx)
        if m >= 16 and it % 2 == 0:
            same(g.verify_cnsm_dtw(q, eps / 2, rho, alpha, beta, iv, shift), o.verify_cnsm_dtw(s, q, eps / 2, rho, alpha, beta, iv, shift), "cnsm-dtw", ctx, eps / 2)
            same(g.verify_dtw(q, eps * sc / 2, rho, iv, shift), o.verify_dtw(s, q, eps * sc / 2, rho, iv, shift), "dtw", ctx, eps * sc / 2)
        if shift == 0:
            qs = np.stack([q, s[(off * 7) % (n - m):(off * 7) % (n - m) + m], np.roll(q, 3)])
            for qq, r in zip(qs, g.verify_cnsm_ed_batch(qs, eps, alpha, beta, iv)):
                same(r, o.verify_cnsm_ed(s, qq, eps, alpha, beta, iv), "query-set", ctx)
        if it % 5 == 0:
            w = int(rng.choice([25, 50, 100, 200, 400]))
            k, f, l, _, _ = g.window_mean_runs(w)
            ek, ef, el = o.window_mean_runs(s, w)
            if not (f.tolist() == ef.tolist() and l.tolist() == el.tolist() and k.view(np.int64).tolist() == ek.view(np.int64).tolist()):
                bad += 1
                print("MISMATCH runs", ctx, w, flush=True)
            ws = (25, 50, 100, 200, 400)
            fused = g.window_mean_runs_all(ws)
            for w2, (k2, f2, l2) in zip(ws, fused.runs):
                ek2, ef2, el2 = o.window_mean_runs(s, w2)
                if not (f2.tolist() == ef2.tolist() and l2.tolist() == el2.tolist() and k2.view(np.int64).tolist() == ek2.view(np.int64).tolist()):
                    bad += 1
                    print("MISMATCH fused runs", ctx, w2, flush=True)
            import tempfile
            with tempfile.TemporaryDirectory() as td:
                path = os.path.join(td, "index")
                g.build_index_file(w, path)
                if open(path, "rb").read() != o.index_file_image(s, w)[0]:
                    bad += 1
                    print("MISMATCH index file", ctx, w, flush=True)
        if verbose and it % 10 == 9:
            print(f"{it + 1} iterations, {bad} mismatches, {time.time() - t0:.0f}s", flush=True)
    g.close()
    return bad, exempt


if __name__ == "__main__":
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    run(iters, int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
    print("FUZZ", "FAILED" if bad else "ok", f"({iters} iterations, {exempt} answers in the reference's eps-tie exemption)")
