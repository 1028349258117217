"""End to end like the reference's README demo: build the five KV-indexes on the GPU (one fused window-mean pass), run
phase 0 / phase 1 over them on the host (kvmatch_b200/phase1.py + kvm_intervals_*), verify the candidates on the GPU.
KV-match guarantees no false dismissals, so the answers must equal a full scan's."""
import numpy as np
import pytest

from kvmatch_b200 import datagen, phase1

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world():
    import kvmatch_b200
    s = datagen.generate(1_000_000)
    g = kvmatch_b200.GpuSeries(0)
    g.load(s)
    images = kvmatch_b200.IndexBuilder(g).build_all()
    yield s, g, images
    g.close()


def test_fused_build_equals_single_width_files(world, oracle):
    s, g, images = world
    for w in (25, 400):
        assert images[w] == oracle.index_file_image(s, w)[0]


@pytest.mark.parametrize("off,length,eps", [(123456, 8192, 10.0), (500_000, 1024, 3.0), (777_000, 512, 1.0), (40_001, 2048, 40.0)])
def test_index_pruned_query_equals_full_scan(world, oracle, off, length, eps):
    import kvmatch_b200
    s, g, images = world
    q = s[off - 1:off - 1 + length].copy()
    eng = kvmatch_b200.QueryEngine(g)
    stats = [kvmatch_b200.StatisticInfo() for _ in range(6)]
    assert eng.query_with_index(stats, q, eps, images) is True
    full = oracle.verify_ed(s, q, eps, [(1, len(s) - length + 1)])
    got = sorted(eng.answers)
    assert [o for o, _ in got] == full.offsets.tolist()
    assert sorted(d for _, d in got) == sorted(full.distances.tolist())
    assert eng.answers[0] == (off, 0.0)  # Best: <offset>, distance: 0.0
    n_cand = sum(r - l + 1 for l, r in eng.valid_positions)
    assert n_cand < 0.2 * len(s)          # the index prunes
    # the candidate list is what sortAndMergeIntervals returns: sorted, disjoint, non-adjacent
    for (l1, r1), (l2, r2) in zip(eng.valid_positions, eng.valid_positions[1:]):
        assert r1 + 1 < l2


def test_query_plan_is_a_set_of_disjoint_windows(world):
    """determineQueryPlan's DP: disjoint windows of the enabled widths inside the query.  (Like the reference's, the plan
    need not reach back to the query's first point: its dp[i][1] accepts a first window anywhere, because
    (j - 1) * Double.MAX_VALUE is 0 for j = 1, K/QueryEngine.java:466.)"""
    s, g, images = world
    indexes = [phase1.IndexFile(images[w]) for w in phase1.WU_LIST]
    for off, length in ((200_000, 1000), (5, 8192), (700_000, 512)):
        q = s[off:off + length]
        plan = phase1.determine_query_plan(q, 5.0, {w: ix.stat for w, ix in zip(phase1.WU_LIST, indexes)})
        covered = sorted((seg.order, seg.wu) for seg in plan)
        assert covered and len(covered) <= 30
        end = 0
        for order, wu in covered:
            assert wu in phase1.WU_LIST and order > end
            end = order + wu // 25 - 1
        assert end == length // 25      # the reconstruction starts from the query's last full block
        assert [seg.count for seg in plan] == sorted(seg.count for seg in plan)  # ENABLE_QUERY_REORDERING


@pytest.mark.parametrize("off,length,eps,alpha,beta", [(123456, 1024, 5.0, 1.5, 5.0), (500_000, 512, 3.0, 1.2, 5.0), (777_000, 2048, 8.0, 2.0, 10.0)])
def test_cnsm_ed_index_pruned_query_equals_full_scan(world, oracle, off, length, eps, alpha, beta):
    """The cNSM-ED engine end to end (K/NormQueryEngine.java:177-545): phases 0 / 1 with beta partitions and the variance
    filter over the GPU-built indexes, phase 2 on the GPU.  Bit-exact against the oracle on the same candidate list;
    the same answer offsets as a full scan (no false dismissals)."""
    import kvmatch_b200
    s, g, images = world
    q = s[off - 1:off - 1 + length].copy()
    eng = kvmatch_b200.NormQueryEngine(g)
    stats = [kvmatch_b200.StatisticInfo() for _ in range(6)]
    assert eng.query_with_index(stats, q, eps, alpha, beta, images) is True
    shift = (eng.last_segment - 1) * 25
    same_list = oracle.verify_cnsm_ed(s, q, eps, alpha, beta, eng.valid_positions, shift)
    assert eng.last.offsets.tolist() == same_list.offsets.tolist() and eng.last.distances.tolist() == same_list.distances.tolist()
    full = oracle.verify_cnsm_ed(s, q, eps, alpha, beta, [(1, len(s) - length + 1)])
    assert eng.last.offsets.tolist() == full.offsets.tolist() and off in full.offsets.tolist()
    assert sum(r - l + 1 for l, r in eng.valid_positions) < len(s)
    assert eng.answers[0][0] == off


@pytest.mark.parametrize("off,length,eps,rho", [(123456, 512, 12.0, 25), (640_000, 1024, 30.0, 51)])
def test_rsm_dtw_index_pruned_query_equals_full_scan(world, oracle, off, length, eps, rho):
    import kvmatch_b200
    s, g, images = world
    q = s[off - 1:off - 1 + length].copy()
    eng = kvmatch_b200.QueryEngineDtw(g)
    assert eng.query_with_index(None, q, eps, rho, images) is True
    full = oracle.verify_dtw(s, q, eps, rho, [(1, len(s) - length + 1)])
    assert eng.last.offsets.tolist() == full.offsets.tolist() and eng.last.distances.tolist() == full.distances.tolist()
    assert eng.answers[0] == (off, 0.0)
    assert sum(r - l + 1 for l, r in eng.valid_positions) < 0.5 * len(s)


@pytest.mark.parametrize("off,length,eps,rho,alpha,beta", [(123456, 512, 3.0, 25, 1.5, 5.0), (500_000, 1024, 5.0, 51, 1.2, 5.0)])
def test_cnsm_dtw_index_pruned_query_equals_full_scan(world, oracle, off, length, eps, rho, alpha, beta):
    import kvmatch_b200
    s, g, images = world
    q = s[off - 1:off - 1 + length].copy()
    eng = kvmatch_b200.NormQueryEngineDtw(g)
    assert eng.query_with_index(None, q, eps, rho, alpha, beta, images) is True
    shift = (eng.last_segment - 1) * 25
    same_list = oracle.verify_cnsm_dtw(s, q, eps, rho, alpha, beta, eng.valid_positions, shift)
    assert eng.last.offsets.tolist() == same_list.offsets.tolist() and eng.last.distances.tolist() == same_list.distances.tolist()
    full = oracle.verify_cnsm_dtw(s, q, eps, rho, alpha, beta, [(1, len(s) - length + 1)])
    assert eng.last.offsets.tolist() == full.offsets.tolist() and off in full.offsets.tolist()


def test_per_shard_index_layout_end_to_end(world, oracle):
    """build_all(shards=3): three files per width from the same fused GPU pass (the per-shard layout of SURVEY 8(f) f2);
    index-pruned queries over them return what the single-file indexes return."""
    import kvmatch_b200
    s, g, images = world
    sharded = kvmatch_b200.IndexBuilder(g).build_all(shards=3)
    assert all(len(sharded[w]) == 3 for w in phase1.WU_LIST)
    for w in (25, 400):
        one, many = phase1.IndexFile(images[w]), phase1.ShardedIndexFile(sharded[w])
        assert many.stat[-1][2] == one.stat[-1][2] == len(s) - w + 1
    off, length = 333_000, 1024          # (a shard boundary of the w = 25 index lies at 333 326: the match straddles it)
    q = s[off - 1:off - 1 + length].copy()
    eng = kvmatch_b200.NormQueryEngine(g)
    assert eng.query_with_index(None, q, 5.0, 1.5, 5.0, sharded) is True
    full = oracle.verify_cnsm_ed(s, q, 5.0, 1.5, 5.0, [(1, len(s) - length + 1)])
    assert eng.last.offsets.tolist() == full.offsets.tolist() and off in full.offsets.tolist()
    rsm = kvmatch_b200.QueryEngine(g)
    assert rsm.query_with_index(None, q, 10.0, sharded) is True
    full = oracle.verify_ed(s, q, 10.0, [(1, len(s) - length + 1)])
    assert rsm.last.offsets.tolist() == full.offsets.tolist() and rsm.last.distances.tolist() == full.distances.tolist()
    assert rsm.answers[0] == (off, 0.0)
