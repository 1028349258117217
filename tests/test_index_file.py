"""IndexBuilder step 2 + index file image: the library's host-side builder against the
oracle's restatement, an independent decoder, and hand-derived byte vectors of the row codec.  No GPU needed: the
product entry point used here (kvm_index_image_from_runs) is host-only."""
import struct

import numpy as np
import pytest

from kvmatch_b200 import _lib, datagen


def decode_image(data: bytes):
    """Reader side of K/operator/file/IndexFileOperator.java:52-62,66-84 and IndexNode.parseBytesCompact (:112-129)."""
    last_line = struct.unpack(">i", data[-4:])[0]
    offs = list(struct.unpack(">%di" % ((len(data) - last_line) // 4), data[last_line:]))
    assert offs[-1] == last_line
    rows = []
    for i in range(len(offs) - 2):
        line = data[offs[i]:offs[i + 1]]
        key = struct.unpack(">d", line[:8])[0]
        v, idx, pos = line[8:], 0, []
        while idx < len(v):
            left = struct.unpack(">i", v[idx:idx + 4])[0]
            idx += 4
            count = struct.unpack("b", v[idx:idx + 1])[0] + 128
            idx += 1
            right = left + struct.unpack("b", v[idx:idx + 1])[0] + 128
            idx += 1
            pos.append((left, right))
            for _ in range(count):
                left = right + struct.unpack("b", v[idx:idx + 1])[0] + 128
                right = left + struct.unpack("b", v[idx + 1:idx + 2])[0] + 128
                idx += 2
                pos.append((left, right))
        rows.append((key, pos))
    stat = data[offs[-2]:offs[-1]]
    table = [struct.unpack(">dii", stat[16 * i:16 * i + 16]) for i in range(len(stat) // 16)]
    return rows, table


@pytest.mark.parametrize("n,w,seed", [(200_000, 50, 1), (200_000, 25, 2), (120_125, 400, 3), (30_000, 2, 4),
                                      (100_003, 100, 5), (1_000, 200, 6)])
def test_image_matches_oracle_and_decodes(oracle, n, w, seed):
    s = datagen.generate(n, seed=seed)
    exp, rows1, rows = oracle.index_file_image(s, w)
    keys, first, last = oracle.window_mean_runs(s, w)
    got, info = _lib.index_image_from_runs(keys, first, last)
    assert got == exp                                   # byte-identical file image
    assert (info.n_rows_step1, info.n_rows, info.n_runs) == (rows1, rows, len(keys))
    dec, table = decode_image(got)
    assert len(dec) == rows == len(table)
    ks = [k for k, _ in dec]
    assert ks == sorted(ks) and len(set(ks)) == len(ks)
    # every window position of step 1 is covered exactly once, no interval longer than a byte can say
    cover = np.zeros(int(last.max()) + 2, dtype=np.int32)
    n_iv = 0
    for _, pos in dec:
        for a, b in pos:
            assert 0 <= b - a <= 255
            cover[a:b + 1] += 1
            n_iv += 1
    step1 = np.zeros_like(cover)
    for a, b in zip(first, last):
        step1[a:b + 1] += 1
    assert np.array_equal(cover, step1) and cover.max() == 1
    assert table[-1][1] == n_iv == info.n_intervals and table[-1][2] == int(cover.sum()) == info.n_offsets
    # a row holds exactly the step-1 runs whose key lies in [row key, next row key)
    bounds = ks + [float("inf")]
    row_of = np.searchsorted(np.array(ks), keys, side="right") - 1
    assert row_of.min() >= 0
    per_row = np.zeros(len(ks), dtype=np.int64)
    np.add.at(per_row, row_of, last - first + 1)
    got_per_row = np.diff([0] + [t[2] for t in table])
    assert per_row.tolist() == got_per_row.tolist()
    assert all(bounds[i] < bounds[i + 1] for i in range(len(ks)))


def test_row_codec_hand_vectors():
    """IndexNode.toBytesCompact (K/common/entity/IndexNode.java:51-96) worked by hand: one key, so step 2 is a no-op and
    the row is {key f64}{left i32}{#follow-128}{len-128}({gap-128}{len-128})*; a gap >= 256 starts a new group."""
    keys = np.array([1.5, 1.5, 1.5, 1.5])
    first = np.array([1, 5, 300, 301], dtype=np.int32)
    last = np.array([3, 5, 300, 400], dtype=np.int32)
    img, info = _lib.index_image_from_runs(keys, first, last)
    row = struct.pack(">d", 1.5) + struct.pack(">i", 1) + bytes([(1 - 128) & 255, (2 - 128) & 255, (2 - 128) & 255,
                                                                 (0 - 128) & 255])
    row += struct.pack(">i", 300) + bytes([(1 - 128) & 255, (0 - 128) & 255, (1 - 128) & 255, (99 - 128) & 255])
    stat = struct.pack(">dii", 1.5, 4, 3 + 1 + 1 + 100)
    offs = struct.pack(">iii", 0, len(row), len(row) + len(stat))
    assert img == row + stat + offs
    assert (info.n_rows_step1, info.n_rows, info.n_intervals, info.n_offsets) == (1, 1, 4, 105)


def test_merge_folds_adjacent_rows_and_resplits():
    """Step 2 (K/IndexBuilder.java:321-343) on a hand case: two keys whose runs interleave perfectly fold into one row
    (4 intervals -> 1 < 0.8 * 4), the folded run is longer than 256 positions and is re-split by addInterval
    (K/utils/IndexNodeUtils.java:82-90), and the row takes the smaller key."""
    keys = np.array([2.0, 1.0, 2.0, 1.0])
    first = np.array([1, 101, 201, 301], dtype=np.int32)
    last = np.array([100, 200, 300, 400], dtype=np.int32)
    img, info = _lib.index_image_from_runs(keys, first, last)
    dec, table = decode_image(img)
    assert (info.n_rows_step1, info.n_rows) == (2, 1)
    assert dec == [(1.0, [(1, 256), (257, 400)])]
    assert table == [(1.0, 2, 400)]


def test_rows_that_do_not_shrink_stay_apart():
    keys = np.array([2.0, 1.0, 2.0, 1.0])
    first = np.array([1, 1001, 2001, 3001], dtype=np.int32)
    last = np.array([100, 1100, 2100, 3100], dtype=np.int32)
    img, info = _lib.index_image_from_runs(keys, first, last)
    dec, table = decode_image(img)
    assert [k for k, _ in dec] == [1.0, 2.0]
    assert dec[0][1] == [(1001, 1100), (3001, 3100)] and dec[1][1] == [(1, 100), (2001, 2100)]
    assert table == [(1.0, 2, 200), (2.0, 4, 400)]     # cumulative, ByteUtils.listTripleToByteArray


def test_oracle_agrees_on_hand_cases(oracle):
    """The same three hand cases cannot be fed to the oracle directly (it starts from a series), so pin the oracle on a
    series engineered to produce the interleaved-keys case: alternating plateaus."""
    w = 2
    s = np.concatenate([np.full(375, 1.0), np.full(375, 2.0), np.full(375, 1.0), np.full(375, 2.0)])  # n % 125 == 0
    exp, rows1, rows = oracle.index_file_image(s, w)
    keys, first, last = oracle.window_mean_runs(s, w)
    got, _ = _lib.index_image_from_runs(keys, first, last)
    assert got == exp
    dec, _ = decode_image(exp)
    assert sum(b - a + 1 for _, pos in dec for a, b in pos) == len(s) - w + 1


def test_library_row_decoder_matches_independent_decoder():
    """kvm_index_row_positions (the reader's half, used by kvmatch_b200/phase1.IndexFile) against decode_image above, plus its
    error behaviour: capacity too small -> KVM_E_ARG with the count reported, truncated row -> KVM_E_RANGE."""
    import ctypes as C
    from kvmatch_b200 import phase1
    s = datagen.generate(150_000, seed=21)
    from oracle import kvm_oracle
    image = kvm_oracle.index_file_image(s, 50)[0]
    rows, table = decode_image(image)
    ix = phase1.IndexFile(image)
    assert ix.n_rows == len(rows) and [tuple(t) for t in ix.stat] == table
    for i, (key, pos) in enumerate(rows):
        k2, arr = ix.row_array(i)
        assert k2 == key and arr.dtype == np.int32 and [tuple(p) for p in arr.tolist()] == pos
    L = _lib.load()
    lo, hi = ix.offsets[0] + 8, ix.offsets[1]
    row = np.frombuffer(image, dtype=np.uint8)[lo:hi].copy()
    k = C.c_int64()
    out = np.zeros((1, 2), dtype=np.int32)
    n_pos = len(rows[0][1])
    assert n_pos > 1
    assert L.kvm_index_row_positions(row.ctypes.data, len(row), out.ctypes.data, 1, C.byref(k)) == _lib.KVM_E_ARG and k.value == n_pos
    assert tuple(out[0]) == rows[0][1][0]
    assert L.kvm_index_row_positions(row.ctypes.data, len(row) - 1, out.ctypes.data, 1, C.byref(k)) == _lib.KVM_E_RANGE
    assert L.kvm_index_row_positions(row.ctypes.data, 0, out.ctypes.data, 1, C.byref(k)) == 0 and k.value == 0
