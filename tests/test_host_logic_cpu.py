"""Host-side logic of the engine that needs no GPU: the solout sampling schedule (exact f64 rule of the reference) and
the pair-symmetric kernel's work-item list (tiles x superchunks above the diagonal, split across ranks)."""
import ctypes as C

import pytest

import ephemeris_explorer_b200 as ee

lib = ee.lib


def stride_ref(delta, period, limit=1 << 16):
    acc = 0.0
    for c in range(1, limit):
        acc = acc + delta
        if acc == period:
            return c
        if abs(acc) > abs(period):
            return 0
    return 0


@pytest.mark.parametrize("delta,count", [(600.0, 1), (600.0, 450), (21600.0, 12), (2.0 ** -10, 7), (0.1, 3), (0.1, 10), (0.7, 3), (1e-3, 5)])
def test_sampling_stride_follows_the_reference_equality_rule(delta, count):
    period = delta * float(count)  # load/mod.rs:325
    got = lib.ee_host_sampling_stride(delta, period)
    assert got == stride_ref(delta, period)
    if delta in (600.0, 21600.0, 2.0 ** -10):
        assert got == count  # exactly representable steps sample on schedule


def test_sampling_stride_reports_never():
    # 0.1 accumulated ten times is 0.9999999999999999, not 1.0: the reference would never sample such a body
    assert stride_ref(0.1, 1.0) == 0 and lib.ee_host_sampling_stride(0.1, 1.0) == 0
    assert lib.ee_host_sampling_stride(1.0, 0.5) == 0


def items(n, js, world=1, rank=0):
    t, lo, hi = C.c_int64(), C.c_int64(), C.c_int64()
    assert lib.ee_host_pair_items(n, js, world, rank, C.byref(t), C.byref(lo), C.byref(hi)) == 0
    return t.value, lo.value, hi.value


def decode(n, js, item):
    ti, sj = C.c_int64(), C.c_int64()
    assert lib.ee_host_pair_item_decode(n, js, item, C.byref(ti), C.byref(sj)) == 0
    return ti.value, sj.value


@pytest.mark.parametrize("n,js", [(32768, 512), (65536, 512), (65536, 256), (4096, 128)])
def test_pair_items_cover_every_block_above_the_diagonal_once(n, js):
    total, lo, hi = items(n, js)
    assert (lo, hi) == (0, total)
    ratio = 1024 // js
    seen = set()
    step = max(1, total // 3000)  # sample densely, plus the ends
    for item in list(range(0, total, step)) + [total - 1]:
        ti, sj = decode(n, js, item)
        assert 0 <= ti < n // 1024 and ratio * ti <= sj < n // js  # superchunk not entirely below the tile
        seen.add((ti, sj))
    # closed form: tile ti has n/js - ratio*ti items
    assert total == sum(n // js - ratio * ti for ti in range(n // 1024))
    assert len(seen) == len(list(range(0, total, step))) + (0 if (total - 1) % step == 0 else 1)
    # consecutive indices walk a tile's superchunks in order, then the next tile
    assert decode(n, js, 0) == (0, 0) and decode(n, js, n // js - 1) == (0, n // js - 1) and decode(n, js, n // js) == (1, ratio)
    with pytest.raises(AssertionError):
        decode(n, js, total)


def test_pair_items_rank_ranges_partition_the_list():
    n, js = 65536, 256
    total, _, _ = items(n, js)
    for world in (2, 4, 8):
        edges = [items(n, js, world, r)[1:] for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == total
        assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
        sizes = [hi - lo for lo, hi in edges]
        assert max(sizes) - min(sizes) <= 1
    assert lib.ee_host_pair_items(1000, 512, 1, 0, None, None, None) == 100  # n must be a multiple of the tile


def test_padding_an_ordered_sum_with_positive_zeros_never_changes_its_bits():
    """k_small_steps adds every partner row and replaces the unwanted ones by +0.0 instead of branching (ee_small.cu):
    a left-to-right sum that starts at +0.0 can never hold -0.0, and x + (+0.0) == x for every other x, so the padded
    sum is bit-identical to the plain one -- including sums of -0.0 terms, cancellations to zero, infinities and NaN."""
    import struct
    import numpy as np
    rng = np.random.default_rng(7)

    def bits(x):
        return struct.pack("<d", float(x))

    specials = [0.0, -0.0, 1e-320, -1e-320, 1.0, -1.0, 1e308, -1e308, np.inf, -np.inf]
    for trial in range(3000):
        k = int(rng.integers(0, 12))
        if trial % 3 == 0:
            terms = [specials[int(i)] for i in rng.integers(0, len(specials), k)]
        elif trial % 3 == 1:
            base = rng.normal(size=k)
            terms = list(base) + list(-base)  # exact cancellations to +0.0
            rng.shuffle(terms)
        else:
            terms = list(rng.normal(size=k) * 10.0 ** rng.integers(-300, 300))
        plain = np.float64(0.0)
        for v in terms:
            plain = plain + np.float64(v)
        padded = np.float64(0.0)
        with np.errstate(invalid="ignore", over="ignore"):
            for v in terms:
                for _ in range(int(rng.integers(0, 3))):
                    padded = padded + np.float64(0.0)
                padded = padded + np.float64(v)
            padded = padded + np.float64(0.0)
        assert bits(plain) == bits(padded) or (np.isnan(plain) and np.isnan(padded)), (terms, plain, padded)
