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


def schedule(n, tile, ctas, world=1, rank=0, max_chunks=64):
    """ee_host_pair_schedule -> (units_total, lo, hi, items[k] = (tile row, first chunk, chunks, slot), row_slot)."""
    import numpy as np
    tot, lo, hi, cnt = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    assert lib.ee_host_pair_schedule(n, tile, ctas, world, rank, max_chunks, C.byref(tot), C.byref(lo), C.byref(hi), C.byref(cnt),
                                     None, 0, None) == 0
    items = np.zeros((max(1, cnt.value), 4), dtype=np.int32)
    row_slot = np.zeros(n // tile + 1, dtype=np.int32)
    assert lib.ee_host_pair_schedule(n, tile, ctas, world, rank, max_chunks, None, None, None, None,
                                     items.ctypes.data_as(C.POINTER(C.c_int32)), cnt.value,
                                     row_slot.ctypes.data_as(C.POINTER(C.c_int32))) == 0
    return tot.value, lo.value, hi.value, items[:cnt.value], row_slot


def row_unit(ti, nch, cpt):
    return ti * nch - cpt * (ti * (ti - 1) // 2)


@pytest.mark.parametrize("n,tile,world", [(65536, 1024, 1), (65536, 1024, 8), (65536, 512, 8), (32768, 512, 2), (4096, 256, 1),
                                          (65536, 1024, 3)])
def test_pair_schedule_covers_every_unit_of_every_rank_exactly_once(n, tile, world):
    """Units = (tile row, 32-body chunk at or above the diagonal), canonical order; ranks own equal contiguous shares; a
    rank's items tile its share in order without crossing a row; guided sizes end in single chunks."""
    import numpy as np
    nch, cpt, nt = n // 32, tile // 32, n // tile
    total = row_unit(nt, nch, cpt)
    assert total == sum(nch - cpt * ti for ti in range(nt))  # closed form: row ti holds the chunks from its diagonal on
    ctas, maxc = 296, 64
    edges = []
    for rank in range(world):
        tot, lo, hi, items, row_slot = schedule(n, tile, ctas, world, rank, maxc)
        assert tot == total
        edges.append((lo, hi))
        u = lo
        for k, (ti, c0, nc, slot) in enumerate(items):
            assert slot == k and 1 <= nc <= maxc
            assert 0 <= ti < nt and ti * cpt <= c0 and c0 + nc <= nch          # inside row ti, at or above the diagonal
            assert row_unit(ti, nch, cpt) + (c0 - ti * cpt) == u                # starts where the previous item ended
            remaining = hi - u
            assert nc == 1 or nc <= remaining // ctas                           # guided: never more than one CTA-share
            u += nc
        assert u == hi
        assert list(items[-8:, 2]) == [1] * 8                                   # the queue drains in single chunks
        counts = np.bincount(items[:, 0], minlength=nt)
        assert np.array_equal(np.diff(row_slot), counts) and row_slot[0] == 0  # slots of a row are consecutive
    assert edges[0][0] == 0 and edges[-1][1] == total
    assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
    sizes = [hi - lo for lo, hi in edges]
    assert max(sizes) - min(sizes) <= 1


def test_pair_schedule_rejects_bad_arguments():
    assert lib.ee_host_pair_schedule(1000, 1024, 296, 1, 0, 64, None, None, None, None, None, 0, None) == 100  # n % tile
    assert lib.ee_host_pair_schedule(4096, 1024, 296, 2, 2, 64, None, None, None, None, None, 0, None) == 100  # rank >= world
    assert lib.ee_host_pair_schedule(4096, 1000, 296, 1, 0, 64, None, None, None, None, None, 0, None) == 100  # tile % 256


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
