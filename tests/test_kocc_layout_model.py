"""Executable statement of the layout contract between the two passes of kernel (c)'s bucketed path
(csrc/tvc_aux.cu: k_occurrence_partition*_kernel -> k_occurrence_bucket_kernel), in NumPy: tile-major 16-bit keys,
transposed bucket boundaries offs[bucket][tile], the bin-less last bucket, the CTA deal derived from the bucket totals,
and pass 2 reading each segment as 32-bit words whose first / last word may hold a neighbour's key.  It pins the index
arithmetic the kernels share (it was written before the kernels and found nothing - the GPU tests in
test_gpu_hubness.py are the parity tests proper); the order of the keys inside a bucket is free, so the model sorts
stably where the kernel ranks by (thread, arrival)."""
import numpy as np
import pytest

TILE, SHIFT, WIDTH = 8192, 15, 32768


def partition(idx, idx_base, n_bins):
    total = idx.size
    n_tiles = (total + TILE - 1) // TILE
    n_buckets = (n_bins + WIDTH - 1) >> SHIFT
    keys = np.full(n_tiles * TILE, 0xABCD, np.uint16)             # stale shared memory behind a tile's keys
    offs = np.zeros((n_buckets + 1, n_tiles), np.uint16)
    totals = np.zeros(n_buckets, np.int64)
    for t in range(n_tiles):
        b = idx[t * TILE:(t + 1) * TILE].astype(np.int64) - idx_base
        ok = (b >= 0) & (b < n_bins)
        bucket = np.where(ok, b >> SHIFT, n_buckets)              # bin-less entries: last bucket, never written out
        order = np.argsort(bucket, kind="stable")
        counts = np.bincount(bucket, minlength=n_buckets + 1)
        base = np.concatenate([[0], np.cumsum(counts)])
        offs[:, t] = base[:n_buckets + 1]
        totals += counts[:n_buckets]
        n_keys = int(base[n_buckets])
        n16 = (n_keys * 2 + 15) >> 4                              # the copy-out moves whole 16-byte units
        tile_sorted = np.full(TILE, 0xDEAD, np.int64)
        tile_sorted[:b.size] = (b[order] & (WIDTH - 1))
        keys[t * TILE:t * TILE + n16 * 8] = tile_sorted[:n16 * 8].astype(np.uint16)
    return keys, offs, totals, n_tiles, n_buckets


def count(keys, offs, totals, n_tiles, n_buckets, n_bins, grid, visit_cost=128):
    out = np.zeros(n_buckets * WIDTH, np.int64)
    weight = np.where(totals > 0, totals + visit_cost * n_tiles, 0)
    n_cta = np.where(weight > 0, 1 + weight * (grid - n_buckets) // max(int(weight.sum()), 1), 0)
    first = np.concatenate([[0], np.cumsum(n_cta)])
    assert first[-1] <= grid
    keys32 = keys.view(np.uint32)
    for me in range(int(first[-1])):
        b = int(np.searchsorted(first, me, side="right")) - 1     # the last bucket whose first CTA is <= me
        while n_cta[b] == 0:
            b -= 1
        s, n = me - int(first[b]), int(n_cta[b])
        for t in range(n_tiles * s // n, n_tiles * (s + 1) // n):
            a, e = int(offs[b, t]), int(offs[b + 1, t])
            if a >= e:
                continue
            words = keys32[t * (TILE // 2) + (a >> 1): t * (TILE // 2) + ((e + 1) >> 1)]
            k = np.stack([words & 0xffff, words >> 16], axis=1).reshape(-1)
            k = k[a - 2 * (a >> 1): e - 2 * (a >> 1)]            # drop a neighbour's key at either end
            np.add.at(out, b * WIDTH + k.astype(np.int64), 1)
    assert out[n_bins:].sum() == 0
    return out[:n_bins]


@pytest.mark.parametrize("total,n_bins,base", [(3, 40000, 0), (8192, 32768, 0), (8193, 32769, 0), (20000, 100000, 0),
                                               (16389, 70000, 1000), (30001, 131072, 0), (9000, 127 * 32768, 5)])
def test_two_pass_layout_counts_like_bincount(total, n_bins, base):
    rng = np.random.default_rng(total)
    idx = (n_bins * rng.random(total) ** 3).astype(np.int64) + base
    idx[rng.random(total) < 0.05] = -1
    idx[rng.random(total) < 0.02] = n_bins + base + 7
    got = count(*partition(idx, base, n_bins), n_bins, grid=444)
    v = idx - base
    assert np.array_equal(got, np.bincount(v[(v >= 0) & (v < n_bins)], minlength=n_bins))
