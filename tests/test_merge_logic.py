"""The run-merge rule of csrc/tef_cm_common.cuh::merge_equal_neighbours, restated lane by lane in numpy and checked on
random warps: per key the issued sums equal the plain sums, no lane both gives and issues, lanes without work never take
part, and a run of L equal keys leaves ceil(L / 2^rounds) issuers.  (The CUDA function itself is exercised by the GPU
parity tests; this pins the bit tricks -- run heads from a ballot, rank from clz, givers at odd rank -- on the CPU.)"""
import numpy as np
import pytest


def clz32(x):
    return 32 - int(x).bit_length()


def merge_warp(keys, vals, rounds=1):
    """keys: [32] uint32, vals: [32, NV] float64 -> (gave [32] bool, vals after the merge)."""
    lanes = np.arange(32)
    prev = np.concatenate([keys[:1], keys[:-1]])                      # __shfl_up_sync(key, 1): lane 0 reads itself
    same = (lanes > 0) & (keys == prev)
    heads = sum(1 << int(l) for l in lanes if not same[l])            # __ballot_sync(!same)
    v = vals.copy()
    gave = np.zeros(32, bool)
    if heads == 0xFFFFFFFF:
        return gave, v
    rank = np.array([l - (31 - clz32(heads & (0xFFFFFFFF >> (31 - l)))) for l in lanes])
    for r in range(rounds):
        d = 1 << r
        give = (rank & (2 * d - 1)) == d
        givers = sum(1 << int(l) for l in lanes if give[l])
        if r > 0 and givers == 0:
            break
        recv = np.array([l + d < 32 and bool((givers >> (l + d)) & 1) for l in lanes])
        down = np.concatenate([v[d:], v[32 - d:]])                    # __shfl_down_sync(v, d): out-of-range lanes read themselves
        v = np.where(recv[:, None], v + down, v)
        gave |= give
    return gave, v


@pytest.mark.parametrize("rounds", [1, 2, 3])
def test_run_merge_preserves_sums_and_halves_runs(rounds):
    rng = np.random.default_rng(7 + rounds)
    for trial in range(400):
        style = trial % 4
        if style == 0:
            keys = rng.integers(0, 6, 32)                             # few distinct keys, unsorted
        elif style == 1:
            keys = np.sort(rng.integers(0, 12, 32))                   # sorted: long runs
        elif style == 2:
            keys = np.repeat(rng.integers(0, 1000, 32), rng.integers(1, 6, 32))[:32]
        else:
            keys = np.full(32, 5)                                     # one run of 32
        keys = keys.astype(np.uint32)
        on = rng.random(32) < (0.8 if trial % 3 else 1.0)
        keys = np.where(on, keys, 0x80000000 | np.arange(32)).astype(np.uint32)   # lanes without work: unique keys
        vals = rng.integers(1, 100, (32, 8)).astype(np.float64)       # integers: sums are exact
        gave, out = merge_warp(keys, vals, rounds)
        assert not gave[~on].any()
        assert np.array_equal(out[~on], vals[~on])                    # idle lanes neither give nor receive
        issued = on & ~gave
        for k in np.unique(keys[on]):
            sel = keys == k
            assert np.array_equal(out[sel & issued].sum(0), vals[sel].sum(0)), (trial, k)
        # every maximal run of L equal keys leaves ceil(L / 2^rounds) issuers
        start = 0
        for l in range(1, 33):
            if l == 32 or keys[l] != keys[l - 1]:
                L = l - start
                if on[start]:
                    assert issued[start:l].sum() == -(-L // (1 << rounds)), (trial, start, L)
                start = l
