"""The selection algorithm of ppc_select_kernel (csrc/sepaihrd_ppc.cu), restated step by step in numpy and run against a full
sort on the columns that are awkward for it: all values equal, a large tie next to a continuous part, mixed signs (keys that
differ in the top bit), NaNs (failed draws), fewer values than the gather capacity, ranks 0 and cnt - 1.  The CUDA kernel itself
is compared with numpy quantiles and with the radix-sort path in tests/test_gpu_parity.py; this test pins the ALGORITHM
(key map, common-prefix start, 8-bit narrowing of every wanted rank at once, gather + counting pick)."""
import numpy as np
import pytest

CAP = 192
M64 = (1 << 64) - 1


def _keys(x):
    b = x.view(np.uint64)
    neg = (b >> np.uint64(63)) == 1
    return np.where(neg, ~b, b | np.uint64(1 << 63))


def _value(k):
    k = int(k)
    b = (k & 0x7FFFFFFFFFFFFFFF) if (k >> 63) else (~k & M64)
    return np.array([b], dtype=np.uint64).view(np.float64)[0]


def select_order_statistics(col, ranks):
    """Values of the sorted non-NaN part of `col` at `ranks`, the way the kernel finds them."""
    x = np.asarray(col, dtype=np.float64)
    k = _keys(x[x == x])
    cnt = len(k)
    lo, hi = int(k.min()), int(k.max())
    diff = lo ^ hi
    hb = diff.bit_length()
    shift = hb
    prefix = {r: (lo >> hb) if hb < 64 else 0 for r in ranks}
    rank = {r: r for r in ranks}
    size = {r: cnt for r in ranks}
    kk = [int(v) for v in k]
    while any(size[r] > CAP for r in ranks) and shift > 0:
        nshift = max(shift - 8, 0)
        width = shift - nshift
        buckets = sorted(set(prefix.values()))
        hist = {b: [0] * 256 for b in buckets}
        for v in kk:
            h = (v >> shift) if shift < 64 else 0
            if h in hist:
                hist[h][(v >> nshift) & ((1 << width) - 1)] += 1
        for r in ranks:
            hrow = hist[prefix[r]]
            d, kr = 0, rank[r]
            while d < (1 << width) - 1 and kr >= hrow[d]:
                kr -= hrow[d]; d += 1
            prefix[r] = (prefix[r] << width) | d
            rank[r] = kr
            size[r] = hrow[d]
        shift = nshift
    out = {}
    if any(size[r] > CAP for r in ranks):                          # bits used up: the prefix IS the key
        for r in ranks:
            out[r] = _value(prefix[r])
        return out
    for r in ranks:
        cand = [v for v in kk if ((v >> shift) if shift < 64 else 0) == prefix[r]]
        assert len(cand) == size[r] <= CAP
        for c in cand:
            less = sum(1 for o in cand if o < c); eq = sum(1 for o in cand if o == c)
            if less <= rank[r] < less + eq:
                out[r] = _value(c)
    return out


def _cases():
    rng = np.random.default_rng(11)
    yield "continuous", rng.gamma(2.0, 30.0, 5000)
    yield "all equal", np.full(3000, 7.25)
    yield "all zero", np.zeros(1000)
    yield "big tie + continuous", np.concatenate([np.zeros(2500), rng.random(2500) * 1e-3, rng.random(50) * 1e6])
    yield "mixed signs", rng.normal(size=4000) * 1e3
    yield "with NaNs", np.where(rng.random(3000) < 0.1, np.nan, rng.lognormal(0, 3, 3000))
    yield "small", rng.random(17)
    yield "two values", np.array([1.0, 2.0])
    yield "one value", np.array([3.5])
    yield "denormals and huge", np.concatenate([rng.random(500) * 1e-310, rng.random(500) * 1e300, [0.0, -0.0]])


@pytest.mark.parametrize("name,col", list(_cases()), ids=[c[0] for c in _cases()])
def test_multi_select_equals_the_sorted_column(name, col):
    valid = np.sort(col[col == col])
    cnt = len(valid)
    probs = (0.0, 0.025, 0.05, 0.5, 0.95, 0.975, 1.0)
    ranks = []
    for p in probs:
        h = (cnt - 1) * p
        i0 = min(max(int(np.floor(h)), 0), cnt - 1)
        for r in (i0, min(i0 + 1, cnt - 1)):
            if r not in ranks:
                ranks.append(r)
    got = select_order_statistics(col, ranks)
    for r in ranks:
        assert got[r] == valid[r] or (got[r] == 0.0 and valid[r] == 0.0), (name, r)
