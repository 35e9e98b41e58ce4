"""SimulationCache of the host mirror (host/epidemic_host.cpp) against a Python restatement of
src/sir_age_structured/caching/SimulationCache.cpp: the 1e-8-quantised hash, linear probing, LFU eviction with LRU tie-break
-- including what the reference's probing does after an eviction punched a hole into a probe chain."""
import numpy as np
import pytest

M64 = (1 << 64) - 1


@pytest.fixture(scope="module")
def host(pkg, cuda_lib):
    import __graft_entry__ as entry
    entry.build()
    from sepaihrd_b200 import hostlib
    hostlib.load_library()
    return hostlib


def _mix(k):
    k ^= k >> 33; k = (k * 0xff51afd7ed558ccd) & M64
    k ^= k >> 33; k = (k * 0xc4ceb9fe1a85ec53) & M64
    k ^= k >> 33
    return k


def _hash(x):                                                      # computeHash, .cpp:35-52
    seed = 0
    for v in x:
        q = int(float(v) * 1e8 + 0.5) & M64                        # static_cast<long long> truncates toward zero; then size_t
        seed ^= (_mix(q) + 0x9e3779b9 + ((seed << 6) & M64) + (seed >> 2)) & M64
    return seed


class PyCache:
    """The reference's slot mechanics, statement by statement (.cpp:58-104, 212-252)."""

    def __init__(self, cap):
        self.cap, self.count, self.tick = cap, 0, 0
        self.keys = [0] * cap; self.vals = [0.0] * cap; self.freq = [0] * cap; self.time = [0] * cap; self.occ = [False] * cap

    def find(self, key):
        idx = start = key % self.cap
        while self.occ[idx]:
            if self.keys[idx] == key:
                return idx
            idx = (idx + 1) % self.cap
            if idx == start:
                break
        return -1

    def get(self, key):
        i = self.find(key)
        if i < 0:
            return None
        self.freq[i] += 1; self.tick += 1; self.time[i] = self.tick
        return self.vals[i]

    def store(self, key, value):
        i = self.find(key)
        if i >= 0:
            self.vals[i] = value; self.freq[i] += 1; self.tick += 1; self.time[i] = self.tick
            return
        if self.count >= self.cap:
            victim, mf, mt = 0, 1 << 32, 1 << 32
            for j in range(self.cap):
                if self.occ[j] and (self.freq[j] < mf or (self.freq[j] == mf and self.time[j] < mt)):
                    victim, mf, mt = j, self.freq[j], self.time[j]
            self.occ[victim] = False; self.count -= 1
        at = key % self.cap
        while self.occ[at]:
            at = (at + 1) % self.cap
        self.keys[at] = key; self.vals[at] = value; self.freq[at] = 1; self.tick += 1; self.time[at] = self.tick; self.occ[at] = True
        self.count += 1


def test_hash_matches_the_reference_formula_and_quantises_at_1e_minus_8(host, problem):
    c = host.Cache(16)
    rng = np.random.default_rng(0)
    for x in (problem.base_params(), rng.normal(size=7) * 100, -rng.random(5), np.zeros(3), np.array([1e9, -1e9, 0.123456789])):
        assert c.hash(x) == _hash(x)
    x = problem.base_params()
    assert c.hash(x) != c.hash(x + 1e-7)                            # one quantum apart
    y = x.copy(); y[3] += 2e-10                                     # same 1e-8 cell (quirk Q6: such vectors share an entry) ...
    q = lambda v: int(v * 1e8 + 0.5)
    if q(x[3]) == q(y[3]):
        assert c.hash(x) == c.hash(y)
    c.set_vector(x, -123.5)
    assert c.get_vector(x) == -123.5 and len(c) == 1
    assert c.get_vector(x + 1.0) is None


def test_lfu_eviction_with_lru_tie_break_follows_the_python_restatement(host):
    rng = np.random.default_rng(5)
    for cap in (1, 3, 8):
        c, ref = host.Cache(cap), PyCache(cap)
        keys = [int(k) for k in rng.integers(0, 1 << 62, size=3 * cap + 2)]
        keys += [keys[0] + cap, keys[0] + 2 * cap]                  # same home slot as keys[0]: probe chains
        for step in range(600):
            k = keys[int(rng.integers(len(keys)))]
            if rng.random() < 0.5:
                v = float(step)
                c.store(k, v); ref.store(k, v)
            else:
                assert c.get(k) == ref.get(k), (cap, step)
            assert len(c) == ref.count
        st = c.stats()
        assert st["get_calls"] + st["store_calls"] == 600 and 0 < st["hits"] <= st["get_calls"]
        c.clear()
        assert len(c) == 0 and all(c.get(k) is None for k in keys)


def test_least_frequently_used_goes_first_then_least_recently_used(host):
    c = host.Cache(3)
    for k in (10, 11, 12):
        c.store(k, float(k))
    assert c.get(10) == 10.0 and c.get(10) == 10.0 and c.get(12) == 12.0      # frequencies: 10 -> 3, 11 -> 1, 12 -> 2
    c.store(13, 13.0)                                               # evicts 11 (lowest frequency)
    assert c.get(11) is None and len(c) == 3
    c.store(14, 14.0)                                               # 13 has frequency 1 and is the oldest of the frequency-1 entries
    assert c.get(13) is None and c.get(14) == 14.0 and c.get(10) == 10.0 and c.get(12) == 12.0


def test_constructor_and_string_keys(host):
    with pytest.raises(host.HostError):
        host.Cache(0)                                               # "SimulationCache: max_size must be > 0."


def test_batches_go_through_the_cache_like_a_sequence_of_calculate_calls(host):
    """evaluateThroughCache (the body of SEPAIHRDObjectiveFunction::calculateBatch) on the CPU with a counting evaluator: probe
    per row, ONE evaluation call for the distinct misses, repeats counted as hits, failures that calculate() returns before its
    store are not stored, batches beyond the capacity evaluated whole."""
    c = host.Cache(32)
    calls = []

    def ev(x):
        calls.append(np.array(x))
        return -np.sum(x * x, axis=1)
    rows = np.array([[float(i), 0.5 * i, -1.0] for i in range(10)])
    out = c.batch(rows[[1, 2, 1, 3, 2]], ev)
    np.testing.assert_array_equal(out, ev(rows[[1, 2, 1, 3, 2]])); calls.pop()
    assert len(calls) == 1 and len(calls[0]) == 3                   # rows 1, 2, 3 once, as one batch
    np.testing.assert_array_equal(calls[0], rows[[1, 2, 3]])
    assert len(c) == 3 and c.stats() == dict(get_calls=5 + 2, hits=2, store_calls=3)
    out = c.batch(rows[[3, 4, 1]], ev)                              # two hits, one miss
    assert len(calls) == 2 and len(calls[1]) == 1 and len(c) == 4
    np.testing.assert_array_equal(out, -np.sum(rows[[3, 4, 1]] ** 2, axis=1))
    out = c.batch(rows[[1, 2, 3, 4]], ev)                           # all hits: the evaluator is not called
    assert len(calls) == 2
    # status words: row 5 fails before calculate()'s store (S overflow = 1), row 6 is the non-finite sentinel (4, stored)
    st = np.zeros(10, dtype=np.uint32); st[5] = 1; st[6] = 4
    c.batch(rows[[5, 6, 7]], ev, status_of_row=st)
    assert len(c) == 4 + 2 and c.get(c.hash(rows[5])) is None and c.get(c.hash(rows[6])) == -np.sum(rows[6] ** 2)
    # a batch larger than the capacity is evaluated whole and leaves the cache alone
    big = np.array([[100.0 + i, 1.0, 2.0] for i in range(40)])
    n_before, s_before = len(c), c.stats()
    out = c.batch(big, ev)
    assert len(calls[-1]) == 40 and len(c) == n_before and c.stats() == s_before
    np.testing.assert_array_equal(out, -np.sum(big * big, axis=1))
