"""The batched Metropolis-Hastings sampler of the host mirror (host/optimizers.cpp) against an independent Python restatement
of src/sir_age_structured/optimizers/MetropolisHastingsSampler.cpp:201-412 -- proposal x + s L z, mirror reflection, log-space
accept with the uniform drawn only for downhill proposals, Robbins-Monro scale, rank-1 covariance updates, periodic full
recomputation + Cholesky -- over Python versions of libstdc++'s seed_seq, mt19937, generate_canonical and the polar
normal_distribution (a fresh distribution object per proposal, as the reference constructs it).  The three libm calls of the
algorithm (log in the polar method and in the accept test, exp in the scale) go through csrc/det_math.h on the host AND on the
device (the device-resident sampler must reproduce the host's bits), restated here in tests/_det_math.py.  Proposals, accept decisions,
chain states and scales must agree BIT FOR BIT for every chain and iteration, through burn-in and through the adaptive phase.

"Seeded" is this build's extension (the reference seeds from std::random_device): chain c draws from
std::mt19937(std::seed_seq{seed, c})."""
import math

import numpy as np
import pytest

from _det_math import det_exp, det_log
from test_nuts import StdNormal
from test_pso_variants import StdMt19937


@pytest.fixture(scope="module")
def host(pkg, cuda_lib):
    import __graft_entry__ as entry
    entry.build()
    from sepaihrd_b200 import hostlib
    hostlib.load_library()
    return hostlib


def seed_seq_generate(v, n=624):
    """std::seed_seq{v...}.generate over n 32-bit words ([rand.util.seedseq])."""
    M = 0xFFFFFFFF
    b = [0x8b8b8b8b] * n
    s = len(v)
    t = 11 if n >= 623 else 7 if n >= 68 else 5 if n >= 39 else 3 if n >= 7 else (n - 1) // 2
    p = (n - t) // 2
    q = p + t
    m = max(s + 1, n)
    T = lambda x: x ^ (x >> 27)
    for k in range(m):
        r1 = (1664525 * T(b[k % n] ^ b[(k + p) % n] ^ b[(k - 1) % n])) & M
        r2 = (r1 + (s if k == 0 else (k % n + v[k - 1]) if k <= s else k % n)) & M
        b[(k + p) % n] = (b[(k + p) % n] + r1) & M
        b[(k + q) % n] = (b[(k + q) % n] + r2) & M
        b[k % n] = r2
    for k in range(m, m + n):
        r3 = (1566083941 * T((b[k % n] + b[(k + p) % n] + b[(k - 1) % n]) & M)) & M
        r4 = (r3 - k % n) & M
        b[(k + p) % n] ^= r3
        b[(k + q) % n] ^= r4
        b[k % n] = r4
    return b


class SeedSeqMt19937(StdMt19937):
    """std::mt19937 seeded from a seed_seq: the generated words ARE the state; the first draw twists."""

    def __init__(self, values):
        self.bg = np.random.MT19937()
        key = np.array(seed_seq_generate(list(values)), dtype=np.uint32)
        self.bg.state = {"bit_generator": "MT19937", "state": {"key": key, "pos": 624}}
        self.buf = np.empty(0, dtype=np.uint64)
        self.i = 0


def reflect(v, lo, hi):                  # reflectBound, SEPAIHRDParameterManager.cpp:302-313
    if lo >= hi:
        return lo
    w = hi - lo
    y = math.fmod(v - lo, 2.0 * w)
    if y < 0:
        y += 2.0 * w
    return lo + y if y <= w else hi - (y - w)


def cholesky_lower(a):
    n = len(a)
    L = [[0.0] * n for _ in range(n)]
    for j in range(n):
        d = a[j][j]
        for k in range(j):
            d -= L[j][k] * L[j][k]
        if not d > 0.0:
            return None
        L[j][j] = math.sqrt(d)
        for i in range(j + 1, n):
            s = a[i][j]
            for k in range(j):
                s -= L[i][k] * L[j][k]
            L[i][j] = s / L[j][j]
    return L


class PyChain:
    def __init__(self, x0, lp0, sigmas, lo, hi, gen, st):
        P = len(x0)
        self.P, self.lo, self.hi, self.gen, self.st = P, lo, hi, gen, st
        self.x, self.lp = list(x0), lp0
        scal = (2.38 * 2.38) / P
        self.cov = [[0.0] * P for _ in range(P)]
        for i in range(P):
            self.cov[i][i] = (sigmas[i] * sigmas[i] if sigmas[i] > 0 else 1e-6) * scal + st["eps"]
        self.L = cholesky_lower(self.cov)
        self.mean = list(x0)
        self.log_scale, self.scale = 0.0, 1.0
        self.recent = []
        self.history = [list(x0)]
        self.accepted = 0
        self.prop = None

    def adapt_covariance(self, t):
        P, st = self.P, self.st
        g = 10.0 / (t + 100.0)                                   # updateCovarianceRank1
        diff = [self.history[-1][i] - self.mean[i] for i in range(P)]
        self.mean = [self.mean[i] + g * diff[i] for i in range(P)]
        self.cov = [[(1.0 - g) * self.cov[i][j] + g * (diff[i] * diff[j]) for j in range(P)] for i in range(P)]
        if t % st["period"] == 0:
            H = self.history
            if len(H) >= P + 10:                                 # recomputeFullCovariance
                mean = [0.0] * P
                for v in H:
                    mean = [mean[i] + v[i] for i in range(P)]
                mean = [m / float(len(H)) for m in mean]
                self.mean = mean
                cov = [[0.0] * P for _ in range(P)]
                for v in H:
                    d = [v[i] - mean[i] for i in range(P)]
                    for i in range(P):
                        for j in range(P):
                            cov[i][j] += d[i] * d[j]
                inv = 1.0 / float(len(H) - 1)
                s = (2.38 * 2.38) / P
                self.cov = [[s * (cov[i][j] * inv) + (st["eps"] if i == j else 0.0) for j in range(P)] for i in range(P)]
                L = cholesky_lower(self.cov)
                if L is not None:
                    self.L = L
            L = cholesky_lower([[self.cov[i][j] + (st["eps"] if i == j else 0.0) for j in range(P)] for i in range(P)])
            if L is not None:
                self.L = L

    def propose(self, t):
        P = self.P
        if t > self.st["burn_in"]:
            self.adapt_covariance(t)
        nrm = StdNormal(det_log)                                        # a fresh distribution per proposal (.cpp:94)
        z = [nrm(self.gen) for _ in range(P)]
        step = [0.0] * P
        for j in range(P):
            for i in range(j, P):
                if self.L[i][j] != 0.0 or i == j:
                    step[i] += self.L[i][j] * z[j]
        y = [self.x[i] + self.scale * step[i] for i in range(P)]
        self.prop = [reflect(y[i], self.lo[i], self.hi[i]) for i in range(P)]
        return self.prop

    def accept(self, plp, t):
        if math.isnan(plp) or math.isinf(plp):
            plp = -1e18
        ratio = plp - self.lp
        acc = True if ratio >= 0.0 else (det_log(self.gen.uniform()) < ratio)
        if acc:
            self.x, self.lp = list(self.prop), plp
            self.accepted += 1
        # adaptGlobalScale (.cpp:104-152)
        self.recent.append(1 if acc else 0)
        if len(self.recent) > 1000:
            self.recent.pop(0)
        rate = sum(self.recent) / len(self.recent)
        tgt = self.st["target"]
        if len(self.recent) >= 1000 and rate < 0.001:
            self.log_scale -= 0.7
        elif rate < 0.02 and len(self.recent) >= 500:
            self.log_scale += min(5.0 / math.sqrt(t + 1.0), 0.3) * (0.0 - tgt)
        else:
            self.log_scale += min(1.0 / math.sqrt(t + 1.0), 0.1) * ((1.0 if acc else 0.0) - tgt)
        if self.scale <= 0.011 and 0.15 < rate < 0.30:
            self.log_scale += 0.01
        self.log_scale = max(min(self.log_scale, 2.3), -6.9)
        self.scale = det_exp(self.log_scale)
        self.history.append(list(self.x))
        return acc


def _target(mu, s):
    def f(x):
        x = np.asarray(x, dtype=float)
        acc = np.zeros(len(x))
        for k in range(len(mu)):
            z = (x[:, k] - mu[k]) / s[k]
            acc = acc + z * z
        return -0.5 * acc
    return f


def test_seed_seq_and_engine_match_libstdcxx(host):
    """First proposals of two chains started far inside wide bounds expose the raw normal draws: z = (y - x) / (scale L_ii)."""
    P = 3
    sig = np.array([0.5, 1.0, 2.0])
    pm = host.ParameterManager(sig, np.full(P, -1e6), np.full(P, 1e6), mode=1)
    mh = host.MultiChainMH(pm, dict(mcmc_iterations=5, burn_in=5, n_chains=2, seed=42))
    mh.begin(np.zeros(P), np.zeros(2))
    y = mh.propose()
    for c in range(2):
        g = SeedSeqMt19937([42, c])
        nrm = StdNormal(det_log)
        z = [nrm(g) for _ in range(P)]
        L = [math.sqrt(sig[i] * sig[i] * ((2.38 * 2.38) / P) + 1e-6) for i in range(P)]
        want = [reflect(0.0 + 1.0 * (L[i] * z[i]), -1e6, 1e6) for i in range(P)]      # reflectBound rounds through fmod even inside the bounds
        np.testing.assert_array_equal(y[c], want)


@pytest.mark.parametrize("burn_in,period,iters", [(400, 50, 120), (20, 15, 140)])
def test_chains_equal_the_python_restatement_bit_for_bit(host, burn_in, period, iters):
    P, n_chains, seed = 4, 3, 2024
    mu = np.array([0.3, -1.0, 2.0, 0.8]); s = np.array([0.4, 0.8, 0.3, 1.2])
    lo, hi = mu - np.array([0.5, 3.0, 0.4, 2.0]), mu + np.array([1.5, 1.0, 2.0, 0.9])      # tight on purpose: reflections happen
    sig = np.array([0.6, 0.9, 0.5, 1.0])
    f = _target(mu, s)
    pm = host.ParameterManager(sig, lo, hi, mode=1)
    st = dict(mcmc_iterations=iters, burn_in=burn_in, adaptation_period=period, n_chains=n_chains, seed=seed,
              regularization_epsilon=1e-6, target_acceptance_rate=0.234)
    mh = host.MultiChainMH(pm, st)
    x0 = mu + 0.1
    lp0 = float(f(x0[None])[0])
    mh.begin(x0, np.full(n_chains, lp0))
    ref = [PyChain(x0, lp0, sig, lo, hi, SeedSeqMt19937([seed, c]), dict(burn_in=burn_in, period=period, eps=1e-6, target=0.234))
           for c in range(n_chains)]
    n_reflected = n_accept = n_reject = 0
    t = 1
    while not mh.done:
        got = mh.propose()
        want = np.array([ch.propose(t) for ch in ref])
        np.testing.assert_array_equal(got, want, err_msg=f"proposals differ at iteration {t}")
        n_reflected += int(np.sum((want == lo) | (want == hi))) + 0
        lp = f(got)
        acc = mh.accept(lp)
        acc_ref = [ch.accept(float(lp[c]), t) for c, ch in enumerate(ref)]
        np.testing.assert_array_equal(acc.astype(bool), acc_ref, err_msg=f"accept decisions differ at iteration {t}")
        n_accept += sum(acc_ref); n_reject += n_chains - sum(acc_ref)
        t += 1
    x, lpc, scale, n_acc = mh.state()
    np.testing.assert_array_equal(x, np.array([ch.x for ch in ref]))
    np.testing.assert_array_equal(lpc, np.array([ch.lp for ch in ref]))
    np.testing.assert_array_equal(scale, np.array([ch.scale for ch in ref]))
    np.testing.assert_array_equal(n_acc, np.array([ch.accepted for ch in ref]))
    assert t == iters and n_accept > 20 and n_reject > 20                       # both branches of the accept rule ran
    if burn_in < iters:
        assert any(any(ch.L[i][j] != 0.0 for i in range(P) for j in range(i)) for ch in ref)   # the adaptive phase produced a dense kernel
