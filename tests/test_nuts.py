"""CPU tests of the NUTS sampler and the forward-difference gradient objective of the host mirror (host/optimizers.cpp,
host/epidemic_host.cpp) -- src/model/optimizers/NUTSSampler.cpp, src/model/objectives/SEPAIHRDGradientObjectiveFunction.cpp.

The reference has no test of either.  Anchor: an independent Python restatement of NUTSSampler::optimize / findReasonableEpsilon /
leapfrog / buildTree / checkNoUTurn over Python versions of libstdc++'s mt19937, generate_canonical, normal (Marsaglia polar
with its saved second value), exponential and uniform_int distributions: the C++ sampler must return the same samples bit for
bit.  Plus a statistical check that the chain samples a Gaussian."""
import math

import numpy as np
import pytest

from test_pso_variants import StdMt19937


@pytest.fixture(scope="module")
def host(pkg, cuda_lib):
    import __graft_entry__ as entry
    entry.build()
    from sepaihrd_b200 import hostlib
    hostlib.load_library()
    return hostlib


class StdNormal:
    """std::normal_distribution<double>(0, 1) of libstdc++: polar method, the second value of a pair is kept for the next call."""

    def __init__(self, log=math.log):
        self.saved = None
        self.log = log            # the Metropolis-Hastings sampler draws with csrc/det_math.h's logarithm (tests/_det_math.py)

    def __call__(self, g: StdMt19937) -> float:
        if self.saved is not None:
            v, self.saved = self.saved, None
            return v
        while True:
            x = 2.0 * g.uniform() - 1.0
            y = 2.0 * g.uniform() - 1.0
            r2 = x * x + y * y
            if not (r2 > 1.0 or r2 == 0.0):
                break
        mult = math.sqrt(-2.0 * self.log(r2) / r2)
        self.saved = x * mult
        return y * mult


def _exponential(g):                 # std::exponential_distribution<>(1.0)
    return -math.log(1.0 - g.uniform()) / 1.0


def _coin(g):                        # std::uniform_int_distribution<>(0, 1) on a 32-bit engine (Lemire's method): the top bit
    return g.raw() >> 31


def _dot(a, b):
    s = 0.0
    for x, y in zip(a, b):
        s += x * y
    return s


class PyNuts:
    def __init__(self, f, sigmas, lb, ub, seed, iterations, window, delta, depth, eps_fd=1e-4):
        self.f, self.sigmas, self.lb, self.ub = f, sigmas, lb, ub
        self.g = StdMt19937(seed)
        self.iterations, self.window, self.delta, self.depth, self.eps_fd = iterations, window, delta, depth, eps_fd
        self.n_grad = 0
        self.memo = None

    def constrain(self, x):
        return np.minimum(np.maximum(x, self.lb), self.ub)

    def grad(self, theta):           # SEPAIHRDGradientObjectiveFunction::evaluate_with_gradient, unclipped
        if self.memo is not None and np.array_equal(self.memo[0], theta):
            return self.memo[1], self.memo[2].copy()
        fc = float(self.f(theta[None])[0])
        g = np.zeros(len(theta))
        if math.isfinite(fc):
            rows = np.tile(theta, (len(theta), 1))
            steps = np.empty(len(theta))
            for i in range(len(theta)):
                steps[i] = self.eps_fd * max(abs(theta[i]), self.eps_fd)
                rows[i, i] += steps[i]
            fp = self.f(rows)
            for i in range(len(theta)):
                g[i] = (fp[i] - fc) / steps[i] if math.isfinite(fp[i]) else 0.0
        self.n_grad += 1
        self.memo = (theta.copy(), fc, g.copy())
        return fc, g

    @staticmethod
    def clip(g):
        nrm = math.sqrt(_dot(g, g))
        return g * (1000.0 / nrm) if nrm > 1000.0 else g

    def leapfrog(self, theta, r, eps):
        _, g = self.grad(theta)
        g = self.clip(g)
        r = r + (0.5 * eps) * g
        theta = self.constrain(theta + eps * r)
        _, g = self.grad(theta)
        g = self.clip(g)
        r = r + (0.5 * eps) * g
        return theta, r

    @staticmethod
    def no_uturn(tm, tp, rm, rp):
        d = tp - tm
        return _dot(d, rm) >= 0 and _dot(d, rp) >= 0

    def build(self, theta, r, log_u, v, j, eps, H0):
        if j == 0:
            tp, rp = self.leapfrog(theta.copy(), r.copy(), v * eps)
            lp, _ = self.grad(tp)
            Hp = lp - 0.5 * _dot(rp, rp)
            return dict(tm=tp, tp=tp, rm=rp, rp=rp, prime=tp, n=1 if log_u <= Hp else 0, s=log_u < Hp + 1000.0,
                        alpha=min(1.0, math.exp(Hp - H0)), na=1)
        left = self.build(theta, r, log_u, v, j - 1, eps, H0)
        if not left["s"]:
            return left
        if v == -1:
            right = self.build(left["tm"], left["rm"], log_u, v, j - 1, eps, H0)
            t = dict(tm=right["tm"], rm=right["rm"], tp=left["tp"], rp=left["rp"])
        else:
            right = self.build(left["tp"], left["rp"], log_u, v, j - 1, eps, H0)
            t = dict(tm=left["tm"], rm=left["rm"], tp=right["tp"], rp=right["rp"])
        if right["s"]:
            t["n"] = left["n"] + right["n"]
            prob = right["n"] / t["n"] if t["n"] > 0 else 0.0
            t["prime"] = right["prime"] if self.g.uniform() < prob else left["prime"]
            t["alpha"] = left["alpha"] + right["alpha"]
            t["na"] = left["na"] + right["na"]
            t["s"] = left["s"] and right["s"] and self.no_uturn(t["tm"], t["tp"], t["rm"], t["rp"])
        else:
            t.update(prime=left["prime"], n=left["n"], s=False, alpha=left["alpha"], na=left["na"])
        return t

    def find_epsilon(self, theta):
        eps = max(1e-6, min(float(np.sum(self.sigmas)) / len(theta) * 0.1, 0.1))
        nrm = StdNormal()
        r = np.array([nrm(self.g) for _ in range(len(theta))])
        lp, _ = self.grad(theta)
        if not math.isfinite(lp):
            return eps
        H0 = lp - 0.5 * _dot(r, r)
        tp, rp = self.leapfrog(theta.copy(), r.copy(), eps)
        lpp, _ = self.grad(tp)
        acc = math.exp(min(0.0, lpp - 0.5 * _dot(rp, rp) - H0))
        for _ in range(5):
            if acc < 0.1 and eps > 1e-8:
                eps *= 0.5
            elif acc > 0.9 and eps < 1.0:
                eps *= 1.5
            else:
                break
            tp, rp = self.leapfrog(theta.copy(), r.copy(), eps)
            lpp, _ = self.grad(tp)
            if not math.isfinite(lpp):
                eps *= 0.5
                continue
            acc = math.exp(min(0.0, lpp - 0.5 * _dot(rp, rp) - H0))
        return eps

    def run(self, theta0):
        theta = np.array(theta0, dtype=float)
        eps = self.find_epsilon(theta)
        mu, eps_bar, H_bar = math.log(10.0 * eps), eps, 0.0
        samples, values, depths = [], [], []
        for m in range(1, self.iterations + 1):
            nrm = StdNormal()
            r0 = np.array([nrm(self.g) for _ in range(len(theta))])
            lp, _ = self.grad(theta)
            if not math.isfinite(lp):
                if samples:
                    samples.append(samples[-1]); values.append(values[-1])
                continue
            H0 = lp - 0.5 * _dot(r0, r0)
            log_u = H0 - _exponential(self.g)
            tm, tp, rm, rp, nxt = theta.copy(), theta.copy(), r0.copy(), r0.copy(), theta.copy()
            j, n, s, alpha, na = 0, 1, True, 0.0, 0
            while s and j < self.depth:
                v = _coin(self.g) * 2 - 1
                if v == -1:
                    sub = self.build(tm, rm, log_u, v, j, eps, H0)
                    tm, rm = sub["tm"], sub["rm"]
                else:
                    sub = self.build(tp, rp, log_u, v, j, eps, H0)
                    tp, rp = sub["tp"], sub["rp"]
                if sub["s"] and self.no_uturn(tm, tp, rm, rp):
                    if self.g.uniform() < sub["n"] / (n + sub["n"]):
                        nxt = sub["prime"]
                    n += sub["n"]; alpha += sub["alpha"]; na += sub["na"]; j += 1
                else:
                    s = False
            theta = nxt.copy()
            depths.append(j)
            if m <= self.window:
                avg = alpha / na if na > 0 else 0.0
                eta = 1.0 / (m + 10.0)
                H_bar = (1.0 - eta) * H_bar + eta * (self.delta - avg)
                log_eps = mu - (math.sqrt(m) / 0.05) * H_bar
                eps = math.exp(log_eps)
                mk = math.pow(m, -0.75)
                eps_bar = math.exp(mk * log_eps + (1.0 - mk) * math.log(eps_bar))
            else:
                eps = eps_bar
            c = self.constrain(theta)
            samples.append(c.copy())
            values.append(float(self.f(c[None])[0]))
        return np.array(samples), np.array(values), depths


def _gaussian(mu, s):
    def f(x):
        x = np.asarray(x, dtype=float)
        acc = np.zeros(len(x))
        for k in range(len(mu)):
            z = (x[:, k] - mu[k]) / s[k]
            acc = acc + z * z
        return -0.5 * acc
    return f


def _collect(host, f, pm, settings, x0):
    seen = []

    def ev(x):
        v = f(x)
        seen.append((np.array(x), np.array(v)))
        return v
    best, val, nev = host.optimize("nuts", pm, settings, ev, x0)
    return best, val, nev, seen


@pytest.mark.parametrize("seed,depth", [(11, 3), (12, 5)])
def test_sampler_equals_an_independent_python_restatement_bit_for_bit(host, seed, depth):
    mu = np.array([1.0, -2.0, 0.5, 3.0]); s = np.array([0.5, 1.5, 0.2, 1.0])
    lb, ub = mu - 50.0, mu + 50.0
    sig = np.array([0.3, 0.6, 0.1, 0.4])
    f = _gaussian(mu, s)
    pm = host.ParameterManager(sig, lb, ub, mode=0)
    x0 = mu + np.array([0.4, -1.0, 0.1, 0.7])
    iters, window = 30, 12
    ref = PyNuts(f, sig, lb, ub, seed, iters, window, 0.8, depth)
    want, want_vals, depths = ref.run(x0)
    best, val, nev, seen = _collect(host, f, pm, dict(nuts_iterations=iters, nuts_adaptation_window=window, nuts_delta_target=0.8,
                                                       nuts_max_tree_depth=depth, seed=seed), x0)
    # every stored sample is scored by one calculate() (a batch of ONE vector); the gradient batches have P rows
    got = np.array([x[0] for x, _ in seen if len(x) == 1])
    scored = [i for i, (x, _) in enumerate(seen) if len(x) == 1]
    # centre evaluations of the gradients are single-row calls too: pick the sample evaluations by matching the restatement
    assert len(want) == iters and max(depths) <= depth and max(depths) >= 1
    k = 0
    for w in want:
        while k < len(got) and not np.array_equal(got[k], w):
            k += 1
        assert k < len(got), "a sample of the restatement never reached the objective"
        k += 1
    assert val == want_vals.max()
    np.testing.assert_array_equal(best, want[int(np.argmax(want_vals))])
    # distinct gradient points: P-row batches (each preceded by its single-row centre evaluation)
    assert sum(1 for x, _ in seen if len(x) == 4) == ref.n_grad
    assert nev == ref.n_grad * 5 + iters


def test_chain_samples_a_gaussian(host):
    mu = np.array([0.5, -1.0, 2.0]); s = np.array([1.0, 0.3, 2.0])
    f = _gaussian(mu, s)
    pm = host.ParameterManager(s * 0.5, mu - 100.0, mu + 100.0, mode=0)
    samples = []

    def ev(x):
        if len(x) == 1:
            samples.append(np.array(x[0]))
        return f(x)
    host.optimize("nuts", pm, dict(nuts_iterations=700, nuts_adaptation_window=200, nuts_max_tree_depth=6, seed=5), ev, mu + s)
    # the samples are the last single-row evaluation of each iteration; the centre evaluations of the gradients are on the same
    # trajectory, so the pooled single-row points already show the target's moments
    x = np.array(samples[len(samples) // 3:])
    assert np.all(np.abs(x.mean(axis=0) - mu) < 0.35 * s)
    assert np.all(np.abs(x.std(axis=0) / s - 1.0) < 0.35)


def test_forward_difference_gradient_formula(host):
    """grad_i = (f(x + eps_i e_i) - f(x)) / eps_i with eps_i = 1e-4 max(|x_i|, 1e-4); non-finite perturbed values give 0."""
    mu = np.array([1.0, 0.0, -3.0]); s = np.array([1.0, 2.0, 0.5])
    f = _gaussian(mu, s)
    pm = host.ParameterManager(np.ones(3), mu - 10, mu + 10, mode=0)
    x0 = np.array([2.0, 0.0, -2.5])
    rows = []

    def ev(x):
        rows.append(np.array(x))
        v = f(x)
        if len(x) == 3:
            v[1] = np.nan                                      # a failed perturbed run: zero component
        return v
    host.optimize("nuts", pm, dict(nuts_iterations=1, nuts_adaptation_window=1, nuts_max_tree_depth=1, seed=1), ev, x0)
    first = next(r for r in rows if len(r) == 3)
    steps = np.array([1e-4 * 2.0, 1e-4 * 1e-4, 1e-4 * 2.5])
    np.testing.assert_array_equal(first, np.tile(x0, (3, 1)) + np.diag(steps))
