"""CPU tests of the whole ParticleSwarmOptimization class of the host mirror (host/optimizers.cpp): the five variants, four
topologies, opposition-based initialisation, evolutionary-state adaptation, stagnation restart and elitist learning of
src/model/optimizers/ParticleSwarmOptimizer.cpp.

The reference has no test of its optimizer and seeds it from std::random_device, so the anchors are:
  * an independent Python restatement of initializeSwarm + updateParticles + standardPSOUpdate + getNeighborhoodBest over a
    Python std::mt19937 / std::uniform_real_distribution (numpy's MT19937 with init_genrand seeding is the same engine): every
    batch the C++ swarm hands to the objective must equal the restatement's positions BIT FOR BIT;
  * the batches themselves: the objective callback sees every position the swarm evaluates, so selection, restart and elitist
    learning are checked on what was actually evaluated, not on internal state.
"""
import math

import numpy as np
import pytest


@pytest.fixture(scope="module")
def host(pkg, cuda_lib):
    import __graft_entry__ as entry
    entry.build()
    from sepaihrd_b200 import hostlib
    hostlib.load_library()
    return hostlib


# ---- std::mt19937 + std::uniform_real_distribution<double>(0, 1) as libstdc++ evaluates them -----------------------------
class StdMt19937:
    def __init__(self, seed: int):
        self.bg = np.random.MT19937()
        self.bg._legacy_seeding(int(seed) & 0xFFFFFFFF)          # init_genrand(seed) == std::mt19937(seed)
        self.buf = np.empty(0, dtype=np.uint64)
        self.i = 0

    def raw(self) -> int:
        if self.i == len(self.buf):
            self.buf = self.bg.random_raw(256)
            self.i = 0
        v = int(self.buf[self.i])
        self.i += 1
        return v

    def uniform(self) -> float:
        # generate_canonical<double, 53>: two 32-bit draws, sum = r0 + r1 * 2^32 (one rounding), / 2^64
        r0, r1 = self.raw(), self.raw()
        x = (float(r0) + float(r1) * 4294967296.0) / 18446744073709551616.0
        return math.nextafter(1.0, 0.0) if x >= 1.0 else x


def _neighbors(topology: int, i: int, N: int):
    if topology == 1:                                           # ring, two on each side
        nb = [i]
        for j in (1, 2):
            nb += [(i - j + N) % N, (i + j) % N]
        return nb
    g = int(math.ceil(math.sqrt(N)))                            # von Neumann grid without wrap-around
    row, col = divmod(i, g)
    nb = [i]
    if row > 0 and (row - 1) * g + col < N: nb.append((row - 1) * g + col)
    if row < g - 1 and (row + 1) * g + col < N: nb.append((row + 1) * g + col)
    if col > 0 and row * g + col - 1 < N: nb.append(row * g + col - 1)
    if col < g - 1 and row * g + col + 1 < N: nb.append(row * g + col + 1)
    return nb


def _python_standard_pso(ev, lb, ub, N, iterations, seed, topology, initial, st):
    """initializeSwarm (.cpp:249-328) + `iterations` x updateParticles (.cpp:330-425) with standardPSOUpdate (.cpp:576-618),
    linear coefficient schedules, synchronous neighbourhood bests.  Returns the list of evaluated batches."""
    n = len(lb)
    master = StdMt19937(seed)
    seeds = [master.raw() for _ in range(N)]
    pos = np.empty((N, n)); vel = np.empty((N, n))
    for i in range(N):
        g = StdMt19937(seeds[i])
        if i == 0 and initial is not None:
            pos[0] = np.minimum(np.maximum(initial, lb), ub)
        else:
            for k in range(n):
                pos[i, k] = lb[k] + g.uniform() * (ub[k] - lb[k])
        for k in range(n):
            vmax = 0.2 * (ub[k] - lb[k])
            vel[i, k] = -vmax + 2 * vmax * g.uniform()
    batches = [pos.copy()]
    fit = ev(pos)
    pbest, pbest_val = pos.copy(), fit.copy()
    gbest_val, gbest = -np.inf, None
    for i in range(N):
        if pbest_val[i] > gbest_val:
            gbest_val, gbest = pbest_val[i], pbest[i].copy()
    for it in range(iterations):
        ratio = it / (iterations - 1) if iterations > 1 else 0.0
        omega = st["omega_start"] + (st["omega_end"] - st["omega_start"]) * ratio
        c1 = st["c1_initial"] + (st["c1_final"] - st["c1_initial"]) * ratio
        c2 = st["c2_initial"] + (st["c2_final"] - st["c2_initial"]) * ratio
        seeds = [master.raw() for _ in range(N)]
        snap, snap_val = pbest.copy(), pbest_val.copy()
        for i in range(N):
            g = StdMt19937(seeds[i])
            if topology == 0:
                lbest = gbest
            else:
                b, bv = i, snap_val[i]
                for j in _neighbors(topology, i, N):
                    if snap_val[j] > bv:
                        b, bv = j, snap_val[j]
                lbest = snap[b]
            r = [(g.uniform(), g.uniform()) for _ in range(n)]
            for k in range(n):
                cognitive = c1 * (r[k][0] * (snap[i, k] - pos[i, k]))
                social = c2 * (r[k][1] * (lbest[k] - pos[i, k]))
                v = omega * vel[i, k] + cognitive + social
                vmax = 0.2 * (ub[k] - lb[k])
                v = min(max(v, -vmax), vmax)
                p = pos[i, k] + v
                if p < lb[k]:
                    p = lb[k] + abs(p - lb[k]); v *= -0.5
                elif p > ub[k]:
                    p = ub[k] - abs(p - ub[k]); v *= -0.5
                pos[i, k] = min(max(p, lb[k]), ub[k])
                vel[i, k] = v
        batches.append(pos.copy())
        fit = ev(pos)
        for i in range(N):
            if fit[i] > pbest_val[i]:
                pbest_val[i], pbest[i] = fit[i], pos[i]
        for i in range(N):
            if pbest_val[i] > gbest_val:
                gbest_val, gbest = pbest_val[i], pbest[i].copy()
    return batches, gbest_val, gbest


def _box(n=6):
    lb = np.array([0.0, -1.0, 2.0, 10.0, -5.0, 0.5, 1.0, -3.0][:n])
    ub = np.array([1.0, 1.0, 7.0, 30.0, -1.0, 0.75, 9.0, 3.0][:n])
    target = lb + np.array([0.3, 0.8, 0.5, 0.1, 0.6, 0.95, 0.4, 0.2][:n]) * (ub - lb)
    return lb, ub, target


def _recording(ev):
    batches = []

    def f(x):
        batches.append(np.array(x))
        return ev(x)
    return f, batches


LINEAR = dict(omega_start=0.9, omega_end=0.4, c1_initial=2.0, c1_final=0.5, c2_initial=0.5, c2_final=2.0)


@pytest.mark.parametrize("topology", [0, 1, 2])
@pytest.mark.parametrize("with_initial", [False, True])
def test_standard_swarm_equals_an_independent_python_restatement_bit_for_bit(host, topology, with_initial):
    lb, ub, target = _box()
    ev = lambda x: -(((x - target) / (ub - lb)) ** 2).sum(axis=1)
    N, iters, seed = 11, 7, 1234                                 # 11: a 4-wide von Neumann grid with a ragged last row
    x0 = lb + 0.5 * (ub - lb) if with_initial else None
    want, want_val, want_pos = _python_standard_pso(ev, lb, ub, N, iters, seed, topology, x0, LINEAR)
    pm = host.ParameterManager(np.ones_like(lb), lb, ub, mode=0)
    f, got = _recording(ev)
    sw = host.Swarm(pm, dict(LINEAR, iterations=iters, swarm_size=N, seed=seed, variant=0, topology=topology, use_opposition_learning=0,
                             use_adaptive_parameters=0, max_stagnation=1000))
    best, val, stats = sw.run(f, x0)
    assert len(got) == len(want) == iters + 1 and stats["evaluations"] == N * (iters + 1) and stats["restarts"] == 0
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)
    assert val == want_val
    np.testing.assert_array_equal(best, want_pos)


def test_neighbourhoods_of_the_four_topologies(host):
    lb, ub, _ = _box(3)
    pm = host.ParameterManager(np.ones_like(lb), lb, ub, mode=0)
    N = 10
    mk = lambda topo: host.Swarm(pm, dict(iterations=1, swarm_size=N, seed=5, topology=topo))
    assert mk(0).neighbors(3) == list(range(N))
    ring = mk(1)
    assert ring.neighbors(0) == [0, 9, 1, 8, 2] and ring.neighbors(9) == [9, 8, 0, 7, 1]     # (.cpp:850-859) self, then -j / +j
    grid = mk(2)                                                  # ceil(sqrt(10)) = 4 columns: rows 0-1 full, row 2 holds 8, 9
    assert grid.neighbors(0) == [0, 4, 1]
    assert grid.neighbors(5) == [5, 1, 9, 4, 6]                   # up, down, left, right (.cpp:868-884)
    assert grid.neighbors(6) == [6, 2, 5, 7]                      # below would be 10: outside the swarm
    assert grid.neighbors(9) == [9, 5, 8]                         # (2, 1): nothing below (13 is outside), right neighbour 10 is outside
    for i in range(N):
        assert grid.neighbors(i) == _neighbors(2, i, N) and ring.neighbors(i) == _neighbors(1, i, N)
    rnd = mk(3)
    seen = set()
    for _ in range(6):
        nb = rnd.neighbors(4)
        assert nb[0] == 4 and len(nb) == 5 and len(set(nb)) == 5 and all(0 <= j < N for j in nb)
        seen.add(tuple(nb))
    assert len(seen) > 1                                          # redrawn on every call
    with pytest.raises(host.HostError):
        rnd.neighbors(N)


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("topology", [0, 1, 2, 3])
def test_every_variant_on_every_topology_improves_stays_in_bounds_and_repeats(host, variant, topology):
    lb, ub, target = _box(8)
    ev = lambda x: -(((x - target) / (ub - lb)) ** 2).sum(axis=1)
    pm = host.ParameterManager(np.ones_like(lb), lb, ub, mode=0)
    N, iters = 20, 25
    st = dict(iterations=iters, swarm_size=N, seed=77, variant=variant, topology=topology, use_opposition_learning=1, use_adaptive_parameters=1)

    def run(settings):
        f, batches = _recording(ev)
        out = host.Swarm(pm, settings).run(f)
        return out, batches
    (best, val, stats), batches = run(st)
    first = ev(batches[0]).max()
    assert val > first and val > -0.05                            # the initial best of 20 uniform draws in 8 dimensions is ~ -0.3
    for b in batches:
        assert np.all(b >= lb) and np.all(b <= ub)
    np.testing.assert_allclose(ev(best[None])[0], val, rtol=0, atol=0)
    # evaluation count: initial swarm twice (opposition selection re-evaluates it), one swarm per iteration, 3 elitist trials
    # every fifth iteration for ADAPTIVE / HYBRID (.cpp:158-180), (N - 3) per restart
    els = 3 * len(range(0, iters, 5)) if variant in (2, 4) else 0
    assert stats["evaluations"] == 2 * N + iters * N + els + stats["restarts"] * (N - 3)
    assert sum(len(b) for b in batches) == stats["evaluations"]
    assert 0.0 <= stats["diversity"] <= 1.0
    (best2, val2, stats2), batches2 = run(st)                     # same seed: the same run
    assert val2 == val and stats2 == stats
    np.testing.assert_array_equal(best2, best)
    (_, val3, _), batches3 = run(dict(st, seed=78))
    assert not np.array_equal(batches3[0], batches[0])


def test_opposition_selection_reorders_by_fitness_and_admits_opposites_only_for_minus_infinity(host):
    lb, ub, target = _box(5)
    base = lambda x: -(((x - target) / (ub - lb)) ** 2).sum(axis=1)
    pm = host.ParameterManager(np.ones_like(lb), lb, ub, mode=0)
    st = dict(LINEAR, iterations=1, swarm_size=9, seed=3, variant=0, topology=0, use_opposition_learning=1, use_adaptive_parameters=0)
    f, b = _recording(base)
    host.Swarm(pm, st).run(f)
    # the opposite particles are never scored before the selection (their pbest_value is the struct default -inf), so the
    # second evaluation of initializeSwarm sees the ORIGINAL swarm sorted by descending fitness (.cpp:306-315, 507-560)
    order = np.argsort(-base(b[0]), kind="stable")
    np.testing.assert_array_equal(b[1], b[0][order])
    # an original that scored -inf ties with the opposites; the stable selection keeps push order (original 0, opposite 0, original 1, ...)
    calls = {"n": 0}

    def with_holes(x):
        calls["n"] += 1
        v = base(x)
        if calls["n"] == 1:
            v[[2, 6]] = -np.inf
        return v
    f, b = _recording(with_holes)
    host.Swarm(pm, st).run(f)
    v0 = base(b[0]); v0[[2, 6]] = -np.inf
    keep = [i for i in np.argsort(-v0, kind="stable") if np.isfinite(v0[i])]
    np.testing.assert_array_equal(b[1][:7], b[0][keep])
    np.testing.assert_array_equal(b[1][7], lb + ub - b[0][0])                     # first -inf entries in push order: opposite 0,
    np.testing.assert_array_equal(b[1][8], lb + ub - b[0][1])                     # opposite 1 (original 2 would come third)


def test_stagnation_restart_keeps_three_elites_and_redraws_the_rest(host):
    lb, ub, target = _box(5)
    pm = host.ParameterManager(np.ones_like(lb), lb, ub, mode=0)
    N = 12
    # a plateau objective: nothing ever improves on the initial best, so |gbest - previous| < threshold from iteration 1 on
    # and the counter passes max_stagnation = 2 at iteration 3 (.cpp:133-146)
    flat = lambda x: np.where(x[:, 0] > -1e300, -1.0, 0.0)
    st = dict(LINEAR, iterations=5, swarm_size=N, seed=8, variant=0, topology=0, use_opposition_learning=0, use_adaptive_parameters=0,
              max_stagnation=2, restart_threshold=1e-6)
    f, b = _recording(flat)
    best, val, stats = host.Swarm(pm, st).run(f)
    assert stats["restarts"] == 1 and val == -1.0
    sizes = [len(x) for x in b]
    assert sizes == [N, N, N, N, N - 3, N, N]                     # init, iterations 0-2, restart batch (all but 3 elites), iterations 3-4
    restart = b[4]
    assert np.all(restart >= lb) and np.all(restart <= ub)
    # with max_stagnation large the same run never restarts
    f, b = _recording(flat)
    _, _, stats = host.Swarm(pm, dict(st, max_stagnation=50)).run(f)
    assert stats["restarts"] == 0 and [len(x) for x in b] == [N] * 6
    # a global best that improves in every iteration resets the counter each time: no restart although max_stagnation is 1
    calls = {"n": 0}

    def rising(x):
        calls["n"] += 1
        return np.full(len(x), float(calls["n"]))
    _, val, stats = host.Swarm(pm, dict(st, iterations=12, max_stagnation=1)).run(rising)
    assert stats["restarts"] == 0 and val == 13.0


def test_elitist_learning_trials_surround_the_best_particle_with_a_halving_radius(host):
    lb, ub, target = _box(5)
    ev = lambda x: -(((x - target) / (ub - lb)) ** 2).sum(axis=1)
    pm = host.ParameterManager(np.ones_like(lb), lb, ub, mode=0)
    N = 10
    st = dict(iterations=6, swarm_size=N, seed=21, variant=2, topology=0, use_opposition_learning=0, use_adaptive_parameters=0)
    f, b = _recording(ev)
    _, val, stats = host.Swarm(pm, st).run(f)
    sizes = [len(x) for x in b]
    assert sizes == [N, N, 3, N, N, N, N, N, 3]                   # elitist trials after iterations 0 and 5 (iter % 5 == 0)
    assert 1 <= stats["elitist_trials"] <= 6
    # the trials are Gaussian around the best particle's position of the batch before, sigma = 0.1 exp(-2 sr) (ub - lb) * 0.5^attempt
    swarm, trials = b[1], b[2]
    fit = ev(swarm)
    pbest = np.maximum(ev(b[0]), fit)
    center = swarm[int(np.argmax(pbest))]
    z = (trials - center) / (0.1 * (ub - lb))
    inside = (trials > lb) & (trials < ub)
    assert np.abs(z[inside]).max() < 6.0                           # |N(0,1)| * exp(-2 sr) * 0.5^a
    # an unclamped coordinate's spread shrinks with the attempt number on average
    spread = [np.sqrt(np.mean(z[a][inside[a]] ** 2)) for a in range(3)]
    assert spread[2] < spread[0] * 1.5


def test_configure_validates_like_the_reference_and_the_stepwise_form_is_the_basic_swarm(host):
    lb, ub, _ = _box(3)
    pm = host.ParameterManager(np.ones_like(lb), lb, ub, mode=0)
    for bad in (dict(variant=5), dict(variant=-1), dict(topology=4), dict(iterations=0), dict(swarm_size=0), dict(omega_start=-0.1),
                dict(c2_final=-1), dict(max_stagnation=0), dict(report_interval=0)):
        with pytest.raises(host.HostError):
            host.Swarm(pm, bad)
    # begin / tell / step drive the STANDARD / GLOBAL_BEST swarm; any other configuration has to go through run()
    for other in (dict(variant=2), dict(topology=2), dict(use_opposition_learning=1), dict(use_adaptive_parameters=1)):
        sw = host.Swarm(pm, dict(iterations=2, swarm_size=4, seed=1, **other))
        with pytest.raises(host.HostError):
            sw.begin(None)
    sw = host.Swarm(pm, dict(iterations=2, swarm_size=4, seed=1))
    sw.begin(None)
    assert sw.positions().shape == (4, 3)


def test_shipped_reference_settings_file_is_accepted(host, tmp_path):
    """data/configuration/pso_settings.txt of the reference selects topology 2 (von Neumann), opposition learning and adaptive
    parameters: the mirror has to run it as configured (with a usable swarm size instead of the file's 1 x 1 smoke values)."""
    text = "\n".join(["iterations 1", "swarm_size 1", "omega_start 0.9", "omega_end 0.4", "c1_initial 2.0", "c1_final 0.5", "c2_initial 0.5",
                      "c2_final 2.0", "report_interval 1", "variant 0", "topology 2", "use_opposition_learning 1.0", "use_parallel 1.0",
                      "use_adaptive_parameters 1.0", "diversity_threshold 0.1", "restart_threshold 1e-6", "quantum_beta 1.0", "levy_alpha 1.5",
                      "max_stagnation 20", "log_evolutionary_state 1.0"]) + "\n"
    path = tmp_path / "pso_settings.txt"
    path.write_text(text)
    st = host.read_file("settings", str(path))
    assert st["topology"] == 2 and st["use_opposition_learning"] == 1 and st["max_stagnation"] == 20
    lb, ub, target = _box(6)
    ev = lambda x: -(((x - target) / (ub - lb)) ** 2).sum(axis=1)
    pm = host.ParameterManager(np.ones_like(lb), lb, ub, mode=0)
    best, val, stats = host.Swarm(pm, dict(st, seed=1)).run(ev, lb + 0.5 * (ub - lb))      # 1 particle, 1 iteration, as shipped
    assert stats["evaluations"] == 3 and np.isfinite(val)
    best, val, stats = host.Swarm(pm, dict(st, seed=1, iterations=30, swarm_size=25)).run(ev)
    assert val > -0.02


# ---------------------------------------------------------------------------------------------------------------------
# The configuration the reference ships and defaults to -- opposition-based initialisation, evolutionary-state parameter
# adaptation, ADAPTIVE variant with elitist learning, stagnation restart -- restated in Python on top of the standard update
# above, sequentially as the reference runs it (one trial evaluation at a time in the elitist strategy; the C++ mirror draws
# the three trials ahead and rewinds the master generator).
class _Normal:
    """std::normal_distribution<double>(0, 1) of libstdc++ (polar method, second value saved)."""

    def __init__(self):
        self.saved = None

    def __call__(self, g):
        if self.saved is not None:
            v, self.saved = self.saved, None
            return v
        while True:
            x = 2.0 * g.uniform() - 1.0
            y = 2.0 * g.uniform() - 1.0
            r2 = x * x + y * y
            if not (r2 > 1.0 or r2 == 0.0):
                break
        mult = math.sqrt(-2.0 * math.log(r2) / r2)
        self.saved = x * mult
        return y * mult


def _uniform_int(g, m):
    """std::uniform_int_distribution<unsigned long>(0, m - 1) on std::mt19937 (libstdc++ >= 11, Lemire's method, 64-bit product)."""
    M = 0xFFFFFFFF
    product = g.raw() * m
    low = product & M
    if low < m:
        threshold = ((M + 1) - m) % m
        while low < threshold:
            product = g.raw() * m
            low = product & M
    return product >> 32


def _std_shuffle(a, g):
    """std::shuffle of libstdc++ (bits/stl_algo.h): two swap positions from one draw while range^2 fits the engine's range."""
    n = len(a)
    if n == 0:
        return
    if (0xFFFFFFFF // n) >= n:
        i = 1
        if n % 2 == 0:
            j = _uniform_int(g, 2)
            a[i], a[j] = a[j], a[i]
            i += 1
        while i != n:
            swap_range = i + 1
            x = _uniform_int(g, swap_range * (swap_range + 1))
            p0, p1 = x // (swap_range + 1), x % (swap_range + 1)
            a[i], a[p0] = a[p0], a[i]
            i += 1
            a[i], a[p1] = a[p1], a[i]
            i += 1
        return
    for i in range(1, n):
        j = _uniform_int(g, i + 1)
        a[i], a[j] = a[j], a[i]


def _levy_sigma_u(alpha):
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.tgamma.restype = ctypes.c_double; libm.tgamma.argtypes = [ctypes.c_double]
    tg = libm.tgamma                                            # std::tgamma of the C++ side is this function
    return math.pow(tg(1 + alpha) * math.sin(math.pi * alpha / 2) / (tg((1 + alpha) / 2) * alpha * math.pow(2, (alpha - 1) / 2)), 1.0 / alpha)


def _python_full_pso(ev, lb, ub, N, iterations, seed, topology, variant, st, max_stagnation, restart_threshold):
    n = len(lb)
    master = StdMt19937(seed)
    master_normal = _Normal()
    swarm_batches = []

    def evaluate(rows):
        swarm_batches.append(np.array(rows))
        return ev(np.array(rows))
    # initializeSwarm
    seeds = [master.raw() for _ in range(N)]
    pos = np.empty((N, n)); vel = np.empty((N, n))
    for i in range(N):
        g = StdMt19937(seeds[i])
        for k in range(n):
            pos[i, k] = lb[k] + g.uniform() * (ub[k] - lb[k])
        for k in range(n):
            vmax = 0.2 * (ub[k] - lb[k])
            vel[i, k] = -vmax + 2 * vmax * g.uniform()
    fit = evaluate(pos)
    # oppositionBasedInitialization: the opposites are never scored -> the originals, sorted by fitness (stable)
    order = sorted(range(N), key=lambda i: -fit[i])
    pos, vel = pos[order].copy(), vel[order].copy()
    fit = evaluate(pos)
    cur_fit = fit.copy()
    pbest, pbest_val = pos.copy(), fit.copy()
    succ_count = np.zeros(N, dtype=int); total = np.zeros(N, dtype=int); succ_rate = np.zeros(N)
    gbest_val, gbest = -np.inf, None

    def rescan():
        nonlocal gbest_val, gbest
        for i in range(N):
            if pbest_val[i] > gbest_val:
                gbest_val, gbest = pbest_val[i], pbest[i].copy()
    rescan()
    previous, stagnation, restarts, els_trials = -np.inf, 0, 0, 0
    for it in range(iterations):
        if abs(gbest_val - previous) < restart_threshold:
            stagnation += 1
            if stagnation > max_stagnation:                    # restartSwarm
                restarts += 1
                order = sorted(range(N), key=lambda i: -pbest_val[i])
                pos, vel, pbest = pos[order].copy(), vel[order].copy(), pbest[order].copy()
                pbest_val, cur_fit = pbest_val[order].copy(), cur_fit[order].copy()
                succ_count, total, succ_rate = succ_count[order].copy(), total[order].copy(), succ_rate[order].copy()
                elite = pos[:3].copy()
                seeds = [master.raw() for _ in range(N)]
                for i in range(3, N):
                    g = StdMt19937(seeds[i]); nrm = _Normal()
                    el = elite[i % 3]
                    for k in range(n):
                        if g.uniform() < 0.7:
                            rng_ = ub[k] - lb[k]
                            sigma = 0.3 * rng_ * (1.0 + 0.5 * g.uniform())
                            pos[i, k] = el[k] + sigma * nrm(g)
                        else:
                            pos[i, k] = lb[k] + g.uniform() * (ub[k] - lb[k])
                        pos[i, k] = min(max(pos[i, k], lb[k]), ub[k])
                        vmax = 0.2 * (ub[k] - lb[k])
                        vel[i, k] = -vmax + 2 * vmax * g.uniform()
                f2 = evaluate(pos[3:])
                for i in range(3, N):
                    cur_fit[i] = f2[i - 3]; pbest[i] = pos[i]; pbest_val[i] = f2[i - 3]
                    succ_count[i] = 0; total[i] = 0; succ_rate[i] = 0.0
                gbest_val, gbest = pbest_val[0], pbest[0].copy()
                stagnation = 0
        else:
            stagnation = 0
        previous = gbest_val
        # updateParticles: evolutionary state -> coefficients
        md = mx = mf = 0.0; maxf, minf = -np.inf, np.inf
        for i in range(N):
            d2 = 0.0
            for k in range(n):
                d = pos[i, k] - gbest[k]
                d2 += d * d
            dist = math.sqrt(d2)
            md += dist; mx = max(mx, dist)
            mf += cur_fit[i]; maxf = max(maxf, cur_fit[i]); minf = min(minf, cur_fit[i])
        md /= N; mf /= N
        frange = (maxf - minf) if (maxf - minf) > 1e-10 else 1e-10
        ef = 0.5 * ((md / mx) if mx > 0 else 0.0) + 0.5 * (1.0 - (maxf - mf) / frange)
        ratio = it / (iterations - 1) if iterations > 1 else 0.0
        if ef > 0.7:
            omega = 0.9 - 0.2 * ratio; c1 = 1.5 + 0.5 * math.sin(ratio * math.pi); c2 = 1.5 - 0.5 * math.sin(ratio * math.pi)
        elif ef > 0.4:
            omega = 0.7 - 0.3 * ratio; c1 = 2.0 - ratio; c2 = 1.0 + ratio
        elif ef > 0.2:
            omega = 0.4 - 0.3 * ratio; c1 = 1.0 - 0.5 * ratio; c2 = 2.0 + 0.5 * ratio
        else:
            omega = 0.9 + 0.1 * master.uniform(); c1 = 2.5 + master.uniform(); c2 = 0.5 + master.uniform()
        omega = min(max(omega, 0.1), 1.0); c1 = min(max(c1, 0.0), 4.0); c2 = min(max(c2, 0.0), 4.0)
        mean_best = None
        if variant in (1, 4):                                  # calculateMeanBestPosition (the mirror computes it for HYBRID too)
            mean_best = np.zeros(n)
            for i in range(N):
                for k in range(n):
                    mean_best[k] += pbest[i, k]
            mean_best = mean_best / N
        seeds = [master.raw() for _ in range(N)]
        snap, snap_val = pbest.copy(), pbest_val.copy()
        lbest_of = list(range(N))
        if topology != 0:                                      # neighbourhood bests from the snapshot, in particle order
            for i in range(N):
                if topology == 3:                              # RANDOM_DYNAMIC: four of the others, shuffled with the MASTER generator
                    cand = [j for j in range(N) if j != i]
                    _std_shuffle(cand, master)
                    nb = [i] + cand[:min(4, len(cand))]
                else:
                    nb = _neighbors(topology, i, N)
                b, bv = i, snap_val[i]
                for j in nb:
                    if snap_val[j] > bv:
                        b, bv = j, snap_val[j]
                lbest_of[i] = b

        def standard(i, g, lbest):
            r = [(g.uniform(), g.uniform()) for _ in range(n)]
            for k in range(n):
                cognitive = c1 * (r[k][0] * (snap[i, k] - pos[i, k]))
                social = c2 * (r[k][1] * (lbest[k] - pos[i, k]))
                v = omega * vel[i, k] + cognitive + social
                vmax = 0.2 * (ub[k] - lb[k])
                v = min(max(v, -vmax), vmax)
                p = pos[i, k] + v
                if p < lb[k]:
                    p = lb[k] + abs(p - lb[k]); v *= -0.5
                elif p > ub[k]:
                    p = ub[k] - abs(p - ub[k]); v *= -0.5
                pos[i, k] = min(max(p, lb[k]), ub[k]); vel[i, k] = v

        def quantum(i, g):                                     # quantumPSOUpdate
            phi = g.uniform()
            beta = st.get("quantum_beta", 1.0) * (1.0 - 0.5 * float(it) / iterations)
            for k in range(n):
                attractor = phi * snap[i, k] + (1 - phi) * gbest[k]
                u = g.uniform()
                Lq = 2.0 * beta * abs(mean_best[k] - pos[i, k])
                if g.uniform() < 0.5:
                    p = attractor + Lq * math.log(1.0 / u)
                else:
                    p = attractor - Lq * math.log(1.0 / u)
                pos[i, k] = min(max(p, lb[k]), ub[k])

        def levy(i, g):                                        # levyFlightUpdate: towards the GLOBAL best, then an occasional jump
            standard(i, g, gbest)
            if g.uniform() < 0.1 * (1.0 + succ_rate[i]):
                alpha = st.get("levy_alpha", 1.5)
                sigma_u = _levy_sigma_u(alpha)
                steps = []
                for _ in range(n):
                    nrm = _Normal()                            # a fresh distribution per number
                    u = nrm(g) * sigma_u
                    v = max(abs(nrm(g)), 1e-10)
                    steps.append(min(max(u / math.pow(v, 1.0 / alpha), -100.0), 100.0))
                step_scale = 0.01 * (1.0 - stagnation / float(max_stagnation))
                for k in range(n):
                    scale = step_scale * (ub[k] - lb[k])
                    p = pos[i, k] + scale * steps[k]
                    pos[i, k] = min(max(p, lb[k]), ub[k])

        for i in range(N):
            g = StdMt19937(seeds[i])
            lbest = gbest if topology == 0 else snap[lbest_of[i]]
            if variant in (0, 2):
                standard(i, g, lbest)
            elif variant == 1:
                quantum(i, g)
            elif variant == 3:
                levy(i, g)
            else:                                              # HYBRID: the uniform is drawn only when its test is reached
                if succ_rate[i] < 0.3 and g.uniform() < 0.5:
                    levy(i, g)
                elif succ_rate[i] > 0.7 and g.uniform() < 0.3:
                    quantum(i, g)
                else:
                    standard(i, g, lbest)
        cur_fit = evaluate(pos).copy()
        for i in range(N):
            total[i] += 1
            if cur_fit[i] > pbest_val[i]:
                pbest_val[i] = cur_fit[i]; pbest[i] = pos[i]; succ_count[i] += 1
            succ_rate[i] = succ_count[i] / total[i] if total[i] > 0 else 0.0
        rescan()
        if variant in (2, 4) and it % 5 == 0:                 # applyElitistLearningStrategy, one trial at a time
            b = int(np.argmax(pbest_val))
            sigma_scale = 0.1 * math.exp(-2.0 * succ_rate[b])
            for _ in range(3):
                trial = np.array([min(max(pos[b, k] + (sigma_scale * (ub[k] - lb[k])) * master_normal(master), lb[k]), ub[k]) for k in range(n)])
                tf = float(ev(trial[None])[0])
                els_trials += 1
                if tf > pbest_val[b]:
                    pos[b] = trial; pbest[b] = trial; pbest_val[b] = tf; cur_fit[b] = tf
                    break
                sigma_scale *= 0.5
            if pbest_val[b] > gbest_val:
                gbest_val, gbest = pbest_val[b], pbest[b].copy()
    return swarm_batches, gbest_val, gbest, restarts, els_trials


@pytest.mark.parametrize("variant,topology", [(0, 2), (2, 0), (2, 1), (1, 0), (3, 1), (4, 2), (0, 3), (4, 3)])
def test_shipped_and_default_configurations_equal_the_python_restatement_bit_for_bit(host, variant, topology):
    lb, ub, target = _box(5)
    ev = lambda x: -(((x - target) / (ub - lb)) ** 2).sum(axis=1)
    N, iters, seed = 13, 40, 99
    st = dict(iterations=iters, swarm_size=N, seed=seed, variant=variant, topology=topology, use_opposition_learning=1,
              use_adaptive_parameters=1, max_stagnation=2, restart_threshold=2e-3)
    want, want_val, want_pos, restarts, els = _python_full_pso(ev, lb, ub, N, iters, seed, topology, variant, st, 2, 2e-3)
    pm = host.ParameterManager(np.ones_like(lb), lb, ub, mode=0)
    f, got = _recording(ev)
    best, val, stats = host.Swarm(pm, st).run(f)
    got_swarm = [b for b in got if len(b) != 3]               # the mirror scores the three elitist trials as one batch
    assert len(got_swarm) == len(want)
    for k, (a, b) in enumerate(zip(got_swarm, want)):
        np.testing.assert_array_equal(a, b, err_msg=f"swarm batch {k} differs")
    assert val == want_val
    np.testing.assert_array_equal(best, want_pos)
    assert stats["restarts"] == restarts and restarts >= 1     # the stagnation restart ran (batches of N - 3 positions)
    assert stats["elitist_trials"] == els and (els > 0) == (variant in (2, 4))
