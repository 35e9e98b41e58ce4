"""ctypes binding of the C++ host layer (host/host_capi.h -> host/libsepaihrd_host.so).

The host layer mirrors the reference's C++ interfaces (IObjectiveFunction, IParameterManager,
AgeSEPAIHRDSimulator, SEPAIHRDModelCalibration, the batched MetropolisHastingsSampler /
ParticleSwarmOptimization / HillClimbingOptimizer) over the device C ABI.  Python only steps it and does the
cross-GPU exchange between the steps (drivers.py).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "host", "libsepaihrd_host.so")
_lib = None

_vp = C.c_void_p
_vpp = C.POINTER(C.c_void_p)
_dp = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)
_keys = C.POINTER(C.c_char_p)
BATCH_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, _dp, C.c_int64, C.c_int64, _dp)

# every symbol host/host_capi.h declares
SIGNATURES = {
    "sepaihrd_host_last_error": (C.c_char_p, []),
    "sepaihrd_host_set_threads": (C.c_int32, [C.c_int32]),
    "sepaihrd_host_set_trace_directory": (C.c_int32, [C.c_char_p]),
    "sepaihrd_host_pm_create": (C.c_int32, [C.c_int32, _vp, _vp, _vp, C.c_int32, _vpp]),
    "sepaihrd_host_pm_set_mode": (C.c_int32, [_vp, C.c_int32]),
    "sepaihrd_host_pm_apply_constraints": (C.c_int32, [_vp, _vp, _vp]),
    "sepaihrd_host_pm_destroy": (None, [_vp]),
    "sepaihrd_host_mh_create": (C.c_int32, [_vp, C.c_int32, _keys, _vp, _vpp]),
    "sepaihrd_host_mh_set_initial_covariance": (C.c_int32, [_vp, _vp, C.c_int32]),
    "sepaihrd_host_mh_begin": (C.c_int32, [_vp, _vp, _vp]),
    "sepaihrd_host_mh_done": (C.c_int32, [_vp]),
    "sepaihrd_host_mh_iteration": (C.c_int32, [_vp]),
    "sepaihrd_host_mh_propose": (C.c_int32, [_vp, _vp]),
    "sepaihrd_host_mh_accept": (C.c_int32, [_vp, _vp, _vp]),
    "sepaihrd_host_mh_state": (C.c_int32, [_vp, _vp, _vp, _vp, _vp]),
    "sepaihrd_host_mh_best": (C.c_int32, [_vp, _vp, _dp]),
    "sepaihrd_host_mh_shared_cholesky": (C.c_int32, [_vp, _vp]),
    "sepaihrd_host_det_log": (C.c_double, [C.c_double]),
    "sepaihrd_host_det_exp": (C.c_double, [C.c_double]),
    "sepaihrd_host_mh_destroy": (None, [_vp]),
    "sepaihrd_host_pso_create": (C.c_int32, [_vp, C.c_int32, _keys, _vp, _vpp]),
    "sepaihrd_host_pso_begin": (C.c_int32, [_vp, _vp]),
    "sepaihrd_host_pso_local_count": (C.c_int32, [_vp]),
    "sepaihrd_host_pso_positions": (C.c_int32, [_vp, _vp]),
    "sepaihrd_host_pso_tell": (C.c_int32, [_vp, _vp, _dp, _i32p, _vp]),
    "sepaihrd_host_pso_set_global_best": (C.c_int32, [_vp, C.c_double, _vp]),
    "sepaihrd_host_pso_global_best": (C.c_int32, [_vp, _dp, _vp]),
    "sepaihrd_host_pso_step": (C.c_int32, [_vp, C.c_int32]),
    "sepaihrd_host_pso_begin_device": (C.c_int32, [_vp, _vp, _vp]),
    "sepaihrd_host_pso_evaluate_device": (C.c_int32, [_vp, _dp, _i32p, _vp]),
    "sepaihrd_host_pso_step_device": (C.c_int32, [_vp, C.c_int32]),
    "sepaihrd_host_pso_fetch": (C.c_int32, [_vp]),
    "sepaihrd_host_pso_run": (C.c_int32, [_vp, BATCH_FN, _vp, _vp, _vp, _dp, _vp]),
    "sepaihrd_host_pso_values": (C.c_int32, [_vp, _vp, _vp]),
    "sepaihrd_host_pso_neighbors": (C.c_int32, [_vp, C.c_int32, _vp, C.c_int32]),
    "sepaihrd_host_pso_destroy": (None, [_vp]),
    "sepaihrd_host_optimize": (C.c_int32, [C.c_char_p, _vp, C.c_int32, _keys, _vp, BATCH_FN, _vp, _vp, _vp, _dp, _i64p]),
    "sepaihrd_host_calibrate": (C.c_int32, [C.c_char_p, _vp, C.c_int32, _keys, _vp, C.c_int32, _keys, _vp, BATCH_FN, _vp, _vp, _vp, _dp, _i64p, _dp]),
    "sepaihrd_host_model_create": (C.c_int32, [_vp, _keys, _vp, _vpp]),
    "sepaihrd_host_model_calculate": (C.c_int32, [_vp, _vp, _dp]),
    "sepaihrd_host_model_calculate_batch": (C.c_int32, [_vp, _vp, C.c_int64, C.c_int64, _vp]),
    "sepaihrd_host_model_set_constraint_mode": (C.c_int32, [_vp, C.c_int32]),
    "sepaihrd_host_model_current_parameters": (C.c_int32, [_vp, _vp]),
    "sepaihrd_host_model_update_parameters": (C.c_int32, [_vp, _vp]),
    "sepaihrd_host_model_simulate": (C.c_int32, [_vp, _vp, _vp, C.c_int32, _vp]),
    "sepaihrd_host_model_calibrate": (C.c_int32, [_vp, C.c_char_p, C.c_int32, _keys, _vp, C.c_int32, _keys, _vp, _vp, _dp, _i64p]),
    "sepaihrd_host_model_metropolis": (C.c_int32, [_vp, C.c_int32, _keys, _vp, _vp, _vp, _dp, _vp, _vp]),
    "sepaihrd_host_model_posterior_predictive": (C.c_int32, [_vp, _vp, C.c_int64, C.c_int32, C.c_uint32, _vp, _vp, _i64p]),
    "sepaihrd_host_model_gradient": (C.c_int32, [_vp, _vp, C.c_double, _dp, _vp]),
    "sepaihrd_host_write_posterior_predictive": (C.c_int32, [C.c_char_p, C.c_int32, C.c_int32, _vp, _vp, _vp]),
    "sepaihrd_host_write_parameter_posteriors": (C.c_int32, [C.c_char_p, _vp, C.c_int64, C.c_int32, _keys, C.c_int32, C.c_int32]),
    "sepaihrd_host_model_set_cache": (C.c_int32, [_vp, C.c_int64]),
    "sepaihrd_host_model_cache_stats": (C.c_int32, [_vp, _vp]),
    "sepaihrd_host_model_destroy": (None, [_vp]),
    "sepaihrd_host_cache_create": (C.c_int32, [C.c_int64, _vpp]),
    "sepaihrd_host_cache_hash": (C.c_uint64, [_vp, _vp, C.c_int32]),
    "sepaihrd_host_cache_get": (C.c_int32, [_vp, C.c_uint64, _dp]),
    "sepaihrd_host_cache_store": (None, [_vp, C.c_uint64, C.c_double]),
    "sepaihrd_host_cache_get_vector": (C.c_int32, [_vp, _vp, C.c_int32, _dp]),
    "sepaihrd_host_cache_set_vector": (None, [_vp, _vp, C.c_int32, C.c_double]),
    "sepaihrd_host_cache_batch": (C.c_int32, [_vp, C.c_int32, _vp, C.c_int64, C.c_int64, _vp, BATCH_FN, _vp, _vp]),
    "sepaihrd_host_cache_size": (C.c_int64, [_vp]),
    "sepaihrd_host_cache_clear": (None, [_vp]),
    "sepaihrd_host_cache_stats": (None, [_vp, _vp]),
    "sepaihrd_host_cache_destroy": (None, [_vp]),
    "sepaihrd_host_metrics": (C.c_int32, [_vp, _vp, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sepaihrd_host_model_scenarios": (C.c_int32, [_vp, _vp, C.c_int64, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _vp, C.c_char_p]),
    "sepaihrd_host_model_analyze_runs": (C.c_int32, [_vp, _vp, C.c_int64, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _i64p]),
    "sepaihrd_host_read_file_json": (C.c_char_p, [C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.c_char_p, C.c_char_p]),
    "sepaihrd_host_project_json": (C.c_char_p, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32]),
    "sepaihrd_host_resave_parameters": (C.c_int32, [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, _keys, C.c_double, C.c_char_p]),
}


class HostError(RuntimeError):
    pass


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (builds csrc/ then host/)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise HostError((load_library().sepaihrd_host_last_error() or b"").decode())


def set_trace_directory(path: Optional[str]):
    """Where whole Metropolis-Hastings runs write posterior_trace*.csv (None: the reference's project-root rule)."""
    check(load_library().sepaihrd_host_set_trace_directory(path.encode() if path else None))


def set_threads(n: int) -> int:
    """OpenMP threads of the C++ sampler loops (torchrun pins OMP_NUM_THREADS=1 per rank)."""
    return int(load_library().sepaihrd_host_set_threads(int(n)))


def _c64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _settings(settings: Dict[str, float]):
    keys = (C.c_char_p * len(settings))(*[k.encode() for k in settings])
    vals = _c64(list(settings.values()))
    return len(settings), keys, vals


class ParameterManager:
    """IParameterManager over arrays: proposal sigmas, bounds (NaN = none), clamp (0) / reflect (1)."""

    def __init__(self, sigmas, lower, upper, mode: int = 0):
        self.L = load_library()
        self.sigmas, self.lower, self.upper = _c64(sigmas), _c64(lower), _c64(upper)
        self.n = len(self.sigmas)
        h = C.c_void_p()
        check(self.L.sepaihrd_host_pm_create(self.n, self.sigmas.ctypes.data, self.lower.ctypes.data, self.upper.ctypes.data, int(mode), C.byref(h)))
        self._h = h

    def set_mode(self, mode: int):
        check(self.L.sepaihrd_host_pm_set_mode(self._h, int(mode)))

    def apply_constraints(self, x) -> np.ndarray:
        x = _c64(x)
        out = np.empty_like(x)
        check(self.L.sepaihrd_host_pm_apply_constraints(self._h, x.ctypes.data, out.ctypes.data))
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.sepaihrd_host_pm_destroy(self._h)
            self._h = None


class MultiChainMH:
    """Step-wise handle on the batched MetropolisHastingsSampler (host/optimizers.hpp)."""

    def __init__(self, pm: ParameterManager, settings: Dict[str, float]):
        self.L, self.pm = pm.L, pm
        self.n_chains = int(settings.get("n_chains", 1))
        n, keys, vals = _settings(settings)
        h = C.c_void_p()
        check(self.L.sepaihrd_host_mh_create(pm._h, n, keys, vals.ctypes.data, C.byref(h)))
        self._h = h

    def set_initial_covariance(self, cov):
        cov = np.asfortranarray(np.asarray(cov, dtype=np.float64))
        check(self.L.sepaihrd_host_mh_set_initial_covariance(self._h, cov.ctypes.data, cov.shape[0]))

    def begin(self, initial, initial_logpost):
        x = _c64(initial); lp = _c64(initial_logpost)
        assert lp.shape == (self.n_chains,)
        check(self.L.sepaihrd_host_mh_begin(self._h, x.ctypes.data, lp.ctypes.data))

    @property
    def done(self) -> bool:
        return bool(self.L.sepaihrd_host_mh_done(self._h))

    @property
    def iteration(self) -> int:
        return int(self.L.sepaihrd_host_mh_iteration(self._h))

    def propose(self) -> np.ndarray:
        out = np.empty((self.n_chains, self.pm.n))
        check(self.L.sepaihrd_host_mh_propose(self._h, out.ctypes.data))
        return out

    def accept(self, proposed_logpost) -> np.ndarray:
        lp = _c64(proposed_logpost)
        acc = np.zeros(self.n_chains, dtype=np.uint8)
        check(self.L.sepaihrd_host_mh_accept(self._h, lp.ctypes.data, acc.ctypes.data))
        return acc

    def state(self):
        x = np.empty((self.n_chains, self.pm.n)); lp = np.empty(self.n_chains); sc = np.empty(self.n_chains)
        acc = np.zeros(self.n_chains, dtype=np.int64)
        check(self.L.sepaihrd_host_mh_state(self._h, x.ctypes.data, lp.ctypes.data, sc.ctypes.data, acc.ctypes.data))
        return x, lp, sc, acc

    def best(self):
        x = np.empty(self.pm.n); v = C.c_double()
        check(self.L.sepaihrd_host_mh_best(self._h, x.ctypes.data, C.byref(v)))
        return x, v.value

    def shared_cholesky(self) -> np.ndarray:
        """Lower Cholesky factor of the start kernel (after begin()), as the sampler factorised it."""
        out = np.empty((self.pm.n, self.pm.n), order="F")
        check(self.L.sepaihrd_host_mh_shared_cholesky(self._h, out.ctypes.data))
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.sepaihrd_host_mh_destroy(self._h)
            self._h = None


# the swarm the step-wise (shardable) and device-resident interfaces implement; the class defaults are the reference's
# (ADAPTIVE variant, opposition learning, adaptive parameters: ParticleSwarmOptimizer.hpp:201-213)
BASIC_SWARM = dict(variant=0, topology=0, use_opposition_learning=0, use_adaptive_parameters=0)


class Swarm:
    """Handle on the batched ParticleSwarmOptimization.  Step-wise (begin / tell / step and the device forms): this process owns
    one shard of the STANDARD / GLOBAL_BEST swarm -- BASIC_SWARM is the default of this wrapper.  run(): a whole optimize() with
    any of the reference's variants, topologies and strategies (pass them in ``settings``)."""

    def __init__(self, pm: ParameterManager, settings: Dict[str, float]):
        self.L, self.pm = pm.L, pm
        settings = {**BASIC_SWARM, **settings}
        n, keys, vals = _settings(settings)
        h = C.c_void_p()
        check(self.L.sepaihrd_host_pso_create(pm._h, n, keys, vals.ctypes.data, C.byref(h)))
        self._h = h
        self.iterations = int(settings.get("iterations", 100))

    def begin(self, initial=None):
        x = None if initial is None else _c64(initial)
        check(self.L.sepaihrd_host_pso_begin(self._h, None if x is None else x.ctypes.data))
        self.local = int(self.L.sepaihrd_host_pso_local_count(self._h))

    def positions(self) -> np.ndarray:
        out = np.empty((self.local, self.pm.n))
        check(self.L.sepaihrd_host_pso_positions(self._h, out.ctypes.data))
        return out

    def tell(self, fitness):
        f = _c64(fitness)
        v = C.c_double(); i = C.c_int32(); pos = np.empty(self.pm.n)
        check(self.L.sepaihrd_host_pso_tell(self._h, f.ctypes.data, C.byref(v), C.byref(i), pos.ctypes.data))
        return v.value, i.value, pos

    def set_global_best(self, value: float, position):
        p = _c64(position)
        check(self.L.sepaihrd_host_pso_set_global_best(self._h, float(value), p.ctypes.data))

    def global_best(self):
        v = C.c_double(); pos = np.empty(self.pm.n)
        check(self.L.sepaihrd_host_pso_global_best(self._h, C.byref(v), pos.ctypes.data))
        return v.value, pos

    def step(self, it: int):
        check(self.L.sepaihrd_host_pso_step(self._h, int(it)))

    # ---- device-resident form: the shard lives in the HBM of the evaluator's GPU (csrc/sepaihrd_swarm.cu) --------
    def begin_device(self, ctx_handle, initial=None):
        """ctx_handle: the sepaihrd_ctx* of the evaluator that scores the particles (BatchEvaluator.handle)."""
        x = None if initial is None else _c64(initial)
        check(self.L.sepaihrd_host_pso_begin_device(self._h, None if x is None else x.ctypes.data, ctx_handle))
        self.local = int(self.L.sepaihrd_host_pso_local_count(self._h))

    def evaluate_device(self):
        """Objective launch over the shard + personal-best update + arg-max: (best value, local index, position)."""
        v = C.c_double(); i = C.c_int32(); pos = np.empty(self.pm.n)
        check(self.L.sepaihrd_host_pso_evaluate_device(self._h, C.byref(v), C.byref(i), pos.ctypes.data))
        return v.value, i.value, pos

    def step_device(self, it: int):
        check(self.L.sepaihrd_host_pso_step_device(self._h, int(it)))

    def fetch(self):
        """Copy positions / velocities / personal bests back from the device (positions() is then current)."""
        check(self.L.sepaihrd_host_pso_fetch(self._h))

    # ---- whole runs: every variant / topology / strategy of the reference class ---------------------------------------
    def run(self, evaluate: Callable[[np.ndarray], np.ndarray], initial=None):
        """ParticleSwarmOptimization::optimize with ``evaluate`` ([B, P] -> [B]) as the objective.
        Returns (best position, best value, dict(evaluations, restarts, elitist_trials, diversity))."""
        P = self.pm.n
        cb = BATCH_FN(_batch_callback(evaluate, P))
        x0 = None if initial is None else _c64(initial)
        best = np.empty(P); val = C.c_double(); stats = np.zeros(4)
        check(self.L.sepaihrd_host_pso_run(self._h, cb, None, None if x0 is None else x0.ctypes.data, best.ctypes.data, C.byref(val),
                                           stats.ctypes.data))
        return best, val.value, dict(evaluations=int(stats[0]), restarts=int(stats[1]), elitist_trials=int(stats[2]), diversity=float(stats[3]))

    def values(self, swarm_size: int):
        """(personal-best values, current fitness) of the swarm after run()."""
        pb = np.empty(swarm_size); cf = np.empty(swarm_size)
        check(self.L.sepaihrd_host_pso_values(self._h, pb.ctypes.data, cf.ctypes.data))
        return pb, cf

    def neighbors(self, particle: int) -> List[int]:
        """getNeighbors of the configured topology (RANDOM_DYNAMIC draws from the master generator on every call)."""
        out = np.empty(64, dtype=np.int32)
        k = self.L.sepaihrd_host_pso_neighbors(self._h, int(particle), out.ctypes.data, 64)
        if k < 0:
            raise HostError(self.L.sepaihrd_host_last_error().decode())
        return [int(v) for v in out[:min(k, 64)]]

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.sepaihrd_host_pso_destroy(self._h)
            self._h = None


class Cache:
    """SimulationCache of the host mirror (LFU / LRU, open addressing, keys = hash of the vector quantised at 1e-8)."""

    def __init__(self, capacity: int = 1000):
        self.L = load_library()
        h = C.c_void_p()
        check(self.L.sepaihrd_host_cache_create(int(capacity), C.byref(h)))
        self._h = h

    def hash(self, params) -> int:
        x = _c64(params)
        return int(self.L.sepaihrd_host_cache_hash(self._h, x.ctypes.data, len(x)))

    def get(self, key: int):
        v = C.c_double()
        return v.value if self.L.sepaihrd_host_cache_get(self._h, int(key), C.byref(v)) else None

    def store(self, key: int, value: float):
        self.L.sepaihrd_host_cache_store(self._h, int(key), float(value))

    def get_vector(self, params):
        x = _c64(params); v = C.c_double()
        return v.value if self.L.sepaihrd_host_cache_get_vector(self._h, x.ctypes.data, len(x), C.byref(v)) else None

    def set_vector(self, params, value: float):
        x = _c64(params)
        self.L.sepaihrd_host_cache_set_vector(self._h, x.ctypes.data, len(x), float(value))

    def batch(self, params, evaluate, status_of_row=None) -> np.ndarray:
        """What SEPAIHRDObjectiveFunction::calculateBatch does with its cache, with ``evaluate`` ([M, P] -> [M]) as the device:
        status_of_row (optional uint32 array) is indexed by each row's first coordinate."""
        x = _c64(params)
        out = np.empty(len(x))
        cb = BATCH_FN(_batch_callback(evaluate, x.shape[1]))
        st = None if status_of_row is None else np.ascontiguousarray(status_of_row, dtype=np.uint32)
        check(self.L.sepaihrd_host_cache_batch(self._h, x.shape[1], x.ctypes.data, x.shape[0], x.shape[1], out.ctypes.data, cb, None,
                                               None if st is None else st.ctypes.data))
        return out

    def __len__(self) -> int:
        return int(self.L.sepaihrd_host_cache_size(self._h))

    def clear(self):
        self.L.sepaihrd_host_cache_clear(self._h)

    def stats(self):
        out = np.zeros(3, dtype=np.int64)
        self.L.sepaihrd_host_cache_stats(self._h, out.ctypes.data)
        return dict(get_calls=int(out[0]), hits=int(out[1]), store_calls=int(out[2]))

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.sepaihrd_host_cache_destroy(self._h)
            self._h = None


def write_posterior_predictive(output_dir: str, time_points, quantiles, observed=None):
    """AnalysisWriter::savePosteriorPredictiveData: quantiles [6, T, n, 5] (HostModel.posterior_predictive), observed [6, T, n]."""
    q = _c64(quantiles); t = _c64(time_points)
    obs = None if observed is None else _c64(observed)
    check(load_library().sepaihrd_host_write_posterior_predictive(output_dir.encode(), q.shape[1], q.shape[2], t.ctypes.data, q.ctypes.data,
                                                                   None if obs is None else obs.ctypes.data))


def write_parameter_posteriors(output_dir: str, samples, names: Sequence[str], burn_in: int = 0, thinning: int = 1):
    """AnalysisWriter::saveParameterPosteriors: posterior_samples.csv and posterior_summary.csv."""
    x = _c64(samples)
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    check(load_library().sepaihrd_host_write_parameter_posteriors(output_dir.encode(), x.ctypes.data, x.shape[0], x.shape[1], arr, int(burn_in), int(thinning)))


def _batch_callback(evaluate, P):
    def _cb(_user, params, B, ld, out):
        try:
            x = np.ctypeslib.as_array(params, shape=(B, ld))[:, :P]
            np.ctypeslib.as_array(out, shape=(B,))[:] = evaluate(np.ascontiguousarray(x))
            return 0
        except Exception:      # never unwind through the C++ frames
            import traceback
            traceback.print_exc()
            return 1
    return _cb


def optimize(algorithm: str, pm: ParameterManager, settings: Dict[str, float], evaluate: Callable[[np.ndarray], np.ndarray], initial):
    """IOptimizationAlgorithm::optimize ("mh", "pso", "hill") with ``evaluate`` ([B, P] -> [B]) as the objective."""
    L = pm.L
    P = pm.n
    cb = BATCH_FN(_batch_callback(evaluate, P))
    n, keys, vals = _settings(settings)
    x0 = _c64(initial)
    best = np.empty(P); val = C.c_double(); nev = C.c_int64()
    check(L.sepaihrd_host_optimize(algorithm.encode(), pm._h, n, keys, vals.ctypes.data, cb, None, x0.ctypes.data, best.ctypes.data,
                                   C.byref(val), C.byref(nev)))
    return best, val.value, nev.value


def calibrate(phase1: str, pm: ParameterManager, settings1: Dict[str, float], settings2: Dict[str, float],
              evaluate: Callable[[np.ndarray], np.ndarray], initial):
    """ModelCalibrator::calibrate (phase 1 "pso" | "hill", phase 2 Metropolis-Hastings) with ``evaluate`` as the objective.
    Returns (best vector, best value, number of re-scored MCMC samples, best value of phase 1)."""
    L, P = pm.L, pm.n

    def _cb(_user, params, B, ld, out):
        try:
            x = np.ctypeslib.as_array(params, shape=(B, ld))[:, :P]
            np.ctypeslib.as_array(out, shape=(B,))[:] = evaluate(np.ascontiguousarray(x))
            return 0
        except Exception:
            import traceback
            traceback.print_exc()
            return 1

    cb = BATCH_FN(_cb)
    n1, k1, v1 = _settings(settings1); n2, k2, v2 = _settings(settings2)
    x0 = _c64(initial)
    best = np.empty(P); val = C.c_double(); ns = C.c_int64(); p1 = C.c_double()
    check(L.sepaihrd_host_calibrate(phase1.encode(), pm._h, n1, k1, v1.ctypes.data, n2, k2, v2.ctypes.data, cb, None, x0.ctypes.data,
                                    best.ctypes.data, C.byref(val), C.byref(ns), C.byref(p1)))
    return best, val.value, ns.value, p1.value


NUM_METRICS = 12
METRIC_NAMES = ["R0", "overall_IFR", "overall_attack_rate", "peak_hospital_occupancy", "peak_ICU_occupancy", "time_to_peak_hospital",
                "time_to_peak_ICU", "total_cumulative_deaths", "max_Rt", "min_Rt", "final_Rt", "seroprevalence_at_target_day"]


def essential_metrics(problem, times, trajectory, initial_state, trajectories: bool = False):
    """MetricsCalculator::calculateEssentialMetrics on a given trajectory [K, 11 n] of the problem's base model (host
    arithmetic only).  Returns (scalars[12], age[4, n]) or, with trajectories=True, also (Rt[K], seroprevalence[K])."""
    L = load_library()
    cp = problem.as_c()
    t = _c64(times); tr = _c64(trajectory); s0 = _c64(initial_state)
    sc = np.empty(NUM_METRICS); age = np.empty((4, problem.n_ages)); rt = np.empty(len(t)); se = np.empty(len(t))
    check(L.sepaihrd_host_metrics(C.addressof(cp), t.ctypes.data, len(t), tr.ctypes.data, s0.ctypes.data, sc.ctypes.data, age.ctypes.data,
                                  rt.ctypes.data if trajectories else None, se.ctypes.data if trajectories else None))
    return (sc, age, rt, se) if trajectories else (sc, age)


class HostModel:
    """The reference-shaped C++ object graph (AgeSEPAIHRDModel, SEPAIHRDParameterManager, SEPAIHRDObjectiveFunction,
    AgeSEPAIHRDSimulator, SEPAIHRDModelCalibration) over the device evaluator.  Needs a CUDA device."""

    def __init__(self, problem):
        self.L = load_library()
        self.problem = problem
        self._cp = problem.as_c()
        names = (C.c_char_p * problem.n_params)(*[n.encode() for n in problem.param_names])
        sig = _c64(problem.sigmas)
        h = C.c_void_p()
        check(self.L.sepaihrd_host_model_create(C.addressof(self._cp), names, sig.ctypes.data, C.byref(h)))
        self._h = h

    def calculate(self, params) -> float:
        x = _c64(params); v = C.c_double()
        check(self.L.sepaihrd_host_model_calculate(self._h, x.ctypes.data, C.byref(v)))
        return v.value

    def calculate_batch(self, params) -> np.ndarray:
        x = _c64(params); out = np.empty(x.shape[0])
        check(self.L.sepaihrd_host_model_calculate_batch(self._h, x.ctypes.data, x.shape[0], x.shape[1], out.ctypes.data))
        return out

    def gradient(self, params, epsilon: float = 0.0):
        """SEPAIHRDGradientObjectiveFunction::evaluate_with_gradient: (value, forward-difference gradient)."""
        x = _c64(params); g = np.empty(len(x)); v = C.c_double()
        check(self.L.sepaihrd_host_model_gradient(self._h, x.ctypes.data, float(epsilon), C.byref(v), g.ctypes.data))
        return v.value, g

    def set_cache(self, capacity: int):
        """capacity > 0: evaluate through a SimulationCache of that capacity (the reference's main() uses 1000); 0: no cache."""
        check(self.L.sepaihrd_host_model_set_cache(self._h, int(capacity)))

    def cache_stats(self):
        out = np.zeros(4, dtype=np.int64)
        check(self.L.sepaihrd_host_model_cache_stats(self._h, out.ctypes.data))
        return dict(entries=int(out[0]), get_calls=int(out[1]), hits=int(out[2]), store_calls=int(out[3]))

    def set_constraint_mode(self, mode: int):
        check(self.L.sepaihrd_host_model_set_constraint_mode(self._h, int(mode)))

    def current_parameters(self) -> np.ndarray:
        out = np.empty(self.problem.n_params)
        check(self.L.sepaihrd_host_model_current_parameters(self._h, out.ctypes.data))
        return out

    def update_parameters(self, params):
        x = _c64(params)
        check(self.L.sepaihrd_host_model_update_parameters(self._h, x.ctypes.data))

    def simulate(self, initial_state, times) -> np.ndarray:
        s0 = _c64(initial_state); t = _c64(times)
        out = np.empty((len(t), self.problem.state_size))
        check(self.L.sepaihrd_host_model_simulate(self._h, s0.ctypes.data, t.ctypes.data, len(t), out.ctypes.data))
        return out

    def calibrate(self, phase1: str, settings1: Dict[str, float], settings2: Dict[str, float]):
        n1, k1, v1 = _settings(settings1); n2, k2, v2 = _settings(settings2)
        best = np.empty(self.problem.n_params); val = C.c_double(); ns = C.c_int64()
        check(self.L.sepaihrd_host_model_calibrate(self._h, phase1.encode(), n1, k1, v1.ctypes.data, n2, k2, v2.ctypes.data,
                                                   best.ctypes.data, C.byref(val), C.byref(ns)))
        return best, val.value, ns.value

    def metropolis(self, settings: Dict[str, float], initial):
        """MetropolisHastingsSampler::optimize on the device objective, timed in C++.  Returns dict(best, best_value, last,
        last_logpost, ms, evaluations, launches, acceptance_rate, final_scale, iterations)."""
        n, k, v = _settings(settings)
        x0 = _c64(initial); P = self.problem.n_params
        best = np.empty(P); last = np.empty(P + 1); val = C.c_double(); st = np.zeros(6)
        check(self.L.sepaihrd_host_model_metropolis(self._h, n, k, v.ctypes.data, x0.ctypes.data, best.ctypes.data, C.byref(val),
                                                    last.ctypes.data, st.ctypes.data))
        return dict(best=best, best_value=val.value, last=last[:P].copy(), last_logpost=float(last[P]), ms=float(st[0]), evaluations=int(st[1]),
                    launches=int(st[2]), acceptance_rate=float(st[3]), final_scale=float(st[4]), iterations=int(st[5]))

    def posterior_predictive(self, samples, initial_state, num_samples: int = 0, seed: int = 0):
        """ResultAggregator::aggregatePosteriorPredictives: [6, T, n, 5] (lower_95, lower_90, median, upper_90, upper_95)."""
        x = _c64(samples); s0 = _c64(initial_state)
        T = int((self.problem.times >= 0).sum())
        out = np.empty((6, T, self.problem.n_ages, 5)); used = C.c_int64()
        check(self.L.sepaihrd_host_model_posterior_predictive(self._h, x.ctypes.data, x.shape[0], int(num_samples), int(seed), s0.ctypes.data,
                                                              out.ctypes.data, C.byref(used)))
        return out, used.value

    def scenarios(self, samples, initial_state, burn_in: int = 0, thinning: int = 1, csv_path: str = "", trajectories: bool = False):
        """PostCalibrationAnalyser scenario analysis (baseline / stricter_lockdown / weaker_lockdown) as one device batch.
        Returns dict(names, scalars [3, 12], age [3, 4, n], kappa [3, nk][, trajectories [3, K, 11 n]])."""
        x = _c64(samples); s0 = _c64(initial_state)
        n, nk, K = self.problem.n_ages, len(self.problem.kappa_end_times), self.problem.n_times
        sc = np.empty((3, NUM_METRICS)); age = np.empty((3, 4, n)); kap = np.empty((3, nk))
        traj = np.empty((3, K, self.problem.state_size)) if trajectories else None
        check(self.L.sepaihrd_host_model_scenarios(self._h, x.ctypes.data, x.shape[0], int(burn_in), int(thinning), s0.ctypes.data,
                                                   sc.ctypes.data, age.ctypes.data, kap.ctypes.data,
                                                   traj.ctypes.data if trajectories else None, os.fsencode(csv_path)))
        out = dict(names=["baseline", "stricter_lockdown", "weaker_lockdown"], scalars=sc, age=age, kappa=kap)
        if trajectories:
            out["trajectories"] = traj
        return out

    def analyze_runs(self, samples, initial_state, burn_in: int = 0, thinning: int = 1):
        """analyzeMCMCRunsInBatches without the files: metrics [runs, 12], Rt and seroprevalence quantiles [5, K]."""
        x = _c64(samples); s0 = _c64(initial_state)
        runs = len(range(int(burn_in), x.shape[0], max(int(thinning), 1)))
        K = self.problem.n_times
        sc = np.empty((max(runs, 1), NUM_METRICS)); rt = np.empty((5, K)); se = np.empty((5, K)); got = C.c_int64()
        check(self.L.sepaihrd_host_model_analyze_runs(self._h, x.ctypes.data, x.shape[0], int(burn_in), int(thinning), s0.ctypes.data,
                                                      sc.ctypes.data, rt.ctypes.data, se.ctypes.data, C.byref(got)))
        return sc[:got.value], rt, se

    def close(self):
        if getattr(self, "_h", None):
            self.L.sepaihrd_host_model_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


# ---- the reference's on-disk formats through the C++ readers (host/config_io.hpp) ---------------------------------
def read_file(kind: str, path: str, a: int = 0, b: int = 0, start_date: str = "", end_date: str = "") -> dict:
    """kind: parameters (a = ages) | bounds | sigmas | names | settings | matrix (a x b) | data (date window).
    Raises HostError with the C++ exception text (FileIOException / DataFormatException / CSVReadException)."""
    import json
    txt = load_library().sepaihrd_host_read_file_json(kind.encode(), os.fsencode(path), int(a), int(b), start_date.encode(), end_date.encode())
    if txt is None:
        check(1)
    return json.loads(txt.decode())


def load_reference_project(root: str, start_date: str = "2020-03-01", end_date: str = "2020-12-31", n_ages: int = 4) -> dict:
    """loadReferenceProject (main.cpp:188-316) in C++; the dict has the keys of Problem.to_json() plus initial_state."""
    import json
    txt = load_library().sepaihrd_host_project_json(os.fsencode(root), start_date.encode(), end_date.encode(), int(n_ages))
    if txt is None:
        check(1)
    return json.loads(txt.decode())


def resave_parameters(in_file: str, n_ages: int, out_file: str, calibrated: List[str], obj_value: float, timestamp: str = "") -> None:
    """readSEPAIHRDParameters -> saveCalibrationResults."""
    names = (C.c_char_p * max(len(calibrated), 1))(*[c.encode() for c in calibrated])
    check(load_library().sepaihrd_host_resave_parameters(os.fsencode(in_file), int(n_ages), os.fsencode(out_file), len(calibrated), names,
                                                         float(obj_value), timestamp.encode()))
