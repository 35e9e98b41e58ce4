"""ctypes binding of the CPU oracle (oracle/libsepaihrd_oracle.so).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsepaihrd_oracle.so")
_lib = None

_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (g++ -O2 -ffp-contract=off -fopenmp)."""
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("sepaihrd_oracle.cpp", "sepaihrd_oracle.h", "Makefile")):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.sepaihrd_oracle_slot_for_name.restype = C.c_int32
        L.sepaihrd_oracle_slot_for_name.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_char_p]
        L.sepaihrd_oracle_apply_constraints.restype = None
        L.sepaihrd_oracle_apply_constraints.argtypes = [C.c_void_p, C.c_int32, _dp, _dp]
        L.sepaihrd_oracle_rhs.restype = None
        L.sepaihrd_oracle_rhs.argtypes = [C.c_void_p, _dp, _dp, C.c_double, _dp]
        L.sepaihrd_oracle_poisson_ll.restype = C.c_double
        L.sepaihrd_oracle_poisson_ll.argtypes = [_dp, _dp, C.c_int32, C.c_int32]
        L.sepaihrd_oracle_initial_state_from_data.restype = None
        L.sepaihrd_oracle_initial_state_from_data.argtypes = [C.c_int32, _dp, _dp, _dp, _dp, _dp, C.c_double,
                                                              C.c_double, C.c_double, C.c_double, _dp, _dp]
        L.sepaihrd_oracle_eval_one.restype = C.c_double
        L.sepaihrd_oracle_eval_one.argtypes = [C.c_void_p, _dp, C.POINTER(C.c_uint32), _dp,
                                               C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        L.sepaihrd_oracle_eval_batch.restype = C.c_int32
        L.sepaihrd_oracle_eval_batch.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int64, _dp,
                                                 C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.c_int32]
        L.sepaihrd_oracle_simulate_batch.restype = C.c_int32
        L.sepaihrd_oracle_simulate_batch.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                                     _dp, C.POINTER(C.c_uint32), C.c_int32]
        L.sepaihrd_oracle_simulate_from_state.restype = C.c_int32
        L.sepaihrd_oracle_simulate_from_state.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int64, _dp, C.c_int64, C.c_int32,
                                                          C.c_int32, _dp, C.POINTER(C.c_uint32), C.c_int32]
        L.sepaihrd_oracle_jitter_params.restype = None
        L.sepaihrd_oracle_jitter_params.argtypes = [C.c_void_p, _dp, _dp, C.c_uint32, C.c_int64, _dp]
        L.sepaihrd_oracle_uniform_params.restype = None
        L.sepaihrd_oracle_uniform_params.argtypes = [C.c_void_p, C.c_uint32, C.c_int64, _dp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    """CPU oracle bound to one Problem."""

    def __init__(self, problem, constraint_mode=None):
        self.problem = problem
        self._cp = problem.as_c(constraint_mode)
        self._ref = C.byref(self._cp)
        self.L = lib()

    def set_constraint_mode(self, mode: int):
        self._cp.constraint_mode = int(mode)

    def slot_for_name(self, name: str) -> int:
        p = self.problem
        return self.L.sepaihrd_oracle_slot_for_name(p.n_ages, len(p.beta_end_times), len(p.kappa_end_times),
                                                    name.encode())

    def apply_constraints(self, params, mode: int) -> np.ndarray:
        x = _c64(params); out = np.empty_like(x)
        self.L.sepaihrd_oracle_apply_constraints(self._ref, int(mode), _p(x), _p(out))
        return out

    def rhs(self, slots, state, t: float) -> np.ndarray:
        s = _c64(slots); x = _c64(state); out = np.empty_like(x)
        self.L.sepaihrd_oracle_rhs(self._ref, _p(s), _p(x), float(t), _p(out))
        return out

    def eval_one(self, params, want_traj=False, want_interval_steps=False):
        p = self.problem
        x = _c64(params)
        st = C.c_uint32(0)
        traj = np.empty((p.n_times, p.state_size)) if want_traj else None
        isteps = np.zeros((p.n_times - 1, 2), dtype=np.int32) if want_interval_steps else None
        counts = np.zeros(3, dtype=np.int64)
        ll = self.L.sepaihrd_oracle_eval_one(
            self._ref, _p(x), C.byref(st), _p(traj) if traj is not None else None,
            isteps.ctypes.data_as(C.POINTER(C.c_int32)) if isteps is not None else None,
            counts.ctypes.data_as(C.POINTER(C.c_int64)))
        return dict(ll=ll, status=st.value, traj=traj, interval_steps=isteps,
                    accepted=int(counts[0]), rejected=int(counts[1]), rhs_calls=int(counts[2]))

    def trace_one(self, params, cap: int = 8192):
        """(t, dt, err) of every step attempt of one evaluation: [attempts, 3]."""
        x = _c64(params)
        out = np.zeros((cap, 3))
        ll = C.c_double()
        f = self.L.sepaihrd_oracle_trace_one
        f.restype = C.c_int64
        f.argtypes = [C.c_void_p, _dp, C.c_int64, _dp, C.POINTER(C.c_double)]
        n = f(self._ref, _p(x), cap, _p(out), C.byref(ll))
        return out[:min(n, cap)], ll.value

    def eval_batch(self, params, nthreads: int = 0):
        x = _c64(params)
        B, ld = x.shape
        ll = np.empty(B); st = np.zeros(B, dtype=np.uint32); steps = np.zeros((B, 2), dtype=np.int32)
        used = self.L.sepaihrd_oracle_eval_batch(self._ref, _p(x), B, ld, _p(ll),
                                                 st.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                 steps.ctypes.data_as(C.POINTER(C.c_int32)), int(nthreads))
        return ll, st, steps, used

    def simulate_batch(self, params, what: int = 0, stride: int = 1, nthreads: int = 0):
        p = self.problem
        x = _c64(params)
        B, ld = x.shape
        W = p.state_size if what == 0 else 3 * p.n_ages
        Kout = (p.n_times + stride - 1) // stride
        out = np.empty((B, Kout, W)); st = np.zeros(B, dtype=np.uint32)
        self.L.sepaihrd_oracle_simulate_batch(self._ref, _p(x), B, ld, int(what), int(stride), _p(out),
                                              st.ctypes.data_as(C.POINTER(C.c_uint32)), int(nthreads))
        return out, st

    def simulate_from_state(self, params, initial_states, what: int = 0, stride: int = 1, nthreads: int = 0):
        p = self.problem
        x = _c64(params); s0 = _c64(initial_states)
        B, ld = x.shape
        sstride = 0 if s0.ndim == 1 else s0.shape[1]
        W = p.state_size if what == 0 else 3 * p.n_ages
        Kout = (p.n_times + stride - 1) // stride
        out = np.empty((B, Kout, W)); st = np.zeros(B, dtype=np.uint32)
        self.L.sepaihrd_oracle_simulate_from_state(self._ref, _p(x), B, ld, _p(s0), sstride, int(what), int(stride),
                                                   _p(out), st.ctypes.data_as(C.POINTER(C.c_uint32)), int(nthreads))
        return out, st

    def jitter_params(self, B: int, seed: int = 1, base=None, sigmas=None) -> np.ndarray:
        p = self.problem
        base = _c64(p.base_params() if base is None else base)
        sig = _c64(p.sigmas if sigmas is None else sigmas)
        out = np.empty((B, p.n_params))
        self.L.sepaihrd_oracle_jitter_params(self._ref, _p(base), _p(sig), int(seed), int(B), _p(out))
        return out

    def uniform_params(self, B: int, seed: int = 2) -> np.ndarray:
        out = np.empty((B, self.problem.n_params))
        self.L.sepaihrd_oracle_uniform_params(self._ref, int(seed), int(B), _p(out))
        return out


def poisson_ll(simulated, observed) -> float:
    s = _c64(simulated); o = _c64(observed)
    return lib().sepaihrd_oracle_poisson_ll(_p(s), _p(o), s.shape[0], s.shape[1])


def initial_state_from_data(population, cum_confirmed0, cum_deaths0, cum_hosp0, cum_icu0,
                            sigma, gamma_p, gamma_a, gamma_i, p_asym) -> np.ndarray:
    N = _c64(population); n = N.shape[0]
    out = np.empty(11 * n)
    lib().sepaihrd_oracle_initial_state_from_data(n, _p(N), _p(_c64(cum_confirmed0)), _p(_c64(cum_deaths0)),
                                                  _p(_c64(cum_hosp0)), _p(_c64(cum_icu0)), sigma, gamma_p,
                                                  gamma_a, gamma_i, _p(_c64(p_asym)), _p(out))
    return out
