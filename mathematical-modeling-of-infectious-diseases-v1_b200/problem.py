"""Problem description shared by the C ABI, the oracle binding, the tests and bench.py.

A :class:`Problem` is the Python image of ``struct sepaihrd_problem`` (include/sepaihrd_b200.h):
everything that is constant across parameter sets.  It can be

* built from a reference-style project tree (``data/configuration/*.txt``, ``data/contacts.csv``,
  ``data/processed/processed_data.csv``) with :func:`problem_from_reference_tree`, following the
  setup code of the reference (``src/model/main.cpp:188-316``), or
* loaded from / saved to a small JSON file (the committed Spain-2020 fixture lives in
  ``data/spain2020_problem.json`` next to this file).

Only numpy + ctypes here: no torch, no CUDA.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

NUM_COMPARTMENTS = 11          # reference include/model/ModelConstants.hpp:18
MAX_AGES = 16
ABI_VERSION = 1

CLAMP = 0                      # ConstraintMode::OPTIMIZATION_CLAMP (SEPAIHRDParameterManager.hpp:22-25)
REFLECT = 1                    # ConstraintMode::MCMC_REFLECT

ST_OK, ST_S_OVERFLOW, ST_STEP_FAILURE, ST_NONFINITE, ST_INVALID_PARAM = 0, 1, 2, 4, 8

TRAJ_FULL, TRAJ_OBSERVED = 0, 1

LOWEST = -float(np.finfo(np.float64).max)   # std::numeric_limits<double>::lowest()

_MULTIPLIERS = ["E0_multiplier", "P0_multiplier", "A0_multiplier", "I0_multiplier",
                "H0_multiplier", "ICU0_multiplier", "R0_multiplier", "D0_multiplier"]
_AGE_BLOCKS = ["a", "h_infec", "p", "h", "icu", "d_H", "d_ICU", "d_community"]
_SCALARS = ["theta", "sigma", "gamma_p", "gamma_A", "gamma_I", "gamma_H", "gamma_ICU"]


class SlotLayout:
    """Flat layout of the model parameters the hot path reads (see sepaihrd_b200.h)."""

    def __init__(self, n_ages: int, n_beta: int, n_kappa: int):
        self.n, self.nb, self.nk = n_ages, n_beta, n_kappa
        self.beta0 = 0
        self.kappa0 = n_beta
        self.scal0 = n_beta + n_kappa
        self.age0 = self.scal0 + len(_SCALARS)
        self.mult0 = self.age0 + len(_AGE_BLOCKS) * n_ages
        self.seed_exposed = self.mult0 + 8
        self.runup_days = self.mult0 + 9
        self.beta_scalar = self.mult0 + 10
        self.count = self.mult0 + 11

    def scalar(self, name: str) -> int:
        return self.scal0 + _SCALARS.index(name)

    def age(self, block: str, i: int) -> int:
        return self.age0 + _AGE_BLOCKS.index(block) * self.n + i

    def names(self) -> List[str]:
        """Canonical reference name of every slot (kappa_1 is the fixed baseline)."""
        out = [f"beta_{k + 1}" for k in range(self.nb)] + [f"kappa_{k + 1}" for k in range(self.nk)]
        out += list(_SCALARS)
        for blk in _AGE_BLOCKS:
            out += [f"{blk}_{i}" for i in range(self.n)]
        out += list(_MULTIPLIERS) + ["seed_exposed", "runup_days", "beta"]
        return out

    def slot_for_name(self, name: str) -> int:
        """Python restatement of sepaihrd_slot_for_name (same dispatch order as
        SEPAIHRDParameterManager::updateModelParameters, .cpp:197-267). -1 unknown, -2 rejected."""
        def idx(s: str) -> Optional[int]:
            return int(s) if s.isdigit() else None
        if name == "beta":
            return self.beta_scalar
        if name.startswith("beta_"):
            k = idx(name[5:])
            return self.beta0 + k - 1 if k is not None and 1 <= k <= self.nb else -2
        if name in _SCALARS:
            return self.scalar(name)
        for blk in ["a", "h_infec", "p", "h", "icu", "d_H", "d_ICU", "d_community"]:
            if name.startswith(blk + "_"):
                k = idx(name[len(blk) + 1:])
                return self.age(blk, k) if k is not None and k < self.n else -2
        if name == "seed_exposed":
            return self.seed_exposed
        if name == "runup_days":
            return self.runup_days
        if name in _MULTIPLIERS:
            return self.mult0 + _MULTIPLIERS.index(name)
        if name.startswith("kappa_"):
            k = idx(name[6:])
            return self.kappa0 + k - 1 if k is not None and 2 <= k <= self.nk else -2
        return -1


class CProblem(C.Structure):
    """ctypes mirror of struct sepaihrd_problem."""
    _fields_ = [
        ("abi_version", C.c_int32), ("n_ages", C.c_int32), ("n_times", C.c_int32), ("n_obs", C.c_int32),
        ("times", C.POINTER(C.c_double)),
        ("obs_hosp", C.POINTER(C.c_double)), ("obs_icu", C.POINTER(C.c_double)), ("obs_deaths", C.POINTER(C.c_double)),
        ("population", C.POINTER(C.c_double)), ("contact_matrix", C.POINTER(C.c_double)),
        ("n_beta", C.c_int32), ("n_kappa", C.c_int32),
        ("beta_end_times", C.POINTER(C.c_double)), ("kappa_end_times", C.POINTER(C.c_double)),
        ("base_slots", C.POINTER(C.c_double)), ("data_initial_state", C.POINTER(C.c_double)),
        ("n_params", C.c_int32), ("constraint_mode", C.c_int32),
        ("param_slot", C.POINTER(C.c_int32)),
        ("lower_bound", C.POINTER(C.c_double)), ("upper_bound", C.POINTER(C.c_double)),
        ("abs_tol", C.c_double), ("rel_tol", C.c_double), ("dt_hint", C.c_double),
    ]


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@dataclass
class Problem:
    n_ages: int
    times: np.ndarray                 # [K]
    obs_hosp: np.ndarray              # [n_obs, n]
    obs_icu: np.ndarray
    obs_deaths: np.ndarray
    population: np.ndarray            # [n]
    contact_matrix: np.ndarray        # [n, n]  M[i, j] (row i, col j) -- stored column-major for the ABI
    beta_end_times: np.ndarray
    kappa_end_times: np.ndarray
    base_slots: np.ndarray            # [slot_count]
    data_initial_state: np.ndarray    # [11 n]
    param_names: List[str]            # P calibrated names
    lower_bound: np.ndarray           # [P]
    upper_bound: np.ndarray
    sigmas: np.ndarray                # [P] proposal sigmas (used by the drivers / jitter recipe)
    constraint_mode: int = CLAMP
    abs_tol: float = 1e-6
    rel_tol: float = 1e-6
    dt_hint: float = 1.0
    meta: Dict[str, object] = field(default_factory=dict)

    def __post_init__(self):
        self.times = _f64(self.times)
        self.obs_hosp = _f64(self.obs_hosp); self.obs_icu = _f64(self.obs_icu); self.obs_deaths = _f64(self.obs_deaths)
        self.population = _f64(self.population)
        self.contact_matrix = _f64(self.contact_matrix)
        self.beta_end_times = _f64(self.beta_end_times); self.kappa_end_times = _f64(self.kappa_end_times)
        self.base_slots = _f64(self.base_slots)
        self.data_initial_state = _f64(self.data_initial_state)
        self.lower_bound = _f64(self.lower_bound); self.upper_bound = _f64(self.upper_bound)
        self.sigmas = _f64(self.sigmas)
        lay = self.layout
        if self.base_slots.shape != (lay.count,):
            raise ValueError(f"base_slots must have {lay.count} entries, got {self.base_slots.shape}")
        self.param_slot = np.ascontiguousarray([lay.slot_for_name(nm) for nm in self.param_names], dtype=np.int32)
        bad = [nm for nm, s in zip(self.param_names, self.param_slot) if s == -2]
        if bad:
            # the reference throws InvalidParameterException at SEPAIHRDParameterManager construction (.cpp:45-88)
            raise ValueError(f"parameter names rejected by the reference parameter manager: {bad}")
        self._c_keepalive = None

    # ------------------------------------------------------------------------------------------
    @property
    def layout(self) -> SlotLayout:
        return SlotLayout(self.n_ages, len(self.beta_end_times), len(self.kappa_end_times))

    @property
    def n_params(self) -> int:
        return len(self.param_names)

    @property
    def n_times(self) -> int:
        return int(self.times.shape[0])

    @property
    def n_obs(self) -> int:
        return int(self.obs_hosp.shape[0])

    @property
    def state_size(self) -> int:
        return NUM_COMPARTMENTS * self.n_ages

    def base_params(self) -> np.ndarray:
        """SEPAIHRDParameterManager::getCurrentParameters (.cpp:91-158): the calibrated subset of the
        base model parameters, in param_names order."""
        out = np.empty(self.n_params)
        for i, s in enumerate(self.param_slot):
            out[i] = self.base_slots[s] if s >= 0 else 0.0
        return out

    def as_c(self, constraint_mode: Optional[int] = None) -> CProblem:
        """ctypes struct pointing into this object's arrays (kept alive by self)."""
        m_colmajor = np.ascontiguousarray(self.contact_matrix.T).reshape(-1)   # data[j*n+i] = M[i, j]
        keep = dict(times=self.times, oh=self.obs_hosp.reshape(-1), oi=self.obs_icu.reshape(-1),
                    od=self.obs_deaths.reshape(-1), N=self.population, M=m_colmajor,
                    bt=self.beta_end_times, kt=self.kappa_end_times, base=self.base_slots,
                    init=self.data_initial_state, slot=self.param_slot, lo=self.lower_bound, hi=self.upper_bound)
        keep = {k: np.ascontiguousarray(v) for k, v in keep.items()}
        cp = CProblem()
        cp.abi_version = ABI_VERSION
        cp.n_ages = self.n_ages; cp.n_times = self.n_times; cp.n_obs = self.n_obs
        cp.times = _dptr(keep["times"])
        cp.obs_hosp = _dptr(keep["oh"]); cp.obs_icu = _dptr(keep["oi"]); cp.obs_deaths = _dptr(keep["od"])
        cp.population = _dptr(keep["N"]); cp.contact_matrix = _dptr(keep["M"])
        cp.n_beta = len(self.beta_end_times); cp.n_kappa = len(self.kappa_end_times)
        cp.beta_end_times = _dptr(keep["bt"]); cp.kappa_end_times = _dptr(keep["kt"])
        cp.base_slots = _dptr(keep["base"]); cp.data_initial_state = _dptr(keep["init"])
        cp.n_params = self.n_params
        cp.constraint_mode = self.constraint_mode if constraint_mode is None else constraint_mode
        cp.param_slot = keep["slot"].ctypes.data_as(C.POINTER(C.c_int32))
        cp.lower_bound = _dptr(keep["lo"]); cp.upper_bound = _dptr(keep["hi"])
        cp.abs_tol = self.abs_tol; cp.rel_tol = self.rel_tol; cp.dt_hint = self.dt_hint
        cp._keep = keep            # tie the arrays' lifetime to the struct
        return cp

    # ------------------------------------------------------------------------------------------
    def to_json(self) -> dict:
        def lst(a):
            return [float(x) for x in np.asarray(a).reshape(-1)]
        return dict(
            format="sepaihrd_problem/1", n_ages=self.n_ages, times=lst(self.times),
            obs_hosp=lst(self.obs_hosp), obs_icu=lst(self.obs_icu), obs_deaths=lst(self.obs_deaths),
            population=lst(self.population), contact_matrix_rowmajor=lst(self.contact_matrix),
            beta_end_times=lst(self.beta_end_times), kappa_end_times=lst(self.kappa_end_times),
            slot_names=self.layout.names(), base_slots=lst(self.base_slots),
            data_initial_state=lst(self.data_initial_state), param_names=list(self.param_names),
            lower_bound=lst(self.lower_bound), upper_bound=lst(self.upper_bound), sigmas=lst(self.sigmas),
            constraint_mode=self.constraint_mode, abs_tol=self.abs_tol, rel_tol=self.rel_tol,
            dt_hint=self.dt_hint, meta=self.meta)

    def save(self, path: str) -> None:
        with open(path, "w") as f:
            json.dump(self.to_json(), f, indent=0, separators=(",", ":"))
            f.write("\n")

    @staticmethod
    def from_json(d: dict) -> "Problem":
        n = int(d["n_ages"])
        n_obs = len(d["obs_hosp"]) // n
        base = np.array([float("nan") if x is None else x for x in d["base_slots"]], dtype=np.float64)
        return Problem(
            n_ages=n, times=d["times"],
            obs_hosp=np.array(d["obs_hosp"], dtype=np.float64).reshape(n_obs, n),
            obs_icu=np.array(d["obs_icu"], dtype=np.float64).reshape(n_obs, n),
            obs_deaths=np.array(d["obs_deaths"], dtype=np.float64).reshape(n_obs, n),
            population=d["population"],
            contact_matrix=np.array(d["contact_matrix_rowmajor"], dtype=np.float64).reshape(n, n),
            beta_end_times=d["beta_end_times"], kappa_end_times=d["kappa_end_times"],
            base_slots=base, data_initial_state=d["data_initial_state"],
            param_names=list(d["param_names"]), lower_bound=d["lower_bound"], upper_bound=d["upper_bound"],
            sigmas=d["sigmas"], constraint_mode=int(d.get("constraint_mode", CLAMP)),
            abs_tol=float(d.get("abs_tol", 1e-6)), rel_tol=float(d.get("rel_tol", 1e-6)),
            dt_hint=float(d.get("dt_hint", 1.0)), meta=dict(d.get("meta", {})))

    @staticmethod
    def load(path: str) -> "Problem":
        with open(path) as f:
            return Problem.from_json(json.load(f))

    # ------------------------------------------------------------------------------------------
    def expand_ages(self, factor: int) -> "Problem":
        """Synthetic many-age-group variant (BASELINE.json configs[4]):
        every age class is split into ``factor`` equal sub-classes:
        M'[i][j] = M[i/f][j/f]/f, N' = N/f, per-age parameters replicated, observations / f."""
        f = int(factor)
        n, n2 = self.n_ages, self.n_ages * f
        if n2 > MAX_AGES:
            raise ValueError("too many age classes")
        lay, lay2 = self.layout, SlotLayout(n2, len(self.beta_end_times), len(self.kappa_end_times))
        base2 = np.empty(lay2.count)
        base2[:lay.age0] = self.base_slots[:lay.age0]
        for b in range(len(_AGE_BLOCKS)):
            base2[lay2.age0 + b * n2: lay2.age0 + (b + 1) * n2] = np.repeat(self.base_slots[lay.age0 + b * n: lay.age0 + (b + 1) * n], f)
        base2[lay2.mult0:] = self.base_slots[lay.mult0:]
        names2: List[str] = []
        lo2: List[float] = []; hi2: List[float] = []; sg2: List[float] = []
        for nm, lo, hi, sg in zip(self.param_names, self.lower_bound, self.upper_bound, self.sigmas):
            blk = next((b for b in sorted(_AGE_BLOCKS, key=len, reverse=True) if nm.startswith(b + "_") and nm[len(b) + 1:].isdigit()), None)
            if blk is None:
                names2.append(nm); lo2.append(lo); hi2.append(hi); sg2.append(sg)
            else:
                i = int(nm[len(blk) + 1:])
                for s in range(f):
                    names2.append(f"{blk}_{i * f + s}"); lo2.append(lo); hi2.append(hi); sg2.append(sg)
        init2 = np.repeat(self.data_initial_state.reshape(NUM_COMPARTMENTS, n), f, axis=1) / f
        return Problem(
            n_ages=n2, times=self.times.copy(),
            obs_hosp=np.repeat(self.obs_hosp, f, axis=1) / f, obs_icu=np.repeat(self.obs_icu, f, axis=1) / f,
            obs_deaths=np.repeat(self.obs_deaths, f, axis=1) / f,
            population=np.repeat(self.population, f) / f,
            contact_matrix=np.repeat(np.repeat(self.contact_matrix, f, axis=0), f, axis=1) / f,
            beta_end_times=self.beta_end_times.copy(), kappa_end_times=self.kappa_end_times.copy(),
            base_slots=base2, data_initial_state=init2.reshape(-1), param_names=names2,
            lower_bound=lo2, upper_bound=hi2, sigmas=sg2, constraint_mode=self.constraint_mode,
            abs_tol=self.abs_tol, rel_tol=self.rel_tol, dt_hint=self.dt_hint,
            meta=dict(self.meta, expanded_from=n, factor=f))


    def select_ages(self, keep) -> "Problem":
        """The sub-problem of the age classes ``keep`` (indices, in the order given): populations, contact sub-matrix, per-age
        parameters, observations and initial state of those classes only.  Gives problems with any class count (3, 5, 7, ...)
        for the tests of the zero-padded kernel path (sepaihrd_create)."""
        keep = [int(i) for i in keep]
        n, n2 = self.n_ages, len(keep)
        if n2 < 1 or any(i < 0 or i >= n for i in keep):
            raise ValueError("age indices out of range")
        lay, lay2 = self.layout, SlotLayout(n2, len(self.beta_end_times), len(self.kappa_end_times))
        base2 = np.empty(lay2.count)
        base2[:lay.age0] = self.base_slots[:lay.age0]
        for b in range(len(_AGE_BLOCKS)):
            base2[lay2.age0 + b * n2: lay2.age0 + (b + 1) * n2] = self.base_slots[lay.age0 + b * n: lay.age0 + (b + 1) * n][keep]
        base2[lay2.mult0:] = self.base_slots[lay.mult0:]
        names2: List[str] = []
        lo2: List[float] = []; hi2: List[float] = []; sg2: List[float] = []
        for nm, lo, hi, sg in zip(self.param_names, self.lower_bound, self.upper_bound, self.sigmas):
            blk = next((b for b in sorted(_AGE_BLOCKS, key=len, reverse=True) if nm.startswith(b + "_") and nm[len(b) + 1:].isdigit()), None)
            if blk is None:
                names2.append(nm); lo2.append(lo); hi2.append(hi); sg2.append(sg)
            elif int(nm[len(blk) + 1:]) in keep:
                names2.append(f"{blk}_{keep.index(int(nm[len(blk) + 1:]))}"); lo2.append(lo); hi2.append(hi); sg2.append(sg)
        init2 = self.data_initial_state.reshape(NUM_COMPARTMENTS, n)[:, keep]
        return Problem(
            n_ages=n2, times=self.times.copy(),
            obs_hosp=self.obs_hosp[:, keep].copy(), obs_icu=self.obs_icu[:, keep].copy(), obs_deaths=self.obs_deaths[:, keep].copy(),
            population=self.population[keep].copy(), contact_matrix=self.contact_matrix[np.ix_(keep, keep)].copy(),
            beta_end_times=self.beta_end_times.copy(), kappa_end_times=self.kappa_end_times.copy(),
            base_slots=base2, data_initial_state=init2.reshape(-1).copy(), param_names=names2,
            lower_bound=lo2, upper_bound=hi2, sigmas=sg2, constraint_mode=self.constraint_mode,
            abs_tol=self.abs_tol, rel_tol=self.rel_tol, dt_hint=self.dt_hint,
            meta=dict(self.meta, selected_from=n, ages=keep))


def default_problem_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "spain2020_problem.json")


def load_default_problem() -> Problem:
    """The Spain-2020, 4-age-class problem of BASELINE.json configs[0..3]."""
    return Problem.load(default_problem_path())
