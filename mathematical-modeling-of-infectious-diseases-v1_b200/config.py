"""Readers for the reference's on-disk formats, so the evaluator can
be driven from a reference-style project tree without the reference code:

* ``key value...`` text files with ``#`` comments -- ``readSEPAIHRDParameters``, ``readParamBounds``,
  ``readProposalSigmas``, ``readParamsToCalibrate``, settings files
  (reference src/utils/ReadCalibrationConfiguration.cpp:164-420)
* the contact matrix CSV -- ``readMatrixFromCSV`` (src/utils/ReadContactMatrix.cpp:8-82)
* the processed daily data CSV -- ``CalibrationData::readCSVData`` (src/utils/GetCalibrationData.cpp:236-401)

and the assembly of a :class:`Problem` from them, following ``main`` (src/model/main.cpp:188-316).
Error behaviour mirrors the reference: malformed input raises ``ValueError`` (DataFormatException /
CSVReadException there), missing files raise ``FileNotFoundError`` (FileIOException there).
"""
from __future__ import annotations

import csv
import os
from typing import Dict, List, Tuple

import numpy as np

from .problem import (NUM_COMPARTMENTS, Problem, SlotLayout, _AGE_BLOCKS, _MULTIPLIERS, _SCALARS)


def _lines(path: str):
    with open(path) as f:
        for raw in f:
            line = raw.strip(" \t\n\r\f\v")
            if not line or line.startswith("#"):
                continue
            yield line


def read_sepaihrd_parameters(path: str, n_ages: int) -> Dict[str, object]:
    """readSEPAIHRDParameters (ReadCalibrationConfiguration.cpp:164-271).

    Returns a dict with scalar floats, per-age numpy vectors, ``beta_values`` / ``kappa_values``
    lists assembled from ``beta_k`` / ``kappa_k`` (1-based), and ``beta_end_times`` /
    ``kappa_end_times``.  Unknown keys are ignored (the reference logs a warning)."""
    out: Dict[str, object] = {blk: np.zeros(n_ages) for blk in _AGE_BLOCKS}
    beta_map: Dict[int, float] = {}
    kappa_map: Dict[int, float] = {}
    scalars = set(_SCALARS) | set(_MULTIPLIERS) | {"beta", "runup_days", "seed_exposed"}
    for line in _lines(path):
        toks = line.split()
        name, vals = toks[0], []
        for t in toks[1:]:
            try:
                vals.append(float(t))
            except ValueError:
                break                      # `iss >> value` stops at the first non-number
        if not vals:
            continue
        if name.startswith("beta_") and name != "beta_end_times":
            try:
                beta_map[int(name[5:])] = vals[0]
            except ValueError:
                pass
        elif name.startswith("kappa_") and name != "kappa_end_times":
            try:
                kappa_map[int(name[6:])] = vals[0]
            except ValueError:
                pass
        elif name in scalars:
            out[name] = vals[0]
        elif name in ("beta_end_times", "kappa_end_times"):
            out[name] = list(vals)
        elif name in _AGE_BLOCKS:
            if len(vals) != n_ages:
                raise ValueError(f"Incorrect number of values for {name}. Expected {n_ages}, got {len(vals)}")
            out[name] = np.array(vals)
    for key, mp in (("beta_values", beta_map), ("kappa_values", kappa_map)):
        if mp:
            arr = [0.0] * max(mp)
            for k, v in mp.items():
                arr[k - 1] = v
            out[key] = arr
        else:
            out[key] = []
    out.setdefault("beta_end_times", [])
    out.setdefault("kappa_end_times", [])
    return out


def read_param_bounds(path: str) -> Dict[str, Tuple[float, float]]:
    """readParamBounds (.cpp:273-305): ``name low high``; anything else is an error."""
    out = {}
    for line in _lines(path):
        toks = line.split()
        if len(toks) != 3:
            raise ValueError(f"Invalid line in bounds file: {line}")
        out[toks[0]] = (float(toks[1]), float(toks[2]))
    return out


def read_proposal_sigmas(path: str) -> Dict[str, float]:
    """readProposalSigmas (.cpp:308-339): ``name sigma``."""
    out = {}
    for line in _lines(path):
        toks = line.split()
        if len(toks) != 2:
            raise ValueError(f"Invalid line in proposal sigmas file: {line}")
        out[toks[0]] = float(toks[1])
    return out


def read_params_to_calibrate(path: str) -> List[str]:
    """readParamsToCalibrate (.cpp:342-370): first word of every non-comment line, in file order."""
    return [line.split()[0] for line in _lines(path)]


def read_settings(path: str) -> Dict[str, float]:
    """readSettingsFile (.cpp:373-405): ``name value``."""
    out = {}
    for line in _lines(path):
        toks = line.split()
        if len(toks) != 2:
            raise ValueError(f"Invalid line in settings file: {line}")
        out[toks[0]] = float(toks[1])
    return out


def read_matrix_csv(path: str, rows: int, cols: int) -> np.ndarray:
    """readMatrixFromCSV (ReadContactMatrix.cpp:8-82): leading ``//`` comment lines are skipped,
    then ``rows`` non-empty lines with at least ``cols`` comma-separated numbers."""
    mat = np.zeros((rows, cols))
    with open(path) as f:
        lines = [ln.rstrip("\r\n") for ln in f]
    i = 0
    while i < len(lines) and lines[i] and lines[i].startswith("//"):
        i += 1
    data = [ln for ln in lines[i:] if ln]
    if len(data) < rows:
        raise ValueError(f"expected {rows} rows, found {len(data)} in {path}")
    for r in range(rows):
        cells = data[r].split(",")
        if len(cells) < cols:
            raise ValueError(f"not enough columns in row {r + 1} of {path}")
        for c in range(cols):
            mat[r, c] = float(cells[c])
    return mat


_AGE_SUFFIXES = ["0_30", "30_60", "60_80", "80_plus"]


class CalibrationData:
    """CalibrationData(filename, start_date, end_date) (GetCalibrationData.cpp:15-22, 236-401):
    4 fixed age bands, rows kept when start <= date <= end (string comparison), population from the
    first kept row."""

    def __init__(self, path: str, start_date: str = "", end_date: str = ""):
        with open(path, newline="") as f:
            rd = csv.reader(f)
            header = next(rd)
            col = {name: i for i, name in enumerate(header)}

            def cols(prefix):
                try:
                    return [col[f"{prefix}_{s}"] for s in _AGE_SUFFIXES]
                except KeyError as e:
                    raise ValueError(f"Missing required column: {e.args[0]}")
            idx = dict(conf=cols("new_confirmed"), dec=cols("new_deceased"), hosp=cols("new_hospitalized_patients"),
                       icu=cols("new_intensive_care_patients"), pop=cols("population"),
                       cconf=cols("cumulative_confirmed"), cdec=cols("cumulative_deceased"),
                       chosp=cols("cumulative_hospitalized_patients"), cicu=cols("cumulative_intensive_care_patients"))
            rows = []
            for row in rd:
                if not row:
                    continue
                d = row[col["date"]]
                if start_date and d < start_date:
                    continue
                if end_date and d > end_date:
                    continue
                rows.append(row)
        if not rows:
            raise ValueError("No data points found in specified date range.")

        def mat(ix):
            return np.array([[float(r[i]) for i in ix] for r in rows])
        self.dates = [r[col["date"]] for r in rows]
        self.new_confirmed = mat(idx["conf"]); self.new_deaths = mat(idx["dec"])
        self.new_hospitalizations = mat(idx["hosp"]); self.new_icu = mat(idx["icu"])
        self.cumulative_confirmed = mat(idx["cconf"]); self.cumulative_deaths = mat(idx["cdec"])
        self.cumulative_hospitalizations = mat(idx["chosp"]); self.cumulative_icu = mat(idx["cicu"])
        self.population = np.array([float(rows[0][i]) for i in idx["pop"]])
        self.n_data_points = len(rows)


def initial_state_from_data(data: CalibrationData, sigma, gamma_p, gamma_a, gamma_i, p_asym) -> np.ndarray:
    """CalibrationData::getInitialSEPAIHRDState (GetCalibrationData.cpp:107-234) in numpy."""
    N = data.population
    n = len(N)
    D0 = np.maximum(data.cumulative_deaths[0], 0.0)
    H0 = np.maximum(data.cumulative_hospitalizations[0], 0.0)
    ICU0 = np.maximum(data.cumulative_icu[0], 0.0)
    CumH0, CumICU0 = H0.copy(), ICU0.copy()
    I0 = np.maximum(data.cumulative_confirmed[0] - D0, 0.0)
    E0 = np.zeros(n); P0 = np.zeros(n); A0 = np.zeros(n); R0 = np.zeros(n)
    for i in range(n):
        p_i = min(max(p_asym[i], 0.0), 1.0)
        omp = 1.0 - p_i
        P0[i] = I0[i] * gamma_i / (omp * gamma_p) if (gamma_p > 1e-9 and omp > 1e-9) else I0[i]
        A0[i] = P0[i] * p_i * gamma_p / gamma_a if gamma_a > 1e-9 else P0[i] * p_i
        E0[i] = P0[i] * gamma_p / sigma if sigma > 1e-9 else P0[i]
    E0 = np.maximum(E0, 0.0); P0 = np.maximum(P0, 0.0); A0 = np.maximum(A0, 0.0)
    for i in range(n):
        D0[i] = min(D0[i], N[i])
        ICU0[i] = min(ICU0[i], max(0.0, N[i] - D0[i]))
        H0[i] = min(H0[i], max(0.0, N[i] - D0[i] - ICU0[i]))
        I0[i] = min(I0[i], max(0.0, N[i] - D0[i] - ICU0[i] - H0[i]))
        R0[i] = min(R0[i], max(0.0, N[i] - D0[i] - ICU0[i] - H0[i] - I0[i]))
    for i in range(n):
        s_set = I0[i] + H0[i] + ICU0[i] + R0[i] + D0[i]
        s_inf = E0[i] + P0[i] + A0[i]
        avail = max(N[i] - s_set, 0.0)
        if s_inf > avail:
            sc = avail / s_inf if s_inf > 1e-9 else 0.0
            E0[i] *= sc; P0[i] *= sc; A0[i] *= sc
    st = np.zeros(NUM_COMPARTMENTS * n)
    for c, v in ((1, E0), (2, P0), (3, A0), (4, I0), (5, H0), (6, ICU0), (7, R0), (8, D0), (9, CumH0), (10, CumICU0)):
        st[c * n:(c + 1) * n] = v
    for i in range(n):
        s = 0.0
        for j in range(1, 9):
            s += st[j * n + i]
        st[i] = max(0.0, N[i] - s)
    return st


def problem_from_reference_tree(root: str, start_date: str = "2020-03-01", end_date: str = "2020-12-31",
                                n_ages: int = 4) -> Problem:
    """Assemble the problem exactly as ``main`` does (src/model/main.cpp:188-316, 377-380):
    data window, contact matrix, initial_guess.txt, bounds, sigmas, params_to_calibrate,
    time grid ``-int(runup_days) .. num_days-1``, tolerances 1e-6 / 1e-6."""
    cfg = os.path.join(root, "data", "configuration")
    data = CalibrationData(os.path.join(root, "data", "processed", "processed_data.csv"), start_date, end_date)
    M = read_matrix_csv(os.path.join(root, "data", "contacts.csv"), n_ages, n_ages)
    prm = read_sepaihrd_parameters(os.path.join(cfg, "initial_guess.txt"), n_ages)
    bounds = read_param_bounds(os.path.join(cfg, "param_bounds.txt"))
    sigmas = read_proposal_sigmas(os.path.join(cfg, "proposal_sigmas.txt"))
    names = read_params_to_calibrate(os.path.join(cfg, "params_to_calibrate.txt"))
    if len(prm["kappa_values"]) != len(prm["kappa_end_times"]) or len(prm["beta_values"]) != len(prm["beta_end_times"]):
        raise ValueError("Mismatch between end times and values for kappa or beta schedules.")   # main.cpp:226-228

    lay = SlotLayout(n_ages, len(prm["beta_end_times"]), len(prm["kappa_end_times"]))
    base = np.zeros(lay.count)
    base[lay.beta0:lay.beta0 + lay.nb] = prm["beta_values"]
    base[lay.kappa0:lay.kappa0 + lay.nk] = prm["kappa_values"]
    for s in _SCALARS:
        base[lay.scalar(s)] = prm[s]
    for blk in _AGE_BLOCKS:
        base[lay.age(blk, 0):lay.age(blk, 0) + n_ages] = prm[blk]
    for m, nm in enumerate(_MULTIPLIERS):
        base[lay.mult0 + m] = prm[nm]
    base[lay.seed_exposed] = prm["seed_exposed"]
    base[lay.runup_days] = prm["runup_days"]
    base[lay.beta_scalar] = prm.get("beta", 0.0)   # quirk Q1: uninitialised in the reference, never read with a schedule

    runup = int(prm["runup_days"])                  # static_cast<int> truncation, main.cpp:247-253
    times = np.arange(-runup, data.n_data_points, dtype=np.float64)
    init = initial_state_from_data(data, prm["sigma"], prm["gamma_p"], prm["gamma_A"], prm["gamma_I"], prm["p"])
    for nm in names:
        if nm not in sigmas:
            raise ValueError(f"Missing proposal sigma for parameter: {nm}")   # ParameterManager.cpp:48-50
        if nm not in bounds:
            raise ValueError(f"Missing bounds for parameter: {nm}")           # .cpp:51-53
    return Problem(
        n_ages=n_ages, times=times, obs_hosp=data.new_hospitalizations, obs_icu=data.new_icu,
        obs_deaths=data.new_deaths, population=data.population, contact_matrix=M,
        beta_end_times=prm["beta_end_times"], kappa_end_times=prm["kappa_end_times"], base_slots=base,
        data_initial_state=init, param_names=names, lower_bound=[bounds[nm][0] for nm in names],
        upper_bound=[bounds[nm][1] for nm in names], sigmas=[sigmas[nm] for nm in names],
        meta=dict(source="adjo0043/Mathematical-Modeling-Of-Infectious-Diseases-V1 data/ tree",
                  window=[start_date, end_date], dates=[data.dates[0], data.dates[-1]]))


def write_reference_tree(problem: Problem, root: str, start_date: str = "2020-03-01") -> None:
    """Write ``problem`` as a reference-style project tree (data/processed/processed_data.csv, data/contacts.csv,
    data/configuration/{initial_guess,params_to_calibrate,param_bounds,proposal_sigmas}.txt) that the readers above --
    and the C++ ones of host/config_io.hpp, e.g. through host/sepaihrd_objective_benchmark -- load back into the same
    problem.  Numbers are written with 17 significant digits, so the round trip is exact for everything the hot path
    reads; the cumulative columns of the first day are chosen such that the data-derived initial state is reproduced
    (to rounding), later cumulative rows are running sums.  4 age classes (the CSV format has 4 fixed bands)."""
    if problem.n_ages != 4:
        raise ValueError("the reference's data format has 4 fixed age bands")
    n, lay = 4, problem.layout
    os.makedirs(os.path.join(root, "data", "processed"), exist_ok=True)
    os.makedirs(os.path.join(root, "data", "configuration"), exist_ok=True)
    r = lambda v: repr(float(v))
    st = problem.data_initial_state.reshape(NUM_COMPARTMENTS, n)
    cum = dict(cumulative_deceased=st[8].copy(), cumulative_hospitalized_patients=st[9].copy(),
               cumulative_intensive_care_patients=st[10].copy(), cumulative_confirmed=st[4] + st[8])
    new = dict(new_confirmed=np.zeros((problem.n_obs, n)), new_deceased=problem.obs_deaths,
               new_hospitalized_patients=problem.obs_hosp, new_intensive_care_patients=problem.obs_icu)
    series = list(new) + list(cum) + ["population"]
    header = ["date"] + [f"{s}_{b}" for s in series for b in _AGE_SUFFIXES]
    day0 = np.datetime64(start_date)
    with open(os.path.join(root, "data", "processed", "processed_data.csv"), "w") as f:
        f.write(",".join(header) + "\n")
        for d in range(problem.n_obs):
            if d > 0:
                for s in cum:
                    cum[s] = cum[s] + np.nan_to_num(np.maximum(new["new" + s[len("cumulative"):]][d], 0.0))
            cells = [str(day0 + d)]
            for s in series:
                row = problem.population if s == "population" else (new[s][d] if s in new else cum[s])
                cells += [r(v) if np.isfinite(v) else "-1.0" for v in row]
            f.write(",".join(cells) + "\n")
    with open(os.path.join(root, "data", "contacts.csv"), "w") as f:
        f.write("// contact matrix M(i, j), row i = contacted age class\n")
        for i in range(n):
            f.write(",".join(r(v) for v in problem.contact_matrix[i]) + "\n")
    s = problem.base_slots
    with open(os.path.join(root, "data", "configuration", "initial_guess.txt"), "w") as f:
        f.write("# written by sepaihrd_b200.config.write_reference_tree\n")
        f.write("beta_end_times " + " ".join(r(v) for v in problem.beta_end_times) + "\n")
        for k in range(lay.nb):
            f.write(f"beta_{k + 1} {r(s[lay.beta0 + k])}\n")
        if np.isfinite(s[lay.beta_scalar]):
            f.write(f"beta {r(s[lay.beta_scalar])}\n")
        for nm in _SCALARS:
            f.write(f"{nm} {r(s[lay.scalar(nm)])}\n")
        for blk in _AGE_BLOCKS:
            f.write(blk + " " + " ".join(r(v) for v in s[lay.age(blk, 0):lay.age(blk, 0) + n]) + "\n")
        for m, nm in enumerate(_MULTIPLIERS):
            f.write(f"{nm} {r(s[lay.mult0 + m])}\n")
        f.write(f"runup_days {r(s[lay.runup_days])}\nseed_exposed {r(s[lay.seed_exposed])}\n")
        f.write("kappa_end_times " + " ".join(r(v) for v in problem.kappa_end_times) + "\n")
        for k in range(lay.nk):
            f.write(f"kappa_{k + 1} {r(s[lay.kappa0 + k])}\n")
    cfg = os.path.join(root, "data", "configuration")
    with open(os.path.join(cfg, "params_to_calibrate.txt"), "w") as f:
        f.write("\n".join(problem.param_names) + "\n")
    with open(os.path.join(cfg, "param_bounds.txt"), "w") as f:
        for nm, lo, hi in zip(problem.param_names, problem.lower_bound, problem.upper_bound):
            f.write(f"{nm} {r(lo)} {r(hi)}\n")
    with open(os.path.join(cfg, "proposal_sigmas.txt"), "w") as f:
        for nm, sg in zip(problem.param_names, problem.sigmas):
            f.write(f"{nm} {r(sg)}\n")
