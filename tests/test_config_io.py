"""CPU tests of the C++ readers / writer for the reference's on-disk formats (host/config_io.{hpp,cpp}, SURVEY.md
section 8f row 4), through the host C API.

Three anchors:
  * the Python readers of config.py (an independent restatement of the same reference functions) on generated files:
    every parsed number must be bit-identical;
  * the cases of the reference's own reader tests that still match its current sources
    (tests/utils/FileUtilsTests.cpp:146-176, 192-217, 304-319; tests/utils/ReadContactMatrixTests.cpp:57-140) --
    the reference suite also holds cases written for an older reader (messages such as "Error reading scalar value")
    that its current src/utils/ReadCalibrationConfiguration.cpp no longer produces; those follow the current source;
  * the committed Spain-2020 problem (data/spain2020_problem.json, extracted from the reference tree by
    tools/extract_reference_problem.py): when the reference tree is present, the C++ assembly of main()'s setup must
    reproduce it exactly.
"""
import json
import os

import numpy as np
import pytest

BANDS = ["0_30", "30_60", "60_80", "80_plus"]
SERIES = ["new_confirmed", "new_deceased", "new_hospitalized_patients", "new_intensive_care_patients",
          "cumulative_confirmed", "cumulative_deceased", "cumulative_hospitalized_patients",
          "cumulative_intensive_care_patients", "population"]


@pytest.fixture(scope="module")
def host(pkg):
    from sepaihrd_b200 import hostlib
    hostlib.load_library()
    return hostlib


@pytest.fixture(scope="module")
def cfg(pkg):
    from sepaihrd_b200 import config
    return config


def _write(path, text):
    with open(path, "w", newline="") as f:
        f.write(text)
    return str(path)


def _make_tree(root, n_days=40, seed=3, runup="2.05541965e+01", seed_exposed="1.5e+01"):
    """A small reference-style project tree with awkward but legal formatting."""
    rng = np.random.default_rng(seed)
    os.makedirs(root / "data" / "processed")
    os.makedirs(root / "data" / "configuration")
    cols = ["date", "location_key"] + [f"{s}_{b}" for s in SERIES for b in BANDS] + ["trailing_note"]
    middle = cols[2:-1]
    rng.shuffle(middle)                 # the reader finds columns by name, wherever they are
    cols = cols[:2] + middle + cols[-1:]
    pop = [14075720.0, 20948387.0, 9032069.0, 2880884.0]
    lines = [",".join(cols)]
    cum = {s: np.zeros(4) for s in SERIES if s.startswith("cumulative")}
    for d in range(n_days + 10):
        date = str(np.datetime64("2020-02-25") + d)
        new = {s: rng.poisson(30.0 * (1 + d), 4).astype(float) for s in SERIES if s.startswith("new")}
        for s in new:
            cum["cumulative" + s[3:]] += new[s]
        row = {"date": date, "location_key": "ES", "trailing_note": ""}
        for s in SERIES:
            for k, b in enumerate(BANDS):
                v = pop[k] if s == "population" else (new[s][k] if s in new else cum[s][k])
                row[f"{s}_{b}"] = repr(float(v)) if d % 3 else f"{v:.1f}"
        lines.append(",".join(row[c] for c in cols))
        if d == 5:
            lines.append("")            # empty lines are skipped
    _write(root / "data" / "processed" / "processed_data.csv", "\n".join(lines) + "\n")
    M = rng.uniform(0.1, 9.0, (4, 4))
    _write(root / "data" / "contacts.csv", "// contact matrix\n// second comment\n" +
           "\n".join(",".join(repr(float(x)) for x in r) for r in M) + "\n")
    _write(root / "data" / "configuration" / "initial_guess.txt", f"""# initial guess
beta_end_times 13.0 63.0 84.0
   beta_1   2.5e-01   # [C]
beta_2 1.25e-01
beta_3 0.2 # [C]
theta 3.0e-01
sigma\t2.6e-01
gamma_p 4.9e-01
gamma_A 1.9e-01 # [C]
gamma_I 1.7e-01
gamma_H 1.5e-01
gamma_ICU 1.1e-01
p 4.5e-01 3.5e-01 2.5e-01 1.5e-01 # [C]
a 5.0e-01 1.0e+00 1.0e+00 7.0e-01
h_infec 1 1 1 1
h 1.0e-03 1.0e-02 5.0e-02 1.0e-01
icu 1.0e-02 5.0e-02 1.5e-01 5.0e-02
d_H 1.0e-03 5.0e-03 5.0e-02 2.0e-01
d_ICU 5.0e-02 1.0e-01 3.0e-01 6.0e-01
d_community 0 0 1.0e-03 2.0e-02
E0_multiplier 1.5
P0_multiplier 1.25
A0_multiplier 1.0
I0_multiplier 0.75
H0_multiplier 1.0
ICU0_multiplier 1.0
R0_multiplier 1.0
D0_multiplier 1.0
runup_days {runup}
seed_exposed {seed_exposed}
unknown_key 7
kappa_end_times 13.0 40.0 84.0
kappa_1 1.0
kappa_2 3.5e-01
kappa_3 6.0e-01
""")
    names = ["beta_1", "beta_3", "kappa_2", "kappa_3", "gamma_A", "p_0", "p_3", "h_1", "d_community_3", "seed_exposed",
             "runup_days", "E0_multiplier", "theta"]
    _write(root / "data" / "configuration" / "params_to_calibrate.txt",
           "# names\n" + "\n".join(nm + (" extra words" if i == 2 else "") for i, nm in enumerate(names)) + "\n\n")
    _write(root / "data" / "configuration" / "param_bounds.txt",
           "# bounds\n" + "\n".join(f"{nm}   {0.01 * (i + 1):.3e}\t{2.0 + i}" for i, nm in enumerate(names)) + "\nnot_calibrated 0 1\n")
    _write(root / "data" / "configuration" / "proposal_sigmas.txt",
           "\n".join(f"{nm} {0.001 * (i + 1)!r}" for i, nm in enumerate(names)) + "\n")
    return names


def _lst(a):
    return [float(x) for x in np.asarray(a, dtype=np.float64).reshape(-1)]


# ---------------------------------------------------------------------------------------------------------------------
def test_project_assembly_matches_the_python_readers(tmp_path, host, cfg):
    _make_tree(tmp_path)
    got = host.load_reference_project(str(tmp_path), "2020-03-01", "2020-03-31")
    want = cfg.problem_from_reference_tree(str(tmp_path), "2020-03-01", "2020-03-31").to_json()
    assert got["param_names"] == want["param_names"]
    assert got["n_ages"] == 4
    for key in ("times", "obs_hosp", "obs_icu", "obs_deaths", "population", "contact_matrix_rowmajor", "beta_end_times",
                "kappa_end_times", "data_initial_state", "lower_bound", "upper_bound", "sigmas"):
        assert got[key] == want[key], key                      # bit-identical doubles
    # base slots: the C++ SEPAIHRDParameters leaves `beta` NaN when the file has no such line (quirk Q1; the reference
    # leaves it uninitialised), the Python reader stores 0.0; every other slot is identical
    g = [float("nan") if x is None else x for x in got["base_slots"]]
    assert len(g) == len(want["base_slots"])
    assert np.isnan(g[-1]) and g[:-1] == want["base_slots"][:-1]
    assert len(got["times"]) == 20 + 31 and got["times"][0] == -20.0          # int(20.55) days of run-up (quirk Q3)


def test_initial_state_of_main(tmp_path, host, cfg):
    """main.cpp:268-316: run-up seeding, else multipliers; S is the remainder."""
    _make_tree(tmp_path)
    prj = host.load_reference_project(str(tmp_path), "2020-03-01", "2020-03-31")
    x = np.array(prj["initial_state"]).reshape(11, 4)
    N = np.array(prj["population"])
    np.testing.assert_array_equal(x[1], 15.0 * (N / N.sum()))
    assert not x[2:].any()
    np.testing.assert_array_equal(x[0], N - x[1])

    other = tmp_path / "mult"
    _make_tree(other, seed_exposed="0.0")
    prj = host.load_reference_project(str(other), "2020-03-01", "2020-03-31")
    x = np.array(prj["initial_state"]).reshape(11, 4)
    d = np.array(prj["data_initial_state"]).reshape(11, 4)
    mult = [1.5, 1.25, 1.0, 0.75, 1.0, 1.0, 1.0, 1.0]
    for c in range(1, 9):
        np.testing.assert_array_equal(x[c], d[c] * mult[c - 1])
    np.testing.assert_array_equal(x[9:], d[9:])
    np.testing.assert_allclose(x[:9].sum(axis=0), N, rtol=1e-15)


def test_data_reader_matches_python(tmp_path, host, cfg):
    _make_tree(tmp_path)
    path = str(tmp_path / "data" / "processed" / "processed_data.csv")
    for window in (("2020-03-01", "2020-03-31"), ("", "2020-03-03"), ("2020-04-01", ""), ("", "")):
        got = host.read_file("data", path, start_date=window[0], end_date=window[1])
        ref = cfg.CalibrationData(path, *window)
        assert got["dates"] == ref.dates
        assert got["population"] == _lst(ref.population)
        for key, arr in (("new_confirmed", ref.new_confirmed), ("new_hospitalizations", ref.new_hospitalizations),
                         ("new_icu", ref.new_icu), ("new_deaths", ref.new_deaths),
                         ("cumulative_confirmed", ref.cumulative_confirmed), ("cumulative_deaths", ref.cumulative_deaths),
                         ("cumulative_hospitalizations", ref.cumulative_hospitalizations), ("cumulative_icu", ref.cumulative_icu)):
            assert got[key] == _lst(arr), key


def test_data_reader_errors(tmp_path, host):
    _make_tree(tmp_path)
    path = str(tmp_path / "data" / "processed" / "processed_data.csv")
    with pytest.raises(host.HostError, match="no data points"):
        host.read_file("data", path, start_date="2021-01-01", end_date="2021-02-01")
    with pytest.raises(host.HostError, match="unable to open"):
        host.read_file("data", str(tmp_path / "nope.csv"))
    text = open(path).read()
    with pytest.raises(host.HostError, match="Missing required column: new_deceased_60_80"):
        host.read_file("data", _write(tmp_path / "nocol.csv", text.replace("new_deceased_60_80", "renamed", 1)))
    lines = text.split("\n")
    lines[3] = lines[3].replace(lines[3].split(",")[5], "abc", 1)
    with pytest.raises(host.HostError, match="Failed to parse value"):
        host.read_file("data", _write(tmp_path / "badnum.csv", "\n".join(lines)))
    lines = text.split("\n")
    lines[4] = ",".join(lines[4].split(",")[:6])
    with pytest.raises(host.HostError, match="insufficient columns"):
        host.read_file("data", _write(tmp_path / "short.csv", "\n".join(lines)))


def test_parameter_file_reference_cases(tmp_path, host):
    """tests/utils/FileUtilsTests.cpp:146-176 (valid file, 2 age classes), :192-217 (age-vector length), :304-319
    (whitespace; the later duplicate wins)."""
    valid = _write(tmp_path / "params.txt", "# This is a comment\nbeta 0.5\ntheta 0.1\nsigma 0.2\ngamma_p 0.3\ngamma_A 0.4\n"
                   "gamma_I 0.5\ngamma_H 0.6\ngamma_ICU 0.7\ncontact_matrix_scaling_factor 1.0\np 0.1 0.2 # Age-specific\n"
                   "h 0.3 0.4\nicu 0.05 0.1\nd_H 0.01 0.02\nd_ICU 0.03 0.04\n")
    p = host.read_file("parameters", valid, 2)
    assert (p["beta"], p["sigma"], p["gamma_ICU"]) == (0.5, 0.2, 0.7)
    assert p["p"] == [0.1, 0.2] and p["d_ICU"] == [0.03, 0.04] and p["h"] == [0.3, 0.4]
    assert p["a"] == [0.0, 0.0] and p["d_community"] == [0.0, 0.0]           # absent vectors are pre-sized zeros
    assert p["beta_values"] == [] and p["kappa_values"] == []
    with pytest.raises(host.HostError, match="Unable to open parameters file"):
        host.read_file("parameters", str(tmp_path / "does_not_exist.txt"), 2)
    with pytest.raises(host.HostError, match="Incorrect number of values for p. Expected 2, got 1"):
        host.read_file("parameters", _write(tmp_path / "missing.txt", "p 0.1 # Only one value for 2 age classes\n"), 2)
    with pytest.raises(host.HostError, match="Incorrect number of values for p. Expected 2, got 3"):
        host.read_file("parameters", _write(tmp_path / "extra.txt", "p 0.1 0.2 0.3 # Three values for 2 age classes\n"), 2)
    ws = _write(tmp_path / "ws.txt", "   beta    0.5   # Comment with spaces   \n\tbeta\t0.6\t#\tComment with tabs\t\np\t0.1  0.2   # Mixed whitespace\n")
    p = host.read_file("parameters", ws, 2)
    assert p["beta"] == 0.6 and p["p"] == [0.1, 0.2]


def test_parameter_file_current_source_behaviour(tmp_path, host, cfg):
    """What src/utils/ReadCalibrationConfiguration.cpp:164-271 does today with odd lines: non-numeric values and unknown
    names are skipped with a warning, indexed schedule values may arrive in any order and leave gaps as zeros."""
    f = _write(tmp_path / "odd.txt", "beta not_a_number\nsigma\nkappa_3 0.9\nkappa_1 1.0\nbeta_2 0.25\nbeta_x 3\n"
               "beta_end_times 10 20\nkappa_end_times 5 6 7\ngamma_p 0.4 0.5\nwhatever 1 2 3\ntheta 1e-1junk\nh_infec 1 2\n")
    p = host.read_file("parameters", f, 2)
    assert p["beta"] is None                       # never set: NaN (the reference leaves it uninitialised, quirk Q1)
    assert p["sigma"] == 0.0
    assert p["kappa_values"] == [1.0, 0.0, 0.9] and p["beta_values"] == [0.0, 0.25]
    assert p["beta_end_times"] == [10.0, 20.0] and p["kappa_end_times"] == [5.0, 6.0, 7.0]
    assert p["gamma_p"] == 0.0                     # two numbers on a scalar line: the reference's scalar stays 0.0
    assert p["theta"] == 0.1                       # formatted extraction stops inside the token
    assert p["h_infec"] == [1.0, 2.0]
    ref = cfg.read_sepaihrd_parameters(f, 2)
    assert p["kappa_values"] == ref["kappa_values"] and p["beta_values"] == ref["beta_values"]


def test_bounds_sigmas_names_settings(tmp_path, host, cfg):
    names = _make_tree(tmp_path)
    c = tmp_path / "data" / "configuration"
    b = host.read_file("bounds", str(c / "param_bounds.txt"))
    assert {k: tuple(v) for k, v in b.items()} == cfg.read_param_bounds(str(c / "param_bounds.txt"))
    assert host.read_file("sigmas", str(c / "proposal_sigmas.txt")) == cfg.read_proposal_sigmas(str(c / "proposal_sigmas.txt"))
    assert host.read_file("names", str(c / "params_to_calibrate.txt"))["names"] == names
    s = _write(tmp_path / "mcmc.txt", "# settings\nmcmc_iterations 1000\nburn_in\t200\n  thinning 2  \nstep_size 5e-2\n")
    assert host.read_file("settings", s) == cfg.read_settings(s) == dict(mcmc_iterations=1000.0, burn_in=200.0, thinning=2.0, step_size=0.05)
    for kind, bad, msg in (("bounds", "x 1\n", "Invalid line in bounds file"), ("bounds", "x 1 2 3\n", "Too many values on line in bounds file"),
                           ("bounds", "x 1 2 y\n", "Too many values"), ("bounds", "x one 2\n", "Invalid line"),
                           ("sigmas", "x\n", "Invalid line in proposal sigmas file"), ("sigmas", "x 1 2\n", "Too many values on line in sigmas file"),
                           ("settings", "x\n", "Invalid line in settings file"), ("settings", "x 1 # note\n", "Too many values on line in settings file")):
        with pytest.raises(host.HostError, match=msg):
            host.read_file(kind, _write(tmp_path / "bad.txt", bad))
    for kind, msg in (("bounds", "Error opening param bounds file"), ("sigmas", "Error opening proposal sigmas file"),
                      ("names", "Error opening params_to_calibrate file"), ("settings", "Error opening settings file")):
        with pytest.raises(host.HostError, match=msg):
            host.read_file(kind, str(tmp_path / "absent.txt"))


def test_contact_matrix_reference_cases(tmp_path, host, cfg):
    """tests/utils/ReadContactMatrixTests.cpp:57-140."""
    ok = _write(tmp_path / "m.csv", "1.0,2.5,3.0\n4.0,5.0,6.5\n")
    assert host.read_file("matrix", ok, 2, 3)["rowmajor"] == [1.0, 2.5, 3.0, 4.0, 5.0, 6.5]
    np.testing.assert_array_equal(cfg.read_matrix_csv(ok, 2, 3).reshape(-1), [1.0, 2.5, 3.0, 4.0, 5.0, 6.5])
    commented = _write(tmp_path / "c.csv", "// header\n//more\n1,2\n\n3,4,99\n")
    assert host.read_file("matrix", commented, 2, 2)["rowmajor"] == [1.0, 2.0, 3.0, 4.0]
    with pytest.raises(host.HostError, match="Could not open file"):
        host.read_file("matrix", str(tmp_path / "non_existent_file.csv"), 2, 2)
    with pytest.raises(host.HostError, match="Invalid number format at row 2, column 1: 'abc'"):
        host.read_file("matrix", _write(tmp_path / "inv.csv", "1.0,2.0\nabc,4.0\n"), 2, 2)
    with pytest.raises(host.HostError, match="Not enough rows: expected 3 rows, found 2"):
        host.read_file("matrix", _write(tmp_path / "rows.csv", "1.0,2.0\n3.0,4.0\n"), 3, 2)
    with pytest.raises(host.HostError, match="Not enough columns in row 2"):
        host.read_file("matrix", _write(tmp_path / "cols.csv", "1.0,2.0,3.0\n4.0,5.0\n"), 2, 3)
    with pytest.raises(host.HostError, match="Not enough rows"):
        host.read_file("matrix", _write(tmp_path / "empty.csv", ""), 2, 2)


def test_save_calibration_results_round_trip(tmp_path, host):
    names = _make_tree(tmp_path)
    src = str(tmp_path / "data" / "configuration" / "initial_guess.txt")
    out = str(tmp_path / "calibrated.txt")
    host.resave_parameters(src, 4, out, names, -1.2068696767e6, "2020-12-31 23:59:59")
    text = open(out).read().split("\n")
    assert text[0] == "# Calibrated SEPAIHRD Model Parameters"
    assert text[1] == "# Calibration completed: 2020-12-31 23:59:59"
    assert text[2] == "# Best objective function value: -1.20686968e+06"
    assert "beta_end_times 13.0 63.0 84.0" in text and "kappa_end_times 13.0 40.0 84.0" in text
    assert "beta_1 2.50000000e-01 # [C]" in text and "beta_2 1.25000000e-01" in text
    assert "p 4.50000000e-01 3.50000000e-01 2.50000000e-01 1.50000000e-01 # [C]" in text      # p_0 / p_3 were calibrated
    assert "a 5.00000000e-01 1.00000000e+00 1.00000000e+00 7.00000000e-01" in text
    assert "seed_exposed 1.50000000e+01 # [C]" in text
    a, b = host.read_file("parameters", src, 4), host.read_file("parameters", out, 4)
    assert a.keys() == b.keys()
    for k in a:
        if a[k] is None:                     # beta: NaN is written as "nan", which the reader skips like any non-number
            assert b[k] is None
        else:
            np.testing.assert_allclose(np.array(b[k], dtype=float), np.array(a[k], dtype=float), rtol=5e-9, atol=0)
    with pytest.raises(host.HostError, match="Unable to open file for writing"):
        host.resave_parameters(src, 4, str(tmp_path / "no_such_dir" / "x.txt"), names, 0.0)


def test_reference_tree_reproduces_the_committed_problem(host, problem):
    """The C++ readers on the real reference files give exactly data/spain2020_problem.json (skipped where the
    reference tree is absent, e.g. on the GPU box)."""
    root = os.environ.get("SEPAIHRD_REFERENCE_ROOT", "/root/reference")
    if not os.path.exists(os.path.join(root, "data", "processed", "processed_data.csv")):
        pytest.skip("reference tree not present")
    got = host.load_reference_project(root)
    want = problem.to_json()
    assert got["param_names"] == want["param_names"] and len(got["param_names"]) == 62
    for key in ("times", "obs_hosp", "obs_icu", "obs_deaths", "population", "contact_matrix_rowmajor", "beta_end_times",
                "kappa_end_times", "data_initial_state", "lower_bound", "upper_bound", "sigmas"):
        assert got[key] == want[key], key
    g = np.array([np.nan if x is None else x for x in got["base_slots"]])
    w = np.array(want["base_slots"], dtype=float)
    both = ~(np.isnan(g) | np.isnan(w))
    assert both.sum() >= len(g) - 1 and (g[both] == w[both]).all()
    assert len(got["times"]) == 326 and got["times"][0] == -20.0 and got["times"][-1] == 305.0


def test_written_tree_round_trips_through_both_readers(tmp_path, host, cfg, problem):
    """config.write_reference_tree -> the Python readers and the C++ readers give back the committed Spain-2020 problem
    exactly (this is the tree host/sepaihrd_objective_benchmark is run on where the reference tree is absent)."""
    cfg.write_reference_tree(problem, str(tmp_path))
    py = cfg.problem_from_reference_tree(str(tmp_path)).to_json()
    cc = host.load_reference_project(str(tmp_path))
    want = problem.to_json()
    assert py["param_names"] == cc["param_names"] == want["param_names"]
    for key in ("times", "obs_hosp", "obs_icu", "obs_deaths", "population", "contact_matrix_rowmajor", "beta_end_times",
                "kappa_end_times", "data_initial_state", "lower_bound", "upper_bound", "sigmas"):
        assert py[key] == want[key], key
        assert cc[key] == want[key], key
    assert py["base_slots"] == want["base_slots"] and cc["base_slots"] == want["base_slots"]
