"""CPU tests of the host-side logic: reference-format readers, problem assembly, slot mapping in all three
implementations (Python, oracle, product library), and that the C-ABI library loads and exports every
symbol include/sepaihrd_b200.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(cuda_lib):
    hdr = open(os.path.join(ROOT, "include", "sepaihrd_b200.h")).read()
    declared = set(re.findall(r"\b(sepaihrd_[a-z0-9_]+)\s*\(", hdr)) - {"sepaihrd_ctx", "sepaihrd_problem", "sepaihrd_rc"}
    from sepaihrd_b200 import capi
    assert declared == set(capi.SIGNATURES), (declared ^ set(capi.SIGNATURES))
    for name in declared:
        assert getattr(cuda_lib, name) is not None
    assert b"sm_100a" in cuda_lib.sepaihrd_version()


def test_no_cpu_fallback(problem, cuda_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sepaihrd_b200.capi import SepaihrdError
    from sepaihrd_b200.evaluator import BatchEvaluator
    with pytest.raises(SepaihrdError) as ei:
        BatchEvaluator(problem)
    assert ei.value.rc == 2 and "no CPU fallback" in str(ei.value)


def test_create_validates_like_the_reference(problem, cuda_lib):
    """Argument validation runs before the device is touched, so it is testable without a GPU; messages
    follow the reference's InvalidParameterException texts."""
    import copy
    def rc_and_msg(mutate):
        p = copy.deepcopy(problem)
        mutate(p)
        cp = p.as_c()
        h = C.c_void_p()
        rc = cuda_lib.sepaihrd_create(C.byref(cp), 0, C.byref(h))
        return rc, cuda_lib.sepaihrd_last_error().decode()
    rc, msg = rc_and_msg(lambda p: p.times.__setitem__(5, p.times[4]))
    assert rc == 1 and "strictly increasing" in msg                      # Simulator.cpp:82-90
    rc, msg = rc_and_msg(lambda p: p.kappa_end_times.__setitem__(0, -1.0))
    assert rc == 1 and "non-negative" in msg                             # NPI.cpp:24-26
    rc, msg = rc_and_msg(lambda p: setattr(p, "dt_hint", 0.0))
    assert rc == 1 and "positive" in msg                                 # Simulator.cpp:39-41
    rc, msg = rc_and_msg(lambda p: setattr(p, "abs_tol", -1e-6))
    assert rc == 1 and "negative" in msg                                 # Simulator.cpp:46-52


def test_slot_mapping_agrees_in_python_oracle_and_library(problem, oracle, cuda_lib):
    lay = problem.layout
    names = lay.names() + ["beta", "foo", "kappa_1", "kappa_baseline", "kappa_8", "beta_0", "beta_8", "a_4", "h_infec_x",
                           "p_", "d_community_3", "gamma_p", "p_2", "h_2", "icu_0", "d_ICU_1", "runup_days"]
    n, nb, nk = problem.n_ages, len(problem.beta_end_times), len(problem.kappa_end_times)
    for nm in names:
        a = lay.slot_for_name(nm)
        b = oracle.slot_for_name(nm)
        c = cuda_lib.sepaihrd_slot_for_name(n, nb, nk, nm.encode())
        assert a == b == c, (nm, a, b, c)
    assert cuda_lib.sepaihrd_slot_count(n, nb, nk) == lay.count == 64
    # prefix-dispatch order of the reference: "h_infec_1" must not be taken for "h_..."
    assert lay.slot_for_name("h_infec_1") == lay.age("h_infec", 1) != lay.age("h", 1)
    # kappa_1 is the fixed baseline: the reference rejects it at construction (ParameterManager.cpp:80-84)
    assert lay.slot_for_name("kappa_1") == -2


def test_default_problem_shape(problem):
    assert problem.n_ages == 4 and problem.n_times == 326 and problem.n_obs == 306 and problem.n_params == 62
    assert problem.times[0] == -20 and problem.times[-1] == 305                # main.cpp:244-253, runup 20.55 -> int 20
    np.testing.assert_array_equal(problem.population, [14075720, 20948387, 9032069, 2880884])
    np.testing.assert_array_equal(problem.beta_end_times, [13, 63, 84, 111, 183, 237, 305])
    assert problem.base_params().shape == (62,)
    # fixture round trip
    from sepaihrd_b200 import Problem
    p2 = Problem.from_json(problem.to_json())
    for f in ("times", "obs_hosp", "obs_icu", "obs_deaths", "contact_matrix", "base_slots", "lower_bound", "sigmas"):
        np.testing.assert_array_equal(getattr(problem, f), getattr(p2, f))
    # column-major contact matrix in the C struct: data[j*n+i] = M(i, j)
    cp = problem.as_c()
    M = np.ctypeslib.as_array(cp.contact_matrix, shape=(16,))
    assert M[1 * 4 + 0] == problem.contact_matrix[0, 1]


def _write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(textwrap.dedent(text))
    return str(p)


def test_reference_format_readers(tmp_path, pkg):
    cfg = pkg.config
    f = _write(tmp_path, "init.txt", """
        # comment
        beta_end_times  13.0 63.0
        kappa_end_times 13.0 63.0
        beta_1 0.4
        beta_2 0.3
        kappa_1 1.0
        kappa_2 0.5
        a   0.5 0.8 0.9 1.2
        p   0.6 0.3 0.1 0.01
          sigma 0.3   # trailing text is ignored by `iss >> value`
        unknown_key 3
        runup_days 20.55
        """)
    prm = cfg.read_sepaihrd_parameters(f, 4)
    assert prm["beta_values"] == [0.4, 0.3] and prm["kappa_values"] == [1.0, 0.5]
    assert prm["beta_end_times"] == [13.0, 63.0] and prm["sigma"] == 0.3 and prm["runup_days"] == 20.55
    np.testing.assert_array_equal(prm["a"], [0.5, 0.8, 0.9, 1.2])
    with pytest.raises(ValueError):
        cfg.read_sepaihrd_parameters(_write(tmp_path, "bad.txt", "a 1 2 3\n"), 4)      # DataFormatException
    b = cfg.read_param_bounds(_write(tmp_path, "b.txt", "# c\nbeta_1 0.1 0.9\n\nsigma 0.15 0.30\n"))
    assert b == {"beta_1": (0.1, 0.9), "sigma": (0.15, 0.30)}
    with pytest.raises(ValueError):
        cfg.read_param_bounds(_write(tmp_path, "b2.txt", "beta_1 0.1 0.9 7\n"))         # too many values
    s = cfg.read_proposal_sigmas(_write(tmp_path, "s.txt", "beta_1 0.02\nsigma 0.01\n"))
    assert s == {"beta_1": 0.02, "sigma": 0.01}
    assert cfg.read_params_to_calibrate(_write(tmp_path, "c.txt", "# x\nbeta_1\n  sigma  extra\n")) == ["beta_1", "sigma"]
    assert cfg.read_settings(_write(tmp_path, "m.txt", "mcmc_iterations 100\nburn_in 5\n")) == {"mcmc_iterations": 100.0, "burn_in": 5.0}
    m = cfg.read_matrix_csv(_write(tmp_path, "m.csv", "// header comment\n1,2\n\n3,4\n"), 2, 2)
    np.testing.assert_array_equal(m, [[1, 2], [3, 4]])
    with pytest.raises(ValueError):
        cfg.read_matrix_csv(_write(tmp_path, "m2.csv", "1,2\n"), 2, 2)                 # NotEnoughRows


def test_calibration_data_window_and_problem_assembly(tmp_path, pkg):
    """CalibrationData(filename, start, end): string date filter, population from the first kept row."""
    cfg = pkg.config
    sfx = ["0_30", "30_60", "60_80", "80_plus"]
    cols = ["date"] + [f"{pre}_{s}" for pre in ("new_confirmed", "new_deceased", "new_hospitalized_patients",
            "new_intensive_care_patients", "population", "cumulative_confirmed", "cumulative_deceased",
            "cumulative_hospitalized_patients", "cumulative_intensive_care_patients") for s in sfx]
    rows = []
    for d in range(1, 11):
        rows.append([f"2020-03-{d:02d}"] + [str(float(d * 10 + k % 7)) for k in range(len(cols) - 1)])
    path = tmp_path / "data.csv"
    path.write_text(",".join(cols) + "\n" + "\n".join(",".join(r) for r in rows) + "\n")
    cd = cfg.CalibrationData(str(path), "2020-03-03", "2020-03-07")
    assert cd.n_data_points == 5 and cd.dates[0] == "2020-03-03" and cd.dates[-1] == "2020-03-07"
    assert cd.new_hospitalizations.shape == (5, 4)
    assert cd.population[0] == float(rows[2][1 + 16])
    with pytest.raises(ValueError):
        cfg.CalibrationData(str(path), "2021-01-01", "2021-02-01")


def test_expand_ages_preserves_totals(problem):
    p16 = problem.expand_ages(4)
    assert p16.n_ages == 16 and p16.layout.count == 7 + 7 + 7 + 8 * 16 + 11
    np.testing.assert_allclose(p16.population.reshape(4, 4).sum(1), problem.population)
    np.testing.assert_allclose(p16.obs_hosp.reshape(-1, 4, 4).sum(2), problem.obs_hosp)
    # contact structure: sum over the sub-classes of j reproduces M(i, j)
    np.testing.assert_allclose(p16.contact_matrix.reshape(4, 4, 4, 4)[:, 0].sum(-1), problem.contact_matrix)
    assert "h_infec_15" in p16.param_names and len(p16.param_names) == 62 + 3 * 32


def test_select_ages_builds_consistent_sub_problems(problem, orc):
    """Problem.select_ages (fixtures for the zero-padded kernel path): the identity selection is the problem itself, a
    permutation leaves the likelihood unchanged up to summation order, and a sub-problem drops exactly the per-age parameters
    of the removed classes."""
    same = problem.select_ages([0, 1, 2, 3])
    assert same.param_names == problem.param_names
    np.testing.assert_array_equal(same.base_slots, problem.base_slots)
    np.testing.assert_array_equal(same.contact_matrix, problem.contact_matrix)
    x = problem.base_params()
    ll = orc.Oracle(problem).eval_one(x)["ll"]
    perm = problem.select_ages([3, 1, 0, 2])
    xp = perm.base_params()
    assert sorted(perm.param_names) == sorted(problem.param_names)
    assert abs(orc.Oracle(perm).eval_one(xp)["ll"] - ll) <= 1e-9 * abs(ll)
    sub = problem.select_ages([0, 3])
    assert sub.n_ages == 2 and sub.n_params == problem.n_params - 2 * 8 and sub.layout.count == problem.layout.count - 2 * 8
    assert "h_infec_1" in sub.param_names and "h_infec_2" not in sub.param_names
    i_old, i_new = problem.param_names.index("h_infec_3"), sub.param_names.index("h_infec_1")
    assert problem.base_params()[i_old] == sub.base_params()[i_new]
    with pytest.raises(ValueError):
        problem.select_ages([0, 4])


def test_device_window_length_follows_the_launch_cost_steps(pkg):
    """resident.window_length: iterations per look-ahead window of the device-resident chains from the measured steps of the launch
    cost (592 warp tiles of 8 sets at one warp per scheduler, 1184 at two): more chains per GPU -> shorter windows, never more
    proposals than a second wave would start for, an explicit request is clamped to 1 .. 64."""
    from sepaihrd_b200 import resident
    ks = [resident.window_length(n) for n in (1, 64, 256, 512, 1024, 2048, 4096, 8192, 65536)]
    assert ks == sorted(ks, reverse=True) and ks[0] >= 8 and ks[-1] == 1
    assert resident.window_length(512) * 512 <= 4736 and resident.window_length(2048) * 2048 <= 9472 and resident.window_length(4096) == 2
    assert resident.window_length(100, 0) == 1 and resident.window_length(100, 1000) == 64 and resident.window_length(100, 7) == 7
    assert resident.window_length(512, rate=0.05) >= resident.window_length(512, rate=0.5)          # rare accepts: longer windows pay
