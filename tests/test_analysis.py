"""CPU tests of the post-calibration analysis host code (host/analysis.{hpp,cpp}): MetricsCalculator and
ReproductionNumberCalculator on oracle trajectories against an independent numpy restatement (tests/_metrics_ref.py).
The device side of the scenario analysis is covered by tests/test_gpu_host.py."""
import numpy as np
import pytest

from _metrics_ref import essential_metrics, unpack


@pytest.fixture(scope="module")
def host(pkg):
    from sepaihrd_b200 import hostlib
    hostlib.load_library()
    return hostlib


def _with_beta_scalar(problem, value):
    d = problem.to_json()
    d["base_slots"] = list(d["base_slots"])
    d["base_slots"][problem.layout.beta_scalar] = value          # None -> NaN
    return problem.__class__.from_json(d)


@pytest.mark.parametrize("beta_scalar", [None, 0.0, 0.35])
def test_essential_metrics_match_numpy(host, problem, oracle, beta_scalar):
    """R0 / Rt (power iteration on the n x n block vs numpy eigenvalues of the full 4n x 4n next-generation matrix), peaks,
    Euler-summed infections with the scalar beta (finite) or the schedule fallback (NaN: quirk Q1), seroprevalence at day
    64, per-age ratios."""
    prob = _with_beta_scalar(problem, beta_scalar)
    traj, st = oracle.simulate_batch(problem.base_params()[None])
    assert st[0] == 0
    traj = traj[0]
    x0 = traj[0]
    sc, age, rt, se = host.essential_metrics(prob, prob.times, traj, x0, trajectories=True)
    ref_sc, ref_age, ref_rt, ref_se = essential_metrics(unpack(prob), prob.times, traj, x0)
    np.testing.assert_allclose(rt, ref_rt, rtol=1e-11)
    np.testing.assert_allclose(se, ref_se, rtol=1e-13, atol=1e-18)
    np.testing.assert_allclose(sc, ref_sc, rtol=1e-11)
    np.testing.assert_allclose(age, ref_age, rtol=1e-11)
    names = host.METRIC_NAMES
    assert sc[names.index("R0")] > 1.0 and sc[names.index("final_Rt")] < sc[names.index("max_Rt")]
    assert sc[names.index("time_to_peak_hospital")] > 0 and sc[names.index("total_cumulative_deaths")] > 1e3
    if beta_scalar == 0.0:
        # beta = 0: no infections are added to the initially infected; the ratios are switched off below one infection
        assert sc[names.index("overall_attack_rate")] == pytest.approx(x0.reshape(11, -1)[1:8].sum() / problem.population.sum())
    else:
        assert 1e-3 < sc[names.index("overall_attack_rate")] < 1.0 and (age[:3] <= 1.0).all() and (age[0] > 0).all()


def test_metrics_argument_errors(host, problem):
    with pytest.raises(host.HostError, match="bad argument"):
        host.essential_metrics(problem, [], np.zeros(0), problem.data_initial_state)


def test_sixteen_age_variant_metrics(host, problem, orc):
    p16 = problem.expand_ages(4)
    traj, st = orc.Oracle(p16).simulate_batch(p16.base_params()[None])
    sc, age = host.essential_metrics(p16, p16.times, traj[0], traj[0][0])
    ref_sc, ref_age, _, _ = essential_metrics(unpack(p16), p16.times, traj[0], traj[0][0])
    np.testing.assert_allclose(sc, ref_sc, rtol=1e-10)
    np.testing.assert_allclose(age, ref_age, rtol=1e-10)


def test_analysis_writer_files_have_the_reference_formats(host, tmp_path):
    """AnalysisWriter::writePosteriorPredictiveData (.cpp:283-347) and writeParameterPosteriors (.cpp:201-281): file names,
    headers and number formats -- stream state included (the first time value is printed before the stream turns fixed)."""
    rng = np.random.default_rng(3)
    T, n = 5, 3
    q = np.sort(rng.random((6, T, n, 5)) * 100, axis=-1)
    obs = rng.integers(0, 50, (6, T, n)).astype(float)
    times = np.arange(T, dtype=float) * 1.0
    out = tmp_path / "posterior_predictive"; out.mkdir()
    host.write_posterior_predictive(str(out), times, q, obs)
    names = sorted(p.name for p in out.iterdir())
    series = ["cumulative_deaths", "cumulative_hospitalizations", "cumulative_icu_admissions", "daily_deaths", "daily_hospitalizations", "daily_icu_admissions"]
    assert names == sorted(f"{s}_{w}.csv" for s in series for w in ("median", "lower90", "upper90", "lower95", "upper95", "observed"))
    order = ["daily_hospitalizations", "daily_icu_admissions", "daily_deaths", "cumulative_hospitalizations", "cumulative_icu_admissions", "cumulative_deaths"]
    which = {"lower95": 0, "lower90": 1, "median": 2, "upper90": 3, "upper95": 4}
    for si, s in enumerate(order):
        for w, k in which.items():
            lines = (out / f"{s}_{w}.csv").read_text().splitlines()
            assert lines[0] == "time,age_0,age_1,age_2" and len(lines) == 1 + T
            assert lines[1].split(",")[0] == "0" and lines[2].split(",")[0] == "1.000000"      # default format, then fixed / 6
            got = np.array([[float(v) for v in ln.split(",")[1:]] for ln in lines[1:]])
            np.testing.assert_allclose(got, q[si, :, :, k], atol=5e-7)
            assert all(len(v.split(".")[1]) == 6 for v in lines[1].split(",")[1:])
        lines = (out / f"{s}_observed.csv").read_text().splitlines()
        np.testing.assert_array_equal(np.array([[float(v) for v in ln.split(",")[1:]] for ln in lines[1:]]), obs[si])
    # without observed matrices the files keep the time column only (a matrix with 0 columns)
    out2 = tmp_path / "ppc2"; out2.mkdir()
    host.write_posterior_predictive(str(out2), times, q)
    assert (out2 / "daily_deaths_observed.csv").read_text().splitlines() == ["time", "0", "1", "2", "3", "4"]

    # parameter posteriors
    S, P = 41, 3
    samples = rng.normal(size=(S, P)) * np.array([1.0, 10.0, 1e-3]) + np.array([0.5, -3.0, 2e-2])
    pnames = ["beta_1", "theta", "a_0"]
    out3 = tmp_path / "parameter_posteriors"; out3.mkdir()
    host.write_parameter_posteriors(str(out3), samples, pnames, burn_in=5, thinning=3)
    kept = samples[5::3]
    lines = (out3 / "posterior_samples.csv").read_text().splitlines()
    assert lines[0] == "sample_index,beta_1,theta,a_0" and len(lines) == 1 + len(kept)
    assert lines[1].split(",")[0] == "0" and lines[-1].split(",")[0] == str(len(kept) - 1)
    assert lines[1].split(",")[1] == "%.8e" % kept[0, 0]
    summ = (out3 / "posterior_summary.csv").read_text().splitlines()
    assert summ[0] == "parameter,mean,median,std_dev,lower_95_ci,upper_95_ci" and len(summ) == 1 + P
    for j, ln in enumerate(summ[1:]):
        f = ln.split(",")
        v = np.sort(kept[:, j])
        want = [v.sum() / len(v), v[len(v) // 2], np.sqrt(((v - v.sum() / len(v)) ** 2).sum() / len(v)), v[int(0.025 * len(v))], v[int(0.975 * len(v))]]
        assert f[0] == pnames[j]
        np.testing.assert_allclose([float(x) for x in f[1:]], want, atol=6e-9)
        assert all(len(x.split(".")[1]) == 8 for x in f[1:])
