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
