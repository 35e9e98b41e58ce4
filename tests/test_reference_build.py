"""Pins the oracle against the UNMODIFIED reference when it can be built (VERDICT r1 "What's missing" 3).

`oracle/build_ref.sh` (needs BOOST_ROOT and EIGEN_ROOT: neither library exists in this repository's build image) compiles the
reference's own sources into `oracle/_ref/ref_driver`; this test then evaluates the same seeded parameter sets with the
reference -- its own Boost.Odeint controlled Dopri5 -- and with the oracle, and requires

  * identical (accepted, rejected) step counts for every set (the reference side counts them from integrate_times' return value
    and the number of right-hand-side calls), and
  * log-likelihoods equal to 1e-12 relative (both are unfused IEEE builds of the same arithmetic over the same libm).

Skipped while `oracle/_ref/ref_driver` is absent; with it, row a4 of SURVEY.md section 8 stops being "parity unpinned".
"""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
REFERENCE = os.environ.get("REFERENCE_ROOT", "/root/reference")

pytestmark = pytest.mark.skipif(not (os.path.exists(DRIVER) and os.path.isdir(os.path.join(REFERENCE, "data"))),
                                reason="oracle/_ref/ref_driver not built (run oracle/build_ref.sh with BOOST_ROOT and EIGEN_ROOT) "
                                       "or no reference tree")


def _run_reference(params, mode, tmp_path):
    B, P = params.shape
    fin, fout = tmp_path / "params.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("<qq", B, P))
        f.write(np.ascontiguousarray(params, dtype="<f8").tobytes())
    subprocess.run([DRIVER, REFERENCE, str(fin), str(fout), mode], check=True, timeout=1800)
    raw = open(fout, "rb").read()
    ll = np.frombuffer(raw, dtype="<f8", count=B, offset=0)
    acc = np.frombuffer(raw, dtype="<i8", count=B, offset=8 * B)
    rej = np.frombuffer(raw, dtype="<i8", count=B, offset=16 * B)
    ll2 = np.frombuffer(raw, dtype="<f8", count=B, offset=24 * B)
    return ll, acc, rej, ll2


@pytest.mark.parametrize("mode", ["clamp", "reflect"])
def test_oracle_equals_the_reference_build(problem, orc, tmp_path, mode):
    oracle = orc.Oracle(problem, constraint_mode=0 if mode == "clamp" else 1)
    params = np.vstack([problem.base_params()[None, :], oracle.jitter_params(255, seed=1), oracle.uniform_params(256, seed=2)])
    params[300:] += 3.0 * problem.sigmas * np.random.default_rng(5).standard_normal(params[300:].shape)   # some outside the bounds
    ll_ref, acc, rej, ll_counting = _run_reference(params, mode, tmp_path)
    np.testing.assert_array_equal(ll_ref, ll_counting)        # the counting solver strategy is the reference's, plus counters
    ll, st, steps, _ = oracle.eval_batch(params)
    ok = st == 0
    assert ok.sum() > 400
    np.testing.assert_array_equal(steps[ok, 0], acc[ok])
    np.testing.assert_array_equal(steps[ok, 1], rej[ok])
    assert (np.abs(ll[ok] - ll_ref[ok]) / np.abs(ll_ref[ok])).max() < 1e-12
    np.testing.assert_array_equal(ll[~ok], ll_ref[~ok])       # both return numeric_limits<double>::lowest()
    # the survey's anchor for the shipped parameters
    assert (acc[0], rej[0]) == (441, 45)
