"""csrc/det_math.h -- the log / exp shared by the host and the device-resident Metropolis-Hastings samplers -- against its
plain-Python restatement (bit for bit) and against libm (within one ulp)."""
import math

import numpy as np
import pytest

from _det_math import det_exp, det_log


@pytest.fixture(scope="module")
def L(pkg, cuda_lib):
    import __graft_entry__ as entry
    entry.build()
    from sepaihrd_b200 import hostlib
    return hostlib.load_library()


def test_host_build_equals_the_python_restatement_bit_for_bit(L):
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.random(20000), np.ldexp(rng.random(5000), rng.integers(-70, 3, 5000)), [1.0, 0.5, 2.0, 5e-324, 2.2e-308, 1 - 2 ** -53]])
    for x in xs:
        if x > 0:
            assert L.sepaihrd_host_det_log(float(x)) == det_log(float(x))
    for y in np.concatenate([rng.uniform(-6.9, 2.3, 20000), [0.0, -6.9, 2.3, 1e-300, -1e-17]]):
        assert L.sepaihrd_host_det_exp(float(y)) == det_exp(float(y))
    assert L.sepaihrd_host_det_log(0.0) == -math.inf


def test_accuracy_is_within_one_ulp_of_libm():
    rng = np.random.default_rng(6)
    for x in np.concatenate([rng.random(5000), np.ldexp(rng.random(2000), rng.integers(-64, 0, 2000))]):
        if x > 0:
            ref = math.log(x)
            assert abs(det_log(float(x)) - ref) <= 1.0 * math.ulp(ref) if ref != 0 else det_log(float(x)) == 0.0
    for y in rng.uniform(-6.9, 2.3, 5000):
        ref = math.exp(y)
        assert abs(det_exp(float(y)) - ref) <= 1.0 * math.ulp(ref)
