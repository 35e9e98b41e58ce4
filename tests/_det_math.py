"""Plain-Python restatement of csrc/det_math.h (the log / exp both Metropolis-Hastings samplers use): Python floats are IEEE
binary64 and every +, -, *, / below rounds once, like the host's unfused operations and the device's __d*_rn intrinsics, so the
results agree bit for bit (tests/test_det_math.py)."""
import struct


def _bits(x: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def _from_bits(u: int) -> float:
    return struct.unpack("<d", struct.pack("<Q", u & 0xFFFFFFFFFFFFFFFF))[0]


def det_log(x: float) -> float:
    ln2_hi, ln2_lo = 6.93147180369123816490e-01, 1.90821492927058770002e-10
    Lg1, Lg2, Lg3, Lg4 = 6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01
    Lg5, Lg6, Lg7 = 1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01
    if x == 0.0:
        return float("-inf")
    k = 0
    u = _bits(x)
    if (u >> 52) == 0:
        x = x * 18014398509481984.0
        u = _bits(x)
        k = -54
    hx = (u >> 32) & 0xFFFFFFFF
    k += (hx >> 20) - 1023
    hx &= 0x000FFFFF
    i = (hx + 0x95F64) & 0x100000
    u = ((hx | (i ^ 0x3FF00000)) << 32) | (u & 0xFFFFFFFF)
    k += i >> 20
    f = _from_bits(u) - 1.0
    dk = float(k)
    s = f / (2.0 + f)
    z = s * s
    w = z * z
    t1 = w * (Lg2 + w * (Lg4 + w * Lg6))
    t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)))
    R = t2 + t1
    hfsq = (0.5 * f) * f
    inner = s * (hfsq + R) + dk * ln2_lo
    return dk * ln2_hi - ((hfsq - inner) - f)


def det_exp(x: float) -> float:
    ln2_hi, ln2_lo, invln2 = 6.93147180369123816490e-01, 1.90821492927058770002e-10, 1.44269504088896338700e+00
    P1, P2, P3 = 1.66666666666666019037e-01, -2.77777777770155933842e-03, 6.61375632143793436117e-05
    P4, P5 = -1.65339022054652515390e-06, 4.13813679705723846039e-08
    k = int(invln2 * x + (-0.5 if x < 0.0 else 0.5))          # int() truncates towards zero like the C cast
    t = float(k)
    hi = x - t * ln2_hi
    lo = t * ln2_lo
    r = hi - lo
    r2 = r * r
    c = r - r2 * (P1 + r2 * (P2 + r2 * (P3 + r2 * (P4 + r2 * P5))))
    y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi)
    return y * _from_bits((1023 + k) << 52)
