// det_math.h -- log and exp that give THE SAME BITS on the host and on the device.
//
// The Metropolis-Hastings sampler (reference src/sir_age_structured/optimizers/MetropolisHastingsSampler.cpp) calls libm in
// three places: log(r2) inside std::normal_distribution's polar method (proposal, :91-102), log(u) in the accept test
// (:323-329) and exp(log_scale) in the Robbins-Monro scale (:104-152).  glibc's and CUDA's log / exp agree only to an ulp, so a
// chain sampled on the device would leave the host chain's path at the first draw.  Both samplers of this repository -- the
// host one (host/optimizers.cpp) and the device-resident one (csrc/sepaihrd_mh.cu) -- therefore call these two functions
// instead: the classic argument-reduction + minimax-polynomial kernels (error < 1 ulp), written with IEEE add / mul / div only,
// every operation individually rounded (no FMA contraction on either side), hence bit-identical wherever IEEE binary64 holds.
// tests/_det_math.py restates them in plain Python floats for the sampler restatement tests.
#pragma once

#include <cstdint>
#include <cstring>

#if defined(__CUDA_ARCH__)
#define DETM_FN __device__ __forceinline__
#define DETM_MUL(a, b) __dmul_rn((a), (b))
#define DETM_ADD(a, b) __dadd_rn((a), (b))
#define DETM_SUB(a, b) __dsub_rn((a), (b))
#define DETM_DIV(a, b) __ddiv_rn((a), (b))
#else
#if defined(__CUDACC__)
#define DETM_FN __host__ inline
#else
#define DETM_FN inline
#endif
// Host: every product passes through an empty asm statement, so the compiler cannot contract it with a following add into an
// FMA (-march=native builds would otherwise round once where the device rounds twice); adds, subs and divisions never fuse.
namespace detm {
inline double opaque(double v) {
#if defined(__GNUC__) && (defined(__x86_64__) || defined(__i386__))
    __asm__ volatile("" : "+x"(v));
#elif defined(__GNUC__) && defined(__aarch64__)
    __asm__ volatile("" : "+w"(v));
#elif defined(__GNUC__)
    __asm__ volatile("" : "+m"(v));
#endif
    return v;
}
}  // namespace detm
#define DETM_MUL(a, b) detm::opaque((a) * (b))
#define DETM_ADD(a, b) ((a) + (b))
#define DETM_SUB(a, b) ((a) - (b))
#define DETM_DIV(a, b) ((a) / (b))
#endif

namespace detm {

DETM_FN uint64_t bits_of(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; std::memcpy(&u, &x, 8); return u;
#endif
}
DETM_FN double from_bits(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x; std::memcpy(&x, &u, 8); return x;
#endif
}

// natural logarithm of a finite x >= 0 (0 -> -inf).  x = 2^k (1 + f), sqrt(1/2) <= 1 + f < sqrt(2); s = f / (2 + f);
// log(1 + f) = f - f^2/2 + s (f^2/2 + R(s^2)), R a degree-14 even minimax polynomial; k ln2 added in two pieces.
DETM_FN double log(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    if (x == 0.0) return from_bits(0xfff0000000000000ull);
    int k = 0;
    uint64_t u = bits_of(x);
    if ((u >> 52) == 0) {                                   // subnormal: scale by 2^54 (exact)
        x = DETM_MUL(x, 18014398509481984.0);
        u = bits_of(x);
        k = -54;
    }
    uint32_t hx = (uint32_t)(u >> 32);
    k += (int)(hx >> 20) - 1023;
    hx &= 0x000fffffu;
    const uint32_t i = (hx + 0x95f64u) & 0x100000u;         // mantissa above sqrt(2): halve it, k + 1
    u = ((uint64_t)(hx | (i ^ 0x3ff00000u)) << 32) | (u & 0xffffffffull);
    k += (int)(i >> 20);
    const double f = DETM_SUB(from_bits(u), 1.0);
    const double dk = (double)k;
    const double s = DETM_DIV(f, DETM_ADD(2.0, f));
    const double z = DETM_MUL(s, s), w = DETM_MUL(z, z);
    const double t1 = DETM_MUL(w, DETM_ADD(Lg2, DETM_MUL(w, DETM_ADD(Lg4, DETM_MUL(w, Lg6)))));
    const double t2 = DETM_MUL(z, DETM_ADD(Lg1, DETM_MUL(w, DETM_ADD(Lg3, DETM_MUL(w, DETM_ADD(Lg5, DETM_MUL(w, Lg7)))))));
    const double R = DETM_ADD(t2, t1);
    const double hfsq = DETM_MUL(DETM_MUL(0.5, f), f);
    // k ln2_hi - ((hfsq - (s (hfsq + R) + k ln2_lo)) - f)
    const double inner = DETM_ADD(DETM_MUL(s, DETM_ADD(hfsq, R)), DETM_MUL(dk, ln2_lo));
    return DETM_SUB(DETM_MUL(dk, ln2_hi), DETM_SUB(DETM_SUB(hfsq, inner), f));
}

// exp(x) for |x| < 700: x = k ln2 + r, |r| <= ln2/2; exp(r) = 1 + r + r c / (2 - c), c = r - r^2 P(r^2); result scaled by 2^k.
DETM_FN double exp(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10, invln2 = 1.44269504088896338700e+00;
    const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                 P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    const int k = (int)DETM_ADD(DETM_MUL(invln2, x), (x < 0.0) ? -0.5 : 0.5);      // truncation towards zero = round half away
    const double t = (double)k;
    const double hi = DETM_SUB(x, DETM_MUL(t, ln2_hi));
    const double lo = DETM_MUL(t, ln2_lo);
    const double r = DETM_SUB(hi, lo);
    const double r2 = DETM_MUL(r, r);
    const double c = DETM_SUB(r, DETM_MUL(r2, DETM_ADD(P1, DETM_MUL(r2, DETM_ADD(P2, DETM_MUL(r2, DETM_ADD(P3, DETM_MUL(r2, DETM_ADD(P4, DETM_MUL(r2, P5))))))))));
    const double y = DETM_SUB(1.0, DETM_SUB(DETM_SUB(lo, DETM_DIV(DETM_MUL(r, c), DETM_SUB(2.0, c))), hi));
    return DETM_MUL(y, from_bits((uint64_t)(1023 + k) << 52));                      // exact scaling, |k| <= 1010
}

}  // namespace detm
