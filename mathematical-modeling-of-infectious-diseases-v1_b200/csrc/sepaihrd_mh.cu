// sepaihrd_mh.cu -- Metropolis-Hastings chains that LIVE ON THE DEVICE: chain states, their std::mt19937 generators, the
// Robbins-Monro scales and the accept history never leave HBM; an iteration is three launches on the ctx stream
// (propose -> the fused likelihood kernel -> accept) and needs no host work at all.
//
// Replaces, for chains scored by this evaluator, the per-iteration host work of the reference's sampler
// (src/sir_age_structured/optimizers/MetropolisHastingsSampler.cpp), one instance per chain:
//   generateProposal   :91-102    y = x + s L z, z ~ N(0, I) from std::normal_distribution over the chain's std::mt19937
//   applyConstraints   src/model/parameters/SEPAIHRDParameterManager.cpp:302-347 (mirror reflection in MCMC mode)
//   accept             :318-330   log-space test, the uniform drawn ONLY for downhill proposals
//   adaptGlobalScale   :104-152   Robbins-Monro on the last <= 1000 accept decisions
// Scope: the phase with a FIXED proposal kernel (t <= burn_in: the start kernel built from the proposal sigmas, or the
// covariance handed over by phase 1); the Haario covariance adaptation after burn-in (:154-199) stays in the host sampler
// (host/optimizers.cpp), and sepaihrd_mh_create refuses a run that would need it.
//
// Bit-compatibility with the host sampler (host/optimizers.cpp, itself pinned against a Python restatement of the reference):
// chain c draws from std::mt19937(std::seed_seq{seed, c}) with c the GLOBAL chain index -- the device runs seed_seq::generate,
// the MT19937 twist / tempering, libstdc++'s generate_canonical<double, 53> and its polar normal_distribution -- every FP64
// operation is individually rounded in the host's order, and log / exp are csrc/det_math.h on both sides.  A device-resident
// run therefore makes EXACTLY the host run's accept decisions and visits its states (tests/test_gpu_host.py).
//
// Mapping: one WARP per chain.  The polar method consumes exactly four generator words per attempt, so the attempts of a
// proposal are independent given their position in the stream: the 32 lanes evaluate 32 attempts at once (uniforms, r^2, log,
// sqrt), a ballot + prefix count hands the accepted ones their place among the P normals, and the position of the last one
// needed tells how many words the chain consumed -- the same numbers, in the same order, as the sequential host loop.  The
// generator state is chain-major ([chains][624]); a warp twists 32 consecutive words per step (word k needs the old k, k + 1 and
// the word 397 ahead or 227 behind: never one of the same step), loads before stores.  ~15 us per iteration for 4096 chains
// against >= 550 us for the likelihood launch (the first version, one thread per chain, took 220 us).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <vector>

#include "det_math.h"
#include "sepaihrd_constraints.cuh"
#include "sepaihrd_internal.h"

struct sepaihrd_mh {
    sepaihrd_ctx* ctx = nullptr;
    int P = 0;
    long long n_chains = 0, offset = 0, local = 0;
    sepaihrd_mh_settings cfg{};
    int t = 1;                                 // next iteration (1-based like the reference loop)
    bool begun = false, diagonal = true;
    char* d_arena = nullptr;
    double *d_x = nullptr, *d_prop = nullptr, *d_lp = nullptr, *d_plp = nullptr, *d_log_scale = nullptr, *d_scale = nullptr;
    double *d_best_lp = nullptr, *d_best_x = nullptr, *d_chol = nullptr, *d_lo = nullptr, *d_hi = nullptr, *d_init = nullptr;
    double* d_trace = nullptr;                 // [trace_cap]: max over all chains' log-posteriors per iteration (sepaihrd_mh_note_gathered)
    unsigned *d_mt = nullptr, *d_C = nullptr, *d_T = nullptr, *d_recent = nullptr, *d_status = nullptr;
    int *d_recent_n = nullptr, *d_recent_sum = nullptr, *d_emergency = nullptr;
    long long* d_accepted = nullptr;
    unsigned char* d_accepts = nullptr;        // [iterations - 1][local] when record_accepts
    int trace_cap = 0;
    // look-ahead windows (sepaihrd_mh_window_*): every chain has its own iteration counter
    bool windowed = false;
    int win_cap = 0, win_K = 0;                // proposals per chain the window buffers hold / of the window in flight
    char* d_win = nullptr;
    int* d_t = nullptr;                        // [local] next iteration of every chain
    int* d_tmin = nullptr;                     // min over d_t after the last commit
    unsigned* d_ghost = nullptr;               // [local][624] generator copies the proposals of a window are drawn from
    double *d_wprop = nullptr, *d_wplp = nullptr, *d_wlogu = nullptr, *d_record = nullptr;
    unsigned *d_wCz = nullptr, *d_wstatus = nullptr;
};

namespace {

using sepaihrd_internal::fail_with;

#define MH_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) return fail_with(SEPAIHRD_ERR_CUDA, cudaGetErrorString(e__));     \
    } while (0)

constexpr int MT_N = 624, MT_M = 397;
constexpr int MH_THREADS = 128;                // 4 chains per block
constexpr int MH_WARPS = MH_THREADS / 32;
constexpr int RECENT_WORDS = 32;               // ring of the last 1000 accept decisions, one bit each
constexpr unsigned FULLM = 0xffffffffu;

// std::mt19937 whose 624-word state lives in global memory ([chains][624]).  C = outputs consumed, T = state words twisted, both
// counted from the seeding (T a multiple of 32; they wrap at 2^32 and only their difference and residues mod 624 are used --
// 2^32 is not a multiple of 624, so a chain may draw at most 2^32 - 1 words: ~2.7e7 iterations of a 62-parameter chain).
// Twisting [T, T + 32) overwrites the outputs T - 624 ..., which must have been consumed: T <= C + 592 (callers keep T - C <= 544).
__device__ __forceinline__ void warp_twist32(unsigned* st, unsigned T, int lane) {
    int k = (int)(T % MT_N) + lane; if (k >= MT_N) k -= MT_N;
    const int k1 = (k + 1 == MT_N) ? 0 : k + 1;
    const int km = (k + MT_M >= MT_N) ? k + MT_M - MT_N : k + MT_M;
    const unsigned a = st[k], b = st[k1], f = st[km];
    __syncwarp();                                                      // every lane has loaded before any lane stores
    const unsigned y = (a & 0x80000000u) | (b & 0x7fffffffu);
    st[k] = f ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    __syncwarp();
}
__device__ __forceinline__ unsigned mt_output(const unsigned* st, unsigned pos) {      // tempered output number `pos`
    unsigned y = st[pos % MT_N];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
// generate_canonical<double, 53>: (lo + hi 2^32) / 2^64, one rounding in the sum; 1.0 -> the largest double below 1
__device__ __forceinline__ double canonical(unsigned lo, unsigned hi) {
    const double sum = __dadd_rn((double)lo, __dmul_rn((double)hi, 4294967296.0));
    const double r = __dmul_rn(sum, 1.0 / 18446744073709551616.0);
    return (r >= 1.0) ? 0.99999999999999988898 : r;
}

// std::seed_seq{seed, chain}.generate over 624 words ([rand.util.seedseq]) = the state of std::mt19937 seeded from it;
// the first output comes after a twist.  Local scratch (lane-interleaved, L1-resident), then one pass to the state.
__global__ void __launch_bounds__(64) mh_seed_kernel(long long local, long long offset, unsigned seed, unsigned* __restrict__ mt,
                                                      unsigned* __restrict__ Cc, unsigned* __restrict__ Tt) {
    const long long c = blockIdx.x * (long long)64 + threadIdx.x;
    if (c >= local) return;
    const unsigned v[2] = {seed, (unsigned)(offset + c)};
    constexpr int n = MT_N, s = 2, t = 11, p = (n - t) / 2, q = p + t, m = n;      // m = max(s + 1, n)
    unsigned b[MT_N];
    for (int i = 0; i < n; ++i) b[i] = 0x8b8b8b8bu;
    for (int k = 0; k < m; ++k) {
        const int kn = k % n, kp = (k + p) % n, kq = (k + q) % n, k1 = (k + n - 1) % n;
        unsigned r1 = b[kn] ^ b[kp] ^ b[k1];
        r1 = 1664525u * (r1 ^ (r1 >> 27));
        const unsigned r2 = r1 + ((k == 0) ? (unsigned)s : (k <= s) ? (unsigned)kn + v[k - 1] : (unsigned)kn);
        b[kp] += r1;
        b[kq] += r2;
        b[kn] = r2;
    }
    for (int k = m; k < m + n; ++k) {
        const int kn = k % n, kp = (k + p) % n, kq = (k + q) % n, k1 = (k + n - 1) % n;
        unsigned r3 = b[kn] + b[kp] + b[k1];
        r3 = 1566083941u * (r3 ^ (r3 >> 27));
        const unsigned r4 = r3 - (unsigned)kn;
        b[kp] ^= r3;
        b[kq] ^= r4;
        b[kn] = r4;
    }
    // mersenne_twister_engine::seed(Sseq&): an all-zero state (upper bit of word 0, all of the others) becomes 2^31
    bool zero = (b[0] & 0x80000000u) == 0u;
    for (int i = 1; i < n && zero; ++i) zero = (b[i] == 0u);
    if (zero) b[0] = 0x80000000u;
    for (int i = 0; i < n; ++i) mt[c * MT_N + i] = b[i];
    Cc[c] = 0u; Tt[c] = 0u;
}

// every chain starts at `init` with the log-posterior of that point (evaluated once, d_plp[0]); safeValue like the host
__global__ void mh_start_kernel(long long local, int P, const double* __restrict__ init, const double* __restrict__ lp0, double* __restrict__ x,
                                double* __restrict__ lp, double* __restrict__ log_scale, double* __restrict__ scale, double* __restrict__ best_lp,
                                double* __restrict__ best_x, unsigned* __restrict__ recent, int* __restrict__ recent_n, int* __restrict__ recent_sum,
                                int* __restrict__ emergency, long long* __restrict__ accepted) {
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (c >= local) return;
    double v = lp0[0];
    if (isnan(v) || isinf(v)) v = -1e18;
    for (int k = 0; k < P; ++k) { x[c * P + k] = init[k]; best_x[c * P + k] = init[k]; }
    lp[c] = v; best_lp[c] = v;
    log_scale[c] = 0.0; scale[c] = 1.0;
    for (int w = 0; w < RECENT_WORDS; ++w) recent[c * RECENT_WORDS + w] = 0u;
    recent_n[c] = 0; recent_sum[c] = 0; emergency[c] = 0; accepted[c] = 0;
}

// std::normal_distribution<double>(0, 1), a FRESH object per proposal (the reference declares it inside generateProposal):
// polar method; the first value of an accepted pair is y * mult, the second (returned by the next call) x * mult.
// Attempt j of this proposal reads the outputs C + 4 j ... C + 4 j + 3 whatever happened to the attempts before it.
// One warp: z[0 .. P) in shared memory, C / T advanced.
__device__ __forceinline__ void warp_draw_normals(unsigned* st, unsigned& C, unsigned& T, int lane, int P, double* z, unsigned* fault) {
    const int npairs = (P + 1) / 2;
    int got = 0, consumed = -1;
    for (int base = 0; base < 128 && consumed < 0; base += 32) {
        while ((unsigned)(T - C) < (unsigned)(4 * (base + 32))) { warp_twist32(st, T, lane); T += 32; }
        const unsigned w0 = C + 4u * (unsigned)(base + lane);
        const double a = __dsub_rn(__dmul_rn(2.0, canonical(mt_output(st, w0), mt_output(st, w0 + 1))), 1.0);
        const double b = __dsub_rn(__dmul_rn(2.0, canonical(mt_output(st, w0 + 2), mt_output(st, w0 + 3))), 1.0);
        const double r2 = __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b));
        const bool ok = !(r2 > 1.0 || r2 == 0.0);
        const unsigned mask = __ballot_sync(FULLM, ok);
        const int rank = got + __popc(mask & ((1u << lane) - 1u));
        if (ok && rank < npairs) {
            const double mult = __dsqrt_rn(__ddiv_rn(__dmul_rn(-2.0, detm::log(r2)), r2));
            z[2 * rank] = __dmul_rn(b, mult);
            if (2 * rank + 1 < P) z[2 * rank + 1] = __dmul_rn(a, mult);
        }
        const int cnt = __popc(mask);
        if (got + cnt >= npairs) consumed = 4 * (base + (int)__fns(mask, 0, npairs - got) + 1);
        got += cnt;
    }
    if (consumed < 0) {                 // 128 attempts without P normals: probability ~1e-40; flagged, never silently wrong
        if (lane == 0) atomicExch(fault, 1u);
        consumed = 4 * 128;
        for (int i = 2 * got + lane; i < P; i += 32) z[i] = 0.0;
    }
    C += (unsigned)consumed;
    __syncwarp();
}

// y = constrain(x + s L z): row i of L z summed over the columns j = 0 .. i in the host's order (the structural zeros add nothing)
__device__ __forceinline__ void warp_make_proposal(int lane, int P, int diagonal, int mode, const double* __restrict__ chol, const double* __restrict__ lo,
                                                   const double* __restrict__ hi, const double* __restrict__ xc, double sc, const double* z,
                                                   double* __restrict__ out) {
    for (int i = lane; i < P; i += 32) {
        double step;
        if (diagonal) {
            step = __dmul_rn(chol[(long long)i * P + i], z[i]);
        } else {
            step = 0.0;
            for (int j = 0; j <= i; ++j) step = __dadd_rn(step, __dmul_rn(chol[(long long)j * P + i], z[j]));
        }
        out[i] = sepaihrd::constrain(__dadd_rn(xc[i], __dmul_rn(sc, step)), lo[i], hi[i], mode);
    }
}

// adaptGlobalScale (.cpp:104-152) for one decision at the reference's 1-based iteration `step`: the ring of the last <= 1000
// decisions (one bit each; the host pushes, then drops the oldest once there are more than 1000), Robbins-Monro on its rate
__device__ __forceinline__ void scale_update(bool acc, int step, double target, unsigned* ring, int& n, int& sum, int& emergency, double& ls, double& sc) {
    const int slot = n % 1000;
    unsigned* wp = ring + (slot >> 5);
    unsigned w = *wp;
    const unsigned bit = 1u << (slot & 31);
    if (n >= 1000) sum -= (w & bit) ? 1 : 0;
    w = acc ? (w | bit) : (w & ~bit);
    *wp = w;
    sum += acc ? 1 : 0;
    n += 1;
    const int size = n < 1000 ? n : 1000;
    const double rate = __ddiv_rn((double)sum, (double)size);
    const double s1 = __dadd_rn((double)step, 1.0);
    if (size >= 1000 && rate < 0.001) {
        ls = __dsub_rn(ls, 0.7);
        emergency += 1;
    } else if (rate < 0.02 && size >= 500) {
        const double g5 = __ddiv_rn(5.0, __dsqrt_rn(s1));
        const double gg = (0.3 < g5) ? 0.3 : g5;                                  // std::min(g5, 0.3)
        ls = __dadd_rn(ls, __dmul_rn(gg, __dsub_rn(0.0, target)));
    } else {
        const double g1 = __ddiv_rn(1.0, __dsqrt_rn(s1));
        const double gg = (0.1 < g1) ? 0.1 : g1;
        ls = __dadd_rn(ls, __dmul_rn(gg, __dsub_rn(acc ? 1.0 : 0.0, target)));
    }
    if (sc <= 0.011 && rate > 0.15 && rate < 0.30) ls = __dadd_rn(ls, 0.01);
    {   // std::max(std::min(ls, 2.3), -6.9)
        const double lo_c = (2.3 < ls) ? 2.3 : ls;
        ls = (lo_c < -6.9) ? -6.9 : lo_c;
    }
    sc = detm::exp(ls);
}

// generateProposal + applyConstraints: one warp per local chain
__global__ void __launch_bounds__(MH_THREADS) mh_propose_kernel(long long local, int P, int diagonal, int mode, const double* __restrict__ chol,
                                                                 const double* __restrict__ lo, const double* __restrict__ hi,
                                                                 const double* __restrict__ x, const double* __restrict__ scale,
                                                                 unsigned* __restrict__ mt, unsigned* __restrict__ Cc, unsigned* __restrict__ Tt,
                                                                 double* __restrict__ prop, unsigned* __restrict__ fault) {
    __shared__ double zs[MH_WARPS][SEPAIHRD_MH_MAX_PARAMS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long c = blockIdx.x * (long long)MH_WARPS + wib;
    if (c >= local) return;
    unsigned* st = mt + c * MT_N;
    unsigned C = Cc[c], T = Tt[c];
    double* z = zs[wib];
    warp_draw_normals(st, C, T, lane, P, z, fault);
    warp_make_proposal(lane, P, diagonal, mode, chol, lo, hi, x + c * P, scale[c], z, prop + c * P);
    if (lane == 0) { Cc[c] = C; Tt[c] = T; }
}

// accept + adaptGlobalScale: one warp per local chain (lane 0 decides and keeps the books, the warp moves the state);
// `step` is the reference's 1-based iteration index
__global__ void __launch_bounds__(MH_THREADS) mh_accept_kernel(long long local, int P, int step, int adapt_scale, double target,
                                                                const double* __restrict__ plp_in, const double* __restrict__ prop,
                                                                double* __restrict__ x, double* __restrict__ lp, unsigned* __restrict__ mt,
                                                                unsigned* __restrict__ Cc, unsigned* __restrict__ Tt, double* __restrict__ log_scale,
                                                                double* __restrict__ scale, unsigned* __restrict__ recent, int* __restrict__ recent_n,
                                                                int* __restrict__ recent_sum, int* __restrict__ emergency,
                                                                long long* __restrict__ accepted, double* __restrict__ best_lp,
                                                                double* __restrict__ best_x, unsigned char* __restrict__ accepts_row) {
    const int lane = threadIdx.x & 31;
    const long long c = blockIdx.x * (long long)MH_WARPS + (threadIdx.x >> 5);
    if (c >= local) return;
    double plp = plp_in[c];
    if (isnan(plp) || isinf(plp)) plp = -1e18;                                  // safeEvaluate, .cpp:65-74
    const double clp = lp[c];
    const double log_ratio = __dsub_rn(plp, clp);
    bool acc = false;
    if (log_ratio >= 0.0) {
        acc = true;
    } else {                                                                    // the uniform is drawn ONLY for downhill proposals
        unsigned* st = mt + c * MT_N;
        unsigned C = Cc[c], T = Tt[c];
        while ((unsigned)(T - C) < 2u) { warp_twist32(st, T, lane); T += 32; }
        const double u = canonical(mt_output(st, C), mt_output(st, C + 1));
        if (detm::log(u) < log_ratio) acc = true;
        __syncwarp();
        if (lane == 0) { Cc[c] = C + 2u; Tt[c] = T; }
    }
    const bool better = acc && plp > best_lp[c];
    __syncwarp();
    if (acc)
        for (int k = lane; k < P; k += 32) {
            const double v = prop[c * P + k];
            x[c * P + k] = v;
            if (better) best_x[c * P + k] = v;
        }
    if (lane != 0) return;
    if (acc) {
        lp[c] = plp;
        accepted[c] += 1;
        if (better) best_lp[c] = plp;
    }
    if (accepts_row) accepts_row[c] = acc ? 1 : 0;
    if (!adapt_scale) return;
    int n = recent_n[c], sum = recent_sum[c], em = emergency[c];
    double ls = log_scale[c], sc = scale[c];
    scale_update(acc, step, target, recent + c * RECENT_WORDS, n, sum, em, ls, sc);
    recent_n[c] = n; recent_sum[c] = sum; emergency[c] = em;
    log_scale[c] = ls;
    scale[c] = sc;
}

// ---- look-ahead windows -----------------------------------------------------------------------------------------------------
// A launch of the likelihood kernel costs the same from 1 to ~4 000 sets, and while a chain rejects it does not move: the
// proposals of its next K iterations are known before any of them is scored -- iteration t + j proposes x + s_j L z_j with z_j the
// generator's next normals (a rejected iteration was a downhill one and has consumed its uniform, .cpp:323-329) and s_j the scale
// after j more rejections (.cpp:104-152).  mh_window_propose_kernel draws them from a COPY of the chain's generator and scale
// state; the likelihood kernel scores local x K proposals in one launch; mh_window_commit_kernel replays the sequential loop up
// to and including the first accepted proposal and leaves the generator where the sequential run would have left it (the
// number of words consumed is all that has to be carried over: the state is twisted forward to it).  Chains advance by
// different amounts per window, so each has its own iteration counter.  Same decisions, same states as the one-iteration-per-
// launch loop above and as the host sampler (tests/test_gpu_resident.py).
__global__ void __launch_bounds__(MH_THREADS) mh_window_propose_kernel(long long local, int P, int K, int iterations, int diagonal, int mode, int adapt_scale,
                                                                        double target, const double* __restrict__ chol, const double* __restrict__ lo,
                                                                        const double* __restrict__ hi, const double* __restrict__ x,
                                                                        const double* __restrict__ log_scale, const double* __restrict__ scale,
                                                                        const unsigned* __restrict__ recent, const int* __restrict__ recent_n,
                                                                        const int* __restrict__ recent_sum, const unsigned* __restrict__ mt,
                                                                        const unsigned* __restrict__ Cc, const unsigned* __restrict__ Tt,
                                                                        const int* __restrict__ t_next, unsigned* __restrict__ ghost,
                                                                        double* __restrict__ wprop, unsigned* __restrict__ wCz, double* __restrict__ wlogu,
                                                                        unsigned* __restrict__ fault) {
    __shared__ double zs[MH_WARPS][SEPAIHRD_MH_MAX_PARAMS];
    __shared__ unsigned rings[MH_WARPS][RECENT_WORDS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long c = blockIdx.x * (long long)MH_WARPS + wib;
    if (c >= local) return;
    const int t0 = t_next[c];
    const int k_eff = max(0, min(K, iterations - t0));
    unsigned* st = ghost + c * MT_N;
    for (int i = lane; i < MT_N; i += 32) st[i] = mt[c * MT_N + i];
    rings[wib][lane] = recent[c * RECENT_WORDS + lane];
    __syncwarp();
    unsigned C = Cc[c], T = Tt[c];
    int n = recent_n[c], sum = recent_sum[c], em = 0;
    double ls = log_scale[c], sc = scale[c];
    double* z = zs[wib];
    const double* xc = x + c * P;
    for (int j = 0; j < K; ++j) {
        double* out = wprop + (c * K + j) * P;
        if (j >= k_eff) {                                     // past the chain's last iteration: a valid point, scored and ignored
            for (int i = lane; i < P; i += 32) out[i] = xc[i];
            continue;
        }
        warp_draw_normals(st, C, T, lane, P, z, fault);
        warp_make_proposal(lane, P, diagonal, mode, chol, lo, hi, xc, sc, z, out);
        while ((unsigned)(T - C) < 2u) { warp_twist32(st, T, lane); T += 32; }
        const double u = canonical(mt_output(st, C), mt_output(st, C + 1));       // the uniform of a rejected (downhill) proposal
        if (lane == 0) { wCz[c * K + j] = C; wlogu[c * K + j] = detm::log(u); }
        C += 2u;
        if (adapt_scale) {
            if (lane == 0) scale_update(false, t0 + j, target, rings[wib], n, sum, em, ls, sc);
            sc = __shfl_sync(FULLM, sc, 0);
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(MH_THREADS) mh_window_commit_kernel(long long local, int P, int K, int iterations, int adapt_scale, double target,
                                                                       const double* __restrict__ wplp, const double* __restrict__ wprop,
                                                                       const unsigned* __restrict__ wCz, const double* __restrict__ wlogu,
                                                                       double* __restrict__ x, double* __restrict__ lp, unsigned* __restrict__ mt,
                                                                       unsigned* __restrict__ Cc, unsigned* __restrict__ Tt, double* __restrict__ log_scale,
                                                                       double* __restrict__ scale, unsigned* __restrict__ recent, int* __restrict__ recent_n,
                                                                       int* __restrict__ recent_sum, int* __restrict__ emergency,
                                                                       long long* __restrict__ accepted, double* __restrict__ best_lp,
                                                                       double* __restrict__ best_x, unsigned char* __restrict__ accepts,
                                                                       int* __restrict__ t_next, int* __restrict__ t_min, double* __restrict__ record) {
    const int lane = threadIdx.x & 31;
    const long long c = blockIdx.x * (long long)MH_WARPS + (threadIdx.x >> 5);
    if (c >= local) return;
    int t = t_next[c];
    const int k_eff = max(0, min(K, iterations - t));
    int jacc = -1, better = 0;
    unsigned Cfin = Cc[c];
    double clp = lp[c];
    if (lane == 0 && k_eff > 0) {                             // the sequential loop, one decision after the other
        int n = recent_n[c], sum = recent_sum[c], em = emergency[c];
        double ls = log_scale[c], sc = scale[c];
        for (int j = 0; j < k_eff; ++j) {
            double plp = wplp[c * K + j];
            if (isnan(plp) || isinf(plp)) plp = -1e18;                          // safeEvaluate, .cpp:65-74
            const double log_ratio = __dsub_rn(plp, clp);
            bool acc;
            if (log_ratio >= 0.0) { acc = true; Cfin = wCz[c * K + j]; }          // uphill: no uniform is drawn
            else { acc = wlogu[c * K + j] < log_ratio; Cfin = wCz[c * K + j] + 2u; }
            if (accepts) accepts[(size_t)(t - 1) * local + c] = acc ? 1 : 0;
            if (adapt_scale) scale_update(acc, t, target, recent + c * RECENT_WORDS, n, sum, em, ls, sc);
            t += 1;
            if (acc) {
                jacc = j;
                better = plp > best_lp[c];
                clp = plp;
                accepted[c] += 1;
                if (better) best_lp[c] = plp;
                break;
            }
        }
        recent_n[c] = n; recent_sum[c] = sum; emergency[c] = em;
        log_scale[c] = ls; scale[c] = sc;
        lp[c] = clp;
        t_next[c] = t;
    }
    jacc = __shfl_sync(FULLM, jacc, 0);
    better = __shfl_sync(FULLM, better, 0);
    Cfin = __shfl_sync(FULLM, Cfin, 0);
    clp = __shfl_sync(FULLM, clp, 0);
    t = __shfl_sync(FULLM, t, 0);
    if (jacc >= 0)
        for (int k = lane; k < P; k += 32) {
            const double v = wprop[(c * K + jacc) * P + k];
            x[c * P + k] = v;
            if (better) best_x[c * P + k] = v;
        }
    // the generator: exactly the words of the committed iterations are consumed; twist the state forward to them
    unsigned T = Tt[c];
    unsigned* st = mt + c * MT_N;
    while ((int)(T - Cfin) < 0) { warp_twist32(st, T, lane); T += 32; }
    if (lane == 0) {
        Cc[c] = Cfin; Tt[c] = T;
        atomicMin(t_min, t);
        if (record) record[c] = clp;
    }
}

// the last slot of the rank's record: the smallest next-iteration index among its chains (a rank without chains is done)
__global__ void mh_window_publish_kernel(int* t_min, int iterations, long long local, double* slot) {
    if (local == 0) *t_min = iterations;
    *slot = (double)min(*t_min, iterations);
}

// trace[slot] = max over the gathered log-posteriors of ALL chains ([world][stride] blocks of counts[r] valid values each)
__global__ void mh_trace_kernel(const double* __restrict__ all_lp, int world, long long stride, long long base, long long extra,
                                double* __restrict__ trace, int slot) {
    __shared__ double s[256];
    double m = -INFINITY;
    for (int r = 0; r < world; ++r) {
        const long long cnt = base + (r < extra ? 1 : 0);
        for (long long i = threadIdx.x; i < cnt; i += blockDim.x) { const double v = all_lp[r * stride + i]; if (v > m) m = v; }
    }
    s[threadIdx.x] = m;
    __syncthreads();
    for (int off = 128; off >= 1; off >>= 1) {
        if (threadIdx.x < off && s[threadIdx.x + off] > s[threadIdx.x]) s[threadIdx.x] = s[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) trace[slot] = s[0];
}

}  // namespace

extern "C" {

sepaihrd_rc sepaihrd_mh_create(sepaihrd_ctx* ctx, int64_t n_chains, int64_t chain_offset, int64_t local_count,
                               const sepaihrd_mh_settings* settings, sepaihrd_mh** out) {
    if (!ctx || !out || !settings) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (n_chains <= 0 || chain_offset < 0 || local_count < 0 || chain_offset + local_count > n_chains)
        return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "chain_offset/local_count outside the set of chains");
    if (settings->iterations < 1) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "iterations must be >= 1");
    if (settings->iterations - 1 > settings->burn_in)
        return fail_with(SEPAIHRD_ERR_UNSUPPORTED, "the device-resident sampler covers the fixed-kernel phase (iterations - 1 <= burn_in); the covariance "
                                                   "adaptation after burn-in runs in the host sampler");
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(ctx);
    if (d.P > SEPAIHRD_MH_MAX_PARAMS) return fail_with(SEPAIHRD_ERR_UNSUPPORTED, "too many parameters for the device-resident sampler");
    MH_TRY(cudaSetDevice(d.device));
    auto* m = new sepaihrd_mh;
    m->ctx = ctx; m->P = d.P; m->n_chains = n_chains; m->offset = chain_offset; m->local = local_count; m->cfg = *settings;
    m->trace_cap = settings->iterations + 1;
    const size_t L = (size_t)local_count, LP = L * (size_t)d.P;
    size_t bytes = 0;
    auto reserve = [&](size_t b) { const size_t at = bytes; bytes += (b + 255) & ~(size_t)255; return at; };
    const size_t o_x = reserve(8 * LP), o_prop = reserve(8 * LP), o_lp = reserve(8 * L), o_plp = reserve(8 * (L + 1)), o_ls = reserve(8 * L),
                 o_sc = reserve(8 * L), o_blp = reserve(8 * L), o_bx = reserve(8 * LP), o_chol = reserve(8 * (size_t)d.P * d.P),
                 o_lo = reserve(8 * (size_t)d.P), o_hi = reserve(8 * (size_t)d.P), o_init = reserve(8 * (size_t)d.P),
                 o_trace = reserve(8 * (size_t)m->trace_cap), o_mt = reserve(4 * L * MT_N), o_C = reserve(4 * L), o_T = reserve(4 * L),
                 o_rec = reserve(4 * L * RECENT_WORDS), o_st = reserve(4 * (L + 2)), o_rn = reserve(4 * L), o_rs = reserve(4 * L),
                 o_em = reserve(4 * L), o_acc = reserve(8 * L),
                 o_accepts = reserve(settings->record_accepts ? L * (size_t)settings->iterations : 1);
    cudaError_t e = cudaMalloc((void**)&m->d_arena, bytes);
    if (e == cudaSuccess) {
        char* D = m->d_arena;
        m->d_x = (double*)(D + o_x); m->d_prop = (double*)(D + o_prop); m->d_lp = (double*)(D + o_lp); m->d_plp = (double*)(D + o_plp);
        m->d_log_scale = (double*)(D + o_ls); m->d_scale = (double*)(D + o_sc); m->d_best_lp = (double*)(D + o_blp); m->d_best_x = (double*)(D + o_bx);
        m->d_chol = (double*)(D + o_chol); m->d_lo = (double*)(D + o_lo); m->d_hi = (double*)(D + o_hi); m->d_init = (double*)(D + o_init);
        m->d_trace = (double*)(D + o_trace); m->d_mt = (unsigned*)(D + o_mt); m->d_C = (unsigned*)(D + o_C); m->d_T = (unsigned*)(D + o_T);
        m->d_recent = (unsigned*)(D + o_rec); m->d_status = (unsigned*)(D + o_st); m->d_recent_n = (int*)(D + o_rn); m->d_recent_sum = (int*)(D + o_rs);
        m->d_emergency = (int*)(D + o_em); m->d_accepted = (long long*)(D + o_acc); m->d_accepts = (unsigned char*)(D + o_accepts);
        e = cudaMemcpy(m->d_lo, sepaihrd_internal::lower_bounds(ctx), 8 * (size_t)d.P, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemset(m->d_status, 0, 4 * (L + 2));       // [L + 1] = the proposal kernel's fault flag
    }
    if (e == cudaSuccess) e = cudaMemcpy(m->d_hi, sepaihrd_internal::upper_bounds(ctx), 8 * (size_t)d.P, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        sepaihrd_mh_destroy(m);
        return fail_with(e == cudaErrorMemoryAllocation ? SEPAIHRD_ERR_OUT_OF_MEMORY : SEPAIHRD_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = m;
    return SEPAIHRD_OK;
}

void sepaihrd_mh_destroy(sepaihrd_mh* m) {
    if (!m) return;
    if (m->d_arena || m->d_win) cudaDeviceSynchronize();
    if (m->d_arena) cudaFree(m->d_arena);
    if (m->d_win) cudaFree(m->d_win);
    delete m;
}

sepaihrd_rc sepaihrd_mh_begin(sepaihrd_mh* m, uint32_t seed, const double* initial, const double* chol_lower) {
    if (!m || !initial || !chol_lower) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    const auto ctx_lock = sepaihrd_internal::lock(m->ctx);
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(m->ctx);
    MH_TRY(cudaSetDevice(d.device));
    cudaStream_t st = sepaihrd_internal::stream(m->ctx);
    const int P = m->P;
    m->diagonal = true;
    for (int j = 0; j < P && m->diagonal; ++j)
        for (int i = 0; i < P; ++i)
            if (i != j && chol_lower[(size_t)j * P + i] != 0.0) { m->diagonal = false; break; }
    MH_TRY(cudaStreamSynchronize(st));
    MH_TRY(cudaMemcpyAsync(m->d_chol, chol_lower, 8 * (size_t)P * P, cudaMemcpyHostToDevice, st));
    MH_TRY(cudaMemcpyAsync(m->d_init, initial, 8 * (size_t)P, cudaMemcpyHostToDevice, st));
    MH_TRY(cudaStreamSynchronize(st));                      // the sources are the caller's pageable buffers
    // log-posterior of the common start: ONE evaluation (like optimize(), host/optimizers.cpp), read by the start kernel
    sepaihrd_rc rc = sepaihrd_internal::eval_batch_device_unordered(m->ctx, m->d_init, 1, P, m->d_plp + m->local, m->d_status + m->local, nullptr);
    if (rc != SEPAIHRD_OK) return rc;
    if (m->local > 0) {
        const unsigned blocks = (unsigned)((m->local + 63) / 64);
        mh_seed_kernel<<<blocks, 64, 0, st>>>(m->local, m->offset, seed, m->d_mt, m->d_C, m->d_T);
        MH_TRY(cudaGetLastError());
        mh_start_kernel<<<blocks, 64, 0, st>>>(m->local, P, m->d_init, m->d_plp + m->local, m->d_x, m->d_lp, m->d_log_scale, m->d_scale,
                                                        m->d_best_lp, m->d_best_x, m->d_recent, m->d_recent_n, m->d_recent_sum, m->d_emergency,
                                                        m->d_accepted);
        MH_TRY(cudaGetLastError());
        sepaihrd_internal::count_launches(m->ctx, 2);
    }
    m->t = 1;
    m->windowed = false; m->win_K = 0;
    m->begun = true;
    return SEPAIHRD_OK;
}

// the three phases of one iteration, each enqueued on the ctx stream (sepaihrd_mh_iterate = the three in a loop; the separate
// entry points let a caller put CUDA events between them)
static sepaihrd_rc mh_phase(sepaihrd_mh* m, int phase) {
    if (!m) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (!m->begun) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "sampler phase before sepaihrd_mh_begin");
    if (m->windowed) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "this run advances in look-ahead windows (sepaihrd_mh_window_*)");
    if (m->t >= m->cfg.iterations) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "all iterations have run");
    const auto ctx_lock = sepaihrd_internal::lock(m->ctx);
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(m->ctx);
    MH_TRY(cudaSetDevice(d.device));
    cudaStream_t st = sepaihrd_internal::stream(m->ctx);
    const unsigned blocks = (unsigned)((m->local + MH_WARPS - 1) / MH_WARPS);      // one warp per chain
    if (phase == 0 && m->local > 0) {
        mh_propose_kernel<<<blocks, MH_THREADS, 0, st>>>(m->local, m->P, m->diagonal ? 1 : 0, sepaihrd_internal::constraint_mode(m->ctx), m->d_chol,
                                                          m->d_lo, m->d_hi, m->d_x, m->d_scale, m->d_mt, m->d_C, m->d_T, m->d_prop, m->d_status + m->local + 1);
        MH_TRY(cudaGetLastError());
        sepaihrd_internal::count_launches(m->ctx, 1);
    } else if (phase == 1 && m->local > 0) {
        return sepaihrd_internal::eval_batch_device_unordered(m->ctx, m->d_prop, m->local, m->P, m->d_plp, m->d_status, nullptr);
    } else if (phase == 2) {
        if (m->local > 0) {
            mh_accept_kernel<<<blocks, MH_THREADS, 0, st>>>(m->local, m->P, m->t, m->cfg.adapt_scale, m->cfg.target_acceptance_rate, m->d_plp, m->d_prop,
                                                             m->d_x, m->d_lp, m->d_mt, m->d_C, m->d_T, m->d_log_scale, m->d_scale, m->d_recent,
                                                             m->d_recent_n, m->d_recent_sum, m->d_emergency, m->d_accepted, m->d_best_lp, m->d_best_x,
                                                             m->cfg.record_accepts ? m->d_accepts + (size_t)(m->t - 1) * m->local : nullptr);
            MH_TRY(cudaGetLastError());
            sepaihrd_internal::count_launches(m->ctx, 1);
        }
        m->t += 1;
    }
    return SEPAIHRD_OK;
}
sepaihrd_rc sepaihrd_mh_propose(sepaihrd_mh* m) { return mh_phase(m, 0); }
sepaihrd_rc sepaihrd_mh_evaluate(sepaihrd_mh* m) { return mh_phase(m, 1); }
sepaihrd_rc sepaihrd_mh_accept(sepaihrd_mh* m) { return mh_phase(m, 2); }

sepaihrd_rc sepaihrd_mh_iterate(sepaihrd_mh* m, int32_t n_iterations) {
    if (!m) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (!m->begun) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "sepaihrd_mh_iterate before sepaihrd_mh_begin");
    for (int it = 0; it < n_iterations && m->t < m->cfg.iterations; ++it)
        for (int phase = 0; phase < 3; ++phase) {
            const sepaihrd_rc rc = mh_phase(m, phase);
            if (rc != SEPAIHRD_OK) return rc;
        }
    return SEPAIHRD_OK;
}


// ---- look-ahead windows: K iterations of every chain per likelihood launch --------------------------------------------------
static sepaihrd_rc window_buffers(sepaihrd_mh* m, int K) {
    if (K <= m->win_cap) return SEPAIHRD_OK;
    if (m->d_win) { cudaDeviceSynchronize(); cudaFree(m->d_win); m->d_win = nullptr; m->win_cap = 0; }
    const size_t L = (size_t)m->local, LK = L * (size_t)K;
    size_t bytes = 0;
    auto reserve = [&](size_t b) { const size_t at = bytes; bytes += (b + 255) & ~(size_t)255; return at; };
    const size_t o_prop = reserve(8 * LK * m->P), o_plp = reserve(8 * LK), o_logu = reserve(8 * LK), o_cz = reserve(4 * LK), o_st = reserve(4 * LK),
                 o_ghost = reserve(4 * L * MT_N), o_t = reserve(4 * L), o_tmin = reserve(4), o_rec = reserve(8 * (L + 2));
    MH_TRY(cudaMalloc((void**)&m->d_win, bytes));
    char* D = m->d_win;
    m->d_wprop = (double*)(D + o_prop); m->d_wplp = (double*)(D + o_plp); m->d_wlogu = (double*)(D + o_logu); m->d_wCz = (unsigned*)(D + o_cz);
    m->d_wstatus = (unsigned*)(D + o_st); m->d_ghost = (unsigned*)(D + o_ghost);
    m->d_t = (int*)(D + o_t); m->d_tmin = (int*)(D + o_tmin); m->d_record = (double*)(D + o_rec);
    m->win_cap = K;
    return SEPAIHRD_OK;
}

static sepaihrd_rc mh_window_phase(sepaihrd_mh* m, int phase, int K, int64_t record_stride) {
    if (!m) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (!m->begun) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "sampler phase before sepaihrd_mh_begin");
    if (!m->windowed && m->t != 1) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "look-ahead windows cannot follow one-iteration phases of the same run");
    const auto ctx_lock = sepaihrd_internal::lock(m->ctx);
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(m->ctx);
    MH_TRY(cudaSetDevice(d.device));
    cudaStream_t st = sepaihrd_internal::stream(m->ctx);
    const unsigned blocks = (unsigned)((m->local + MH_WARPS - 1) / MH_WARPS);      // one warp per chain
    if (phase == 0) {
        if (K < 1 || K > 64) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "a window holds between 1 and 64 iterations per chain");
        if (!m->windowed) {
            if (m->win_cap < K) { const sepaihrd_rc rc = window_buffers(m, K); if (rc != SEPAIHRD_OK) return rc; }
            if (m->local > 0) {
                std::vector<int> ones((size_t)m->local, 1);
                MH_TRY(cudaMemcpyAsync(m->d_t, ones.data(), 4 * (size_t)m->local, cudaMemcpyHostToDevice, st));
                MH_TRY(cudaStreamSynchronize(st));
            }
            m->windowed = true;
        } else if (K > m->win_cap) {
            return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "the first window of a run (or sepaihrd_mh_window_reserve) fixes the largest window length");
        }
        m->win_K = K;
        if (m->local > 0) {
            mh_window_propose_kernel<<<blocks, MH_THREADS, 0, st>>>(m->local, m->P, K, m->cfg.iterations, m->diagonal ? 1 : 0,
                                                                      sepaihrd_internal::constraint_mode(m->ctx), m->cfg.adapt_scale, m->cfg.target_acceptance_rate,
                                                                      m->d_chol, m->d_lo, m->d_hi, m->d_x, m->d_log_scale, m->d_scale, m->d_recent, m->d_recent_n,
                                                                      m->d_recent_sum, m->d_mt, m->d_C, m->d_T, m->d_t, m->d_ghost, m->d_wprop, m->d_wCz,
                                                                      m->d_wlogu, m->d_status + m->local + 1);
            MH_TRY(cudaGetLastError());
            sepaihrd_internal::count_launches(m->ctx, 1);
        }
        return SEPAIHRD_OK;
    }
    if (!m->windowed || m->win_K < 1) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "window phase without sepaihrd_mh_window_propose");
    if (phase == 1) {
        if (m->local == 0) return SEPAIHRD_OK;
        return sepaihrd_internal::eval_batch_device_unordered(m->ctx, m->d_wprop, m->local * m->win_K, m->P, m->d_wplp, m->d_wstatus, nullptr);
    }
    // commit
    if (record_stride != 0 && (record_stride < m->local || record_stride > m->local + 1))
        return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "record stride: local_count or local_count + 1 (the largest shard of the run)");
    MH_TRY(cudaMemsetAsync(m->d_tmin, 0x7f, 4, st));
    if (m->local > 0) {
        mh_window_commit_kernel<<<blocks, MH_THREADS, 0, st>>>(m->local, m->P, m->win_K, m->cfg.iterations, m->cfg.adapt_scale, m->cfg.target_acceptance_rate,
                                                                 m->d_wplp, m->d_wprop, m->d_wCz, m->d_wlogu, m->d_x, m->d_lp, m->d_mt, m->d_C, m->d_T,
                                                                 m->d_log_scale, m->d_scale, m->d_recent, m->d_recent_n, m->d_recent_sum, m->d_emergency,
                                                                 m->d_accepted, m->d_best_lp, m->d_best_x, m->cfg.record_accepts ? m->d_accepts : nullptr,
                                                                 m->d_t, m->d_tmin, m->d_record);
        MH_TRY(cudaGetLastError());
    }
    mh_window_publish_kernel<<<1, 1, 0, st>>>(m->d_tmin, m->cfg.iterations, m->local, m->d_record + (record_stride ? record_stride : m->local));
    MH_TRY(cudaGetLastError());
    sepaihrd_internal::count_launches(m->ctx, m->local > 0 ? 2 : 1);
    m->win_K = 0;
    return SEPAIHRD_OK;
}
sepaihrd_rc sepaihrd_mh_window_reserve(sepaihrd_mh* m, int32_t K) {
    if (!m) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (K < 1 || K > 64) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "a window holds between 1 and 64 iterations per chain");
    if (m->windowed) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "the window buffers of a run in progress cannot be replaced");
    const auto ctx_lock = sepaihrd_internal::lock(m->ctx);
    MH_TRY(cudaSetDevice(sepaihrd_internal::dims(m->ctx).device));
    return window_buffers(m, K);
}
sepaihrd_rc sepaihrd_mh_window_propose(sepaihrd_mh* m, int32_t K) { return mh_window_phase(m, 0, K, 0); }
sepaihrd_rc sepaihrd_mh_window_evaluate(sepaihrd_mh* m) { return mh_window_phase(m, 1, 0, 0); }
sepaihrd_rc sepaihrd_mh_window_commit(sepaihrd_mh* m, int64_t record_stride) { return mh_window_phase(m, 2, 0, record_stride); }

sepaihrd_rc sepaihrd_mh_window_record(sepaihrd_mh* m, const double** d_record) {
    if (!m || !d_record) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (!m->windowed) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "no window has been proposed yet");
    *d_record = m->d_record;
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_mh_window_progress(sepaihrd_mh* m, int32_t* out_min_iteration) {
    if (!m || !out_min_iteration) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (!m->windowed) { *out_min_iteration = m->t; return SEPAIHRD_OK; }
    const auto ctx_lock = sepaihrd_internal::lock(m->ctx);
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(m->ctx);
    MH_TRY(cudaSetDevice(d.device));
    cudaStream_t st = sepaihrd_internal::stream(m->ctx);
    int v = 0;
    MH_TRY(cudaMemcpyAsync(&v, m->d_tmin, 4, cudaMemcpyDeviceToHost, st));
    MH_TRY(cudaStreamSynchronize(st));
    *out_min_iteration = std::min(v, m->cfg.iterations);
    m->t = *out_min_iteration;
    return SEPAIHRD_OK;
}

int32_t sepaihrd_mh_iteration(const sepaihrd_mh* m) { return m ? m->t : 0; }

sepaihrd_rc sepaihrd_mh_logpost_device(sepaihrd_mh* m, const double** d_logpost) {
    if (!m || !d_logpost) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    *d_logpost = m->d_lp;
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_mh_note_gathered(sepaihrd_mh* m, const double* d_all_logpost, int32_t world, int64_t block_stride, int32_t trace_slot) {
    if (!m || !d_all_logpost) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (world < 1 || trace_slot < 0 || trace_slot >= m->trace_cap) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "bad world / trace slot");
    const auto ctx_lock = sepaihrd_internal::lock(m->ctx);
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(m->ctx);
    MH_TRY(cudaSetDevice(d.device));
    mh_trace_kernel<<<1, 256, 0, sepaihrd_internal::stream(m->ctx)>>>(d_all_logpost, world, block_stride, m->n_chains / world, m->n_chains % world,
                                                                        m->d_trace, trace_slot);
    MH_TRY(cudaGetLastError());
    sepaihrd_internal::count_launches(m->ctx, 1);
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_mh_read(sepaihrd_mh* m, int32_t what, void* out) {
    if (!m || !out) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    const auto ctx_lock = sepaihrd_internal::lock(m->ctx);
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(m->ctx);
    MH_TRY(cudaSetDevice(d.device));
    cudaStream_t st = sepaihrd_internal::stream(m->ctx);
    const void* src = nullptr;
    size_t bytes = 0;
    const size_t L = (size_t)m->local;
    switch (what) {
        case SEPAIHRD_MH_POSITIONS: src = m->d_x; bytes = 8 * L * m->P; break;
        case SEPAIHRD_MH_LOGPOST: src = m->d_lp; bytes = 8 * L; break;
        case SEPAIHRD_MH_SCALES: src = m->d_scale; bytes = 8 * L; break;
        case SEPAIHRD_MH_ACCEPTED_COUNTS: src = m->d_accepted; bytes = 8 * L; break;
        case SEPAIHRD_MH_BEST_LOGPOST: src = m->d_best_lp; bytes = 8 * L; break;
        case SEPAIHRD_MH_BEST_POSITIONS: src = m->d_best_x; bytes = 8 * L * m->P; break;
        case SEPAIHRD_MH_ACCEPT_MATRIX:
            if (!m->cfg.record_accepts) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "accept decisions were not recorded (settings.record_accepts)");
            src = m->d_accepts; bytes = L * (size_t)(m->t - 1); break;
        case SEPAIHRD_MH_TRACE: src = m->d_trace; bytes = 8 * (size_t)m->trace_cap; break;
        case SEPAIHRD_MH_PROPOSALS: src = m->d_prop; bytes = 8 * L * m->P; break;
        case SEPAIHRD_MH_FAULT: src = m->d_status + m->local + 1; bytes = 4; break;
        default: return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "bad sampler array selector");
    }
    if (bytes == 0) return SEPAIHRD_OK;
    MH_TRY(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, st));
    MH_TRY(cudaStreamSynchronize(st));
    return SEPAIHRD_OK;
}

}  // extern "C"
