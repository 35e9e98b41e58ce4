// sepaihrd_kernels.cuh -- fused SEPAIHRD Dopri5 + Poisson-likelihood kernel for sm_100a.
//
// One launch = B parameter sets -> B log-likelihoods (MODE_LL) or B trajectories (MODE_TRAJ).
//
// Mapping:
//   * a parameter set is owned by a LANE GROUP of NA lanes, one lane per age class; each lane keeps the
//     11 compartments of its age class, the FSAL derivative and the Dopri5 stages in REGISTERS;
//   * cross-lane traffic is the infectious pressure all-gather (NA shuffles per RHS) and the error-norm
//     max (log2 NA shuffles per step attempt); t, dt and the accept/reject decision are replicated and
//     bit-identical in the NA lanes of a group, so a group never diverges internally;
//   * problem constants (time grid, observed H/ICU/D series, contact matrix, schedule breakpoints,
//     bounds, name->slot table) are staged once per block into shared memory with one TMA bulk copy
//     (cp.async.bulk + mbarrier); each set's constrained model parameters live in a padded shared-memory
//     slot vector, the hot ones are hoisted into registers;
//   * the passive compartments R, D, CumH, CumICU never feed the right-hand side, so their stage
//     derivatives are folded into running sums as they are produced (same association order as
//     Boost's scale_sum, hence bit-identical) and never stored;
//   * the Poisson log-likelihood is accumulated as each output day is reached: only one double (+ status,
//     + optional step counts) per set reaches HBM.
//
// Reference semantics restated here (file:line relative to the reference repository):
//   RHS                 src/model/AgeSEPAIHRDModel.cpp:101-228
//   beta(t), kappa(t)   src/model/PiecewiseConstantParameterStrategy.cpp:37-74, src/model/PieceWiseConstantNPIStrategy.cpp:86-127
//   Dopri5 controller   Boost.Odeint controlled_runge_kutta<runge_kutta_dopri5>, integrate_times
//                       (call site src/sir_age_structured/solvers/Dopri5SolverStrategy.cpp:28-37)
//   constraints         src/model/parameters/SEPAIHRDParameterManager.cpp:302-347
//   initial state       src/model/objectives/SEPAIHRDObjectiveFunction.cpp:124-163
//   incidence + Poisson src/model/objectives/SEPAIHRDObjectiveFunction.cpp:191-225, 241-279
#pragma once

#include <cfloat>
#include <climits>
#include <cstdint>
#include <cuda_runtime.h>

#include "sepaihrd_b200.h"
#include "sepaihrd_constraints.cuh"

namespace sepaihrd {

constexpr int NCOMP = SEPAIHRD_NUM_COMPARTMENTS;  // 11
constexpr int NDYN = 7;                            // S E P A I H ICU feed the RHS
constexpr int NPAS = 4;                            // R D CumH CumICU do not

enum { MODE_LL = 0, MODE_TRAJ = 1, MODE_PPC = 2 };   // MODE_PPC: FAST instantiation that writes TRAJ_PPC_SERIES and nothing else
// internal trajectory selector (next to SEPAIHRD_TRAJ_FULL / _OBSERVED): the six posterior-predictive series on the output days
// t >= 0, draws fastest -- out[6][T][n][B]: daily hospitalisations, ICU admissions, deaths (first differences of CumH / CumICU / D
// clamped at 0, ResultAggregator.cpp:292-335) and their running sums (.cpp:337-351).  The quantile pass reads them column by column.
constexpr int TRAJ_PPC_SERIES = 2;

// Everything the kernel needs besides the staged blob; passed by value (constant bank).
struct KParams {
    const double* blob;        // device copy of the constants blob (16-byte aligned)
    int blob_bytes;            // multiple of 16
    // offsets (in doubles) of the arrays inside the blob
    int o_times, o_obs_h, o_obs_i, o_obs_d, o_pop, o_agefrac, o_invN, o_M, o_bp, o_base, o_init, o_lo, o_hi;
    int o_pslot, o_segb, o_segk;   // int32 arrays, offsets still in doubles
    int o_logtab;              // 128 x (1/c_i, log c_i): table of the FAST-mode logarithm
    int n, K, n_obs, runup_offset, nb, nk, nseg, P, nslots;
    int slot_stride;           // doubles per set in the shared slot table (odd -> conflict-free)
    int seg_stride;            // doubles per set in the shared beta_eff table
    int constraint_mode;       // 0 clamp, 1 reflect
    double abs_tol, rel_tol, dt_hint;
    double hmax;               // longest output interval: growing dt beyond it cannot change the result
    double inv_rel;            // 1 / rel_tol          (FAST error norm is evaluated in units of rel_tol)
    double abs_over_rel;       // abs_tol / rel_tol
    double hc[32];             // Dopri5 coefficients pre-multiplied by hmax (b, c rows) resp. hmax / rel_tol (dc row): the attempt
                               // body of a full-length step reads them straight from the constant bank (T_* indices)
    double grow_max;           // 0.9 * pow(pow(5,-5), -1/5): the step-growth factor once err <= 5^-5 (libm, host)
    // LOOP 6: coarse classification of the error norm from the high words of num/den (units: 2^-20 of a log2)
    int thr_small, thr_nogrow, thr_big;
    // I/O
    const double* params;      // [B][ld]
    long long B, ld;
    double* out_ll;            // [B]
    unsigned* out_status;      // [B] or null
    int* out_steps;            // [B][2] or null
    double* out_traj;          // MODE_TRAJ: [B][traj_rows][W], or [traj_rows][W][B] when traj_draw_minor
    int traj_what, traj_stride, traj_rows;
    int traj_draw_minor;       // draws fastest: the layout the posterior-predictive quantile pass reads column by column
    long long ppc_b0, ppc_B;   // TRAJ_PPC_SERIES: this launch holds draws [ppc_b0, ppc_b0 + B) of ppc_B (chunked launches overlap the H2D copy)
    const double* init_states; // optional [B][11n] (or one shared state when init_stride == 0): Simulator::run semantics
    long long init_stride;
    const int* perm;           // optional [B]: the order in which the sets are handed to the warps (position -> set index); results still go
                               // to the set's own slot.  Built by the ordering pass (sepaihrd_order.cu) so that a warp holds sets that
                               // need similar numbers of step attempts per day
    int* out_profile;          // PROFILE instantiation only: [B][K] attempts (accepted + rejected) made before each grid point
    long long tiles;           // ceil(B / sets_per_tile)
    int sets_per_tile;         // sets a warp takes at a time: 32 / NA, fewer (down to one) when the launch is smaller than the machine
    int active_warps;          // warps per block that take tiles (all of them unless the launch is smaller than the machine)
    unsigned* tile_counter;    // zeroed before every launch (tiles < 2^32 - grid warps)
};

// ------------------------------------------------------------------------------------------------
// Arithmetic flavours.  STRICT: unfused IEEE ops (never contracted by nvcc) in the reference's source
// order.  FAST: explicit FMAs.
template <bool STRICT>
struct Ops;
template <>
struct Ops<true> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mad(double a, double b, double c) { return __dadd_rn(c, __dmul_rn(a, b)); }
};
template <>
struct Ops<false> {
    static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
    static __device__ __forceinline__ double add(double a, double b) { return a + b; }
    static __device__ __forceinline__ double sub(double a, double b) { return a - b; }
    static __device__ __forceinline__ double mad(double a, double b, double c) { return fma(a, b, c); }
};

// ---- TMA bulk copy + mbarrier (PTX) ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- step-size controller powers (FAST) ---------------------------------------------------------------
// x^(-1/3) (cubic) or x^(-1/5) by an FP32 MUFU seed (rel. error ~1e-6) and two Newton steps in FP64 (error constant 2
// resp. 3: 1e-6 -> ~3e-12 -> rounding level, a few ulp; ~14 FP64 instructions against ~130 for pow()), in ONE
// instruction stream: lanes that shrink a rejected step and lanes that grow an accepted one share the chain.
__device__ __forceinline__ double pow_neg_inv(double x, bool cubic) {
    double y = (double)__powf((float)x, cubic ? -0.33333334f : -0.2f);
    const double xs = x * (cubic ? (1.0 / 3.0) : 0.2);
    const double cst = cubic ? (4.0 / 3.0) : 1.2;
#ifndef SEPAIHRD_POW_NEWTON_STEPS
#define SEPAIHRD_POW_NEWTON_STEPS 2      // one step (3e-12) was measured in round 2: -0.5 % time and still 0 of 2,097,152 sets with other step counts,
#endif                                   // but the worst logL error of the batch grows from 8.9e-10 to 3.1e-9 (gate 1e-8): not worth the margin
#pragma unroll
    for (int it = 0; it < SEPAIHRD_POW_NEWTON_STEPS; ++it) {
        const double y2 = y * y;
        const double yn = (cubic ? y2 : (y2 * y2)) * y;
        y = y * fma(-xs, yn, cst);
    }
    return y;
}

// ---- logarithm for the Poisson terms (FAST) -----------------------------------------------------------
// log(x) = e ln2 + log(c_i) + log1p(r),  r = m / c_i - 1 (one FMA, exact rounding), i = top 7 mantissa bits,
// |r| <= 2^-8, log1p by a degree-5 polynomial (truncation 7e-16).  Absolute error ~1e-15: far inside what the
// 1e-8 relative gate on logL needs, at ~9 FP64 instructions instead of ~35 for CUDA's log() (v4 profile: the three
// log() calls per output day were 25% of the kernel time).
// No special-case branch, so that the three Poisson streams of an output day interleave in one basic block.  Inputs
// outside the positive normal range (the incidence went NaN/inf) raise `bad`; the caller turns that into the
// reference's non-finite sentinel.
__device__ __forceinline__ double fast_log_nb(double x, const double2* __restrict__ tab, bool valid, bool& bad) {
    const int hi = __double2hiint(x);
    bad |= valid && ((unsigned)(hi - 0x00100000) >= 0x7fe00000u);
    const int e = (hi >> 20) - 1023;
    const int i = (hi >> 13) & 0x7f;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));   // [1, 2)
    const double2 t = tab[i];
    const double r = fma(m, t.x, -1.0);
    double p = fma(r, 0.2, -0.25);
    p = fma(p, r, 1.0 / 3.0);
    p = fma(p, r, -0.5);
    p = fma(p, r, 1.0);
    return fma((double)e, 0.6931471805599453, fma(p, r, t.y));
}

// a / b for a normal, positive b: MUFU.RCP64H seed (20 bits), two Newton steps, one residual correction.  Replaces
// the IEEE division (slow-path check + call) on the error-norm VALUE, which only sizes the next step.
__device__ __forceinline__ double fast_div_pos(double a, double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}

// OR of every lane group's bits, replicated to all lanes of the group (ballot mask -> group mask).
template <int NA>
__device__ __forceinline__ unsigned expand_groups(unsigned m) {
    if (NA >= 32) return m ? 0xffffffffu : 0u;
    unsigned y = m;
#pragma unroll
    for (int s = 1; s < NA; s *= 2) y |= y >> s;
    unsigned first = 0;
#pragma unroll
    for (int g = 0; g < 32; g += NA) first |= 1u << g;
    return (y & first) * ((NA >= 32) ? 1u : ((1u << NA) - 1u));
}

// ---- per-lane model parameters (registers) ----------------------------------------------------------
template <int NA>
struct LaneParams {
    double theta, sigma, gamma_p, gamma_A, gamma_I, gamma_H, gamma_ICU;
    double a, hinf, invN, p, h, icu, dH, dICU, dcomm;
    double hN, kI, kH, kU;   // FAST: h_infec/N, gamma_I+h+d_comm, gamma_H+d_H+icu, gamma_ICU+d_ICU
    double M[NA];            // row `age` of the contact matrix: M(age, j)
};

// Infectious-pressure all-gather inside a lane group through shared memory: one STS.64 per lane, one
// __syncwarp, then NA/2 LDS.128 (v2 profile: the 2*NA serialized SHFLs cost as much issue time as ~40
// instructions per RHS).  `slot` alternates between two buffers so that consecutive exchanges never
// race (write-after-read) without a second barrier.
template <int NA>
__device__ __forceinline__ void gather_pressure(double* spi, int slot_base, int lane_in_block, double pressure,
                                                double (&pall)[NA]) {
    spi[slot_base + lane_in_block] = pressure;
    __syncwarp();
    const double2* rp = reinterpret_cast<const double2*>(spi + slot_base + (lane_in_block & ~(NA - 1)));
#pragma unroll
    for (int j = 0; j < NA / 2; ++j) {
        const double2 v = rp[j];
        pall[2 * j] = v.x;
        pall[2 * j + 1] = v.y;
    }
}

// AgeSEPAIHRDModel::computeDerivatives for one age class (this lane), inputs y = S E P A I H ICU.
// Outputs: dyn = d(S E P A I H ICU), pas = d(R D CumH CumICU).
// PASSIVE=false skips pas (stage 2: Dopri5 has c2 = dc2 = 0).
// FAST: q.M is the FOLDED row M(age, j) * h_infec_j / N_j * (beta*kappa*a_age) of the schedule segment the step
// runs in, so the lanes exchange u_j = P_j + A_j + theta I_j and lambda is one short dot product (two partial sums).
// FOLDED=false (a step whose stages sit in different segments; never on the Spain-2020 grid) takes the unfolded
// row M(age, j) h_infec_j / N_j from shared memory (`mb`, thread-strided) and multiplies by `ba`.
// With 16 or more age classes the folded row lives in shared memory too (`mb` then points at it, same stride): 2 NA
// registers per lane that the stage vectors need more (row_in_smem()).
// Measured at 16 ages (profiles/r02_v17_16_ages_contact_row_in_smem.txt): it removes the spills everywhere, but only the
// posterior-predictive instantiation gains (-10 %: its observer keeps six more doubles live); the others lose 5-8 % to the
// 16 extra LDS.64 per right-hand side, so only MODE_PPC takes it.
template <int NA, bool STRICT, int MODE>
__host__ __device__ constexpr bool row_in_smem() { return NA >= 16 && !STRICT && MODE == 2; }
template <int NA, bool STRICT, bool PASSIVE, bool FOLDED = true, bool MSM = false>
__device__ __forceinline__ void rhs(const LaneParams<NA>& q, double ba, double* spi, int slot_base, int lane_in_block,
                                    const double (&y)[NDYN], double (&dyn)[NDYN], double (&pas)[NPAS],
                                    const double* mb = nullptr, int mb_stride = 0) {
    using O = Ops<STRICT>;
    const double S = y[0], E = y[1], P = y[2], A = y[3], I = y[4], H = y[5], U = y[6];
    double pressure;
    if (STRICT) {
        const double total_inf = O::add(O::add(P, A), O::mul(q.theta, I));   // :155
        pressure = O::mul(O::mul(total_inf, q.hinf), q.invN);                // :156
    } else {
        pressure = fma(q.theta, I, P + A);
    }
    double pall[NA];
    gather_pressure<NA>(spi, slot_base, lane_in_block, pressure, pall);
    double lam;
    if (STRICT) {
        lam = 0.0;                                                           // :162-174 (j outer, in order)
#pragma unroll
        for (int j = 0; j < NA; ++j) lam = O::mad(q.M[j], pall[j], lam);
        lam = O::mul(lam, ba);                                               // :181-183, ba = (beta*kappa)*a_i
    } else {
        constexpr int NACC = (NA >= 16) ? 4 : 2;
        double acc[NACC];
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            const double m = (FOLDED && !MSM) ? q.M[j] : mb[j * mb_stride];
            acc[j % NACC] = (j < NACC) ? m * pall[j] : fma(m, pall[j], acc[j % NACC]);
        }
        lam = (NACC == 4) ? (acc[0] + acc[1]) + (acc[2] + acc[3]) : acc[0] + acc[1];
        if (!FOLDED) lam *= ba;
    }
    if (STRICT) {
        lam = (0.0 < lam) ? lam : 0.0;                                       // std::max(0.0, lambda) :196
    } else {
        // same selection with integer ops (off the FP64 pipe): negative, -0 and NaN (sign or quiet bit patterns
        // above +inf) become +0; +inf and positive finite values pass.
        const unsigned hi = (unsigned)__double2hiint(lam);
        if (hi > 0x7ff00000u) lam = 0.0;
    }
    const double flow_SE = O::mul(lam, S);
    const double flow_IH = O::mul(q.h, I);
    const double flow_H_ICU = O::mul(q.icu, H);
    dyn[0] = -flow_SE;
    if (STRICT) {
        const double flow_EP = O::mul(q.sigma, E);
        const double flow_P_out = O::mul(q.gamma_p, P);
        const double flow_PA = O::mul(q.p, flow_P_out);
        const double flow_PI = O::sub(flow_P_out, flow_PA);
        const double flow_IR = O::mul(q.gamma_I, I);
        const double flow_IDc = O::mul(q.dcomm, I);
        const double I_out = O::add(O::add(flow_IR, flow_IH), flow_IDc);
        const double gH = O::mul(q.gamma_H, H), dHH = O::mul(q.dH, H);
        const double H_out = O::add(O::add(gH, dHH), flow_H_ICU);
        const double U_out = O::mul(O::add(q.gamma_ICU, q.dICU), U);
        const double gAA = O::mul(q.gamma_A, A);
        dyn[1] = O::sub(flow_SE, flow_EP);
        dyn[2] = O::sub(flow_EP, flow_P_out);
        dyn[3] = O::sub(flow_PA, gAA);
        dyn[4] = O::sub(flow_PI, I_out);
        dyn[5] = O::sub(flow_IH, H_out);
        dyn[6] = O::sub(flow_H_ICU, U_out);
        if (PASSIVE) {
            const double gUU = O::mul(q.gamma_ICU, U), dUU = O::mul(q.dICU, U);
            pas[0] = O::add(O::add(O::add(gAA, flow_IR), gH), gUU);          // :222
            pas[1] = O::add(O::add(dHH, dUU), flow_IDc);                    // :223
            pas[2] = flow_IH;
            pas[3] = flow_H_ICU;
        }
    } else {
        const double flow_P_out = q.gamma_p * P;
        const double flow_PA = q.p * flow_P_out;
        dyn[1] = fma(-q.sigma, E, flow_SE);
        dyn[2] = fma(q.sigma, E, -flow_P_out);
        dyn[3] = fma(-q.gamma_A, A, flow_PA);
        dyn[4] = fma(-q.kI, I, flow_P_out - flow_PA);
        dyn[5] = fma(-q.kH, H, flow_IH);
        dyn[6] = fma(-q.kU, U, flow_H_ICU);
        if (PASSIVE) {
            pas[0] = fma(q.gamma_ICU, U, fma(q.gamma_H, H, fma(q.gamma_I, I, q.gamma_A * A)));
            pas[1] = fma(q.dcomm, I, fma(q.dICU, U, q.dH * H));
            pas[2] = flow_IH;
            pas[3] = flow_H_ICU;
        }
    }
}

// Dopri5 tableau as Boost writes it (runge_kutta_dopri5.hpp): integer ratios evaluated in double.
// Kept in __constant__ memory so that every use is a c[bank][offset] operand of the DMUL/DFMA itself
// (as 64-bit immediates each use costs two UMOVs; v1 profile: 10% of all issued instructions).
enum TabIdx {
    T_A2, T_A3, T_A4, T_A5,
    T_B21, T_B31, T_B32, T_B41, T_B42, T_B43, T_B51, T_B52, T_B53, T_B54, T_B61, T_B62, T_B63, T_B64, T_B65,
    T_C1, T_C3, T_C4, T_C5, T_C6, T_DC1, T_DC3, T_DC4, T_DC5, T_DC6, T_DC7, T_COUNT
};
namespace tabv {
constexpr double c1 = 35.0 / 384.0, c3 = 500.0 / 1113.0, c4 = 125.0 / 192.0, c5 = -2187.0 / 6784.0, c6 = 11.0 / 84.0;
}
#define SEPAIHRD_TABLEAU_VALUES { \
    1.0 / 5.0, 3.0 / 10.0, 4.0 / 5.0, 8.0 / 9.0, \
    1.0 / 5.0, 3.0 / 40.0, 9.0 / 40.0, 44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0, \
    19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0, \
    9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0, \
    tabv::c1, tabv::c3, tabv::c4, tabv::c5, tabv::c6, \
    tabv::c1 - 5179.0 / 57600.0, tabv::c3 - 7571.0 / 16695.0, tabv::c4 - 393.0 / 640.0, \
    tabv::c5 - -92097.0 / 339200.0, tabv::c6 - 187.0 / 2100.0, -1.0 / 40.0}
__constant__ double c_tab[T_COUNT] = SEPAIHRD_TABLEAU_VALUES;
static const double h_tab[T_COUNT] = SEPAIHRD_TABLEAU_VALUES;   // host image of c_tab (same constant expressions)

constexpr unsigned FULL = 0xffffffffu;

template <int NA>
__device__ __forceinline__ double group_max(double v) {
#pragma unroll
    for (int off = NA / 2; off >= 1; off >>= 1) {
        const double o = __shfl_xor_sync(FULL, v, off, NA);
        v = (v < o) ? o : v;
    }
    return v;
}
template <int NA>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int off = NA / 2; off >= 1; off >>= 1) v += __shfl_xor_sync(FULL, v, off, NA);
    return v;
}

// One Dopri5 step attempt (runge_kutta_dopri5::do_step_impl with error estimate) for this lane's age class.
// MIXED=false: every stage of the step lies in one schedule segment (always true when the breakpoints sit
// on output-grid points, e.g. the Spain-2020 configuration) and uses `ba_step`.  MIXED=true: the step
// straddles a breakpoint; each stage looks its segment up from its own time.
struct StepSched {
    double ba_step;       // (beta*kappa)*a_i of the segment of the first stage time
    int s_lo;             // segment of t + dt/5
    const double* bp;     // merged breakpoints (shared memory)
    const double* beff;   // this set's beta*kappa per segment (shared memory)
    int nseg;
    double a;
    const double* mb;     // FAST: this lane's unfolded row M(age, j) h_infec_j / N_j in shared memory, stride mb_stride
    const double* mf;     // FAST, row_in_smem(): this lane's folded row, same stride
    int mb_stride;
};

// UNIT=true: every stepping lane group of the warp takes a step of exactly hmax (the usual case on a uniform output
// grid: 48 % of the warp-attempts of the Spain-2020 jitter batch), so the 27 products step x coefficient are the
// launch constants `hc` instead of 27 DMULs per lane.  hc[i] is computed on the host with the same single rounding.
template <int NA, bool STRICT, bool MIXED, bool UNIT = false, bool MSM = false>
__device__ __forceinline__ void dopri5_attempt(const LaneParams<NA>& q, const StepSched& sc, double* spi, int& pi_slot,
                                               int pi_stride, int lane_in_block, double t, double cur, double t_end,
                                               const double (&x)[NCOMP], const double (&k1)[NCOMP], double (&xn)[NDYN],
                                               double (&k7d)[NDYN], double (&k7p)[NPAS], double (&accN)[NPAS],
                                               double (&xe)[NCOMP], double ecur, const double* hc = nullptr) {
    using O = Ops<STRICT>;
    static_assert(!(UNIT && MIXED), "a full-length step on a breakpoint-free day never mixes segments");
    auto cf = [&](int i) -> double { return UNIT ? hc[i] : O::mul(cur, c_tab[i]); };     // step * b_ij, step * c_j
    auto ef = [&](int i) -> double { return UNIT ? hc[i] : O::mul(ecur, c_tab[i]); };    // step / rel_tol * dc_j
    auto ba_at = [&](int tab_a) -> double {
        if (!MIXED) return sc.ba_step;
        const double ts = (tab_a < 0) ? t_end : O::add(t, O::mul(cur, c_tab[tab_a]));
        int s = sc.s_lo;
        while (s < sc.nseg && ts > sc.bp[s]) ++s;
        return O::mul(sc.beff[s], sc.a);
    };
    auto next_slot = [&]() -> int { pi_slot ^= pi_stride; return pi_slot; };
    const double* mrow = MIXED ? sc.mb : sc.mf;
    double k2[NDYN], k3[NDYN], k4[NDYN], k5[NDYN], k6[NDYN];
    double y[NDYN], kp_[NPAS], accE[NPAS];
    // stage 2
    { const double f1 = cf(T_B21);
#pragma unroll
      for (int c = 0; c < NDYN; ++c) y[c] = O::mad(f1, k1[c], x[c]); }
    rhs<NA, STRICT, false, !MIXED, MSM>(q, ba_at(T_A2), spi, next_slot(), lane_in_block, y, k2, kp_, mrow, sc.mb_stride);
    // stage 3
    { const double f1 = cf(T_B31), f2 = cf(T_B32);
#pragma unroll
      for (int c = 0; c < NDYN; ++c) y[c] = O::mad(f2, k2[c], O::mad(f1, k1[c], x[c])); }
    rhs<NA, STRICT, true, !MIXED, MSM>(q, ba_at(T_A3), spi, next_slot(), lane_in_block, y, k3, kp_, mrow, sc.mb_stride);
    { const double g1 = cf(T_C1), g3 = cf(T_C3);
      const double e1 = ef(T_DC1), e3 = ef(T_DC3);
#pragma unroll
      for (int c = 0; c < NPAS; ++c) {
          accN[c] = O::mad(g3, kp_[c], O::mad(g1, k1[NDYN + c], x[NDYN + c]));
          accE[c] = O::mad(e3, kp_[c], O::mul(e1, k1[NDYN + c]));
      } }
    // stage 4
    { const double f1 = cf(T_B41), f2 = cf(T_B42), f3 = cf(T_B43);
#pragma unroll
      for (int c = 0; c < NDYN; ++c) y[c] = O::mad(f3, k3[c], O::mad(f2, k2[c], O::mad(f1, k1[c], x[c]))); }
    rhs<NA, STRICT, true, !MIXED, MSM>(q, ba_at(T_A4), spi, next_slot(), lane_in_block, y, k4, kp_, mrow, sc.mb_stride);
    { const double g4 = cf(T_C4), e4 = ef(T_DC4);
#pragma unroll
      for (int c = 0; c < NPAS; ++c) { accN[c] = O::mad(g4, kp_[c], accN[c]); accE[c] = O::mad(e4, kp_[c], accE[c]); } }
    // stage 5
    { const double f1 = cf(T_B51), f2 = cf(T_B52), f3 = cf(T_B53),
                   f4 = cf(T_B54);
#pragma unroll
      for (int c = 0; c < NDYN; ++c)
          y[c] = O::mad(f4, k4[c], O::mad(f3, k3[c], O::mad(f2, k2[c], O::mad(f1, k1[c], x[c])))); }
    rhs<NA, STRICT, true, !MIXED, MSM>(q, ba_at(T_A5), spi, next_slot(), lane_in_block, y, k5, kp_, mrow, sc.mb_stride);
    { const double g5 = cf(T_C5), e5 = ef(T_DC5);
#pragma unroll
      for (int c = 0; c < NPAS; ++c) { accN[c] = O::mad(g5, kp_[c], accN[c]); accE[c] = O::mad(e5, kp_[c], accE[c]); } }
    // stage 6
    { const double f1 = cf(T_B61), f2 = cf(T_B62), f3 = cf(T_B63),
                   f4 = cf(T_B64), f5 = cf(T_B65);
#pragma unroll
      for (int c = 0; c < NDYN; ++c)
          y[c] = O::mad(f5, k5[c], O::mad(f4, k4[c], O::mad(f3, k3[c], O::mad(f2, k2[c], O::mad(f1, k1[c], x[c]))))); }
    rhs<NA, STRICT, true, !MIXED, MSM>(q, ba_at(-1), spi, next_slot(), lane_in_block, y, k6, kp_, mrow, sc.mb_stride);
    { const double g6 = cf(T_C6), e6 = ef(T_DC6);
#pragma unroll
      for (int c = 0; c < NPAS; ++c) { accN[c] = O::mad(g6, kp_[c], accN[c]); accE[c] = O::mad(e6, kp_[c], accE[c]); } }
    // solution (dynamic part) and the FSAL derivative
    { const double g1 = cf(T_C1), g3 = cf(T_C3), g4 = cf(T_C4),
                   g5 = cf(T_C5), g6 = cf(T_C6);
#pragma unroll
      for (int c = 0; c < NDYN; ++c)
          xn[c] = O::mad(g6, k6[c], O::mad(g5, k5[c], O::mad(g4, k4[c], O::mad(g3, k3[c], O::mad(g1, k1[c], x[c]))))); }
    rhs<NA, STRICT, true, !MIXED, MSM>(q, ba_at(-1), spi, next_slot(), lane_in_block, xn, k7d, k7p, mrow, sc.mb_stride);
    // error estimate
    { const double e1 = ef(T_DC1), e3 = ef(T_DC3), e4 = ef(T_DC4),
                   e5 = ef(T_DC5), e6 = ef(T_DC6), e7 = ef(T_DC7);
#pragma unroll
      for (int c = 0; c < NDYN; ++c)
          xe[c] = O::mad(e7, k7d[c], O::mad(e6, k6[c], O::mad(e5, k5[c], O::mad(e4, k4[c], O::mad(e3, k3[c], O::mul(e1, k1[c]))))));
#pragma unroll
      for (int c = 0; c < NPAS; ++c) xe[NDYN + c] = O::mad(e7, k7p[c], accE[c]); }
}

// ------------------------------------------------------------------------------------------------
// The kernel is WARP-SYNCHRONOUS: all 32 lanes of a warp execute every step attempt together (lane
// groups that have already reached the next output time run the attempt with a zero-length step and
// discard it), so every shuffle is a plain full-mask SHFL.  This is the "regroup at output-day
// boundaries" policy: a warp spends max-over-its-groups attempts per output interval.
// ONGRID (FAST only): every schedule breakpoint inside the integration window sits on an output-grid point, so no step can
// have its stages in two segments and the mixed-segment attempt body is not instantiated (668 SASS instructions less to
// keep in the instruction cache).  The host decides per problem (sepaihrd_create).
// PROFILE (FAST, log-likelihood mode): additionally records the running attempt count at every grid point -- the pilot of the
// ordering pass; a separate instantiation so that the production kernel carries none of it.
template <int NA, bool STRICT, int MODE, int THREADS, int MINBLOCKS, int LOOP, bool ONGRID = false, bool PROFILE = false>
__global__ void __launch_bounds__(THREADS, MINBLOCKS) sepaihrd_batch_kernel(const KParams kp) {
    static_assert(STRICT ? LOOP == 5 : LOOP == 6, "STRICT keeps the reference-order loop 5; FAST runs loop 6");
    static_assert(!(STRICT && MODE == MODE_PPC), "STRICT writes the posterior-predictive series from its MODE_TRAJ instantiation");
    using O = Ops<STRICT>;
    constexpr int SETS = THREADS / NA;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sblob = reinterpret_cast<double*>(smem_raw);
    const int blob_doubles = kp.blob_bytes >> 3;
    double* sslots = sblob + blob_doubles;
    double* sbeff = sslots + SETS * kp.slot_stride;
    double* spi = sbeff + SETS * ((kp.seg_stride + 1) & ~1);            // 2 x THREADS doubles, 16-byte aligned
    double* smb = spi + 2 * THREADS;                                    // NA x THREADS doubles: unfolded contact rows (FAST)
    constexpr bool MSM = row_in_smem<NA, STRICT, MODE>();                     // the same array then holds the row the attempt reads:
    double* smf = smb;                                                  // folded, or unfolded around a mixed-segment attempt
    uint64_t* bar = reinterpret_cast<uint64_t*>(smb + NA * THREADS);

    // ---- stage the constants blob once per block with one TMA bulk copy -----------------------------
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (uint32_t)kp.blob_bytes);
        tma_bulk_g2s(sblob, kp.blob, (uint32_t)kp.blob_bytes, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);

    const double* s_times = sblob + kp.o_times;
    const double* s_obs_h = sblob + kp.o_obs_h;
    const double* s_obs_i = sblob + kp.o_obs_i;
    const double* s_obs_d = sblob + kp.o_obs_d;
    const double* s_bp = sblob + kp.o_bp;
    const int* s_pslot = reinterpret_cast<const int*>(sblob + kp.o_pslot);
    const int* s_segb = reinterpret_cast<const int*>(sblob + kp.o_segb);
    const int* s_segk = reinterpret_cast<const int*>(sblob + kp.o_segk);
    const double2* s_logtab = reinterpret_cast<const double2*>(sblob + kp.o_logtab);
    int pi_slot = 0;   // toggles between 0 and THREADS before every pressure exchange

    const int age = threadIdx.x % NA;
    const int grp = threadIdx.x / NA;
    constexpr int n = NA;
    const int nseg = kp.nseg;
    const int K = kp.K;
    const double hmax = kp.hmax;
    const double f_abs = kp.abs_tol, f_rel = kp.rel_tol;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);

    double* my_slots = sslots + grp * kp.slot_stride;
    double* my_beff = sbeff + grp * kp.seg_stride;

    // slot layout (sepaihrd_b200.h)
    const int sl_beta0 = 0, sl_kappa0 = kp.nb, sl_scal0 = kp.nb + kp.nk, sl_age0 = sl_scal0 + 7;
    const int sl_mult0 = sl_age0 + 8 * n, sl_seed = sl_mult0 + 8, sl_runup = sl_mult0 + 9, sl_beta = sl_mult0 + 10;

    // Work is handed out per WARP (kp.sets_per_tile <= 32 / NA sets at a time) from a global counter: warps never wait for the
    // other warps of their block and the tail of a launch is one warp-tile long.
    const int grp_in_warp = (threadIdx.x & 31) / NA;
    // A launch with fewer tiles than resident warps (a few thousand sets: multi-chain samplers, line searches) is spread over the
    // SMs and, inside an SM, over the schedulers: the host launches one block per tile until the machine is full and lets only
    // the first `active_warps` warps of every block work (64 tiles: one warp on each of 64 SMs, not eight warps on eight SMs) --
    // its attempts are latency-bound and want a scheduler of their own.
    if ((int)(threadIdx.x >> 5) >= kp.active_warps) return;
    while (true) {
        unsigned wt = 0;
        if ((threadIdx.x & 31) == 0) wt = atomicAdd(kp.tile_counter, 1u);
        wt = __shfl_sync(FULL, wt, 0);
        if ((long long)wt >= kp.tiles) break;
        // a small launch hands out fewer than WSETS sets per tile (a warp then pays for its own sets' attempts only, not for the
        // day-by-day maximum over eight); idle groups shadow the tile's first set, nothing is written for them
        const long long b_first = (long long)wt * kp.sets_per_tile;
        const long long b_raw = b_first + grp_in_warp;
        const bool have = grp_in_warp < kp.sets_per_tile && b_raw < kp.B;
        const long long b_pos = have ? b_raw : b_first;
        const long long b = kp.perm ? (long long)kp.perm[b_pos] : b_pos;

        // ---- updateModelParameters: base slots, then constrained calibrated values -------------------
        __syncwarp();
        for (int s = age; s < kp.nslots; s += NA) my_slots[s] = sblob[kp.o_base + s];
        __syncwarp();
        bool kappa_touched = false;
        {
            const double* prow = kp.params + b * kp.ld;
            for (int i = age; i < kp.P; i += NA) {
                const int sl = s_pslot[i];
                if (sl >= 0) {
                    const double v = constrain(prow[i], sblob[kp.o_lo + i], sblob[kp.o_hi + i], kp.constraint_mode);
                    my_slots[sl] = v;
                    if (sl >= sl_kappa0 && sl < sl_kappa0 + kp.nk) kappa_touched = true;
                }
            }
        }
        __syncwarp();
        unsigned status = 0;
        {   // setCalibratableValues throws on a negative kappa (NPI.cpp:238-242) once any kappa is calibrated
            bool neg = false;
            for (int k = 1 + age; k < kp.nk; k += NA) neg |= (my_slots[sl_kappa0 + k] < 0.0);
            const unsigned gsh = (threadIdx.x & 31) - age;
            const unsigned gm = (NA >= 32) ? FULL : (((1u << NA) - 1u) << gsh);
            const bool any_touch = (__ballot_sync(FULL, kappa_touched) & gm) != 0;
            const bool any_neg = (__ballot_sync(FULL, neg) & gm) != 0;
            if (any_touch && any_neg) status |= SEPAIHRD_ST_INVALID_PARAM;
        }
        // beta_eff per merged schedule segment: beta(t) * kappa(t)   (AgeSEPAIHRDModel.cpp:176-178)
        for (int s = age; s <= nseg; s += NA) {
            const double bv = (kp.nb > 0) ? my_slots[sl_beta0 + s_segb[s]] : my_slots[sl_beta];
            my_beff[s] = O::mul(bv, my_slots[sl_kappa0 + s_segk[s]]);
        }
        __syncwarp();

        LaneParams<NA> q;
        q.theta = my_slots[sl_scal0 + 0]; q.sigma = my_slots[sl_scal0 + 1]; q.gamma_p = my_slots[sl_scal0 + 2];
        q.gamma_A = my_slots[sl_scal0 + 3]; q.gamma_I = my_slots[sl_scal0 + 4]; q.gamma_H = my_slots[sl_scal0 + 5];
        q.gamma_ICU = my_slots[sl_scal0 + 6];
        q.a = my_slots[sl_age0 + 0 * n + age]; q.hinf = my_slots[sl_age0 + 1 * n + age];
        q.p = my_slots[sl_age0 + 2 * n + age]; q.h = my_slots[sl_age0 + 3 * n + age];
        q.icu = my_slots[sl_age0 + 4 * n + age]; q.dH = my_slots[sl_age0 + 5 * n + age];
        q.dICU = my_slots[sl_age0 + 6 * n + age]; q.dcomm = my_slots[sl_age0 + 7 * n + age];
        q.invN = sblob[kp.o_invN + age];
        q.hN = q.hinf * q.invN; q.kI = q.gamma_I + q.h + q.dcomm; q.kH = q.gamma_H + q.dH + q.icu;
        q.kU = q.gamma_ICU + q.dICU;
#pragma unroll
        for (int j = 0; j < NA; ++j) q.M[j] = sblob[kp.o_M + j * n + age];   // column-major M(age, j)
        int mseg = -1;                 // FAST: schedule segment q.M is currently folded for
        // unfolded row element M(age, j) h_infec_j / N_j: kept in shared memory, or (MSM) formed again when it is needed
        auto unfolded = [&](int j) -> double {
            if (MSM) return sblob[kp.o_M + j * n + age] * (my_slots[sl_age0 + 1 * n + j] * sblob[kp.o_invN + j]);
            return smb[j * THREADS + threadIdx.x];
        };
        if (!STRICT && !MSM) {
#pragma unroll
            for (int j = 0; j < NA; ++j) smb[j * THREADS + threadIdx.x] = q.M[j] * __shfl_sync(FULL, q.hN, j, NA);
        }
        auto refold = [&](int s) {     // row[j] = M(age, j) h_infec_j / N_j * (beta*kappa)(segment s) * a_age
            const double f = my_beff[s] * q.a;
#pragma unroll
            for (int j = 0; j < NA; ++j) {
                const double v = unfolded(j) * f;
                if (MSM) smf[j * THREADS + threadIdx.x] = v; else q.M[j] = v;
            }
            mseg = s;
        };
        auto unfold = [&]() {          // MSM, around a mixed-segment attempt: the array holds the unfolded row
#pragma unroll
            for (int j = 0; j < NA; ++j) smf[j * THREADS + threadIdx.x] = unfolded(j);
        };

        // ---- initial state (ObjectiveFunction.cpp:124-163) -------------------------------------------
        double x[NCOMP], k1[NCOMP];
        const double popN = sblob[kp.o_pop + age];
        if (kp.init_states != nullptr) {
            // Simulator::run(initial_state, times): the caller's state is integrated as given (Simulator.cpp:60-150)
            const double* st0 = kp.init_states + b * kp.init_stride;
#pragma unroll
            for (int c = 0; c < NCOMP; ++c) x[c] = st0[c * n + age];
        } else {
            const double runup_days = my_slots[sl_runup], seed_exposed = my_slots[sl_seed];
            if (runup_days > 0 && seed_exposed > 0) {
                x[1] = O::mul(seed_exposed, sblob[kp.o_agefrac + age]);
#pragma unroll
                for (int c = 2; c < NCOMP; ++c) x[c] = 0.0;
            } else {
#pragma unroll
                for (int c = 1; c <= 8; ++c) x[c] = O::mul(sblob[kp.o_init + c * n + age], my_slots[sl_mult0 + c - 1]);
                x[9] = sblob[kp.o_init + 9 * n + age];
                x[10] = sblob[kp.o_init + 10 * n + age];
            }
            double sum = 0;
#pragma unroll
            for (int j = 1; j < 9; ++j) sum = O::add(sum, x[j]);
            const unsigned gsh = (threadIdx.x & 31) - age;
            const unsigned gm = (NA >= 32) ? FULL : (((1u << NA) - 1u) << gsh);
            const bool over = (__ballot_sync(FULL, sum > popN) & gm) != 0;
            if (over && status == 0) status |= SEPAIHRD_ST_S_OVERFLOW;
            x[0] = O::sub(popN, sum);
        }

        double ll_acc_h = 0.0, ll_acc_i = 0.0, ll_acc_d = 0.0;   // STRICT: per stream; FAST: ll_acc_h only
        int n_acc = 0, n_rej = 0;
        double* traj_out = nullptr;
        size_t t_rs = 0, t_cs = 1;     // row / column strides of this set's trajectory block
        int W = 0;
        // FAST: the series are MODE_PPC's only output; STRICT decides at run time inside its MODE_TRAJ instantiation
        const bool is_ppc = STRICT ? (MODE == MODE_TRAJ && kp.traj_what == TRAJ_PPC_SERIES) : (MODE == MODE_PPC);
        if (MODE != MODE_LL) {
            W = (kp.traj_what == SEPAIHRD_TRAJ_FULL) ? NCOMP * n : 3 * n;
            traj_out = kp.traj_draw_minor ? kp.out_traj + (size_t)b : kp.out_traj + (size_t)b * kp.traj_rows * W;
            t_rs = kp.traj_draw_minor ? (size_t)W * (size_t)kp.B : (size_t)W;
            t_cs = kp.traj_draw_minor ? (size_t)kp.B : (size_t)1;
        }
        bool alive = (status == 0);
        // ---- integrate_times: observer at every grid point, adaptive steps in between ---------------------
        double dt = kp.dt_hint;
        double prev_h = x[9], prev_i = x[10], prev_d = x[8];   // row 0 is differenced against the initial state
        // schedule bookkeeping: seg = number of merged breakpoints strictly below t ("t <= end_k" picks k, so a
        // step that STARTS on a breakpoint still has its FSAL k1 in the old segment: quirk Q2)
        double t = s_times[0];
        int seg = 0;
        while (seg < nseg && t > s_bp[seg]) ++seg;
        double bp_next = (seg < nseg) ? s_bp[seg] : INF;
        double ba = O::mul(my_beff[seg], q.a);
        {   // controlled stepper initialise(): dxdt = f(x, t0)
            double y0[NDYN], d0[NDYN], p0[NPAS];
#pragma unroll
            for (int c = 0; c < NDYN; ++c) y0[c] = x[c];
            pi_slot ^= THREADS;
            if (!STRICT) refold(seg);
            rhs<NA, STRICT, true, true, MSM>(q, ba, spi, pi_slot, threadIdx.x, y0, d0, p0, smf + threadIdx.x, THREADS);
#pragma unroll
            for (int c = 0; c < NDYN; ++c) k1[c] = d0[c];
#pragma unroll
            for (int c = 0; c < NPAS; ++c) k1[NDYN + c] = p0[c];
        }

        // TRAJ_PPC_SERIES: element (series s, day r, age) of this draw sits at ((s T + r) n + age) B + b
        const size_t ppc_day = (size_t)n * (size_t)kp.ppc_B, ppc_series = (size_t)kp.traj_rows * ppc_day;
        auto ppc_observe = [&](int idx, double& ph, double& pi, double& pd, double& rh, double& ri, double& rd) {
            // ppc_series rule: d = max(0, v - prev), a NaN value stays NaN (kept out of the quantiles); run += d
            const double vh = x[9], vi = x[10], vd = x[8];
            double dh = vh - ph, di = vi - pi, dd = vd - pd;
            dh = (0.0 < dh) ? dh : 0.0; di = (0.0 < di) ? di : 0.0; dd = (0.0 < dd) ? dd : 0.0;
            if (vh != vh) dh = vh;
            if (vi != vi) di = vi;
            if (vd != vd) dd = vd;
            ph = vh; pi = vi; pd = vd;
            const int r = idx - kp.runup_offset;
            if (r >= 0) {
                rh += dh; ri += di; rd += dd;
                if (have && alive) {
                    double* o = kp.out_traj + ((size_t)r * n + age) * (size_t)kp.ppc_B + (size_t)(kp.ppc_b0 + b);
                    o[0 * ppc_series] = dh; o[1 * ppc_series] = di; o[2 * ppc_series] = dd;
                    o[3 * ppc_series] = rh; o[4 * ppc_series] = ri; o[5 * ppc_series] = rd;
                }
            }
        };

        if constexpr (LOOP == 6 && !STRICT) {
        // ================= LOOP 6 (FAST): one short decision section per attempt ============================
        // Same arithmetic as the loop below; what changes is the control flow around the attempt body:
        //  * a finished lane group idles with a ZERO-length step (cur = 0), so nothing in the body is predicated;
        //  * accept/reject is decided by exact compares (num > den); the VALUE of the error norm is needed only to
        //    size the next step, and three cases need no value at all.  They are recognised from the high words of
        //    num and den (hi(num) - hi(den) = log2(num/den) within +-0.0862, in units of 2^-20):
        //       every ratio surely <= 5^-5   -> growth factor is the constant 0.9 * (5^-5)^(-1/5)
        //       some ratio surely  >= 0.5    -> accepted step does not grow
        //       some ratio surely  >= 91.2   -> rejected step shrinks by exactly 1/5
        //    only the remaining groups run the tournament + division + power (one shared instruction stream);
        //  * all votes of an attempt are independent of each other (issued back to back), the loop-continuation
        //    mask is derived from them with uniform integer arithmetic;
        //  * the three Poisson streams of an output day are evaluated branch-free in one basic block.
        bool bad = false;                                   // a scored likelihood term saw a NaN/inf incidence
        int fail_steps = 0;
        const unsigned lane_bit = 1u << (threadIdx.x & 31);
        for (int idx = 0; idx < K; ++idx) {
            t = s_times[idx];
#ifdef SEPAIHRD_DEBUG_INTERVALS   // diagnostic build (tools/strict_parity_diag.py): out_steps is [B][K][2], running totals at every grid point
            if (MODE == MODE_LL && have && age == 0 && kp.out_steps) { kp.out_steps[(b * K + idx) * 2] = n_acc; kp.out_steps[(b * K + idx) * 2 + 1] = n_rej; }
#endif
            if constexpr (PROFILE) { if (have && age == 0) kp.out_profile[b * K + idx] = n_acc + n_rej; }
            if (MODE != MODE_LL) {
                if (is_ppc) {
                    ppc_observe(idx, prev_h, prev_i, prev_d, ll_acc_h, ll_acc_i, ll_acc_d);     // the likelihood accumulators are free here: running sums
                } else if (have && alive && (idx % kp.traj_stride == 0)) {
                    double* row = traj_out + (size_t)(idx / kp.traj_stride) * t_rs;
                    if (kp.traj_what == SEPAIHRD_TRAJ_FULL) {
#pragma unroll
                        for (int c = 0; c < NCOMP; ++c) row[(size_t)(c * n + age) * t_cs] = x[c];
                    } else {
                        row[(size_t)(0 * n + age) * t_cs] = x[8]; row[(size_t)(1 * n + age) * t_cs] = x[9]; row[(size_t)(2 * n + age) * t_cs] = x[10];
                    }
                }
            } else {
                const double inc_h = std_max(x[9] - prev_h, 0.0);
                const double inc_i = std_max(x[10] - prev_i, 0.0);
                const double inc_d = std_max(x[8] - prev_d, 0.0);
                prev_h = x[9]; prev_i = x[10]; prev_d = x[8];
                const int r = idx - kp.runup_offset;
                if (r >= 0) {
                    const double oh = s_obs_h[r * n + age], oi = s_obs_i[r * n + age], od = s_obs_d[r * n + age];
                    const bool vh = (oh >= 0.0), vi = (oi >= 0.0), vd = (od >= 0.0);   // skipped observations are stored as -1
                    const double sh = inc_h + 1e-10, si = inc_i + 1e-10, sd = inc_d + 1e-10;
                    const double th = fma(oh, fast_log_nb(sh, s_logtab, vh, bad), -sh);
                    const double ti = fma(oi, fast_log_nb(si, s_logtab, vi, bad), -si);
                    const double td = fma(od, fast_log_nb(sd, s_logtab, vd, bad), -sd);
                    ll_acc_h += ((vh ? th : 0.0) + (vi ? ti : 0.0)) + (vd ? td : 0.0);
                }
            }
            if (idx + 1 == K) break;
            const double t_next = s_times[idx + 1];
            double rem = t_next - t;
            bool active = alive && (rem > DBL_EPSILON);            // less_with_sign(t, t_next, dt)
            unsigned m_active = __ballot_sync(FULL, active);
            if (m_active == 0) {
                if (!__any_sync(FULL, alive)) break;
                continue;
            }
            // a breakpoint inside [t, t_next): attempts of this day may leave their schedule segment
            const bool day_bp = __any_sync(FULL, bp_next < t_next);
            bool first_of_day = true;
            while (true) {
                const double cur = active ? std_min(dt, rem) : 0.0;   // min_abs(dt, t_next - t)
                const double t_end = t + cur;
                StepSched sc;
                sc.ba_step = ba; sc.s_lo = seg; sc.bp = s_bp; sc.beff = my_beff; sc.nseg = nseg; sc.a = q.a;
                sc.mb = smb + threadIdx.x; sc.mf = smf + threadIdx.x; sc.mb_stride = THREADS;
                int s_hi = seg;
                bool run_mixed = false;
                if (day_bp) {
                    if (__any_sync(FULL, !(t_end <= bp_next))) {
                        bool mixed = false;
                        if (!(t_end <= bp_next)) {
                            const double t2 = fma(cur, c_tab[T_A2], t);
                            int s_lo = seg;
                            while (s_lo < nseg && t2 > s_bp[s_lo]) ++s_lo;
                            s_hi = s_lo;
                            while (s_hi < nseg && t_end > s_bp[s_hi]) ++s_hi;
                            mixed = (s_lo != s_hi);
                            sc.s_lo = s_lo;
                            sc.ba_step = my_beff[s_lo] * q.a;
                            if (s_lo != mseg) refold(s_lo);   // stages 2..7 of a step that starts ON a breakpoint (quirk Q2)
                        }
                        if (!ONGRID) run_mixed = __any_sync(FULL, mixed);
                    }
                }
                double xn[NDYN], k7d[NDYN], k7p[NPAS], accN[NPAS], xe[NCOMP];
                const double ecur = cur * kp.inv_rel;
                // first attempt of a breakpoint-free day with every stepping group at the full step hmax: constant coefficients
                const bool unit = first_of_day && !day_bp && __all_sync(FULL, !active || cur == hmax);
                first_of_day = false;
                if (unit)
                    dopri5_attempt<NA, false, false, true, MSM>(q, sc, spi, pi_slot, THREADS, threadIdx.x, t, cur, t_end, x, k1, xn, k7d, k7p, accN, xe, ecur, kp.hc);
                else if (!ONGRID && run_mixed) {
                    if (MSM) unfold();
                    dopri5_attempt<NA, false, true, false, MSM>(q, sc, spi, pi_slot, THREADS, threadIdx.x, t, cur, t_end, x, k1, xn, k7d, k7p, accN, xe, ecur);
                    if (MSM) refold(mseg);
                }
                else
                    dopri5_attempt<NA, false, false, false, MSM>(q, sc, spi, pi_slot, THREADS, threadIdx.x, t, cur, t_end, x, k1, xn, k7d, k7p, accN, xe, ecur);
                // ---- decision: exact compares; coarse magnitude of the worst ratio from the high words ---------
                double num[NCOMP], den[NCOMP];
                bool big0 = false, big1 = false, big2 = false;
#pragma unroll
                for (int c = 0; c < NCOMP; ++c) {
                    den[c] = fma(cur, fabs(k1[c]), fabs(x[c])) + kp.abs_over_rel;
                    const bool g = fabs(xe[c]) > den[c];          // |.| is an operand modifier of the DSETP
                    if (c % 3 == 0) big0 |= g; else if (c % 3 == 1) big1 |= g; else big2 |= g;
                }
                // all votes of the attempt sit together after the body (measured: 1 % faster than voting before it)
                const unsigned m_more = __ballot_sync(FULL, active && ((t_next - t_end) > DBL_EPSILON));
                const unsigned bal_big = __ballot_sync(FULL, big0 | big1 | big2);
                if (unit && ((bal_big & m_active) | m_more) == 0) {
                    // The common day: every stepping group took the full step and accepted it.  dt >= hmax, so nothing about
                    // the error norm's value can matter: commit and leave the day.
                    if (active) {
#ifdef SEPAIHRD_DEBUG_INTERVALS
                        if (MODE == MODE_LL && have && age == 0 && kp.out_traj && n_acc + n_rej < 2048) {
                            double* tr = kp.out_traj + ((size_t)b * 2048 + (n_acc + n_rej)) * 3;
                            tr[0] = t; tr[1] = cur; tr[2] = -2.0;
                        }
#endif
                        t = t_end;
                        ++n_acc;
                        fail_steps = 0;
#pragma unroll
                        for (int c = 0; c < NDYN; ++c) { x[c] = xn[c]; k1[c] = k7d[c]; }
#pragma unroll
                        for (int c = 0; c < NPAS; ++c) { x[NDYN + c] = accN[c]; k1[NDYN + c] = k7p[c]; }
                    }
                    break;
                }
                int lm0 = INT_MIN, lm1 = INT_MIN;
                // |xe| as a value is needed only from here on (magnitude classes, tournament): clear the sign bit with an
                // integer op instead of a DADD on the FP64 pipe (v10: +0.6 % A/B)
#pragma unroll
                for (int c = 0; c < NCOMP; ++c) num[c] = __hiloint2double(__double2hiint(xe[c]) & 0x7fffffff, __double2loint(xe[c]));
#pragma unroll
                for (int c = 0; c < NCOMP; ++c) {
                    const int l = __double2hiint(num[c]) - __double2hiint(den[c]);
                    if (c & 1) lm1 = max(lm1, l); else lm0 = max(lm0, l);
                }
                const int lmax = max(lm0, lm1);
                const unsigned m_low = __ballot_sync(FULL, dt < hmax);
                const unsigned bal_ng = __ballot_sync(FULL, lmax > kp.thr_nogrow);
                const unsigned bal_ns = __ballot_sync(FULL, !(lmax < kp.thr_small));
                const unsigned bal_sb = __ballot_sync(FULL, lmax > kp.thr_big);
                const unsigned g_rej = expand_groups<NA>(bal_big) & m_active;
                const unsigned g_ng = expand_groups<NA>(bal_ng), g_ns = expand_groups<NA>(bal_ns), g_sb = expand_groups<NA>(bal_sb);
                const unsigned m_val = (g_rej & ~g_sb) | (m_active & ~g_rej & m_low & g_ns & ~g_ng);
                const bool reject = (g_rej & lane_bit) != 0;
                double err = 0.0, facv = 0.0;
                if (m_val != 0) {
                    // arg-max of num/den by a cross-multiplication tournament, then ONE division per lane (eleven parallel
                    // reciprocal chains instead were measured: no faster, and their 11 extra live doubles spill)
#pragma unroll
                    for (int stride = 1; stride < NCOMP; stride *= 2) {
#pragma unroll
                        for (int c = 0; c + stride < NCOMP; c += 2 * stride) {
                            const bool other = num[c + stride] * den[c] > num[c] * den[c + stride];
                            num[c] = other ? num[c + stride] : num[c];
                            den[c] = other ? den[c + stride] : den[c];
                        }
                    }
                    err = group_max<NA>(fast_div_pos(num[0], den[0]));
                    // decrease_step: 0.9 err^(-1/3) (error_order 4);  increase_step: 0.9 max(5^-5, err)^(-1/5) (stepper_order 5)
                    facv = 0.9 * pow_neg_inv(reject ? err : std_max(3.2e-4 /* pow(5,-5) */, err), reject);
                }
                unsigned m_dead = 0;
#ifdef SEPAIHRD_DEBUG_INTERVALS
                if (MODE == MODE_LL && active && have && age == 0 && kp.out_traj && n_acc + n_rej < 2048) {
                    double* tr = kp.out_traj + ((size_t)b * 2048 + (n_acc + n_rej)) * 3;
                    tr[0] = t; tr[1] = cur; tr[2] = (m_val != 0) ? err : (reject ? 1e300 : -1.0);
                }
#endif
                if (reject) {
                    const double shrink = ((g_sb & lane_bit) || err > 128.0) ? 0.2 : std_max(facv, 0.2);   // 0.9 err^(-1/3) < 0.2 beyond 91.2
                    dt = cur * shrink;
                    ++n_rej;
                    if (fail_steps++ >= 500) { status |= SEPAIHRD_ST_STEP_FAILURE; alive = false; }
                } else if (active) {
                    t = t_end;
                    rem = t_next - t;
                    if (dt < hmax) {
                        double g = 0.0;                                  // err >= 0.5: the step is kept
                        if (!(g_ns & lane_bit)) g = kp.grow_max;          // err <= 5^-5
                        else if (!(g_ng & lane_bit) && err < 0.5) g = facv;
                        dt = std_max(dt, cur * g);                       // max_abs(carried, proposed)
                    }
                    ++n_acc;
                    fail_steps = 0;
                    if (s_hi != seg) {
                        seg = s_hi;
                        bp_next = (seg < nseg) ? s_bp[seg] : INF;
                        ba = my_beff[seg] * q.a;
                        if (seg != mseg) refold(seg);
                    }
#pragma unroll
                    for (int c = 0; c < NDYN; ++c) { x[c] = xn[c]; k1[c] = k7d[c]; }
#pragma unroll
                    for (int c = 0; c < NPAS; ++c) { x[NDYN + c] = accN[c]; k1[NDYN + c] = k7p[c]; }
                }
                if (g_rej != 0) m_dead = __ballot_sync(FULL, !alive);   // only a rejection can exhaust the failed-step budget
                m_active = (m_more | g_rej) & ~m_dead;
                if (m_active == 0) break;
                active = (m_active & lane_bit) != 0;
            }
        }
        if (bad) ll_acc_h = __longlong_as_double(0x7ff8000000000000LL);
        } else {
        // ================= LOOP 5 (STRICT): the reference's arithmetic, operation by operation ====================
        // Unfused IEEE mul/add in source order, libm pow/log, true divisions, likelihood summed row by row per stream.
        for (int idx = 0; idx < K; ++idx) {
            t = s_times[idx];
#ifdef SEPAIHRD_DEBUG_INTERVALS
            if (MODE == MODE_LL && have && age == 0 && kp.out_steps) { kp.out_steps[(b * K + idx) * 2] = n_acc; kp.out_steps[(b * K + idx) * 2 + 1] = n_rej; }
#endif
            // ---- observer -----------------------------------------------------------------------------
            if (MODE != MODE_LL) {
                if (is_ppc) {
                    ppc_observe(idx, prev_h, prev_i, prev_d, ll_acc_h, ll_acc_i, ll_acc_d);
                } else if (have && alive && (idx % kp.traj_stride == 0)) {
                    double* row = traj_out + (size_t)(idx / kp.traj_stride) * t_rs;
                    if (kp.traj_what == SEPAIHRD_TRAJ_FULL) {
#pragma unroll
                        for (int c = 0; c < NCOMP; ++c) row[(size_t)(c * n + age) * t_cs] = x[c];
                    } else {
                        row[(size_t)(0 * n + age) * t_cs] = x[8]; row[(size_t)(1 * n + age) * t_cs] = x[9]; row[(size_t)(2 * n + age) * t_cs] = x[10];
                    }
                }
            } else {
                // daily incidence + Poisson terms (ObjectiveFunction.cpp:191-225, 241-279)
                const double inc_h = std_max(O::sub(x[9], prev_h), 0.0);
                const double inc_i = std_max(O::sub(x[10], prev_i), 0.0);
                const double inc_d = std_max(O::sub(x[8], prev_d), 0.0);
                prev_h = x[9]; prev_i = x[10]; prev_d = x[8];
                const int r = idx - kp.runup_offset;
                if (r >= 0) {
                    const double eps = 1e-10;
                    double term_h = 0.0, term_i = 0.0, term_d = 0.0;
                    const double oh = s_obs_h[r * n + age], oi = s_obs_i[r * n + age], od = s_obs_d[r * n + age];
                    // the host stores every skipped observation (negative, NaN, inf: ObjectiveFunction.cpp:267) as -1
                    if (oh >= 0.0) { double sim = inc_h; if (sim < 0.0) sim = 0.0; sim = O::add(sim, eps); term_h = O::sub(O::mul(oh, log(sim)), sim); }
                    if (oi >= 0.0) { double sim = inc_i; if (sim < 0.0) sim = 0.0; sim = O::add(sim, eps); term_i = O::sub(O::mul(oi, log(sim)), sim); }
                    if (od >= 0.0) { double sim = inc_d; if (sim < 0.0) sim = 0.0; sim = O::add(sim, eps); term_d = O::sub(O::mul(od, log(sim)), sim); }
                    // row_sum over ages in order, then log_likelihood += row_sum (per stream)
                    double rs_h = 0.0, rs_i = 0.0, rs_d = 0.0;
#pragma unroll
                    for (int j = 0; j < NA; ++j) {
                        const double th = __shfl_sync(FULL, term_h, j, NA), ti = __shfl_sync(FULL, term_i, j, NA),
                                     td = __shfl_sync(FULL, term_d, j, NA);
                        const double ohj = s_obs_h[r * n + j], oij = s_obs_i[r * n + j], odj = s_obs_d[r * n + j];
                        if (ohj >= 0.0) rs_h = O::add(rs_h, th);
                        if (oij >= 0.0) rs_i = O::add(rs_i, ti);
                        if (odj >= 0.0) rs_d = O::add(rs_d, td);
                    }
                    ll_acc_h = O::add(ll_acc_h, rs_h); ll_acc_i = O::add(ll_acc_i, rs_i); ll_acc_d = O::add(ll_acc_d, rs_d);
                }
            }
            if (idx + 1 == K) break;
            const double t_next = s_times[idx + 1];
            int fail_steps = 0;
            bool need = alive && ((t_next - t) > DBL_EPSILON);   // less_with_sign(t, t_next, dt)

            // ---- adaptive steps up to t_next: one attempt per iteration for the WHOLE warp ------------------
            while (true) {
                const unsigned m_need = __ballot_sync(FULL, need);   // group-uniform, so whole groups are set
                if (m_need == 0) break;
                double cur = std_min(dt, t_next - t);   // min_abs(dt, t_next - t)
                const double t_end = O::add(t, cur);
                StepSched sc;
                sc.ba_step = ba; sc.s_lo = seg; sc.bp = s_bp; sc.beff = my_beff; sc.nseg = nseg; sc.a = q.a;
                sc.mb = nullptr; sc.mf = nullptr; sc.mb_stride = 0;
                int s_hi = seg;
                bool run_mixed = false;
                if (__any_sync(FULL, !(t_end <= bp_next))) {
                    // Stage times lie in (t, t_end].  Steps that start on a breakpoint (quirk Q2) or straddle one on a
                    // general grid look their segments up.
                    bool mixed = false;
                    if (!(t_end <= bp_next)) {
                        const double t2 = O::add(t, O::mul(cur, c_tab[T_A2]));
                        int s_lo = seg;
                        while (s_lo < nseg && t2 > s_bp[s_lo]) ++s_lo;
                        s_hi = s_lo;
                        while (s_hi < nseg && t_end > s_bp[s_hi]) ++s_hi;
                        mixed = (s_lo != s_hi);
                        sc.s_lo = s_lo;
                        sc.ba_step = O::mul(my_beff[s_lo], q.a);
                    }
                    run_mixed = __any_sync(FULL, mixed);
                }
                double xn[NDYN], k7d[NDYN], k7p[NPAS], accN[NPAS], xe[NCOMP];
                if (run_mixed)
                    dopri5_attempt<NA, STRICT, true>(q, sc, spi, pi_slot, THREADS, threadIdx.x, t, cur, t_end, x, k1, xn, k7d, k7p, accN, xe, cur);
                else
                    dopri5_attempt<NA, STRICT, false>(q, sc, spi, pi_slot, THREADS, threadIdx.x, t, cur, t_end, x, k1, xn, k7d, k7p, accN, xe, cur);
                // error norm (default_error_checker): max_c |xerr_c| / (abs + rel * (|x_c| + dt * |dxdt_c|))
                double m = 0.0;
#pragma unroll
                for (int c = 0; c < NCOMP; ++c) {
                    const double den = O::add(f_abs, O::mul(f_rel, O::add(fabs(x[c]), O::mul(cur, fabs(k1[c])))));
                    const double v = fabs(__ddiv_rn(fabs(xe[c]), den));
                    m = (m < v) ? v : m;
                }
                const double err = group_max<NA>(m);
#ifdef SEPAIHRD_DEBUG_INTERVALS
                if (MODE == MODE_LL && need && have && age == 0 && kp.out_traj && n_acc + n_rej < 2048) {
                    double* tr = kp.out_traj + ((size_t)b * 2048 + (n_acc + n_rej)) * 3;
                    tr[0] = t; tr[1] = cur; tr[2] = err;
                }
#endif
                if (need) {
                    if (err > 1.0) {
                        // decrease_step (error_order 4): dt *= max(0.9 * err^(-1/3), 1/5)
                        cur = O::mul(cur, std_max(O::mul(9.0 / 10.0, pow(err, -1.0 / 3.0)), 1.0 / 5.0));
                        ++n_rej;
                        dt = cur;
                        if (fail_steps++ >= 500) { status |= SEPAIHRD_ST_STEP_FAILURE; alive = false; }
                    } else {
                        // accept: t += dt; increase_step (stepper_order 5) when err < 0.5
                        t = t_end;
                        if (err < 0.5) {
                            const double e2 = std_max(3.2e-4 /* pow(5,-5) */, err);
                            cur = O::mul(cur, O::mul(9.0 / 10.0, pow(e2, -1.0 / 5.0)));
                        }
                        dt = std_max(dt, cur);   // max_abs: keep the larger of the carried and the proposed step
                        ++n_acc;
                        fail_steps = 0;
                        if (s_hi != seg) {
                            seg = s_hi;
                            bp_next = (seg < nseg) ? s_bp[seg] : INF;
                            ba = O::mul(my_beff[seg], q.a);
                        }
#pragma unroll
                        for (int c = 0; c < NDYN; ++c) { x[c] = xn[c]; k1[c] = k7d[c]; }
#pragma unroll
                        for (int c = 0; c < NPAS; ++c) { x[NDYN + c] = accN[c]; k1[NDYN + c] = k7p[c]; }
                    }
                }
                need = alive && ((t_next - t) > DBL_EPSILON);
            }
            if (!__any_sync(FULL, alive)) break;
        }

        }   // LOOP

        // ---- epilogue -------------------------------------------------------------------------------------
        if (MODE == MODE_LL) {
            double total;
            if (STRICT) total = O::add(O::add(ll_acc_h, ll_acc_i), ll_acc_d);   // ll_hosp + ll_icu + ll_deaths
            else total = group_sum<NA>(ll_acc_h);
            if (status != 0) total = -DBL_MAX;
            else if (isnan(total) || isinf(total)) { total = -DBL_MAX; status |= SEPAIHRD_ST_NONFINITE; }
            if (have && age == 0) {
                kp.out_ll[b] = total;
                if (kp.out_status) kp.out_status[b] = status;
#ifndef SEPAIHRD_DEBUG_INTERVALS
                if (kp.out_steps) { kp.out_steps[2 * b] = n_acc; kp.out_steps[2 * b + 1] = n_rej; }
#endif
            }
        } else {
            if (have && status != 0) {   // failed sets: NaN-fill every row
                const double qnan = __longlong_as_double(0x7ff8000000000000LL);
                if (is_ppc) {
                    for (int r = 0; r < kp.traj_rows; ++r)
                        for (int sidx = 0; sidx < 6; ++sidx) kp.out_traj[(size_t)sidx * ppc_series + ((size_t)r * n + age) * (size_t)kp.ppc_B + (size_t)(kp.ppc_b0 + b)] = qnan;
                } else {
                    for (int r = 0; r < kp.traj_rows; ++r)
                        for (int w = age; w < W; w += NA) traj_out[(size_t)r * t_rs + (size_t)w * t_cs] = qnan;
                }
            }
            if (have && age == 0) {
                if (kp.out_status) kp.out_status[b] = status;
                if (kp.out_steps) { kp.out_steps[2 * b] = n_acc; kp.out_steps[2 * b + 1] = n_rej; }
            }
        }
    }
}

}  // namespace sepaihrd
