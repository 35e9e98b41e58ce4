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
#include <cstdint>
#include <cuda_runtime.h>

#include "sepaihrd_b200.h"

namespace sepaihrd {

constexpr int NCOMP = SEPAIHRD_NUM_COMPARTMENTS;  // 11
constexpr int NDYN = 7;                            // S E P A I H ICU feed the RHS
constexpr int NPAS = 4;                            // R D CumH CumICU do not

enum { MODE_LL = 0, MODE_TRAJ = 1 };

// Everything the kernel needs besides the staged blob; passed by value (constant bank).
struct KParams {
    const double* blob;        // device copy of the constants blob (16-byte aligned)
    int blob_bytes;            // multiple of 16
    // offsets (in doubles) of the arrays inside the blob
    int o_times, o_obs_h, o_obs_i, o_obs_d, o_pop, o_agefrac, o_invN, o_M, o_bp, o_base, o_init, o_lo, o_hi;
    int o_pslot, o_segb, o_segk;   // int32 arrays, offsets still in doubles
    int n, K, n_obs, runup_offset, nb, nk, nseg, P, nslots;
    int slot_stride;           // doubles per set in the shared slot table (odd -> conflict-free)
    int seg_stride;            // doubles per set in the shared beta_eff table
    int constraint_mode;       // 0 clamp, 1 reflect
    double abs_tol, rel_tol, dt_hint;
    double hmax;               // longest output interval: growing dt beyond it cannot change the result
    // I/O
    const double* params;      // [B][ld]
    long long B, ld;
    double* out_ll;            // [B]
    unsigned* out_status;      // [B] or null
    int* out_steps;            // [B][2] or null
    double* out_traj;          // MODE_TRAJ: [B][traj_rows][W]
    int traj_what, traj_stride, traj_rows;
    long long tiles;           // ceil(B / sets_per_block)
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

// std::max(a, b) == (a < b) ? b : a   (NaN in b is ignored, NaN in a is returned)
__device__ __forceinline__ double std_max(double a, double b) { return (a < b) ? b : a; }
__device__ __forceinline__ double std_min(double a, double b) { return (b < a) ? b : a; }

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

// ---- constraints ------------------------------------------------------------------------------------
// reflectBound, SEPAIHRDParameterManager.cpp:302-313
__device__ __forceinline__ double reflect_bound(double value, double minb, double maxb) {
    if (minb >= maxb) return minb;
    const double width = maxb - minb;
    double y = fmod(value - minb, 2.0 * width);
    if (y < 0) y += 2.0 * width;
    if (y <= width) return minb + y;
    return maxb - (y - width);
}
// applyConstraints, .cpp:315-347
__device__ __forceinline__ double constrain(double v, double lo, double hi, int mode) {
    if (lo == lo) {   // has a bounds entry
        if (lo > hi) { double t = lo; lo = hi; hi = t; }
        return (mode == 0) ? std_min(std_max(v, lo), hi) : reflect_bound(v, lo, hi);
    }
    return (mode == 0) ? std_max(0.0, v) : fabs(v);
}

// ---- per-lane model parameters (registers) ----------------------------------------------------------
template <int NA>
struct LaneParams {
    double theta, sigma, gamma_p, gamma_A, gamma_I, gamma_H, gamma_ICU;
    double a, hinf, invN, p, h, icu, dH, dICU, dcomm;
    double hN, kI, kH, kU;   // FAST: h_infec/N, gamma_I+h+d_comm, gamma_H+d_H+icu, gamma_ICU+d_ICU
    double M[NA];            // row `age` of the contact matrix: M(age, j)
};

// AgeSEPAIHRDModel::computeDerivatives for one age class (this lane), inputs y = S E P A I H ICU.
// PASSIVE=false skips dR dD dCumH dCumICU (stage 2: Dopri5 has c2 = dc2 = 0).
template <int NA, bool STRICT, bool PASSIVE>
__device__ __forceinline__ void rhs(const LaneParams<NA>& q, unsigned gmask, double ba, const double (&y)[NDYN],
                                    double (&d)[NCOMP]) {
    using O = Ops<STRICT>;
    const double S = y[0], E = y[1], P = y[2], A = y[3], I = y[4], H = y[5], U = y[6];
    double pressure;
    if (STRICT) {
        const double total_inf = O::add(O::add(P, A), O::mul(q.theta, I));   // :155
        pressure = O::mul(O::mul(total_inf, q.hinf), q.invN);                // :156
    } else {
        pressure = fma(q.theta, I, P + A) * q.hN;
    }
    double lam = 0.0;                                                        // :162-174 (j outer, in order)
#pragma unroll
    for (int j = 0; j < NA; ++j) {
        const double pj = __shfl_sync(gmask, pressure, j, NA);
        lam = O::mad(q.M[j], pj, lam);
    }
    lam = O::mul(lam, ba);                                                   // :181-183, ba = (beta*kappa)*a_i
    lam = (0.0 < lam) ? lam : 0.0;                                           // std::max(0.0, lambda) :196
    const double flow_SE = O::mul(lam, S);
    const double flow_IH = O::mul(q.h, I);
    const double flow_H_ICU = O::mul(q.icu, H);
    d[0] = -flow_SE;
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
        d[1] = O::sub(flow_SE, flow_EP);
        d[2] = O::sub(flow_EP, flow_P_out);
        d[3] = O::sub(flow_PA, gAA);
        d[4] = O::sub(flow_PI, I_out);
        d[5] = O::sub(flow_IH, H_out);
        d[6] = O::sub(flow_H_ICU, U_out);
        if (PASSIVE) {
            const double gUU = O::mul(q.gamma_ICU, U), dUU = O::mul(q.dICU, U);
            d[7] = O::add(O::add(O::add(gAA, flow_IR), gH), gUU);            // :222
            d[8] = O::add(O::add(dHH, dUU), flow_IDc);                      // :223
            d[9] = flow_IH;
            d[10] = flow_H_ICU;
        }
    } else {
        const double flow_P_out = q.gamma_p * P;
        const double flow_PA = q.p * flow_P_out;
        d[1] = fma(-q.sigma, E, flow_SE);
        d[2] = fma(q.sigma, E, -flow_P_out);
        d[3] = fma(-q.gamma_A, A, flow_PA);
        d[4] = fma(-q.kI, I, flow_P_out - flow_PA);
        d[5] = fma(-q.kH, H, flow_IH);
        d[6] = fma(-q.kU, U, flow_H_ICU);
        if (PASSIVE) {
            d[7] = fma(q.gamma_ICU, U, fma(q.gamma_H, H, fma(q.gamma_I, I, q.gamma_A * A)));
            d[8] = fma(q.dcomm, I, fma(q.dICU, U, q.dH * H));
            d[9] = flow_IH;
            d[10] = flow_H_ICU;
        }
    }
}

// Dopri5 tableau as Boost writes it (runge_kutta_dopri5.hpp): integer ratios evaluated in double.
struct Tab {
    static constexpr double a2 = 1.0 / 5.0, a3 = 3.0 / 10.0, a4 = 4.0 / 5.0, a5 = 8.0 / 9.0;
    static constexpr double b21 = 1.0 / 5.0;
    static constexpr double b31 = 3.0 / 40.0, b32 = 9.0 / 40.0;
    static constexpr double b41 = 44.0 / 45.0, b42 = -56.0 / 15.0, b43 = 32.0 / 9.0;
    static constexpr double b51 = 19372.0 / 6561.0, b52 = -25360.0 / 2187.0, b53 = 64448.0 / 6561.0, b54 = -212.0 / 729.0;
    static constexpr double b61 = 9017.0 / 3168.0, b62 = -355.0 / 33.0, b63 = 46732.0 / 5247.0, b64 = 49.0 / 176.0,
                            b65 = -5103.0 / 18656.0;
    static constexpr double c1 = 35.0 / 384.0, c3 = 500.0 / 1113.0, c4 = 125.0 / 192.0, c5 = -2187.0 / 6784.0,
                            c6 = 11.0 / 84.0;
    static constexpr double dc1 = c1 - 5179.0 / 57600.0, dc3 = c3 - 7571.0 / 16695.0, dc4 = c4 - 393.0 / 640.0;
    static constexpr double dc5 = c5 - -92097.0 / 339200.0, dc6 = c6 - 187.0 / 2100.0, dc7 = -1.0 / 40.0;
};

template <int NA>
__device__ __forceinline__ double group_max(unsigned gmask, double v) {
#pragma unroll
    for (int off = NA / 2; off >= 1; off >>= 1) {
        const double o = __shfl_xor_sync(gmask, v, off, NA);
        v = (v < o) ? o : v;
    }
    return v;
}
template <int NA>
__device__ __forceinline__ double group_sum(unsigned gmask, double v) {
#pragma unroll
    for (int off = NA / 2; off >= 1; off >>= 1) v += __shfl_xor_sync(gmask, v, off, NA);
    return v;
}

// ------------------------------------------------------------------------------------------------
template <int NA, bool STRICT, int MODE, int THREADS, int MINBLOCKS>
__global__ void __launch_bounds__(THREADS, MINBLOCKS) sepaihrd_batch_kernel(const KParams kp) {
    using O = Ops<STRICT>;
    constexpr int SETS = THREADS / NA;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sblob = reinterpret_cast<double*>(smem_raw);
    const int blob_doubles = kp.blob_bytes >> 3;
    double* sslots = sblob + blob_doubles;
    double* sbeff = sslots + SETS * kp.slot_stride;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sbeff + SETS * kp.seg_stride);

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

    const int lane = threadIdx.x & 31;
    const int age = threadIdx.x % NA;
    const int grp = threadIdx.x / NA;
    const unsigned gmask = (NA >= 32) ? 0xffffffffu : (((1u << NA) - 1u) << (lane - age));
    const int n = NA;
    const int nseg = kp.nseg;

    double* my_slots = sslots + grp * kp.slot_stride;
    double* my_beff = sbeff + grp * kp.seg_stride;

    // slot layout (sepaihrd_b200.h)
    const int sl_beta0 = 0, sl_kappa0 = kp.nb, sl_scal0 = kp.nb + kp.nk, sl_age0 = sl_scal0 + 7;
    const int sl_mult0 = sl_age0 + 8 * n, sl_seed = sl_mult0 + 8, sl_runup = sl_mult0 + 9, sl_beta = sl_mult0 + 10;

    for (long long tile = blockIdx.x; tile < kp.tiles; tile += gridDim.x) {
        const long long b = tile * SETS + grp;
        if (b >= kp.B) continue;   // whole group idle (all group-scoped sync below uses gmask)

        // ---- updateModelParameters: base slots, then constrained calibrated values -------------------
        __syncwarp(gmask);
        for (int s = age; s < kp.nslots; s += NA) my_slots[s] = sblob[kp.o_base + s];
        __syncwarp(gmask);
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
        __syncwarp(gmask);
        unsigned status = 0;
        kappa_touched = __any_sync(gmask, kappa_touched);
        if (kappa_touched) {   // setCalibratableValues throws on a negative kappa (NPI.cpp:238-242)
            bool neg = false;
            for (int k = 1 + age; k < kp.nk; k += NA) neg |= (my_slots[sl_kappa0 + k] < 0.0);
            if (__any_sync(gmask, neg)) status |= SEPAIHRD_ST_INVALID_PARAM;
        }
        // beta_eff per merged schedule segment: beta(t) * kappa(t)   (AgeSEPAIHRDModel.cpp:176-178)
        for (int s = age; s <= nseg; s += NA) {
            const double bv = (kp.nb > 0) ? my_slots[sl_beta0 + s_segb[s]] : my_slots[sl_beta];
            my_beff[s] = O::mul(bv, my_slots[sl_kappa0 + s_segk[s]]);
        }
        __syncwarp(gmask);

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

        // ---- initial state (ObjectiveFunction.cpp:124-163) -------------------------------------------
        double x[NCOMP], k1[NCOMP];
        const double popN = sblob[kp.o_pop + age];
        {
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
            if (__any_sync(gmask, sum > popN) && status == 0) status |= SEPAIHRD_ST_S_OVERFLOW;
            x[0] = O::sub(popN, sum);
        }

        double ll_acc_h = 0.0, ll_acc_i = 0.0, ll_acc_d = 0.0;   // STRICT: per stream; FAST: ll_acc_h only
        int n_acc = 0, n_rej = 0;
        double* traj_out = nullptr;
        int W = 0;
        if (MODE == MODE_TRAJ) {
            W = (kp.traj_what == SEPAIHRD_TRAJ_FULL) ? NCOMP * n : 3 * n;
            traj_out = kp.out_traj + (size_t)b * kp.traj_rows * W;
        }

        if (status == 0) {
            // ---- integrate_times: observer at every grid point, adaptive steps in between -----------------
            double dt = kp.dt_hint;
            double prev_h = x[9], prev_i = x[10], prev_d = x[8];   // row 0 is differenced against the initial state
            int seg = 0;          // number of merged breakpoints strictly below the current time
            int seg_ba = -1;      // segment for which `ba` is current
            double ba = 0.0;
            bool have_k1 = false;
            const double f_abs = kp.abs_tol, f_rel = kp.rel_tol;

            for (int idx = 0; idx < kp.K; ++idx) {
                double t = s_times[idx];
                // ---- observer -----------------------------------------------------------------------------
                if (MODE == MODE_TRAJ) {
                    if (idx % kp.traj_stride == 0) {
                        double* row = traj_out + (size_t)(idx / kp.traj_stride) * W;
                        if (kp.traj_what == SEPAIHRD_TRAJ_FULL) {
#pragma unroll
                            for (int c = 0; c < NCOMP; ++c) row[c * n + age] = x[c];
                        } else {
                            row[0 * n + age] = x[8]; row[1 * n + age] = x[9]; row[2 * n + age] = x[10];
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
                        const bool vh = (oh >= 0.0) && isfinite(oh), vi = (oi >= 0.0) && isfinite(oi),
                                   vd = (od >= 0.0) && isfinite(od);
                        if (vh) { double sim = inc_h; if (sim < 0.0) sim = 0.0; sim = O::add(sim, eps); term_h = O::sub(O::mul(oh, log(sim)), sim); }
                        if (vi) { double sim = inc_i; if (sim < 0.0) sim = 0.0; sim = O::add(sim, eps); term_i = O::sub(O::mul(oi, log(sim)), sim); }
                        if (vd) { double sim = inc_d; if (sim < 0.0) sim = 0.0; sim = O::add(sim, eps); term_d = O::sub(O::mul(od, log(sim)), sim); }
                        if (STRICT) {
                            // row_sum over ages in order, then log_likelihood += row_sum (per stream)
                            double rs_h = 0.0, rs_i = 0.0, rs_d = 0.0;
#pragma unroll
                            for (int j = 0; j < NA; ++j) {
                                const double th = __shfl_sync(gmask, term_h, j, NA), ti = __shfl_sync(gmask, term_i, j, NA),
                                             td = __shfl_sync(gmask, term_d, j, NA);
                                const double ohj = s_obs_h[r * n + j], oij = s_obs_i[r * n + j], odj = s_obs_d[r * n + j];
                                if ((ohj >= 0.0) && isfinite(ohj)) rs_h = O::add(rs_h, th);
                                if ((oij >= 0.0) && isfinite(oij)) rs_i = O::add(rs_i, ti);
                                if ((odj >= 0.0) && isfinite(odj)) rs_d = O::add(rs_d, td);
                            }
                            ll_acc_h = O::add(ll_acc_h, rs_h); ll_acc_i = O::add(ll_acc_i, rs_i); ll_acc_d = O::add(ll_acc_d, rs_d);
                        } else {
                            ll_acc_h += (term_h + term_i) + term_d;
                        }
                    }
                }
                if (idx + 1 == kp.K) break;
                const double t_next = s_times[idx + 1];
                int fail_steps = 0;

                // ---- adaptive steps up to t_next --------------------------------------------------------------
                while ((t_next - t) > DBL_EPSILON) {
                    double cur = std_min(dt, t_next - t);   // min_abs(dt, t_next - t)
                    // schedule segment bookkeeping: stage times increase within a step
                    int s_stage = seg;
                    auto ba_at = [&](double ts) -> double {
                        while (s_stage < nseg && ts > s_bp[s_stage]) ++s_stage;
                        if (s_stage != seg_ba) { ba = O::mul(my_beff[s_stage], q.a); seg_ba = s_stage; }
                        return ba;
                    };
                    if (!have_k1) {   // controlled stepper initialise(): dxdt = f(x, t0)
                        double y0[NDYN];
#pragma unroll
                        for (int c = 0; c < NDYN; ++c) y0[c] = x[c];
                        rhs<NA, STRICT, true>(q, gmask, ba_at(t), y0, k1);
                        have_k1 = true;
                    }
                    double k2[NDYN], k3[NDYN], k4[NDYN], k5[NDYN], k6[NDYN];
                    double kk[NCOMP], y[NDYN];
                    double accN[NPAS], accE[NPAS];   // running solution / error sums of the passive compartments
                    // stage 2
                    { const double f1 = O::mul(cur, Tab::b21);
#pragma unroll
                      for (int c = 0; c < NDYN; ++c) y[c] = O::mad(f1, k1[c], x[c]); }
                    rhs<NA, STRICT, false>(q, gmask, ba_at(O::add(t, O::mul(cur, Tab::a2))), y, kk);
#pragma unroll
                    for (int c = 0; c < NDYN; ++c) k2[c] = kk[c];
                    // stage 3
                    { const double f1 = O::mul(cur, Tab::b31), f2 = O::mul(cur, Tab::b32);
#pragma unroll
                      for (int c = 0; c < NDYN; ++c) y[c] = O::mad(f2, k2[c], O::mad(f1, k1[c], x[c])); }
                    rhs<NA, STRICT, true>(q, gmask, ba_at(O::add(t, O::mul(cur, Tab::a3))), y, kk);
                    { const double g1 = O::mul(cur, Tab::c1), g3 = O::mul(cur, Tab::c3);
                      const double e1 = O::mul(cur, Tab::dc1), e3 = O::mul(cur, Tab::dc3);
#pragma unroll
                      for (int c = 0; c < NDYN; ++c) k3[c] = kk[c];
#pragma unroll
                      for (int c = 0; c < NPAS; ++c) {
                          accN[c] = O::mad(g3, kk[NDYN + c], O::mad(g1, k1[NDYN + c], x[NDYN + c]));
                          accE[c] = O::mad(e3, kk[NDYN + c], O::mul(e1, k1[NDYN + c]));
                      } }
                    // stage 4
                    { const double f1 = O::mul(cur, Tab::b41), f2 = O::mul(cur, Tab::b42), f3 = O::mul(cur, Tab::b43);
#pragma unroll
                      for (int c = 0; c < NDYN; ++c) y[c] = O::mad(f3, k3[c], O::mad(f2, k2[c], O::mad(f1, k1[c], x[c]))); }
                    rhs<NA, STRICT, true>(q, gmask, ba_at(O::add(t, O::mul(cur, Tab::a4))), y, kk);
                    { const double g4 = O::mul(cur, Tab::c4), e4 = O::mul(cur, Tab::dc4);
#pragma unroll
                      for (int c = 0; c < NDYN; ++c) k4[c] = kk[c];
#pragma unroll
                      for (int c = 0; c < NPAS; ++c) { accN[c] = O::mad(g4, kk[NDYN + c], accN[c]); accE[c] = O::mad(e4, kk[NDYN + c], accE[c]); } }
                    // stage 5
                    { const double f1 = O::mul(cur, Tab::b51), f2 = O::mul(cur, Tab::b52), f3 = O::mul(cur, Tab::b53),
                                   f4 = O::mul(cur, Tab::b54);
#pragma unroll
                      for (int c = 0; c < NDYN; ++c)
                          y[c] = O::mad(f4, k4[c], O::mad(f3, k3[c], O::mad(f2, k2[c], O::mad(f1, k1[c], x[c])))); }
                    rhs<NA, STRICT, true>(q, gmask, ba_at(O::add(t, O::mul(cur, Tab::a5))), y, kk);
                    { const double g5 = O::mul(cur, Tab::c5), e5 = O::mul(cur, Tab::dc5);
#pragma unroll
                      for (int c = 0; c < NDYN; ++c) k5[c] = kk[c];
#pragma unroll
                      for (int c = 0; c < NPAS; ++c) { accN[c] = O::mad(g5, kk[NDYN + c], accN[c]); accE[c] = O::mad(e5, kk[NDYN + c], accE[c]); } }
                    // stage 6
                    { const double f1 = O::mul(cur, Tab::b61), f2 = O::mul(cur, Tab::b62), f3 = O::mul(cur, Tab::b63),
                                   f4 = O::mul(cur, Tab::b64), f5 = O::mul(cur, Tab::b65);
#pragma unroll
                      for (int c = 0; c < NDYN; ++c)
                          y[c] = O::mad(f5, k5[c], O::mad(f4, k4[c], O::mad(f3, k3[c], O::mad(f2, k2[c], O::mad(f1, k1[c], x[c]))))); }
                    const double t_end = O::add(t, cur);
                    rhs<NA, STRICT, true>(q, gmask, ba_at(t_end), y, kk);
                    { const double g6 = O::mul(cur, Tab::c6), e6 = O::mul(cur, Tab::dc6);
#pragma unroll
                      for (int c = 0; c < NDYN; ++c) k6[c] = kk[c];
#pragma unroll
                      for (int c = 0; c < NPAS; ++c) { accN[c] = O::mad(g6, kk[NDYN + c], accN[c]); accE[c] = O::mad(e6, kk[NDYN + c], accE[c]); } }
                    // solution (dynamic part) and the FSAL derivative
                    double xn[NDYN];
                    { const double g1 = O::mul(cur, Tab::c1), g3 = O::mul(cur, Tab::c3), g4 = O::mul(cur, Tab::c4),
                                   g5 = O::mul(cur, Tab::c5), g6 = O::mul(cur, Tab::c6);
#pragma unroll
                      for (int c = 0; c < NDYN; ++c)
                          xn[c] = O::mad(g6, k6[c], O::mad(g5, k5[c], O::mad(g4, k4[c], O::mad(g3, k3[c], O::mad(g1, k1[c], x[c]))))); }
                    rhs<NA, STRICT, true>(q, gmask, ba_at(t_end), xn, kk);   // kk = k7 = dxdt_new
                    const int seg_end = s_stage;                            // segment of t + dt (for the next step)
                    // error estimate and its scaled max-norm (default_error_checker)
                    double err;
                    {
                        const double e1 = O::mul(cur, Tab::dc1), e3 = O::mul(cur, Tab::dc3), e4 = O::mul(cur, Tab::dc4),
                                     e5 = O::mul(cur, Tab::dc5), e6 = O::mul(cur, Tab::dc6), e7 = O::mul(cur, Tab::dc7);
                        double xe[NCOMP];
#pragma unroll
                        for (int c = 0; c < NDYN; ++c)
                            xe[c] = O::mad(e7, kk[c], O::mad(e6, k6[c], O::mad(e5, k5[c], O::mad(e4, k4[c], O::mad(e3, k3[c], O::mul(e1, k1[c]))))));
#pragma unroll
                        for (int c = 0; c < NPAS; ++c) xe[NDYN + c] = O::mad(e7, kk[NDYN + c], accE[c]);
                        if (STRICT) {
                            double m = 0.0;
#pragma unroll
                            for (int c = 0; c < NCOMP; ++c) {
                                const double den = O::add(f_abs, O::mul(f_rel, O::add(fabs(x[c]), O::mul(cur, fabs(k1[c])))));
                                const double v = fabs(__ddiv_rn(fabs(xe[c]), den));
                                m = (m < v) ? v : m;
                            }
                            err = group_max<NA>(gmask, m);
                        } else {
                            // arg-max by cross multiplication, then ONE exact division per lane
                            double bn = 0.0, bd = 1.0;
#pragma unroll
                            for (int c = 0; c < NCOMP; ++c) {
                                const double den = fma(f_rel, fma(cur, fabs(k1[c]), fabs(x[c])), f_abs);
                                const double num = fabs(xe[c]);
                                if (num * bd > bn * den) { bn = num; bd = den; }
                            }
                            err = group_max<NA>(gmask, bn / bd);
                        }
                    }
                    if (err > 1.0) {
                        // reject: decrease_step (error_order 4): dt *= max(0.9 * err^(-1/3), 1/5)
                        cur = O::mul(cur, std_max(O::mul(9.0 / 10.0, pow(err, -1.0 / 3.0)), 1.0 / 5.0));
                        ++n_rej;
                        dt = cur;
                        if (fail_steps++ >= 500) { status |= SEPAIHRD_ST_STEP_FAILURE; break; }
                    } else {
                        // accept: t += dt; increase_step (stepper_order 5) when err < 0.5
                        t = t_end;
                        if (STRICT || dt < kp.hmax) {
                            if (err < 0.5) {
                                const double e2 = std_max(3.2e-4 /* pow(5,-5) */, err);
                                cur = O::mul(cur, O::mul(9.0 / 10.0, pow(e2, -1.0 / 5.0)));
                            }
                            dt = std_max(dt, cur);   // max_abs: keep the larger of the carried and the proposed step
                        }
                        ++n_acc;
                        fail_steps = 0;
                        seg = seg_end;
#pragma unroll
                        for (int c = 0; c < NDYN; ++c) { x[c] = xn[c]; k1[c] = kk[c]; }
#pragma unroll
                        for (int c = 0; c < NPAS; ++c) { x[NDYN + c] = accN[c]; k1[NDYN + c] = kk[NDYN + c]; }
                    }
                }
                if (status & SEPAIHRD_ST_STEP_FAILURE) break;
            }
        }

        // ---- epilogue -------------------------------------------------------------------------------------
        if (MODE == MODE_LL) {
            double total;
            if (STRICT) total = O::add(O::add(ll_acc_h, ll_acc_i), ll_acc_d);   // ll_hosp + ll_icu + ll_deaths
            else total = group_sum<NA>(gmask, ll_acc_h);
            if (status != 0) total = -DBL_MAX;
            else if (isnan(total) || isinf(total)) { total = -DBL_MAX; status |= SEPAIHRD_ST_NONFINITE; }
            if (age == 0) {
                kp.out_ll[b] = total;
                if (kp.out_status) kp.out_status[b] = status;
                if (kp.out_steps) { kp.out_steps[2 * b] = n_acc; kp.out_steps[2 * b + 1] = n_rej; }
            }
        } else {
            if (status != 0) {   // failed sets: NaN-fill every row
                const double qnan = __longlong_as_double(0x7ff8000000000000LL);
                for (int r = 0; r < kp.traj_rows; ++r)
                    for (int w = age; w < W; w += NA) traj_out[(size_t)r * W + w] = qnan;
            }
            if (age == 0) {
                if (kp.out_status) kp.out_status[b] = status;
                if (kp.out_steps) { kp.out_steps[2 * b] = n_acc; kp.out_steps[2 * b + 1] = n_rej; }
            }
        }
    }
}

}  // namespace sepaihrd
