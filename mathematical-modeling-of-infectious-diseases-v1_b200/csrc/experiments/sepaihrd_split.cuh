// sepaihrd_split.cuh -- the fused likelihood kernel with every parameter set split over a PAIR OF WARPS by data flow.
//
// The SEPAIHRD right-hand side has a one-way structure: S, E, P, A, I form the infection loop (the force of infection reads
// P, A, I and drives S -> E -> P -> A / I), while H, ICU and the four passive compartments R, D, CumH, CumICU only CONSUME I and
// A and never feed back (reference src/model/AgeSEPAIHRDModel.cpp:198-226).  So the state of a lane (one age class of one
// parameter set) is cut in two:
//     UPSTREAM warp    S E P A I           5 of the 11 compartments, the pressure exchange, the step-size controller
//     DOWNSTREAM warp  H ICU R D CumH CumICU   6 compartments, the daily incidences and the Poisson log-likelihood
// Lane l of both warps works on the same (set, age).  Per Dopri5 stage the upstream warp posts two doubles per lane (I and A
// of the stage state) into shared memory and never waits for its partner inside an attempt; the downstream warp follows one
// stage behind.  Once per attempt they meet: downstream posts its share of the error test (any |xe| > den, the coarse
// log-ratio, three candidates of the arg-max tournament), upstream decides accept / reject and the next step and sends the
// verdict with the next command.  Hand-offs are mbarriers in shared memory (one elected lane arrives after a __syncwarp).
//
// Why: (1) latency -- an attempt's critical instruction stream shrinks from ~800 to ~400 instructions per warp and the Poisson
// terms leave the controller's path, which is what a launch of a few thousand sets (one warp per scheduler, nothing to overlap
// with) is bound by: single-chain Metropolis-Hastings, line searches, 512 chains per GPU (reference callers
// MetropolisHastingsSampler.cpp:283-384, HillClimbingOptimizer.cpp:58-103); (2) occupancy -- a lane carries half the state, so
// the kernel fits 170 registers and three warps per scheduler instead of two at 255.
//
// ARITHMETIC IS THE FAST KERNEL'S, OPERATION FOR OPERATION (sepaihrd_kernels.cuh, loop 6): same FMA chains per compartment,
// same controller, same tournament bracket, same Poisson terms -- results are bit-identical to sepaihrd_batch_kernel
// (tests/test_gpu_parity.py), so which kernel a batch runs on never shows in a log-likelihood or an accept decision.
// Scope: FAST arithmetic, log-likelihood mode, 4 age classes, schedule breakpoints on output-grid days (the ONGRID case).
//
// STATUS: EXPERIMENT, NOT SHIPPED (compiled only with -DSEPAIHRD_WITH_SPLIT).  Measured on a B200 (round 2,
// profiles/r02_split_kernel_experiment.txt): bit-identical to FAST on 65,536 mixed sets, but 0.62 ms against 0.51 ms for one
// set and 0.64x of FAST's throughput on 32 k sets.  The reason is in the data flow, not in the code: an attempt is a chain of
// seven DEPENDENT right-hand sides (stage s + 1 needs dE of stage s, which needs lambda of stage s: ~9 dependent FP64
// operations of 8.9 cycles each plus the pressure exchange per stage, >= 1,100 cycles per attempt), so a lone warp is bound by
// that chain and not by its issue slots; taking 45 % of the instructions off it buys nothing, and the hand-offs (six posts, one
// join and one command per attempt) add ~450 cycles.  With many warps the FAST kernel already overlaps two such chains per
// scheduler and its smaller instruction count wins.
#pragma once

#include "../sepaihrd_kernels.cuh"

namespace sepaihrd {

constexpr int UPN = 5;                    // S E P A I
enum : unsigned { CMD_BEGIN = 1u, CMD_COMMIT = 2u, CMD_OBSERVE = 4u, CMD_ATTEMPT = 8u, CMD_UNIT = 16u, CMD_FINISH = 32u, CMD_EXIT = 64u };

struct alignas(16) PairBox {              // one per warp pair, in shared memory
    uint64_t bar_cmd, bar_ack, bar_join, bar_stage[6];
    unsigned flags, commit_mask, big_mask;
    int obs_idx;
    long long tile;                        // CMD_BEGIN: first set of the tile is tile * 8
    double cur[32];                        // CMD_ATTEMPT: step length per lane (0 = the lane group idles)
    int fin_status[32], fin_acc[32], fin_rej[32];
    double sI[6][32], sA[6][32];           // I and A of the stage states 2..6 and of the new solution
    int lmax[32];
    double num[3][32], den[3][32];         // downstream's tournament candidates: component 5, winner(6, 7), winner(8, 9, 10)
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// mbarrier wait with a deadlock guard: a partner that never arrives traps the kernel (a CUDA error at the next
// synchronisation) instead of hanging the GPU.  The first probe succeeds in the steady state.
__device__ __forceinline__ void pair_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (unsigned it = 0;; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        if (it > (1u << 22)) __trap();
    }
}
// every lane has written its part of a message: publish it (one arrival per message)
__device__ __forceinline__ void post(uint64_t* bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

struct UpParams { double theta, sigma, gamma_p, gamma_A, kI, p, a; double M[4]; };
struct DnParams { double h, icu, kH, kU, gamma_A, gamma_I, gamma_H, gamma_ICU, dH, dICU, dcomm; };

// upstream half of rhs<4, false, ., true>: y = S E P A I
__device__ __forceinline__ void rhs_up(const UpParams& q, double* spi, int slot_base, int lane_in_block, const double (&y)[UPN], double (&d)[UPN]) {
    const double S = y[0], E = y[1], P = y[2], A = y[3], I = y[4];
    const double pressure = fma(q.theta, I, P + A);
    double pall[4];
    gather_pressure<4>(spi, slot_base, lane_in_block, pressure, pall);
    double acc[2];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j % 2] = (j < 2) ? q.M[j] * pall[j] : fma(q.M[j], pall[j], acc[j % 2]);
    double lam = acc[0] + acc[1];
    const unsigned hi = (unsigned)__double2hiint(lam);
    if (hi > 0x7ff00000u) lam = 0.0;
    const double flow_SE = lam * S;
    d[0] = -flow_SE;
    const double flow_P_out = q.gamma_p * P;
    const double flow_PA = q.p * flow_P_out;
    d[1] = fma(-q.sigma, E, flow_SE);
    d[2] = fma(q.sigma, E, -flow_P_out);
    d[3] = fma(-q.gamma_A, A, flow_PA);
    d[4] = fma(-q.kI, I, flow_P_out - flow_PA);
}
// downstream half: dyn = d(H ICU), pas = d(R D CumH CumICU)
template <bool PASSIVE>
__device__ __forceinline__ void rhs_dn(const DnParams& q, double I, double A, double H, double U, double (&d)[2], double (&pas)[NPAS]) {
    const double flow_IH = q.h * I;
    const double flow_H_ICU = q.icu * H;
    d[0] = fma(-q.kH, H, flow_IH);
    d[1] = fma(-q.kU, U, flow_H_ICU);
    if (PASSIVE) {
        pas[0] = fma(q.gamma_ICU, U, fma(q.gamma_H, H, fma(q.gamma_I, I, q.gamma_A * A)));
        pas[1] = fma(q.dcomm, I, fma(q.dICU, U, q.dH * H));
        pas[2] = flow_IH;
        pas[3] = flow_H_ICU;
    }
}

// One Dopri5 attempt, upstream half (dopri5_attempt<4, false, false, UNIT> restricted to S E P A I).
template <bool UNIT>
__device__ __forceinline__ void attempt_up(const UpParams& q, PairBox& box, int lane, double* spi, int& pi_slot, int pi_stride, int lane_in_block,
                                           double cur, double ecur, const double (&x)[UPN], const double (&k1)[UPN], double (&xn)[UPN],
                                           double (&k7)[UPN], double (&xe)[UPN], const double* hc) {
    auto cf = [&](int i) -> double { return UNIT ? hc[i] : cur * c_tab[i]; };
    auto ef = [&](int i) -> double { return UNIT ? hc[i] : ecur * c_tab[i]; };
    auto next_slot = [&]() -> int { pi_slot ^= pi_stride; return pi_slot; };
    auto send = [&](int s, const double (&y)[UPN]) { box.sI[s][lane] = y[4]; box.sA[s][lane] = y[3]; post(&box.bar_stage[s], lane); };
    double k2[UPN], k3[UPN], k4[UPN], k5[UPN], k6[UPN], y[UPN];
    { const double f1 = cf(T_B21);
#pragma unroll
      for (int c = 0; c < UPN; ++c) y[c] = fma(f1, k1[c], x[c]); }
    send(0, y);
    rhs_up(q, spi, next_slot(), lane_in_block, y, k2);
    { const double f1 = cf(T_B31), f2 = cf(T_B32);
#pragma unroll
      for (int c = 0; c < UPN; ++c) y[c] = fma(f2, k2[c], fma(f1, k1[c], x[c])); }
    send(1, y);
    rhs_up(q, spi, next_slot(), lane_in_block, y, k3);
    { const double f1 = cf(T_B41), f2 = cf(T_B42), f3 = cf(T_B43);
#pragma unroll
      for (int c = 0; c < UPN; ++c) y[c] = fma(f3, k3[c], fma(f2, k2[c], fma(f1, k1[c], x[c]))); }
    send(2, y);
    rhs_up(q, spi, next_slot(), lane_in_block, y, k4);
    { const double f1 = cf(T_B51), f2 = cf(T_B52), f3 = cf(T_B53), f4 = cf(T_B54);
#pragma unroll
      for (int c = 0; c < UPN; ++c) y[c] = fma(f4, k4[c], fma(f3, k3[c], fma(f2, k2[c], fma(f1, k1[c], x[c])))); }
    send(3, y);
    rhs_up(q, spi, next_slot(), lane_in_block, y, k5);
    { const double f1 = cf(T_B61), f2 = cf(T_B62), f3 = cf(T_B63), f4 = cf(T_B64), f5 = cf(T_B65);
#pragma unroll
      for (int c = 0; c < UPN; ++c) y[c] = fma(f5, k5[c], fma(f4, k4[c], fma(f3, k3[c], fma(f2, k2[c], fma(f1, k1[c], x[c]))))); }
    send(4, y);
    rhs_up(q, spi, next_slot(), lane_in_block, y, k6);
    { const double g1 = cf(T_C1), g3 = cf(T_C3), g4 = cf(T_C4), g5 = cf(T_C5), g6 = cf(T_C6);
#pragma unroll
      for (int c = 0; c < UPN; ++c) xn[c] = fma(g6, k6[c], fma(g5, k5[c], fma(g4, k4[c], fma(g3, k3[c], fma(g1, k1[c], x[c]))))); }
    send(5, xn);
    rhs_up(q, spi, next_slot(), lane_in_block, xn, k7);
    { const double e1 = ef(T_DC1), e3 = ef(T_DC3), e4 = ef(T_DC4), e5 = ef(T_DC5), e6 = ef(T_DC6), e7 = ef(T_DC7);
#pragma unroll
      for (int c = 0; c < UPN; ++c) xe[c] = fma(e7, k7[c], fma(e6, k6[c], fma(e5, k5[c], fma(e4, k4[c], fma(e3, k3[c], e1 * k1[c]))))); }
}

// One Dopri5 attempt, downstream half: x = H ICU | R D CumH CumICU (indices 0..1 dynamic, 2..5 passive).
template <bool UNIT>
__device__ __forceinline__ void attempt_dn(const DnParams& q, PairBox& box, int lane, unsigned par, double cur, double ecur, const double (&x)[6],
                                           const double (&k1)[6], double (&xn)[6], double (&k7)[6], double (&xe)[6], const double* hc) {
    auto cf = [&](int i) -> double { return UNIT ? hc[i] : cur * c_tab[i]; };
    auto ef = [&](int i) -> double { return UNIT ? hc[i] : ecur * c_tab[i]; };
    double k2[2], k3[2], k4[2], k5[2], k6[2], y[2], kp_[NPAS], accN[NPAS], accE[NPAS];
    { const double f1 = cf(T_B21);
#pragma unroll
      for (int c = 0; c < 2; ++c) y[c] = fma(f1, k1[c], x[c]); }
    pair_wait(&box.bar_stage[0], par);
    rhs_dn<false>(q, box.sI[0][lane], box.sA[0][lane], y[0], y[1], k2, kp_);
    { const double f1 = cf(T_B31), f2 = cf(T_B32);
#pragma unroll
      for (int c = 0; c < 2; ++c) y[c] = fma(f2, k2[c], fma(f1, k1[c], x[c])); }
    pair_wait(&box.bar_stage[1], par);
    rhs_dn<true>(q, box.sI[1][lane], box.sA[1][lane], y[0], y[1], k3, kp_);
    { const double g1 = cf(T_C1), g3 = cf(T_C3), e1 = ef(T_DC1), e3 = ef(T_DC3);
#pragma unroll
      for (int c = 0; c < NPAS; ++c) {
          accN[c] = fma(g3, kp_[c], fma(g1, k1[2 + c], x[2 + c]));
          accE[c] = fma(e3, kp_[c], e1 * k1[2 + c]);
      } }
    { const double f1 = cf(T_B41), f2 = cf(T_B42), f3 = cf(T_B43);
#pragma unroll
      for (int c = 0; c < 2; ++c) y[c] = fma(f3, k3[c], fma(f2, k2[c], fma(f1, k1[c], x[c]))); }
    pair_wait(&box.bar_stage[2], par);
    rhs_dn<true>(q, box.sI[2][lane], box.sA[2][lane], y[0], y[1], k4, kp_);
    { const double g4 = cf(T_C4), e4 = ef(T_DC4);
#pragma unroll
      for (int c = 0; c < NPAS; ++c) { accN[c] = fma(g4, kp_[c], accN[c]); accE[c] = fma(e4, kp_[c], accE[c]); } }
    { const double f1 = cf(T_B51), f2 = cf(T_B52), f3 = cf(T_B53), f4 = cf(T_B54);
#pragma unroll
      for (int c = 0; c < 2; ++c) y[c] = fma(f4, k4[c], fma(f3, k3[c], fma(f2, k2[c], fma(f1, k1[c], x[c])))); }
    pair_wait(&box.bar_stage[3], par);
    rhs_dn<true>(q, box.sI[3][lane], box.sA[3][lane], y[0], y[1], k5, kp_);
    { const double g5 = cf(T_C5), e5 = ef(T_DC5);
#pragma unroll
      for (int c = 0; c < NPAS; ++c) { accN[c] = fma(g5, kp_[c], accN[c]); accE[c] = fma(e5, kp_[c], accE[c]); } }
    { const double f1 = cf(T_B61), f2 = cf(T_B62), f3 = cf(T_B63), f4 = cf(T_B64), f5 = cf(T_B65);
#pragma unroll
      for (int c = 0; c < 2; ++c) y[c] = fma(f5, k5[c], fma(f4, k4[c], fma(f3, k3[c], fma(f2, k2[c], fma(f1, k1[c], x[c]))))); }
    pair_wait(&box.bar_stage[4], par);
    rhs_dn<true>(q, box.sI[4][lane], box.sA[4][lane], y[0], y[1], k6, kp_);
    { const double g6 = cf(T_C6), e6 = ef(T_DC6);
#pragma unroll
      for (int c = 0; c < NPAS; ++c) { accN[c] = fma(g6, kp_[c], accN[c]); accE[c] = fma(e6, kp_[c], accE[c]); } }
    { const double g1 = cf(T_C1), g3 = cf(T_C3), g4 = cf(T_C4), g5 = cf(T_C5), g6 = cf(T_C6);
#pragma unroll
      for (int c = 0; c < 2; ++c) xn[c] = fma(g6, k6[c], fma(g5, k5[c], fma(g4, k4[c], fma(g3, k3[c], fma(g1, k1[c], x[c]))))); }
    double k7d[2], k7p[NPAS];
    pair_wait(&box.bar_stage[5], par);
    rhs_dn<true>(q, box.sI[5][lane], box.sA[5][lane], xn[0], xn[1], k7d, k7p);
    { const double e1 = ef(T_DC1), e3 = ef(T_DC3), e4 = ef(T_DC4), e5 = ef(T_DC5), e6 = ef(T_DC6), e7 = ef(T_DC7);
#pragma unroll
      for (int c = 0; c < 2; ++c) xe[c] = fma(e7, k7d[c], fma(e6, k6[c], fma(e5, k5[c], fma(e4, k4[c], fma(e3, k3[c], e1 * k1[c])))));
#pragma unroll
      for (int c = 0; c < NPAS; ++c) xe[2 + c] = fma(e7, k7p[c], accE[c]); }
#pragma unroll
    for (int c = 0; c < 2; ++c) k7[c] = k7d[c];
#pragma unroll
    for (int c = 0; c < NPAS; ++c) { xn[2 + c] = accN[c]; k7[2 + c] = k7p[c]; }
}

// cross-multiplication comparison of the tournament: is b's ratio larger than a's?  (kernels.cuh: num[c + s] * den[c] > num[c] * den[c + s])
__device__ __forceinline__ void take_larger(double& na, double& da, double nb, double db) {
    const bool other = nb * da > na * db;
    na = other ? nb : na;
    da = other ? db : da;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) sepaihrd_split_kernel(const KParams kp) {
    constexpr int WARPS = THREADS / 32, PAIRS = WARPS / 2, SETS = PAIRS * 8;
    constexpr int n = 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sblob = reinterpret_cast<double*>(smem_raw);
    const int blob_doubles = kp.blob_bytes >> 3;
    double* sslots = sblob + blob_doubles;
    double* sbeff = sslots + SETS * kp.slot_stride;
    double* spi = sbeff + SETS * ((kp.seg_stride + 1) & ~1);
    double* smb = spi + 2 * THREADS;
    PairBox* boxes = reinterpret_cast<PairBox*>(smb + 4 * THREADS);
    uint64_t* bar = reinterpret_cast<uint64_t*>(boxes + PAIRS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, pair = warp >> 1;
    const bool upstream = (warp & 1) == 0;
    PairBox& box = boxes[pair];
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (uint32_t)kp.blob_bytes);
        tma_bulk_g2s(sblob, kp.blob, (uint32_t)kp.blob_bytes, bar);
    }
    if (upstream && lane == 0) {
        mbar_init(&box.bar_cmd, 1); mbar_init(&box.bar_ack, 1); mbar_init(&box.bar_join, 1);
#pragma unroll
        for (int s = 0; s < 6; ++s) mbar_init(&box.bar_stage[s], 1);
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
    const int age = lane & 3, grp_in_warp = lane >> 2;
    const int nseg = kp.nseg, K = kp.K;
    const double hmax = kp.hmax;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double* my_slots = sslots + (pair * 8 + grp_in_warp) * kp.slot_stride;
    double* my_beff = sbeff + (pair * 8 + grp_in_warp) * kp.seg_stride;
    const int sl_kappa0 = kp.nb, sl_scal0 = kp.nb + kp.nk, sl_age0 = sl_scal0 + 7;
    const int sl_mult0 = sl_age0 + 8 * n, sl_seed = sl_mult0 + 8, sl_runup = sl_mult0 + 9, sl_beta = sl_mult0 + 10;
    const unsigned lane_bit = 1u << lane;
    unsigned cmd_par = 0, ack_par = 0, att_par = 0;

    if (upstream) {
        // ================================ UPSTREAM: S E P A I, controller, schedule ================================
        int pi_slot = 0;
        const int lib = threadIdx.x;                 // this lane's cell of the pressure exchange buffers
        bool ack_outstanding = false;
        auto send_cmd = [&](unsigned flags, int obs_idx, unsigned commit_mask, double cur_lane) {
            if (ack_outstanding) { pair_wait(&box.bar_ack, ack_par); ack_par ^= 1u; ack_outstanding = false; }
            if (lane == 0) { box.flags = flags; box.obs_idx = obs_idx; box.commit_mask = commit_mask; }
            if (flags & CMD_ATTEMPT) box.cur[lane] = cur_lane;
            post(&box.bar_cmd, lane);
            if (!(flags & CMD_ATTEMPT)) ack_outstanding = true;
        };
        while (true) {
            unsigned wt = 0;
            if (lane == 0) wt = atomicAdd(kp.tile_counter, 1u);
            wt = __shfl_sync(FULL, wt, 0);
            if ((long long)wt >= kp.tiles) {
                send_cmd(CMD_EXIT, 0, 0u, 0.0);
                break;
            }
            const long long b_raw = (long long)wt * 8 + grp_in_warp;
            const bool have = b_raw < kp.B;
            const long long b = have ? b_raw : (kp.B - 1);
            // ---- updateModelParameters (as in sepaihrd_batch_kernel) ----------------------------------------------
            if (ack_outstanding) { pair_wait(&box.bar_ack, ack_par); ack_par ^= 1u; ack_outstanding = false; }   // the partner has left the previous tile's slots
            __syncwarp();
            for (int s = age; s < kp.nslots; s += n) my_slots[s] = sblob[kp.o_base + s];
            __syncwarp();
            bool kappa_touched = false;
            {
                const double* prow = kp.params + b * kp.ld;
                for (int i = age; i < kp.P; i += n) {
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
            const unsigned gm = 0xfu << (lane - age);
            {
                bool neg = false;
                for (int k = 1 + age; k < kp.nk; k += n) neg |= (my_slots[sl_kappa0 + k] < 0.0);
                const bool any_touch = (__ballot_sync(FULL, kappa_touched) & gm) != 0;
                const bool any_neg = (__ballot_sync(FULL, neg) & gm) != 0;
                if (any_touch && any_neg) status |= SEPAIHRD_ST_INVALID_PARAM;
            }
            for (int s = age; s <= nseg; s += n) {
                const double bv = (kp.nb > 0) ? my_slots[s_segb[s]] : my_slots[sl_beta];
                my_beff[s] = bv * my_slots[sl_kappa0 + s_segk[s]];
            }
            __syncwarp();
            if (lane == 0) box.tile = (long long)wt;
            send_cmd(CMD_BEGIN, 0, 0u, 0.0);                       // the slot vectors are ready: the partner loads its rates

            UpParams q;
            q.theta = my_slots[sl_scal0 + 0]; q.sigma = my_slots[sl_scal0 + 1]; q.gamma_p = my_slots[sl_scal0 + 2];
            q.gamma_A = my_slots[sl_scal0 + 3];
            const double gamma_I = my_slots[sl_scal0 + 4];
            q.a = my_slots[sl_age0 + 0 * n + age];
            const double hinf = my_slots[sl_age0 + 1 * n + age];
            q.p = my_slots[sl_age0 + 2 * n + age];
            const double h_ = my_slots[sl_age0 + 3 * n + age], dcomm = my_slots[sl_age0 + 7 * n + age];
            const double invN = sblob[kp.o_invN + age];
            const double hN = hinf * invN;
            q.kI = gamma_I + h_ + dcomm;
            double Mrow[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) Mrow[j] = sblob[kp.o_M + j * n + age];
#pragma unroll
            for (int j = 0; j < 4; ++j) smb[j * THREADS + threadIdx.x] = Mrow[j] * __shfl_sync(FULL, hN, j, 4);
            int mseg = -1;
            auto refold = [&](int s) {
                const double f = my_beff[s] * q.a;
#pragma unroll
                for (int j = 0; j < 4; ++j) q.M[j] = smb[j * THREADS + threadIdx.x] * f;
                mseg = s;
            };
            // ---- initial state (ObjectiveFunction.cpp:124-163): this warp keeps S E P A I, but S needs the whole sum -----
            double x[UPN], k1[UPN];
            const double popN = sblob[kp.o_pop + age];
            {
                double xa[9];
                const double runup_days = my_slots[sl_runup], seed_exposed = my_slots[sl_seed];
                if (runup_days > 0 && seed_exposed > 0) {
                    xa[1] = seed_exposed * sblob[kp.o_agefrac + age];
#pragma unroll
                    for (int c = 2; c < 9; ++c) xa[c] = 0.0;
                } else {
#pragma unroll
                    for (int c = 1; c <= 8; ++c) xa[c] = sblob[kp.o_init + c * n + age] * my_slots[sl_mult0 + c - 1];
                }
                double sum = 0;
#pragma unroll
                for (int j = 1; j < 9; ++j) sum = sum + xa[j];
                const bool over = (__ballot_sync(FULL, sum > popN) & gm) != 0;
                if (over && status == 0) status |= SEPAIHRD_ST_S_OVERFLOW;
                x[0] = popN - sum;
#pragma unroll
                for (int c = 1; c < UPN; ++c) x[c] = xa[c];
            }
            int n_acc = 0, n_rej = 0;
            bool alive = (status == 0);
            double dt = kp.dt_hint;
            double t = s_times[0];
            int seg = 0;
            while (seg < nseg && t > s_bp[seg]) ++seg;
            double bp_next = (seg < nseg) ? s_bp[seg] : INF;
            double ba = my_beff[seg] * q.a;
            (void)ba;
            {
                pi_slot ^= THREADS;
                refold(seg);
                rhs_up(q, spi, pi_slot, lib, x, k1);
            }
            int fail_steps = 0;
            int obs_pending = -1;
            unsigned commit_pending = 0;            // lanes whose last attempt was accepted and not yet committed downstream
            bool have_commit = false;
            for (int idx = 0; idx < K; ++idx) {
                t = s_times[idx];
                if (obs_pending >= 0) {              // a grid point without an attempt after it (degenerate interval)
                    send_cmd((have_commit ? CMD_COMMIT : 0u) | CMD_OBSERVE, obs_pending, commit_pending, 0.0);
                    have_commit = false; commit_pending = 0;
                }
                obs_pending = idx;
                if (idx + 1 == K) break;
                const double t_next = s_times[idx + 1];
                double rem = t_next - t;
                bool active = alive && (rem > DBL_EPSILON);
                unsigned m_active = __ballot_sync(FULL, active);
                if (m_active == 0) {
                    if (!__any_sync(FULL, alive)) break;
                    continue;
                }
                const bool day_bp = __any_sync(FULL, bp_next < t_next);
                bool first_of_day = true;
                while (true) {
                    const double cur = active ? std_min(dt, rem) : 0.0;
                    const double t_end = t + cur;
                    int s_hi = seg;
                    if (day_bp) {
                        if (__any_sync(FULL, !(t_end <= bp_next))) {
                            if (!(t_end <= bp_next)) {
                                const double t2 = fma(cur, c_tab[T_A2], t);
                                int s_lo = seg;
                                while (s_lo < nseg && t2 > s_bp[s_lo]) ++s_lo;
                                s_hi = s_lo;
                                while (s_hi < nseg && t_end > s_bp[s_hi]) ++s_hi;
                                if (s_lo != mseg) refold(s_lo);
                            }
                        }
                    }
                    double xn[UPN], k7[UPN], xe[UPN];
                    const double ecur = cur * kp.inv_rel;
                    const bool unit = first_of_day && !day_bp && __all_sync(FULL, !active || cur == hmax);
                    first_of_day = false;
                    send_cmd(CMD_ATTEMPT | (unit ? CMD_UNIT : 0u) | (have_commit ? CMD_COMMIT : 0u) | (obs_pending >= 0 ? CMD_OBSERVE : 0u),
                             obs_pending, commit_pending, cur);
                    have_commit = false; commit_pending = 0; obs_pending = -1;
                    if (unit) attempt_up<true>(q, box, lane, spi, pi_slot, THREADS, lib, cur, ecur, x, k1, xn, k7, xe, kp.hc);
                    else attempt_up<false>(q, box, lane, spi, pi_slot, THREADS, lib, cur, ecur, x, k1, xn, k7, xe, kp.hc);
                    // ---- decision (loop 6 of sepaihrd_batch_kernel), with the downstream half's share joined in ---------
                    double num[UPN], den[UPN];
                    bool big = false;
#pragma unroll
                    for (int c = 0; c < UPN; ++c) {
                        den[c] = fma(cur, fabs(k1[c]), fabs(x[c])) + kp.abs_over_rel;
                        big |= fabs(xe[c]) > den[c];
                    }
                    const unsigned m_more = __ballot_sync(FULL, active && ((t_next - t_end) > DBL_EPSILON));
                    const unsigned bal_up = __ballot_sync(FULL, big);
                    pair_wait(&box.bar_join, att_par);
                    att_par ^= 1u;
                    const unsigned bal_big = bal_up | box.big_mask;
                    if (unit && ((bal_big & m_active) | m_more) == 0) {
                        if (active) {
                            t = t_end;
                            ++n_acc;
                            fail_steps = 0;
#pragma unroll
                            for (int c = 0; c < UPN; ++c) { x[c] = xn[c]; k1[c] = k7[c]; }
                        }
                        commit_pending = m_active; have_commit = true;
                        break;
                    }
                    int lmax = box.lmax[lane];
#pragma unroll
                    for (int c = 0; c < UPN; ++c) num[c] = __hiloint2double(__double2hiint(xe[c]) & 0x7fffffff, __double2loint(xe[c]));
#pragma unroll
                    for (int c = 0; c < UPN; ++c) lmax = max(lmax, __double2hiint(num[c]) - __double2hiint(den[c]));
                    const unsigned m_low = __ballot_sync(FULL, dt < hmax);
                    const unsigned bal_ng = __ballot_sync(FULL, lmax > kp.thr_nogrow);
                    const unsigned bal_ns = __ballot_sync(FULL, !(lmax < kp.thr_small));
                    const unsigned bal_sb = __ballot_sync(FULL, lmax > kp.thr_big);
                    const unsigned g_rej = expand_groups<4>(bal_big) & m_active;
                    const unsigned g_ng = expand_groups<4>(bal_ng), g_ns = expand_groups<4>(bal_ns), g_sb = expand_groups<4>(bal_sb);
                    const unsigned m_val = (g_rej & ~g_sb) | (m_active & ~g_rej & m_low & g_ns & ~g_ng);
                    const bool reject = (g_rej & lane_bit) != 0;
                    double err = 0.0, facv = 0.0;
                    if (m_val != 0) {
                        // the bracket of sepaihrd_batch_kernel's tournament over components 0..10: (0,1) (2,3) (4,5) (6,7) (8,9) |
                        // (0,2) (4,6) (8,10) | (0,4) | (0,8); components 5..10 live downstream, which sends 5, w(6,7), w(8,9,10)
                        take_larger(num[0], den[0], num[1], den[1]);
                        take_larger(num[2], den[2], num[3], den[3]);
                        take_larger(num[4], den[4], box.num[0][lane], box.den[0][lane]);
                        take_larger(num[0], den[0], num[2], den[2]);
                        take_larger(num[4], den[4], box.num[1][lane], box.den[1][lane]);
                        take_larger(num[0], den[0], num[4], den[4]);
                        take_larger(num[0], den[0], box.num[2][lane], box.den[2][lane]);
                        err = group_max<4>(fast_div_pos(num[0], den[0]));
                        facv = 0.9 * pow_neg_inv(reject ? err : std_max(3.2e-4, err), reject);
                    }
                    unsigned m_dead = 0;
                    if (reject) {
                        const double shrink = ((g_sb & lane_bit) || err > 128.0) ? 0.2 : std_max(facv, 0.2);
                        dt = cur * shrink;
                        ++n_rej;
                        if (fail_steps++ >= 500) { status |= SEPAIHRD_ST_STEP_FAILURE; alive = false; }
                    } else if (active) {
                        t = t_end;
                        rem = t_next - t;
                        if (dt < hmax) {
                            double g = 0.0;
                            if (!(g_ns & lane_bit)) g = kp.grow_max;
                            else if (!(g_ng & lane_bit) && err < 0.5) g = facv;
                            dt = std_max(dt, cur * g);
                        }
                        ++n_acc;
                        fail_steps = 0;
                        if (s_hi != seg) {
                            seg = s_hi;
                            bp_next = (seg < nseg) ? s_bp[seg] : INF;
                            if (seg != mseg) refold(seg);
                        }
#pragma unroll
                        for (int c = 0; c < UPN; ++c) { x[c] = xn[c]; k1[c] = k7[c]; }
                    }
                    commit_pending = m_active & ~g_rej; have_commit = true;
                    if (g_rej != 0) m_dead = __ballot_sync(FULL, !alive);
                    m_active = (m_more | g_rej) & ~m_dead;
                    if (m_active == 0) break;
                    active = (m_active & lane_bit) != 0;
                }
            }
            // ---- the tile is done: last verdict, last grid point, status and step counts go downstream ----------------
            if (ack_outstanding) { pair_wait(&box.bar_ack, ack_par); ack_par ^= 1u; ack_outstanding = false; }
            box.fin_status[lane] = (int)status; box.fin_acc[lane] = n_acc; box.fin_rej[lane] = n_rej;
            send_cmd(CMD_FINISH | (have_commit ? CMD_COMMIT : 0u) | (obs_pending >= 0 ? CMD_OBSERVE : 0u), obs_pending, commit_pending, 0.0);
        }
    } else {
        // ================================ DOWNSTREAM: H ICU R D CumH CumICU, observer, likelihood ====================
        DnParams q{};
        double x[6], k1[6], xn[6], k7[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) { x[c] = 0.0; k1[c] = 0.0; xn[c] = 0.0; k7[c] = 0.0; }
        double prev_h = 0.0, prev_i = 0.0, prev_d = 0.0, ll_acc = 0.0;
        bool bad = false, have = false;
        long long b = 0;
        while (true) {
            pair_wait(&box.bar_cmd, cmd_par);
            cmd_par ^= 1u;
            const unsigned flags = box.flags;
            if (flags & CMD_EXIT) break;
            if (flags & CMD_BEGIN) {
                const long long b_raw = box.tile * 8 + grp_in_warp;
                have = b_raw < kp.B;
                b = have ? b_raw : (kp.B - 1);
                q.gamma_A = my_slots[sl_scal0 + 3]; q.gamma_I = my_slots[sl_scal0 + 4]; q.gamma_H = my_slots[sl_scal0 + 5];
                q.gamma_ICU = my_slots[sl_scal0 + 6];
                q.h = my_slots[sl_age0 + 3 * n + age]; q.icu = my_slots[sl_age0 + 4 * n + age];
                q.dH = my_slots[sl_age0 + 5 * n + age]; q.dICU = my_slots[sl_age0 + 6 * n + age]; q.dcomm = my_slots[sl_age0 + 7 * n + age];
                q.kH = q.gamma_H + q.dH + q.icu;
                q.kU = q.gamma_ICU + q.dICU;
                // initial H ICU R D CumH CumICU, and the I, A the first derivative needs
                double I0, A0;
                const double runup_days = my_slots[sl_runup], seed_exposed = my_slots[sl_seed];
                if (runup_days > 0 && seed_exposed > 0) {
#pragma unroll
                    for (int c = 0; c < 6; ++c) x[c] = 0.0;
                    I0 = 0.0; A0 = 0.0;
                } else {
                    A0 = sblob[kp.o_init + 3 * n + age] * my_slots[sl_mult0 + 2];
                    I0 = sblob[kp.o_init + 4 * n + age] * my_slots[sl_mult0 + 3];
#pragma unroll
                    for (int c = 5; c <= 8; ++c) x[c - 5] = sblob[kp.o_init + c * n + age] * my_slots[sl_mult0 + c - 1];
                    x[4] = sblob[kp.o_init + 9 * n + age];
                    x[5] = sblob[kp.o_init + 10 * n + age];
                }
                double d0[2], p0[NPAS];
                rhs_dn<true>(q, I0, A0, x[0], x[1], d0, p0);
                k1[0] = d0[0]; k1[1] = d0[1];
#pragma unroll
                for (int c = 0; c < NPAS; ++c) k1[2 + c] = p0[c];
                prev_h = x[4]; prev_i = x[5]; prev_d = x[3];
                ll_acc = 0.0; bad = false;
            }
            if ((flags & CMD_COMMIT) && (box.commit_mask & lane_bit)) {
#pragma unroll
                for (int c = 0; c < 6; ++c) { x[c] = xn[c]; k1[c] = k7[c]; }
            }
            if (flags & CMD_OBSERVE) {
                const int idx = box.obs_idx;
                const double inc_h = std_max(x[4] - prev_h, 0.0);
                const double inc_i = std_max(x[5] - prev_i, 0.0);
                const double inc_d = std_max(x[3] - prev_d, 0.0);
                prev_h = x[4]; prev_i = x[5]; prev_d = x[3];
                const int r = idx - kp.runup_offset;
                if (r >= 0) {
                    const double oh = s_obs_h[r * n + age], oi = s_obs_i[r * n + age], od = s_obs_d[r * n + age];
                    const bool vh = (oh >= 0.0), vi = (oi >= 0.0), vd = (od >= 0.0);
                    const double sh = inc_h + 1e-10, si = inc_i + 1e-10, sd = inc_d + 1e-10;
                    const double th = fma(oh, fast_log_nb(sh, s_logtab, vh, bad), -sh);
                    const double ti = fma(oi, fast_log_nb(si, s_logtab, vi, bad), -si);
                    const double td = fma(od, fast_log_nb(sd, s_logtab, vd, bad), -sd);
                    ll_acc += ((vh ? th : 0.0) + (vi ? ti : 0.0)) + (vd ? td : 0.0);
                }
            }
            if (flags & CMD_ATTEMPT) {
                const double cur = box.cur[lane];
                const double ecur = cur * kp.inv_rel;
                double xe[6];
                if (flags & CMD_UNIT) attempt_dn<true>(q, box, lane, att_par, cur, ecur, x, k1, xn, k7, xe, kp.hc);
                else attempt_dn<false>(q, box, lane, att_par, cur, ecur, x, k1, xn, k7, xe, kp.hc);
                att_par ^= 1u;
                double num[6], den[6];
                bool big = false;
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    den[c] = fma(cur, fabs(k1[c]), fabs(x[c])) + kp.abs_over_rel;
                    big |= fabs(xe[c]) > den[c];
                }
                int lmax = INT_MIN;
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    num[c] = __hiloint2double(__double2hiint(xe[c]) & 0x7fffffff, __double2loint(xe[c]));
                    lmax = max(lmax, __double2hiint(num[c]) - __double2hiint(den[c]));
                }
                const unsigned bal = __ballot_sync(FULL, big);
                // local rounds of the tournament: components 5 | (6, 7) | ((8, 9), 10) in the numbering of sepaihrd_batch_kernel
                take_larger(num[1], den[1], num[2], den[2]);
                take_larger(num[3], den[3], num[4], den[4]);
                take_larger(num[3], den[3], num[5], den[5]);
                box.num[0][lane] = num[0]; box.den[0][lane] = den[0];
                box.num[1][lane] = num[1]; box.den[1][lane] = den[1];
                box.num[2][lane] = num[3]; box.den[2][lane] = den[3];
                box.lmax[lane] = lmax;
                if (lane == 0) box.big_mask = bal;
                post(&box.bar_join, lane);
            } else {
                if (flags & CMD_FINISH) {
                    double total = group_sum<4>(bad ? __longlong_as_double(0x7ff8000000000000LL) : ll_acc);
                    unsigned status = (unsigned)box.fin_status[lane];
                    if (status != 0) total = -DBL_MAX;
                    else if (isnan(total) || isinf(total)) { total = -DBL_MAX; status |= SEPAIHRD_ST_NONFINITE; }
                    if (have && age == 0) {
                        kp.out_ll[b] = total;
                        if (kp.out_status) kp.out_status[b] = status;
                        if (kp.out_steps) { kp.out_steps[2 * b] = box.fin_acc[lane]; kp.out_steps[2 * b + 1] = box.fin_rej[lane]; }
                    }
                }
                post(&box.bar_ack, lane);       // the command (and, for CMD_BEGIN, the slot vectors) have been read
            }
        }
    }
}

}  // namespace sepaihrd
