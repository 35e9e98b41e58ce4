// sepaihrd_ppc.cu -- posterior-predictive aggregation on the device: the consumer of the trajectory kernel.
//
// Replaces ResultAggregator::aggregatePosteriorPredictives (reference src/model/ResultAggregator.cpp:174-412): for B
// posterior draws, simulate from ONE fixed initial state (quirk Q9), form the daily incidence of hospitalisations, ICU
// admissions and deaths on the output days t >= 0 (first difference of CumH / CumICU / D against the previous output day,
// or against the initial state for the first one; clamped at 0: .cpp:292-335) and their running sums ("cumulative from
// flows", .cpp:337-351), then reduce every (series, day, age) column over the draws to quantiles.
//
// Deliberate, documented difference: the reference feeds the draws through Boost's extended P-square streaming
// estimator (.cpp:28-33), whose result depends on the order of the draws; here every quantile is the EXACT sample
// quantile with linear interpolation between order statistics (numpy's default).
//
// Data flow (HBM-bound byte work -- no tensor cores): the trajectory kernel itself forms the six series while it integrates
// (TRAJ_PPC_SERIES: the incidences and running sums are three subtractions and three additions per output day on values it
// holds in registers anyway) and writes them draws-fastest, out[6][T][n][B]; ppc_select_kernel then SELECTS, per (series, day,
// age) column, the <= 2Q order statistics the quantiles need -- 8 probabilities per sweep group, any number of probabilities
// in groups -- and writes [6][T][n][Q].  Round 1 wrote D / CumH / CumICU trajectories (3.1 GB for 100 k draws x 4 ages), read
// them back in a series kernel that wrote 5.9 GB, and fell back to cub's segmented radix sort above 8 probabilities; now the
// trajectories are never materialised, the series are written once and no library sort is left.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <vector>

#include "sepaihrd_internal.h"

namespace {

#define PPC_TRY(expr)                                                                                        \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess) { cleanup(); return sepaihrd_internal::fail_with(SEPAIHRD_ERR_CUDA, cudaGetErrorString(e__)); } \
    } while (0)

// ---- exact order statistics without sorting -------------------------------------------------------------------------------
// A column needs the values at <= 2Q ranks (the two neighbours of every quantile position), not its full order.  One block per
// column: a first sweep finds the count, minimum and maximum of the non-NaN keys (everything above their common bit prefix is
// already decided); then most-significant-digit radix steps of 8 bits narrow, for ALL wanted ranks at once, the bucket each rank
// lies in (one 256-bin histogram per distinct bucket prefix, <= 16 of them), until every bucket holds at most SEL_CAP keys or
// the bits are used up; the survivors are gathered into shared memory and the wanted rank is picked by counting.  A column is
// read 3-5 times (its later sweeps mostly from L2) instead of being read and written 8 times by the radix sort.
// Keys: the usual order-preserving map of a double's bits to an unsigned integer; NaNs (failed draws) are left out.
constexpr int SEL_MAXT = 16, SEL_CAP = 192;
constexpr int PPC_DEFAULT_GRID = 4 * 148, PPC_DEFAULT_THREADS = 1024;

__device__ __forceinline__ unsigned long long sel_key(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
// SEL_UNROLL coalesced loads issued together (the sweeps are latency-bound otherwise); past the end: NaN, which every sweep skips
constexpr int SEL_UNROLL = 4;
template <int SEL_THREADS>
__device__ __forceinline__ void sel_load(const double* __restrict__ v, long long i0, long long B, double (&xs)[SEL_UNROLL]) {
#pragma unroll
    for (int u = 0; u < SEL_UNROLL; ++u) {
        const long long i = i0 + (long long)u * SEL_THREADS;
        xs[u] = (i < B) ? __ldg(v + i) : __longlong_as_double(0x7ff8000000000000LL);
    }
}
__device__ __forceinline__ double sel_value(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
    return __longlong_as_double((long long)b);
}

// Q probabilities probs[0..Q) of this launch land in out[col * Q_total + q_first + q]
template <int SEL_THREADS>
__global__ void __launch_bounds__(SEL_THREADS) ppc_select_kernel(const double* __restrict__ series, long long B, long long n_cols, int Q,
                                                                 const double* __restrict__ probs, double* __restrict__ out_all, int Q_total,
                                                                 int q_first) {
    probs += q_first;
    __shared__ unsigned hist[SEL_MAXT][256];
    __shared__ unsigned long long cand[SEL_MAXT][SEL_CAP];
    __shared__ unsigned cand_n[SEL_MAXT];
    __shared__ unsigned long long t_prefix[SEL_MAXT], t_value[SEL_MAXT], b_prefix[SEL_MAXT];
    __shared__ long long t_want[SEL_MAXT], t_rank[SEL_MAXT];
    __shared__ unsigned t_size[SEL_MAXT];
    __shared__ int t_bucket[SEL_MAXT];
    __shared__ unsigned long long s_min, s_max;          // AND and OR of the column's keys
    __shared__ unsigned s_bloom[128];                    // which 12-bit prefix tails belong to a live bucket (4096 bits)
    __shared__ unsigned long long s_cnt;
    __shared__ int n_targets, n_buckets, s_shift, s_go;
    const int tid = threadIdx.x;
    for (long long col = blockIdx.x; col < n_cols; col += gridDim.x) {
        const double* v = series + (size_t)col * B;
        if (tid == 0) { s_min = ~0ULL; s_max = 0ULL; s_cnt = 0ULL; }
        __syncthreads();
        {   // sweep 0: how many non-NaN keys, and which leading bits they all share (AND / OR of the keys: two logic ops per key,
            // where a minimum and a maximum of 64-bit keys cost a dozen)
            unsigned long long lo = ~0ULL, hi = 0ULL, c = 0ULL;
            for (long long i0 = tid; i0 < B; i0 += SEL_UNROLL * SEL_THREADS) {
                double xs[SEL_UNROLL];
                sel_load<SEL_THREADS>(v, i0, B, xs);
#pragma unroll
                for (int u = 0; u < SEL_UNROLL; ++u) {
                    const double x = xs[u];
                    if (x == x) { const unsigned long long k = sel_key(x); lo &= k; hi |= k; ++c; }
                }
            }
            for (int o = 16; o >= 1; o >>= 1) {
                lo &= __shfl_xor_sync(0xffffffffu, lo, o); hi |= __shfl_xor_sync(0xffffffffu, hi, o); c += __shfl_xor_sync(0xffffffffu, c, o);
            }
            if ((tid & 31) == 0) { atomicAnd(&s_min, lo); atomicOr(&s_max, hi); atomicAdd(&s_cnt, c); }
        }
        __syncthreads();
        const long long cnt = (long long)s_cnt;
        if (cnt == 0) {                                            // no valid draw: every quantile is NaN
            for (int q = tid; q < Q; q += SEL_THREADS) out_all[(size_t)col * Q_total + q_first + q] = nan("");
            __syncthreads();
            continue;
        }
        if (tid == 0) {
            // the ranks wanted: floor((cnt - 1) p) and its right neighbour for every probability, without repeats
            int nt = 0;
            for (int q = 0; q < Q; ++q) {
                const double h = (double)(cnt - 1) * probs[q];
                long long i0 = (long long)floor(h);
                if (i0 < 0) i0 = 0;
                if (i0 > cnt - 1) i0 = cnt - 1;
                const long long i1 = (i0 + 1 < cnt) ? i0 + 1 : i0;
                for (int w = 0; w < 2; ++w) {
                    const long long r = w ? i1 : i0;
                    bool seen = false;
                    for (int t = 0; t < nt; ++t) seen = seen || (t_want[t] == r);
                    if (!seen) { t_want[nt] = r; ++nt; }
                }
            }
            n_targets = nt;
            const unsigned long long diff = s_min ^ s_max;
            const int hb = diff ? (64 - __clzll((long long)diff)) : 0;     // low bits in which the keys differ at all
            s_shift = hb;
            for (int t = 0; t < nt; ++t) {
                t_prefix[t] = (hb < 64) ? (s_min >> hb) : 0ULL;
                t_rank[t] = t_want[t];
                t_size[t] = (unsigned)((cnt > 0xffffffffLL) ? 0xffffffffLL : cnt);
            }
        }
        __syncthreads();
        const int first_shift = s_shift;
        while (true) {
            if (tid == 0) {
                // go on while some bucket is still too large for the gather and bits remain; buckets = distinct prefixes
                int go = 0;
                for (int t = 0; t < n_targets; ++t) go = go || (t_size[t] > (unsigned)SEL_CAP);
                s_go = go && (s_shift > 0);
                int nb = 0;
                for (int t = 0; t < n_targets; ++t) {
                    int a = -1;
                    for (int j = 0; j < nb; ++j) if (b_prefix[j] == t_prefix[t]) a = j;
                    if (a < 0) { a = nb; b_prefix[nb] = t_prefix[t]; ++nb; }
                    t_bucket[t] = a;
                }
                n_buckets = nb;
            }
            __syncthreads();
            if (!s_go) break;
            // digit width: as many bits as the 4096 histogram cells allow for the buckets alive -- 12 bits while all ranks still
            // share one bucket (the first step: 100 k keys fall to ~25 per cell, so the usual column needs no second step)
            const int nb = n_buckets;
            const int dig = (nb == 1) ? 12 : (nb <= 4) ? 10 : 8;
            const int shift = s_shift, nshift = (shift > dig) ? shift - dig : 0, width = shift - nshift;
            unsigned* hflat = &hist[0][0];
            for (int i = tid; i < (nb << width); i += SEL_THREADS) hflat[i] = 0u;
            if (tid < 128) s_bloom[tid] = 0u;
            __syncthreads();
            if (tid < nb) atomicOr(&s_bloom[(unsigned)(b_prefix[tid] & 4095ULL) >> 5], 1u << (unsigned)(b_prefix[tid] & 31ULL));
            __syncthreads();
            const bool everyone = (shift == first_shift);      // first step: the prefix is the one ALL keys share, nothing to test
            for (long long i0 = tid; i0 < B; i0 += SEL_UNROLL * SEL_THREADS) {
                double xs[SEL_UNROLL];
                sel_load<SEL_THREADS>(v, i0, B, xs);
#pragma unroll
                for (int u = 0; u < SEL_UNROLL; ++u) {
                    const double x = xs[u];
                    if (x == x) {
                        const unsigned long long k = sel_key(x);
                        const unsigned d = (unsigned)(k >> nshift) & ((1u << width) - 1u);
                        if (everyone) {
                            atomicAdd(&hflat[d], 1u);
                        } else {
                            const unsigned long long hi = (shift < 64) ? (k >> shift) : 0ULL;
                            const unsigned tail = (unsigned)hi & 4095u;
                            if ((s_bloom[tail >> 5] >> (tail & 31u)) & 1u)
                                for (int a = 0; a < nb; ++a) if (hi == b_prefix[a]) atomicAdd(&hflat[((unsigned)a << width) + d], 1u);
                        }
                    }
                }
            }
            __syncthreads();
            // one warp per wanted rank: the cell that holds it (lanes sum 1/32 of the cells each, a shuffle scan finds the lane,
            // the lane walks its cells)
            for (int t = tid >> 5; t < n_targets; t += SEL_THREADS / 32) {
                const int lane = tid & 31;
                const unsigned* h = hflat + ((unsigned)t_bucket[t] << width);
                const int cells = 1 << width, per = (cells + 31) / 32, c0 = lane * per, c1 = (c0 + per < cells) ? c0 + per : cells;
                long long mine = 0;
                for (int c = c0; c < c1; ++c) mine += h[c];
                long long incl = mine;
                for (int o = 1; o < 32; o <<= 1) { const long long up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
                const long long k = t_rank[t], before = incl - mine;
                const bool owner = (k >= before && k < incl) || (lane == 31 && k >= incl);      // (a rank past the end cannot happen; the last lane would take it)
                if (owner) {
                    long long r = k - before;
                    int d = c0;
                    for (; d < c1 - 1; ++d) { const unsigned hc = h[d]; if (r < (long long)hc) break; r -= hc; }
                    t_prefix[t] = (t_prefix[t] << width) | (unsigned long long)d;
                    t_rank[t] = r;
                    t_size[t] = h[d];
                }
            }
            __syncthreads();
            if (tid == 0) s_shift = nshift;
            __syncthreads();
        }
        const int shift = s_shift;
        bool small = true;
        for (int t = 0; t < n_targets; ++t) small = small && (t_size[t] <= (unsigned)SEL_CAP);
        if (!small) {
            // bits used up with a large bucket: all its keys are equal, the prefix IS the key
            if (tid < n_targets) t_value[tid] = t_prefix[tid];
        } else {
            const int nb = n_buckets;
            if (tid < nb) cand_n[tid] = 0u;
            if (tid < 128) s_bloom[tid] = 0u;
            __syncthreads();
            if (tid < nb) atomicOr(&s_bloom[(unsigned)(b_prefix[tid] & 4095ULL) >> 5], 1u << (unsigned)(b_prefix[tid] & 31ULL));
            __syncthreads();
            for (long long i0 = tid; i0 < B; i0 += SEL_UNROLL * SEL_THREADS) {
                double xs[SEL_UNROLL];
                sel_load<SEL_THREADS>(v, i0, B, xs);
#pragma unroll
                for (int u = 0; u < SEL_UNROLL; ++u) {
                    const double x = xs[u];
                    if (x == x) {
                        const unsigned long long k = sel_key(x);
                        const unsigned long long hi = (shift < 64) ? (k >> shift) : 0ULL;
                        const unsigned tail = (unsigned)hi & 4095u;
                        if ((s_bloom[tail >> 5] >> (tail & 31u)) & 1u)
                            for (int a = 0; a < nb; ++a)
                                if (hi == b_prefix[a]) { const unsigned pos = atomicAdd(&cand_n[a], 1u); if (pos < (unsigned)SEL_CAP) cand[a][pos] = k; }
                    }
                }
            }
            __syncthreads();
            for (int t = 0; t < n_targets; ++t) {
                const int a = t_bucket[t];
                const unsigned m = (cand_n[a] < (unsigned)SEL_CAP) ? cand_n[a] : (unsigned)SEL_CAP;
                const long long k = t_rank[t];
                for (unsigned c = tid; c < m; c += SEL_THREADS) {
                    const unsigned long long key = cand[a][c];
                    long long less = 0, eq = 0;
                    for (unsigned j = 0; j < m; ++j) { const unsigned long long o = cand[a][j]; less += (o < key); eq += (o == key); }
                    if (less <= k && k < less + eq) t_value[t] = key;       // every candidate of that value writes the same bits
                }
            }
        }
        __syncthreads();
        for (int q = tid; q < Q; q += SEL_THREADS) {
            const double h = (double)(cnt - 1) * probs[q];
            long long i0 = (long long)floor(h);
            if (i0 < 0) i0 = 0;
            if (i0 > cnt - 1) i0 = cnt - 1;
            const long long i1 = (i0 + 1 < cnt) ? i0 + 1 : i0;
            double a = 0.0, c = 0.0;
            for (int t = 0; t < n_targets; ++t) {
                if (t_want[t] == i0) a = sel_value(t_value[t]);
                if (t_want[t] == i1) c = sel_value(t_value[t]);
            }
            out_all[(size_t)col * Q_total + q_first + q] = a + (h - (double)i0) * (c - a);
        }
        __syncthreads();
    }
}

__global__ void ppc_count_valid_kernel(const unsigned* st, long long B, unsigned long long* out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const unsigned ok = (i < B && st[i] == 0u) ? 1u : 0u;
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

}  // namespace

extern "C" sepaihrd_rc sepaihrd_posterior_predictive(sepaihrd_ctx* ctx, const double* params, int64_t B, int64_t ld,
                                                     const double* initial_state, int32_t n_probs, const double* probs,
                                                     double* out_quantiles, int64_t* out_valid_draws) {
    using namespace sepaihrd_internal;
    if (!ctx || !params || !initial_state || !probs || !out_quantiles) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    const Dims d = dims(ctx);
    if (B <= 0 || ld < d.P) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "Parameter vector size mismatch.");
    if (n_probs < 1 || n_probs > 64) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "between 1 and 64 quantile probabilities");
    for (int q = 0; q < n_probs; ++q)
        if (!(probs[q] >= 0.0 && probs[q] <= 1.0)) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "quantile probabilities must lie in [0, 1]");
    const int n = d.n, T = d.n_nonneg;
    if (T <= 0) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "No non-negative time points for PPC.");   // ResultAggregator.cpp:197-200
    const long long n_cols = (long long)T * n;            // columns per series
    if ((double)n_cols * (double)B > 2.0e9) return fail_with(SEPAIHRD_ERR_UNSUPPORTED, "more than 2e9 values per series: split the draws");

    const auto ctx_lock = lock(ctx);
    cudaSetDevice(d.device);
    cudaStream_t s = stream(ctx);
    // work buffers live in the ctx (grow-only scratch slots): repeated aggregations of the same size allocate nothing
    auto cleanup = [] {};
    const size_t col_elems = (size_t)n_cols * (size_t)B;
    int slot = 0;
    bool oom = false;
    auto buf = [&](size_t bytes) { void* p = scratch(ctx, slot++, bytes); oom = oom || (p == nullptr); return p; };
    double* d_params = (double*)buf(sizeof(double) * (size_t)B * ld);
    double* d_init = (double*)buf(sizeof(double) * SEPAIHRD_NUM_COMPARTMENTS * n);
    double* d_series = (double*)buf(sizeof(double) * 6 * col_elems);
    double* d_probs = (double*)buf(sizeof(double) * n_probs);
    double* d_q = (double*)buf(sizeof(double) * 6 * (size_t)n_cols * n_probs);
    unsigned* d_status = (unsigned*)buf(sizeof(unsigned) * (size_t)B);
    unsigned long long* d_cnt = (unsigned long long*)buf(sizeof(unsigned long long));
    if (oom) return fail_with(SEPAIHRD_ERR_OUT_OF_MEMORY, "posterior-predictive work buffers do not fit: split the draws");
    // padded age classes (sepaihrd_create): the caller's state has d.n_user classes per compartment, the pass runs with n
    std::vector<double> wide_init;
    if (n != d.n_user) {
        wide_init.assign((size_t)SEPAIHRD_NUM_COMPARTMENTS * n, 0.0);
        for (int cpt = 0; cpt < SEPAIHRD_NUM_COMPARTMENTS; ++cpt)
            for (int a = 0; a < d.n_user; ++a) wide_init[(size_t)cpt * n + a] = initial_state[(size_t)cpt * d.n_user + a];
        initial_state = wide_init.data();
    }
    PPC_TRY(cudaMemcpyAsync(d_init, initial_state, sizeof(double) * SEPAIHRD_NUM_COMPARTMENTS * n, cudaMemcpyHostToDevice, s));
    if (!wide_init.empty()) PPC_TRY(cudaStreamSynchronize(s));      // the staging vector is pageable and local
    PPC_TRY(cudaMemcpyAsync(d_probs, probs, sizeof(double) * n_probs, cudaMemcpyHostToDevice, s));
    PPC_TRY(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), s));

    // 1. the six series, formed by the trajectory kernel while it integrates.  The draws go to the device in growing chunks on a
    //    copy stream of their own: the kernel of chunk c runs while chunk c + 1 is copied (a draw costs the kernel ~90 ns and the
    //    link ~10-40 ns, pinned or pageable), so only the first, small copy is exposed.
    {
        cudaStream_t copy_stream = sepaihrd_internal::copy_stream(ctx);      // the ctx's own copy stream and chunk events (the ctx lock is held)
        cudaEvent_t* ev_copy = sepaihrd_internal::copy_events(ctx);          // [8]
        cudaEvent_t ev_free = sepaihrd_internal::chunk_events(ctx)[0];
        PPC_TRY(cudaEventRecord(ev_free, s));                         // the copy stream may not overwrite d_params while earlier work on s reads it
        PPC_TRY(cudaStreamWaitEvent(copy_stream, ev_free, 0));
        long long ends[8];
        int nc = 0;
        // chunk ends at 1, 4 and 16 waves of the trajectory kernel (a wave = every resident warp holding one tile: 148 SMs x 8 warps x
        // 32 / n draws), so no chunk pays for a mostly empty last wave
        const long long wave = 148LL * 8 * (32 / n);
        for (long long w = 1; w <= 16; w *= 4) if (w * wave * 2 <= B) ends[nc++] = w * wave;
        ends[nc++] = B;
        long long b0 = 0;
        for (int c = 0; c < nc; ++c) {
            // copy and launch alternate: a copy from PAGEABLE memory blocks the host until it is staged, and it should block it
            // while the previous chunk's kernel is already running
            PPC_TRY(cudaMemcpyAsync(d_params + b0 * ld, params + b0 * ld, sizeof(double) * (size_t)(ends[c] - b0) * ld, cudaMemcpyHostToDevice, copy_stream));
            PPC_TRY(cudaEventRecord(ev_copy[c], copy_stream));
            PPC_TRY(cudaStreamWaitEvent(s, ev_copy[c], 0));
            const sepaihrd_rc rc = simulate_ppc_series(ctx, d_params, b0, ends[c] - b0, B, ld, d_init, d_series, d_status);
            if (rc != SEPAIHRD_OK) { cudaStreamSynchronize(copy_stream); cleanup(); return rc; }
            b0 = ends[c];
        }
    }
    ppc_count_valid_kernel<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(d_status, B, d_cnt);
    PPC_TRY(cudaGetLastError());
    // 2. the order statistics the quantiles need, column by column, 8 probabilities (<= 16 ranks) per group
    {
        const long long all_cols = 6 * n_cols;
        // Columns in flight x column size is what the later sweeps of a column find in L2 (126 MB): one 1024-thread block per SM
        // keeps 148 columns of 100 k draws (118 MB) resident, so only the first sweep of a column reads HBM.
        static const int env_grid = std::getenv("SEPAIHRD_PPC_GRID") ? std::atoi(std::getenv("SEPAIHRD_PPC_GRID")) : 0;
        static const int env_threads = std::getenv("SEPAIHRD_PPC_THREADS") ? std::atoi(std::getenv("SEPAIHRD_PPC_THREADS")) : 0;
        const int threads = env_threads ? env_threads : PPC_DEFAULT_THREADS;
        const unsigned grid = (unsigned)std::min<long long>(all_cols, env_grid ? env_grid : PPC_DEFAULT_GRID);
        for (int q0 = 0; q0 < n_probs; q0 += SEL_MAXT / 2) {
            const int qg = std::min(SEL_MAXT / 2, n_probs - q0);
            if (threads == 1024) ppc_select_kernel<1024><<<grid, 1024, 0, s>>>(d_series, B, all_cols, qg, d_probs, d_q, n_probs, q0);
            else if (threads == 256) ppc_select_kernel<256><<<grid, 256, 0, s>>>(d_series, B, all_cols, qg, d_probs, d_q, n_probs, q0);
            else ppc_select_kernel<512><<<grid, 512, 0, s>>>(d_series, B, all_cols, qg, d_probs, d_q, n_probs, q0);
            PPC_TRY(cudaGetLastError());
            count_launches(ctx, 1);
        }
    }
    unsigned long long cnt = 0;
    std::vector<double> wide_q;
    double* q_dst = out_quantiles;
    if (n != d.n_user) { wide_q.resize(6 * (size_t)n_cols * n_probs); q_dst = wide_q.data(); }
    PPC_TRY(cudaMemcpyAsync(q_dst, d_q, sizeof(double) * 6 * (size_t)n_cols * n_probs, cudaMemcpyDeviceToHost, s));
    PPC_TRY(cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, s));
    PPC_TRY(cudaStreamSynchronize(s));
    if (n != d.n_user)                                                  // [6][T][n][q] -> [6][T][n_user][q]
        for (size_t st = 0; st < 6 * (size_t)T; ++st)
            for (int a = 0; a < d.n_user; ++a)
                for (int q = 0; q < n_probs; ++q) out_quantiles[(st * d.n_user + a) * n_probs + q] = wide_q[(st * n + a) * n_probs + q];
    if (out_valid_draws) *out_valid_draws = (int64_t)cnt;
    cleanup();
    return SEPAIHRD_OK;
}
