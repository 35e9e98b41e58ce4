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
// Data flow (all in HBM, HBM-bound integer/byte-style work -- no tensor cores):
//   trajectory kernel --[K][3n][B], draws fastest--> series kernel --[6][T][n][B]--> per column: the <= 2Q order statistics
//   the quantiles need, SELECTED by ppc_select_kernel (up to 8 probabilities) or read off a cub segmented radix sort (more)
//   --[6][T][n][Q]--> host
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <vector>

#include <cub/device/device_segmented_radix_sort.cuh>

#include "sepaihrd_internal.h"

namespace {

#define PPC_TRY(expr)                                                                                        \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess) { cleanup(); return sepaihrd_internal::fail_with(SEPAIHRD_ERR_CUDA, cudaGetErrorString(e__)); } \
    } while (0)

// One thread per (draw b, stream, age): walks the output days t >= 0 in order.  Reads and writes are coalesced over b.
//   traj   [K][3n][B]: w = 0*n+age -> D, 1*n+age -> CumH, 2*n+age -> CumICU   (TRAJ_OBSERVED order)
//   series [6][T][n][B]: daily hosp, daily icu, daily deaths, cumulative hosp, cumulative icu, cumulative deaths
__global__ void ppc_series_kernel(const double* __restrict__ traj, const double* __restrict__ init_state, int n, int first_pos, int T,
                                  long long B, double* __restrict__ series) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int stream = blockIdx.y / n, age = blockIdx.y % n;      // stream: 0 hosp, 1 icu, 2 deaths
    const int w = ((stream == 0) ? 1 : (stream == 1) ? 2 : 0) * n + age;
    const int comp = (stream == 0) ? 9 : (stream == 1) ? 10 : 8;
    const size_t W = (size_t)3 * n;
    double prev = (first_pos > 0) ? traj[((size_t)(first_pos - 1) * W + w) * B + b] : init_state[comp * n + age];
    double run = 0.0;
    double* daily = series + (((size_t)stream * T) * n + age) * B + b;
    double* cum = series + (((size_t)(3 + stream) * T) * n + age) * B + b;
    const size_t step = (size_t)n * B;
    for (int t = 0; t < T; ++t) {
        const double v = traj[((size_t)(first_pos + t) * W + w) * B + b];
        double d = v - prev;
        d = (0.0 < d) ? d : 0.0;          // std::max(0.0, diff)
        if (v != v) d = v;                // a failed draw is NaN-filled by the trajectory kernel: keep it out of the quantiles
        prev = v;
        run += d;
        daily[(size_t)t * step] = d;
        cum[(size_t)t * step] = run;
    }
}

// One warp per sorted column: number of non-NaN entries by binary search (NaNs sort last), then linear interpolation.
__global__ void ppc_quantile_kernel(const double* __restrict__ sorted, long long B, long long n_cols, int Q, const double* __restrict__ probs,
                                    double* __restrict__ out) {
    const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (col >= n_cols) return;
    const double* v = sorted + (size_t)col * B;
    long long lo = 0, hi = B;             // first index whose value is NaN
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (v[mid] != v[mid]) hi = mid; else lo = mid + 1;
    }
    const long long cnt = lo;
    for (int q = 0; q < Q; ++q) {
        double r = nan("");
        if (cnt > 0) {
            const double h = (double)(cnt - 1) * probs[q];
            long long i0 = (long long)floor(h);
            if (i0 < 0) i0 = 0;
            if (i0 > cnt - 1) i0 = cnt - 1;
            const long long i1 = (i0 + 1 < cnt) ? i0 + 1 : i0;
            const double a = v[i0], c = v[i1];
            r = a + (h - (double)i0) * (c - a);
        }
        out[(size_t)col * Q + q] = r;
    }
}

// ---- exact order statistics without sorting -------------------------------------------------------------------------------
// A column needs the values at <= 2Q ranks (the two neighbours of every quantile position), not its full order.  One block per
// column: a first sweep finds the count, minimum and maximum of the non-NaN keys (everything above their common bit prefix is
// already decided); then most-significant-digit radix steps of 8 bits narrow, for ALL wanted ranks at once, the bucket each rank
// lies in (one 256-bin histogram per distinct bucket prefix, <= 16 of them), until every bucket holds at most SEL_CAP keys or
// the bits are used up; the survivors are gathered into shared memory and the wanted rank is picked by counting.  A column is
// read 3-5 times (its later sweeps mostly from L2) instead of being read and written 8 times by the radix sort.
// Keys: the usual order-preserving map of a double's bits to an unsigned integer; NaNs (failed draws) are left out.
constexpr int SEL_THREADS = 512, SEL_MAXT = 16, SEL_CAP = 192;

__device__ __forceinline__ unsigned long long sel_key(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
// SEL_UNROLL coalesced loads issued together (the sweeps are latency-bound otherwise); past the end: NaN, which every sweep skips
constexpr int SEL_UNROLL = 4;
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

__global__ void __launch_bounds__(SEL_THREADS) ppc_select_kernel(const double* __restrict__ series, long long B, long long n_cols, int Q,
                                                                 const double* __restrict__ probs, double* __restrict__ out) {
    __shared__ unsigned hist[SEL_MAXT][256];
    __shared__ unsigned long long cand[SEL_MAXT][SEL_CAP];
    __shared__ unsigned cand_n[SEL_MAXT];
    __shared__ unsigned long long t_prefix[SEL_MAXT], t_value[SEL_MAXT], b_prefix[SEL_MAXT];
    __shared__ long long t_want[SEL_MAXT], t_rank[SEL_MAXT];
    __shared__ unsigned t_size[SEL_MAXT];
    __shared__ int t_bucket[SEL_MAXT];
    __shared__ unsigned long long s_min, s_max;
    __shared__ unsigned long long s_cnt;
    __shared__ int n_targets, n_buckets, s_shift, s_go;
    const int tid = threadIdx.x;
    for (long long col = blockIdx.x; col < n_cols; col += gridDim.x) {
        const double* v = series + (size_t)col * B;
        if (tid == 0) { s_min = ~0ULL; s_max = 0ULL; s_cnt = 0ULL; }
        __syncthreads();
        {   // sweep 0: how many non-NaN keys, and between which bounds
            unsigned long long lo = ~0ULL, hi = 0ULL, c = 0ULL;
            for (long long i0 = tid; i0 < B; i0 += SEL_UNROLL * SEL_THREADS) {
                double xs[SEL_UNROLL];
                sel_load(v, i0, B, xs);
#pragma unroll
                for (int u = 0; u < SEL_UNROLL; ++u) {
                    const double x = xs[u];
                    if (x == x) { const unsigned long long k = sel_key(x); lo = (k < lo) ? k : lo; hi = (k > hi) ? k : hi; ++c; }
                }
            }
            for (int o = 16; o >= 1; o >>= 1) {
                const unsigned long long lo2 = __shfl_xor_sync(0xffffffffu, lo, o), hi2 = __shfl_xor_sync(0xffffffffu, hi, o);
                lo = (lo2 < lo) ? lo2 : lo; hi = (hi2 > hi) ? hi2 : hi; c += __shfl_xor_sync(0xffffffffu, c, o);
            }
            if ((tid & 31) == 0) { atomicMin(&s_min, lo); atomicMax(&s_max, hi); atomicAdd(&s_cnt, c); }
        }
        __syncthreads();
        const long long cnt = (long long)s_cnt;
        if (cnt == 0) {                                            // no valid draw: every quantile is NaN
            for (int q = tid; q < Q; q += SEL_THREADS) out[(size_t)col * Q + q] = nan("");
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
            const int shift = s_shift, nshift = (shift > 8) ? shift - 8 : 0, width = shift - nshift, nb = n_buckets;
            for (int i = tid; i < nb * 256; i += SEL_THREADS) hist[i >> 8][i & 255] = 0u;
            __syncthreads();
            for (long long i0 = tid; i0 < B; i0 += SEL_UNROLL * SEL_THREADS) {
                double xs[SEL_UNROLL];
                sel_load(v, i0, B, xs);
#pragma unroll
                for (int u = 0; u < SEL_UNROLL; ++u) {
                    const double x = xs[u];
                    if (x == x) {
                        const unsigned long long k = sel_key(x);
                        const unsigned long long hi = (shift < 64) ? (k >> shift) : 0ULL;
                        const unsigned d = (unsigned)((k >> nshift) & ((1ULL << width) - 1ULL));
                        for (int a = 0; a < nb; ++a) if (hi == b_prefix[a]) atomicAdd(&hist[a][d], 1u);
                    }
                }
            }
            __syncthreads();
            if (tid < n_targets) {
                const int a = t_bucket[tid];
                long long k = t_rank[tid];
                unsigned d = 0;
                for (; d < (1u << width) - 1u; ++d) {
                    const unsigned hcount = hist[a][d];
                    if (k < (long long)hcount) break;
                    k -= hcount;
                }
                t_prefix[tid] = (t_prefix[tid] << width) | d;
                t_rank[tid] = k;
                t_size[tid] = hist[a][d];
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
            __syncthreads();
            for (long long i0 = tid; i0 < B; i0 += SEL_UNROLL * SEL_THREADS) {
                double xs[SEL_UNROLL];
                sel_load(v, i0, B, xs);
#pragma unroll
                for (int u = 0; u < SEL_UNROLL; ++u) {
                    const double x = xs[u];
                    if (x == x) {
                        const unsigned long long k = sel_key(x);
                        const unsigned long long hi = (shift < 64) ? (k >> shift) : 0ULL;
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
            out[(size_t)col * Q + q] = a + (h - (double)i0) * (c - a);
        }
        __syncthreads();
    }
}

__global__ void ppc_offsets_kernel(long long* off, long long n_cols, long long B) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i <= n_cols) off[i] = i * B;
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
    const int n = d.n, K = d.K, T = d.n_nonneg, first_pos = K - T;
    if (T <= 0) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "No non-negative time points for PPC.");   // ResultAggregator.cpp:197-200
    const long long n_cols = (long long)T * n;            // columns per series
    if ((double)n_cols * (double)B > 2.0e9) return fail_with(SEPAIHRD_ERR_UNSUPPORTED, "more than 2e9 values per series: split the draws");

    const auto ctx_lock = lock(ctx);
    cudaSetDevice(d.device);
    cudaStream_t s = stream(ctx);
    // work buffers live in the ctx (grow-only scratch slots): repeated aggregations of the same size allocate nothing
    auto cleanup = [] {};
    const size_t traj_elems = (size_t)K * 3 * n * (size_t)B, col_elems = (size_t)n_cols * (size_t)B;
    int slot = 0;
    bool oom = false;
    auto buf = [&](size_t bytes) { void* p = scratch(ctx, slot++, bytes); oom = oom || (p == nullptr); return p; };
    double* d_params = (double*)buf(sizeof(double) * (size_t)B * ld);
    double* d_init = (double*)buf(sizeof(double) * SEPAIHRD_NUM_COMPARTMENTS * n);
    double* d_traj = (double*)buf(sizeof(double) * traj_elems);
    double* d_series = (double*)buf(sizeof(double) * 6 * col_elems);
    static const bool force_sort = std::getenv("SEPAIHRD_PPC_SORT") != nullptr;
    const bool use_select = 2 * n_probs <= SEL_MAXT && !force_sort;
    double* d_sorted = use_select ? nullptr : (double*)buf(sizeof(double) * col_elems);     // only the sort path needs a second copy of a series
    double* d_probs = (double*)buf(sizeof(double) * n_probs);
    double* d_q = (double*)buf(sizeof(double) * 6 * (size_t)n_cols * n_probs);
    unsigned* d_status = (unsigned*)buf(sizeof(unsigned) * (size_t)B);
    long long* d_off = (long long*)buf(sizeof(long long) * (size_t)(n_cols + 1));
    unsigned long long* d_cnt = (unsigned long long*)buf(sizeof(unsigned long long));
    if (oom) return fail_with(SEPAIHRD_ERR_OUT_OF_MEMORY, "posterior-predictive work buffers do not fit: split the draws");
    void* d_tmp = nullptr;
    PPC_TRY(cudaMemcpyAsync(d_params, params, sizeof(double) * (size_t)B * ld, cudaMemcpyHostToDevice, s));
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

    // 1. trajectories of D, CumH, CumICU, draws fastest
    sepaihrd_rc rc = simulate_observed_draw_minor(ctx, d_params, B, ld, d_init, d_traj, d_status);
    if (rc != SEPAIHRD_OK) { cleanup(); return rc; }
    // 2. daily incidence + cumulative-from-flows series
    {
        const int threads = 256;
        dim3 grid((unsigned)((B + threads - 1) / threads), (unsigned)(3 * n));
        ppc_series_kernel<<<grid, threads, 0, s>>>(d_traj, d_init, n, first_pos, T, B, d_series);
        PPC_TRY(cudaGetLastError());
        ppc_count_valid_kernel<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(d_status, B, d_cnt);
        ppc_offsets_kernel<<<(unsigned)((n_cols + 256) / 256), 256, 0, s>>>(d_off, n_cols, B);
        PPC_TRY(cudaGetLastError());
    }
    // 3 + 4. the order statistics the quantiles need, column by column, without sorting (few probabilities: the usual case)
    if (use_select) {
        const long long all_cols = 6 * n_cols;
        const unsigned grid = (unsigned)std::min<long long>(all_cols, 4 * 148);
        ppc_select_kernel<<<grid, SEL_THREADS, 0, s>>>(d_series, B, all_cols, n_probs, d_probs, d_q);
        PPC_TRY(cudaGetLastError());
        count_launches(ctx, 1);
    } else {
    // 3. sort every column (one segment per (day, age)), one series at a time; 4. gather the quantiles
    size_t tmp_bytes = 0;
    PPC_TRY(cub::DeviceSegmentedRadixSort::SortKeys(nullptr, tmp_bytes, d_series, d_sorted, (long long)col_elems, (long long)n_cols, d_off, d_off + 1,
                                                     0, 64, s));
    d_tmp = scratch(ctx, slot++, tmp_bytes ? tmp_bytes : 16);
    if (!d_tmp) return fail_with(SEPAIHRD_ERR_OUT_OF_MEMORY, "posterior-predictive sort buffer does not fit: split the draws");
    for (int ser = 0; ser < 6; ++ser) {
        PPC_TRY(cub::DeviceSegmentedRadixSort::SortKeys(d_tmp, tmp_bytes, d_series + (size_t)ser * col_elems, d_sorted, (long long)col_elems,
                                                         (long long)n_cols, d_off, d_off + 1, 0, 64, s));
        ppc_quantile_kernel<<<(unsigned)((n_cols + 127) / 128), 128, 0, s>>>(d_sorted, B, n_cols, n_probs, d_probs, d_q + (size_t)ser * n_cols * n_probs);
        PPC_TRY(cudaGetLastError());
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
