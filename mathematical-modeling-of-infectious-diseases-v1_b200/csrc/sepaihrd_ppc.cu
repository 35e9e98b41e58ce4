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
// quantile with linear interpolation between order statistics (numpy's default), from a full sort of the column.
//
// Data flow (all in HBM, HBM-bound integer/byte-style work -- no tensor cores):
//   trajectory kernel --[K][3n][B], draws fastest--> series kernel --[6][T][n][B]--> segmented radix sort (cub) -->
//   quantile gather --[6][T][n][Q]--> host
#include <cmath>
#include <cstdint>
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
    double* d_sorted = (double*)buf(sizeof(double) * col_elems);
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
