// sepaihrd_order.cu -- the ordering pass in front of a large likelihood launch.
//
// The fused kernel steps the 8 parameter sets of a warp in lockstep: per output day a warp pays the LARGEST number of step
// attempts among its sets ("warp-synchronous stepping").  On a widely spread batch -- uniform in the
// bounds, the particle-swarm initialisation recipe of the reference (ParticleSwarmOptimizer.cpp:291) -- 22 % of the
// lane-attempts are such idle repeats.  Which sets need how many attempts on which day is a smooth function of the parameters,
// so it can be PREDICTED before the launch and the sets handed to the warps in an order that puts alike ones together:
//
//   fit (once per distribution, ~15 ms):  a pilot of 2048 sets runs through the PROFILE instantiation of the kernel, which
//       records the attempts made before every grid point; the first two principal components of the per-day attempt profile
//       (power iteration) are regressed on the standardised parameters (ridge-regularised normal equations).
//   order (every large launch, ~0.3 ms per 1M sets): one warp per set evaluates the two linear predictors, the sets are
//       dropped into 4096 x 16 buckets (first component fine, second coarse) by a counting sort, and the kernel takes its tiles
//       through the resulting index list (KParams::perm); every result still lands in its set's own slot.
//
// Measured on a B200, 1,048,576 uniform-in-bounds sets: 113.4 -> ~100 ms per launch (0.463 -> ~0.52 of the FP64 peak), nothing
// lost on the jittered batch; the results are bit-identical to the unordered launch (a set's arithmetic does not depend on its
// neighbours).  This is HBM / integer work in front of a compute-bound kernel; no part of the reference corresponds to it.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "sepaihrd_internal.h"

namespace {

using sepaihrd_internal::fail_with;

#define ORD_TRY(expr)                                                                             \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) return fail_with(SEPAIHRD_ERR_CUDA, cudaGetErrorString(e__));     \
    } while (0)

constexpr long long ORDER_MIN_BATCH = 32768;       // below this a launch is a few waves: nothing to gain
constexpr int PILOT = 2048;
constexpr int NB1 = 4096, NB2 = 16, NBUCKETS = NB1 * NB2;

struct OrderModel {
    int P = 0, K = 0;
    bool fitted = false;
    std::vector<double> mu, inv_sd, sd;             // standardisation of the parameters (pilot statistics)
    double* d_model = nullptr;                      // mu[P] | inv_sd[P] | w1[P + 1] | w2[P + 1]  (intercept last)
    // Work buffers of one ordering pass.  TWO sets, used alternately: the host-buffer evaluation runs consecutive chunks on two
    // streams so that they overlap, and a launch reads its index list until it ends -- a set is reused only after the launch
    // that last read it has finished (its `done` event, awaited on the new launch's stream).
    struct Work {
        float2* d_keys = nullptr;                   // [cap] predicted components
        int* d_perm = nullptr;                      // [cap]
        unsigned* d_bucket = nullptr;               // [cap]
        unsigned* d_hist = nullptr;                 // [NBUCKETS + 1]
        double* d_stats = nullptr;                  // [4]: sum1, sumsq1, sum2, sumsq2
        long long cap = 0;
        cudaEvent_t done = nullptr;
        bool pending = false;
    } work[2];
    unsigned seq = 0;
    int last = -1;                                  // set handed out by the latest order_batch, until order_mark_done records its event
    long long fits = 0;
};

OrderModel* model_of(sepaihrd_ctx* ctx, bool create) {
    void** slot = sepaihrd_internal::order_slot(ctx);
    if (!*slot && create) *slot = new OrderModel();
    return static_cast<OrderModel*>(*slot);
}

// rows[i] = params[(i * B) / M]: an evenly spaced pilot, whatever the order of the batch
__global__ void order_gather_kernel(const double* __restrict__ params, long long B, long long ld, int M, int P, double* __restrict__ rows) {
    const long long i = blockIdx.x;
    const double* src = params + ((i * B) / M) * ld;
    for (int j = threadIdx.x; j < P; j += blockDim.x) rows[i * P + j] = src[j];
}

// one warp per set: the two linear predictors; block sums of the keys and their squares for the bucket ranges
__global__ void __launch_bounds__(256) order_keys_kernel(const double* __restrict__ params, long long B, long long ld, int P,
                                                          const double* __restrict__ model, float2* __restrict__ keys, double* __restrict__ stats) {
    __shared__ double s_part[8][4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long b = blockIdx.x * 8LL + w;
    const double* mu = model; const double* inv_sd = model + P; const double* w1 = model + 2 * P; const double* w2 = w1 + (P + 1);
    double a1 = 0.0, a2 = 0.0;
    if (b < B) {
        const double* row = params + b * ld;
        for (int j = lane; j < P; j += 32) {
            const double z = (row[j] - mu[j]) * inv_sd[j];
            a1 = fma(w1[j], z, a1);
            a2 = fma(w2[j], z, a2);
        }
    }
    for (int o = 16; o >= 1; o >>= 1) { a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o); }
    if (lane == 0) {
        const bool ok = b < B && a1 == a1 && a2 == a2;
        if (b < B) keys[b] = make_float2(ok ? (float)(a1 + w1[P]) : 0.0f, ok ? (float)(a2 + w2[P]) : 0.0f);
        const double k1 = ok ? a1 + w1[P] : 0.0, k2 = ok ? a2 + w2[P] : 0.0;
        s_part[w][0] = k1; s_part[w][1] = k1 * k1; s_part[w][2] = k2; s_part[w][3] = k2 * k2;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += s_part[i][threadIdx.x];
        atomicAdd(&stats[threadIdx.x], t);
    }
}

// bucket = (first component in NB1 cells over mean +- 3.5 sigma) x (second component in NB2 cells)
__global__ void order_bucket_kernel(const float2* __restrict__ keys, long long B, const double* __restrict__ stats, unsigned* __restrict__ bucket,
                                    unsigned* __restrict__ hist) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double n = (double)B;
    const double m1 = stats[0] / n, m2 = stats[2] / n;
    const double s1 = sqrt(fmax(stats[1] / n - m1 * m1, 1e-300)), s2 = sqrt(fmax(stats[3] / n - m2 * m2, 1e-300));
    const float2 k = keys[b];
    double u1 = ((double)k.x - (m1 - 3.5 * s1)) / (7.0 * s1), u2 = ((double)k.y - (m2 - 3.5 * s2)) / (7.0 * s2);
    u1 = fmin(fmax(u1, 0.0), 0.999999); u2 = fmin(fmax(u2, 0.0), 0.999999);
    const unsigned c = (unsigned)(u1 * NB1) * NB2 + (unsigned)(u2 * NB2);
    bucket[b] = c;
    atomicAdd(&hist[c], 1u);
}

// exclusive scan of the NBUCKETS counts, one block
__global__ void __launch_bounds__(1024) order_scan_kernel(unsigned* __restrict__ hist) {
    __shared__ unsigned s_tot[1024];
    constexpr int PER = NBUCKETS / 1024;
    unsigned local[PER];
    unsigned sum = 0;
    for (int i = 0; i < PER; ++i) { local[i] = hist[threadIdx.x * PER + i]; sum += local[i]; }
    s_tot[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned v = (threadIdx.x >= (unsigned)o) ? s_tot[threadIdx.x - o] : 0u;
        __syncthreads();
        s_tot[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned run = s_tot[threadIdx.x] - sum;
    for (int i = 0; i < PER; ++i) { hist[threadIdx.x * PER + i] = run; run += local[i]; }
}

__global__ void order_scatter_kernel(const unsigned* __restrict__ bucket, long long B, unsigned* __restrict__ cursor, int* __restrict__ perm) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= B) return;
    perm[atomicAdd(&cursor[bucket[b]], 1u)] = (int)b;
}

sepaihrd_rc ensure_buffers(OrderModel::Work& w, long long B) {
    if (!w.done) ORD_TRY(cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming));
    if (!w.d_hist) ORD_TRY(cudaMalloc((void**)&w.d_hist, sizeof(unsigned) * (NBUCKETS + 1)));
    if (!w.d_stats) ORD_TRY(cudaMalloc((void**)&w.d_stats, sizeof(double) * 4));
    if (B <= w.cap) return SEPAIHRD_OK;
    ORD_TRY(cudaDeviceSynchronize());
    if (w.d_keys) cudaFree(w.d_keys);
    if (w.d_perm) cudaFree(w.d_perm);
    if (w.d_bucket) cudaFree(w.d_bucket);
    w.d_keys = nullptr; w.d_perm = nullptr; w.d_bucket = nullptr; w.cap = 0; w.pending = false;
    ORD_TRY(cudaMalloc((void**)&w.d_keys, sizeof(float2) * (size_t)B));
    ORD_TRY(cudaMalloc((void**)&w.d_perm, sizeof(int) * (size_t)B));
    ORD_TRY(cudaMalloc((void**)&w.d_bucket, sizeof(unsigned) * (size_t)B));
    w.cap = B;
    return SEPAIHRD_OK;
}

// lower Cholesky factor of the symmetric positive definite n x n matrix a (row-major, overwritten by L); false if it breaks down
bool cholesky(std::vector<double>& a, int n) {
    for (int j = 0; j < n; ++j) {
        double d = a[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) d -= a[(size_t)j * n + k] * a[(size_t)j * n + k];
        if (!(d > 0.0)) return false;
        d = std::sqrt(d);
        a[(size_t)j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double v = a[(size_t)i * n + j];
            for (int k = 0; k < j; ++k) v -= a[(size_t)i * n + k] * a[(size_t)j * n + k];
            a[(size_t)i * n + j] = v / d;
        }
    }
    return true;
}
void cholesky_solve(const std::vector<double>& L, int n, std::vector<double>& b) {
    for (int i = 0; i < n; ++i) { double v = b[i]; for (int k = 0; k < i; ++k) v -= L[(size_t)i * n + k] * b[k]; b[i] = v / L[(size_t)i * n + i]; }
    for (int i = n - 1; i >= 0; --i) { double v = b[i]; for (int k = i + 1; k < n; ++k) v -= L[(size_t)k * n + i] * b[k]; b[i] = v / L[(size_t)i * n + i]; }
}

// the fit proper, on host copies of the pilot: rows [M][P], prof [M][K] running attempt counts, status [M]
bool fit_model(OrderModel* m, const std::vector<double>& rows, const std::vector<int>& prof, const std::vector<unsigned>& status, int M, int P, int K,
               std::vector<double>& w1, std::vector<double>& w2) {
    const int D = K - 1;
    std::vector<int> good;
    for (int i = 0; i < M; ++i) if (status[i] == 0) good.push_back(i);
    const int G = (int)good.size();
    if (G < 4 * (P + 1) || D < 1) return false;
    std::vector<float> X((size_t)G * D);
    std::vector<double> colmean(D, 0.0);
    for (int g = 0; g < G; ++g)
        for (int d = 0; d < D; ++d) { const float v = (float)(prof[(size_t)good[g] * K + d + 1] - prof[(size_t)good[g] * K + d]); X[(size_t)g * D + d] = v; colmean[d] += v; }
    for (int d = 0; d < D; ++d) colmean[d] /= G;
    for (int g = 0; g < G; ++g) for (int d = 0; d < D; ++d) X[(size_t)g * D + d] -= (float)colmean[d];
    // two leading principal components by power iteration on X^T X (the second orthogonalised against the first)
    std::vector<double> v[2], score[2];
    for (int c = 0; c < 2; ++c) {
        v[c].assign(D, 0.0);
        for (int d = 0; d < D; ++d) v[c][d] = 1.0 + 0.37 * std::sin(1.0 + d * (c + 1));
        score[c].assign(G, 0.0);
        for (int it = 0; it < 40; ++it) {
            if (c == 1) { double dot = 0; for (int d = 0; d < D; ++d) dot += v[1][d] * v[0][d]; for (int d = 0; d < D; ++d) v[1][d] -= dot * v[0][d]; }
            double nrm = 0; for (int d = 0; d < D; ++d) nrm += v[c][d] * v[c][d];
            nrm = std::sqrt(nrm);
            if (!(nrm > 0)) return false;
            for (int d = 0; d < D; ++d) v[c][d] /= nrm;
            for (int g = 0; g < G; ++g) { double s = 0; const float* x = &X[(size_t)g * D]; for (int d = 0; d < D; ++d) s += x[d] * v[c][d]; score[c][g] = s; }
            std::vector<double> nv(D, 0.0);
            for (int g = 0; g < G; ++g) { const double s = score[c][g]; const float* x = &X[(size_t)g * D]; for (int d = 0; d < D; ++d) nv[d] += s * x[d]; }
            v[c] = nv;
        }
        if (c == 1) { double dot = 0; for (int d = 0; d < D; ++d) dot += v[1][d] * v[0][d]; for (int d = 0; d < D; ++d) v[1][d] -= dot * v[0][d]; }
        double nrm = 0; for (int d = 0; d < D; ++d) nrm += v[c][d] * v[c][d];
        nrm = std::sqrt(nrm);
        if (!(nrm > 0)) { if (c == 0) return false; std::fill(score[1].begin(), score[1].end(), 0.0); continue; }
        for (int d = 0; d < D; ++d) v[c][d] /= nrm;
        for (int g = 0; g < G; ++g) { double s = 0; const float* x = &X[(size_t)g * D]; for (int d = 0; d < D; ++d) s += x[d] * v[c][d]; score[c][g] = s; }
    }
    // standardise the parameters, then least squares of both scores on [z, 1]
    m->mu.assign(P, 0.0); m->sd.assign(P, 0.0); m->inv_sd.assign(P, 0.0);
    for (int g = 0; g < G; ++g) for (int j = 0; j < P; ++j) m->mu[j] += rows[(size_t)good[g] * P + j];
    for (int j = 0; j < P; ++j) m->mu[j] /= G;
    for (int g = 0; g < G; ++g) for (int j = 0; j < P; ++j) { const double dlt = rows[(size_t)good[g] * P + j] - m->mu[j]; m->sd[j] += dlt * dlt; }
    for (int j = 0; j < P; ++j) { m->sd[j] = std::sqrt(m->sd[j] / G); m->inv_sd[j] = (m->sd[j] > 1e-300 * (1.0 + std::fabs(m->mu[j]))) ? 1.0 / m->sd[j] : 0.0; }
    const int n = P + 1;
    std::vector<double> A((size_t)n * n, 0.0), r1(n, 0.0), r2(n, 0.0), f(n);
    for (int g = 0; g < G; ++g) {
        for (int j = 0; j < P; ++j) f[j] = (rows[(size_t)good[g] * P + j] - m->mu[j]) * m->inv_sd[j];
        f[P] = 1.0;
        for (int i = 0; i < n; ++i) { for (int j = 0; j <= i; ++j) A[(size_t)i * n + j] += f[i] * f[j]; r1[i] += f[i] * score[0][g]; r2[i] += f[i] * score[1][g]; }
    }
    for (int i = 0; i < n; ++i) { A[(size_t)i * n + i] += 1e-6 * G + 1e-12; for (int j = 0; j < i; ++j) A[(size_t)j * n + i] = A[(size_t)i * n + j]; }
    if (!cholesky(A, n)) return false;
    cholesky_solve(A, n, r1); cholesky_solve(A, n, r2);
    w1 = r1; w2 = r2;
    return true;
}

sepaihrd_rc fit_from_rows(sepaihrd_ctx* ctx, OrderModel* m, double* d_rows, int M) {
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(ctx);
    cudaStream_t s = sepaihrd_internal::stream(ctx);
    const int P = d.P, K = d.K;
    double* d_ll = nullptr; unsigned* d_st = nullptr; int* d_prof = nullptr;
    ORD_TRY(cudaMalloc((void**)&d_ll, sizeof(double) * M));
    ORD_TRY(cudaMalloc((void**)&d_st, sizeof(unsigned) * M));
    ORD_TRY(cudaMalloc((void**)&d_prof, sizeof(int) * (size_t)M * K));
    sepaihrd_rc rc = sepaihrd_internal::eval_profile(ctx, d_rows, M, P, d_ll, d_st, d_prof);
    std::vector<double> rows((size_t)M * P); std::vector<int> prof((size_t)M * K); std::vector<unsigned> st(M);
    cudaError_t e = cudaSuccess;
    if (rc == SEPAIHRD_OK) {
        e = cudaMemcpyAsync(rows.data(), d_rows, sizeof(double) * rows.size(), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(prof.data(), d_prof, sizeof(int) * prof.size(), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(st.data(), d_st, sizeof(unsigned) * st.size(), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    }
    cudaFree(d_ll); cudaFree(d_st); cudaFree(d_prof);
    if (rc != SEPAIHRD_OK) return rc;
    if (e != cudaSuccess) return fail_with(SEPAIHRD_ERR_CUDA, cudaGetErrorString(e));
    std::vector<double> w1, w2;
    m->P = P; m->K = K;
    if (!fit_model(m, rows, prof, st, M, P, K, w1, w2)) { m->fitted = false; return SEPAIHRD_OK; }      // degenerate pilot: launches stay unordered
    std::vector<double> blob;
    blob.insert(blob.end(), m->mu.begin(), m->mu.end());
    blob.insert(blob.end(), m->inv_sd.begin(), m->inv_sd.end());
    blob.insert(blob.end(), w1.begin(), w1.end());
    blob.insert(blob.end(), w2.begin(), w2.end());
    if (!m->d_model) ORD_TRY(cudaMalloc((void**)&m->d_model, sizeof(double) * blob.size()));
    ORD_TRY(cudaMemcpyAsync(m->d_model, blob.data(), sizeof(double) * blob.size(), cudaMemcpyHostToDevice, s));
    ORD_TRY(cudaStreamSynchronize(s));
    m->fitted = true;
    m->fits += 1;
    return SEPAIHRD_OK;
}

}  // namespace

namespace sepaihrd_internal {

void order_release(sepaihrd_ctx* ctx) {
    OrderModel* m = model_of(ctx, false);
    if (!m) return;
    cudaDeviceSynchronize();
    if (m->d_model) cudaFree(m->d_model);
    for (OrderModel::Work& w : m->work) {
        for (void* p : {(void*)w.d_keys, (void*)w.d_perm, (void*)w.d_bucket, (void*)w.d_hist, (void*)w.d_stats}) if (p) cudaFree(p);
        if (w.done) cudaEventDestroy(w.done);
    }
    delete m;
    *order_slot(ctx) = nullptr;
}

sepaihrd_rc order_batch(sepaihrd_ctx* ctx, const double* d_params, long long B, long long ld, const int** perm) {
    OrderModel* m = model_of(ctx, false);
    if (!m || !m->fitted || order_mode(ctx) == 0 || !order_applicable(ctx) || B < ORDER_MIN_BATCH || B > 0x7fffffffLL) return SEPAIHRD_OK;
    const Dims d = dims(ctx);
    if (m->P != d.P) return SEPAIHRD_OK;
    const int which = (int)(m->seq++ & 1u);
    OrderModel::Work& w = m->work[which];
    sepaihrd_rc rc = ensure_buffers(w, B);
    if (rc != SEPAIHRD_OK) return rc;
    cudaStream_t s = stream(ctx);
    if (w.pending) ORD_TRY(cudaStreamWaitEvent(s, w.done, 0));        // the launch that last read this set's index list (maybe on another stream)
    ORD_TRY(cudaMemsetAsync(w.d_stats, 0, sizeof(double) * 4, s));
    ORD_TRY(cudaMemsetAsync(w.d_hist, 0, sizeof(unsigned) * (NBUCKETS + 1), s));
    order_keys_kernel<<<(unsigned)((B + 7) / 8), 256, 0, s>>>(d_params, B, ld, d.P, m->d_model, w.d_keys, w.d_stats);
    order_bucket_kernel<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(w.d_keys, B, w.d_stats, w.d_bucket, w.d_hist);
    order_scan_kernel<<<1, 1024, 0, s>>>(w.d_hist);
    order_scatter_kernel<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(w.d_bucket, B, w.d_hist, w.d_perm);
    ORD_TRY(cudaGetLastError());
    count_launches(ctx, 4);
    *perm = w.d_perm;
    m->last = which;
    return SEPAIHRD_OK;
}

// the launch that reads the index list of the latest order_batch has been enqueued on the ctx stream
sepaihrd_rc order_mark_done(sepaihrd_ctx* ctx) {
    OrderModel* m = model_of(ctx, false);
    if (!m || m->last < 0) return SEPAIHRD_OK;
    OrderModel::Work& w = m->work[m->last];
    m->last = -1;
    ORD_TRY(cudaEventRecord(w.done, stream(ctx)));
    w.pending = true;
    return SEPAIHRD_OK;
}

}  // namespace sepaihrd_internal

extern "C" {

sepaihrd_rc sepaihrd_set_ordering(sepaihrd_ctx* ctx, int32_t mode) {
    if (!ctx || (mode != 0 && mode != 1)) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "bad ordering mode");
    const auto ctx_lock = sepaihrd_internal::lock(ctx);
    sepaihrd_internal::set_order_mode(ctx, mode);
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_fit_ordering(sepaihrd_ctx* ctx, const double* params, int64_t B, int64_t ld, int32_t params_on_device) {
    if (!ctx || !params) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(ctx);
    if (B < 1 || ld < d.P) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "Parameter vector size mismatch.");
    const auto ctx_lock = sepaihrd_internal::lock(ctx);
    ORD_TRY(cudaSetDevice(d.device));
    if (!sepaihrd_internal::order_applicable(ctx)) return SEPAIHRD_OK;   // no profiling instantiation (16 lanes per set, STRICT arithmetic): such launches stay unordered
    OrderModel* m = model_of(ctx, true);
    cudaStream_t s = sepaihrd_internal::stream(ctx);
    const int M = (int)std::min<int64_t>(PILOT, B);
    double* d_rows = nullptr;
    ORD_TRY(cudaMalloc((void**)&d_rows, sizeof(double) * (size_t)M * d.P));
    sepaihrd_rc rc = SEPAIHRD_OK;
    if (params_on_device) {
        order_gather_kernel<<<M, 64, 0, s>>>(params, B, ld, M, d.P, d_rows);
        if (cudaGetLastError() != cudaSuccess) rc = fail_with(SEPAIHRD_ERR_CUDA, "ordering pilot gather failed");
    } else {
        std::vector<double> rows((size_t)M * d.P);
        for (int i = 0; i < M; ++i) std::memcpy(&rows[(size_t)i * d.P], params + (((int64_t)i * B) / M) * ld, sizeof(double) * d.P);
        if (cudaMemcpyAsync(d_rows, rows.data(), sizeof(double) * rows.size(), cudaMemcpyHostToDevice, s) != cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess) rc = fail_with(SEPAIHRD_ERR_CUDA, "ordering pilot copy failed");
    }
    if (rc == SEPAIHRD_OK) rc = fit_from_rows(ctx, m, d_rows, M);
    cudaFree(d_rows);
    return rc;
}

sepaihrd_rc sepaihrd_ordering_state(const sepaihrd_ctx* ctx, int32_t* out_fitted, int64_t* out_fits) {
    if (!ctx) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null ctx");
    const OrderModel* m = static_cast<const OrderModel*>(*sepaihrd_internal::order_slot(const_cast<sepaihrd_ctx*>(ctx)));
    if (out_fitted) *out_fitted = (m && m->fitted) ? 1 : 0;
    if (out_fits) *out_fits = m ? m->fits : 0;
    return SEPAIHRD_OK;
}

}  // extern "C"

namespace sepaihrd_internal {

// Host-pointer evaluations fit by themselves: no model yet, or the batch's parameters sit elsewhere / are spread differently than
// the pilot the model came from (a strided sample of 2048 rows: mean moved by more than half a pilot sigma, or sigma changed 2x).
sepaihrd_rc order_autofit_host(sepaihrd_ctx* ctx, const double* params, long long B, long long ld) {
    if (order_mode(ctx) == 0 || B < ORDER_MIN_BATCH) return SEPAIHRD_OK;
    const Dims d = dims(ctx);
    if (!order_applicable(ctx)) return SEPAIHRD_OK;
    OrderModel* m = model_of(ctx, true);
    bool refit = !m->fitted && m->fits == 0;
    if (m->fitted) {
        const int S = 1024;
        std::vector<double> mean(d.P, 0.0), var(d.P, 0.0);
        for (int i = 0; i < S; ++i) { const double* r = params + (((long long)i * B) / S) * ld; for (int j = 0; j < d.P; ++j) mean[j] += r[j]; }
        for (int j = 0; j < d.P; ++j) mean[j] /= S;
        for (int i = 0; i < S; ++i) { const double* r = params + (((long long)i * B) / S) * ld; for (int j = 0; j < d.P; ++j) { const double t = r[j] - mean[j]; var[j] += t * t; } }
        for (int j = 0; j < d.P && !refit; ++j) {
            const double sd = std::sqrt(var[j] / S), ref = m->sd[j];
            if (ref <= 0.0 && sd <= 0.0) continue;
            if (std::fabs(mean[j] - m->mu[j]) > 0.5 * std::max(ref, sd) || sd > 2.0 * ref || ref > 2.0 * sd) refit = true;
        }
    }
    return refit ? sepaihrd_fit_ordering(ctx, params, B, ld, 0) : SEPAIHRD_OK;
}

}  // namespace sepaihrd_internal
