// sepaihrd_capi.cu -- C ABI (include/sepaihrd_b200.h) over the fused CUDA kernels.
// No torch types, no CPU fallback: every compute entry point needs a CUDA device.

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#include "sepaihrd_internal.h"
#include "sepaihrd_kernels.cuh"
#ifdef SEPAIHRD_WITH_SPLIT      // experiment, not part of the shipped library (tools/build_variant.sh split -DSEPAIHRD_WITH_SPLIT)
#include "experiments/sepaihrd_split.cuh"
#endif

#ifndef SEPAIHRD_SPLIT_THREADS
#define SEPAIHRD_SPLIT_THREADS 384                 // 6 warp pairs per block, one block per SM: <= 170 registers per thread
#endif
#ifndef SEPAIHRD_E2E_DEFAULT_SPLIT
// Chunk ends of the host-buffer evaluation at B/d for each d listed (descending): every chunk is 3x the one before it.  The kernel
// consumes a set in ~83 ns; a PCIe 5 x16 link alone delivers one in ~9 ns, but with 8 ranks copying at once a rank gets 23-35 GB/s
// (~20 ns per set, measured: bench.py e2e.h2d_gbs_per_rank_all_ranks_copying), so a chunk's copy hides under the previous chunk's
// kernel only if it is at most ~4x larger.  Round 1's split (128, 32, 8: last chunk 9x the one before) was right for one GPU and
// exposed 5-10 ms of copy per step at 8 GPUs; this one measures 87.60 ms per 1M-set step on one GPU (87.76 before).
#define SEPAIHRD_E2E_DEFAULT_SPLIT 243, 81, 27, 9, 3
#endif

namespace {

thread_local std::string g_last_error;

sepaihrd_rc fail(sepaihrd_rc rc, const std::string& msg) {
    g_last_error = msg;
    return rc;
}

#define CUDA_TRY(expr)                                                                                 \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return fail(_e == cudaErrorMemoryAllocation ? SEPAIHRD_ERR_OUT_OF_MEMORY : SEPAIHRD_ERR_CUDA, \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                           \
    } while (0)

// ---- slot layout ------------------------------------------------------------------------------------
struct Layout {
    int n, nb, nk;
    int scal0() const { return nb + nk; }
    int age0() const { return scal0() + 7; }
    int mult0() const { return age0() + 8 * n; }
    int count() const { return mult0() + 11; }
};

bool has_prefix(const char* s, const char* prefix, size_t* len) {
    size_t l = std::strlen(prefix);
    *len = l;
    return std::strncmp(s, prefix, l) == 0;
}

// decimal index after a prefix, as std::stoul would read it (digits required)
bool read_index(const char* s, long* out) {
    if (*s < '0' || *s > '9') return false;
    char* end = nullptr;
    *out = std::strtol(s, &end, 10);
    return end != s;
}

}  // namespace

// One host-pointer evaluation request waiting to be served (sepaihrd_eval_batch): concurrent callers of one ctx are merged
// into ONE launch by whichever of them currently leads.
struct EvalRequest {
    const double* params; int64_t B, ld;
    double* out_ll; uint32_t* out_status; int32_t* out_steps;
    sepaihrd_rc rc = SEPAIHRD_OK;
    std::string error;
    bool done = false, lead = false;     // served / told to lead the next round
    std::condition_variable cv;          // its owner sleeps here: a round wakes exactly the owners it served and ONE next leader
};

struct sepaihrd_ctx {
    // Every entry point that touches the ctx takes `mu` (recursive: the host-pointer calls go through the device-pointer ones):
    // the reference's objective is called concurrently from OpenMP loops (ParticleSwarmOptimizer.cpp:368-424,
    // HillClimbingOptimizer.cpp:228-234), so a drop-in behind calculate() has to tolerate that.
    std::recursive_mutex mu;
    // request coalescing of sepaihrd_eval_batch
    std::mutex q_mu;
    std::vector<EvalRequest*> queue;
    bool leader = false;
    double* h_stage = nullptr; size_t cap_stage = 0;          // pinned: packed rows | logL | status | steps of a merged launch
    double* h_small = nullptr; size_t cap_small = 0;          // pinned: the same four pieces of ONE small request with pageable buffers
    long long merged_launches = 0, merged_requests = 0;
    int device = 0;
    int n_user = 0;                // the caller's age-class count; n below is what the kernels run with (4 or 16, zero-padded classes)
    int n = 0, K = 0, n_obs = 0, nb = 0, nk = 0, P = 0, nslots = 0, nseg = 0, runup_offset = 0, n_nonneg = 0;
    int constraint_mode = 0, math_mode = SEPAIHRD_MATH_FAST;
    bool obs_mismatch = false;
    double abs_tol = 1e-6, rel_tol = 1e-6, dt_hint = 1.0, hmax = 1.0;
    bool bp_on_grid = false;       // no schedule breakpoint strictly inside an output interval
    static constexpr int N_SCRATCH = 16;
    void* scratch[N_SCRATCH] = {};        // grow-only work buffers of the aggregation passes (sepaihrd_internal::scratch)
    size_t scratch_bytes[N_SCRATCH] = {};
    std::vector<double> blob;   // host image
    sepaihrd::KParams kp{};     // offsets etc. (I/O fields filled per call)
    double* d_blob = nullptr;
    static constexpr unsigned N_TILE_COUNTERS = 64;
    unsigned* d_tile_counter = nullptr;   // ring of N_TILE_COUNTERS work counters, one per launch in flight
    unsigned launch_seq = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // host-pointer entry points: H2D / early D2H next to the compute stream
    cudaStream_t stream2 = nullptr;       // host-pointer evaluation: odd chunks run here, so a chunk's first blocks start while the previous chunk's last warps drain
    static constexpr int MAX_CHUNKS = 8;
    cudaEvent_t ev_chunk[MAX_CHUNKS] = {};
    cudaEvent_t ev_copy[MAX_CHUNKS] = {};
    cudaStream_t stream = nullptr;
    int num_sms = 0;
    // scratch for the host-pointer entry points (grown on demand)
    double* d_params = nullptr; size_t cap_params = 0;
    double* d_out = nullptr; size_t cap_out = 0;
    unsigned* d_status = nullptr; size_t cap_status = 0;
    int* d_steps = nullptr; size_t cap_steps = 0;
    long long launches = 0, sets = 0;
    double* dbg_trace = nullptr;          // SEPAIHRD_DEBUG_INTERVALS builds only
    void* order_model = nullptr;          // sepaihrd_order.cu: the fitted predictor + work buffers of the ordering pass
    int order_mode = 1;                   // 0 off, 1 on (when a model has been fitted and the batch is large enough)
};

namespace {

template <class T>
sepaihrd_rc grow(T** ptr, size_t* cap, size_t need) {
    if (need <= *cap) return SEPAIHRD_OK;
    if (*ptr) { cudaDeviceSynchronize(); cudaFree(*ptr); }   // nothing queued on ANY stream may still use the old buffer
    *ptr = nullptr; *cap = 0;
    CUDA_TRY(cudaMalloc((void**)ptr, need * sizeof(T)));
    *cap = need;
    return SEPAIHRD_OK;
}

struct LaunchCfg { int threads, minblocks; };

template <int NA, bool STRICT, int MODE, int THREADS, int MINBLOCKS, int LOOP, bool ONGRID = false, bool PROFILE = false>
sepaihrd_rc launch_t(sepaihrd_ctx* ctx, const sepaihrd::KParams& kp_in) {
    using namespace sepaihrd;
    KParams kp = kp_in;
    constexpr int SETS = THREADS / NA;
    constexpr int WSETS = (32 / NA) > 0 ? (32 / NA) : 1;
    // A warp pays the day-by-day MAXIMUM of its sets' attempts.  While there are fewer sets than warp slots with a scheduler of
    // their own (4 per SM), a warp takes fewer sets per tile, down to one: 8 sets cost what 1 set costs (0.51 ms, not 0.57).
    kp.sets_per_tile = (int)std::min<long long>(WSETS, std::max<long long>(1, (kp.B + 4LL * ctx->num_sms - 1) / (4LL * ctx->num_sms)));
    {   // experiments: SEPAIHRD_SETS_PER_TILE=n forces the tile width of launches without an index list
        static const int forced = std::getenv("SEPAIHRD_SETS_PER_TILE") ? std::atoi(std::getenv("SEPAIHRD_SETS_PER_TILE")) : 0;
        if (forced > 0) kp.sets_per_tile = std::min(forced, WSETS);
    }
    if (PROFILE || kp.perm) kp.sets_per_tile = WSETS;
    kp.tiles = (kp.B + kp.sets_per_tile - 1) / kp.sets_per_tile;
    if (kp.tiles > 0xffff0000LL) return fail(SEPAIHRD_ERR_UNSUPPORTED, "batch too large for one launch");
    // Every launch draws its tiles from its OWN counter (a ring of N_TILE_COUNTERS, zeroed on the launching stream right before
    // the kernel): launches of one ctx that are in flight on different streams never share one.
    kp.tile_counter = ctx->d_tile_counter + (ctx->launch_seq++ % sepaihrd_ctx::N_TILE_COUNTERS);
    CUDA_TRY(cudaMemsetAsync(kp.tile_counter, 0, sizeof(unsigned), ctx->stream));
    auto kern = sepaihrd_batch_kernel<NA, STRICT, MODE, THREADS, MINBLOCKS, LOOP, ONGRID, PROFILE>;
    const size_t smem = (size_t)kp.blob_bytes + sizeof(double) * (SETS * (size_t)(kp.slot_stride + ((kp.seg_stride + 1) & ~1)) + 2 * THREADS + (size_t)NA * THREADS) + 64;
    {   // once per device and instantiation, whichever thread / ctx gets here first
        static std::once_flag attr_once[64];
        cudaError_t attr_err = cudaSuccess;
        std::call_once(attr_once[ctx->device & 63], [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
        CUDA_TRY(attr_err);
    }
    if (smem > 200 * 1024) return fail(SEPAIHRD_ERR_UNSUPPORTED, "problem constants do not fit in shared memory");
    int occ = 0;
    {   // asked once per (instantiation, device, shared-memory size): a small launch should not pay for the query every time
        struct OccCache { int device = -1; size_t smem = 0; int occ = 0; };
        static thread_local OccCache cache;
        if (cache.device != ctx->device || cache.smem != smem) {
            int o = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, THREADS, smem));
            cache.device = ctx->device; cache.smem = smem; cache.occ = o;
        }
        occ = cache.occ;
    }
    if (occ < 1) return fail(SEPAIHRD_ERR_CUDA, "kernel does not fit on an SM");
    // one block per tile until the machine is full; a launch smaller than the machine uses only as many warps per block as it needs
    long long grid = std::min<long long>(kp.tiles, (long long)ctx->num_sms * occ);
    if (grid < 1) grid = 1;
    kp.active_warps = (int)std::min<long long>(THREADS / 32, (kp.tiles + grid - 1) / grid);
    kern<<<(unsigned)grid, THREADS, smem, ctx->stream>>>(kp);
    CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    ctx->sets += kp.B;
    return SEPAIHRD_OK;
}

template <int NA, int THREADS, int MINBLOCKS>
sepaihrd_rc launch_na(sepaihrd_ctx* ctx, const sepaihrd::KParams& kp, int mode) {
    using namespace sepaihrd;
    const bool strict = ctx->math_mode == SEPAIHRD_MATH_STRICT;
    if (strict)   // STRICT keeps the reference-order loop (the posterior-predictive series come out of its trajectory instantiation)
        return (mode == MODE_LL) ? launch_t<NA, true, MODE_LL, THREADS, MINBLOCKS, 5>(ctx, kp)
                                 : launch_t<NA, true, MODE_TRAJ, THREADS, MINBLOCKS, 5>(ctx, kp);
#ifndef SEPAIHRD_EXP_NO_ONGRID
    if (ctx->bp_on_grid && ctx->math_mode != SEPAIHRD_MATH_FAST_GENERAL) {   // breakpoints on output-grid points (e.g. Spain 2020): the build without the mixed-segment attempt body
        if (mode == MODE_PPC) return launch_t<NA, false, MODE_PPC, THREADS, MINBLOCKS, 6, true>(ctx, kp);
        return (mode == MODE_LL) ? launch_t<NA, false, MODE_LL, THREADS, MINBLOCKS, 6, true>(ctx, kp)
                                 : launch_t<NA, false, MODE_TRAJ, THREADS, MINBLOCKS, 6, true>(ctx, kp);
    }
#endif
    if (mode == MODE_PPC) return launch_t<NA, false, MODE_PPC, THREADS, MINBLOCKS, 6>(ctx, kp);
    return (mode == MODE_LL) ? launch_t<NA, false, MODE_LL, THREADS, MINBLOCKS, 6>(ctx, kp)
                             : launch_t<NA, false, MODE_TRAJ, THREADS, MINBLOCKS, 6>(ctx, kp);
}

#ifdef SEPAIHRD_WITH_SPLIT
// The warp-pair kernel (experiments/sepaihrd_split.cuh): same arithmetic, half the state per lane.
template <int THREADS>
sepaihrd_rc launch_split(sepaihrd_ctx* ctx, const sepaihrd::KParams& kp_in) {
    using namespace sepaihrd;
    KParams kp = kp_in;
    constexpr int PAIRS = THREADS / 64, SETS = PAIRS * 8;
    kp.tiles = (kp.B + 7) / 8;
    if (kp.tiles > 0xffff0000LL) return fail(SEPAIHRD_ERR_UNSUPPORTED, "batch too large for one launch");
    kp.tile_counter = ctx->d_tile_counter + (ctx->launch_seq++ % sepaihrd_ctx::N_TILE_COUNTERS);
    CUDA_TRY(cudaMemsetAsync(kp.tile_counter, 0, sizeof(unsigned), ctx->stream));
    auto kern = sepaihrd_split_kernel<THREADS>;
    const size_t smem = (size_t)kp.blob_bytes + sizeof(double) * (SETS * (size_t)(kp.slot_stride + ((kp.seg_stride + 1) & ~1)) + 2 * THREADS + (size_t)4 * THREADS) +
                        PAIRS * sizeof(PairBox) + 16;
    {
        static std::once_flag attr_once[64];
        cudaError_t attr_err = cudaSuccess;
        std::call_once(attr_once[ctx->device & 63], [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
        CUDA_TRY(attr_err);
    }
    if (smem > 200 * 1024) return fail(SEPAIHRD_ERR_UNSUPPORTED, "problem constants do not fit in shared memory");
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
    if (occ < 1) return fail(SEPAIHRD_ERR_CUDA, "kernel does not fit on an SM");
    long long grid = std::min<long long>((kp.tiles + PAIRS - 1) / PAIRS, (long long)ctx->num_sms * occ);
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, THREADS, smem, ctx->stream>>>(kp);
    CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    ctx->sets += kp.B;
    return SEPAIHRD_OK;
}

#endif

sepaihrd_rc launch(sepaihrd_ctx* ctx, const sepaihrd::KParams& kp, int mode) {
    if (ctx->math_mode == SEPAIHRD_MATH_FAST_SPLIT) {
#ifdef SEPAIHRD_WITH_SPLIT
        if (!(ctx->n == 4 && mode == sepaihrd::MODE_LL && ctx->bp_on_grid))
            return fail(SEPAIHRD_ERR_UNSUPPORTED, "the warp-pair kernel covers log-likelihoods of 4-age problems with breakpoints on grid days");
        return launch_split<SEPAIHRD_SPLIT_THREADS>(ctx, kp);
#else
        return fail(SEPAIHRD_ERR_UNSUPPORTED, "SEPAIHRD_MATH_FAST_SPLIT needs an experimental build (-DSEPAIHRD_WITH_SPLIT): measured slower than FAST, not shipped");
#endif
    }
    switch (ctx->n) {
#ifdef SEPAIHRD_EXP_TWO_BLOCKS
        case 4: return launch_na<4, 128, 2>(ctx, kp, mode);      // round 1's shape: two blocks of four warps per SM
#else
        case 4: return launch_na<4, 256, 1>(ctx, kp, mode);      // v16: ONE block of eight warps per SM (the constants are staged once per SM): -1.1 % time
#endif
        case 16: return launch_na<16, 256, 1>(ctx, kp, mode);   // the 16-age observation block (117 KB) allows one block per SM: make it 8 warps
        default: return fail(SEPAIHRD_ERR_UNSUPPORTED, "GPU kernels are instantiated for 4 and 16 lanes per set (sepaihrd_create pads other age-class counts)");
    }
}

__global__ void fill_kernel(double* ll, unsigned* st, int* steps, long long B, double v, unsigned s) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < B) {
        ll[i] = v;
        if (st) st[i] = s;
        if (steps) { steps[2 * i] = 0; steps[2 * i + 1] = 0; }
    }
}

// ---- FP64 pipe microbenchmark --------------------------------------------------------------------------
constexpr int PEAK_CHAINS = 8, PEAK_ITERS = 4096;
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, double a, double b) {
    double v[PEAK_CHAINS];
#pragma unroll
    for (int c = 0; c < PEAK_CHAINS; ++c) v[c] = a + c + threadIdx.x * 1e-9;
#pragma unroll 1
    for (int it = 0; it < PEAK_ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < PEAK_CHAINS; ++c) v[c] = fma(v[c], b, a);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < PEAK_CHAINS; ++c) s += v[c];
    if (s == 123.456) out[0] = s;   // never true: keeps the chains alive
}

}  // namespace

namespace {

// A copy of a problem description with its age classes padded to `np` (see sepaihrd_create).
struct PaddedProblem {
    sepaihrd_problem pb{};
    std::vector<double> obs_h, obs_i, obs_d, pop, M, base, init;
    std::vector<int32_t> slot;
    static int map_slot(int s, const Layout& from, const Layout& to) {
        if (s < from.age0()) return s;                                  // schedules and scalar rates: same place
        if (s < from.mult0()) {                                          // per-age blocks a, h_infec, p, h, icu, d_H, d_ICU, d_community
            const int k = (s - from.age0()) / from.n, age = (s - from.age0()) % from.n;
            return to.age0() + k * to.n + age;
        }
        return to.mult0() + (s - from.mult0());                          // multipliers, seed_exposed, runup_days, beta
    }
    void build(const sepaihrd_problem& src, int np) {
        pb = src;
        const int n = src.n_ages;
        const Layout from{n, src.n_beta, src.n_kappa}, to{np, src.n_beta, src.n_kappa};
        auto pad_rows = [&](const double* a, std::vector<double>& out) {   // [n_obs][n] -> [n_obs][np], padding = skipped observation
            out.assign((size_t)src.n_obs * np, -1.0);
            for (int r = 0; r < src.n_obs; ++r) for (int j = 0; j < n; ++j) out[(size_t)r * np + j] = a[(size_t)r * n + j];
        };
        pad_rows(src.obs_hosp, obs_h); pad_rows(src.obs_icu, obs_i); pad_rows(src.obs_deaths, obs_d);
        pop.assign(np, 0.0);
        for (int j = 0; j < n; ++j) pop[j] = src.population[j];
        M.assign((size_t)np * np, 0.0);                                    // column-major M(i, j) = M[j * n + i]
        for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) M[(size_t)j * np + i] = src.contact_matrix[(size_t)j * n + i];
        base.assign(to.count(), 0.0);
        for (int s = 0; s < from.count(); ++s) base[map_slot(s, from, to)] = src.base_slots[s];
        init.assign((size_t)SEPAIHRD_NUM_COMPARTMENTS * np, 0.0);
        for (int cpt = 0; cpt < SEPAIHRD_NUM_COMPARTMENTS; ++cpt) for (int j = 0; j < n; ++j) init[(size_t)cpt * np + j] = src.data_initial_state[(size_t)cpt * n + j];
        slot.resize(src.n_params);
        for (int i = 0; i < src.n_params; ++i) slot[i] = src.param_slot[i] < 0 ? src.param_slot[i] : map_slot(src.param_slot[i], from, to);
        pb.n_ages = np;
        pb.obs_hosp = obs_h.data(); pb.obs_icu = obs_i.data(); pb.obs_deaths = obs_d.data();
        pb.population = pop.data(); pb.contact_matrix = M.data(); pb.base_slots = base.data();
        pb.data_initial_state = init.data(); pb.param_slot = slot.data();
    }
};

sepaihrd_rc create_impl(const sepaihrd_problem* pb, int n_user, int n, int32_t device, sepaihrd_ctx** out_ctx);

// `count` compartment-major state vectors (C compartments) from n_in to n_out age classes: extra classes are dropped / zero-filled
__global__ void repack_ages_kernel(const double* __restrict__ in, long long in_stride, double* __restrict__ out, long long out_stride,
                                   long long count, int C, int n_in, int n_out) {
    const long long per = (long long)C * n_out, total = count * per;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long v = i / per;
        const int r = (int)(i % per), cpt = r / n_out, age = r % n_out;
        out[v * out_stride + r] = (age < n_in) ? in[v * in_stride + (long long)cpt * n_in + age] : 0.0;
    }
}

}  // namespace

extern "C" {

int32_t sepaihrd_slot_count(int32_t n, int32_t nb, int32_t nk) { return Layout{n, nb, nk}.count(); }

int32_t sepaihrd_slot_for_name(int32_t n, int32_t nb, int32_t nk, const char* name) {
    if (!name) return -1;
    const Layout L{n, nb, nk};
    size_t len = 0;
    long idx = 0;
    // exact scalar names first where the reference tests equality before prefixes
    if (std::strcmp(name, "beta") == 0) return L.mult0() + 10;
    if (has_prefix(name, "beta_", &len)) {
        if (!read_index(name + len, &idx) || idx < 1 || idx > nb) return -2;
        return (int)idx - 1;
    }
    static const char* scalars[7] = {"theta", "sigma", "gamma_p", "gamma_A", "gamma_I", "gamma_H", "gamma_ICU"};
    for (int s = 0; s < 7; ++s)
        if (std::strcmp(name, scalars[s]) == 0) return L.scal0() + s;
    // per-age blocks, in the reference's if/else order (a_, h_infec_, p_, h_, icu_, d_H_, d_ICU_, d_community_)
    static const char* blocks[8] = {"a_", "h_infec_", "p_", "h_", "icu_", "d_H_", "d_ICU_", "d_community_"};
    for (int b = 0; b < 8; ++b)
        if (has_prefix(name, blocks[b], &len)) {
            if (!read_index(name + len, &idx) || idx >= n) return -2;
            return L.age0() + b * n + (int)idx;
        }
    if (std::strcmp(name, "seed_exposed") == 0) return L.mult0() + 8;
    if (std::strcmp(name, "runup_days") == 0) return L.mult0() + 9;
    static const char* mult[8] = {"E0_multiplier", "P0_multiplier", "A0_multiplier", "I0_multiplier",
                                  "H0_multiplier", "ICU0_multiplier", "R0_multiplier", "D0_multiplier"};
    for (int m = 0; m < 8; ++m)
        if (std::strcmp(name, mult[m]) == 0) return L.mult0() + m;
    if (has_prefix(name, "kappa_", &len)) {
        if (!read_index(name + len, &idx) || idx < 2 || idx > nk) return -2;   // kappa_1 is the fixed baseline
        return nb + (int)idx - 1;
    }
    return -1;
}

const char* sepaihrd_last_error(void) { return g_last_error.c_str(); }
const char* sepaihrd_version(void) { return "sepaihrd_b200 0.1.0 (sm_100a)"; }

sepaihrd_rc sepaihrd_create(const sepaihrd_problem* pb, int32_t device, sepaihrd_ctx** out_ctx) {
    if (!pb || !out_ctx) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    *out_ctx = nullptr;
    if (pb->abi_version != SEPAIHRD_ABI_VERSION) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "ABI version mismatch");
    const int n = pb->n_ages, K = pb->n_times, nb = pb->n_beta, nk = pb->n_kappa, P = pb->n_params;
    if (n < 1 || n > SEPAIHRD_MAX_AGES) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "n_ages out of range");
    if (K < 1) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "Output time points vector cannot be empty.");   // Simulator.cpp:70-72
    for (int i = 1; i < K; ++i)
        if (!(pb->times[i] > pb->times[i - 1]))
            return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "Output time points must be strictly increasing.");   // Simulator.cpp:82-90
    if (nk < 1) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "NPI strategy needs at least the baseline period");
    if (pb->kappa_end_times[0] < 0.0)
        return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "Baseline period end time must be non-negative.");   // NPI.cpp:24-26
    for (int k = 1; k < nk; ++k)
        if (!(pb->kappa_end_times[k] > pb->kappa_end_times[k - 1]))
            return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "NPI end times must be strictly increasing.");   // NPI.cpp:36-46
    for (int k = 1; k < nb; ++k)
        if (!(pb->beta_end_times[k] > pb->beta_end_times[k - 1]))
            return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "beta end times must be strictly increasing.");  // PiecewiseConstantParameterStrategy.cpp:22-34
    if (P < 1) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "Parameter names list cannot be empty.");      // ParameterManager.cpp:26-28
    if (pb->abs_tol < 0 || pb->rel_tol < 0) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "Error tolerances cannot be negative.");   // Simulator.cpp:46-52
    if (!(pb->dt_hint > 0)) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "Time step hint must be positive.");   // Simulator.cpp:39-41
    {
        const Layout Lu{n, nb, nk};
        for (int i = 0; i < P; ++i)
            if (pb->param_slot[i] < -1 || pb->param_slot[i] >= Lu.count())
                return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "param_slot out of range (use sepaihrd_slot_for_name)");
    }
    // The kernels are instantiated for 4 and 16 lanes per parameter set.  Any other age-class count runs in the next larger one
    // with EMPTY extra classes: population 0 (hence 1/N = 0 and age fraction 0, see below), zero contact rows and columns, zero
    // rates, zero initial state, observations marked as skipped.  Such a class has zero derivatives, a zero error estimate (it
    // never decides a step) and contributes no likelihood term, so the real classes compute exactly what they would alone.
    const int n_user = n;
    PaddedProblem padded;
    if (n != 4 && n != 16) {
        padded.build(*pb, n <= 4 ? 4 : 16);
        pb = &padded.pb;
    }
    const int n_run = pb->n_ages;
    return create_impl(pb, n_user, n_run, device, out_ctx);
}

}  // extern "C"

namespace {

sepaihrd_rc create_impl(const sepaihrd_problem* pb, int n_user, int n, int32_t device, sepaihrd_ctx** out_ctx) {
    const int K = pb->n_times, nb = pb->n_beta, nk = pb->n_kappa, P = pb->n_params;
    const Layout L{n, nb, nk};

    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1)
        return fail(SEPAIHRD_ERR_NO_DEVICE, "no CUDA device: sepaihrd_b200 has no CPU fallback");
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= count) return fail(SEPAIHRD_ERR_NO_DEVICE, "CUDA device index out of range");
    CUDA_TRY(cudaSetDevice(device));

    sepaihrd_ctx* ctx = new sepaihrd_ctx();
    ctx->device = device;
    ctx->n_user = n_user;
    ctx->n = n; ctx->K = K; ctx->n_obs = pb->n_obs; ctx->nb = nb; ctx->nk = nk; ctx->P = P; ctx->nslots = L.count();
    ctx->constraint_mode = pb->constraint_mode; ctx->abs_tol = pb->abs_tol; ctx->rel_tol = pb->rel_tol; ctx->dt_hint = pb->dt_hint;
    // runup_offset_ = first index with t >= 0 (ObjectiveFunction.cpp:39-46)
    ctx->runup_offset = 0;
    for (int i = 0; i < K; ++i) if (pb->times[i] >= 0.0) { ctx->runup_offset = i; break; }
    ctx->n_nonneg = 0;
    for (int i = 0; i < K; ++i) ctx->n_nonneg += (pb->times[i] >= 0.0) ? 1 : 0;
    ctx->obs_mismatch = (K - ctx->runup_offset) != pb->n_obs;   // calculate() then returns lowest() (.cpp:176-178)
    ctx->hmax = 0.0;
    for (int i = 1; i < K; ++i) ctx->hmax = std::max(ctx->hmax, pb->times[i] - pb->times[i - 1]);

    // merged schedule segments: segment s = number of breakpoints strictly below t ("t <= end_k" picks k)
    std::vector<double> bp;
    for (int k = 0; k < nb; ++k) bp.push_back(pb->beta_end_times[k]);
    for (int k = 0; k < nk; ++k) bp.push_back(pb->kappa_end_times[k]);
    std::sort(bp.begin(), bp.end());
    bp.erase(std::unique(bp.begin(), bp.end()), bp.end());
    const int nseg = (int)bp.size();
    if (nseg > SEPAIHRD_MAX_SEGMENTS) { delete ctx; return fail(SEPAIHRD_ERR_UNSUPPORTED, "too many schedule breakpoints"); }
    ctx->nseg = nseg;
    // a breakpoint strictly inside an output interval can split the stages of a step over two segments; one on a grid
    // point (or outside the window) cannot, because no step crosses a grid point
    ctx->bp_on_grid = true;
    for (double b : bp)
        for (int i = 0; i + 1 < K; ++i)
            if (b > pb->times[i] && b < pb->times[i + 1]) ctx->bp_on_grid = false;
    std::vector<int> segb(nseg + 1), segk(nseg + 1);
    for (int s = 0; s <= nseg; ++s) {
        const double tr = (s < nseg) ? bp[s] : bp[nseg - 1] + 1.0;
        int ib = 0, ik = 0;
        for (int k = 0; k < nb; ++k) if (tr > pb->beta_end_times[k]) ++ib;
        for (int k = 0; k < nk; ++k) if (tr > pb->kappa_end_times[k]) ++ik;
        segb[s] = std::min(ib, std::max(nb - 1, 0));
        segk[s] = std::min(ik, nk - 1);
    }

    // later duplicates of a slot win in updateModelParameters' loop: disable the earlier ones
    std::vector<int> pslot(P);
    for (int i = 0; i < P; ++i) pslot[i] = pb->param_slot[i];
    for (int i = 0; i < P; ++i)
        for (int j = i + 1; j < P; ++j)
            if (pslot[i] >= 0 && pslot[i] == pslot[j]) { pslot[i] = -1; break; }

    // ---- pack the constants blob ---------------------------------------------------------------------
    std::vector<double>& B = ctx->blob;
    auto put = [&](const double* src, size_t cnt) { int off = (int)B.size(); B.insert(B.end(), src, src + cnt); return off; };
    auto put_ints = [&](const int* src, size_t cnt) {
        int off = (int)B.size();
        B.resize(B.size() + (cnt + 1) / 2, 0.0);
        std::memcpy(B.data() + off, src, cnt * sizeof(int));
        return off;
    };
    sepaihrd::KParams& kp = ctx->kp;
    kp.o_times = put(pb->times, K);
    // observations: entries the likelihood skips (negative or non-finite, ObjectiveFunction.cpp:267) are stored as -1
    auto put_obs = [&](const double* src) {
        std::vector<double> v(src, src + (size_t)pb->n_obs * n);
        for (double& o : v) if (!(o >= 0.0 && std::isfinite(o))) o = -1.0;
        return put(v.data(), v.size());
    };
    kp.o_obs_h = put_obs(pb->obs_hosp);
    kp.o_obs_i = put_obs(pb->obs_icu);
    kp.o_obs_d = put_obs(pb->obs_deaths);
    kp.o_pop = put(pb->population, n);
    {
        // age_fraction = N / N.sum() (ObjectiveFunction.cpp:100-107); inv_N (AgeSEPAIHRDModel.cpp:46-49)
        double total = 0.0;
        for (int i = 0; i < n; ++i) total += pb->population[i];
        std::vector<double> frac(n), inv(n);
        for (int i = 0; i < n; ++i) {
            frac[i] = (total > 0.0) ? pb->population[i] / total : 0.0;
            inv[i] = (pb->population[i] > 1e-9) ? 1.0 / pb->population[i] : 0.0;
        }
        kp.o_agefrac = put(frac.data(), n);
        kp.o_invN = put(inv.data(), n);
    }
    kp.o_M = put(pb->contact_matrix, (size_t)n * n);
    kp.o_bp = put(bp.data(), nseg);
    kp.o_base = put(pb->base_slots, L.count());
    kp.o_init = put(pb->data_initial_state, (size_t)SEPAIHRD_NUM_COMPARTMENTS * n);
    kp.o_lo = put(pb->lower_bound, P);
    kp.o_hi = put(pb->upper_bound, P);
    {
        // FAST-mode logarithm table: c_i = 1 + (i + 1/2)/128; (1/c_i, -log(1/c_i)) with the SAME rounded 1/c_i
        if (B.size() % 2) B.push_back(0.0);   // 16-byte alignment for LDS.128
        std::vector<double> tab(256);
        for (int i = 0; i < 128; ++i) {
            const double invc = 1.0 / (1.0 + (i + 0.5) / 128.0);
            tab[2 * i] = invc;
            tab[2 * i + 1] = (double)(-std::log((long double)invc));
        }
        kp.o_logtab = put(tab.data(), tab.size());
    }
    kp.o_pslot = put_ints(pslot.data(), P);
    kp.o_segb = put_ints(segb.data(), nseg + 1);
    kp.o_segk = put_ints(segk.data(), nseg + 1);
    if (B.size() % 2) B.push_back(0.0);   // multiple of 16 bytes for cp.async.bulk
    kp.blob_bytes = (int)(B.size() * sizeof(double));
    kp.n = n; kp.K = K; kp.n_obs = pb->n_obs; kp.runup_offset = ctx->runup_offset; kp.nb = nb; kp.nk = nk;
    kp.nseg = nseg; kp.P = P; kp.nslots = L.count();
    kp.slot_stride = L.count() | 1;
    kp.seg_stride = (nseg + 1) | 1;
    kp.abs_tol = pb->abs_tol; kp.rel_tol = pb->rel_tol; kp.dt_hint = pb->dt_hint; kp.hmax = ctx->hmax;
    kp.inv_rel = (pb->rel_tol > 0.0) ? 1.0 / pb->rel_tol : 0.0;
    kp.abs_over_rel = (pb->rel_tol > 0.0) ? pb->abs_tol / pb->rel_tol : 0.0;
    kp.grow_max = 9.0 / 10.0 * std::pow(std::pow(5.0, -5.0), -1.0 / 5);
    {   // hmax * coefficient with the device's single rounding (the kernel would compute cur * c_tab[i], cur == hmax)
        const double* tab = sepaihrd::h_tab;
        static_assert(sepaihrd::T_COUNT <= 32, "hc table too small");
        for (int i = 0; i < sepaihrd::T_COUNT; ++i) {
            const bool is_dc = i >= sepaihrd::T_DC1;
            const volatile double step = is_dc ? ctx->hmax * kp.inv_rel : ctx->hmax;   // volatile: no FMA contraction / reassociation on the host
            const volatile double prod = step * tab[i];
            kp.hc[i] = prod;
        }
    }
    // hi(num) - hi(den) equals log2(num/den) within +-0.0862 (units 2^-20) for normal den; margins on top of that
    if (kp.abs_over_rel >= 1e-150) {
        kp.thr_small = (int)std::lround(-11.75 * 1048576.0);    // ratio < 2^-11.66 = 3.09e-4 < 5^-5
        kp.thr_nogrow = (int)std::lround(-0.90 * 1048576.0);    // ratio > 2^-0.987 = 0.5046 > 1/2
        kp.thr_big = (int)std::lround(6.65 * 1048576.0);        // ratio > 2^6.56 = 94.6 > (0.9/0.2)^3 = 91.125
    } else {   // the denominators may be zero/denormal: never trust the coarse classification
        kp.thr_small = INT_MIN; kp.thr_nogrow = INT_MAX; kp.thr_big = INT_MAX;
    }

    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_blob, kp.blob_bytes);
    if (e == cudaSuccess) e = cudaMemcpy(ctx->d_blob, B.data(), kp.blob_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_tile_counter, sepaihrd_ctx::N_TILE_COUNTERS * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking);
    for (int i = 0; i < sepaihrd_ctx::MAX_CHUNKS && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming);
    for (int i = 0; i < sepaihrd_ctx::MAX_CHUNKS && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        std::string msg = std::string("CUDA setup failed: ") + cudaGetErrorString(e);
        sepaihrd_destroy(ctx);
        return fail(SEPAIHRD_ERR_CUDA, msg);
    }
    ctx->num_sms = prop.multiProcessorCount;
    ctx->stream = ctx->own_stream;
    kp.blob = ctx->d_blob;
    if (pb->abs_tol <= 0.0 || pb->rel_tol <= 0.0) ctx->math_mode = SEPAIHRD_MATH_STRICT;   // FAST error norm works in units of rel_tol, positive denominator
    *out_ctx = ctx;
    return SEPAIHRD_OK;
}

}  // namespace

extern "C" {

void sepaihrd_destroy(sepaihrd_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->d_blob) cudaFree(ctx->d_blob);
    if (ctx->d_tile_counter) cudaFree(ctx->d_tile_counter);
    if (ctx->d_params) cudaFree(ctx->d_params);
    if (ctx->d_out) cudaFree(ctx->d_out);
    if (ctx->d_status) cudaFree(ctx->d_status);
    if (ctx->d_steps) cudaFree(ctx->d_steps);
    sepaihrd_internal::order_release(ctx);
    for (void* p : ctx->scratch) if (p) cudaFree(p);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->h_small) cudaFreeHost(ctx->h_small);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    for (int i = 0; i < sepaihrd_ctx::MAX_CHUNKS; ++i) if (ctx->ev_chunk[i]) cudaEventDestroy(ctx->ev_chunk[i]);
    for (int i = 0; i < sepaihrd_ctx::MAX_CHUNKS; ++i) if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]);
    delete ctx;
}

sepaihrd_rc sepaihrd_set_constraint_mode(sepaihrd_ctx* ctx, int32_t mode) {
    if (!ctx || (mode != 0 && mode != 1)) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "bad constraint mode");
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    ctx->constraint_mode = mode;
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_set_math_mode(sepaihrd_ctx* ctx, int32_t mode) {
    if (!ctx || (mode != SEPAIHRD_MATH_FAST && mode != SEPAIHRD_MATH_STRICT && mode != SEPAIHRD_MATH_FAST_GENERAL && mode != SEPAIHRD_MATH_FAST_SPLIT)) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "bad math mode");
    if (mode != SEPAIHRD_MATH_STRICT && (ctx->abs_tol <= 0.0 || ctx->rel_tol <= 0.0)) return fail(SEPAIHRD_ERR_UNSUPPORTED, "FAST math needs abs_tol > 0 and rel_tol > 0");
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    ctx->math_mode = mode;
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_set_stream(sepaihrd_ctx* ctx, void* cuda_stream) {
    if (!ctx) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "null ctx");
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    ctx->stream = (cuda_stream == SEPAIHRD_STREAM_OWN) ? ctx->own_stream : (cudaStream_t)cuda_stream;
    return SEPAIHRD_OK;
}

#ifdef SEPAIHRD_DEBUG_INTERVALS
sepaihrd_rc sepaihrd_debug_set_trace(sepaihrd_ctx* ctx, double* d_trace) { ctx->dbg_trace = d_trace; return SEPAIHRD_OK; }
#endif

sepaihrd_rc sepaihrd_release_scratch(sepaihrd_ctx* ctx) {
    if (!ctx) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    sepaihrd_internal::release_scratch(ctx);
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_alloc_pinned(size_t bytes, void** out) {
    if (!out) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    const cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return fail(SEPAIHRD_ERR_NO_DEVICE, cudaGetErrorString(e));
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? SEPAIHRD_ERR_OUT_OF_MEMORY : SEPAIHRD_ERR_CUDA, cudaGetErrorString(e));
    return SEPAIHRD_OK;
}

void sepaihrd_free_pinned(void* ptr) {
    if (ptr) cudaFreeHost(ptr);
}

sepaihrd_rc sepaihrd_synchronize(sepaihrd_ctx* ctx) {
    if (!ctx) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "null ctx");
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_get_counters(const sepaihrd_ctx* ctx, int64_t* launches, int64_t* sets) {
    if (!ctx) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "null ctx");
    if (launches) *launches = ctx->launches;
    if (sets) *sets = ctx->sets;
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_get_merge_counters(const sepaihrd_ctx* ctx, int64_t* merged_launches, int64_t* merged_requests) {
    if (!ctx) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "null ctx");
    if (merged_launches) *merged_launches = ctx->merged_launches;
    if (merged_requests) *merged_requests = ctx->merged_requests;
    return SEPAIHRD_OK;
}

static sepaihrd_rc eval_device_impl(sepaihrd_ctx* ctx, const double* d_params, int64_t B, int64_t ld, double* d_out_ll, uint32_t* d_out_status,
                                    int32_t* d_out_steps, bool allow_order);

sepaihrd_rc sepaihrd_eval_batch_device(sepaihrd_ctx* ctx, const double* d_params, int64_t B, int64_t ld,
                                       double* d_out_ll, uint32_t* d_out_status, int32_t* d_out_steps) {
    return eval_device_impl(ctx, d_params, B, ld, d_out_ll, d_out_status, d_out_steps, true);
}

static sepaihrd_rc eval_device_impl(sepaihrd_ctx* ctx, const double* d_params, int64_t B, int64_t ld, double* d_out_ll, uint32_t* d_out_status,
                                    int32_t* d_out_steps, bool allow_order) {
    if (!ctx || !d_out_ll || (B > 0 && !d_params)) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (B < 0 || ld < ctx->P) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "Parameter vector size mismatch.");   // ParameterManager.cpp:165-167
    if (B == 0) return SEPAIHRD_OK;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (ctx->obs_mismatch) {
        fill_kernel<<<(unsigned)((B + 255) / 256), 256, 0, ctx->stream>>>(d_out_ll, d_out_status, d_out_steps, B, -DBL_MAX,
                                                                        SEPAIHRD_ST_INVALID_PARAM);
        CUDA_TRY(cudaGetLastError());
        ctx->launches += 1;
        return SEPAIHRD_OK;
    }
    sepaihrd::KParams kp = ctx->kp;
    kp.constraint_mode = ctx->constraint_mode;
    kp.params = d_params; kp.B = B; kp.ld = ld;
    kp.out_ll = d_out_ll; kp.out_status = d_out_status; kp.out_steps = d_out_steps;
    kp.out_traj = nullptr; kp.traj_what = 0; kp.traj_stride = 1; kp.traj_rows = 0; kp.traj_draw_minor = 0;
#ifdef SEPAIHRD_DEBUG_INTERVALS
    kp.out_traj = ctx->dbg_trace;      // diagnostic build: [B][2048][3] (t, step, err) of every attempt
#endif
    kp.init_states = nullptr; kp.init_stride = 0;
    kp.perm = nullptr; kp.out_profile = nullptr;
    if (allow_order) {
        const sepaihrd_rc orc = sepaihrd_internal::order_batch(ctx, d_params, B, ld, &kp.perm);      // no model / small batch: perm stays null
        if (orc != SEPAIHRD_OK) return orc;
    }
    const sepaihrd_rc lrc = launch(ctx, kp, sepaihrd::MODE_LL);
    if (kp.perm) { const sepaihrd_rc mrc = sepaihrd_internal::order_mark_done(ctx); if (lrc == SEPAIHRD_OK && mrc != SEPAIHRD_OK) return mrc; }
    return lrc;
}

}  // extern "C"

namespace {

// the host-pointer evaluation itself; the caller holds no lock
sepaihrd_rc eval_batch_serial(sepaihrd_ctx* ctx, const double* params, int64_t B, int64_t ld, double* out_ll,
                              uint32_t* out_status, int32_t* out_steps) {
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    sepaihrd_rc rc;
    if ((rc = grow(&ctx->d_params, &ctx->cap_params, (size_t)B * ld)) != SEPAIHRD_OK) return rc;
    if ((rc = grow(&ctx->d_out, &ctx->cap_out, (size_t)B)) != SEPAIHRD_OK) return rc;
    if ((rc = grow(&ctx->d_status, &ctx->cap_status, (size_t)B)) != SEPAIHRD_OK) return rc;
    if (out_steps && (rc = grow(&ctx->d_steps, &ctx->cap_steps, (size_t)B * 2)) != SEPAIHRD_OK) return rc;
    if ((rc = sepaihrd_internal::order_autofit_host(ctx, params, B, ld)) != SEPAIHRD_OK) return rc;      // large batches: keep the ordering model current (sepaihrd_order.cu)
    // Chunks on two streams: while the kernel works on a chunk, the copy stream brings the next ones over (H2D moves several
    // times more sets per second than the kernel consumes), so only the FIRST chunk's copy and the last chunk's 12-byte-per-set
    // results are exposed; the chunks grow geometrically so that each copy still finishes under the kernel before it.  Small
    // batches go as one piece.  SEPAIHRD_E2E_SPLIT="d1,d2,..." (chunk ends at B/d1 < B/d2 < ...) overrides the default for
    // experiments.
    static const std::vector<int> split = [] {
        std::vector<int> v;
        if (const char* e = std::getenv("SEPAIHRD_E2E_SPLIT")) {
            for (const char* p = e; *p;) { const long d = std::strtol(p, const_cast<char**>(&p), 10); if (d > 1) v.push_back((int)d); while (*p == ',' || *p == ' ') ++p; }
        }
        if (v.empty()) v = {SEPAIHRD_E2E_DEFAULT_SPLIT};
        if (v.size() > (size_t)sepaihrd_ctx::MAX_CHUNKS - 1) v.resize(sepaihrd_ctx::MAX_CHUNKS - 1);
        std::sort(v.begin(), v.end(), [](int a, int b) { return a > b; });
        return v;
    }();
    int64_t off[sepaihrd_ctx::MAX_CHUNKS + 1];
    off[0] = 0;
    for (int i = 1; i <= sepaihrd_ctx::MAX_CHUNKS; ++i) off[i] = B;
    int n_chunks = 1;
    if (B >= (1 << 16)) {
        for (int d : split) {
            const int64_t end = ((B / d + 255) / 256) * 256;
            if (end > off[n_chunks - 1] && end < B) { off[n_chunks] = end; ++n_chunks; }
        }
        off[n_chunks] = B;
    }
    struct DrainCopies {     // declared before the first copy is queued: an early return below must not leave one in flight
        sepaihrd_ctx* ctx;
        ~DrainCopies() { cudaStreamSynchronize(ctx->copy_stream); }
    } drain_copies{ctx};
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t b0 = off[c], nb = off[c + 1] - off[c];
        CUDA_TRY(cudaMemcpyAsync(ctx->d_params + b0 * ld, params + b0 * ld, sizeof(double) * (size_t)nb * ld, cudaMemcpyHostToDevice, ctx->copy_stream));
        CUDA_TRY(cudaEventRecord(ctx->ev_copy[c], ctx->copy_stream));
    }
    // Chunk c runs on the ctx stream (even c) or on a second compute stream (odd c); every launch has its own tile counter: nothing orders
    // the kernels among themselves, so the first blocks of chunk c + 1 start on the SMs the last warps of chunk c have left.
    // On EVERY way out -- error returns included -- the three streams are drained before the caller gets its buffers back:
    // copies of earlier chunks may still be reading `params` or writing `out_*`.
    struct Restore {
        sepaihrd_ctx* ctx; cudaStream_t stream;
        ~Restore() {
            ctx->stream = stream;
            cudaStreamSynchronize(ctx->stream2); cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(stream);
        }
    } restore{ctx, ctx->stream};
    static const bool two_streams = std::getenv("SEPAIHRD_E2E_ONE_STREAM") == nullptr;
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t b0 = off[c], nb = off[c + 1] - off[c];
        cudaStream_t s = (two_streams && (c & 1)) ? ctx->stream2 : restore.stream;
        ctx->stream = s;
        CUDA_TRY(cudaStreamWaitEvent(s, ctx->ev_copy[c], 0));
        rc = sepaihrd_eval_batch_device(ctx, ctx->d_params + b0 * ld, nb, ld, ctx->d_out + b0, ctx->d_status + b0,
                                        out_steps ? ctx->d_steps + 2 * b0 : nullptr);
        if (rc != SEPAIHRD_OK) return rc;
        // results go back on the copy stream as soon as their kernel is done (under the following kernels)
        CUDA_TRY(cudaEventRecord(ctx->ev_chunk[c], s));
        CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_chunk[c], 0));
        CUDA_TRY(cudaMemcpyAsync(out_ll + b0, ctx->d_out + b0, sizeof(double) * (size_t)nb, cudaMemcpyDeviceToHost, ctx->copy_stream));
        if (out_status) CUDA_TRY(cudaMemcpyAsync(out_status + b0, ctx->d_status + b0, sizeof(unsigned) * (size_t)nb, cudaMemcpyDeviceToHost, ctx->copy_stream));
        if (out_steps) CUDA_TRY(cudaMemcpyAsync(out_steps + 2 * b0, ctx->d_steps + 2 * b0, sizeof(int) * (size_t)nb * 2, cudaMemcpyDeviceToHost, ctx->copy_stream));
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream2));
    ctx->stream = restore.stream;
    CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SEPAIHRD_OK;
}

constexpr int64_t COALESCE_MAX_SETS = 4096;      // larger requests fill the machine on their own

bool page_locked(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// A small request (<= COALESCE_MAX_SETS: a look-ahead window of the one-chain sampler, a line search, a few thousand chains) is
// latency, not throughput: ONE piece on the ctx stream -- copy in, launch, copy out, one synchronisation -- without the copy
// stream, the chunk events and the three stream synchronisations of the large path.  Pageable caller buffers (an
// Eigen::VectorXd, a std::vector) are staged through a page-locked buffer of the ctx: a cudaMemcpyAsync on pageable memory
// is a synchronous, driver-staged copy, three of them per call cost more than the rest of the call's overhead together
// (measured through the C++ host layer: 0.606 ms per calculate() against 0.538 ms from page-locked buffers).
sepaihrd_rc eval_batch_small(sepaihrd_ctx* ctx, const double* params, int64_t B, int64_t ld, double* out_ll,
                             uint32_t* out_status, int32_t* out_steps) {
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int P = ctx->P;
    sepaihrd_rc rc;
    if ((rc = grow(&ctx->d_params, &ctx->cap_params, (size_t)B * ld)) != SEPAIHRD_OK) return rc;
    if ((rc = grow(&ctx->d_out, &ctx->cap_out, (size_t)B)) != SEPAIHRD_OK) return rc;
    if ((rc = grow(&ctx->d_status, &ctx->cap_status, (size_t)B)) != SEPAIHRD_OK) return rc;
    if (out_steps && (rc = grow(&ctx->d_steps, &ctx->cap_steps, (size_t)B * 2)) != SEPAIHRD_OK) return rc;
    const bool stage_in = !page_locked(params);
    const bool stage_out = !(page_locked(out_ll) && (!out_status || page_locked(out_status)) && (!out_steps || page_locked(out_steps)));
    const size_t doubles = (size_t)B * P + (size_t)B /* logL */ + (size_t)(B + 1) / 2 /* status */ + (size_t)B /* steps */;
    if ((stage_in || stage_out) && doubles > ctx->cap_small) {
        if (ctx->h_small) cudaFreeHost(ctx->h_small);
        ctx->h_small = nullptr; ctx->cap_small = 0;
        const size_t want = std::max<size_t>(doubles * 2, 1 << 14);
        CUDA_TRY(cudaMallocHost((void**)&ctx->h_small, want * sizeof(double)));
        ctx->cap_small = want;
    }
    double* h_rows = ctx->h_small;
    double* h_ll = stage_out ? h_rows + (size_t)B * P : out_ll;
    uint32_t* h_st = stage_out ? reinterpret_cast<uint32_t*>(ctx->h_small + (size_t)B * P + B) : out_status;
    int32_t* h_steps = stage_out ? reinterpret_cast<int32_t*>(ctx->h_small + (size_t)B * P + B + (B + 1) / 2) : out_steps;
    cudaStream_t s = ctx->stream;
    struct Drain {           // an early return must not leave a copy from / into the caller's buffers in flight
        cudaStream_t s;
        ~Drain() { cudaStreamSynchronize(s); }
    } drain{s};
    int64_t dev_ld = ld;
    if (stage_in) {          // rows packed to P doubles
        for (int64_t b = 0; b < B; ++b) std::memcpy(h_rows + b * P, params + b * ld, sizeof(double) * (size_t)P);
        dev_ld = P;
        CUDA_TRY(cudaMemcpyAsync(ctx->d_params, h_rows, sizeof(double) * (size_t)B * P, cudaMemcpyHostToDevice, s));
    } else {
        CUDA_TRY(cudaMemcpyAsync(ctx->d_params, params, sizeof(double) * (size_t)B * ld, cudaMemcpyHostToDevice, s));
    }
    rc = sepaihrd_eval_batch_device(ctx, ctx->d_params, B, dev_ld, ctx->d_out, ctx->d_status, out_steps ? ctx->d_steps : nullptr);
    if (rc != SEPAIHRD_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(h_ll, ctx->d_out, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, s));
    if (out_status) CUDA_TRY(cudaMemcpyAsync(h_st, ctx->d_status, sizeof(unsigned) * (size_t)B, cudaMemcpyDeviceToHost, s));
    if (out_steps) CUDA_TRY(cudaMemcpyAsync(h_steps, ctx->d_steps, sizeof(int) * (size_t)B * 2, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (stage_out) {
        std::memcpy(out_ll, h_ll, sizeof(double) * (size_t)B);
        if (out_status) std::memcpy(out_status, h_st, sizeof(uint32_t) * (size_t)B);
        if (out_steps) std::memcpy(out_steps, h_steps, sizeof(int32_t) * 2 * (size_t)B);
    }
    return SEPAIHRD_OK;
}

// Serve the requests of `batch` (>= 2) with ONE launch: rows packed into a page-locked staging buffer (leading dimension P),
// results scattered back.
sepaihrd_rc eval_merged(sepaihrd_ctx* ctx, const std::vector<EvalRequest*>& batch) {
    const int P = ctx->P;
    int64_t total = 0;
    bool want_status = false, want_steps = false;
    for (const EvalRequest* r : batch) { total += r->B; want_status |= r->out_status != nullptr; want_steps |= r->out_steps != nullptr; }
    const size_t doubles = (size_t)total * P + (size_t)total /* logL */ + (size_t)(total + 1) / 2 /* status */ + (size_t)total /* steps */;
    {
        std::lock_guard<std::recursive_mutex> lock(ctx->mu);
        if (doubles > ctx->cap_stage) {
            CUDA_TRY(cudaSetDevice(ctx->device));
            if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
            ctx->h_stage = nullptr; ctx->cap_stage = 0;
            const size_t want = std::max<size_t>(doubles * 2, 1 << 16);
            CUDA_TRY(cudaMallocHost((void**)&ctx->h_stage, want * sizeof(double)));
            ctx->cap_stage = want;
        }
    }
    double* rows = ctx->h_stage;
    double* ll = rows + (size_t)total * P;
    uint32_t* st = reinterpret_cast<uint32_t*>(ll + total);
    int32_t* steps = reinterpret_cast<int32_t*>(ll + total + (total + 1) / 2);
    int64_t at = 0;
    for (const EvalRequest* r : batch)
        for (int64_t b = 0; b < r->B; ++b, ++at) std::memcpy(rows + at * P, r->params + b * r->ld, sizeof(double) * (size_t)P);
    const sepaihrd_rc rc = (total <= COALESCE_MAX_SETS) ? eval_batch_small(ctx, rows, total, P, ll, want_status ? st : nullptr, want_steps ? steps : nullptr)
                                                        : eval_batch_serial(ctx, rows, total, P, ll, want_status ? st : nullptr, want_steps ? steps : nullptr);
    if (rc != SEPAIHRD_OK) return rc;
    at = 0;
    for (EvalRequest* r : batch) {
        std::memcpy(r->out_ll, ll + at, sizeof(double) * (size_t)r->B);
        if (r->out_status) std::memcpy(r->out_status, st + at, sizeof(uint32_t) * (size_t)r->B);
        if (r->out_steps) std::memcpy(r->out_steps, steps + 2 * at, sizeof(int32_t) * 2 * (size_t)r->B);
        at += r->B;
    }
    ctx->merged_launches += 1;
    ctx->merged_requests += (long long)batch.size();
    return SEPAIHRD_OK;
}

}  // namespace

extern "C" {

// B calls of calculate().  Thread-safe; small requests that arrive while another one is being served are MERGED: the caller
// that finds no leader takes everything queued so far (its own request included) to the device as one launch, wakes the owners
// and hands the lead on.  An OpenMP loop over calculate() -- the reference's optimizers, unchanged -- therefore costs one launch
// per round of its threads instead of one launch per call.
sepaihrd_rc sepaihrd_eval_batch(sepaihrd_ctx* ctx, const double* params, int64_t B, int64_t ld, double* out_ll,
                                uint32_t* out_status, int32_t* out_steps) {
    if (!ctx || !out_ll || (B > 0 && !params)) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (B < 0 || ld < ctx->P) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "Parameter vector size mismatch.");
    if (B == 0) return SEPAIHRD_OK;
    if (B > COALESCE_MAX_SETS) return eval_batch_serial(ctx, params, B, ld, out_ll, out_status, out_steps);
    EvalRequest me;
    me.params = params; me.B = B; me.ld = ld; me.out_ll = out_ll; me.out_status = out_status; me.out_steps = out_steps;
    std::unique_lock<std::mutex> ql(ctx->q_mu);
    ctx->queue.push_back(&me);
    if (ctx->leader) me.cv.wait(ql, [&] { return me.done || me.lead; });
    if (!me.done) {
        // lead ONE round: everything queued up to now, this request among it
        ctx->leader = true;
        std::vector<EvalRequest*> batch;
        batch.swap(ctx->queue);
        ql.unlock();
        sepaihrd_rc rc;
        if (batch.size() == 1) rc = eval_batch_small(ctx, params, B, ld, out_ll, out_status, out_steps);    // nobody else
        else rc = eval_merged(ctx, batch);
        const std::string err = (rc != SEPAIHRD_OK) ? g_last_error : std::string();
        ql.lock();
        // hand the lead to the oldest request that arrived meanwhile (it starts its round while the owners below wake up)
        EvalRequest* next = ctx->queue.empty() ? nullptr : ctx->queue.front();
        if (next) next->lead = true; else ctx->leader = false;
        for (EvalRequest* r : batch) { r->rc = rc; r->error = err; r->done = true; }
        if (next) next->cv.notify_one();
        for (EvalRequest* r : batch) if (r != &me) r->cv.notify_one();      // under the lock: a woken owner's request lives on its stack
        ql.unlock();
        return rc;
    }
    ql.unlock();
    if (me.rc != SEPAIHRD_OK) g_last_error = me.error;
    return me.rc;
}

static sepaihrd_rc simulate_device_impl(sepaihrd_ctx* ctx, const double* d_params, int64_t B, int64_t ld,
                                        const double* d_init, int64_t init_stride,
                                        int32_t what, int32_t stride, double* d_out, uint32_t* d_out_status, bool draw_minor = false) {
    if (!ctx || !d_out || (B > 0 && !d_params)) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (B < 0 || ld < ctx->P) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "Parameter vector size mismatch.");
    if (what != SEPAIHRD_TRAJ_FULL && what != SEPAIHRD_TRAJ_OBSERVED) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "bad trajectory selector");
    if (stride < 1) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "stride must be >= 1");
    if (B == 0) return SEPAIHRD_OK;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int rows = (ctx->K + stride - 1) / stride;
    // padded age classes (sepaihrd_create): the caller's states are widened on the way in, the trajectories narrowed on the way
    // out; the draw-minor layout is internal (posterior-predictive pass) and stays padded
    const bool repack = ctx->n != ctx->n_user && !draw_minor;
    const int C = (what == SEPAIHRD_TRAJ_FULL) ? SEPAIHRD_NUM_COMPARTMENTS : 3;
    double* d_kernel_out = d_out;
    if (repack) {
        if (d_init) {
            const long long n_vec = (init_stride == 0) ? 1 : B;
            const long long wide = (long long)SEPAIHRD_NUM_COMPARTMENTS * ctx->n;
            double* d_wide = (double*)sepaihrd_internal::scratch(ctx, 14, sizeof(double) * (size_t)(n_vec * wide));
            if (!d_wide) return fail(SEPAIHRD_ERR_OUT_OF_MEMORY, "no room for the padded initial states");
            repack_ages_kernel<<<(unsigned)std::min<long long>((n_vec * wide + 255) / 256, 65535), 256, 0, ctx->stream>>>(
                d_init, init_stride, d_wide, wide, n_vec, SEPAIHRD_NUM_COMPARTMENTS, ctx->n_user, ctx->n);
            CUDA_TRY(cudaGetLastError());
            d_init = d_wide;
            if (init_stride != 0) init_stride = wide;
        }
        d_kernel_out = (double*)sepaihrd_internal::scratch(ctx, 15, sizeof(double) * (size_t)B * rows * C * ctx->n);
        if (!d_kernel_out) return fail(SEPAIHRD_ERR_OUT_OF_MEMORY, "no room for the padded trajectories: split the batch");
    }
    sepaihrd::KParams kp = ctx->kp;
    kp.constraint_mode = ctx->constraint_mode;
    kp.params = d_params; kp.B = B; kp.ld = ld;
    kp.out_ll = nullptr; kp.out_status = d_out_status; kp.out_steps = nullptr;
    kp.out_traj = d_kernel_out; kp.traj_what = what; kp.traj_stride = stride; kp.traj_rows = rows;
    kp.traj_draw_minor = draw_minor ? 1 : 0;
    kp.init_states = d_init; kp.init_stride = init_stride;
    kp.perm = nullptr; kp.out_profile = nullptr;
    const sepaihrd_rc rc = launch(ctx, kp, sepaihrd::MODE_TRAJ);
    if (rc != SEPAIHRD_OK || !repack) return rc;
    const long long n_vec = (long long)B * rows, narrow = (long long)C * ctx->n_user;
    repack_ages_kernel<<<(unsigned)std::min<long long>((n_vec * narrow + 255) / 256, 148 * 32), 256, 0, ctx->stream>>>(
        d_kernel_out, (long long)C * ctx->n, d_out, narrow, n_vec, C, ctx->n, ctx->n_user);
    CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_simulate_batch_device(sepaihrd_ctx* ctx, const double* d_params, int64_t B, int64_t ld,
                                           int32_t what, int32_t stride, double* d_out, uint32_t* d_out_status) {
    return simulate_device_impl(ctx, d_params, B, ld, nullptr, 0, what, stride, d_out, d_out_status);
}

sepaihrd_rc sepaihrd_simulate_from_state_device(sepaihrd_ctx* ctx, const double* d_params, int64_t B, int64_t ld,
                                                const double* d_initial_states, int64_t state_stride, int32_t what,
                                                int32_t stride, double* d_out, uint32_t* d_out_status) {
    if (!d_initial_states) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "null initial state");
    return simulate_device_impl(ctx, d_params, B, ld, d_initial_states, state_stride, what, stride, d_out, d_out_status);
}

sepaihrd_rc sepaihrd_simulate_batch(sepaihrd_ctx* ctx, const double* params, int64_t B, int64_t ld, int32_t what,
                                    int32_t stride, double* out, uint32_t* out_status) {
    if (!ctx || !out || (B > 0 && !params)) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (B < 0 || ld < ctx->P) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "Parameter vector size mismatch.");
    if (stride < 1) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "stride must be >= 1");
    if (B == 0) return SEPAIHRD_OK;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int W = (what == SEPAIHRD_TRAJ_FULL) ? SEPAIHRD_NUM_COMPARTMENTS * ctx->n_user : 3 * ctx->n_user;
    const size_t rows = (size_t)(ctx->K + stride - 1) / stride;
    const size_t total = (size_t)B * rows * W;
    sepaihrd_rc rc;
    if ((rc = grow(&ctx->d_params, &ctx->cap_params, (size_t)B * ld)) != SEPAIHRD_OK) return rc;
    if ((rc = grow(&ctx->d_out, &ctx->cap_out, total)) != SEPAIHRD_OK) return rc;
    if ((rc = grow(&ctx->d_status, &ctx->cap_status, (size_t)B)) != SEPAIHRD_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(ctx->d_params, params, sizeof(double) * (size_t)B * ld, cudaMemcpyHostToDevice, ctx->stream));
    rc = sepaihrd_simulate_batch_device(ctx, ctx->d_params, B, ld, what, stride, ctx->d_out, ctx->d_status);
    if (rc != SEPAIHRD_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, ctx->d_out, sizeof(double) * total, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_status) CUDA_TRY(cudaMemcpyAsync(out_status, ctx->d_status, sizeof(unsigned) * (size_t)B, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_simulate_from_state(sepaihrd_ctx* ctx, const double* params, int64_t B, int64_t ld,
                                         const double* initial_states, int64_t state_stride, int32_t what, int32_t stride,
                                         double* out, uint32_t* out_status) {
    if (!ctx || !out || !initial_states || (B > 0 && !params)) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (B < 0 || ld < ctx->P) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "Parameter vector size mismatch.");
    if (stride < 1) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "stride must be >= 1");
    if (state_stride != 0 && state_stride < (int64_t)SEPAIHRD_NUM_COMPARTMENTS * ctx->n_user)
        return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "Initial state size does not match model state size.");   // Simulator.cpp:64-69
    if (B == 0) return SEPAIHRD_OK;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int W = (what == SEPAIHRD_TRAJ_FULL) ? SEPAIHRD_NUM_COMPARTMENTS * ctx->n_user : 3 * ctx->n_user;
    const size_t rows = (size_t)(ctx->K + stride - 1) / stride;
    const size_t total = (size_t)B * rows * W;
    const size_t n_init = (state_stride == 0) ? (size_t)SEPAIHRD_NUM_COMPARTMENTS * ctx->n_user : (size_t)B * state_stride;
    sepaihrd_rc rc;
    if ((rc = grow(&ctx->d_params, &ctx->cap_params, (size_t)B * ld + n_init)) != SEPAIHRD_OK) return rc;
    if ((rc = grow(&ctx->d_out, &ctx->cap_out, total)) != SEPAIHRD_OK) return rc;
    if ((rc = grow(&ctx->d_status, &ctx->cap_status, (size_t)B)) != SEPAIHRD_OK) return rc;
    double* d_init = ctx->d_params + (size_t)B * ld;
    CUDA_TRY(cudaMemcpyAsync(ctx->d_params, params, sizeof(double) * (size_t)B * ld, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_init, initial_states, sizeof(double) * n_init, cudaMemcpyHostToDevice, ctx->stream));
    rc = simulate_device_impl(ctx, ctx->d_params, B, ld, d_init, state_stride, what, stride, ctx->d_out, ctx->d_status);
    if (rc != SEPAIHRD_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, ctx->d_out, sizeof(double) * total, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_status) CUDA_TRY(cudaMemcpyAsync(out_status, ctx->d_status, sizeof(unsigned) * (size_t)B, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_measure_fp64_peak(int32_t device, double* out_dfma_per_second) {
    if (!out_dfma_per_second) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) return fail(SEPAIHRD_ERR_NO_DEVICE, "no CUDA device");
    if (device < 0) device = 0;
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    double* d_out = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d_out, sizeof(double)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CUDA_TRY(cudaEventRecord(e0));
        dfma_peak_kernel<<<blocks, threads>>>(d_out, 1.0000001, 0.9999999);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double ops = (double)blocks * threads * PEAK_CHAINS * PEAK_ITERS;
        if (rep > 0) best = std::max(best, ops / (ms * 1e-3));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    *out_dfma_per_second = best;
    return SEPAIHRD_OK;
}

}  // extern "C"

// ---- accessors for the other translation units of the library (sepaihrd_internal.h) -----------------------------
namespace sepaihrd_internal {
Dims dims(const sepaihrd_ctx* ctx) { return Dims{ctx->n, ctx->K, ctx->runup_offset, ctx->n_nonneg, ctx->P, ctx->device, ctx->n_user}; }
cudaStream_t stream(const sepaihrd_ctx* ctx) { return ctx->stream; }
cudaStream_t copy_stream(const sepaihrd_ctx* ctx) { return ctx->copy_stream; }
cudaEvent_t* copy_events(sepaihrd_ctx* ctx) { return ctx->ev_copy; }
cudaEvent_t* chunk_events(sepaihrd_ctx* ctx) { return ctx->ev_chunk; }
sepaihrd_rc fail_with(sepaihrd_rc rc, const char* msg) { return fail(rc, msg); }
const double* lower_bounds(const sepaihrd_ctx* ctx) { return ctx->blob.data() + ctx->kp.o_lo; }
const double* upper_bounds(const sepaihrd_ctx* ctx) { return ctx->blob.data() + ctx->kp.o_hi; }
void count_launches(sepaihrd_ctx* ctx, int n) { ctx->launches += n; }
int constraint_mode(const sepaihrd_ctx* ctx) { return ctx->constraint_mode; }
void** order_slot(sepaihrd_ctx* ctx) { return &ctx->order_model; }
int order_mode(const sepaihrd_ctx* ctx) { return ctx->order_mode; }
void set_order_mode(sepaihrd_ctx* ctx, int mode) { ctx->order_mode = mode; }
int num_sms(const sepaihrd_ctx* ctx) { return ctx->num_sms; }
bool order_applicable(const sepaihrd_ctx* ctx) { return ctx->n == 4 && ctx->math_mode != SEPAIHRD_MATH_STRICT && !ctx->obs_mismatch; }
sepaihrd_rc eval_batch_device_unordered(sepaihrd_ctx* ctx, const double* d_params, long long B, long long ld, double* d_ll, unsigned* d_status, int* d_steps) {
    return eval_device_impl(ctx, d_params, B, ld, d_ll, d_status, d_steps, false);
}
// the pilot of the ordering pass: one launch of the PROFILE instantiation (FAST arithmetic, 4 lanes per set)
sepaihrd_rc eval_profile(sepaihrd_ctx* ctx, const double* d_params, long long B, long long ld, double* d_ll, unsigned* d_status, int* d_profile) {
    if (ctx->n != 4 || ctx->math_mode == SEPAIHRD_MATH_STRICT || ctx->obs_mismatch) return fail(SEPAIHRD_ERR_UNSUPPORTED, "no attempt profile for this configuration");
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    sepaihrd::KParams kp = ctx->kp;
    kp.constraint_mode = ctx->constraint_mode;
    kp.params = d_params; kp.B = B; kp.ld = ld;
    kp.out_ll = d_ll; kp.out_status = d_status; kp.out_steps = nullptr;
    kp.out_traj = nullptr; kp.traj_what = 0; kp.traj_stride = 1; kp.traj_rows = 0; kp.traj_draw_minor = 0;
    kp.init_states = nullptr; kp.init_stride = 0;
    kp.perm = nullptr; kp.out_profile = d_profile;
    if (ctx->bp_on_grid && ctx->math_mode != SEPAIHRD_MATH_FAST_GENERAL) return launch_t<4, false, sepaihrd::MODE_LL, 128, 2, 6, true, true>(ctx, kp);
    return launch_t<4, false, sepaihrd::MODE_LL, 128, 2, 6, false, true>(ctx, kp);
}
std::unique_lock<std::recursive_mutex> lock(sepaihrd_ctx* ctx) { return std::unique_lock<std::recursive_mutex>(ctx->mu); }
void* scratch(sepaihrd_ctx* ctx, int slot, size_t bytes) {
    if (slot < 0 || slot >= sepaihrd_ctx::N_SCRATCH) return nullptr;
    if (ctx->scratch_bytes[slot] >= bytes && ctx->scratch[slot]) return ctx->scratch[slot];
    if (ctx->scratch[slot]) {
        cudaDeviceSynchronize();                    // the ctx stream may have been re-pointed since the buffer was last used: drain the whole device
        cudaFree(ctx->scratch[slot]);
    }
    ctx->scratch[slot] = nullptr; ctx->scratch_bytes[slot] = 0;
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    ctx->scratch[slot] = p; ctx->scratch_bytes[slot] = bytes;
    return p;
}
void release_scratch(sepaihrd_ctx* ctx) {
    cudaDeviceSynchronize();
    for (int i = 0; i < sepaihrd_ctx::N_SCRATCH; ++i) {
        if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
        ctx->scratch[i] = nullptr; ctx->scratch_bytes[i] = 0;
    }
}
sepaihrd_rc simulate_ppc_series(sepaihrd_ctx* ctx, const double* d_params, long long b0, long long nb, long long B_total, long long ld,
                                const double* d_init, double* d_series, unsigned* d_status) {
    if (!ctx || !d_series || !d_params || !d_init) return fail(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    sepaihrd::KParams kp = ctx->kp;
    kp.constraint_mode = ctx->constraint_mode;
    kp.params = d_params + b0 * ld; kp.B = nb; kp.ld = ld;
    kp.out_ll = nullptr; kp.out_status = d_status + b0; kp.out_steps = nullptr;
    kp.out_traj = d_series; kp.traj_what = sepaihrd::TRAJ_PPC_SERIES; kp.traj_stride = 1; kp.traj_rows = ctx->n_nonneg;
    kp.traj_draw_minor = 1;
    kp.ppc_b0 = b0; kp.ppc_B = B_total;
    kp.init_states = d_init; kp.init_stride = 0;
    kp.perm = nullptr; kp.out_profile = nullptr;
    return launch(ctx, kp, sepaihrd::MODE_PPC);
}
}  // namespace sepaihrd_internal
