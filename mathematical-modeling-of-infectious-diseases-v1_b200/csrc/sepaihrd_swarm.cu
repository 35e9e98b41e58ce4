// sepaihrd_swarm.cu -- a particle swarm that LIVES ON THE DEVICE: positions, velocities and personal bests never leave
// HBM between iterations; per iteration the host sends one 32-bit seed per particle and the global best position, and
// reads back one (value, index, position) triple.
//
// Replaces, for swarms whose objective is the device evaluator, the per-iteration host work of
// ParticleSwarmOptimization (reference src/model/optimizers/ParticleSwarmOptimizer.cpp):
//   initializeSwarm       :249-328   uniform-in-bounds positions, velocities in [-vmax, vmax], vmax = 0.2 (ub - lb)
//   updateParticles       :330-425   per-particle std::mt19937 seeded from the master generator (:365-371)
//   standardPSOUpdate     :576-618   v = w v + c1 r1 (pbest - x) + c2 r2 (gbest - x), clamp, reflect at the bounds
//   personal / global best :301-303, :417-421, :149-156  (first maximum wins)
//
// Bit-compatibility with the host implementation (host/optimizers.cpp, itself tested against the oracle): the device
// runs the same generator -- MT19937 with libstdc++'s seeding, tempering and generate_canonical<double, 53> (two 32-bit
// draws per double, low word first) -- and the same unfused FP64 operations in the same order, so a device-resident
// swarm visits EXACTLY the positions of the host swarm (tests/test_gpu_host.py).
//
// Mapping: one thread per particle; the generator state (624 words) is a per-thread local-memory array (lane-interleaved,
// hence coalesced, L1/L2-resident); the row-major [particle][P] arrays are read and written once per iteration (HBM traffic
// 7 x 8 P bytes per particle, against ~1.8 MFLOP of evaluation: irrelevant to the iteration time).
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <vector>

#include "sepaihrd_internal.h"

struct sepaihrd_swarm {
    sepaihrd_ctx* ctx = nullptr;
    int P = 0;
    long long swarm_size = 0, offset = 0, local = 0;
    double *d_pos = nullptr, *d_vel = nullptr, *d_pbest = nullptr, *d_pbest_val = nullptr, *d_fit = nullptr;
    double *d_lb = nullptr, *d_ub = nullptr, *d_gbest = nullptr, *d_init = nullptr;
    unsigned* d_seeds = nullptr;
    unsigned* d_status = nullptr;
    double* d_block_val = nullptr;      // per-block (value, index) candidates of the arg-max
    long long* d_block_idx = nullptr;
    double* d_best = nullptr;           // [2 + P]: value, index (as double), position
    double* h_best = nullptr;           // pinned mirror
    unsigned* h_seeds = nullptr;        // pinned staging
    double* h_gbest = nullptr;
    bool evaluated_once = false;
    int blocks_tell = 0;
    // asynchronous form: seeds of every iteration resident, global best and its trace on the device
    unsigned* d_seed_sets = nullptr;    // [n_seed_sets][swarm_size], a separate allocation (sepaihrd_swarm_upload_seeds)
    long long n_seed_sets = 0;
    double* d_gbest_val = nullptr;      // [1]
    double* d_trace = nullptr;          // [SEPAIHRD_SWARM_TRACE_CAPACITY]
    char* d_arena = nullptr;            // the one device allocation all d_* pointers point into
    char* h_arena = nullptr;            // the one pinned allocation all h_* pointers point into
};

namespace {

using sepaihrd_internal::fail_with;

#define SW_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) return fail_with(SEPAIHRD_ERR_CUDA, cudaGetErrorString(e__));     \
    } while (0)

constexpr int MT_N = 624, MT_M = 397;
constexpr int RNG_THREADS = 128;
// Block size of the generator kernels: they are latency-bound (a 624-step seeding recurrence per particle), so a shard that
// does not fill the GPU with 128-thread blocks is spread over more SMs with smaller ones (8,192 particles: 256 blocks of one
// warp instead of 64 blocks of four -- every SM gets work and a block's generator states fit its L1).
inline int rng_block(long long local) { return local <= 148LL * 32 * 4 ? 32 : (local <= 148LL * 64 * 4 ? 64 : RNG_THREADS); }

// std::mt19937, one generator per thread.  The 624-word state is a per-thread array in LOCAL memory: the hardware interleaves
// local memory by lane, so word i of the 32 generators of a warp is one coalesced 128-byte line, served from L1/L2.  (v10
// kept the states in shared memory -- 156 KB for 64 generators -- which left 2 warps per SM to hide the latencies of the
// seeding recurrence and of the row-major parameter traffic: 0.44 ms per update of 65 536 particles.)
// The state is refilled lazily in blocks of 16 words, in increasing order: a step of 4 P = 248 draws twists 256 words, not 624.
struct Mt19937 {
    unsigned st[MT_N];
    int idx, filled;
    __device__ void seed(unsigned s) {
        unsigned x = s;
        st[0] = x;
#pragma unroll 4
        for (int i = 1; i < MT_N; ++i) {
            x = 1812433253u * (x ^ (x >> 30)) + (unsigned)i;
            st[i] = x;
        }
        idx = MT_N; filled = MT_N;
    }
    // twist words [from, from + 16): word k needs the OLD k + 1 and the word k + 397 (mod 624), which for k >= 227 is the
    // already twisted k - 227 -- exactly the order of the reference algorithm
    __device__ void refill16(int from) {
#pragma unroll 4
        for (int k = from; k < from + 16; ++k) {
            const int k1 = (k + 1 == MT_N) ? 0 : k + 1;
            const int km = (k + MT_M >= MT_N) ? k + MT_M - MT_N : k + MT_M;
            const unsigned y = (st[k] & 0x80000000u) | (st[k1] & 0x7fffffffu);
            st[k] = st[km] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
    }
    __device__ unsigned next() {
        if (idx >= MT_N) { idx = 0; filled = 0; }
        if (idx >= filled) { refill16(filled); filled += 16; }      // 624 = 39 x 16
        unsigned y = st[idx++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    // std::uniform_real_distribution<double>(0, 1): generate_canonical<double, 53> = (lo + hi 2^32) / 2^64 with ONE rounding in
    // the sum, and the (never reached in practice) 1.0 result mapped to the largest double below 1
    __device__ double uniform01() {
        const unsigned lo = next(), hi = next();
        const double sum = __dadd_rn((double)lo, __dmul_rn((double)hi, 4294967296.0));
        const double r = __dmul_rn(sum, 1.0 / 18446744073709551616.0);      // x 2^-64: exact, like the host's sum / 2^64
        return (r >= 1.0) ? 0.99999999999999988898 : r;
    }
};

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return (v < lo) ? lo : (hi < v) ? hi : v; }   // std::clamp

// initializeSwarm: positions (unless this is global particle 0 with a caller-supplied start), then velocities
__global__ void __launch_bounds__(RNG_THREADS) swarm_init_kernel(long long local, long long offset, int P, const unsigned* __restrict__ seeds,
                                                                 const double* __restrict__ lb, const double* __restrict__ ub,
                                                                 const double* __restrict__ init, double* __restrict__ pos,
                                                                 double* __restrict__ vel) {
    const long long li = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (li >= local) return;
    const long long gi = offset + li;
    Mt19937 rng;
    rng.seed(seeds[gi]);
    double* p = pos + li * P;
    double* v = vel + li * P;
    if (gi == 0 && init != nullptr) {
        for (int k = 0; k < P; ++k) p[k] = clampd(init[k], lb[k], ub[k]);
    } else {
        for (int k = 0; k < P; ++k) p[k] = __dadd_rn(lb[k], __dmul_rn(rng.uniform01(), __dsub_rn(ub[k], lb[k])));
    }
    for (int k = 0; k < P; ++k) {
        const double vmax = __dmul_rn(0.2, __dsub_rn(ub[k], lb[k]));
        v[k] = __dadd_rn(-vmax, __dmul_rn(__dmul_rn(2.0, vmax), rng.uniform01()));
    }
}

// standardPSOUpdate for every local particle
__global__ void __launch_bounds__(RNG_THREADS) swarm_step_kernel(long long local, long long offset, int P, const unsigned* __restrict__ seeds,
                                                                 const double* __restrict__ lb, const double* __restrict__ ub,
                                                                 const double* __restrict__ gbest, const double* __restrict__ pbest,
                                                                 double omega, double c1, double c2, double* __restrict__ pos,
                                                                 double* __restrict__ vel) {
    const long long li = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (li >= local) return;
    Mt19937 rng;
    rng.seed(seeds[offset + li]);
    double* p = pos + li * P;
    double* v = vel + li * P;
    const double* pb = pbest + li * P;
    for (int k = 0; k < P; ++k) {
        const double r1 = rng.uniform01(), r2 = rng.uniform01();     // r1_k, r2_k interleaved per dimension
        const double lo = lb[k], hi = ub[k], x = p[k];
        const double cognitive = __dmul_rn(c1, __dmul_rn(r1, __dsub_rn(pb[k], x)));
        const double social = __dmul_rn(c2, __dmul_rn(r2, __dsub_rn(gbest[k], x)));
        double vk = __dadd_rn(__dadd_rn(__dmul_rn(omega, v[k]), cognitive), social);
        const double vmax = __dmul_rn(0.2, __dsub_rn(hi, lo));
        vk = clampd(vk, -vmax, vmax);
        double pk = __dadd_rn(x, vk);
        if (pk < lo) { pk = __dadd_rn(lo, fabs(__dsub_rn(pk, lo))); vk = __dmul_rn(vk, -0.5); }
        else if (pk > hi) { pk = __dsub_rn(hi, fabs(__dsub_rn(pk, hi))); vk = __dmul_rn(vk, -0.5); }
        p[k] = clampd(pk, lo, hi);
        v[k] = vk;
    }
}

// personal bests + per-block arg-max of the personal-best values (ties: the lower index, i.e. "first maximum wins")
constexpr int TELL_THREADS = 256;
__device__ __forceinline__ void take_better(double& v, long long& i, double ov, long long oi) {
    if (oi >= 0 && (i < 0 || ov > v || (ov == v && oi < i))) { v = ov; i = oi; }
}
__global__ void __launch_bounds__(TELL_THREADS) swarm_tell_kernel(long long local, int P, int first, const double* __restrict__ fit,
                                                                  const double* __restrict__ pos, double* __restrict__ pbest,
                                                                  double* __restrict__ pbest_val, double* __restrict__ block_val,
                                                                  long long* __restrict__ block_idx) {
    __shared__ double s_val[TELL_THREADS / 32];
    __shared__ long long s_idx[TELL_THREADS / 32];
    const long long li = blockIdx.x * (long long)TELL_THREADS + threadIdx.x;
    double v = 0.0;
    long long i = -1;
    if (li < local) {
        const double f = fit[li];
        double pv = pbest_val[li];
        if (first || f > pv) {
            pv = f;
            pbest_val[li] = f;
            for (int k = 0; k < P; ++k) pbest[li * P + k] = pos[li * P + k];
        }
        // host rule: `if (pbest_val > best)` starting from -inf -- a NaN or -inf value is never selected
        if (pv > -INFINITY) { v = pv; i = li; }
    }
    for (int off = 16; off >= 1; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, v, off);
        const long long oi = __shfl_down_sync(0xffffffffu, i, off);
        take_better(v, i, ov, oi);
    }
    if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = v; s_idx[threadIdx.x >> 5] = i; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < TELL_THREADS / 32; ++w) take_better(v, i, s_val[w], s_idx[w]);
        block_val[blockIdx.x] = v;
        block_idx[blockIdx.x] = i;
    }
}
// best = [value, particle index + index_offset (-1: none), position[P]]: index_offset 0 gives the shard-local index of the
// synchronous form, the shard's particle_offset the GLOBAL index of the record the ranks exchange
__global__ void __launch_bounds__(TELL_THREADS) swarm_best_kernel(int blocks, int P, const double* __restrict__ block_val,
                                                                  const long long* __restrict__ block_idx, const double* __restrict__ pbest,
                                                                  double* __restrict__ best, long long index_offset) {
    __shared__ double s_val[TELL_THREADS];
    __shared__ long long s_idx[TELL_THREADS];
    double v = 0.0;
    long long i = -1;
    for (int b = threadIdx.x; b < blocks; b += TELL_THREADS) take_better(v, i, block_val[b], block_idx[b]);
    s_val[threadIdx.x] = v; s_idx[threadIdx.x] = i;
    __syncthreads();
    for (int off = TELL_THREADS / 2; off >= 1; off >>= 1) {
        if (threadIdx.x < off) {
            double a = s_val[threadIdx.x]; long long ai = s_idx[threadIdx.x];
            take_better(a, ai, s_val[threadIdx.x + off], s_idx[threadIdx.x + off]);
            s_val[threadIdx.x] = a; s_idx[threadIdx.x] = ai;
        }
        __syncthreads();
    }
    const long long bi = s_idx[0];
    if (threadIdx.x == 0) { best[0] = (bi >= 0) ? s_val[0] : -INFINITY; best[1] = (bi >= 0) ? (double)(bi + index_offset) : -1.0; }
    if (bi >= 0)
        for (int k = threadIdx.x; k < P; k += TELL_THREADS) best[2 + k] = pbest[bi * P + k];
}

// Global best from one record per rank: the arg-max over the records (a NaN value never wins; ties go to the lowest particle
// index, i.e. what a serial scan of the whole swarm finds first) replaces the resident global best when STRICTLY better
// (ParticleSwarmOptimizer.cpp:149-156 / setGlobalBest of the host swarm); trace[slot] = the global best value after that.
__global__ void swarm_adopt_kernel(const double* __restrict__ records, int n_records, long long stride, int P, double* __restrict__ gbest,
                                   double* __restrict__ gbest_val, double* __restrict__ trace, int slot) {
    __shared__ int s_owner;
    if (threadIdx.x == 0) {
        int owner = -1;
        double best = -INFINITY, best_idx = 0.0;
        for (int r = 0; r < n_records; ++r) {
            const double v = records[r * stride], idx = records[r * stride + 1];
            if (!(idx >= 0.0) || !(v == v)) continue;
            if (owner < 0 || v > best || (v == best && idx < best_idx)) { owner = r; best = v; best_idx = idx; }
        }
        if (owner >= 0 && !(best > gbest_val[0])) owner = -1;
        if (owner >= 0) gbest_val[0] = best;
        if (trace) trace[slot] = gbest_val[0];
        s_owner = owner;
    }
    __syncthreads();
    const int owner = s_owner;
    if (owner >= 0)
        for (int k = threadIdx.x; k < P; k += blockDim.x) gbest[k] = records[owner * stride + 2 + k];
}
__global__ void swarm_reset_best_kernel(double* gbest_val) { gbest_val[0] = -INFINITY; }

}  // namespace

extern "C" {

sepaihrd_rc sepaihrd_swarm_create(sepaihrd_ctx* ctx, int64_t swarm_size, int64_t particle_offset, int64_t local_count,
                                  sepaihrd_swarm** out) {
    if (!ctx || !out) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (swarm_size <= 0 || particle_offset < 0 || local_count < 0 || particle_offset + local_count > swarm_size)
        return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "particle_offset/local_count outside the swarm");
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(ctx);
    const double* lo = sepaihrd_internal::lower_bounds(ctx);
    const double* hi = sepaihrd_internal::upper_bounds(ctx);
    for (int k = 0; k < d.P; ++k)
        if (!(std::isfinite(lo[k]) && std::isfinite(hi[k]))) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "a particle swarm needs finite bounds for every parameter");
    SW_TRY(cudaSetDevice(d.device));
    auto* s = new sepaihrd_swarm;
    s->ctx = ctx; s->P = d.P; s->swarm_size = swarm_size; s->offset = particle_offset; s->local = local_count;
    const size_t tot = (size_t)local_count * d.P;
    s->blocks_tell = (int)((local_count + TELL_THREADS - 1) / TELL_THREADS);
    // ONE device allocation and ONE pinned allocation, carved into the arrays (allocation and release cost far more than the
    // arithmetic of an iteration; 256-byte aligned pieces)
    cudaError_t e = cudaSuccess;
    size_t dev_bytes = 0, host_bytes = 0;
    auto reserve = [](size_t& total, size_t bytes) { const size_t at = total; total += (bytes + 255) & ~(size_t)255; return at; };
    const size_t o_pos = reserve(dev_bytes, sizeof(double) * tot), o_vel = reserve(dev_bytes, sizeof(double) * tot),
                 o_pbest = reserve(dev_bytes, sizeof(double) * tot), o_pval = reserve(dev_bytes, sizeof(double) * (size_t)local_count),
                 o_fit = reserve(dev_bytes, sizeof(double) * (size_t)local_count), o_status = reserve(dev_bytes, sizeof(unsigned) * (size_t)local_count),
                 o_lb = reserve(dev_bytes, sizeof(double) * d.P), o_ub = reserve(dev_bytes, sizeof(double) * d.P),
                 o_gbest = reserve(dev_bytes, sizeof(double) * d.P), o_init = reserve(dev_bytes, sizeof(double) * d.P),
                 o_seeds = reserve(dev_bytes, sizeof(unsigned) * (size_t)swarm_size),
                 o_bval = reserve(dev_bytes, sizeof(double) * (size_t)(s->blocks_tell + 1)),
                 o_bidx = reserve(dev_bytes, sizeof(long long) * (size_t)(s->blocks_tell + 1)),
                 o_best = reserve(dev_bytes, sizeof(double) * ((size_t)d.P + 2)),
                 o_gval = reserve(dev_bytes, sizeof(double)), o_trace = reserve(dev_bytes, sizeof(double) * SEPAIHRD_SWARM_TRACE_CAPACITY);
    const size_t h_best = reserve(host_bytes, sizeof(double) * ((size_t)d.P + 2)), h_seeds = reserve(host_bytes, sizeof(unsigned) * (size_t)swarm_size),
                 h_gbest = reserve(host_bytes, sizeof(double) * (size_t)d.P);
    e = cudaMalloc((void**)&s->d_arena, dev_bytes);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&s->h_arena, host_bytes);
    if (e == cudaSuccess) {
        char* D = s->d_arena; char* H = s->h_arena;
        s->d_pos = (double*)(D + o_pos); s->d_vel = (double*)(D + o_vel); s->d_pbest = (double*)(D + o_pbest);
        s->d_pbest_val = (double*)(D + o_pval); s->d_fit = (double*)(D + o_fit); s->d_status = (unsigned*)(D + o_status);
        s->d_lb = (double*)(D + o_lb); s->d_ub = (double*)(D + o_ub); s->d_gbest = (double*)(D + o_gbest); s->d_init = (double*)(D + o_init);
        s->d_seeds = (unsigned*)(D + o_seeds); s->d_block_val = (double*)(D + o_bval); s->d_block_idx = (long long*)(D + o_bidx);
        s->d_best = (double*)(D + o_best); s->d_gbest_val = (double*)(D + o_gval); s->d_trace = (double*)(D + o_trace);
        s->h_best = (double*)(H + h_best); s->h_seeds = (unsigned*)(H + h_seeds); s->h_gbest = (double*)(H + h_gbest);
    }
    if (e == cudaSuccess) e = cudaMemcpy(s->d_lb, lo, sizeof(double) * d.P, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(s->d_ub, hi, sizeof(double) * d.P, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        sepaihrd_swarm_destroy(s);
        return fail_with(e == cudaErrorMemoryAllocation ? SEPAIHRD_ERR_OUT_OF_MEMORY : SEPAIHRD_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = s;
    return SEPAIHRD_OK;
}

void sepaihrd_swarm_destroy(sepaihrd_swarm* s) {
    if (!s) return;
    if (s->d_arena) { cudaDeviceSynchronize(); cudaFree(s->d_arena); }
    if (s->d_seed_sets) cudaFree(s->d_seed_sets);
    if (s->h_arena) cudaFreeHost(s->h_arena);
    delete s;
}

static sepaihrd_rc upload_seeds(sepaihrd_swarm* s, const uint32_t* seeds, cudaStream_t st) {
    for (long long i = 0; i < s->swarm_size; ++i) s->h_seeds[i] = seeds[i];
    SW_TRY(cudaMemcpyAsync(s->d_seeds, s->h_seeds, sizeof(unsigned) * (size_t)s->swarm_size, cudaMemcpyHostToDevice, st));
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_swarm_init(sepaihrd_swarm* s, const uint32_t* seeds, const double* initial) {
    if (!s || !seeds) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    const auto ctx_lock = sepaihrd_internal::lock(s->ctx);
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(s->ctx);
    SW_TRY(cudaSetDevice(d.device));
    cudaStream_t st = sepaihrd_internal::stream(s->ctx);
    SW_TRY(cudaStreamSynchronize(st));                       // the pinned staging buffers are reused
    sepaihrd_rc rc = upload_seeds(s, seeds, st);
    if (rc != SEPAIHRD_OK) return rc;
    if (initial) {
        for (int k = 0; k < s->P; ++k) s->h_gbest[k] = initial[k];
        SW_TRY(cudaMemcpyAsync(s->d_init, s->h_gbest, sizeof(double) * s->P, cudaMemcpyHostToDevice, st));
    }
    s->evaluated_once = false;
    if (s->local > 0) {
        const int threads = rng_block(s->local);
        const unsigned blocks = (unsigned)((s->local + threads - 1) / threads);
        swarm_init_kernel<<<blocks, threads, 0, st>>>(s->local, s->offset, s->P, s->d_seeds, s->d_lb, s->d_ub,
                                                                 initial ? s->d_init : nullptr, s->d_pos, s->d_vel);
        SW_TRY(cudaGetLastError());
        sepaihrd_internal::count_launches(s->ctx, 1);
    }
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_swarm_evaluate(sepaihrd_swarm* s, double* out_best_value, int64_t* out_best_local_index, double* out_best_position) {
    if (!s) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    const auto ctx_lock = sepaihrd_internal::lock(s->ctx);
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(s->ctx);
    SW_TRY(cudaSetDevice(d.device));
    cudaStream_t st = sepaihrd_internal::stream(s->ctx);
    if (s->local == 0) {
        if (out_best_value) *out_best_value = -INFINITY;
        if (out_best_local_index) *out_best_local_index = -1;
        return SEPAIHRD_OK;
    }
    sepaihrd_rc rc = sepaihrd_internal::eval_batch_device_unordered(s->ctx, s->d_pos, s->local, s->P, s->d_fit, s->d_status, nullptr);
    if (rc != SEPAIHRD_OK) return rc;
    swarm_tell_kernel<<<s->blocks_tell, TELL_THREADS, 0, st>>>(s->local, s->P, s->evaluated_once ? 0 : 1, s->d_fit, s->d_pos, s->d_pbest,
                                                              s->d_pbest_val, s->d_block_val, s->d_block_idx);
    SW_TRY(cudaGetLastError());
    swarm_best_kernel<<<1, TELL_THREADS, 0, st>>>(s->blocks_tell, s->P, s->d_block_val, s->d_block_idx, s->d_pbest, s->d_best, 0);
    SW_TRY(cudaGetLastError());
    sepaihrd_internal::count_launches(s->ctx, 2);
    s->evaluated_once = true;
    SW_TRY(cudaMemcpyAsync(s->h_best, s->d_best, sizeof(double) * ((size_t)s->P + 2), cudaMemcpyDeviceToHost, st));
    SW_TRY(cudaStreamSynchronize(st));
    const long long bi = (long long)s->h_best[1];
    if (out_best_value) *out_best_value = s->h_best[0];
    if (out_best_local_index) *out_best_local_index = bi;
    if (out_best_position && bi >= 0)
        for (int k = 0; k < s->P; ++k) out_best_position[k] = s->h_best[2 + k];
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_swarm_step(sepaihrd_swarm* s, const uint32_t* seeds, double omega, double c1, double c2, const double* global_best) {
    if (!s || !seeds || !global_best) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (!s->evaluated_once) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "sepaihrd_swarm_step before the first sepaihrd_swarm_evaluate");
    const auto ctx_lock = sepaihrd_internal::lock(s->ctx);
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(s->ctx);
    SW_TRY(cudaSetDevice(d.device));
    cudaStream_t st = sepaihrd_internal::stream(s->ctx);
    SW_TRY(cudaStreamSynchronize(st));
    sepaihrd_rc rc = upload_seeds(s, seeds, st);
    if (rc != SEPAIHRD_OK) return rc;
    for (int k = 0; k < s->P; ++k) s->h_gbest[k] = global_best[k];
    SW_TRY(cudaMemcpyAsync(s->d_gbest, s->h_gbest, sizeof(double) * s->P, cudaMemcpyHostToDevice, st));
    if (s->local > 0) {
        const int threads = rng_block(s->local);
        const unsigned blocks = (unsigned)((s->local + threads - 1) / threads);
        swarm_step_kernel<<<blocks, threads, 0, st>>>(s->local, s->offset, s->P, s->d_seeds, s->d_lb, s->d_ub, s->d_gbest, s->d_pbest,
                                                                 omega, c1, c2, s->d_pos, s->d_vel);
        SW_TRY(cudaGetLastError());
        sepaihrd_internal::count_launches(s->ctx, 1);
    }
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_swarm_read(sepaihrd_swarm* s, int32_t what, double* out) {
    if (!s || !out) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    const auto ctx_lock = sepaihrd_internal::lock(s->ctx);
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(s->ctx);
    SW_TRY(cudaSetDevice(d.device));
    cudaStream_t st = sepaihrd_internal::stream(s->ctx);
    const double* src = nullptr;
    size_t count = (size_t)s->local * s->P;
    switch (what) {
        case SEPAIHRD_SWARM_POSITIONS: src = s->d_pos; break;
        case SEPAIHRD_SWARM_VELOCITIES: src = s->d_vel; break;
        case SEPAIHRD_SWARM_PERSONAL_BEST: src = s->d_pbest; break;
        case SEPAIHRD_SWARM_PERSONAL_BEST_VALUES: src = s->d_pbest_val; count = (size_t)s->local; break;
        case SEPAIHRD_SWARM_FITNESS: src = s->d_fit; count = (size_t)s->local; break;
        default: return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "bad swarm array selector");
    }
    if (count == 0) return SEPAIHRD_OK;
    SW_TRY(cudaMemcpyAsync(out, src, sizeof(double) * count, cudaMemcpyDeviceToHost, st));
    SW_TRY(cudaStreamSynchronize(st));
    return SEPAIHRD_OK;
}

// ---- asynchronous form: nothing below synchronises or copies per iteration -------------------------------------------------
sepaihrd_rc sepaihrd_swarm_upload_seeds(sepaihrd_swarm* s, const uint32_t* seeds, int64_t n_sets) {
    if (!s || !seeds || n_sets < 1) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument / no seed sets");
    const auto ctx_lock = sepaihrd_internal::lock(s->ctx);
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(s->ctx);
    SW_TRY(cudaSetDevice(d.device));
    cudaStream_t st = sepaihrd_internal::stream(s->ctx);
    if (s->d_seed_sets) { SW_TRY(cudaDeviceSynchronize()); cudaFree(s->d_seed_sets); s->d_seed_sets = nullptr; s->n_seed_sets = 0; }
    const size_t bytes = sizeof(unsigned) * (size_t)n_sets * (size_t)s->swarm_size;
    SW_TRY(cudaMalloc((void**)&s->d_seed_sets, bytes));
    SW_TRY(cudaMemcpyAsync(s->d_seed_sets, seeds, bytes, cudaMemcpyHostToDevice, st));
    SW_TRY(cudaStreamSynchronize(st));                       // the source is the caller's buffer
    s->n_seed_sets = n_sets;
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_swarm_init_async(sepaihrd_swarm* s, const double* initial) {
    if (!s) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (s->n_seed_sets < 1) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "sepaihrd_swarm_init_async before sepaihrd_swarm_upload_seeds");
    const auto ctx_lock = sepaihrd_internal::lock(s->ctx);
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(s->ctx);
    SW_TRY(cudaSetDevice(d.device));
    cudaStream_t st = sepaihrd_internal::stream(s->ctx);
    if (initial) {
        SW_TRY(cudaStreamSynchronize(st));                   // pinned staging reused
        for (int k = 0; k < s->P; ++k) s->h_gbest[k] = initial[k];
        SW_TRY(cudaMemcpyAsync(s->d_init, s->h_gbest, sizeof(double) * s->P, cudaMemcpyHostToDevice, st));
    }
    s->evaluated_once = false;
    swarm_reset_best_kernel<<<1, 1, 0, st>>>(s->d_gbest_val);
    SW_TRY(cudaGetLastError());
    if (s->local > 0) {
        const int threads = rng_block(s->local);
        const unsigned blocks = (unsigned)((s->local + threads - 1) / threads);
        swarm_init_kernel<<<blocks, threads, 0, st>>>(s->local, s->offset, s->P, s->d_seed_sets, s->d_lb, s->d_ub,
                                                                 initial ? s->d_init : nullptr, s->d_pos, s->d_vel);
        SW_TRY(cudaGetLastError());
    }
    sepaihrd_internal::count_launches(s->ctx, 2);
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_swarm_evaluate_async(sepaihrd_swarm* s) {
    if (!s) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    const auto ctx_lock = sepaihrd_internal::lock(s->ctx);
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(s->ctx);
    SW_TRY(cudaSetDevice(d.device));
    cudaStream_t st = sepaihrd_internal::stream(s->ctx);
    if (s->local > 0) {
        sepaihrd_rc rc = sepaihrd_internal::eval_batch_device_unordered(s->ctx, s->d_pos, s->local, s->P, s->d_fit, s->d_status, nullptr);
        if (rc != SEPAIHRD_OK) return rc;
        swarm_tell_kernel<<<s->blocks_tell, TELL_THREADS, 0, st>>>(s->local, s->P, s->evaluated_once ? 0 : 1, s->d_fit, s->d_pos, s->d_pbest,
                                                                  s->d_pbest_val, s->d_block_val, s->d_block_idx);
        SW_TRY(cudaGetLastError());
    }
    // an empty shard still publishes a record: (-inf, -1)
    swarm_best_kernel<<<1, TELL_THREADS, 0, st>>>(s->blocks_tell, s->P, s->d_block_val, s->d_block_idx, s->d_pbest, s->d_best, s->offset);
    SW_TRY(cudaGetLastError());
    sepaihrd_internal::count_launches(s->ctx, 2);
    s->evaluated_once = true;
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_swarm_record_device(sepaihrd_swarm* s, const double** d_record, int32_t* record_doubles) {
    if (!s || !d_record) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    *d_record = s->d_best;
    if (record_doubles) *record_doubles = s->P + 2;
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_swarm_adopt_global_best(sepaihrd_swarm* s, const double* d_records, int32_t n_records, int64_t record_stride,
                                             int32_t trace_slot) {
    if (!s || !d_records || n_records < 1 || record_stride < s->P + 2) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "bad records");
    if (trace_slot >= SEPAIHRD_SWARM_TRACE_CAPACITY) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "trace slot beyond SEPAIHRD_SWARM_TRACE_CAPACITY");
    const auto ctx_lock = sepaihrd_internal::lock(s->ctx);
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(s->ctx);
    SW_TRY(cudaSetDevice(d.device));
    swarm_adopt_kernel<<<1, 64, 0, sepaihrd_internal::stream(s->ctx)>>>(d_records, n_records, record_stride, s->P, s->d_gbest, s->d_gbest_val,
                                                                          trace_slot >= 0 ? s->d_trace : nullptr, trace_slot);
    SW_TRY(cudaGetLastError());
    sepaihrd_internal::count_launches(s->ctx, 1);
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_swarm_step_async(sepaihrd_swarm* s, int32_t iteration, double omega, double c1, double c2) {
    if (!s) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (!s->evaluated_once) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "sepaihrd_swarm_step_async before the first evaluation");
    if (iteration < 0 || iteration + 1 >= s->n_seed_sets) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "no seed set uploaded for this iteration");
    const auto ctx_lock = sepaihrd_internal::lock(s->ctx);
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(s->ctx);
    SW_TRY(cudaSetDevice(d.device));
    if (s->local > 0) {
        const int threads = rng_block(s->local);
        const unsigned blocks = (unsigned)((s->local + threads - 1) / threads);
        swarm_step_kernel<<<blocks, threads, 0, sepaihrd_internal::stream(s->ctx)>>>(
            s->local, s->offset, s->P, s->d_seed_sets + (size_t)(iteration + 1) * (size_t)s->swarm_size, s->d_lb, s->d_ub, s->d_gbest, s->d_pbest,
            omega, c1, c2, s->d_pos, s->d_vel);
        SW_TRY(cudaGetLastError());
        sepaihrd_internal::count_launches(s->ctx, 1);
    }
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_swarm_read_trace(sepaihrd_swarm* s, double* out, int32_t n) {
    if (!s || !out || n < 0 || n > SEPAIHRD_SWARM_TRACE_CAPACITY) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "bad trace request");
    const auto ctx_lock = sepaihrd_internal::lock(s->ctx);
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(s->ctx);
    SW_TRY(cudaSetDevice(d.device));
    cudaStream_t st = sepaihrd_internal::stream(s->ctx);
    if (n > 0) SW_TRY(cudaMemcpyAsync(out, s->d_trace, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    SW_TRY(cudaStreamSynchronize(st));
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_swarm_read_global_best(sepaihrd_swarm* s, double* out_value, double* out_position) {
    if (!s) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    const auto ctx_lock = sepaihrd_internal::lock(s->ctx);
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(s->ctx);
    SW_TRY(cudaSetDevice(d.device));
    cudaStream_t st = sepaihrd_internal::stream(s->ctx);
    if (out_value) SW_TRY(cudaMemcpyAsync(out_value, s->d_gbest_val, sizeof(double), cudaMemcpyDeviceToHost, st));
    if (out_position) SW_TRY(cudaMemcpyAsync(out_position, s->d_gbest, sizeof(double) * (size_t)s->P, cudaMemcpyDeviceToHost, st));
    SW_TRY(cudaStreamSynchronize(st));
    return SEPAIHRD_OK;
}

}  // extern "C"
