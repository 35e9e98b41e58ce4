// sepaihrd_exchange.cu -- all-gather of small records between the GPUs of one node, written as ONE kernel over NVLink peer
// memory (CUDA IPC mailboxes): the collective the callers of the hot path need once per iteration.
//
//   Metropolis-Hastings, 4096 chains on G GPUs: every rank contributes its B/G current log-likelihoods (<= 32 KiB in total)
//       reference call site  src/sir_age_structured/optimizers/MetropolisHastingsSampler.cpp:312-330
//   particle swarm, 65,536 particles on G GPUs: every rank contributes [best value, particle index, position[P]] (512 B)
//       reference call sites src/model/optimizers/ParticleSwarmOptimizer.cpp:149-156, 417-421
//
// These messages are latency-bound.  Round 1 moved them through numpy -> H2D -> dist.all_gather on a
// list of tensors -> one .cpu() per rank: 500-670 us per iteration.  Here every rank owns a mailbox in its own HBM
//     slots[2][world][cap] doubles + flags[2][world] 64-bit sequence numbers
// which every peer maps with cudaIpcOpenMemHandle.  An all-gather is a single kernel of `world` blocks on the ctx stream:
//   block p  1. stores this rank's record into slots[seq & 1][rank] of PEER p (plain stores over NVLink / NVSwitch),
//            2. makes them visible system-wide (__threadfence_system) and releases flags[seq & 1][rank] = seq at peer p,
//            3. spins (acquire, bounded by a timeout) on ITS OWN flags[seq & 1][p] until rank p's record of this round has
//               landed, and copies it to the destination.
// Every put precedes every wait inside a rank, so the exchange cannot deadlock as long as all ranks launch it; two slot
// parities suffice because a rank can only start round k + 2 after it has seen every peer's round k + 1, which the peer sent
// after it had consumed round k.  A peer that never shows up raises a status flag after the timeout instead of hanging the GPU.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sepaihrd_internal.h"

struct sepaihrd_exchange {
    sepaihrd_ctx* ctx = nullptr;
    int world = 1, rank = 0;
    long long cap = 0;                       // doubles per record slot
    char* d_box = nullptr;                   // own mailbox (cudaMalloc): slots | flags | status
    size_t box_bytes = 0, flags_off = 0, status_off = 0;
    std::vector<char*> peer;                 // mailbox base of every rank in THIS process's address space (own entry = d_box)
    char** d_peer = nullptr;                 // device copy of `peer`
    unsigned long long seq = 0;
    bool connected = false;
    double timeout_s = 10.0;
};

namespace {

using sepaihrd_internal::fail_with;

#define EX_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) return fail_with(SEPAIHRD_ERR_CUDA, (std::string(#expr) + ": " + cudaGetErrorString(e__)).c_str()); \
    } while (0)

__device__ __forceinline__ void store_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long load_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr int EX_THREADS = 128;

__global__ void __launch_bounds__(EX_THREADS) exchange_all_gather_kernel(char* const* __restrict__ peer, int world, int rank, long long cap,
                                                                          size_t flags_off, size_t status_off, unsigned long long seq,
                                                                          const double* __restrict__ src, long long count, double* __restrict__ dst,
                                                                          unsigned long long timeout_ns) {
    const int p = blockIdx.x;                           // the peer this block sends to and receives from
    const int par = (int)(seq & 1ull);
    // ---- put: my record into peer p's mailbox --------------------------------------------------------------------
    {
        double* slot = reinterpret_cast<double*>(peer[p]) + ((long long)par * world + rank) * cap;
        for (long long i = threadIdx.x; i < count; i += EX_THREADS) slot[i] = src[i];
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0)
            store_release_sys(reinterpret_cast<unsigned long long*>(peer[p] + flags_off) + (par * world + rank), seq);
    }
    // ---- get: rank p's record from my own mailbox ------------------------------------------------------------------
    __shared__ int ok;
    if (threadIdx.x == 0) {
        const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(peer[rank] + flags_off) + (par * world + p);
        const unsigned long long t0 = global_timer_ns();
        int good = 1;
        while (load_acquire_sys(flag) != seq) {
            if (global_timer_ns() - t0 > timeout_ns) { good = 0; break; }
            __nanosleep(64);
        }
        if (!good) atomicExch(reinterpret_cast<unsigned*>(peer[rank] + status_off), 1u + (unsigned)p);
        ok = good;
    }
    __syncthreads();
    if (ok) {
        const double* slot = reinterpret_cast<const double*>(peer[rank]) + ((long long)par * world + p) * cap;
        for (long long i = threadIdx.x; i < count; i += EX_THREADS) dst[(long long)p * count + i] = __ldcv(slot + i);
    }
}

}  // namespace

extern "C" {

sepaihrd_rc sepaihrd_exchange_create(sepaihrd_ctx* ctx, int32_t world, int32_t rank, int64_t max_doubles, sepaihrd_exchange** out,
                                     unsigned char* out_handle) {
    if (!ctx || !out || !out_handle) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (world < 1 || rank < 0 || rank >= world || max_doubles < 1) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "bad world / rank / record size");
    static_assert(sizeof(cudaIpcMemHandle_t) == SEPAIHRD_EXCHANGE_HANDLE_BYTES, "IPC handle size");
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(ctx);
    EX_TRY(cudaSetDevice(d.device));
    auto* ex = new sepaihrd_exchange;
    ex->ctx = ctx; ex->world = world; ex->rank = rank;
    ex->cap = (max_doubles + 31) & ~31ll;                  // 256-byte slots
    const size_t slots = sizeof(double) * 2 * (size_t)world * (size_t)ex->cap;
    ex->flags_off = slots;
    ex->status_off = ex->flags_off + sizeof(unsigned long long) * 2 * (size_t)world;
    // a multiple of 2 MiB: the mailbox is then an allocation of its own (small cudaMallocs may share one, and an IPC handle
    // always opens at the base of the underlying allocation)
    ex->box_bytes = (ex->status_off + 256 + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    if (const char* t = std::getenv("SEPAIHRD_EXCHANGE_TIMEOUT_S")) { const double v = std::atof(t); if (v > 0) ex->timeout_s = v; }
    cudaError_t e = cudaMalloc((void**)&ex->d_box, ex->box_bytes);
    if (e == cudaSuccess) e = cudaMemset(ex->d_box, 0, ex->box_bytes);          // flags start at sequence 0; the first round is 1
    if (e == cudaSuccess) e = cudaMalloc((void**)&ex->d_peer, sizeof(char*) * (size_t)world);
    cudaIpcMemHandle_t h;
    std::memset(&h, 0, sizeof(h));
    if (e == cudaSuccess && world > 1) e = cudaIpcGetMemHandle(&h, ex->d_box);
    if (e != cudaSuccess) {
        const std::string msg = std::string("exchange mailbox: ") + cudaGetErrorString(e);
        cudaGetLastError();
        sepaihrd_exchange_destroy(ex);
        return fail_with(e == cudaErrorMemoryAllocation ? SEPAIHRD_ERR_OUT_OF_MEMORY : SEPAIHRD_ERR_CUDA, msg.c_str());
    }
    std::memcpy(out_handle, &h, sizeof(h));
    ex->peer.assign((size_t)world, nullptr);
    ex->peer[(size_t)rank] = ex->d_box;
    if (world == 1) {
        e = cudaMemcpy(ex->d_peer, ex->peer.data(), sizeof(char*), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { sepaihrd_exchange_destroy(ex); return fail_with(SEPAIHRD_ERR_CUDA, cudaGetErrorString(e)); }
        ex->connected = true;
    }
    *out = ex;
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_exchange_connect(sepaihrd_exchange* ex, const unsigned char* handles) {
    if (!ex || !handles) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(ex->ctx);
    EX_TRY(cudaSetDevice(d.device));
    for (int r = 0; r < ex->world; ++r) {
        if (r == ex->rank || ex->peer[(size_t)r]) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)r * SEPAIHRD_EXCHANGE_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail_with(SEPAIHRD_ERR_CUDA, (std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(r) + "): " + cudaGetErrorString(e)).c_str());
        }
        ex->peer[(size_t)r] = static_cast<char*>(p);
    }
    EX_TRY(cudaMemcpy(ex->d_peer, ex->peer.data(), sizeof(char*) * (size_t)ex->world, cudaMemcpyHostToDevice));
    ex->connected = true;
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_exchange_all_gather(sepaihrd_exchange* ex, const double* d_src, int64_t count, double* d_dst) {
    if (!ex || !d_src || !d_dst) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (!ex->connected) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "sepaihrd_exchange_all_gather before sepaihrd_exchange_connect");
    if (count < 1 || count > ex->cap) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "record larger than the mailbox slots");
    const auto ctx_lock = sepaihrd_internal::lock(ex->ctx);
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(ex->ctx);
    EX_TRY(cudaSetDevice(d.device));
    ex->seq += 1;
    exchange_all_gather_kernel<<<ex->world, EX_THREADS, 0, sepaihrd_internal::stream(ex->ctx)>>>(
        ex->d_peer, ex->world, ex->rank, ex->cap, ex->flags_off, ex->status_off, ex->seq, d_src, count, d_dst,
        (unsigned long long)(ex->timeout_s * 1e9));
    EX_TRY(cudaGetLastError());
    sepaihrd_internal::count_launches(ex->ctx, 1);
    return SEPAIHRD_OK;
}

sepaihrd_rc sepaihrd_exchange_status(sepaihrd_exchange* ex, int32_t* out_status) {
    if (!ex || !out_status) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    const auto ctx_lock = sepaihrd_internal::lock(ex->ctx);
    const sepaihrd_internal::Dims d = sepaihrd_internal::dims(ex->ctx);
    EX_TRY(cudaSetDevice(d.device));
    EX_TRY(cudaStreamSynchronize(sepaihrd_internal::stream(ex->ctx)));
    unsigned v = 0;
    EX_TRY(cudaMemcpy(&v, ex->d_box + ex->status_off, sizeof(v), cudaMemcpyDeviceToHost));
    *out_status = (int32_t)v;
    return SEPAIHRD_OK;
}

void sepaihrd_exchange_destroy(sepaihrd_exchange* ex) {
    if (!ex) return;
    cudaDeviceSynchronize();
    for (int r = 0; r < (int)ex->peer.size(); ++r)
        if (r != ex->rank && ex->peer[(size_t)r]) cudaIpcCloseMemHandle(ex->peer[(size_t)r]);
    if (ex->d_peer) cudaFree(ex->d_peer);
    if (ex->d_box) cudaFree(ex->d_box);
    delete ex;
}

}  // extern "C"
