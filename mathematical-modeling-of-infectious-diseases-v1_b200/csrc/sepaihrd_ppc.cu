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
#include <mutex>
#include <vector>

#include <cooperative_groups.h>

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
            out_all[(size_t)col * Q_total + q_first + q] = __dadd_rn(a, __dmul_rn(h - (double)i0, c - a));   // unfused: the same bits as the host formula
        }
        __syncthreads();
    }
}

// ---- the same selection with the column held ON CHIP ---------------------------------------------------------------------
// ppc_select_kernel above is instruction-bound (ncu, 100 k draws x 4 ages: 142 thread-instructions per key, 56 % issue slots
// busy) and reads every column 3.6 x from DRAM, because 2 x 148 resident columns of 800 KB do not fit the L2.  Here a CLUSTER
// of CL thread blocks owns a column: block r loads draws [r S, (r + 1) S) ONCE, as order-preserving keys (high and low words in
// separate arrays), into its shared memory (CL = 4: 4 x 200 KB hold 100 k draws); every later sweep runs on shared memory.
// The narrowing works on rel = key - (smallest high word << 32): its first 12-bit digit covers exactly the spread of the
// column (2048-4096 cells in use whatever binade boundary the values straddle), so ~25-50 keys share a cell and one step is
// enough for the usual column; while that digit lies in the high word a sweep reads 4 bytes per key.
// The blocks exchange only small things through distributed shared memory:
//   * after the load: min / max high word and count of each slice (every block combines all CL triples itself);
//   * per narrowing step: each block histograms its slice, then every thread sums ITS four cells over the CL histograms
//     (one 16-byte remote load per block), a block-wide scan finds the cell of every wanted rank -- redundantly and
//     identically in every block, so no result has to be sent back;
//   * the <= CS_CAP survivors of a rank's bucket are PUSHED (remote atomic + remote store) into the list of the block that
//     owns the bucket (bucket a -> block a mod CL), which picks the rank by counting and writes the order statistic.
// The linear interpolation between order statistics is a separate tiny kernel (ppc_finalize_kernel), so no block waits for
// another one's pick.  Barriers per column: 3 cluster-wide (after load, after histogram, after push) + 2 per extra step.
// The bookkeeping between sweeps (wanted ranks, distinct buckets) is done by one WARP with match / ballot, not by one thread.
namespace cg = cooperative_groups;
constexpr int CS_THREADS = 1024, CS_CELLS = 4096, CS_MAXT = 16, CS_CAP = 192, CS_LOADS = 8;
constexpr unsigned CS_NAN_HI = 0xffffffffu;              // high word of the key NaNs are stored as (~0: no double maps to it)

struct CsCtl {
    unsigned st_min[2], st_max[2], st_cnt[2];            // this block's slice, double-buffered by column parity (read by the others)
    unsigned long long t_prefix[CS_MAXT], b_prefix[CS_MAXT];
    long long t_want[CS_MAXT];
    unsigned t_rank[CS_MAXT], t_size[CS_MAXT], n_rank[CS_MAXT], n_size[CS_MAXT], n_digit[CS_MAXT], base[CS_MAXT], bucket_base[CS_MAXT];
    int t_bucket[CS_MAXT];
    unsigned own_n[CS_MAXT];                              // keys pushed into the lists this block owns
    unsigned warp_tot[32];
    unsigned cnt, base_hi;
    int n_targets, n_buckets, shift, go;
};
constexpr size_t cs_ctl_bytes() { return (sizeof(CsCtl) + 15) & ~(size_t)15; }
template <int CL> constexpr int cs_own_slots() { return (CS_MAXT + CL - 1) / CL; }
// control block, histogram, first-digit -> bucket map, owned candidate lists; the keys follow
template <int CL> constexpr size_t cs_fixed_bytes() {
    return cs_ctl_bytes() + sizeof(unsigned) * CS_CELLS + CS_CELLS + sizeof(unsigned long long) * cs_own_slots<CL>() * CS_CAP;
}

__device__ __forceinline__ unsigned long long cs_rel(unsigned hi, unsigned lo, unsigned base_hi) {
    return ((unsigned long long)(hi - base_hi) << 32) | (unsigned long long)lo;
}

template <int CL>
__global__ void __launch_bounds__(CS_THREADS, 1) ppc_select_cluster_kernel(const double* __restrict__ series, long long B, long long n_cols, int Q,
                                                                           const double* __restrict__ probs, double* __restrict__ tvals,
                                                                           long long* __restrict__ col_cnt) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned cr = cluster.block_rank();
    const long long cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
    extern __shared__ __align__(16) unsigned char cs_raw[];
    CsCtl* ctl = reinterpret_cast<CsCtl*>(cs_raw);
    unsigned* hist = reinterpret_cast<unsigned*>(cs_raw + cs_ctl_bytes());
    unsigned char* cellmap = reinterpret_cast<unsigned char*>(hist + CS_CELLS);   // first digit -> 0 (no live bucket), bucket + 1, or 255 (several)
    unsigned long long* lists = reinterpret_cast<unsigned long long*>(cellmap + CS_CELLS);
    const int tid = threadIdx.x, lane = tid & 31;
    const long long S = (B + CL - 1) / CL;
    const long long s0 = (cr * S < B) ? cr * S : B, s1 = (s0 + S < B) ? s0 + S : B;
    const int nl = (int)(s1 - s0);
    unsigned* khi = reinterpret_cast<unsigned*>(lists + cs_own_slots<CL>() * CS_CAP);
    unsigned* klo = khi + ((S + 3) & ~3LL);
    for (int i = tid; i < CS_CELLS / 4; i += CS_THREADS) reinterpret_cast<unsigned*>(cellmap)[i] = 0u;
    // the only read of a column: CS_LOADS coalesced loads in flight per thread, keys to shared memory, slice statistics
    auto load_column = [&](long long col, int parity) {
        const double* v = series + (size_t)col * B + s0;
        if (tid == 0) { ctl->st_min[parity] = 0xffffffffu; ctl->st_max[parity] = 0u; ctl->st_cnt[parity] = 0u; }
        __syncthreads();
        unsigned lo_hi = 0xffffffffu, hi_hi = 0u, c = 0u;
        auto take = [&](int i, double x) {
            const unsigned long long b = (unsigned long long)__double_as_longlong(x);
            unsigned h = (unsigned)(b >> 32), l = (unsigned)b;
            const unsigned sgn = (unsigned)((int)h >> 31);          // all ones for a negative value
            h ^= sgn | 0x80000000u; l ^= sgn;                          // sel_key, word by word
            if (x != x) { h = CS_NAN_HI; l = 0xffffffffu; } else { lo_hi = min(lo_hi, h); hi_hi = max(hi_hi, h); ++c; }
            khi[i] = h; klo[i] = l;
        };
        int i0 = tid;
        for (; i0 + (CS_LOADS - 1) * CS_THREADS < nl; i0 += CS_LOADS * CS_THREADS) {
            double xs[CS_LOADS];
#pragma unroll
            for (int u = 0; u < CS_LOADS; ++u) xs[u] = __ldg(v + i0 + u * CS_THREADS);
#pragma unroll
            for (int u = 0; u < CS_LOADS; ++u) take(i0 + u * CS_THREADS, xs[u]);
        }
        if (i0 < nl) {
            double xs[CS_LOADS];
#pragma unroll
            for (int u = 0; u < CS_LOADS; ++u) { const int i = i0 + u * CS_THREADS; xs[u] = (i < nl) ? __ldg(v + i) : 0.0; }
#pragma unroll
            for (int u = 0; u < CS_LOADS; ++u) { const int i = i0 + u * CS_THREADS; if (i < nl) take(i, xs[u]); }
        }
        lo_hi = __reduce_min_sync(0xffffffffu, lo_hi); hi_hi = __reduce_max_sync(0xffffffffu, hi_hi); c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) { atomicMin(&ctl->st_min[parity], lo_hi); atomicMax(&ctl->st_max[parity], hi_hi); atomicAdd(&ctl->st_cnt[parity], c); }
    };
    if (tid < CS_MAXT) ctl->own_n[tid] = 0u;
    int parity = 0;
    if (cluster_id < n_cols) load_column(cluster_id, 0);
    // Software pipeline: the keys of a column are dead once its survivors are pushed, so the NEXT column is loaded between the
    // arrival at the "lists complete" barrier and the wait on it -- the load latency covers the barrier skew and vice versa.
    for (long long col = cluster_id; col < n_cols; col += n_clusters, parity ^= 1) {
        const long long next_col = col + n_clusters;
        bool pushed = false;
        cluster.sync();                                                                    // every slice is loaded and summarised
        if (tid < 32) {                                                                    // one warp: column statistics, wanted ranks
            unsigned mn = 0xffffffffu, mx = 0u, cn = 0u;
            if (lane < CL) {
                const CsCtl* rc = cluster.map_shared_rank(ctl, lane);
                mn = rc->st_min[parity]; mx = rc->st_max[parity]; cn = rc->st_cnt[parity];
            }
            mn = __reduce_min_sync(0xffffffffu, mn); mx = __reduce_max_sync(0xffffffffu, mx); cn = __reduce_add_sync(0xffffffffu, cn);
            const long long cnt = (long long)cn;
            long long r = -1 - lane;                                                       // lanes without a rank: distinct, negative
            const bool has = (cnt > 0) && (lane < 2 * Q);
            if (has) {
                const double h = (double)(cnt - 1) * probs[lane >> 1];
                long long i0 = (long long)floor(h);
                if (i0 < 0) i0 = 0;
                if (i0 > cnt - 1) i0 = cnt - 1;
                const long long i1 = (i0 + 1 < cnt) ? i0 + 1 : i0;
                r = (lane & 1) ? i1 : i0;
            }
            const unsigned same = __match_any_sync(0xffffffffu, r);
            const bool first = has && ((same & ((1u << lane) - 1u)) == 0u);               // ranks without repeats, in order of appearance
            const unsigned fm = __ballot_sync(0xffffffffu, first);
            if (first) {
                const int t = __popc(fm & ((1u << lane) - 1u));
                ctl->t_want[t] = r; ctl->t_rank[t] = (unsigned)r; ctl->t_prefix[t] = 0ULL; ctl->t_size[t] = cn;
            }
            if (lane == 0) {
                ctl->n_targets = __popc(fm); ctl->cnt = cn; ctl->base_hi = mn;
                ctl->shift = (cn > 0) ? 32 + (32 - __clz(mx - mn)) : 0;                    // rel < 2^shift
                if (cr == 0) col_cnt[col] = cnt;
            }
        }
        __syncthreads();
        int first_nshift = 0;                                                              // where the first digit sits: what cellmap is keyed by
        int round = 0;
        if (ctl->cnt != 0u) {                                                              // (no valid draw: nothing to select; every block sees the same count)
        const unsigned base_hi = ctl->base_hi;
        while (true) {
            if (tid < 32) {                                                                // one warp: go on?  distinct buckets
                const int nt = ctl->n_targets;
                const bool valid = lane < nt;
                const unsigned long long p = valid ? ctl->t_prefix[lane] : 0ULL;
                const bool large = valid && (ctl->t_size[lane] > (unsigned)CS_CAP);
                const unsigned vm = __ballot_sync(0xffffffffu, valid);
                const unsigned any_large = __ballot_sync(0xffffffffu, large);
                unsigned same = 0u;
                if (valid) same = __match_any_sync(vm, p);
                const int leader = valid ? (__ffs(same) - 1) : 0;
                const unsigned lm = __ballot_sync(0xffffffffu, valid && leader == lane);
                if (valid) {
                    const int a = __popc(lm & ((1u << leader) - 1u));
                    ctl->t_bucket[lane] = a;
                    if (leader == lane) ctl->b_prefix[a] = p;
                }
                if (lane == 0) { ctl->n_buckets = __popc(lm); ctl->go = (any_large != 0u) && (ctl->shift > 0); }
            }
            __syncthreads();
            if (!ctl->go) break;
            const int nb = ctl->n_buckets;
            const int dig = (nb == 1) ? 12 : (nb <= 4) ? 10 : 8;
            const int shift = ctl->shift, nshift = (shift > dig) ? shift - dig : 0, width = shift - nshift;
            const int ncells = nb << width;
            const unsigned dmask = (1u << width) - 1u;
            if (round == 0) first_nshift = nshift;
            else cluster.sync();                                                           // nobody reads the previous step's histograms any more
            for (int i = tid; i < ncells; i += CS_THREADS) hist[i] = 0u;
            if (round > 0 && tid == 0) {                                                   // first digit -> live bucket(s)
                for (int a = 0; a < nb; ++a) {
                    const unsigned d1 = (unsigned)(ctl->b_prefix[a] >> (first_nshift - shift));
                    cellmap[d1] = (cellmap[d1] == 0) ? (unsigned char)(a + 1) : (unsigned char)255;
                }
            }
            __syncthreads();
            if (round == 0) {                                                              // every valid key: rel < 2^shift, one bucket
                if (nshift >= 32) {
                    const int sh = nshift - 32;
#pragma unroll 4
                    for (int i = tid; i < nl; i += CS_THREADS) {
                        const unsigned h = khi[i];
                        if (h != CS_NAN_HI) atomicAdd(&hist[(h - base_hi) >> sh], 1u);
                    }
                } else {
#pragma unroll 4
                    for (int i = tid; i < nl; i += CS_THREADS) {
                        const unsigned h = khi[i];
                        if (h != CS_NAN_HI) atomicAdd(&hist[(unsigned)(cs_rel(h, klo[i], base_hi) >> nshift)], 1u);
                    }
                }
            } else {
#pragma unroll 4
                for (int i = tid; i < nl; i += CS_THREADS) {
                    const unsigned h = khi[i];
                    if (h == CS_NAN_HI) continue;
                    const unsigned long long rel = cs_rel(h, klo[i], base_hi);
                    const unsigned cm = cellmap[(unsigned)(rel >> first_nshift)];
                    if (cm != 0u) {
                        const unsigned long long hi = (shift < 64) ? (rel >> shift) : 0ULL;
                        const unsigned d = (unsigned)(rel >> nshift) & dmask;
                        if (cm != 255u) { if (hi == ctl->b_prefix[cm - 1]) atomicAdd(&hist[((cm - 1) << width) + d], 1u); }
                        else for (int a = 0; a < nb; ++a) if (hi == ctl->b_prefix[a]) atomicAdd(&hist[((unsigned)a << width) + d], 1u);
                    }
                }
            }
            cluster.sync();                                                                // all CL histograms are complete
            if (round > 0 && tid < nb) cellmap[(unsigned)(ctl->b_prefix[tid] >> (first_nshift - shift))] = 0;     // leave the map clean
            // my four cells, summed over the blocks; exclusive scan over all cells
            unsigned h4[4] = {0u, 0u, 0u, 0u};
            const int c0 = 4 * tid;
            if (c0 < ncells) {
#pragma unroll
                for (int r = 0; r < CL; ++r) {
                    const uint4 hv = *reinterpret_cast<const uint4*>(cluster.map_shared_rank(hist, r) + c0);
                    h4[0] += hv.x; h4[1] += hv.y; h4[2] += hv.z; h4[3] += hv.w;
                }
                if (c0 + 1 >= ncells) h4[1] = 0u;
                if (c0 + 2 >= ncells) h4[2] = 0u;
                if (c0 + 3 >= ncells) h4[3] = 0u;
            }
            const unsigned mine = h4[0] + h4[1] + h4[2] + h4[3];
            unsigned incl = mine;
            for (int o = 1; o < 32; o <<= 1) { const unsigned up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
            if (lane == 31) ctl->warp_tot[tid >> 5] = incl;
            __syncthreads();
            if (tid < 32) {
                unsigned w = ctl->warp_tot[tid], wi = w;
                for (int o = 1; o < 32; o <<= 1) { const unsigned up = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += up; }
                ctl->warp_tot[tid] = wi - w;
            }
            __syncthreads();
            const unsigned excl = ctl->warp_tot[tid >> 5] + (incl - mine);                  // counts fit 32 bits: a column has < 2^31 draws
            // where every bucket starts in the scan: its first cell belongs to the thread that holds cell (a << width)
            if (c0 < ncells) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ci = c0 + j;
                    if (ci < ncells && (ci & (int)dmask) == 0) {
                        unsigned e = excl;
                        for (int jj = 0; jj < j; ++jj) e += h4[jj];
                        ctl->bucket_base[ci >> width] = e;
                    }
                }
            }
            __syncthreads();
            const int nt = ctl->n_targets;
            if (tid < CS_MAXT) ctl->base[tid] = (tid < nt) ? ctl->bucket_base[ctl->t_bucket[tid]] + ctl->t_rank[tid] : 0xffffffffu;
            __syncthreads();
            if (mine != 0u) {
#pragma unroll
                for (int t = 0; t < CS_MAXT; ++t) {
                    const unsigned pos = ctl->base[t];
                    if (pos - excl < mine) {                                               // excl <= pos < excl + mine (unsigned wrap-around)
                        unsigned e = excl;
                        int j = 0;
                        for (; j < 3; ++j) { if (pos - e < h4[j]) break; e += h4[j]; }
                        ctl->n_rank[t] = pos - e;
                        ctl->n_size[t] = h4[j];
                        ctl->n_digit[t] = (unsigned)(c0 + j) & dmask;
                    }
                }
            }
            __syncthreads();
            if (tid < nt) {
                ctl->t_prefix[tid] = (ctl->t_prefix[tid] << width) | (unsigned long long)ctl->n_digit[tid];
                ctl->t_rank[tid] = ctl->n_rank[tid];
                ctl->t_size[tid] = ctl->n_size[tid];
            }
            if (tid == 0) ctl->shift = nshift;
            __syncthreads();
            ++round;
        }
        const int shift = ctl->shift, nt = ctl->n_targets, nb = ctl->n_buckets;
        bool small = true;
        for (int t = 0; t < nt; ++t) small = small && (ctl->t_size[t] <= (unsigned)CS_CAP);
        if (!small) {
            // bits used up with a large bucket: all its keys are equal, the prefix IS rel
            if (cr == 0 && tid < nt) tvals[(size_t)col * CS_MAXT + tid] = sel_value(ctl->t_prefix[tid] + ((unsigned long long)base_hi << 32));
        } else {
        if (round > 0) {                                                                   // first digit -> live bucket(s)
            if (tid == 0) {
                for (int a = 0; a < nb; ++a) {
                    const unsigned d1 = (unsigned)(ctl->b_prefix[a] >> (first_nshift - shift));
                    cellmap[d1] = (cellmap[d1] == 0) ? (unsigned char)(a + 1) : (unsigned char)255;
                }
            }
            __syncthreads();
        }
        {   // push the survivors to the blocks that own their buckets
            auto push = [&](int a, unsigned h, unsigned l) {
                const int owner = a % CL, slot = a / CL;
                const unsigned pos = atomicAdd(cluster.map_shared_rank(&ctl->own_n[slot], owner), 1u);
                if (pos < (unsigned)CS_CAP) cluster.map_shared_rank(lists, owner)[slot * CS_CAP + pos] = ((unsigned long long)h << 32) | (unsigned long long)l;
            };
            if (round == 0) {                                                              // a column of <= CS_CAP valid draws, or of one value
                for (int i = tid; i < nl; i += CS_THREADS) { const unsigned h = khi[i]; if (h != CS_NAN_HI) push(0, h, klo[i]); }
            } else if (first_nshift >= 32) {
                const int sh = first_nshift - 32;
#pragma unroll 4
                for (int i = tid; i < nl; i += CS_THREADS) {
                    const unsigned h = khi[i];
                    if (h == CS_NAN_HI) continue;
                    const unsigned cm = cellmap[(h - base_hi) >> sh];
                    if (cm != 0u) {
                        const unsigned l = klo[i];
                        const unsigned long long rel = cs_rel(h, l, base_hi);
                        const unsigned long long hi = (shift < 64) ? (rel >> shift) : 0ULL;
                        if (cm != 255u) { if (hi == ctl->b_prefix[cm - 1]) push((int)cm - 1, h, l); }
                        else for (int a = 0; a < nb; ++a) if (hi == ctl->b_prefix[a]) push(a, h, l);
                    }
                }
            } else {
#pragma unroll 4
                for (int i = tid; i < nl; i += CS_THREADS) {
                    const unsigned h = khi[i];
                    if (h == CS_NAN_HI) continue;
                    const unsigned l = klo[i];
                    const unsigned long long rel = cs_rel(h, l, base_hi);
                    const unsigned cm = cellmap[(unsigned)(rel >> first_nshift)];
                    if (cm != 0u) {
                        const unsigned long long hi = (shift < 64) ? (rel >> shift) : 0ULL;
                        if (cm != 255u) { if (hi == ctl->b_prefix[cm - 1]) push((int)cm - 1, h, l); }
                        else for (int a = 0; a < nb; ++a) if (hi == ctl->b_prefix[a]) push(a, h, l);
                    }
                }
            }
        }
        cluster.barrier_arrive();                                                          // my pushes are out (release)
        pushed = true;
        }   // small
        }   // cnt != 0
        if (next_col < n_cols) load_column(next_col, parity ^ 1);
        if (pushed) {
            cluster.barrier_wait();                                                        // every list is complete
            const int shift = ctl->shift, nt = ctl->n_targets, nb = ctl->n_buckets;
            if (round > 0 && tid < nb) cellmap[(unsigned)(ctl->b_prefix[tid] >> (first_nshift - shift))] = 0;     // leave the map clean
            // every rank whose bucket this block owns is picked by its own group of CS_CAP threads, all groups at once
            const int grp = tid / CS_CAP, c = tid % CS_CAP;
            int seen = 0;
            for (int t = 0; t < nt; ++t) {
                const int a = ctl->t_bucket[t];
                if (a % CL != (int)cr) continue;
                if (seen++ % (CS_THREADS / CS_CAP) != grp) continue;
                const unsigned long long* cand = lists + (a / CL) * CS_CAP;
                const unsigned m = (ctl->own_n[a / CL] < (unsigned)CS_CAP) ? ctl->own_n[a / CL] : (unsigned)CS_CAP;
                const unsigned k = ctl->t_rank[t];
                if ((unsigned)c < m) {
                    const unsigned long long key = cand[c];
                    unsigned less = 0, eq = 0;
                    for (unsigned j = 0; j < m; ++j) { const unsigned long long o = cand[j]; less += (o < key); eq += (o == key); }
                    if (less <= k && k < less + eq) tvals[(size_t)col * CS_MAXT + t] = sel_value(key);   // every candidate of that value writes the same bits
                }
            }
        }
        __syncthreads();
        if (tid < CS_MAXT) ctl->own_n[tid] = 0u;                                           // refilled only after the next column's barriers
    }
    cluster.sync();                                                                        // no block leaves while another may still read its shared memory
}

// quantile q of column col = linear interpolation between the two order statistics the select left in tvals (target order
// rebuilt the way the select built it)
__global__ void ppc_finalize_kernel(const double* __restrict__ tvals, const long long* __restrict__ col_cnt, long long n_cols, int Q,
                                    const double* __restrict__ probs, double* __restrict__ out_all, int Q_total, int q_first) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= n_cols * Q) return;
    const long long col = idx / Q;
    const int q = (int)(idx % Q);
    const long long cnt = col_cnt[col];
    if (cnt <= 0) { out_all[(size_t)col * Q_total + q_first + q] = nan(""); return; }
    long long want[CS_MAXT];
    int nt = 0;
    double hq = 0.0;
    long long q_i0 = 0, q_i1 = 0;
    for (int qq = 0; qq <= q; ++qq) {
        const double h = (double)(cnt - 1) * probs[qq];
        long long i0 = (long long)floor(h);
        if (i0 < 0) i0 = 0;
        if (i0 > cnt - 1) i0 = cnt - 1;
        const long long i1 = (i0 + 1 < cnt) ? i0 + 1 : i0;
        for (int w = 0; w < 2; ++w) {
            const long long r = w ? i1 : i0;
            bool seen = false;
            for (int t = 0; t < nt; ++t) seen = seen || (want[t] == r);
            if (!seen) { want[nt] = r; ++nt; }
        }
        if (qq == q) { hq = h; q_i0 = i0; q_i1 = i1; }
    }
    double a = 0.0, c = 0.0;
    for (int t = 0; t < nt; ++t) {
        if (want[t] == q_i0) a = tvals[(size_t)col * CS_MAXT + t];
        if (want[t] == q_i1) c = tvals[(size_t)col * CS_MAXT + t];
    }
    out_all[(size_t)col * Q_total + q_first + q] = __dadd_rn(a, __dmul_rn(hq - (double)q_i0, c - a));   // unfused: the same bits as the host formula
}

__global__ void ppc_count_valid_kernel(const unsigned* st, long long B, unsigned long long* out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const unsigned ok = (i < B && st[i] == 0u) ? 1u : 0u;
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}


template <int CL>
cudaError_t launch_cluster_select(cudaStream_t s, int device, long long B, long long n_cols, int Q, const double* d_series, const double* d_probs,
                                  double* d_tvals, long long* d_colcnt) {
    auto kern = ppc_select_cluster_kernel<CL>;
    const size_t smem = cs_fixed_bytes<CL>() + 2 * sizeof(unsigned) * (size_t)((((B + CL - 1) / CL) + 3) & ~3LL);   // high and low key words
    static std::once_flag attr_once[64];
    cudaError_t err = cudaSuccess;
    std::call_once(attr_once[device & 63], [&] {
        int max_optin = 0;
        err = cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
        if (err == cudaSuccess) err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin);
    });
    if (err != cudaSuccess) return err;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.blockDim = dim3(CS_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cfg.gridDim = dim3(CL);
    int n_clusters = 0;
    err = cudaOccupancyMaxActiveClusters(&n_clusters, kern, &cfg);
    if (err != cudaSuccess) return err;
    if (n_clusters < 1) return cudaErrorLaunchOutOfResources;
    cfg.gridDim = dim3((unsigned)(std::min<long long>(n_cols, n_clusters) * CL));
    return cudaLaunchKernelEx(&cfg, kern, d_series, B, n_cols, Q, d_probs, d_tvals, d_colcnt);
}

// how many draws a block of the cluster kernel can hold
template <int CL>
long long cluster_slice_cap(int device) {
    int max_optin = 0;
    if (cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) return 0;
    const long long room = (long long)max_optin - (long long)cs_fixed_bytes<CL>();
    return room > 0 ? (room / (long long)sizeof(unsigned long long)) & ~3LL : 0;
}

// Quantiles of n_cols device columns of B values each: d_q[col][n_probs].  path: 0 = by size (the cluster kernel while a column fits
// the shared memory of 4 or 8 blocks, the one-block-per-column kernel beyond), 1 = one block per column, 2 = cluster kernel or fail.
sepaihrd_rc select_quantiles(sepaihrd_ctx* ctx, cudaStream_t s, int device, const double* d_series, long long B, long long all_cols, int n_probs,
                             const double* d_probs, double* d_q, double* d_tvals, long long* d_colcnt, int path) {
    using sepaihrd_internal::fail_with;
    static const char* env_path = std::getenv("SEPAIHRD_PPC_SELECT");
    if (path == 0 && env_path) path = (env_path[0] == 'b') ? 1 : (env_path[0] == 'c') ? 2 : 0;
    int cl = 0;
    if (path != 1) {
        if ((B + 3) / 4 <= cluster_slice_cap<4>(device)) cl = 4;
        else if ((B + 7) / 8 <= cluster_slice_cap<8>(device)) cl = 8;
        if (cl == 0 && path == 2) return fail_with(SEPAIHRD_ERR_UNSUPPORTED, "a column of this many draws does not fit the shared memory of 8 thread blocks");
    }
    static const int env_grid = std::getenv("SEPAIHRD_PPC_GRID") ? std::atoi(std::getenv("SEPAIHRD_PPC_GRID")) : 0;
    static const int env_threads = std::getenv("SEPAIHRD_PPC_THREADS") ? std::atoi(std::getenv("SEPAIHRD_PPC_THREADS")) : 0;
    for (int q0 = 0; q0 < n_probs; q0 += SEL_MAXT / 2) {
        const int qg = std::min(SEL_MAXT / 2, n_probs - q0);
        if (cl != 0) {
            const cudaError_t e = (cl == 4) ? launch_cluster_select<4>(s, device, B, all_cols, qg, d_series, d_probs + q0, d_tvals, d_colcnt)
                                            : launch_cluster_select<8>(s, device, B, all_cols, qg, d_series, d_probs + q0, d_tvals, d_colcnt);
            if (e != cudaSuccess) return fail_with(SEPAIHRD_ERR_CUDA, cudaGetErrorString(e));
            const long long work = all_cols * qg;
            ppc_finalize_kernel<<<(unsigned)((work + 255) / 256), 256, 0, s>>>(d_tvals, d_colcnt, all_cols, qg, d_probs + q0, d_q, n_probs, q0);
            if (cudaGetLastError() != cudaSuccess) return fail_with(SEPAIHRD_ERR_CUDA, "ppc_finalize_kernel launch failed");
            sepaihrd_internal::count_launches(ctx, 2);
        } else {
            // Columns in flight x column size is what the later sweeps of a column find in L2 (126 MB)
            const int threads = env_threads ? env_threads : PPC_DEFAULT_THREADS;
            const unsigned grid = (unsigned)std::min<long long>(all_cols, env_grid ? env_grid : PPC_DEFAULT_GRID);
            if (threads == 1024) ppc_select_kernel<1024><<<grid, 1024, 0, s>>>(d_series, B, all_cols, qg, d_probs, d_q, n_probs, q0);
            else if (threads == 256) ppc_select_kernel<256><<<grid, 256, 0, s>>>(d_series, B, all_cols, qg, d_probs, d_q, n_probs, q0);
            else ppc_select_kernel<512><<<grid, 512, 0, s>>>(d_series, B, all_cols, qg, d_probs, d_q, n_probs, q0);
            const cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return fail_with(SEPAIHRD_ERR_CUDA, cudaGetErrorString(e));
            sepaihrd_internal::count_launches(ctx, 1);
        }
    }
    return SEPAIHRD_OK;
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
    double* d_tvals = (double*)buf(sizeof(double) * CS_MAXT * 6 * (size_t)n_cols);      // order statistics per column (cluster select)
    long long* d_colcnt = (long long*)buf(sizeof(long long) * 6 * (size_t)n_cols);
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
        const sepaihrd_rc rc = select_quantiles(ctx, s, d.device, d_series, B, 6 * n_cols, n_probs, d_probs, d_q, d_tvals, d_colcnt, 0);
        if (rc != SEPAIHRD_OK) { cleanup(); return rc; }
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

extern "C" sepaihrd_rc sepaihrd_column_quantiles_device(sepaihrd_ctx* ctx, const double* d_columns, int64_t B, int64_t n_cols, int32_t n_probs,
                                                        const double* probs, double* d_out, int32_t path) {
    using namespace sepaihrd_internal;
    if (!ctx || !d_columns || !probs || !d_out) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "null argument");
    if (B <= 0 || n_cols <= 0) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "empty columns");
    if (n_probs < 1 || n_probs > 64) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "between 1 and 64 quantile probabilities");
    for (int q = 0; q < n_probs; ++q)
        if (!(probs[q] >= 0.0 && probs[q] <= 1.0)) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "quantile probabilities must lie in [0, 1]");
    if (path < 0 || path > 2) return fail_with(SEPAIHRD_ERR_INVALID_ARGUMENT, "path: 0 by size, 1 one block per column, 2 cluster");
    const Dims d = dims(ctx);
    const auto ctx_lock = lock(ctx);
    cudaSetDevice(d.device);
    cudaStream_t s = stream(ctx);
    // scratch slots 9..11: the aggregation pass above uses 0..8 and may be called with the same ctx in between
    double* d_probs = (double*)scratch(ctx, 9, sizeof(double) * 64);
    double* d_tvals = (double*)scratch(ctx, 10, sizeof(double) * CS_MAXT * (size_t)n_cols);
    long long* d_colcnt = (long long*)scratch(ctx, 11, sizeof(long long) * (size_t)n_cols);
    if (!d_probs || !d_tvals || !d_colcnt) return fail_with(SEPAIHRD_ERR_OUT_OF_MEMORY, "work buffers of the column quantiles do not fit");
    auto cleanup = [] {};
    PPC_TRY(cudaMemcpyAsync(d_probs, probs, sizeof(double) * n_probs, cudaMemcpyHostToDevice, s));
    PPC_TRY(cudaStreamSynchronize(s));                                   // probs may be a temporary of the caller
    return select_quantiles(ctx, s, d.device, d_columns, B, n_cols, n_probs, d_probs, d_out, d_tvals, d_colcnt, path);
}
