// sepaihrd_internal.h -- the few ctx accessors other translation units of the library need (not part of the ABI).
#pragma once

#include <mutex>

#include <cuda_runtime.h>

#include "sepaihrd_b200.h"

namespace sepaihrd_internal {

struct Dims { int n, K, runup_offset, n_nonneg, P, device, n_user; };   // n: age classes the kernels run with (n_user zero-padded to 4 or 16); n_nonneg: output times >= 0 (the last ones)
Dims dims(const sepaihrd_ctx* ctx);
cudaStream_t stream(const sepaihrd_ctx* ctx);
cudaStream_t copy_stream(const sepaihrd_ctx* ctx);     // the ctx's H2D / D2H stream next to its compute stream
cudaEvent_t* copy_events(sepaihrd_ctx* ctx);           // [8] timing-free events for chunked copies (callers hold the ctx lock)
cudaEvent_t* chunk_events(sepaihrd_ctx* ctx);          // [8]
sepaihrd_rc fail_with(sepaihrd_rc rc, const char* msg);
const double* lower_bounds(const sepaihrd_ctx* ctx);   // host copies, [P], as given at creation
const double* upper_bounds(const sepaihrd_ctx* ctx);
void count_launches(sepaihrd_ctx* ctx, int n);
int constraint_mode(const sepaihrd_ctx* ctx);         // 0 clamp, 1 reflect (sepaihrd_set_constraint_mode)
int num_sms(const sepaihrd_ctx* ctx);
// ---- ordering pass (sepaihrd_order.cu) ----------------------------------------------------------------------------------
void** order_slot(sepaihrd_ctx* ctx);                 // where the ctx keeps the fitted model (owned by sepaihrd_order.cu)
int order_mode(const sepaihrd_ctx* ctx);
bool order_applicable(const sepaihrd_ctx* ctx);       // 4 lanes per set, FAST arithmetic, a scorable problem: where the profiling instantiation exists
void set_order_mode(sepaihrd_ctx* ctx, int mode);
sepaihrd_rc order_autofit_host(sepaihrd_ctx* ctx, const double* params, long long B, long long ld);   // host-pointer evaluations: fit when there is no model or the batch looks different
void order_release(sepaihrd_ctx* ctx);
// *perm = the order in which to hand the B sets to the warps (device, owned by the model), or left null (no model, batch too
// small, ordering off); enqueues its kernels on the ctx stream
sepaihrd_rc order_batch(sepaihrd_ctx* ctx, const double* d_params, long long B, long long ld, const int** perm);
sepaihrd_rc order_mark_done(sepaihrd_ctx* ctx);      // call right after the launch that reads *perm has been enqueued (same stream)
// one launch of the likelihood kernel that also records, per set, the attempts made before every grid point: d_profile [B][K]
sepaihrd_rc eval_profile(sepaihrd_ctx* ctx, const double* d_params, long long B, long long ld, double* d_ll, unsigned* d_status, int* d_profile);
// the evaluation as sepaihrd_eval_batch_device does it, but never reordered (callers whose batches change every iteration)
sepaihrd_rc eval_batch_device_unordered(sepaihrd_ctx* ctx, const double* d_params, long long B, long long ld, double* d_ll, unsigned* d_status, int* d_steps);
// the ctx mutex (recursive): every entry point of another translation unit that touches the ctx holds it for its duration
std::unique_lock<std::recursive_mutex> lock(sepaihrd_ctx* ctx);
// Grow-only device work buffer `slot` (0..15) of at least `bytes`, owned by the ctx and reused across calls; nullptr when the
// allocation fails.  sepaihrd_release_scratch() drops them all.
void* scratch(sepaihrd_ctx* ctx, int slot, size_t bytes);
void release_scratch(sepaihrd_ctx* ctx);          // kernels another translation unit enqueued on the ctx stream
// The six posterior-predictive series of B draws straight from the trajectory kernel: out[6][T][n][B] (T = output days t >= 0),
// all draws from the one state d_init (quirk Q9).  Failed draws are NaN in every series.
// One launch covers draws [b0, b0 + nb) of B_total (d_params, d_status and the columns of d_series are indexed by the GLOBAL draw),
// so the caller can feed the draws in chunks while later ones are still on their way from the host.
sepaihrd_rc simulate_ppc_series(sepaihrd_ctx* ctx, const double* d_params, long long b0, long long nb, long long B_total, long long ld,
                                const double* d_init, double* d_series, unsigned* d_status);

}  // namespace sepaihrd_internal
