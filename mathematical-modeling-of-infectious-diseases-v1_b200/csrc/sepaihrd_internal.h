// sepaihrd_internal.h -- the few ctx accessors other translation units of the library need (not part of the ABI).
#pragma once

#include <mutex>

#include <cuda_runtime.h>

#include "sepaihrd_b200.h"

namespace sepaihrd_internal {

struct Dims { int n, K, runup_offset, n_nonneg, P, device, n_user; };   // n: age classes the kernels run with (n_user zero-padded to 4 or 16); n_nonneg: output times >= 0 (the last ones)
Dims dims(const sepaihrd_ctx* ctx);
cudaStream_t stream(const sepaihrd_ctx* ctx);
sepaihrd_rc fail_with(sepaihrd_rc rc, const char* msg);
const double* lower_bounds(const sepaihrd_ctx* ctx);   // host copies, [P], as given at creation
const double* upper_bounds(const sepaihrd_ctx* ctx);
void count_launches(sepaihrd_ctx* ctx, int n);
int constraint_mode(const sepaihrd_ctx* ctx);         // 0 clamp, 1 reflect (sepaihrd_set_constraint_mode)
// the ctx mutex (recursive): every entry point of another translation unit that touches the ctx holds it for its duration
std::unique_lock<std::recursive_mutex> lock(sepaihrd_ctx* ctx);
// Grow-only device work buffer `slot` (0..15) of at least `bytes`, owned by the ctx and reused across calls; nullptr when the
// allocation fails.  sepaihrd_release_scratch() drops them all.
void* scratch(sepaihrd_ctx* ctx, int slot, size_t bytes);
void release_scratch(sepaihrd_ctx* ctx);          // kernels another translation unit enqueued on the ctx stream
// D, CumH, CumICU of B draws in the DRAW-MINOR layout [K][3n][B] (device pointers); d_init: one shared state or null
sepaihrd_rc simulate_observed_draw_minor(sepaihrd_ctx* ctx, const double* d_params, long long B, long long ld, const double* d_init,
                                         double* d_out, unsigned* d_status);

}  // namespace sepaihrd_internal
