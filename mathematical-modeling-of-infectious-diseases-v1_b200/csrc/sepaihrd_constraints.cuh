// sepaihrd_constraints.cuh -- SEPAIHRDParameterManager::applyConstraints on the device (clamp / mirror reflection), shared by
// the evaluation kernel's prologue and the device-resident samplers.
//   reflectBound      reference src/model/parameters/SEPAIHRDParameterManager.cpp:302-313
//   applyConstraints  reference src/model/parameters/SEPAIHRDParameterManager.cpp:315-347
#pragma once

#include <cuda_runtime.h>

namespace sepaihrd {

// std::max(a, b) == (a < b) ? b : a   (NaN in b is ignored, NaN in a is returned)
__device__ __forceinline__ double std_max(double a, double b) { return (a < b) ? b : a; }
__device__ __forceinline__ double std_min(double a, double b) { return (b < a) ? b : a; }

// ---- constraints ------------------------------------------------------------------------------------
// reflectBound, SEPAIHRDParameterManager.cpp:302-313
__device__ __forceinline__ double reflect_bound(double value, double minb, double maxb) {
    if (minb >= maxb) return minb;
    const double width = maxb - minb;
    double y = fmod(value - minb, 2.0 * width);
    if (y < 0) y += 2.0 * width;
    if (y <= width) return minb + y;
    return maxb - (y - width);
}
// applyConstraints, .cpp:315-347
__device__ __forceinline__ double constrain(double v, double lo, double hi, int mode) {
    if (lo == lo) {   // has a bounds entry
        if (lo > hi) { double t = lo; lo = hi; hi = t; }
        return (mode == 0) ? std_min(std_max(v, lo), hi) : reflect_bound(v, lo, hi);
    }
    return (mode == 0) ? std_max(0.0, v) : fabs(v);
}

}  // namespace sepaihrd
