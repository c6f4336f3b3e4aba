// Device helpers shared by the RANSAC-EPnP kernels.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "epnp_math.cuh"
#include "ransac.cuh"

namespace spe {

constexpr unsigned kFullMask = 0xffffffffu;

// RANSACUpdateNumIters: the iteration budget after a model with outlier ratio ep
SPE_HD __forceinline__ int update_num_iters(double p, double ep, int max_iters) {
  p = fmin(fmax(p, 0.0), 1.0);
  ep = fmin(fmax(ep, 0.0), 1.0);
  double num = fmax(1.0 - p, 2.2250738585072014e-308);
  double denom = 1.0 - pow(1.0 - ep, (double)kModelPoints);
  if (denom < 2.2250738585072014e-308) return 0;
  num = log(num);
  denom = log(denom);
  return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : (int)rint(num / denom);
}

// U(n, limit): how many distinct minimal sets the first `limit` draws for n points contain = number of entries of the
// (ascending, 0xffff-padded) first-occurrence list below `limit`
__device__ __forceinline__ int unique_below(const DevModel& m, int n, int limit) {
  const uint16_t* u = m.uniq + (size_t)(n - 6) * m.max_hyp;
  int lo = 0, hi = m.max_hyp;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((int)u[mid] < limit) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

}  // namespace spe
