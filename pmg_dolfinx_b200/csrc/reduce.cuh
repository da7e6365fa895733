// Warp-shuffle block reduction + deterministic last-block finish (north_star item 3:
// "the CG dot-product/axpy path uses warp-level reductions").  Replaces
// thrust::inner_product + host return (src/vector.hpp:345-347).
#pragma once
#include <cuda_runtime.h>

namespace pmgx
{
constexpr int RED_THREADS = 256;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Reduce NV values per thread across the block, then across the grid: every block writes its
// partial, the last block to finish (atomic ticket) sums the partials in a fixed order, so the
// result does not depend on scheduling.  out[k] receives value k.  Must be called by all
// threads of a RED_THREADS-sized block.
template <int NV, bool MAX = false>
__device__ __forceinline__ void grid_reduce(double (&v)[NV], double* __restrict__ partials,
                                            unsigned int* __restrict__ counter,
                                            double* __restrict__ out)
{
  __shared__ double sh[NV][RED_THREADS / 32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k)
  {
    double w = MAX ? warp_max(v[k]) : warp_sum(v[k]);
    if (lane == 0)
      sh[k][warp] = w;
  }
  __syncthreads();
  if (warp == 0)
  {
#pragma unroll
    for (int k = 0; k < NV; ++k)
    {
      double w = lane < RED_THREADS / 32 ? sh[k][lane] : (MAX ? -1.0 : 0.0);
      w = MAX ? warp_max(w) : warp_sum(w);
      if (lane == 0)
        partials[(size_t)blockIdx.x * NV + k] = w;
    }
  }
  if (threadIdx.x == 0)
  {
    __threadfence();
    unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last)
    return;
  __threadfence();
  // the last block sums the partials in a fixed (thread, warp) order with all of its threads
#pragma unroll
  for (int k = 0; k < NV; ++k)
  {
    double acc = MAX ? -1.0 : 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += RED_THREADS)
    {
      double pv = __ldcg(&partials[(size_t)b * NV + k]);
      acc = MAX ? fmax(acc, pv) : acc + pv;
    }
    acc = MAX ? warp_max(acc) : warp_sum(acc);
    if (lane == 0)
      sh[k][warp] = acc;
  }
  __syncthreads();
  if (warp == 0)
  {
#pragma unroll
    for (int k = 0; k < NV; ++k)
    {
      double w = lane < RED_THREADS / 32 ? sh[k][lane] : (MAX ? -1.0 : 0.0);
      w = MAX ? warp_max(w) : warp_sum(w);
      if (lane == 0)
        out[k] = w;
    }
    if (lane == 0)
      *counter = 0u;
  }
}
} // namespace pmgx
