// Warp-shuffle block reduction + deterministic last-block finish (north_star item 3:
// "the CG dot-product/axpy path uses warp-level reductions").  Replaces
// thrust::inner_product + host return (src/vector.hpp:345-347).
#pragma once
#include <cuda_runtime.h>

struct pmgx_ctx;

namespace pmgx
{
constexpr int RED_THREADS = 256;
constexpr int AR_MAX = 4; // operands per peer-memory all-reduce slot (p2p.cu)

// Cross-GPU epilogue of a grid reduction over NVLink peer memory (p2p.cu): when nranks > 1 the
// last block does not stop at the local sums but stores them into every peer's slot, releases an
// epoch flag, waits for the peers' flags and combines the slots in rank order -- the reduction
// kernel IS the all-reduce, no second launch.  nranks == 0: local reduction only.
struct PeerReduce
{
  double* const* peers = nullptr; // every rank's slot buffer, mapped into this process
  int myrank = 0, nranks = 0;
  // device counter of completed all-reduces: the kernel takes *epoch_ptr + 1 and stores it back,
  // so a launch carries no per-call state and can be replayed from a CUDA graph
  unsigned long long* epoch_ptr = nullptr;
  unsigned long long timeout_ns = 0; // 0: wait for ever (like NCCL); else trap after this long (wait_epoch)
};

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *p >= epoch.  Benign rank skew (a peer busy with host work, a first-use graph
// instantiation, a debugger) must not kill the context, so the wait is bounded in WALL time
// (%globaltimer), not in polls: timeout_ns = 0 waits for ever like NCCL would; otherwise a peer
// that has not arrived after timeout_ns (default 10 min, PMGX_P2P_TIMEOUT_S) is taken for dead
// and the kernel traps instead of hanging the GPU.
__device__ __forceinline__ void wait_epoch(const unsigned long long* p, unsigned long long epoch,
                                           unsigned long long timeout_ns)
{
  if (ld_acquire_sys_u64(p) >= epoch)
    return;
  const unsigned long long t0 = globaltimer_ns();
  unsigned int spins = 0;
  while (ld_acquire_sys_u64(p) < epoch)
  {
    __nanosleep(spins < 64 ? 32 : 256);
    if ((++spins & 1023u) == 0 && timeout_ns != 0 && globaltimer_ns() - t0 > timeout_ns)
      __trap();
  }
}

// One CTA: vals[0..count) (shared or global, visible to the CTA) are combined over all ranks and
// written to out[0..count).  Thread r < nranks talks to rank r.
template <bool MAX>
__device__ __forceinline__ void peer_allreduce(const PeerReduce& pr, const double* vals, int count, double* out)
{
  const int r = threadIdx.x;
  const unsigned long long epoch = *pr.epoch_ptr + 1; // read by every thread before anyone stores it back
  const int buf = (int)(epoch & 1ull);
  const size_t flag_off = (size_t)2 * pr.nranks * AR_MAX; // in 8-byte units
  if (r < pr.nranks)
  {
    double* dst = pr.peers[r] + ((size_t)buf * pr.nranks + pr.myrank) * AR_MAX;
    for (int k = 0; k < count; ++k)
      dst[k] = vals[k];
    __threadfence_system();
    st_release_sys_u64(reinterpret_cast<unsigned long long*>(pr.peers[r] + flag_off) + pr.myrank, epoch);
    wait_epoch(reinterpret_cast<const unsigned long long*>(pr.peers[pr.myrank] + flag_off) + r, epoch, pr.timeout_ns);
  }
  __syncthreads();
  if (r < count)
  {
    const double* src = pr.peers[pr.myrank] + (size_t)buf * pr.nranks * AR_MAX;
    double acc = __ldcg(src + r);
    for (int q = 1; q < pr.nranks; ++q)
    {
      const double v = __ldcg(src + (size_t)q * AR_MAX + r);
      acc = MAX ? fmax(acc, v) : acc + v;
    }
    out[r] = acc;
  }
  if (r == 0)
    *pr.epoch_ptr = epoch; // after the barrier above: every thread of the CTA has read the old value
}

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
                                            double* __restrict__ out, const PeerReduce pr = PeerReduce())
{
  __shared__ double sh[NV][RED_THREADS / 32];
  __shared__ double res[NV];
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
      {
        if (pr.nranks > 1)
          res[k] = w;
        else
          out[k] = w;
      }
    }
    if (lane == 0)
      *counter = 0u;
  }
  if (pr.nranks > 1)
  {
    __syncthreads();
    peer_allreduce<MAX>(pr, res, NV, out);
  }
}
namespace p2p
{
// descriptor of an all-reduce on the context's compute stream;
// nranks == 0 when the peer-memory path is off (single rank or NCCL fallback)
PeerReduce next_epoch(pmgx_ctx* c);
} // namespace p2p
} // namespace pmgx
