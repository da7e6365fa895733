// NVLink peer-memory plumbing for one-process-per-GPU runs on a single NVSwitch node.
//
// The reference hands device pointers to GPU-aware MPI (src/vector.hpp:203-215) and blocks the
// host on every exchange; NCCL send/recv removes the host from the path but still costs a
// 20-30 us proxy round trip per call, which is what the coarse levels of the V-cycle are made of
// (60 SpMV halos + 60 two-double all-reduces per cycle).  Here every rank maps its peers'
// exchange buffers once with CUDA IPC, and our own kernels store halo values / all-reduce
// operands straight into the peer's HBM over NVLink, followed by a release store of an epoch
// flag; the receiver spins on its local flag with acquire loads.  NCCL remains the bootstrap
// (exchange of IPC handles) and the fallback when IPC is unavailable (PMGX_P2P=0 forces it).
#include "common.hpp"
#include "operator.hpp"

#include <cstdlib>
#include <cstring>

namespace pmgx
{
namespace p2p
{
namespace
{
constexpr int AR_MAX = 4; // operands per all-reduce

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// One CTA.  Thread r < nranks stores this rank's operands into rank r's slot [buf][myrank] and
// then releases flag[myrank] = epoch there; it then waits for rank r's operands to arrive here.
// Thread k < count finally combines the slots in rank order, so every rank gets the same bits.
__global__ void k_allreduce_p2p(double* const* __restrict__ peers, double* __restrict__ vals, int count,
                                int myrank, int nranks, unsigned long long epoch, bool is_max)
{
  const int r = threadIdx.x;
  const int buf = (int)(epoch & 1ull);
  const size_t flag_off = (size_t)2 * nranks * AR_MAX; // in doubles (flags are 8 bytes too)
  if (r < nranks)
  {
    double* dst = peers[r] + ((size_t)buf * nranks + myrank) * AR_MAX;
    for (int k = 0; k < count; ++k)
      dst[k] = vals[k];
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned long long*>(peers[r] + flag_off) + myrank, epoch);
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(peers[myrank] + flag_off) + r;
    while (ld_acquire_sys(mine) < epoch)
      __nanosleep(40);
  }
  __syncthreads();
  if (r < count)
  {
    const double* src = peers[myrank] + (size_t)buf * nranks * AR_MAX;
    double acc = __ldcg(src + r);
    for (int q = 1; q < nranks; ++q)
    {
      const double v = __ldcg(src + (size_t)q * AR_MAX + r);
      acc = is_max ? fmax(acc, v) : acc + v;
    }
    vals[r] = acc;
  }
}
} // namespace

void allgather_bytes(pmgx_ctx* c, const void* mine, size_t bytes, std::vector<char>& all)
{
  all.assign(bytes * c->nranks, 0);
  if (c->nranks == 1)
  {
    std::memcpy(all.data(), mine, bytes);
    return;
  }
  DevBuf<char> send, recv;
  send.alloc(bytes);
  recv.alloc(bytes * c->nranks);
  PMGX_CUDA(cudaMemcpyAsync(send.p, mine, bytes, cudaMemcpyHostToDevice, c->stream));
  PMGX_NCCL(ncclAllGather(send.p, recv.p, bytes, ncclChar, c->comm, c->stream));
  PMGX_CUDA(cudaMemcpyAsync(all.data(), recv.p, bytes * c->nranks, cudaMemcpyDeviceToHost, c->stream));
  PMGX_CUDA(cudaStreamSynchronize(c->stream));
}

bool all_agree(pmgx_ctx* c, bool mine)
{
  char m = mine ? 1 : 0;
  std::vector<char> all;
  allgather_bytes(c, &m, 1, all);
  for (char v : all)
    if (!v)
      return false;
  return true;
}

void ctx_setup(pmgx_ctx* c)
{
  c->p2p = false;
  if (c->nranks == 1)
    return;
  const char* env = getenv("PMGX_P2P");
  const bool want = !(env && std::strcmp(env, "0") == 0) && c->nranks <= 32;
  const size_t bytes = ((size_t)2 * c->nranks * AR_MAX + c->nranks) * sizeof(double);
  bool ok = want;
  cudaIpcMemHandle_t h;
  std::memset(&h, 0, sizeof(h));
  if (ok)
  {
    ok = cudaMalloc(&c->ar_local, bytes) == cudaSuccess && cudaMemset(c->ar_local, 0, bytes) == cudaSuccess
         && cudaIpcGetMemHandle(&h, c->ar_local) == cudaSuccess;
    cudaGetLastError();
  }
  std::vector<char> all;
  allgather_bytes(c, &h, sizeof(h), all);
  ok = all_agree(c, ok);
  std::vector<double*> peers(c->nranks, nullptr);
  if (ok)
  {
    for (int r = 0; r < c->nranks && ok; ++r)
    {
      if (r == c->rank)
      {
        peers[r] = c->ar_local;
        continue;
      }
      cudaIpcMemHandle_t hr;
      std::memcpy(&hr, all.data() + (size_t)r * sizeof(hr), sizeof(hr));
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
      {
        cudaGetLastError();
        ok = false;
        break;
      }
      c->p2p_mapped.push_back(p);
      peers[r] = static_cast<double*>(p);
    }
  }
  ok = all_agree(c, ok);
  if (!ok)
  {
    ctx_teardown(c);
    return;
  }
  PMGX_CUDA(cudaMalloc(&c->d_ar_peers, c->nranks * sizeof(double*)));
  PMGX_CUDA(cudaMemcpy(c->d_ar_peers, peers.data(), c->nranks * sizeof(double*), cudaMemcpyHostToDevice));
  c->p2p = true;
}

void ctx_teardown(pmgx_ctx* c)
{
  for (void* p : c->p2p_mapped)
    cudaIpcCloseMemHandle(p);
  c->p2p_mapped.clear();
  if (c->ar_local)
    cudaFree(c->ar_local);
  if (c->d_ar_peers)
    cudaFree(c->d_ar_peers);
  c->ar_local = nullptr;
  c->d_ar_peers = nullptr;
  c->p2p = false;
  cudaGetLastError();
}

void allreduce(pmgx_ctx* c, int slot, int count, bool is_max)
{
  if (c->nranks == 1)
    return;
  if (!c->p2p || count > AR_MAX)
  {
    PMGX_NCCL(ncclAllReduce(c->d_scalars + slot, c->d_scalars + slot, count, ncclDouble, is_max ? ncclMax : ncclSum,
                            c->comm, c->stream));
    return;
  }
  ++c->ar_epoch;
  k_allreduce_p2p<<<1, 32, 0, c->stream>>>(c->d_ar_peers, c->d_scalars + slot, count, c->rank, c->nranks, c->ar_epoch,
                                           is_max);
  check_launch("k_allreduce_p2p");
  count_launch(c);
}
} // namespace p2p
} // namespace pmgx
