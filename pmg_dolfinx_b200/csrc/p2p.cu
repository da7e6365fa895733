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
#include "reduce.cuh"

#include <cstdlib>
#include <cstring>

namespace pmgx
{
namespace p2p
{
namespace
{
// Stand-alone all-reduce of device scalars (for reductions whose kernel does not carry the
// peer epilogue): one CTA running peer_allreduce.
__global__ void k_allreduce_p2p(PeerReduce pr, double* __restrict__ vals, int count, bool is_max)
{
  __shared__ double in[AR_MAX];
  if ((int)threadIdx.x < count)
    in[threadIdx.x] = vals[threadIdx.x];
  __syncthreads();
  if (is_max)
    peer_allreduce<true>(pr, in, count, vals);
  else
    peer_allreduce<false>(pr, in, count, vals);
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
  if (const char* t = getenv("PMGX_P2P_TIMEOUT_S"))
    c->p2p_timeout_ns = (unsigned long long)(std::max(atof(t), 0.0) * 1e9);
  const char* env = getenv("PMGX_P2P");
  const bool want = !(env && std::strcmp(env, "0") == 0) && c->nranks <= 32;
  const size_t bytes = ((size_t)2 * c->nranks * AR_MAX + c->nranks) * sizeof(double);
  bool ok = want;
  cudaIpcMemHandle_t h;
  std::memset(&h, 0, sizeof(h));
  if (ok)
  {
    ok = cudaMalloc(&c->ar_local, bytes) == cudaSuccess && cudaMemsetAsync(c->ar_local, 0, bytes, c->stream) == cudaSuccess
         && cudaStreamSynchronize(c->stream) == cudaSuccess
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
  PMGX_CUDA(cudaMalloc(&c->d_ar_epoch, sizeof(unsigned long long)));
  PMGX_CUDA(cudaMemsetAsync(c->d_ar_epoch, 0, sizeof(unsigned long long), c->stream));
  PMGX_CUDA(cudaStreamSynchronize(c->stream));
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
  if (c->d_ar_epoch)
    cudaFree(c->d_ar_epoch);
  c->d_ar_epoch = nullptr;
  c->ar_local = nullptr;
  c->d_ar_peers = nullptr;
  c->p2p = false;
  cudaGetLastError();
}

// descriptor of an all-reduce on this context's compute stream (the epoch lives on the device);
// nranks == 0 when the peer path is off
PeerReduce next_epoch(pmgx_ctx* c)
{
  PeerReduce pr;
  if (c->p2p)
  {
    pr.peers = c->d_ar_peers;
    pr.myrank = c->rank;
    pr.nranks = c->nranks;
    pr.epoch_ptr = c->d_ar_epoch;
    pr.timeout_ns = c->p2p_timeout_ns;
  }
  return pr;
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
  k_allreduce_p2p<<<1, 32, 0, c->stream>>>(next_epoch(c), c->d_scalars + slot, count, is_max);
  check_launch("k_allreduce_p2p");
  count_launch(c);
}
} // namespace p2p
} // namespace pmgx
