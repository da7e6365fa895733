// Context: device selection, compute + comm streams, NCCL communicator, reduction scratch.
// Replaces the reference's implicit "one MPI rank per GPU, stream 0, MPI_COMM_WORLD"
// (examples/pmg/select_gpu.sh, src/vector.hpp:350) by an explicit per-GPU context.
#include "common.hpp"

#include <cstdlib>
#include <cstring>

extern "C"
{
int pmgx_nccl_unique_id(void* id_h)
{
  PMGX_API_BEGIN
  static_assert(sizeof(ncclUniqueId) <= PMGX_NCCL_ID_BYTES, "ncclUniqueId too large");
  PMGX_REQUIRE(id_h != nullptr, "nccl_unique_id: null output");
  ncclUniqueId id;
  PMGX_NCCL(ncclGetUniqueId(&id));
  std::memset(id_h, 0, PMGX_NCCL_ID_BYTES);
  std::memcpy(id_h, &id, sizeof(id));
  PMGX_API_END
}

int pmgx_ctx_create(int device, int rank, int nranks, const void* nccl_id_h, pmgx_ctx** out)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(out != nullptr, "ctx_create: null output");
  PMGX_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "ctx_create: bad rank %d/%d", rank, nranks);
  PMGX_REQUIRE(nranks == 1 || nccl_id_h != nullptr, "ctx_create: nranks > 1 needs an NCCL id");
  int ndev = 0;
  PMGX_CUDA(cudaGetDeviceCount(&ndev));
  PMGX_REQUIRE(device >= 0 && device < ndev, "ctx_create: device %d not available (%d devices)", device, ndev);
  PMGX_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PMGX_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
  {
    pmgx::set_error("pmgx is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    return PMGX_ERR_UNSUPPORTED;
  }
  pmgx_ctx* c = new pmgx_ctx();
  c->device = device;
  c->rank = rank;
  c->nranks = nranks;
  c->num_sms = prop.multiProcessorCount;
  PMGX_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {
    // the halo stream outranks the compute stream: its small pack / unpack kernels must not queue
    // behind the CTAs of a long interior-cell kernel
    int lo = 0, hi = 0;
    PMGX_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    PMGX_CUDA(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
  }
  PMGX_CUDA(cudaMalloc(&c->d_scalars, 64 * sizeof(double)));
  PMGX_CUDA(cudaMemsetAsync(c->d_scalars, 0, 64 * sizeof(double), c->stream));
  PMGX_CUDA(cudaHostAlloc(&c->h_scalars, 64 * sizeof(double), cudaHostAllocMapped));
  PMGX_CUDA(cudaHostGetDevicePointer(&c->h_scalars_dev, c->h_scalars, 0));
  c->max_red_blocks = 8 * c->num_sms;
  PMGX_CUDA(cudaMalloc(&c->d_partials, (size_t)c->max_red_blocks * 4 * sizeof(double)));
  PMGX_CUDA(cudaMalloc(&c->d_counter, 16 * sizeof(unsigned int)));
  PMGX_CUDA(cudaMemsetAsync(c->d_counter, 0, 16 * sizeof(unsigned int), c->stream));
  // the zero fills must have landed before any kernel on a (non-blocking) stream reads them
  PMGX_CUDA(cudaStreamSynchronize(c->stream));
  if (nranks > 1)
  {
    ncclUniqueId id;
    std::memcpy(&id, nccl_id_h, sizeof(id));
    PMGX_NCCL(ncclCommInitRank(&c->comm, nranks, id, rank));
    pmgx::p2p::ctx_setup(c); // NVLink peer-memory path; NCCL stays the fallback
  }
  *out = c;
  PMGX_API_END
}

int pmgx_ctx_destroy(pmgx_ctx* c)
{
  PMGX_API_BEGIN
  if (!c)
    return PMGX_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  pmgx::p2p::ctx_teardown(c);
  if (c->comm)
    ncclCommDestroy(c->comm);
  cudaFree(c->d_scalars);
  cudaFreeHost(c->h_scalars);
  cudaFree(c->d_partials);
  cudaFree(c->d_counter);
  cudaStreamDestroy(c->stream);
  cudaStreamDestroy(c->comm_stream);
  delete c;
  PMGX_API_END
}

int pmgx_ctx_sync(pmgx_ctx* c)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(c, "ctx_sync: null ctx");
  PMGX_CUDA(cudaStreamSynchronize(c->comm_stream));
  PMGX_CUDA(cudaStreamSynchronize(c->stream));
  PMGX_API_END
}

int pmgx_ctx_profile(pmgx_ctx* c, int on)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(c, "ctx_profile: null ctx");
  c->profiling = on != 0;
  PMGX_API_END
}

int pmgx_ctx_profile_read(pmgx_ctx* c, int degree, double* ms_total_h, long long* launches_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(c && degree >= 1 && degree <= PMGX_MAX_DEGREE, "ctx_profile_read: bad arguments");
  PMGX_CUDA(cudaSetDevice(c->device));
  PMGX_CUDA(cudaStreamSynchronize(c->stream));
  double total = 0.0;
  for (auto& pr : c->prof[degree])
  {
    float ms = 0.f;
    PMGX_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
    total += ms;
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  if (ms_total_h)
    *ms_total_h = total;
  if (launches_h)
    *launches_h = (long long)c->prof[degree].size();
  c->prof[degree].clear();
  PMGX_API_END
}

void* pmgx_ctx_stream(pmgx_ctx* c) { return c ? (void*)c->stream : nullptr; }
int pmgx_ctx_rank(pmgx_ctx* c) { return c ? c->rank : -1; }
int pmgx_ctx_uses_p2p(pmgx_ctx* c) { return c && c->p2p ? 1 : 0; }
int pmgx_ctx_nranks(pmgx_ctx* c) { return c ? c->nranks : -1; }
long long pmgx_ctx_launch_count(pmgx_ctx* c) { return c ? c->launches : -1; }
}
