// Shared internals of libpmgx: context, error handling, launch helpers.
// Replaces src/util.hpp (err_check / device_synchronize) of the reference: errors are
// returned through the C ABI instead of printf + exit(1) (src/util.hpp:21-29).
#pragma once

#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <algorithm>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/pmgx.h"

namespace pmgx
{
void set_error(const char* fmt, ...);

struct Error
{
  int code;
};

#define PMGX_CUDA(call)                                                                            \
  do                                                                                               \
  {                                                                                                \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess)                                                                         \
    {                                                                                              \
      pmgx::set_error("%s:%d CUDA error: %s (%s)", __FILE__, __LINE__, cudaGetErrorString(_e),     \
                      #call);                                                                      \
      throw pmgx::Error{PMGX_ERR_CUDA};                                                            \
    }                                                                                              \
  } while (0)

#define PMGX_NCCL(call)                                                                            \
  do                                                                                               \
  {                                                                                                \
    ncclResult_t _e = (call);                                                                      \
    if (_e != ncclSuccess)                                                                         \
    {                                                                                              \
      pmgx::set_error("%s:%d NCCL error: %s (%s)", __FILE__, __LINE__, ncclGetErrorString(_e),     \
                      #call);                                                                      \
      throw pmgx::Error{PMGX_ERR_NCCL};                                                            \
    }                                                                                              \
  } while (0)

#define PMGX_REQUIRE(cond, ...)                                                                    \
  do                                                                                               \
  {                                                                                                \
    if (!(cond))                                                                                   \
    {                                                                                              \
      pmgx::set_error(__VA_ARGS__);                                                                \
      throw pmgx::Error{PMGX_ERR_ARG};                                                             \
    }                                                                                              \
  } while (0)

// Wrap the body of every extern "C" entry point.
#define PMGX_API_BEGIN try {
#define PMGX_API_END                                                                               \
  }                                                                                                \
  catch (const pmgx::Error& e) { return e.code; }                                                  \
  catch (const std::exception& e)                                                                  \
  {                                                                                                \
    pmgx::set_error("exception: %s", e.what());                                                    \
    return PMGX_ERR_ARG;                                                                           \
  }                                                                                                \
  return PMGX_OK;

template <typename T>
struct DevBuf
{
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n)
  {
    o.p = nullptr;
    o.n = 0;
  }
  ~DevBuf() { release(); }
  void alloc(size_t count)
  {
    release();
    n = count;
    if (count > 0)
      PMGX_CUDA(cudaMalloc(&p, count * sizeof(T)));
  }
  void release()
  {
    if (p)
      cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void upload(const T* h, size_t count, cudaStream_t s)
  {
    alloc(count);
    if (count > 0)
    {
      PMGX_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
      PMGX_CUDA(cudaStreamSynchronize(s));
    }
  }
};
} // namespace pmgx

// Opaque context: one per GPU/rank.
struct pmgx_ctx
{
  int device = 0;
  int rank = 0;
  int nranks = 1;
  int num_sms = 148;
  cudaStream_t stream = nullptr;      // compute stream
  cudaStream_t comm_stream = nullptr; // halo pack / NCCL / unpack
  ncclComm_t comm = nullptr;
  // scratch for reductions: device partials + pinned host result
  double* d_scalars = nullptr;        // [64] device scalars (dot results, alpha, beta, ...)
  double* h_scalars = nullptr;        // pinned, device-mapped mirror (written by k_publish: no copy engine)
  double* h_scalars_dev = nullptr;    // device view of h_scalars
  unsigned int* d_counter = nullptr;  // last-block-done counters
  double* d_partials = nullptr;       // [max_blocks * 4]
  int max_red_blocks = 0;
  // NVLink peer-memory path (one process per GPU on one node, p2p.cu): every rank maps the
  // others' exchange buffers with CUDA IPC; halo values and all-reduce operands are stored
  // straight into the peer's memory by our own kernels, completion is signalled with epoch flags.
  bool p2p = false;
  double* ar_local = nullptr;             // [2][nranks][4] operands + nranks uint64 epoch flags
  double** d_ar_peers = nullptr;          // device array [nranks]: every rank's ar_local, mapped here
  std::vector<void*> p2p_mapped;          // IPC mappings to close at destroy
  unsigned long long* d_ar_epoch = nullptr; // device counter of completed all-reduces
  // wall-time bound of the in-kernel waits on peer flags (0 = unbounded); PMGX_P2P_TIMEOUT_S, default 600 s
  unsigned long long p2p_timeout_ns = 600ull * 1000000000ull;
  long long launches = 0;
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof[PMGX_MAX_DEGREE + 1];
};

namespace pmgx
{
inline void count_launch(pmgx_ctx* ctx, int n = 1) { ctx->launches += n; }
inline void check_launch(const char* what)
{
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
  {
    set_error("kernel launch failed (%s): %s", what, cudaGetErrorString(e));
    throw Error{PMGX_ERR_CUDA};
  }
}

// Host GLL tables (gll.cpp)
void gll_points_weights(int n, std::vector<double>& x, std::vector<double>& w);
void gll_deriv_matrix(const std::vector<double>& x, std::vector<double>& D); // D[q*n+i]
void gll_interp_matrix(int pc, int pf, std::vector<double>& M);              // M[f*(pc+1)+c]

// NVLink peer-memory plumbing (p2p.cu)
namespace p2p
{
// collective over all ranks: every rank contributes `bytes`, all[r*bytes ...] is rank r's block
void allgather_bytes(pmgx_ctx* c, const void* mine, size_t bytes, std::vector<char>& all);
// collective: true iff `mine` is true on every rank
bool all_agree(pmgx_ctx* c, bool mine);
void ctx_setup(pmgx_ctx* c);    // maps the all-reduce slots of all ranks; leaves c->p2p false on failure
void ctx_teardown(pmgx_ctx* c);
// sum / max of d_scalars[slot .. slot+count) over ranks, in place, on the compute stream
void allreduce(pmgx_ctx* c, int slot, int count, bool is_max);
} // namespace p2p

// vector kernels (vector_ops.cu) -- device-scalar flavoured internals used by the solvers
namespace vec
{
void set(pmgx_ctx* c, double* x, long long n, double v);
void copy(pmgx_ctx* c, double* a, const double* b, long long n);
void axpy(pmgx_ctx* c, double* r, double alpha, const double* x, const double* y, long long n);
void scale(pmgx_ctx* c, double* r, double alpha, long long n);
void pointwise_mult(pmgx_ctx* c, double* w, const double* x, const double* y, long long n);
void mask_bc(pmgx_ctx* c, double* b, const int8_t* bc, long long n);
void allreduce_scalars(pmgx_ctx* c, int slot, int count, bool is_max);
// local dot into device scalar slot (no host sync); allreduce over ranks when nranks > 1
void dot_device(pmgx_ctx* c, const double* a, const double* b, long long n, int slot);
double read_scalar(pmgx_ctx* c, int slot); // blocking read of one scalar
// d_scalars[slot .. slot+count) -> h_scalars (same indices), then wait for the compute stream.  A
// kernel stores into mapped host memory: a cudaMemcpy of a few bytes would queue behind bulk
// transfers on the copy engines (measured: +9 ms per V-cycle under concurrent PCIe traffic).
void publish_scalars(pmgx_ctx* c, int slot, int count);
double dot(pmgx_ctx* c, const double* a, const double* b, long long n);
double norm_linf(pmgx_ctx* c, const double* a, long long n);
} // namespace vec
} // namespace pmgx
