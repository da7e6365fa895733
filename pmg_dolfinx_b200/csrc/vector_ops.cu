// BLAS-1 kernels on raw device arrays: the free functions of dolfinx::acc
// (src/vector.hpp:333-454) without Thrust.  Streaming kernels, 128-bit loads, grid sized to
// a multiple of the SM count; reductions via warp shuffles (reduce.cuh).
#include "common.hpp"
#include "reduce.cuh"

namespace pmgx
{
namespace
{
constexpr int VT = 256; // threads per block for streaming kernels

inline int stream_grid(pmgx_ctx* c, long long n, int per_thread)
{
  long long blocks = (n + (long long)VT * per_thread - 1) / ((long long)VT * per_thread);
  long long cap = (long long)c->num_sms * 16;
  if (blocks > cap)
    blocks = cap;
  if (blocks < 1)
    blocks = 1;
  return (int)blocks;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// OP: 0 set, 1 copy, 2 axpy (r = a*x + y), 3 scale, 4 pointwise mult
template <int OP>
__global__ void __launch_bounds__(VT) k_stream(double* __restrict__ r, const double* x,
                                               const double* y, double a, long long n, bool vec2)
{
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  if (vec2)
  {
    const long long n2 = n >> 1;
    double2* r2 = reinterpret_cast<double2*>(r);
    const double2* x2 = reinterpret_cast<const double2*>(x);
    const double2* y2 = reinterpret_cast<const double2*>(y);
    for (long long i = tid; i < n2; i += nth)
    {
      double2 o;
      if (OP == 0)
        o = make_double2(a, a);
      else if (OP == 1)
        o = x2[i];
      else if (OP == 2)
      {
        double2 xv = x2[i], yv = y2[i];
        o = make_double2(xv.x * a + yv.x, xv.y * a + yv.y);
      }
      else if (OP == 3)
      {
        double2 rv = r2[i];
        o = make_double2(rv.x * a, rv.y * a);
      }
      else
      {
        double2 xv = x2[i], yv = y2[i];
        o = make_double2(xv.x * yv.x, xv.y * yv.y);
      }
      r2[i] = o;
    }
    if (tid == 0 && (n & 1))
    {
      const long long i = n - 1;
      if (OP == 0)
        r[i] = a;
      else if (OP == 1)
        r[i] = x[i];
      else if (OP == 2)
        r[i] = x[i] * a + y[i];
      else if (OP == 3)
        r[i] = r[i] * a;
      else
        r[i] = x[i] * y[i];
    }
  }
  else
  {
    for (long long i = tid; i < n; i += nth)
    {
      if (OP == 0)
        r[i] = a;
      else if (OP == 1)
        r[i] = x[i];
      else if (OP == 2)
        r[i] = x[i] * a + y[i];
      else if (OP == 3)
        r[i] = r[i] * a;
      else
        r[i] = x[i] * y[i];
    }
  }
}

__global__ void __launch_bounds__(VT) k_mask_bc(double* __restrict__ b,
                                                const int8_t* __restrict__ bc, long long n)
{
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth)
    b[i] = b[i] * (1 - bc[i]);
}

__global__ void __launch_bounds__(RED_THREADS) k_dot(const double* __restrict__ a,
                                                     const double* __restrict__ b, long long n,
                                                     bool vec2, double* partials,
                                                     unsigned int* counter, double* out)
{
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  double s0 = 0.0, s1 = 0.0;
  if (vec2)
  {
    const long long n2 = n >> 1;
    const double2* a2 = reinterpret_cast<const double2*>(a);
    const double2* b2 = reinterpret_cast<const double2*>(b);
    long long i = tid;
    for (; i + nth < n2; i += 2 * nth)
    {
      double2 av = a2[i], bv = b2[i], aw = a2[i + nth], bw = b2[i + nth];
      s0 = fma(av.x, bv.x, s0);
      s1 = fma(av.y, bv.y, s1);
      s0 = fma(aw.x, bw.x, s0);
      s1 = fma(aw.y, bw.y, s1);
    }
    for (; i < n2; i += nth)
    {
      double2 av = a2[i], bv = b2[i];
      s0 = fma(av.x, bv.x, s0);
      s1 = fma(av.y, bv.y, s1);
    }
    if (tid == 0 && (n & 1))
      s0 = fma(a[n - 1], b[n - 1], s0);
  }
  else
  {
    for (long long i = tid; i < n; i += nth)
      s0 = fma(a[i], b[i], s0);
  }
  double v[1] = {s0 + s1};
  grid_reduce<1>(v, partials, counter, out);
}

__global__ void __launch_bounds__(RED_THREADS) k_absmax(const double* __restrict__ a, long long n,
                                                        double* partials, unsigned int* counter,
                                                        double* out)
{
  const long long nth = (long long)gridDim.x * blockDim.x;
  double m = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth)
    m = fmax(m, fabs(a[i]));
  double v[1] = {m};
  grid_reduce<1, true>(v, partials, counter, out);
}

template <int OP>
void launch_stream(pmgx_ctx* c, double* r, const double* x, const double* y, double a, long long n)
{
  if (n <= 0)
    return;
  cudaSetDevice(c->device);
  const bool v2 = aligned16(r) && (x == nullptr || aligned16(x)) && (y == nullptr || aligned16(y));
  const int grid = stream_grid(c, v2 ? (n + 1) / 2 : n, 4);
  k_stream<OP><<<grid, VT, 0, c->stream>>>(r, x, y, a, n, v2);
  check_launch("k_stream");
  count_launch(c);
}
} // namespace

namespace vec
{
void set(pmgx_ctx* c, double* x, long long n, double v)
{
  if (n <= 0)
    return;
  // a fill KERNEL, never cudaMemsetAsync: memsets may be executed by a copy engine and then queue
  // behind bulk host<->device transfers of other streams (measured: V-cycle +6 ms under PCIe load)
  launch_stream<0>(c, x, nullptr, nullptr, v, n);
}
void copy(pmgx_ctx* c, double* a, const double* b, long long n)
{
  if (n <= 0)
    return;
  cudaSetDevice(c->device);
  PMGX_CUDA(cudaMemcpyAsync(a, b, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
}
void axpy(pmgx_ctx* c, double* r, double alpha, const double* x, const double* y, long long n)
{
  launch_stream<2>(c, r, x, y, alpha, n);
}
void scale(pmgx_ctx* c, double* r, double alpha, long long n)
{
  launch_stream<3>(c, r, nullptr, nullptr, alpha, n);
}
void pointwise_mult(pmgx_ctx* c, double* w, const double* x, const double* y, long long n)
{
  launch_stream<4>(c, w, x, y, 0.0, n);
}
void mask_bc(pmgx_ctx* c, double* b, const int8_t* bc, long long n)
{
  if (n <= 0)
    return;
  cudaSetDevice(c->device);
  k_mask_bc<<<stream_grid(c, n, 4), VT, 0, c->stream>>>(b, bc, n);
  check_launch("k_mask_bc");
  count_launch(c);
}

static int red_grid(pmgx_ctx* c, long long n)
{
  long long blocks = (n + RED_THREADS * 8 - 1) / (RED_THREADS * 8);
  if (blocks > c->max_red_blocks)
    blocks = c->max_red_blocks;
  if (blocks < 1)
    blocks = 1;
  return (int)blocks;
}

void allreduce_scalars(pmgx_ctx* c, int slot, int count, bool is_max)
{
  p2p::allreduce(c, slot, count, is_max); // peer-memory kernel over NVLink, NCCL when unavailable
}

void dot_device(pmgx_ctx* c, const double* a, const double* b, long long n, int slot)
{
  cudaSetDevice(c->device);
  const bool v2 = aligned16(a) && aligned16(b);
  k_dot<<<red_grid(c, n), RED_THREADS, 0, c->stream>>>(a, b, n, v2, c->d_partials, c->d_counter,
                                                       c->d_scalars + slot);
  check_launch("k_dot");
  count_launch(c);
  allreduce_scalars(c, slot, 1, false);
}

namespace
{
__global__ void k_publish(const double* __restrict__ src, double* __restrict__ dst_host, int count)
{
  if ((int)threadIdx.x < count)
    dst_host[threadIdx.x] = src[threadIdx.x];
  __threadfence_system();
}
} // namespace

void publish_scalars(pmgx_ctx* c, int slot, int count)
{
  cudaSetDevice(c->device);
  k_publish<<<1, 32, 0, c->stream>>>(c->d_scalars + slot, c->h_scalars_dev + slot, count);
  check_launch("k_publish");
  count_launch(c);
  PMGX_CUDA(cudaStreamSynchronize(c->stream));
}

double read_scalar(pmgx_ctx* c, int slot)
{
  publish_scalars(c, slot, 1);
  return c->h_scalars[slot];
}

double dot(pmgx_ctx* c, const double* a, const double* b, long long n)
{
  dot_device(c, a, b, n, 0);
  return read_scalar(c, 0);
}

double norm_linf(pmgx_ctx* c, const double* a, long long n)
{
  cudaSetDevice(c->device);
  k_absmax<<<red_grid(c, n), RED_THREADS, 0, c->stream>>>(a, n, c->d_partials, c->d_counter,
                                                          c->d_scalars);
  check_launch("k_absmax");
  count_launch(c);
  allreduce_scalars(c, 0, 1, true);
  return read_scalar(c, 0);
}
} // namespace vec
} // namespace pmgx

extern "C"
{
int pmgx_vec_set(pmgx_ctx* c, double* x, long long n, double v)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(c && (x || n == 0) && n >= 0, "vec_set: bad arguments");
  pmgx::vec::set(c, x, n, v);
  PMGX_API_END
}
int pmgx_vec_copy(pmgx_ctx* c, double* a, const double* b, long long n)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(c && n >= 0, "vec_copy: bad arguments");
  pmgx::vec::copy(c, a, b, n);
  PMGX_API_END
}
int pmgx_vec_axpy(pmgx_ctx* c, double* r, double alpha, const double* x, const double* y, long long n)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(c && n >= 0, "vec_axpy: bad arguments");
  pmgx::vec::axpy(c, r, alpha, x, y, n);
  PMGX_API_END
}
int pmgx_vec_scale(pmgx_ctx* c, double* r, double alpha, long long n)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(c && n >= 0, "vec_scale: bad arguments");
  pmgx::vec::scale(c, r, alpha, n);
  PMGX_API_END
}
int pmgx_vec_pointwise_mult(pmgx_ctx* c, double* w, const double* x, const double* y, long long n)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(c && n >= 0, "vec_pointwise_mult: bad arguments");
  pmgx::vec::pointwise_mult(c, w, x, y, n);
  PMGX_API_END
}
int pmgx_vec_mask_bc(pmgx_ctx* c, double* b, const int8_t* bc, long long n)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(c && n >= 0, "vec_mask_bc: bad arguments");
  pmgx::vec::mask_bc(c, b, bc, n);
  PMGX_API_END
}
int pmgx_vec_dot(pmgx_ctx* c, const double* a, const double* b, long long n, double* result_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(c && n >= 0 && result_h, "vec_dot: bad arguments");
  *result_h = pmgx::vec::dot(c, a, b, n);
  PMGX_API_END
}
int pmgx_vec_norm(pmgx_ctx* c, const double* a, long long n, int linf, double* result_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(c && n >= 0 && result_h, "vec_norm: bad arguments");
  if (linf)
    *result_h = pmgx::vec::norm_linf(c, a, n);
  else
    *result_h = sqrt(pmgx::vec::dot(c, a, a, n));
  PMGX_API_END
}
}
