// extern "C" launchers around the reference's own kernels (compiled from /root/reference by the
// Makefile next to this file).  TEST INFRASTRUCTURE ONLY: used by tests/ and bench.py to pin
// parity against, and time, the reference's kernels on the same GPU.  All pointers are device
// pointers; launch shapes are the reference's (cited in the Makefile header).
#include <cuda_runtime.h>
#include REF_GEN

#define REF_DISPATCH_P(P, ...)                                                                     \
  switch (P)                                                                                       \
  {                                                                                                \
  case 1: { constexpr int PP = 1; __VA_ARGS__; break; }                                                   \
  case 2: { constexpr int PP = 2; __VA_ARGS__; break; }                                                   \
  case 3: { constexpr int PP = 3; __VA_ARGS__; break; }                                                   \
  case 4: { constexpr int PP = 4; __VA_ARGS__; break; }                                                   \
  case 5: { constexpr int PP = 5; __VA_ARGS__; break; }                                                   \
  case 6: { constexpr int PP = 6; __VA_ARGS__; break; }                                                   \
  case 7: { constexpr int PP = 7; __VA_ARGS__; break; }                                                   \
  case 8: { constexpr int PP = 8; __VA_ARGS__; break; }                                                   \
  default: return -1;                                                                              \
  }

static int finish()
{
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess)
    e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

extern "C"
{
// src/laplacian.hpp:352-371 (compute_geometry): grid = cells, block = nq, 24 doubles of smem
int ref_geometry(int P, const double* xgeom, double* G, const int32_t* geom_dofmap, const double* dphi,
                 const double* weights, const int* entities, int n)
{
  const int nq = (P + 1) * (P + 1) * (P + 1);
  REF_DISPATCH_P(P, (geometry_computation<double, PP><<<n, nq, 24 * sizeof(double)>>>(
                        xgeom, G, geom_dofmap, dphi, weights, entities, n)));
  return finish();
}

// src/laplacian.hpp:398-409: grid = cells, block = (P+1,P+1,P+1), smem = 4 (P+1)^3 doubles
int ref_stiffness(int P, const double* x, const double* constants, double* y, const double* G,
                  const int32_t* dofmap, const double* dphi, const int* entities, int n, const int8_t* bc,
                  int sync)
{
  const int nd = P + 1;
  dim3 block(nd, nd, nd);
  const size_t shm = 4 * nd * nd * nd * sizeof(double);
  REF_DISPATCH_P(P, {
    if (shm > 48 * 1024)
      cudaFuncSetAttribute(stiffness_operator<double, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    stiffness_operator<double, PP><<<n, block, shm>>>(x, constants, y, G, dofmap, dphi, entities, n, bc);
  });
  return sync ? finish() : 0;
}

int ref_pack(int n, const int32_t* idx, const double* in, double* out)
{
  pack<double><<<(n + 511) / 512, 512>>>(n, idx, in, out);
  return finish();
}
int ref_unpack(int n, const int32_t* idx, const double* in, double* out)
{
  unpack<double><<<(n + 511) / 512, 512>>>(n, idx, in, out);
  return finish();
}
int ref_unpack_add(int n, const int32_t* idx, const double* in, double* out)
{
  unpack_add<double><<<(n + 511) / 512, 512>>>(n, idx, in, out);
  return finish();
}

// src/interpolate.hpp:198-208 / :260-271: one thread per cell, 256-thread blocks
int ref_interpolate_Q1Q2(int n, const int32_t* cells, const int32_t* dm1, int nd1, const int32_t* dm2, int nd2,
                         const double* v1, double* v2, const int32_t* Mptr, const int32_t* Mcols,
                         const double* Mvals)
{
  interpolate_Q1Q2<double><<<(n + 255) / 256, 256>>>(n, cells, dm1, nd1, dm2, nd2, v1, v2, Mptr, Mcols, Mvals);
  return finish();
}
int ref_interpolate_Q2Q1(int n, const int32_t* cells, const int32_t* dm1, int nd1, const int32_t* dm2, int nd2,
                         double* v1, const double* v2, const int32_t* MptrT, const int32_t* McolsT,
                         const double* MvalsT, const double* mult)
{
  interpolate_Q2Q1<double><<<(n + 255) / 256, 256>>>(n, cells, dm1, nd1, dm2, nd2, v1, v2, MptrT, McolsT, MvalsT, mult);
  return finish();
}

// src/csr.hpp:253-268: y += A x, one thread per row, 256-thread blocks
int ref_spmv(int n, const double* values, const int32_t* row_begin, const int32_t* row_end,
             const int32_t* cols, const double* x, double* y)
{
  spmv_impl<double><<<(n + 255) / 256, 256>>>(n, values, row_begin, row_end, cols, x, y);
  return finish();
}
int ref_spmvT(int n, const double* values, const int32_t* row_begin, const int32_t* row_end,
              const int32_t* cols, const double* x, double* y)
{
  spmvT_impl<double><<<(n + 255) / 256, 256>>>(n, values, row_begin, row_end, cols, x, y);
  return finish();
}

// host: src/cg.hpp:56-84
int ref_tqli(double* d, double* e, int n) { return tqli<double>(d, e, n); }
}
