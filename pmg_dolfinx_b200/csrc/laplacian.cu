// Matrix-free hexahedral Laplacian for P1..P8 on sm_100a.
//
// Replaces (reference paths):
//   geometry_computation<T,P>   src/laplacian.hpp:22-113
//   stiffness_operator<T,P>     src/laplacian.hpp:143-278
//   MatFreeLaplacian<T>         src/laplacian.hpp:283-526
//   diagonal via assembled CSR  examples/pmg/main.cpp:274-279, src/csr.hpp:101-112
//
// Design (not a port): cells are re-laid-out at create time in launch order (lcells then
// bcells); the dofmap is snapshotted with the Dirichlet marker folded into the sign bit, so
// the apply never gathers bc_marker; G is stored component-major per cell, G[p][6][nq], so
// every warp load of a G component is one contiguous run.  The apply kernel maps a thread to
// one (iy,iz) column of a cell and keeps the x-direction in registers: the x contractions
// never touch shared memory, the y/z contractions read planes from shared memory, and the
// element is streamed plane by plane so only 2 small plane buffers are exchanged between
// threads.  G is streamed with evict-first loads and software-prefetched one plane ahead.
#include "common.hpp"
#include "operator.hpp"
#include "csr.hpp"

#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/reduce.h>
#include <thrust/sort.h>
#include <thrust/scan.h>

#include <cmath>
#include <cstring>

namespace pmgx
{
namespace
{
constexpr int MAXN = PMGX_MAX_DEGREE + 1;

// 1-D tables for every degree; index [P][...]
__constant__ double c_D[PMGX_MAX_DEGREE + 1][MAXN * MAXN]; // D[q*n+i] = l_i'(x_q)
__constant__ double c_pts[PMGX_MAX_DEGREE + 1][MAXN];
__constant__ double c_wts[PMGX_MAX_DEGREE + 1][MAXN];

__device__ __forceinline__ double ldg_stream(const double* p)
{
  // streamed exactly once per apply: ld.global.cs (evict-first) keeps the gathered /
  // scattered x and y lines resident in L2 instead of the one-shot G / dofmap stream
  return __ldcs(p);
}
__device__ __forceinline__ int ldg_stream_i32(const int* p) { return __ldcs(p); }

// ------------------------------------------------------------------ set-up kernels --
__global__ void k_encode_dofmap(const int32_t* __restrict__ dofmap, const int32_t* __restrict__ perm,
                                const int8_t* __restrict__ bc, int32_t* __restrict__ enc, int n3,
                                long long total)
{
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += nth)
  {
    const long long p = t / n3;
    const int a = (int)(t - p * n3);
    const int32_t d = dofmap[(long long)perm[p] * n3 + a];
    enc[t] = bc[d] ? ~d : d;
  }
}

// One thread per (cell position, quadrature point).  J = sum_k x_k dphi_k, K = adj(J),
// G = w K K^T / detJ with 6 unique entries (xx,xy,xz,yy,yz,zz).  Also used for detJ only.
template <bool WRITE_G>
__global__ void k_geometry(int P, const double* __restrict__ xgeom,
                           const int32_t* __restrict__ geom_dofmap,
                           const int32_t* __restrict__ perm, double* __restrict__ G,
                           double* __restrict__ detj_w, int n_list, bool literal_detj)
{
  const int n = P + 1, n2 = n * n, n3 = n2 * n;
  const long long total = (long long)n_list * n3;
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += nth)
  {
    const long long p = t / n3;
    const int q = (int)(t - p * n3);
    const int ix = q / n2, iy = (q / n) % n, iz = q % n;
    const double xi[3] = {c_pts[P][ix], c_pts[P][iy], c_pts[P][iz]};
    const int32_t* gd = geom_dofmap + (long long)perm[p] * 8;
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int k = 0; k < 8; ++k)
    {
      const int a = (k >> 2) & 1, b = (k >> 1) & 1, c = k & 1;
      const double la = a ? xi[0] : 1.0 - xi[0], lb = b ? xi[1] : 1.0 - xi[1],
                   lc = c ? xi[2] : 1.0 - xi[2];
      const double da = a ? 1.0 : -1.0, db = b ? 1.0 : -1.0, dc = c ? 1.0 : -1.0;
      const double dphi[3] = {da * lb * lc, la * db * lc, la * lb * dc};
      const double* xv = xgeom + 3ll * gd[k];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
          J[i][j] += xv[i] * dphi[j];
    }
    double K[3][3];
    K[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    K[0][1] = -J[0][1] * J[2][2] + J[0][2] * J[2][1];
    K[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
    K[1][0] = -J[1][0] * J[2][2] + J[1][2] * J[2][0];
    K[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
    K[1][2] = -J[0][0] * J[1][2] + J[0][2] * J[1][0];
    K[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    K[2][1] = -J[0][0] * J[2][1] + J[0][1] * J[2][0];
    K[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    double detJ;
    if (literal_detj) // expression of src/laplacian.hpp:97 (quirk Q17)
      detJ = J[0][0] * K[0][0] - J[1][0] * K[0][1] + J[0][2] * K[2][0];
    else
      detJ = J[0][0] * K[0][0] + J[0][1] * K[1][0] + J[0][2] * K[2][0];
    const double w = c_wts[P][ix] * c_wts[P][iy] * c_wts[P][iz];
    if (WRITE_G)
    {
      const double s = w / detJ;
      double* g = G + p * 6 * n3 + q;
      g[0 * n3] = (K[0][0] * K[0][0] + K[0][1] * K[0][1] + K[0][2] * K[0][2]) * s;
      g[1 * n3] = (K[1][0] * K[0][0] + K[1][1] * K[0][1] + K[1][2] * K[0][2]) * s;
      g[2 * n3] = (K[2][0] * K[0][0] + K[2][1] * K[0][1] + K[2][2] * K[0][2]) * s;
      g[3 * n3] = (K[1][0] * K[1][0] + K[1][1] * K[1][1] + K[1][2] * K[1][2]) * s;
      g[4 * n3] = (K[2][0] * K[1][0] + K[2][1] * K[1][1] + K[2][2] * K[1][2]) * s;
      g[5 * n3] = (K[2][0] * K[2][0] + K[2][1] * K[2][1] + K[2][2] * K[2][2]) * s;
    }
    else
      detj_w[t] = w * fabs(detJ);
  }
}

// diag(A) contributions, thread per (cell position, local dof); see DESIGN.md "diagonal".
__global__ void k_diag(int P, const double* __restrict__ G, const int32_t* __restrict__ enc,
                       const int32_t* __restrict__ perm, const double* __restrict__ kappa,
                       double* __restrict__ diag, int n_list)
{
  const int n = P + 1, n2 = n * n, n3 = n2 * n;
  const long long total = (long long)n_list * n3;
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += nth)
  {
    const long long p = t / n3;
    const int a = (int)(t - p * n3);
    const int32_t d = enc[t];
    if (d < 0)
      continue;
    const int i = a / n2, j = (a / n) % n, k = a % n;
    const double* g = G + p * 6 * n3;
    const double* D = c_D[P];
    double s = 0.0;
    for (int q = 0; q < n; ++q)
    {
      const double dx = D[q * n + i], dy = D[q * n + j], dz = D[q * n + k];
      s += dx * dx * g[0 * n3 + q * n2 + j * n + k];
      s += dy * dy * g[3 * n3 + i * n2 + q * n + k];
      s += dz * dz * g[5 * n3 + i * n2 + j * n + q];
    }
    const double dii = D[i * n + i], djj = D[j * n + j], dkk = D[k * n + k];
    s += 2.0 * (g[1 * n3 + a] * dii * djj + g[2 * n3 + a] * dii * dkk + g[4 * n3 + a] * djj * dkk);
    atomicAdd(&diag[d], kappa[perm[p]] * s);
  }
}

__global__ void k_invert_diag(const double* __restrict__ diag, const int8_t* __restrict__ bc,
                              double* __restrict__ dinv, int n)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    dinv[i] = bc[i] ? 1.0 : 1.0 / diag[i];
}

__global__ void k_G_to_reference_layout(const double* __restrict__ G, double* __restrict__ out,
                                        int n3, long long total)
{
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += nth)
  {
    const long long p = t / (6 * n3);
    const int r = (int)(t - p * 6 * n3);
    const int q = r / 6, c = r % 6;
    out[t] = G[p * 6 * n3 + (long long)c * n3 + q];
  }
}

__global__ void k_rhs(const double* __restrict__ detj_w, const int32_t* __restrict__ enc,
                      const double* __restrict__ fvals, double* __restrict__ b, long long total)
{
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += nth)
  {
    const int32_t d = enc[t];
    if (d >= 0)
      atomicAdd(&b[d], fvals[d] * detj_w[t]);
  }
}

__global__ void k_set_bc_value(double* __restrict__ b, const int8_t* __restrict__ bc, double g, int n)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && bc[i])
    b[i] = g;
}

// ------------------------------------------------------------- CSR assembly kernels --
// Entry (i,j) of the element matrix kappa * B^T G B with the collocated gradient table
// (phi = identity at the GLL points, src/laplacian.hpp:200-202): only index pairs that
// agree in at least one direction couple.
__device__ double element_entry(int P, const double* __restrict__ g, int i, int j)
{
  const int n = P + 1, n2 = n * n, n3 = n2 * n;
  const double* D = c_D[P];
  const int ix = i / n2, iy = (i / n) % n, iz = i % n;
  const int jx = j / n2, jy = (j / n) % n, jz = j % n;
  double s = 0.0;
  if (iy == jy && iz == jz)
    for (int q = 0; q < n; ++q)
      s += D[q * n + ix] * D[q * n + jx] * g[0 * n3 + q * n2 + iy * n + iz];
  if (ix == jx && iz == jz)
    for (int q = 0; q < n; ++q)
      s += D[q * n + iy] * D[q * n + jy] * g[3 * n3 + ix * n2 + q * n + iz];
  if (ix == jx && iy == jy)
    for (int q = 0; q < n; ++q)
      s += D[q * n + iz] * D[q * n + jz] * g[5 * n3 + ix * n2 + iy * n + q];
  if (iz == jz)
    s += D[jx * n + ix] * D[iy * n + jy] * g[1 * n3 + jx * n2 + iy * n + iz]
         + D[jy * n + iy] * D[ix * n + jx] * g[1 * n3 + ix * n2 + jy * n + iz];
  if (iy == jy)
    s += D[jx * n + ix] * D[iz * n + jz] * g[2 * n3 + jx * n2 + iy * n + iz]
         + D[jz * n + iz] * D[ix * n + jx] * g[2 * n3 + ix * n2 + iy * n + jz];
  if (ix == jx)
    s += D[jy * n + iy] * D[iz * n + jz] * g[4 * n3 + ix * n2 + jy * n + iz]
         + D[jz * n + iz] * D[iy * n + jy] * g[4 * n3 + ix * n2 + iy * n + jz];
  return s;
}

// One (key, value) per (cell, i, j); key = row * ntot + col, sentinel for dropped entries
// (ghost rows, Dirichlet rows / columns: fem::assemble_matrix with bcs, src/csr.hpp:84).
__global__ void k_element_triplets(int P, const double* __restrict__ G, const int32_t* __restrict__ enc,
                                   const int32_t* __restrict__ perm, const double* __restrict__ kappa,
                                   long long n_list, int n_owned, long long ntot,
                                   unsigned long long* __restrict__ keys, double* __restrict__ vals)
{
  const int n = P + 1, n3 = n * n * n;
  const long long per_cell = (long long)n3 * n3;
  const long long total = n_list * per_cell;
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += nth)
  {
    const long long p = t / per_cell;
    const int r = (int)(t - p * per_cell);
    const int i = r / n3, j = r - i * n3;
    const int32_t di = enc[p * n3 + i], dj = enc[p * n3 + j];
    if (di < 0 || dj < 0 || di >= n_owned)
    {
      keys[t] = ~0ull;
      vals[t] = 0.0;
      continue;
    }
    keys[t] = (unsigned long long)di * (unsigned long long)ntot + (unsigned long long)dj;
    vals[t] = kappa[perm[p]] * element_entry(P, G + p * 6 * n3, i, j);
  }
}

// Dirichlet rows: unit diagonal (fem::set_diagonal, src/csr.hpp:86)
__global__ void k_bc_diag_triplets(const int8_t* __restrict__ bc, int n_owned, long long ntot,
                                   unsigned long long* __restrict__ keys, double* __restrict__ vals)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_owned)
    return;
  keys[i] = bc[i] ? (unsigned long long)i * (unsigned long long)ntot + (unsigned long long)i : ~0ull;
  vals[i] = bc[i] ? 1.0 : 0.0;
}

__global__ void k_split_keys(const unsigned long long* __restrict__ keys, long long nnz, long long ntot,
                             int n_owned, int32_t* __restrict__ cols, int32_t* __restrict__ row_count,
                             int32_t* __restrict__ owned_count)
{
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += nth)
  {
    const unsigned long long k = keys[t];
    const int row = (int)(k / (unsigned long long)ntot);
    const int col = (int)(k - (unsigned long long)row * (unsigned long long)ntot);
    cols[t] = col;
    atomicAdd(&row_count[row], 1);
    if (col < n_owned)
      atomicAdd(&owned_count[row], 1);
  }
}

__global__ void k_offdiag(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ owned_count,
                          int n_rows, int32_t* __restrict__ off_diag)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_rows)
    off_diag[i] = row_ptr[i] + owned_count[i];
}

// ----------------------------------------------------------------- the apply kernel --
// smem row stride: pad when a (j -> j+1) step of n doubles would alias banks (n % 8 == 0)
template <int P>
struct ApplyCfg
{
  static constexpr int n = P + 1;
  static constexpr int n2 = n * n;
  static constexpr int n3 = n2 * n;
  static constexpr int row = (n % 8 == 0) ? n + 1 : n;
  static constexpr int plane = n * row;
  // threads per block: the smallest of {128,160,192,256} wasting the fewest lanes
  static constexpr int tpb = (P == 5 || P == 6) ? 160 : (P == 8 ? 256 : 128);
  static constexpr int cpb = tpb / n2; // cells per block
  static constexpr int smem_doubles = cpb * (n * plane + 4 * plane);
};

template <int P>
__global__ void __launch_bounds__(ApplyCfg<P>::tpb)
k_apply(const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ G,
        const int32_t* __restrict__ enc, const int32_t* __restrict__ perm,
        const double* __restrict__ kappa, int first, int count)
{
  using C = ApplyCfg<P>;
  constexpr int n = C::n, n2 = C::n2, n3 = C::n3, ROW = C::row, PL = C::plane, CPB = C::cpb;
  extern __shared__ double smem[];
  double* su = smem;                      // [CPB][n][PL]
  double* sf = smem + CPB * n * PL;       // [2][CPB][2][PL]

  const int tid = threadIdx.x;
  const int cl = tid / n2;
  const int jk = tid - cl * n2;
  const int j = jk / n, k = jk - j * n;
  const long long pl = (long long)blockIdx.x * CPB + cl;
  const bool active = (cl < CPB) && (pl < count);
  const long long p = first + pl;
  const int sjk = j * ROW + k;

  int d[n];
  double u[n];
  double kap = 0.0;
  const double* Gp = G + p * 6 * n3 + jk;
  double g[6];
  if (active)
  {
    const int32_t* e = enc + p * n3 + jk;
#pragma unroll
    for (int i = 0; i < n; ++i)
      d[i] = ldg_stream_i32(e + i * n2);
#pragma unroll
    for (int c = 0; c < 6; ++c)
      g[c] = ldg_stream(Gp + c * n3);
    kap = kappa[perm[p]];
#pragma unroll
    for (int i = 0; i < n; ++i)
    {
      const int idx = d[i] < 0 ? ~d[i] : d[i];
      const double xv = x[idx];
      if (d[i] < 0)
        y[idx] = xv; // Dirichlet row: y = x (src/laplacian.hpp:273-274)
      u[i] = d[i] < 0 ? 0.0 : xv;
      su[(cl * n + i) * PL + sjk] = u[i];
    }
  }
  else
  {
#pragma unroll
    for (int i = 0; i < n; ++i)
      u[i] = 0.0, d[i] = -1;
#pragma unroll
    for (int c = 0; c < 6; ++c)
      g[c] = 0.0;
  }
  // rows of D needed with a runtime index
  double Dj[n], Dk[n], DTj[n], DTk[n];
#pragma unroll
  for (int l = 0; l < n; ++l)
  {
    Dj[l] = c_D[P][j * n + l];
    Dk[l] = c_D[P][k * n + l];
    DTj[l] = c_D[P][l * n + j];
    DTk[l] = c_D[P][l * n + k];
  }
  double acc[n];
#pragma unroll
  for (int i = 0; i < n; ++i)
    acc[i] = 0.0;
  __syncthreads();

  const int cls = active ? cl : 0;
#pragma unroll
  for (int i = 0; i < n; ++i)
  {
    // prefetch next plane's geometry
    double gn[6];
    if (i + 1 < n)
    {
      if (active)
      {
#pragma unroll
        for (int c = 0; c < 6; ++c)
          gn[c] = ldg_stream(Gp + c * n3 + (i + 1) * n2);
      }
      else
      {
#pragma unroll
        for (int c = 0; c < 6; ++c)
          gn[c] = 0.0;
      }
    }
    // gradient at quadrature point (i, j, k)
    double gx = 0.0, gy = 0.0, gz = 0.0;
    const double* sp = su + (cls * n + i) * PL;
#pragma unroll
    for (int l = 0; l < n; ++l)
    {
      gx = fma(c_D[P][i * n + l], u[l], gx);
      gy = fma(Dj[l], sp[l * ROW + k], gy);
      gz = fma(Dk[l], sp[j * ROW + l], gz);
    }
    const double fx = kap * (g[0] * gx + g[1] * gy + g[2] * gz);
    const double fy = kap * (g[1] * gx + g[3] * gy + g[4] * gz);
    const double fz = kap * (g[2] * gx + g[4] * gy + g[5] * gz);
#pragma unroll
    for (int l = 0; l < n; ++l)
      acc[l] = fma(c_D[P][i * n + l], fx, acc[l]);
    double* sfy = sf + (((i & 1) * CPB + cls) * 2 + 0) * PL;
    double* sfz = sfy + PL;
    if (active)
    {
      sfy[sjk] = fy;
      sfz[sjk] = fz;
    }
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < n; ++q)
    {
      t = fma(DTj[q], sfy[q * ROW + k], t);
      t = fma(DTk[q], sfz[j * ROW + q], t);
    }
    acc[i] += t;
    if (i + 1 < n)
    {
#pragma unroll
      for (int c = 0; c < 6; ++c)
        g[c] = gn[c];
    }
  }
  if (active)
  {
#pragma unroll
    for (int i = 0; i < n; ++i)
      if (d[i] >= 0)
        atomicAdd(&y[d[i]], acc[i]);
  }
}

template <int P>
void launch_apply(pmgx_ctx* c, const double* x, double* y, const double* G, const int32_t* enc,
                  const int32_t* perm, const double* kappa, int first, int count)
{
  if (count <= 0)
    return;
  using C = ApplyCfg<P>;
  const size_t smem = (size_t)C::smem_doubles * sizeof(double);
  static bool configured[64] = {false};
  if (!configured[c->device])
  {
    PMGX_CUDA(cudaFuncSetAttribute(k_apply<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[c->device] = true;
  }
  const int grid = (count + C::cpb - 1) / C::cpb;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (c->profiling)
  {
    PMGX_CUDA(cudaEventCreate(&e0));
    PMGX_CUDA(cudaEventCreate(&e1));
    PMGX_CUDA(cudaEventRecord(e0, c->stream));
  }
  k_apply<P><<<grid, C::tpb, smem, c->stream>>>(x, y, G, enc, perm, kappa, first, count);
  check_launch("k_apply");
  count_launch(c);
  if (c->profiling)
  {
    PMGX_CUDA(cudaEventRecord(e1, c->stream));
    c->prof[P].emplace_back(e0, e1);
  }
}

void upload_tables(pmgx_ctx* c)
{
  static double hD[PMGX_MAX_DEGREE + 1][MAXN * MAXN];
  static double hp[PMGX_MAX_DEGREE + 1][MAXN], hw[PMGX_MAX_DEGREE + 1][MAXN];
  std::memset(hD, 0, sizeof(hD));
  std::memset(hp, 0, sizeof(hp));
  std::memset(hw, 0, sizeof(hw));
  for (int P = 1; P <= PMGX_MAX_DEGREE; ++P)
  {
    std::vector<double> xs, ws, D;
    gll_points_weights(P + 1, xs, ws);
    gll_deriv_matrix(xs, D);
    for (int i = 0; i <= P; ++i)
      hp[P][i] = xs[i], hw[P][i] = ws[i];
    for (size_t i = 0; i < D.size(); ++i)
      hD[P][i] = D[i];
  }
  PMGX_CUDA(cudaMemcpyToSymbolAsync(c_D, hD, sizeof(hD), 0, cudaMemcpyHostToDevice, c->stream));
  PMGX_CUDA(cudaMemcpyToSymbolAsync(c_pts, hp, sizeof(hp), 0, cudaMemcpyHostToDevice, c->stream));
  PMGX_CUDA(cudaMemcpyToSymbolAsync(c_wts, hw, sizeof(hw), 0, cudaMemcpyHostToDevice, c->stream));
  PMGX_CUDA(cudaStreamSynchronize(c->stream));
}

inline int setup_grid(pmgx_ctx* c, long long total)
{
  long long b = (total + 255) / 256;
  long long cap = (long long)c->num_sms * 32;
  return (int)std::max<long long>(1, std::min(b, cap));
}
} // namespace

struct Laplacian : pmgx_operator
{
  int P = 0, n3 = 0;
  int n_cells = 0, n_l = 0, n_b = 0;
  const double* kappa = nullptr; // borrowed (cell_constants, src/laplacian.hpp:501)
  const int8_t* bc = nullptr;    // borrowed
  const double* xgeom = nullptr;
  const int32_t* geom_dofmap = nullptr;
  int flags = 0;
  DevBuf<int32_t> perm; // launch position -> caller cell index (lcells then bcells)
  DevBuf<int32_t> enc;  // [n_list][n3] BC-encoded dofmap
  DevBuf<double> G;     // [n_list][6][n3]

  int n_list() const { return n_l + n_b; }

  template <int PP>
  void apply_t(double* x, double* y)
  {
    const long long ntot = (long long)n_owned + n_ghost;
    PMGX_CUDA(cudaMemsetAsync(y, 0, (size_t)ntot * sizeof(double), ctx->stream)); // out.set(0) :466
    if (halo)
      halo_fwd_begin(halo, x);                                                    // :378
    launch_apply<PP>(ctx, x, y, G.p, enc.p, perm.p, kappa, 0, n_l);               // :406-409
    if (halo)
      halo_fwd_end(halo, x);                                                      // :425
    launch_apply<PP>(ctx, x, y, G.p, enc.p, perm.p, kappa, n_l, n_b);             // :449-452
  }

  void apply(double* x, double* y) override
  {
    cudaSetDevice(ctx->device);
    switch (P)
    {
    case 1: apply_t<1>(x, y); break;
    case 2: apply_t<2>(x, y); break;
    case 3: apply_t<3>(x, y); break;
    case 4: apply_t<4>(x, y); break;
    case 5: apply_t<5>(x, y); break;
    case 6: apply_t<6>(x, y); break;
    case 7: apply_t<7>(x, y); break;
    case 8: apply_t<8>(x, y); break;
    default:
      set_error("Unsupported degree");
      throw Error{PMGX_ERR_UNSUPPORTED};
    }
  }
};
} // namespace pmgx

using pmgx::Laplacian;

extern "C"
{
int pmgx_laplacian_create(pmgx_ctx* ctx, int degree, int n_cells, const int32_t* dofmap,
                          const double* xgeom, int n_points, const int32_t* geom_dofmap,
                          const double* kappa, const int32_t* lcells_h, int n_lcells,
                          const int32_t* bcells_h, int n_bcells, const int8_t* bc_marker,
                          int n_owned, int n_ghost, pmgx_halo* halo, int flags, pmgx_operator** out)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(ctx && out, "laplacian_create: null ctx/out");
  if (degree < 1 || degree > PMGX_MAX_DEGREE)
  {
    pmgx::set_error("Unsupported degree"); // same message as src/laplacian.hpp:346,479
    return PMGX_ERR_UNSUPPORTED;
  }
  PMGX_REQUIRE(n_cells >= 0 && n_lcells >= 0 && n_bcells >= 0 && n_lcells + n_bcells <= n_cells,
               "laplacian_create: inconsistent cell counts");
  PMGX_REQUIRE(n_cells == 0 || (dofmap && xgeom && geom_dofmap && kappa && bc_marker),
               "laplacian_create: null array");
  PMGX_REQUIRE(n_owned >= 0 && n_ghost >= 0 && n_points >= 0, "laplacian_create: bad sizes");
  PMGX_REQUIRE(!halo || (halo->n_owned == n_owned && halo->n_ghost == n_ghost),
               "laplacian_create: halo does not match the vector layout");
  PMGX_CUDA(cudaSetDevice(ctx->device));
  pmgx::upload_tables(ctx);

  auto* L = new Laplacian();
  std::unique_ptr<Laplacian> guard(L);
  L->ctx = ctx;
  L->kind = pmgx_operator::LAPLACIAN;
  L->P = degree;
  const int n = degree + 1;
  L->n3 = n * n * n;
  L->n_cells = n_cells;
  L->n_l = n_lcells;
  L->n_b = n_bcells;
  L->n_owned = n_owned;
  L->n_ghost = n_ghost;
  L->halo = halo;
  L->kappa = kappa;
  L->bc = bc_marker;
  L->xgeom = xgeom;
  L->geom_dofmap = geom_dofmap;
  L->flags = flags;

  const int n_list = L->n_list();
  std::vector<int32_t> perm_h((size_t)n_list);
  for (int i = 0; i < n_lcells; ++i)
    perm_h[i] = lcells_h[i];
  for (int i = 0; i < n_bcells; ++i)
    perm_h[n_lcells + i] = bcells_h[i];
  for (int32_t v : perm_h)
    PMGX_REQUIRE(v >= 0 && v < n_cells, "laplacian_create: cell index %d out of range", v);
  L->perm.upload(perm_h.data(), perm_h.size(), ctx->stream);

  const long long total = (long long)n_list * L->n3;
  L->enc.alloc((size_t)total);
  L->G.alloc((size_t)total * 6);
  if (total > 0)
  {
    pmgx::k_encode_dofmap<<<pmgx::setup_grid(ctx, total), 256, 0, ctx->stream>>>(
        dofmap, L->perm.p, bc_marker, L->enc.p, L->n3, total);
    pmgx::check_launch("k_encode_dofmap");
    pmgx::k_geometry<true><<<pmgx::setup_grid(ctx, total), 256, 0, ctx->stream>>>(
        degree, xgeom, geom_dofmap, L->perm.p, L->G.p, nullptr, n_list,
        (flags & PMGX_LAP_LITERAL_DETJ) != 0);
    pmgx::check_launch("k_geometry");
    pmgx::count_launch(ctx, 2);
  }
  L->diag_inv.alloc((size_t)n_owned);
  if (!(flags & PMGX_LAP_NO_DIAG) && n_owned > 0)
  {
    pmgx::DevBuf<double> diag;
    diag.alloc((size_t)n_owned + n_ghost);
    PMGX_CUDA(cudaMemsetAsync(diag.p, 0, diag.n * sizeof(double), ctx->stream));
    if (total > 0)
    {
      pmgx::k_diag<<<pmgx::setup_grid(ctx, total), 256, 0, ctx->stream>>>(
          degree, L->G.p, L->enc.p, L->perm.p, kappa, diag.p, n_list);
      pmgx::check_launch("k_diag");
    }
    pmgx::k_invert_diag<<<(n_owned + 255) / 256, 256, 0, ctx->stream>>>(diag.p, bc_marker,
                                                                         L->diag_inv.p, n_owned);
    pmgx::check_launch("k_invert_diag");
    pmgx::count_launch(ctx, 2);
    PMGX_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  else if (n_owned > 0)
    PMGX_CUDA(cudaMemsetAsync(L->diag_inv.p, 0, (size_t)n_owned * sizeof(double), ctx->stream));
  PMGX_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = guard.release();
  PMGX_API_END
}

int pmgx_laplacian_get_G(pmgx_operator* op, double* G_out)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(op && op->kind == pmgx_operator::LAPLACIAN && G_out, "laplacian_get_G: bad arguments");
  auto* L = static_cast<Laplacian*>(op);
  const long long total = (long long)L->n_list() * L->n3 * 6;
  if (total > 0)
  {
    PMGX_CUDA(cudaSetDevice(L->ctx->device));
    pmgx::k_G_to_reference_layout<<<pmgx::setup_grid(L->ctx, total), 256, 0, L->ctx->stream>>>(
        L->G.p, G_out, L->n3, total);
    pmgx::check_launch("k_G_to_reference_layout");
    pmgx::count_launch(L->ctx);
  }
  PMGX_API_END
}

int pmgx_laplacian_rhs(pmgx_operator* op, const double* fvals, double g, double* b)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(op && op->kind == pmgx_operator::LAPLACIAN && fvals && b, "laplacian_rhs: bad arguments");
  auto* L = static_cast<Laplacian*>(op);
  pmgx_ctx* ctx = L->ctx;
  PMGX_CUDA(cudaSetDevice(ctx->device));
  const long long total = (long long)L->n_list() * L->n3;
  const int ntot = L->n_owned + L->n_ghost;
  PMGX_CUDA(cudaMemsetAsync(b, 0, (size_t)ntot * sizeof(double), ctx->stream));
  if (total > 0)
  {
    pmgx::DevBuf<double> dw;
    dw.alloc((size_t)total);
    pmgx::k_geometry<false><<<pmgx::setup_grid(ctx, total), 256, 0, ctx->stream>>>(
        L->P, L->xgeom, L->geom_dofmap, L->perm.p, nullptr, dw.p, L->n_list(), false);
    pmgx::check_launch("k_geometry(detJ)");
    pmgx::k_rhs<<<pmgx::setup_grid(ctx, total), 256, 0, ctx->stream>>>(dw.p, L->enc.p, fvals, b, total);
    pmgx::check_launch("k_rhs");
    pmgx::count_launch(ctx, 2);
    PMGX_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  if (ntot > 0)
  {
    pmgx::k_set_bc_value<<<(ntot + 255) / 256, 256, 0, ctx->stream>>>(b, L->bc, g, ntot);
    pmgx::check_launch("k_set_bc_value");
    pmgx::count_launch(ctx);
  }
  PMGX_API_END
}

int pmgx_csr_from_laplacian(pmgx_operator* op, pmgx_operator** out)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(op && op->kind == pmgx_operator::LAPLACIAN && out, "csr_from_laplacian: bad arguments");
  auto* L = static_cast<Laplacian*>(op);
  pmgx_ctx* ctx = L->ctx;
  PMGX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int n_owned = L->n_owned;
  const long long ntot = (long long)L->n_owned + L->n_ghost;
  const long long n_list = L->n_list();
  const long long per_cell = (long long)L->n3 * L->n3;
  const long long n_trip = n_list * per_cell + n_owned;
  pmgx::DevBuf<unsigned long long> keys, ukeys;
  pmgx::DevBuf<double> vals, uvals;
  keys.alloc((size_t)std::max<long long>(n_trip, 1));
  vals.alloc((size_t)std::max<long long>(n_trip, 1));
  if (n_list > 0)
  {
    pmgx::k_element_triplets<<<pmgx::setup_grid(ctx, n_list * per_cell), 256, 0, st>>>(
        L->P, L->G.p, L->enc.p, L->perm.p, L->kappa, n_list, n_owned, ntot, keys.p, vals.p);
    pmgx::check_launch("k_element_triplets");
  }
  if (n_owned > 0)
  {
    pmgx::k_bc_diag_triplets<<<(n_owned + 255) / 256, 256, 0, st>>>(
        L->bc, n_owned, ntot, keys.p + n_list * per_cell, vals.p + n_list * per_cell);
    pmgx::check_launch("k_bc_diag_triplets");
  }
  auto pol = thrust::cuda::par.on(st);
  thrust::device_ptr<unsigned long long> kp(keys.p);
  thrust::device_ptr<double> vp(vals.p);
  thrust::stable_sort_by_key(pol, kp, kp + n_trip, vp);
  ukeys.alloc((size_t)std::max<long long>(n_trip, 1));
  uvals.alloc((size_t)std::max<long long>(n_trip, 1));
  thrust::device_ptr<unsigned long long> ukp(ukeys.p);
  thrust::device_ptr<double> uvp(uvals.p);
  auto ends = thrust::reduce_by_key(pol, kp, kp + n_trip, vp, ukp, uvp);
  long long nnz = ends.first - ukp;
  if (nnz > 0)
  {
    unsigned long long last;
    PMGX_CUDA(cudaMemcpyAsync(&last, ukeys.p + nnz - 1, sizeof(last), cudaMemcpyDeviceToHost, st));
    PMGX_CUDA(cudaStreamSynchronize(st));
    if (last == ~0ull)
      --nnz; // the sentinel group sorts last
  }
  PMGX_REQUIRE(nnz < (1ll << 31), "csr_from_laplacian: more than 2^31 non-zeros");
  keys.release();
  vals.release();

  std::unique_ptr<pmgx::CsrOperator> A(new pmgx::CsrOperator());
  A->ctx = ctx;
  A->kind = pmgx_operator::CSR;
  A->n_owned = n_owned;
  A->n_ghost = L->n_ghost;
  A->halo = L->halo;
  A->nnz = nnz;
  A->row_ptr.alloc((size_t)n_owned + 1);
  A->off_diag.alloc((size_t)n_owned);
  A->cols.alloc((size_t)std::max<long long>(nnz, 1));
  A->values.alloc((size_t)std::max<long long>(nnz, 1));
  pmgx::DevBuf<int32_t> row_count, owned_count;
  row_count.alloc((size_t)n_owned + 1);
  owned_count.alloc((size_t)n_owned + 1);
  PMGX_CUDA(cudaMemsetAsync(row_count.p, 0, ((size_t)n_owned + 1) * sizeof(int32_t), st));
  PMGX_CUDA(cudaMemsetAsync(owned_count.p, 0, ((size_t)n_owned + 1) * sizeof(int32_t), st));
  if (nnz > 0)
  {
    pmgx::k_split_keys<<<pmgx::setup_grid(ctx, nnz), 256, 0, st>>>(ukeys.p, nnz, ntot, n_owned, A->cols.p,
                                                                 row_count.p, owned_count.p);
    pmgx::check_launch("k_split_keys");
    PMGX_CUDA(cudaMemcpyAsync(A->values.p, uvals.p, (size_t)nnz * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  thrust::device_ptr<int32_t> rc(row_count.p), rp(A->row_ptr.p);
  thrust::exclusive_scan(pol, rc, rc + n_owned + 1, rp);
  if (n_owned > 0)
  {
    pmgx::k_offdiag<<<(n_owned + 255) / 256, 256, 0, st>>>(A->row_ptr.p, owned_count.p, n_owned, A->off_diag.p);
    pmgx::check_launch("k_offdiag");
  }
  pmgx::count_launch(ctx, 6);
  PMGX_CUDA(cudaStreamSynchronize(st));
  // does any row address ghost columns?
  {
    thrust::device_ptr<int32_t> oc(owned_count.p);
    const long long owned_total = thrust::reduce(pol, oc, oc + n_owned, (long long)0);
    A->has_ghost_cols = owned_total < nnz;
  }
  A->finish_setup();
  *out = A.release();
  PMGX_API_END
}

// ----------------------------------------------------------------- generic operator --
int pmgx_operator_apply(pmgx_operator* op, double* x, double* y)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(op && x && y, "operator_apply: null argument");
  PMGX_REQUIRE(x != y, "operator_apply: in-place apply is not supported");
  op->apply(x, y);
  PMGX_API_END
}

int pmgx_operator_get_diag_inverse(pmgx_operator* op, double* out)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(op && out, "get_diag_inverse: null argument");
  pmgx::vec::copy(op->ctx, out, op->diag_inv.p, op->n_owned);
  PMGX_API_END
}

int pmgx_operator_set_diag_inverse(pmgx_operator* op, const double* in)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(op && in, "set_diag_inverse: null argument");
  pmgx::vec::copy(op->ctx, op->diag_inv.p, in, op->n_owned);
  PMGX_API_END
}

int pmgx_operator_n_owned(pmgx_operator* op) { return op ? op->n_owned : -1; }
int pmgx_operator_n_ghost(pmgx_operator* op) { return op ? op->n_ghost : -1; }

int pmgx_operator_destroy(pmgx_operator* op)
{
  PMGX_API_BEGIN
  if (op)
  {
    cudaSetDevice(op->ctx->device);
    cudaStreamSynchronize(op->ctx->stream);
    delete op;
  }
  PMGX_API_END
}
}
