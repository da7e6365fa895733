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
// the apply never gathers bc_marker; G is stored component-major per batch of cells so that one
// (ix) plane of a batch is one contiguous cp.async.bulk transaction.  Apply kernels:
//   k_apply_affine / k_apply_affine_shfl   every cell affine: one geometry 6-vector per cell
//   k_apply_tma                            streamed G through a TMA ring (P1..P6)
//   k_apply_slab                           same thread mapping, register-streamed G (A/B runs)
//   k_apply                                column kernel (P7, P8)
// In the slab kernels a thread owns one z-index of a cell and keeps the (ix,iy) slab of the
// element in registers: the x and y contractions are register FMAs against constant-memory
// derivative entries, only the z contraction crosses threads.
#include "common.hpp"
#include "operator.hpp"
#include "csr.hpp"

#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/reduce.h>
#include <thrust/sort.h>
#include <thrust/scan.h>

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace pmgx
{
namespace
{
constexpr int MAXN = PMGX_MAX_DEGREE + 1;
constexpr int SLAB_MAX_DEGREE = 6; // slab kernels (2 (P+1)^2 doubles of registers) up to here; above: column kernel (measured: P7 53 % vs 47 %, P8 24 % vs 27 %)

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


// How the operator-private arrays (BC-encoded dofmap, geometry factors) are laid out.  The
// reference keeps G private (src/laplacian.hpp:512), so the layout is free: it follows the
// thread mapping of the apply kernel so that every warp load is one aligned, contiguous run.
//   COLUMN: enc[p][n3], G[p][6][n3]                         (thread = (iy,iz) column, P >= 5)
//   SLAB  : enc[batch][n2][S], G[batch][n2][6][S]           (thread = iz slab,        P <= 4)
//           batch = CPB consecutive cells of the launch list, slot = cell_in_batch*n + iz,
//           S = 128 slots (padded); lcells and bcells batches are padded separately.
struct Lay
{
  int mode; // 0 COLUMN, 1 SLAB
  int n, n2, n3;
  int cpb, S, SE; // SE: slot stride of the int32 dofmap rows (multiple of 4 -> 16-byte rows)
  int n_l, nb_l;
  long long n_batches;
  __host__ __device__ long long enc_size() const { return mode == 0 ? 0 : n_batches * n2 * SE; }
  __host__ __device__ long long g_size() const { return mode == 0 ? 0 : n_batches * n2 * 6 * S; }
  __host__ __device__ long long enc_index(long long p, int a) const
  {
    if (mode == 0)
      return p * n3 + a;
    const long long pl = p < n_l ? p : p - n_l;
    const long long batch = (p < n_l ? 0 : nb_l) + pl / cpb;
    const int slot = (int)(pl % cpb) * n + a % n;
    return (batch * n2 + a / n) * SE + slot;
  }
  __host__ __device__ long long g_index(long long p, int comp, int q) const
  {
    if (mode == 0)
      return (p * 6 + comp) * n3 + q;
    const long long pl = p < n_l ? p : p - n_l;
    const long long batch = (p < n_l ? 0 : nb_l) + pl / cpb;
    const int slot = (int)(pl % cpb) * n + q % n;
    return ((batch * n2 + q / n) * 6 + comp) * S + slot;
  }
};

// ------------------------------------------------------------------ set-up kernels --
__global__ void k_encode_dofmap(const int32_t* __restrict__ dofmap, const int32_t* __restrict__ perm,
                                const int8_t* __restrict__ bc, int32_t* __restrict__ enc, int n3,
                                long long total, Lay lay)
{
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += nth)
  {
    const long long p = t / n3;
    const int a = (int)(t - p * n3);
    const int32_t d = dofmap[(long long)perm[p] * n3 + a];
    enc[lay.enc_index(p, a)] = bc[d] ? ~d : d;
  }
}

// One thread per (cell position, quadrature point).  J = sum_k x_k dphi_k, K = adj(J),
// G = w K K^T / detJ with 6 unique entries (xx,xy,xz,yy,yz,zz).  Also used for detJ only.
template <bool WRITE_G>
__global__ void k_geometry(int P, const double* __restrict__ xgeom,
                           const int32_t* __restrict__ geom_dofmap,
                           const int32_t* __restrict__ perm, double* __restrict__ G,
                           double* __restrict__ detj_w, int n_list, bool literal_detj, Lay lay)
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
      G[lay.g_index(p, 0, q)] = (K[0][0] * K[0][0] + K[0][1] * K[0][1] + K[0][2] * K[0][2]) * s;
      G[lay.g_index(p, 1, q)] = (K[1][0] * K[0][0] + K[1][1] * K[0][1] + K[1][2] * K[0][2]) * s;
      G[lay.g_index(p, 2, q)] = (K[2][0] * K[0][0] + K[2][1] * K[0][1] + K[2][2] * K[0][2]) * s;
      G[lay.g_index(p, 3, q)] = (K[1][0] * K[1][0] + K[1][1] * K[1][1] + K[1][2] * K[1][2]) * s;
      G[lay.g_index(p, 4, q)] = (K[2][0] * K[1][0] + K[2][1] * K[1][1] + K[2][2] * K[1][2]) * s;
      G[lay.g_index(p, 5, q)] = (K[2][0] * K[2][0] + K[2][1] * K[2][1] + K[2][2] * K[2][2]) * s;
    }
    else
      detj_w[t] = w * fabs(detJ);
  }
}

// Per cell: Gc = K K^T / det J of the trilinear map at the cell centre (weight 1), and the affinity
// test: a cell is affine iff every vertex equals v000 + a e1 + b e2 + c e3; the deviation is
// measured against the longest edge.  One thread per launch position.
__global__ void k_cell_geometry(const double* __restrict__ xgeom, const int32_t* __restrict__ geom_dofmap,
                                const int32_t* __restrict__ perm, double* __restrict__ Gc, int n_list,
                                bool literal_detj, double tol, int* __restrict__ n_nonaffine)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_list)
    return;
  const int32_t* gd = geom_dofmap + (long long)perm[p] * 8;
  double v[8][3];
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int i = 0; i < 3; ++i)
      v[k][i] = xgeom[3ll * gd[k] + i];
  // tensor-product vertex order: k = 4a + 2b + c (src/mesh.hpp:75-84)
  double scale = 0.0, dev = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    const double e1 = v[4][i] - v[0][i], e2 = v[2][i] - v[0][i], e3 = v[1][i] - v[0][i];
    scale = fmax(scale, fmax(fabs(e1), fmax(fabs(e2), fabs(e3))));
    dev = fmax(dev, fabs(v[6][i] - (v[0][i] + e1 + e2)));
    dev = fmax(dev, fabs(v[5][i] - (v[0][i] + e1 + e3)));
    dev = fmax(dev, fabs(v[3][i] - (v[0][i] + e2 + e3)));
    dev = fmax(dev, fabs(v[7][i] - (v[0][i] + e1 + e2 + e3)));
  }
  if (!(dev <= tol * scale))
    atomicAdd(n_nonaffine, 1);
  double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
  for (int k = 0; k < 8; ++k)
  {
    const int a = (k >> 2) & 1, b = (k >> 1) & 1, c = k & 1;
    const double dphi[3] = {(a ? 1.0 : -1.0) * 0.25, (b ? 1.0 : -1.0) * 0.25, (c ? 1.0 : -1.0) * 0.25};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        J[i][j] += v[k][i] * dphi[j];
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
  const double detJ = literal_detj ? J[0][0] * K[0][0] - J[1][0] * K[0][1] + J[0][2] * K[2][0]
                                   : J[0][0] * K[0][0] + J[0][1] * K[1][0] + J[0][2] * K[2][0];
  const double s = 1.0 / detJ;
  double* g = Gc + (size_t)p * 6;
  g[0] = (K[0][0] * K[0][0] + K[0][1] * K[0][1] + K[0][2] * K[0][2]) * s;
  g[1] = (K[1][0] * K[0][0] + K[1][1] * K[0][1] + K[1][2] * K[0][2]) * s;
  g[2] = (K[2][0] * K[0][0] + K[2][1] * K[0][1] + K[2][2] * K[0][2]) * s;
  g[3] = (K[1][0] * K[1][0] + K[1][1] * K[1][1] + K[1][2] * K[1][2]) * s;
  g[4] = (K[2][0] * K[1][0] + K[2][1] * K[1][1] + K[2][2] * K[1][2]) * s;
  g[5] = (K[2][0] * K[2][0] + K[2][1] * K[2][1] + K[2][2] * K[2][2]) * s;
}

// diag(A) contributions, thread per (cell position, local dof); see DESIGN.md "diagonal".
__global__ void k_diag(int P, const double* __restrict__ G, const int32_t* __restrict__ enc,
                       const int32_t* __restrict__ perm, const double* __restrict__ kappa,
                       double* __restrict__ diag, int n_list, Lay lay)
{
  const int n = P + 1, n2 = n * n, n3 = n2 * n;
  const long long total = (long long)n_list * n3;
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += nth)
  {
    const long long p = t / n3;
    const int a = (int)(t - p * n3);
    const int32_t d = enc[lay.enc_index(p, a)];
    if (d < 0)
      continue;
    const int i = a / n2, j = (a / n) % n, k = a % n;
    const double* D = c_D[P];
    double s = 0.0;
    for (int q = 0; q < n; ++q)
    {
      const double dx = D[q * n + i], dy = D[q * n + j], dz = D[q * n + k];
      s += dx * dx * G[lay.g_index(p, 0, q * n2 + j * n + k)];
      s += dy * dy * G[lay.g_index(p, 3, i * n2 + q * n + k)];
      s += dz * dz * G[lay.g_index(p, 5, i * n2 + j * n + q)];
    }
    const double dii = D[i * n + i], djj = D[j * n + j], dkk = D[k * n + k];
    s += 2.0 * (G[lay.g_index(p, 1, a)] * dii * djj + G[lay.g_index(p, 2, a)] * dii * dkk
                + G[lay.g_index(p, 4, a)] * djj * dkk);
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
                                        int n3, long long total, Lay lay)
{
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += nth)
  {
    const long long p = t / (6 * n3);
    const int r = (int)(t - p * 6 * n3);
    const int q = r / 6, c = r % 6;
    out[t] = G[lay.g_index(p, c, q)];
  }
}

__global__ void k_rhs(const double* __restrict__ detj_w, const int32_t* __restrict__ enc,
                      const double* __restrict__ fvals, double* __restrict__ b, long long total,
                      int n3, Lay lay)
{
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += nth)
  {
    const int32_t d = enc[lay.enc_index(t / n3, (int)(t % n3))];
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

// gbc = bc ? g : 0
__global__ void k_mask_to_bc(const double* __restrict__ g, const int8_t* __restrict__ bc, double* __restrict__ gbc, int n)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    gbc[i] = bc[i] ? g[i] : 0.0;
}

// b = bc ? g : b - y   (lifting on the free rows, set_bc on the marked ones)
__global__ void k_lift(double* __restrict__ b, const double* __restrict__ y, const double* __restrict__ g,
                       const int8_t* __restrict__ bc, int n)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    b[i] = bc[i] ? g[i] : y[i] * (-1.0) + b[i];
}

__global__ void k_fill(double* __restrict__ v, double a, int n)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    v[i] = a;
}

// ------------------------------------------------------------- CSR assembly kernels --
// Entry (i,j) of the element matrix kappa * B^T G B with the collocated gradient table
// (phi = identity at the GLL points, src/laplacian.hpp:200-202): only index pairs that
// agree in at least one direction couple.
__device__ double element_entry(int P, const double* __restrict__ G, const Lay& lay, long long p, int i, int j)
{
  const int n = P + 1, n2 = n * n, n3 = n2 * n;
  const double* D = c_D[P];
  const int ix = i / n2, iy = (i / n) % n, iz = i % n;
  const int jx = j / n2, jy = (j / n) % n, jz = j % n;
  double s = 0.0;
  if (iy == jy && iz == jz)
    for (int q = 0; q < n; ++q)
      s += D[q * n + ix] * D[q * n + jx] * G[lay.g_index(p, 0, q * n2 + iy * n + iz)];
  if (ix == jx && iz == jz)
    for (int q = 0; q < n; ++q)
      s += D[q * n + iy] * D[q * n + jy] * G[lay.g_index(p, 3, ix * n2 + q * n + iz)];
  if (ix == jx && iy == jy)
    for (int q = 0; q < n; ++q)
      s += D[q * n + iz] * D[q * n + jz] * G[lay.g_index(p, 5, ix * n2 + iy * n + q)];
  if (iz == jz)
    s += D[jx * n + ix] * D[iy * n + jy] * G[lay.g_index(p, 1, jx * n2 + iy * n + iz)]
         + D[jy * n + iy] * D[ix * n + jx] * G[lay.g_index(p, 1, ix * n2 + jy * n + iz)];
  if (iy == jy)
    s += D[jx * n + ix] * D[iz * n + jz] * G[lay.g_index(p, 2, jx * n2 + iy * n + iz)]
         + D[jz * n + iz] * D[ix * n + jx] * G[lay.g_index(p, 2, ix * n2 + iy * n + jz)];
  if (ix == jx)
    s += D[jy * n + iy] * D[iz * n + jz] * G[lay.g_index(p, 4, ix * n2 + jy * n + iz)]
         + D[jz * n + iz] * D[iy * n + jy] * G[lay.g_index(p, 4, ix * n2 + iy * n + jz)];
  return s;
}

// One (key, value) per (cell, i, j); key = row * ntot + col, sentinel for dropped entries
// (ghost rows, Dirichlet rows / columns: fem::assemble_matrix with bcs, src/csr.hpp:84).
__global__ void k_element_triplets(int P, const double* __restrict__ G, const int32_t* __restrict__ enc,
                                   const int32_t* __restrict__ perm, const double* __restrict__ kappa,
                                   long long n_list, int n_owned, long long ntot,
                                   unsigned long long* __restrict__ keys, double* __restrict__ vals, Lay lay)
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
    const int32_t di = enc[lay.enc_index(p, i)], dj = enc[lay.enc_index(p, j)];
    if (di < 0 || dj < 0 || di >= n_owned)
    {
      keys[t] = ~0ull;
      vals[t] = 0.0;
      continue;
    }
    keys[t] = (unsigned long long)di * (unsigned long long)ntot + (unsigned long long)dj;
    vals[t] = kappa[perm[p]] * element_entry(P, G, lay, p, i, j);
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
void launch_apply(pmgx_ctx* c, cudaStream_t st, const double* x, double* y, const double* G, const int32_t* enc,
                  const int32_t* perm, const double* kappa, int first, int count)
{
  const bool timed = c->profiling && first == 0; // per-kernel timing covers the interior-cell launch only
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
  if (timed)
  {
    PMGX_CUDA(cudaEventCreate(&e0));
    PMGX_CUDA(cudaEventCreate(&e1));
    PMGX_CUDA(cudaEventRecord(e0, st));
  }
  k_apply<P><<<grid, C::tpb, smem, st>>>(x, y, G, enc, perm, kappa, first, count);
  check_launch("k_apply");
  count_launch(c);
  if (timed)
  {
    PMGX_CUDA(cudaEventRecord(e1, st));
    c->prof[P].emplace_back(e0, e1);
  }
}

// ------------------------------------------------------------- the slab apply kernel --
// Thread = one z-index k of one cell; it keeps the whole (ix,iy) slab of the element in
// registers, so the x and y contractions are pure register FMAs with compile-time D entries
// and only the z contraction exchanges data through shared memory (rows of n values that all
// n threads of a cell read as a broadcast).  This moves ~1/3 of the bytes of the column
// kernel through the LSU/shared-memory pipe, which ncu showed to be the limiter there
// (l1tex data-pipe 80 % busy at 61 % DRAM, profiles/r1_apply_p4_column.txt).
template <int P>
struct SlabCfg
{
  static constexpr int n = P + 1;
  static constexpr int n2 = n * n;
  static constexpr int tpb = 128;
  static constexpr int cpb = tpb / n;      // cells per block
  static constexpr int S = (cpb * n + 1) & ~1; // slots per (batch, ix, iy) row of G
  static constexpr int SE = (cpb * n + 3) & ~3; // ... and of the dofmap
  static constexpr int kp = n + (n & 1);   // padded z-row (16-byte rows for double2 loads)
  static constexpr int su_doubles = cpb * n2 * kp;
  static constexpr int sf_doubles = 2 * cpb * n * kp;
  static constexpr size_t smem = (size_t)(su_doubles + sf_doubles) * sizeof(double);
};

template <int P, int MINB>
__global__ void __launch_bounds__(SlabCfg<P>::tpb, MINB)
k_apply_slab(const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ G,
             const int32_t* __restrict__ enc, const int32_t* __restrict__ perm,
             const double* __restrict__ kappa, long long batch0, int cell0, int count)
{
  using C = SlabCfg<P>;
  constexpr int n = C::n, n2 = C::n2, CPB = C::cpb, S = C::S, SE = C::SE, KP = C::kp;
  extern __shared__ __align__(16) double smem[];
  double* su = smem;                  // [CPB][n2][KP]
  double* sf = smem + C::su_doubles;  // [2][CPB][n][KP]

  const int tid = threadIdx.x;
  const int cl = tid / n;
  const int k = tid - cl * n;
  const int pl = blockIdx.x * CPB + cl;
  const bool in_block = cl < CPB;
  const bool active = in_block && pl < count;
  const long long gb = batch0 + blockIdx.x;
  const int32_t* e = enc + gb * (n2 * SE) + tid;
  const double* g = G + gb * ((long long)n2 * 6 * S) + tid;
  const int cls = in_block ? cl : 0;

  int d[n2];
  double u[n2];
  double kap = 0.0;
  if (active)
  {
#pragma unroll
    for (int a = 0; a < n2; ++a)
      d[a] = ldg_stream_i32(e + a * SE);
    kap = kappa[perm[cell0 + pl]];
#pragma unroll
    for (int a = 0; a < n2; ++a)
    {
      const int idx = d[a] < 0 ? ~d[a] : d[a];
      const double xv = x[idx];
      if (d[a] < 0)
        y[idx] = xv; // Dirichlet row: y = x (src/laplacian.hpp:273-274)
      u[a] = d[a] < 0 ? 0.0 : xv;
    }
  }
  else
  {
#pragma unroll
    for (int a = 0; a < n2; ++a)
      d[a] = -1, u[a] = 0.0;
  }
  if (in_block)
  {
#pragma unroll
    for (int a = 0; a < n2; ++a)
      su[(cl * n2 + a) * KP + k] = u[a];
  }
  double Dk[n], DTk[n];
#pragma unroll
  for (int l = 0; l < n; ++l)
  {
    Dk[l] = c_D[P][k * n + l];
    DTk[l] = c_D[P][l * n + k];
  }
  double acc[n2];
#pragma unroll
  for (int a = 0; a < n2; ++a)
    acc[a] = 0.0;
  __syncthreads();

#pragma unroll
  for (int i = 0; i < n; ++i)
  {
    double gg[n][6];
    if (active)
    {
#pragma unroll
      for (int j = 0; j < n; ++j)
#pragma unroll
        for (int c = 0; c < 6; ++c)
          gg[j][c] = ldg_stream(g + ((i * n + j) * 6 + c) * S);
    }
    else
    {
#pragma unroll
      for (int j = 0; j < n; ++j)
#pragma unroll
        for (int c = 0; c < 6; ++c)
          gg[j][c] = 0.0;
    }
    double fz[n];
#pragma unroll
    for (int j = 0; j < n; ++j)
    {
      double gx = 0.0, gy = 0.0, gz = 0.0;
      const double2* row = reinterpret_cast<const double2*>(su + (cls * n2 + i * n + j) * KP);
#pragma unroll
      for (int l2 = 0; l2 < KP / 2; ++l2)
      {
        const double2 r = row[l2];
        gz = fma(Dk[2 * l2], r.x, gz);
        if (2 * l2 + 1 < n)
          gz = fma(Dk[2 * l2 + 1], r.y, gz);
      }
#pragma unroll
      for (int l = 0; l < n; ++l)
      {
        gx = fma(c_D[P][i * n + l], u[l * n + j], gx);
        gy = fma(c_D[P][j * n + l], u[i * n + l], gy);
      }
      const double fx = kap * (gg[j][0] * gx + gg[j][1] * gy + gg[j][2] * gz);
      const double fy = kap * (gg[j][1] * gx + gg[j][3] * gy + gg[j][4] * gz);
      fz[j] = kap * (gg[j][2] * gx + gg[j][4] * gy + gg[j][5] * gz);
#pragma unroll
      for (int l = 0; l < n; ++l)
      {
        acc[l * n + j] = fma(c_D[P][i * n + l], fx, acc[l * n + j]);
        acc[i * n + l] = fma(c_D[P][j * n + l], fy, acc[i * n + l]);
      }
    }
    double* sfz = sf + (((i & 1) * CPB + cls) * n) * KP;
    if (in_block)
    {
#pragma unroll
      for (int j = 0; j < n; ++j)
        sfz[j * KP + k] = fz[j];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < n; ++j)
    {
      const double2* row = reinterpret_cast<const double2*>(sfz + j * KP);
      double t = 0.0;
#pragma unroll
      for (int q2 = 0; q2 < KP / 2; ++q2)
      {
        const double2 r = row[q2];
        t = fma(DTk[2 * q2], r.x, t);
        if (2 * q2 + 1 < n)
          t = fma(DTk[2 * q2 + 1], r.y, t);
      }
      acc[i * n + j] += t;
    }
  }
  if (active)
  {
#pragma unroll
    for (int a = 0; a < n2; ++a)
      if (d[a] >= 0)
        atomicAdd(&y[d[a]], acc[a]);
  }
}

template <int P>
constexpr int slab_minb()
{
  return P <= 2 ? 4 : (P == 3 ? 3 : 2);
}

template <int P>
void launch_apply_slab(pmgx_ctx* c, cudaStream_t st, const double* x, double* y, const double* G, const int32_t* enc,
                       const int32_t* perm, const double* kappa, long long batch0, int cell0, int count)
{
  const bool timed = c->profiling && cell0 == 0; // per-kernel timing covers the interior-cell launch only
  if (count <= 0)
    return;
  using C = SlabCfg<P>;
  constexpr int MINB = slab_minb<P>();
  static bool configured[64] = {false};
  if (!configured[c->device])
  {
    PMGX_CUDA(cudaFuncSetAttribute(k_apply_slab<P, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)C::smem));
    configured[c->device] = true;
  }
  const int grid = (count + C::cpb - 1) / C::cpb;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (timed)
  {
    PMGX_CUDA(cudaEventCreate(&e0));
    PMGX_CUDA(cudaEventCreate(&e1));
    PMGX_CUDA(cudaEventRecord(e0, st));
  }
  k_apply_slab<P, MINB><<<grid, C::tpb, C::smem, st>>>(x, y, G, enc, perm, kappa, batch0, cell0, count);
  check_launch("k_apply_slab");
  count_launch(c);
  if (timed)
  {
    PMGX_CUDA(cudaEventRecord(e1, st));
    c->prof[P].emplace_back(e0, e1);
  }
}

// --------------------------------------------- the TMA-pipelined slab apply kernel --
// Same thread mapping and arithmetic as k_apply_slab, but the one-shot streams (geometry
// factors, encoded dofmap) no longer pass through registers with the latency exposed to a
// handful of resident warps: a persistent CTA walks over its batches and an elected thread
// keeps a ring of R geometry planes (and the next batch's dofmap) in flight with
// cp.async.bulk (TMA) into shared memory, completion tracked by mbarriers.  Memory-level
// parallelism is then set by the ring depth, not by occupancy or register count.
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile("{\n"
               ".reg .pred p;\n"
               "WAIT_LOOP:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra WAIT_DONE;\n"
               "bra WAIT_LOOP;\n"
               "WAIT_DONE:\n"
               "}\n" ::"r"(smem_u32(bar)),
               "r"(parity)
               : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
               "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}

template <int P, int TPB, int R>
struct TmaCfg
{
  static constexpr int n = P + 1;
  static constexpr int n2 = n * n;
  static constexpr int tpb = TPB;
  static constexpr int cpb = tpb / n;
  static constexpr int S = (cpb * n + 1) & ~1;  // G rows: 16-byte multiples, (almost) no padding streamed
  static constexpr int SE = (cpb * n + 3) & ~3; // dofmap rows: 16-byte multiples of int32
  static constexpr int minb = P == 1 ? 4 : (P == 2 ? 3 : (TPB <= 64 ? 4 : 2)); // resident CTAs per SM
  static constexpr int kp = n + (n & 1);
  // per-cell stride of a z-row plane buffer: an odd number of 16-byte chunks, so the cells of
  // a warp start in different bank groups (profiles/r1_apply_p3_slab.txt: 75 % of the shared
  // wavefronts were conflict replays with the unpadded stride)
  static constexpr int cs = ((n * kp / 2) & 1) ? n * kp : n * kp + 2;
  static constexpr int plane_doubles = n * 6 * S;
  static constexpr uint32_t plane_bytes = plane_doubles * sizeof(double);
  static constexpr uint32_t enc_bytes = n2 * SE * sizeof(int32_t);
  static constexpr int buf_doubles = 2 * cpb * cs; // double-buffered plane rows (su and sf each)
  static constexpr size_t off_enc = (size_t)R * plane_bytes;
  // dofmap buffers: double-buffered (read again at scatter time, next batch prefetched meanwhile);
  // a single buffer at P >= 6, where the saved 12.5 KB let a third CTA (+50 % warps) fit on the SM
  static constexpr int EB = P >= 6 ? 1 : 2;
  static constexpr size_t off_su = off_enc + EB * enc_bytes;
  static constexpr size_t off_sf = off_su + (size_t)buf_doubles * sizeof(double);
  static constexpr size_t off_bar = off_sf + (size_t)buf_doubles * sizeof(double);
  static constexpr size_t smem = off_bar + (R + 2) * sizeof(uint64_t);
};

template <int P, int TPB, int R>
__global__ void __launch_bounds__(TPB, TmaCfg<P, TPB, R>::minb)
k_apply_tma(const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ G,
            const int32_t* __restrict__ enc, const int32_t* __restrict__ perm,
            const double* __restrict__ kappa, long long batch0, int cell0, int count, int nbatch)
{
  using C = TmaCfg<P, TPB, R>;
  constexpr int n = C::n, n2 = C::n2, CPB = C::cpb, S = C::S, SE = C::SE, KP = C::kp, CS = C::cs;
  extern __shared__ __align__(128) unsigned char smraw[];
  double* sG = reinterpret_cast<double*>(smraw);
  const int32_t* sE = reinterpret_cast<const int32_t*>(smraw + C::off_enc);
  double* su = reinterpret_cast<double*>(smraw + C::off_su);
  double* sf = reinterpret_cast<double*>(smraw + C::off_sf);
  uint64_t* fullG = reinterpret_cast<uint64_t*>(smraw + C::off_bar);
  uint64_t* fullE = fullG + R;

  const int tid = threadIdx.x;
  const int cl = tid / n;
  const int k = tid - cl * n;
  const bool in_block = cl < CPB;
  const int cls = in_block ? cl : 0;
  const int my_nb = ((int)blockIdx.x < nbatch) ? (nbatch - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int total_planes = my_nb * n;

  uint64_t pol = 0;
  if (tid == 0)
  {
    for (int s = 0; s < R + 2; ++s)
      mbar_init(&fullG[s], 1);
    mbar_fence_init();
    pol = policy_evict_first();
  }
  __syncthreads();

  // producer state (thread 0 only): next plane to issue
  int issue_t = 0;
  auto issue_plane = [&](int t)
  {
    const int it = t / n, i = t - it * n;
    const long long gb = batch0 + blockIdx.x + (long long)it * gridDim.x;
    const int s = t % R;
    mbar_expect_tx(&fullG[s], C::plane_bytes);
    bulk_g2s(sG + (size_t)s * C::plane_doubles, G + (gb * n2 + (long long)i * n) * 6 * S, C::plane_bytes,
             &fullG[s], pol);
  };
  constexpr int EB = C::EB;
  auto issue_enc = [&](int it)
  {
    const long long gb = batch0 + blockIdx.x + (long long)it * gridDim.x;
    const int eb = EB == 2 ? (it & 1) : 0;
    mbar_expect_tx(&fullE[eb], C::enc_bytes);
    bulk_g2s(const_cast<int32_t*>(sE) + eb * (n2 * SE), enc + gb * (long long)(n2 * SE), C::enc_bytes, &fullE[eb], pol);
  };
  if (tid == 0 && my_nb > 0)
  {
    issue_enc(0);
    for (; issue_t < R && issue_t < total_planes; ++issue_t)
      issue_plane(issue_t);
  }

  double Dk[n], DTk[n];
#pragma unroll
  for (int l = 0; l < n; ++l)
  {
    Dk[l] = c_D[P][k * n + l];
    DTk[l] = c_D[P][l * n + k];
  }

  int stage = 0;
  uint32_t stage_parity = 0;
  for (int it = 0; it < my_nb; ++it)
  {
    const int b = blockIdx.x + it * gridDim.x;
    const int pl = b * CPB + cl;
    const bool active = in_block && pl < count;

    const int eb = EB == 2 ? (it & 1) : 0;
    mbar_wait(&fullE[eb], EB == 2 ? (it >> 1) & 1 : it & 1);
    const int32_t* dE = sE + eb * (n2 * SE) + tid; // this thread's n2 encoded dofs
    double u[n2];
    double kap = 0.0;
    if (active)
    {
      kap = kappa[perm[cell0 + pl]];
#pragma unroll
      for (int a = 0; a < n2; ++a)
      {
        const int da = dE[a * SE];
        const int idx = da < 0 ? ~da : da;
        const double xv = x[idx];
        if (da < 0)
          y[idx] = xv; // Dirichlet row: y = x (src/laplacian.hpp:273-274)
        u[a] = da < 0 ? 0.0 : xv;
      }
    }
    else
    {
#pragma unroll
      for (int a = 0; a < n2; ++a)
        u[a] = 0.0;
    }
    if (in_block)
    {
#pragma unroll
      for (int j = 0; j < n; ++j)
        su[cl * CS + j * KP + k] = u[j];
    }
    __syncthreads(); // z-rows of plane 0 are visible; the other dofmap buffer (batch it-1) is free
    if (EB == 2 && tid == 0 && it + 1 < my_nb)
      issue_enc(it + 1);

    double acc[n2];
#pragma unroll
    for (int a = 0; a < n2; ++a)
      acc[a] = 0.0;

#pragma unroll
    for (int i = 0; i < n; ++i)
    {
      mbar_wait(&fullG[stage], stage_parity);
      const double* gs = sG + (size_t)stage * C::plane_doubles + tid;
      const double* sup = su + ((i & 1) * CPB + cls) * CS;
      double* sfz = sf + ((i & 1) * CPB + cls) * CS;
      double fz[n];
#pragma unroll
      for (int j = 0; j < n; ++j)
      {
        double gx = 0.0, gy = 0.0, gz = 0.0;
        const double2* row = reinterpret_cast<const double2*>(sup + j * KP);
#pragma unroll
        for (int l2 = 0; l2 < KP / 2; ++l2)
        {
          const double2 r = row[l2];
          gz = fma(Dk[2 * l2], r.x, gz);
          if (2 * l2 + 1 < n)
            gz = fma(Dk[2 * l2 + 1], r.y, gz);
        }
#pragma unroll
        for (int l = 0; l < n; ++l)
        {
          gx = fma(c_D[P][i * n + l], u[l * n + j], gx);
          gy = fma(c_D[P][j * n + l], u[i * n + l], gy);
        }
        const double g0 = gs[(j * 6 + 0) * S], g1 = gs[(j * 6 + 1) * S], g2 = gs[(j * 6 + 2) * S];
        const double g3 = gs[(j * 6 + 3) * S], g4 = gs[(j * 6 + 4) * S], g5 = gs[(j * 6 + 5) * S];
        const double fx = kap * (g0 * gx + g1 * gy + g2 * gz);
        const double fy = kap * (g1 * gx + g3 * gy + g4 * gz);
        fz[j] = kap * (g2 * gx + g4 * gy + g5 * gz);
#pragma unroll
        for (int l = 0; l < n; ++l)
        {
          acc[l * n + j] = fma(c_D[P][i * n + l], fx, acc[l * n + j]);
          acc[i * n + l] = fma(c_D[P][j * n + l], fy, acc[i * n + l]);
        }
      }
      if (in_block)
      {
#pragma unroll
        for (int j = 0; j < n; ++j)
          sfz[j * KP + k] = fz[j];
        if (i + 1 < n)
        {
          double* sun = su + (((i + 1) & 1) * CPB + cl) * CS;
#pragma unroll
          for (int j = 0; j < n; ++j)
            sun[j * KP + k] = u[(i + 1) * n + j];
        }
      }
      __syncthreads(); // geometry stage is consumed; fz rows and next z-rows are visible
      if (tid == 0)
      {
        if (issue_t < total_planes)
          issue_plane(issue_t++);
      }
      if (++stage == R)
      {
        stage = 0;
        stage_parity ^= 1u;
      }
#pragma unroll
      for (int j = 0; j < n; ++j)
      {
        const double2* row = reinterpret_cast<const double2*>(sfz + j * KP);
        double t = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < KP / 2; ++q2)
        {
          const double2 r = row[q2];
          t = fma(DTk[2 * q2], r.x, t);
          if (2 * q2 + 1 < n)
            t = fma(DTk[2 * q2 + 1], r.y, t);
        }
        acc[i * n + j] += t;
      }
    }
    if (active)
    {
#pragma unroll
      for (int a = 0; a < n2; ++a)
      {
        const int da = dE[a * SE];
        if (da >= 0)
          atomicAdd(&y[da], acc[a]);
      }
    }
    if (EB == 1 && it + 1 < my_nb)
    {
      __syncthreads(); // every thread has read its scatter indices: the single dofmap buffer is free
      if (tid == 0)
        issue_enc(it + 1);
    }
  }
}

template <int P, int TPB, int R>
void launch_apply_tma_t(pmgx_ctx* c, cudaStream_t st, const double* x, double* y, const double* G, const int32_t* enc,
                        const int32_t* perm, const double* kappa, long long batch0, int cell0, int count)
{
  const bool timed = c->profiling && cell0 == 0; // per-kernel timing covers the interior-cell launch only
  using C = TmaCfg<P, TPB, R>;
  static int ctas_per_sm[64] = {0};
  if (ctas_per_sm[c->device] == 0)
  {
    PMGX_CUDA(cudaFuncSetAttribute(k_apply_tma<P, TPB, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)C::smem));
    int nb = 0;
    PMGX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_apply_tma<P, TPB, R>, TPB, C::smem));
    PMGX_REQUIRE(nb >= 1, "k_apply_tma<%d,%d,%d> does not fit on an SM", P, TPB, R);
    ctas_per_sm[c->device] = nb;
  }
  const int nbatch = (count + C::cpb - 1) / C::cpb;
  // persistent CTAs: exactly the number that is co-resident, so every SM streams all the time
  const int grid = std::min(nbatch, ctas_per_sm[c->device] * c->num_sms);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (timed)
  {
    PMGX_CUDA(cudaEventCreate(&e0));
    PMGX_CUDA(cudaEventCreate(&e1));
    PMGX_CUDA(cudaEventRecord(e0, st));
  }
  k_apply_tma<P, TPB, R><<<grid, TPB, C::smem, st>>>(x, y, G, enc, perm, kappa, batch0, cell0, count, nbatch);
  check_launch("k_apply_tma");
  count_launch(c);
  if (timed)
  {
    PMGX_CUDA(cudaEventRecord(e1, st));
    c->prof[P].emplace_back(e0, e1);
  }
}

// ------------------------------------------------ the affine-geometry apply kernel --
// On a cell whose trilinear map is affine (every cell of a box / parallelepiped mesh, i.e. all
// the benchmark meshes) J is constant, so G(q) = w_q * Gc with ONE 6-vector Gc = K K^T / det J per
// cell: the 48 B per quadrature point that make up 79 % of the streamed bytes of the apply collapse
// to 48 B per cell.  When ALL cells of an operator pass the affinity test at create
// (k_cell_geometry, tolerance 1e-13 h) this kernel replaces k_apply_tma: same thread mapping and
// arithmetic (thread = z-index of a cell, (ix,iy) slab in registers, z contraction through
// shared memory), but nothing is streamed except the encoded dofmap (cp.async.bulk double
// buffer); the kernel is then bound by the gather / atomic traffic of x and y in L2, not by HBM.
// Meshes with any non-affine cell keep the streamed-G kernels; PMGX_LAP_STREAM_G forces them.
template <int P, int TPB>
struct AffCfg
{
  static constexpr int n = P + 1;
  static constexpr int n2 = n * n;
  static constexpr int tpb = TPB;
  static constexpr int cpb = tpb / n;
  static constexpr int SE = (cpb * n + 3) & ~3;
  static constexpr int kp = n + (n & 1);
  static constexpr int cs = ((n * kp / 2) & 1) ? n * kp : n * kp + 2;
  static constexpr uint32_t enc_bytes = n2 * SE * sizeof(int32_t);
  static constexpr int buf_doubles = 2 * cpb * cs;
  static constexpr size_t off_su = 2 * (size_t)enc_bytes;
  static constexpr size_t off_sf = off_su + (size_t)buf_doubles * sizeof(double);
  static constexpr size_t off_bar = off_sf + (size_t)buf_doubles * sizeof(double);
  static constexpr size_t smem = off_bar + 2 * sizeof(uint64_t);
#ifndef PMGX_AFF_MINB_P4
#define PMGX_AFF_MINB_P4 4
#endif
  static constexpr int minb = P <= 2 ? 6 : (TPB <= 64 ? (P == 4 ? PMGX_AFF_MINB_P4 : 4) : 2);
};

template <int P, int TPB>
__global__ void __launch_bounds__(TPB, AffCfg<P, TPB>::minb)
k_apply_affine(const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ Gc,
               const int32_t* __restrict__ enc, const int32_t* __restrict__ perm,
               const double* __restrict__ kappa, long long batch0, int cell0, int count, int nbatch)
{
  using C = AffCfg<P, TPB>;
  constexpr int n = C::n, n2 = C::n2, CPB = C::cpb, SE = C::SE, KP = C::kp, CS = C::cs;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int32_t* sE = reinterpret_cast<const int32_t*>(smraw);
  double* su = reinterpret_cast<double*>(smraw + C::off_su);
  double* sf = reinterpret_cast<double*>(smraw + C::off_sf);
  uint64_t* fullE = reinterpret_cast<uint64_t*>(smraw + C::off_bar);

  const int tid = threadIdx.x;
  const int cl = tid / n;
  const int k = tid - cl * n;
  const bool in_block = cl < CPB;
  const int cls = in_block ? cl : 0;
  const int my_nb = ((int)blockIdx.x < nbatch) ? (nbatch - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  uint64_t pol = 0;
  if (tid == 0)
  {
    mbar_init(&fullE[0], 1);
    mbar_init(&fullE[1], 1);
    mbar_fence_init();
    pol = policy_evict_first();
  }
  __syncthreads();
  auto issue_enc = [&](int it)
  {
    const long long gb = batch0 + blockIdx.x + (long long)it * gridDim.x;
    mbar_expect_tx(&fullE[it & 1], C::enc_bytes);
    bulk_g2s(const_cast<int32_t*>(sE) + (it & 1) * (n2 * SE), enc + gb * (long long)(n2 * SE), C::enc_bytes,
             &fullE[it & 1], pol);
  };
  if (tid == 0 && my_nb > 0)
    issue_enc(0);

  double Dk[n], DTk[n];
#pragma unroll
  for (int l = 0; l < n; ++l)
  {
    Dk[l] = c_D[P][k * n + l];
    DTk[l] = c_D[P][l * n + k];
  }
  const double wk = c_wts[P][k];

  for (int it = 0; it < my_nb; ++it)
  {
    const int b = blockIdx.x + it * gridDim.x;
    const int pl = b * CPB + cl;
    const bool active = in_block && pl < count;

    mbar_wait(&fullE[it & 1], (it >> 1) & 1);
    const int32_t* dE = sE + (it & 1) * (n2 * SE) + tid;
    double u[n2];
    double g[6];
    if (active)
    {
      const double kw = kappa[perm[cell0 + pl]] * wk; // kappa re-read on every apply (src/laplacian.hpp:230)
      const double* gp = Gc + (size_t)(cell0 + pl) * 6;
#pragma unroll
      for (int c = 0; c < 6; ++c)
        g[c] = gp[c] * kw;
#pragma unroll
      for (int a = 0; a < n2; ++a)
      {
        const int da = dE[a * SE];
        const int idx = da < 0 ? ~da : da;
        const double xv = x[idx];
        if (da < 0)
          y[idx] = xv; // Dirichlet row: y = x (src/laplacian.hpp:273-274)
        u[a] = da < 0 ? 0.0 : xv;
      }
    }
    else
    {
#pragma unroll
      for (int c = 0; c < 6; ++c)
        g[c] = 0.0;
#pragma unroll
      for (int a = 0; a < n2; ++a)
        u[a] = 0.0;
    }
    if (in_block)
    {
#pragma unroll
      for (int j = 0; j < n; ++j)
        su[cl * CS + j * KP + k] = u[j];
    }
    __syncthreads(); // z-rows of plane 0 are visible; the other dofmap buffer (batch it-1) is free
    if (tid == 0 && it + 1 < my_nb)
      issue_enc(it + 1);

    double acc[n2];
#pragma unroll
    for (int a = 0; a < n2; ++a)
      acc[a] = 0.0;

#pragma unroll
    for (int i = 0; i < n; ++i)
    {
      const double* sup = su + ((i & 1) * CPB + cls) * CS;
      double* sfz = sf + ((i & 1) * CPB + cls) * CS;
      double fz[n];
#pragma unroll
      for (int j = 0; j < n; ++j)
      {
        double gx = 0.0, gy = 0.0, gz = 0.0;
        const double2* row = reinterpret_cast<const double2*>(sup + j * KP);
#pragma unroll
        for (int l2 = 0; l2 < n / 2; ++l2)
        {
          const double2 r = row[l2];
          gz = fma(Dk[2 * l2], r.x, gz);
          gz = fma(Dk[2 * l2 + 1], r.y, gz);
        }
        if (n & 1) // odd n: the last entry alone (a 64-bit load: half the wavefronts of a padded 128-bit one)
          gz = fma(Dk[n - 1], sup[j * KP + n - 1], gz);
#pragma unroll
        for (int l = 0; l < n; ++l)
        {
          gx = fma(c_D[P][i * n + l], u[l * n + j], gx);
          gy = fma(c_D[P][j * n + l], u[i * n + l], gy);
        }
        const double wij = c_wts[P][i] * c_wts[P][j]; // G(q) = w_i w_j w_k Gc
        const double fx = wij * (g[0] * gx + g[1] * gy + g[2] * gz);
        const double fy = wij * (g[1] * gx + g[3] * gy + g[4] * gz);
        fz[j] = wij * (g[2] * gx + g[4] * gy + g[5] * gz);
#pragma unroll
        for (int l = 0; l < n; ++l)
        {
          acc[l * n + j] = fma(c_D[P][i * n + l], fx, acc[l * n + j]);
          acc[i * n + l] = fma(c_D[P][j * n + l], fy, acc[i * n + l]);
        }
      }
      if (in_block)
      {
#pragma unroll
        for (int j = 0; j < n; ++j)
          sfz[j * KP + k] = fz[j];
        if (i + 1 < n)
        {
          double* sun = su + (((i + 1) & 1) * CPB + cl) * CS;
#pragma unroll
          for (int j = 0; j < n; ++j)
            sun[j * KP + k] = u[(i + 1) * n + j];
        }
      }
      __syncthreads(); // fz rows and the next z-rows are visible
#pragma unroll
      for (int j = 0; j < n; ++j)
      {
        const double2* row = reinterpret_cast<const double2*>(sfz + j * KP);
        double t = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < n / 2; ++q2)
        {
          const double2 r = row[q2];
          t = fma(DTk[2 * q2], r.x, t);
          t = fma(DTk[2 * q2 + 1], r.y, t);
        }
        if (n & 1)
          t = fma(DTk[n - 1], sfz[j * KP + n - 1], t);
        acc[i * n + j] += t;
      }
    }
    if (active)
    {
#pragma unroll
      for (int a = 0; a < n2; ++a)
      {
        const int da = dE[a * SE];
        if (da >= 0)
          atomicAdd(&y[da], acc[a]);
      }
    }
  }
}

// Shuffle variant of the affine kernel: a warp holds 32/n whole cells (lanes beyond that idle), so
// the z contractions are warp shuffles from the n lanes of the thread's own cell -- no shared-memory
// rows, no block barrier per plane.  ncu on the shared-memory variant at P4: L1/shared data pipe 78 %
// busy with 20 %-efficient broadcast reads; n double shuffles move the same data in ~40 % fewer
// pipe cycles and the warps of a CTA no longer wait for each other.
template <int P, int TPB>
struct AffShCfg
{
  static constexpr int n = P + 1;
  static constexpr int n2 = n * n;
  static constexpr int tpb = TPB;
  static constexpr int cpw = 32 / n;               // cells per warp
  static constexpr int cpb = (TPB / 32) * cpw;     // cells per batch
  static constexpr int SE = (cpb * n + 3) & ~3;
  static constexpr uint32_t enc_bytes = n2 * SE * sizeof(int32_t);
  static constexpr size_t off_bar = 2 * (size_t)enc_bytes;
  static constexpr size_t smem = off_bar + 2 * sizeof(uint64_t);
  static constexpr int minb = P <= 2 ? 6 : (TPB <= 64 ? 4 : 2);
};

template <int P, int TPB>
__global__ void __launch_bounds__(TPB, AffShCfg<P, TPB>::minb)
k_apply_affine_shfl(const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ Gc,
                    const int32_t* __restrict__ enc, const int32_t* __restrict__ perm,
                    const double* __restrict__ kappa, long long batch0, int cell0, int count, int nbatch)
{
  using C = AffShCfg<P, TPB>;
  constexpr int n = C::n, n2 = C::n2, CPB = C::cpb, CPW = C::cpw, SE = C::SE;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int32_t* sE = reinterpret_cast<const int32_t*>(smraw);
  uint64_t* fullE = reinterpret_cast<uint64_t*>(smraw + C::off_bar);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int cw = lane / n;                 // cell within the warp
  const int k = lane - cw * n;             // iz
  const bool lane_ok = cw < CPW;
  const int cl = warp * CPW + (lane_ok ? cw : 0);
  const int base = (lane_ok ? cw : 0) * n; // first lane of this thread's cell
  const int slot = cl * n + (lane_ok ? k : 0);
  const int my_nb = ((int)blockIdx.x < nbatch) ? (nbatch - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  uint64_t pol = 0;
  if (tid == 0)
  {
    mbar_init(&fullE[0], 1);
    mbar_init(&fullE[1], 1);
    mbar_fence_init();
    pol = policy_evict_first();
  }
  __syncthreads();
  auto issue_enc = [&](int it)
  {
    const long long gb = batch0 + blockIdx.x + (long long)it * gridDim.x;
    mbar_expect_tx(&fullE[it & 1], C::enc_bytes);
    bulk_g2s(const_cast<int32_t*>(sE) + (it & 1) * (n2 * SE), enc + gb * (long long)(n2 * SE), C::enc_bytes,
             &fullE[it & 1], pol);
  };
  if (tid == 0 && my_nb > 0)
  {
    issue_enc(0);
    if (my_nb > 1)
      issue_enc(1);
  }

  const int kk = lane_ok ? k : 0;
  double Dk[n], DTk[n];
#pragma unroll
  for (int l = 0; l < n; ++l)
  {
    Dk[l] = c_D[P][kk * n + l];
    DTk[l] = c_D[P][l * n + kk];
  }
  const double wk = c_wts[P][kk];

  for (int it = 0; it < my_nb; ++it)
  {
    const int b = blockIdx.x + it * gridDim.x;
    const int pl = b * CPB + cl;
    const bool active = lane_ok && pl < count;

    mbar_wait(&fullE[it & 1], (it >> 1) & 1);
    const int32_t* dE = sE + (it & 1) * (n2 * SE) + slot;
    double u[n2];
    double g[6];
    if (active)
    {
      const double kw = kappa[perm[cell0 + pl]] * wk; // kappa re-read on every apply (src/laplacian.hpp:230)
      const double* gp = Gc + (size_t)(cell0 + pl) * 6;
#pragma unroll
      for (int c = 0; c < 6; ++c)
        g[c] = gp[c] * kw;
#pragma unroll
      for (int a = 0; a < n2; ++a)
      {
        const int da = dE[a * SE];
        const int idx = da < 0 ? ~da : da;
        const double xv = x[idx];
        if (da < 0)
          y[idx] = xv; // Dirichlet row: y = x (src/laplacian.hpp:273-274)
        u[a] = da < 0 ? 0.0 : xv;
      }
    }
    else
    {
#pragma unroll
      for (int c = 0; c < 6; ++c)
        g[c] = 0.0;
#pragma unroll
      for (int a = 0; a < n2; ++a)
        u[a] = 0.0;
    }
    double acc[n2];
#pragma unroll
    for (int a = 0; a < n2; ++a)
      acc[a] = 0.0;

#pragma unroll
    for (int i = 0; i < n; ++i)
    {
#pragma unroll
      for (int j = 0; j < n; ++j)
      {
        double gx = 0.0, gy = 0.0, gz = 0.0;
        const double uij = u[i * n + j];
#pragma unroll
        for (int l = 0; l < n; ++l)
        {
          gz = fma(Dk[l], __shfl_sync(0xffffffffu, uij, base + l), gz);
          gx = fma(c_D[P][i * n + l], u[l * n + j], gx);
          gy = fma(c_D[P][j * n + l], u[i * n + l], gy);
        }
        const double wij = c_wts[P][i] * c_wts[P][j]; // G(q) = w_i w_j w_k Gc
        const double fx = wij * (g[0] * gx + g[1] * gy + g[2] * gz);
        const double fy = wij * (g[1] * gx + g[3] * gy + g[4] * gz);
        const double fz = wij * (g[2] * gx + g[4] * gy + g[5] * gz);
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < n; ++q)
          t = fma(DTk[q], __shfl_sync(0xffffffffu, fz, base + q), t);
#pragma unroll
        for (int l = 0; l < n; ++l)
        {
          acc[l * n + j] = fma(c_D[P][i * n + l], fx, acc[l * n + j]);
          acc[i * n + l] = fma(c_D[P][j * n + l], fy, acc[i * n + l]);
        }
        acc[i * n + j] += t;
      }
    }
    if (active)
    {
#pragma unroll
      for (int a = 0; a < n2; ++a)
      {
        const int da = dE[a * SE];
        if (da >= 0)
          atomicAdd(&y[da], acc[a]);
      }
    }
    // every thread is done with this batch's dofmap buffer before the one after next may land in it
    __syncthreads();
    if (tid == 0 && it + 2 < my_nb)
      issue_enc(it + 2);
  }
}

template <int P, int TPB>
void launch_apply_affine_shfl_t(pmgx_ctx* c, cudaStream_t st, const double* x, double* y, const double* Gc,
                                const int32_t* enc, const int32_t* perm, const double* kappa, long long batch0,
                                int cell0, int count)
{
  using C = AffShCfg<P, TPB>;
  const bool timed = c->profiling && cell0 == 0; // per-kernel timing covers the interior-cell launch only
  static int ctas_per_sm[64] = {0};
  if (ctas_per_sm[c->device] == 0)
  {
    PMGX_CUDA(cudaFuncSetAttribute(k_apply_affine_shfl<P, TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)C::smem));
    int nb = 0;
    PMGX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_apply_affine_shfl<P, TPB>, TPB, C::smem));
    PMGX_REQUIRE(nb >= 1, "k_apply_affine_shfl<%d,%d> does not fit on an SM", P, TPB);
    ctas_per_sm[c->device] = nb;
  }
  const int nbatch = (count + C::cpb - 1) / C::cpb;
  const int grid = std::min(nbatch, ctas_per_sm[c->device] * c->num_sms);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (timed)
  {
    PMGX_CUDA(cudaEventCreate(&e0));
    PMGX_CUDA(cudaEventCreate(&e1));
    PMGX_CUDA(cudaEventRecord(e0, st));
  }
  k_apply_affine_shfl<P, TPB><<<grid, TPB, C::smem, st>>>(x, y, Gc, enc, perm, kappa, batch0, cell0, count, nbatch);
  check_launch("k_apply_affine_shfl");
  count_launch(c);
  if (timed)
  {
    PMGX_CUDA(cudaEventRecord(e1, st));
    c->prof[P].emplace_back(e0, e1);
  }
}

template <int P, int TPB>
void launch_apply_affine_t(pmgx_ctx* c, cudaStream_t st, const double* x, double* y, const double* Gc,
                           const int32_t* enc, const int32_t* perm, const double* kappa, long long batch0, int cell0,
                           int count)
{
  using C = AffCfg<P, TPB>;
  const bool timed = c->profiling && cell0 == 0; // per-kernel timing covers the interior-cell launch only
  static int ctas_per_sm[64] = {0};
  if (ctas_per_sm[c->device] == 0)
  {
    PMGX_CUDA(cudaFuncSetAttribute(k_apply_affine<P, TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem));
    int nb = 0;
    PMGX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_apply_affine<P, TPB>, TPB, C::smem));
    PMGX_REQUIRE(nb >= 1, "k_apply_affine<%d,%d> does not fit on an SM", P, TPB);
    ctas_per_sm[c->device] = nb;
  }
  const int nbatch = (count + C::cpb - 1) / C::cpb;
  const int grid = std::min(nbatch, ctas_per_sm[c->device] * c->num_sms);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (timed)
  {
    PMGX_CUDA(cudaEventCreate(&e0));
    PMGX_CUDA(cudaEventCreate(&e1));
    PMGX_CUDA(cudaEventRecord(e0, st));
  }
  k_apply_affine<P, TPB><<<grid, TPB, C::smem, st>>>(x, y, Gc, enc, perm, kappa, batch0, cell0, count, nbatch);
  check_launch("k_apply_affine");
  count_launch(c);
  if (timed)
  {
    PMGX_CUDA(cudaEventRecord(e1, st));
    c->prof[P].emplace_back(e0, e1);
  }
}


// ---------------------------------------- the re-slabbed affine-geometry apply kernel --
// Same thread mapping as k_apply_affine (thread t of a cell holds the (ix,iy) slab at iz = t in
// registers, x and y contractions are register FMAs), but the z direction no longer re-reads every
// z-row of the element n times as a broadcast (2 n^3 + 2 n^2 shared-memory doubles per thread; ncu:
// L1/shared data pipe 78 % busy at 42 % FP64, profiles/r1_apply_p4_affine.txt).  Instead the data is
// RE-SLABBED: every value crosses shared memory once per phase, written by the thread that owns it
// and read by the thread that needs it --
//   1. u(i,j,t)        -> thread j reads the slab u(i,j,.) and forms gz(i,j,l) = sum_m D[l][m] u(i,j,m)
//   2. gz(i,j,l)       -> thread l reads gz(.,.,l): all three gradients at its points; G-transform;
//                         x and y transposed contractions into the register accumulator
//   3. fz(i,j,t)       -> thread j reads fz(i,j,.) and forms az(i,j,l) = sum_q D[q][l] fz(i,j,q)
//   4. az(i,j,l)       -> thread l adds az(.,.,l) to its accumulator
// 8 n^2 doubles per thread instead of 2 n^3 + 2 n^2 (P4: 200 vs 300; P6: 392 vs 784), four barriers
// per batch instead of n + 1, and no broadcast reads.  Strides: slab stride SS = 1 (mod 16) and cell
// stride CS = n (mod 16) doubles make the 8-byte bank of BOTH access patterns (unit stride in t when
// writing, stride SS in t when reading) equal to lane + const: every 64-bit access is conflict-free.
template <int P, int TPB, bool WL>
struct Aff2Cfg
{
  static constexpr int n = P + 1;
  static constexpr int n2 = n * n;
  static constexpr int tpb = TPB;
  // WL (warp-local): a warp holds 32 / n WHOLE cells (the lanes beyond that idle), so every exchange stays
  // inside a warp and the phases are separated by __syncwarp() instead of CTA barriers: the warps of a CTA
  // run free of each other (with 8-10 warps per SM a CTA barrier per phase leaves the schedulers empty)
  static constexpr int cpw = 32 / n;
  static constexpr int cpb = WL ? (TPB / 32) * cpw : TPB / n;
  static constexpr int SE = (cpb * n + 3) & ~3;
  static constexpr int ss = ((n2 - 1 + 15) / 16) * 16 + 1;                 // >= n2, = 1 mod 16
  static constexpr int cs = ((n * ss - n + 15) / 16) * 16 + n;             // >= n * ss, = n mod 16
  static constexpr uint32_t enc_bytes = n2 * SE * sizeof(int32_t);
  static constexpr int buf_doubles = cpb * cs;
  static constexpr size_t off_a = 2 * (size_t)enc_bytes;
  static constexpr size_t off_b = off_a + (size_t)buf_doubles * sizeof(double);
  static constexpr size_t off_bar = off_b + (size_t)buf_doubles * sizeof(double);
  static constexpr size_t smem = off_bar + 2 * sizeof(uint64_t);
  static constexpr int minb = P <= 2 ? 6 : (P == 4 && TPB <= 64 ? PMGX_AFF_MINB_P4 : (P <= 4 ? 4 : 2));
};

template <int P, int TPB, bool WL>
__global__ void __launch_bounds__(TPB, Aff2Cfg<P, TPB, WL>::minb)
k_apply_affine2(const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ Gc,
                const int32_t* __restrict__ enc, const int32_t* __restrict__ perm,
                const double* __restrict__ kappa, long long batch0, int cell0, int count, int nbatch)
{
  using C = Aff2Cfg<P, TPB, WL>;
  constexpr int n = C::n, n2 = C::n2, CPB = C::cpb, SE = C::SE, SS = C::ss, CS = C::cs;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int32_t* sE = reinterpret_cast<const int32_t*>(smraw);
  double* sA = reinterpret_cast<double*>(smraw + C::off_a);
  double* sB = reinterpret_cast<double*>(smraw + C::off_b);
  uint64_t* fullE = reinterpret_cast<uint64_t*>(smraw + C::off_bar);

  const int tid = threadIdx.x;
  int cl, t;
  bool in_block;
  if (WL)
  {
    const int warp = tid >> 5, lane = tid & 31;
    const int cw = lane / n;
    in_block = cw < C::cpw;
    cl = warp * C::cpw + (in_block ? cw : 0);
    t = in_block ? lane - cw * n : 0;
  }
  else
  {
    cl = tid / n;
    t = tid - cl * n;
    in_block = cl < CPB;
  }
  const int cls = in_block ? cl : 0;
  const int slot = WL ? cls * n + t : tid;
  const int my_nb = ((int)blockIdx.x < nbatch) ? (nbatch - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto phase_sync = [&]()
  {
    if (WL)
      __syncwarp();
    else
      __syncthreads();
  };

  uint64_t pol = 0;
  if (tid == 0)
  {
    mbar_init(&fullE[0], 1);
    mbar_init(&fullE[1], 1);
    mbar_fence_init();
    pol = policy_evict_first();
  }
  __syncthreads();
  auto issue_enc = [&](int it)
  {
    const long long gb = batch0 + blockIdx.x + (long long)it * gridDim.x;
    mbar_expect_tx(&fullE[it & 1], C::enc_bytes);
    bulk_g2s(const_cast<int32_t*>(sE) + (it & 1) * (n2 * SE), enc + gb * (long long)(n2 * SE), C::enc_bytes,
             &fullE[it & 1], pol);
  };
  if (tid == 0 && my_nb > 0)
    issue_enc(0);

  const double wt = c_wts[P][t];
  // writer view: element (slab s, row r) of this thread's z-index; reader view: this thread's slab
  double* const wA = sA + cls * CS + t;
  double* const wB = sB + cls * CS + t;
  const double* const rA = sA + cls * CS + t * SS;
  const double* const rB = sB + cls * CS + t * SS;

  for (int it = 0; it < my_nb; ++it)
  {
    const int b = blockIdx.x + it * gridDim.x;
    const int pl = b * CPB + cl;
    const bool active = in_block && pl < count;

    if (WL)
    {
      // the only CTA barrier of a batch: every warp is done with the dofmap buffer the next load will overwrite
      __syncthreads();
      if (tid == 0 && it + 1 < my_nb)
        issue_enc(it + 1);
    }
    mbar_wait(&fullE[it & 1], (it >> 1) & 1);
    const int32_t* dE = sE + (it & 1) * (n2 * SE) + slot;
    double u[n2];
    double g[6];
    if (active)
    {
      const double kw = kappa[perm[cell0 + pl]] * wt; // kappa re-read on every apply (src/laplacian.hpp:230)
      const double* gp = Gc + (size_t)(cell0 + pl) * 6;
#pragma unroll
      for (int c = 0; c < 6; ++c)
        g[c] = gp[c] * kw;
#pragma unroll
      for (int a = 0; a < n2; ++a)
      {
        const int da = dE[a * SE];
        const int idx = da < 0 ? ~da : da;
        const double xv = x[idx];
        if (da < 0)
          y[idx] = xv; // Dirichlet row: y = x (src/laplacian.hpp:273-274)
        u[a] = da < 0 ? 0.0 : xv;
      }
    }
    else
    {
#pragma unroll
      for (int c = 0; c < 6; ++c)
        g[c] = 0.0;
#pragma unroll
      for (int a = 0; a < n2; ++a)
        u[a] = 0.0;
    }
    // phase 1: u(i,j,t) -> A[slab j][i][t]
    if (in_block)
    {
#pragma unroll
      for (int i = 0; i < n; ++i)
#pragma unroll
        for (int j = 0; j < n; ++j)
          wA[j * SS + i * n] = u[i * n + j];
    }
    phase_sync(); // (CTA-wide variant: every thread is also past the previous batch's reads of B and of the other dofmap buffer)
    if (!WL && tid == 0 && it + 1 < my_nb)
      issue_enc(it + 1);
    // phase 2: slab j = t: gz(i,t,l) = sum_m D[l][m] u(i,t,m) -> B[slab l][i][t]
#pragma unroll
    for (int i = 0; i < n; ++i)
    {
      double v[n];
#pragma unroll
      for (int m = 0; m < n; ++m)
        v[m] = rA[i * n + m];
#pragma unroll
      for (int l = 0; l < n; ++l)
      {
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < n; ++m)
          s = fma(c_D[P][l * n + m], v[m], s);
        if (in_block)
          wB[l * SS + i * n] = s;
      }
    }
    phase_sync();
    // phase 3: all three gradients at the points (i,j,t); transform; x / y transposed contractions;
    // fz(i,j,t) -> A[slab j][i][t]
    double acc[n2];
#pragma unroll
    for (int a = 0; a < n2; ++a)
      acc[a] = 0.0;
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int j = 0; j < n; ++j)
      {
        double gx = 0.0, gy = 0.0;
        const double gz = rB[i * n + j];
#pragma unroll
        for (int l = 0; l < n; ++l)
        {
          gx = fma(c_D[P][i * n + l], u[l * n + j], gx);
          gy = fma(c_D[P][j * n + l], u[i * n + l], gy);
        }
        const double wij = c_wts[P][i] * c_wts[P][j]; // G(q) = w_i w_j w_k Gc
        const double fx = wij * (g[0] * gx + g[1] * gy + g[2] * gz);
        const double fy = wij * (g[1] * gx + g[3] * gy + g[4] * gz);
        const double fz = wij * (g[2] * gx + g[4] * gy + g[5] * gz);
#pragma unroll
        for (int l = 0; l < n; ++l)
        {
          acc[l * n + j] = fma(c_D[P][i * n + l], fx, acc[l * n + j]);
          acc[i * n + l] = fma(c_D[P][j * n + l], fy, acc[i * n + l]);
        }
        if (in_block)
          wA[j * SS + i * n] = fz;
      }
    phase_sync();
    // phase 4: slab j = t: az(i,t,l) = sum_q D[q][l] fz(i,t,q) -> B[slab l][i][t]
#pragma unroll
    for (int i = 0; i < n; ++i)
    {
      double v[n];
#pragma unroll
      for (int q = 0; q < n; ++q)
        v[q] = rA[i * n + q];
#pragma unroll
      for (int l = 0; l < n; ++l)
      {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < n; ++q)
          s = fma(c_D[P][q * n + l], v[q], s);
        if (in_block)
          wB[l * SS + i * n] = s;
      }
    }
    phase_sync();
    // phase 5: acc(i,j,t) += az(i,j,t); scatter
    if (active)
    {
#pragma unroll
      for (int a = 0; a < n2; ++a)
      {
        const int da = dE[a * SE];
        if (da >= 0)
          atomicAdd(&y[da], acc[a] + rB[a]);
      }
    }
    if (WL)
      __syncwarp(); // B is rewritten in the next batch's phase 2 only after this warp's reads above
  }
}

template <int P, int TPB, bool WL>
void launch_apply_affine2_t(pmgx_ctx* c, cudaStream_t st, const double* x, double* y, const double* Gc,
                            const int32_t* enc, const int32_t* perm, const double* kappa, long long batch0, int cell0,
                            int count)
{
  using C = Aff2Cfg<P, TPB, WL>;
  const bool timed = c->profiling && cell0 == 0; // per-kernel timing covers the interior-cell launch only
  static int ctas_per_sm[64] = {0};
  if (ctas_per_sm[c->device] == 0)
  {
    PMGX_CUDA(cudaFuncSetAttribute(k_apply_affine2<P, TPB, WL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem));
    int nb = 0;
    PMGX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_apply_affine2<P, TPB, WL>, TPB, C::smem));
    PMGX_REQUIRE(nb >= 1, "k_apply_affine2<%d,%d> does not fit on an SM", P, TPB);
    ctas_per_sm[c->device] = nb;
  }
  const int nbatch = (count + C::cpb - 1) / C::cpb;
  const int grid = std::min(nbatch, ctas_per_sm[c->device] * c->num_sms);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (timed)
  {
    PMGX_CUDA(cudaEventCreate(&e0));
    PMGX_CUDA(cudaEventCreate(&e1));
    PMGX_CUDA(cudaEventRecord(e0, st));
  }
  k_apply_affine2<P, TPB, WL><<<grid, TPB, C::smem, st>>>(x, y, Gc, enc, perm, kappa, batch0, cell0, count, nbatch);
  check_launch("k_apply_affine2");
  count_launch(c);
  if (timed)
  {
    PMGX_CUDA(cudaEventRecord(e1, st));
    c->prof[P].emplace_back(e0, e1);
  }
}

template <int P>
void launch_apply_affine(pmgx_ctx* c, cudaStream_t st, bool shfl, int tpb, const double* x, double* y,
                         const double* Gc, const int32_t* enc, const int32_t* perm, const double* kappa,
                         long long batch0, int cell0, int count, int reslab = 0)
{
  if (count <= 0)
    return;
  // reslab: 1 = CTA-wide phases (batch layout of k_apply_affine), 2 = warp-local phases (batch layout of the shuffle kernel)
  if (reslab == 2 && tpb == 64)
    return launch_apply_affine2_t<P, 64, true>(c, st, x, y, Gc, enc, perm, kappa, batch0, cell0, count);
  if (reslab == 2)
    return launch_apply_affine2_t<P, 128, true>(c, st, x, y, Gc, enc, perm, kappa, batch0, cell0, count);
  if (reslab && tpb == 64)
    return launch_apply_affine2_t<P, 64, false>(c, st, x, y, Gc, enc, perm, kappa, batch0, cell0, count);
  if (reslab)
    return launch_apply_affine2_t<P, 128, false>(c, st, x, y, Gc, enc, perm, kappa, batch0, cell0, count);
  if (shfl && tpb == 64)
    return launch_apply_affine_shfl_t<P, 64>(c, st, x, y, Gc, enc, perm, kappa, batch0, cell0, count);
  if (shfl)
    return launch_apply_affine_shfl_t<P, 128>(c, st, x, y, Gc, enc, perm, kappa, batch0, cell0, count);
  if (tpb == 64)
    return launch_apply_affine_t<P, 64>(c, st, x, y, Gc, enc, perm, kappa, batch0, cell0, count);
  return launch_apply_affine_t<P, 128>(c, st, x, y, Gc, enc, perm, kappa, batch0, cell0, count);
}

// default tuning (threads per CTA, geometry planes in flight); PMGX_TMA_TPB / PMGX_TMA_R override
// (measured on B200, scripts/sweep_tma.sh: more, smaller CTAs with a 2-plane ring win for P3/P4)
inline int tma_default_tpb(int P) { return (P == 1 || P >= 5) ? 64 : 128; }
inline int tma_default_r(int P) { return 2; }
// affine kernel (no geometry ring, so more small CTAs fit): measured at 100 M dofs, P1 3.45 vs 4.07 ms
// (128 vs 64 threads), P2 1.67 vs 1.79, P3 1.61 vs 1.54, P4 1.34 vs 1.28, P5 1.20 vs 1.19, P6 2.06 vs 2.03
inline int affine_default_tpb(int P) { return P <= 2 ? 128 : 64; } // (P3 re-slab kernel: 1.201 / 1.202 ms at 64 / 128)

template <int P>
void launch_apply_tma(pmgx_ctx* c, cudaStream_t st, int tpb, int r, const double* x, double* y, const double* G,
                      const int32_t* enc, const int32_t* perm, const double* kappa, long long batch0, int cell0,
                      int count)
{
  if (count <= 0)
    return;
#define PMGX_TMA_CASE(T, RR)                                                                       \
  if (tpb == T && r == RR)                                                                         \
    return launch_apply_tma_t<P, T, RR>(c, st, x, y, G, enc, perm, kappa, batch0, cell0, count);
  PMGX_TMA_CASE(128, 2)
  PMGX_TMA_CASE(128, 3)
  PMGX_TMA_CASE(64, 2)
  PMGX_TMA_CASE(64, 3)
  PMGX_TMA_CASE(64, 4)
#undef PMGX_TMA_CASE
  set_error("unsupported TMA kernel configuration tpb=%d r=%d", tpb, r);
  throw Error{PMGX_ERR_ARG};
}

#include "laplacian_mma.cuh"

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
  const int32_t* dofmap_borrowed = nullptr; // the caller's dofmap (borrowed like the reference's span, src/laplacian.hpp:502)
  int flags = 0;
  DevBuf<int32_t> perm; // launch position -> caller cell index (lcells then bcells)
  DevBuf<int32_t> enc;  // BC-encoded dofmap, layout `lay`
  DevBuf<double> G;     // geometry factors, layout `lay`
  Lay lay;
  int tma_tpb = 128, tma_r = 2;
  bool use_tma = true; // TMA-pipelined slab kernel (default); PMGX_APPLY_KERNEL=slab|column for A/B runs
  DevBuf<double> Gc;   // [n_list][6] per-cell geometry factor of affine cells
  bool affine = false; // every cell affine: k_apply_affine replaces the streamed-G kernels
  bool aff_shfl = false; // z contractions by warp shuffles instead of shared-memory rows (default for P <= 2)
  int aff_reslab = 0; // re-slabbed z direction (k_apply_affine2): 1 CTA-wide phases, 2 warp-local phases
  bool use_mma = false; // P6 / P7: FP64 tensor-core kernel (k_apply_affine_mma; G per cell if affine, else streamed), plain [p][n3] layout

  int n_list() const { return n_l + n_b; }

  template <int PP>
  void apply_t(double* x, double* y)
  {
    const long long ntot = (long long)n_owned + n_ghost;
    vec::set(ctx, y, ntot, 0.0); // out.set(0) :466 (fill kernel, see vec::set)
    if (halo)
      halo_fwd_begin(halo, x);                                                    // :378
    // interior cells on the compute stream; boundary + ghost cells right behind the exchange on
    // the halo stream: they start when the ghosts are in place and fill the interior kernel's tail
    // (both only add into y, which was zeroed before the exchange started)
    static const bool on_compute = getenv("PMGX_BOUNDARY_ON_COMPUTE") != nullptr; // A/B switch
    // measured at 4 GPUs (scripts/ab_apply_mgpu.py): riding the halo stream wins at P1 (0.114 vs
    // 0.121 ms), is neutral at P2 and loses 2 % at P4, where the interior kernel is long enough to
    // hide the join anyway
    const bool side = halo && !on_compute && PP <= 2;
    cudaStream_t cs = ctx->stream, bs = side ? halo_stream(halo, ctx) : ctx->stream;
    auto join_before_boundary = [&]()
    {
      if (halo && !side)
        halo_fwd_end(halo, x); // reference order: wait for the ghosts, then the boundary launch (:425)
    };
    bool done = false;
    if constexpr (PP == 6 || PP == 7)
    {
      if (use_mma)
      {
        if (affine)
          launch_apply_affine_mma<PP, false>(ctx, cs, x, y, Gc.p, enc.p, perm.p, kappa, 0, n_l);
        else
          launch_apply_affine_mma<PP, true>(ctx, cs, x, y, G.p, enc.p, perm.p, kappa, 0, n_l);
        join_before_boundary();
        if (affine)
          launch_apply_affine_mma<PP, false>(ctx, bs, x, y, Gc.p, enc.p, perm.p, kappa, n_l, n_b);
        else
          launch_apply_affine_mma<PP, true>(ctx, bs, x, y, G.p, enc.p, perm.p, kappa, n_l, n_b);
        done = true;
      }
    }
    if constexpr (PP <= SLAB_MAX_DEGREE)
    {
      if (done)
        ;
      else if (lay.mode == 1 && affine)
      {
        launch_apply_affine<PP>(ctx, cs, aff_shfl, tma_tpb, x, y, Gc.p, enc.p, perm.p, kappa, 0, 0, n_l, aff_reslab);
        join_before_boundary();
        launch_apply_affine<PP>(ctx, bs, aff_shfl, tma_tpb, x, y, Gc.p, enc.p, perm.p, kappa, lay.nb_l, n_l, n_b, aff_reslab);
        done = true;
      }
      else if (lay.mode == 1 && use_tma)
      {
        launch_apply_tma<PP>(ctx, cs, tma_tpb, tma_r, x, y, G.p, enc.p, perm.p, kappa, 0, 0, n_l); // :406-409
        join_before_boundary();
        launch_apply_tma<PP>(ctx, bs, tma_tpb, tma_r, x, y, G.p, enc.p, perm.p, kappa, lay.nb_l, n_l, n_b); // :449-452
        done = true;
      }
      else if (lay.mode == 1)
      {
        launch_apply_slab<PP>(ctx, cs, x, y, G.p, enc.p, perm.p, kappa, 0, 0, n_l);
        join_before_boundary();
        launch_apply_slab<PP>(ctx, bs, x, y, G.p, enc.p, perm.p, kappa, lay.nb_l, n_l, n_b);
        done = true;
      }
    }
    if (!done)
    {
      launch_apply<PP>(ctx, cs, x, y, G.p, enc.p, perm.p, kappa, 0, n_l);
      join_before_boundary();
      launch_apply<PP>(ctx, bs, x, y, G.p, enc.p, perm.p, kappa, n_l, n_b);
    }
    if (side)
      halo_fwd_end(halo, x);                                                      // :425 (join)
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
  L->dofmap_borrowed = dofmap;
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
  pmgx::Lay& lay = L->lay;
  lay.n = n, lay.n2 = n * n, lay.n3 = L->n3;
  lay.n_l = n_lcells;
  const char* force = getenv("PMGX_APPLY_KERNEL"); // "column" forces the column kernel (A/B runs)
  lay.mode = (degree <= pmgx::SLAB_MAX_DEGREE && !(force && std::strcmp(force, "column") == 0)) ? 1 : 0;
  // P6 / P7: the tensor-core kernel, with G per cell on affine meshes and G streamed per quadrature point
  // otherwise (PMGX_APPLY_MMA=0: the slab / column kernels)
  const bool mma_ok = (degree == 6 || degree == 7) && !force
                      && !(getenv("PMGX_APPLY_MMA") && atoi(getenv("PMGX_APPLY_MMA")) == 0);
  L->use_mma = mma_ok;
  if (mma_ok)
    lay.mode = 0; // plain enc[p][n3] / G[p][6][n3]
  // affine cells: one geometry 6-vector per cell instead of one per quadrature point (decided
  // first: the batch layout follows the kernel that will run)
  if (n_list > 0 && (degree <= pmgx::SLAB_MAX_DEGREE || mma_ok) && !force && !(flags & PMGX_LAP_STREAM_G))
  {
    L->Gc.alloc((size_t)n_list * 6);
    pmgx::DevBuf<int> bad;
    bad.alloc(1);
    PMGX_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), ctx->stream));
    pmgx::k_cell_geometry<<<(n_list + 255) / 256, 256, 0, ctx->stream>>>(
        xgeom, geom_dofmap, L->perm.p, L->Gc.p, n_list, (flags & PMGX_LAP_LITERAL_DETJ) != 0, 1e-13, bad.p);
    pmgx::check_launch("k_cell_geometry");
    pmgx::count_launch(ctx);
    int n_bad = 0;
    PMGX_CUDA(cudaMemcpyAsync(&n_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PMGX_CUDA(cudaStreamSynchronize(ctx->stream));
    L->affine = n_bad == 0;
    if (!L->affine)
      L->Gc.release();
    if (degree > pmgx::SLAB_MAX_DEGREE && !L->use_mma)
    {
      L->affine = false; // no affine kernel above the slab degrees
      L->Gc.release();
    }
  }
  L->use_tma = !(force && std::strcmp(force, "slab") == 0);
  L->tma_tpb = L->affine ? pmgx::affine_default_tpb(degree) : pmgx::tma_default_tpb(degree);
  L->tma_r = pmgx::tma_default_r(degree);
  if (const char* e = getenv("PMGX_TMA_TPB"))
    L->tma_tpb = atoi(e);
  if (const char* e = getenv("PMGX_TMA_R"))
    L->tma_r = atoi(e);
  if (!L->use_tma)
    L->tma_tpb = 128;
  PMGX_REQUIRE(L->tma_tpb == 64 || L->tma_tpb == 128, "PMGX_TMA_TPB must be 64 or 128");
  // measured at 100 M dofs (shuffle vs shared-memory z contraction): P1 3.52 vs 3.45 ms, P2 1.58 vs 1.67,
  // P3 1.63 vs 1.54, P4 1.36 vs 1.28, P5 1.57 vs 1.19, P6 2.27 vs 2.03 -- double shuffles cost two issue
  // slots each and only pay where the element is small
  L->aff_shfl = degree == 2;
  if (const char* e = getenv("PMGX_AFFINE_SHFL"))
    L->aff_shfl = atoi(e) != 0;
  // re-slabbed z direction (k_apply_affine2, warp-local phases): measured at 100 M dofs against the broadcast-row
  // kernel -- P1 4.16 vs 3.44 ms, P2 1.76 vs 1.58, P3 1.20 vs 1.54, P4 1.41 vs 1.28, P5 2.40 vs 1.19, P6 3.09 vs 2.03:
  // it halves the shared-memory wavefronts (ncu: L1/shared pipe 78 -> 53 %) but both kernels sit at 2 warps per
  // scheduler with fixed-latency (DFMA dependency) stalls on top, so it only wins where its register count drops
  // far enough to matter (P3)
  L->aff_reslab = degree == 3 ? 2 : 0;
  if (const char* e = getenv("PMGX_AFFINE_RESLAB"))
    L->aff_reslab = atoi(e);
  if (L->aff_reslab)
    L->aff_shfl = false;
  lay.cpb = (L->affine && (L->aff_shfl || L->aff_reslab == 2)) ? (L->tma_tpb / 32) * (32 / n) : L->tma_tpb / n;
  lay.S = (lay.cpb * n + 1) & ~1;
  lay.SE = (lay.cpb * n + 3) & ~3;
  lay.nb_l = (n_lcells + lay.cpb - 1) / lay.cpb;
  lay.n_batches = lay.nb_l + (n_bcells + lay.cpb - 1) / lay.cpb;
  const long long enc_size = lay.mode == 1 ? lay.enc_size() : total;
  const long long g_size = lay.mode == 1 ? lay.g_size() : total * 6;
  L->enc.alloc((size_t)enc_size);
  L->G.alloc((size_t)g_size);
  if (lay.mode == 1 && enc_size > 0)
  { // padded slots: inert
    PMGX_CUDA(cudaMemsetAsync(L->enc.p, 0, (size_t)enc_size * sizeof(int32_t), ctx->stream));
    PMGX_CUDA(cudaMemsetAsync(L->G.p, 0, (size_t)g_size * sizeof(double), ctx->stream));
  }
  if (total > 0)
  {
    pmgx::k_encode_dofmap<<<pmgx::setup_grid(ctx, total), 256, 0, ctx->stream>>>(
        dofmap, L->perm.p, bc_marker, L->enc.p, L->n3, total, lay);
    pmgx::check_launch("k_encode_dofmap");
    pmgx::k_geometry<true><<<pmgx::setup_grid(ctx, total), 256, 0, ctx->stream>>>(
        degree, xgeom, geom_dofmap, L->perm.p, L->G.p, nullptr, n_list,
        (flags & PMGX_LAP_LITERAL_DETJ) != 0, lay);
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
          degree, L->G.p, L->enc.p, L->perm.p, kappa, diag.p, n_list, lay);
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
        L->G.p, G_out, L->n3, total, L->lay);
    pmgx::check_launch("k_G_to_reference_layout");
    pmgx::count_launch(L->ctx);
  }
  PMGX_API_END
}

int pmgx_laplacian_is_affine(pmgx_operator* op)
{
  return op && op->kind == pmgx_operator::LAPLACIAN && static_cast<Laplacian*>(op)->affine ? 1 : 0;
}

int pmgx_laplacian_kernel_name(pmgx_operator* op, char* name_h, int cap)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(op && op->kind == pmgx_operator::LAPLACIAN && name_h && cap > 0, "laplacian_kernel_name: bad arguments");
  auto* L = static_cast<Laplacian*>(op);
  if (L->use_mma)
    snprintf(name_h, cap, "k_apply_affine_mma<%d,2,%s>", L->P, L->affine ? "affine" : "streamed");
  else if (L->lay.mode == 1 && L->affine)
    snprintf(name_h, cap, "%s<%d,%d%s>", L->aff_reslab ? "k_apply_affine2" : (L->aff_shfl ? "k_apply_affine_shfl" : "k_apply_affine"),
             L->P, L->tma_tpb, L->aff_reslab == 2 ? ",warp-local" : "");
  else if (L->lay.mode == 1 && L->use_tma)
    snprintf(name_h, cap, "k_apply_tma<%d,%d,%d>", L->P, L->tma_tpb, L->tma_r);
  else if (L->lay.mode == 1)
    snprintf(name_h, cap, "k_apply_slab<%d>", L->P);
  else
    snprintf(name_h, cap, "k_apply<%d>", L->P);
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
        L->P, L->xgeom, L->geom_dofmap, L->perm.p, nullptr, dw.p, L->n_list(), false, L->lay);
    pmgx::check_launch("k_geometry(detJ)");
    pmgx::k_rhs<<<pmgx::setup_grid(ctx, total), 256, 0, ctx->stream>>>(dw.p, L->enc.p, fvals, b, total, L->n3, L->lay);
    pmgx::check_launch("k_rhs");
    pmgx::count_launch(ctx, 2);
    PMGX_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  if (ntot > 0 && g != 0.0)
  {
    // inhomogeneous data: b -= A_full g_bc, then set_bc (examples/pmg/main.cpp:293-295)
    pmgx::DevBuf<double> gv;
    gv.alloc((size_t)ntot);
    pmgx::k_fill<<<(ntot + 255) / 256, 256, 0, ctx->stream>>>(gv.p, g, ntot);
    pmgx::check_launch("k_fill");
    const int rc = pmgx_laplacian_lift(op, gv.p, b);
    if (rc != PMGX_OK)
      return rc;
    PMGX_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  else if (ntot > 0)
  {
    pmgx::k_set_bc_value<<<(ntot + 255) / 256, 256, 0, ctx->stream>>>(b, L->bc, g, ntot);
    pmgx::check_launch("k_set_bc_value");
    pmgx::count_launch(ctx);
  }
  PMGX_API_END
}

int pmgx_laplacian_lift(pmgx_operator* op, const double* gvals, double* b)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(op && op->kind == pmgx_operator::LAPLACIAN && gvals && b, "laplacian_lift: bad arguments");
  auto* L = static_cast<Laplacian*>(op);
  pmgx_ctx* ctx = L->ctx;
  PMGX_CUDA(cudaSetDevice(ctx->device));
  const int ntot = L->n_owned + L->n_ghost;
  if (ntot == 0)
    return PMGX_OK;
  // the same cells, geometry and kappa with an all-zero Dirichlet marker: A_full.  No halo: g_bc is
  // given on owned AND ghost dofs, and the ghost cells make the owned rows complete.
  pmgx::DevBuf<int8_t> nobc;
  nobc.alloc((size_t)ntot);
  PMGX_CUDA(cudaMemsetAsync(nobc.p, 0, (size_t)ntot, ctx->stream));
  std::vector<int32_t> perm_h((size_t)L->n_list());
  if (!perm_h.empty())
    PMGX_CUDA(cudaMemcpy(perm_h.data(), L->perm.p, perm_h.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  pmgx_operator* full = nullptr;
  const int n_points_unknown = 0; // only used for argument checks
  const int rc = pmgx_laplacian_create(ctx, L->P, L->n_cells, L->dofmap_borrowed, L->xgeom, n_points_unknown, L->geom_dofmap,
                                       L->kappa, perm_h.data(), (int)perm_h.size(), nullptr, 0, nobc.p, L->n_owned,
                                       L->n_ghost, nullptr, (L->flags & PMGX_LAP_LITERAL_DETJ) | PMGX_LAP_NO_DIAG, &full);
  if (rc != PMGX_OK)
    return rc;
  pmgx::DevBuf<double> gbc, y;
  gbc.alloc((size_t)ntot);
  y.alloc((size_t)ntot);
  pmgx::k_mask_to_bc<<<(ntot + 255) / 256, 256, 0, ctx->stream>>>(gvals, L->bc, gbc.p, ntot);
  pmgx::check_launch("k_mask_to_bc");
  full->apply(gbc.p, y.p);
  pmgx::k_lift<<<(L->n_owned + 255) / 256, 256, 0, ctx->stream>>>(b, y.p, gvals, L->bc, L->n_owned);
  pmgx::check_launch("k_lift");
  pmgx::count_launch(ctx, 2);
  PMGX_CUDA(cudaStreamSynchronize(ctx->stream));
  pmgx_operator_destroy(full);
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
        L->P, L->G.p, L->enc.p, L->perm.p, L->kappa, n_list, n_owned, ntot, keys.p, vals.p, L->lay);
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
