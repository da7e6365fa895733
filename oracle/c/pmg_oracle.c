/* C/OpenMP restatement of the hot-path kernels (oracle; TEST INFRASTRUCTURE ONLY).
 *
 * Used (a) as the checker at sizes numpy cannot reach and (b) as the timed CPU stand-in
 * ("port") for bench.py's cpu_baseline / --impl reference legs: the reference's own CPU path
 * (python_tests/pmg.py on DOLFINx + PETSc) cannot be installed here.  Never linked into or
 * called from the product (pmg_dolfinx_b200/).
 *
 * Follows, function by function:
 *   orc_geometry   src/laplacian.hpp:72-111 (J, K = adj J, G = w K K^T / detJ), exact detJ
 *   orc_apply      src/laplacian.hpp:182-277 (gather with BC zeroing, 3 forward contractions
 *                  with phi = identity, G transform, 3 transposed contractions, scatter-add,
 *                  BC rows y = x); one cell per loop iteration instead of one thread block
 *   orc_spmv       src/csr.hpp:20-36
 *   orc_axpy/dot   src/vector.hpp:333-352,397-407
 * All tables (dphi, points, weights) are passed in from oracle/gll.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXN 9

int orc_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* the launcher may have exported OMP_NUM_THREADS=1 (torch.distributed.run does): the timed CPU arm
 * sets the thread count explicitly */
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
  if (n > 0)
    omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* G[c][q][6] (reference layout), detj[c][q]; pts/w1: 1-D GLL points/weights, n = P+1 */
void orc_geometry(int n, const double* pts, const double* w1, const double* xgeom,
                  const int32_t* geom_dofmap, int64_t ncells, double* G, double* detj)
{
  const int nq = n * n * n;
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < ncells; ++c)
  {
    double v[8][3];
    for (int k = 0; k < 8; ++k)
      for (int d = 0; d < 3; ++d)
        v[k][d] = xgeom[3 * (int64_t)geom_dofmap[c * 8 + k] + d];
    for (int q = 0; q < nq; ++q)
    {
      const int ix = q / (n * n), iy = (q / n) % n, iz = q % n;
      const double xi[3] = {pts[ix], pts[iy], pts[iz]};
      double J[3][3] = {{0}};
      for (int k = 0; k < 8; ++k)
      {
        const int a = (k >> 2) & 1, b = (k >> 1) & 1, cc = k & 1;
        const double la = a ? xi[0] : 1 - xi[0], lb = b ? xi[1] : 1 - xi[1], lc = cc ? xi[2] : 1 - xi[2];
        const double dphi[3] = {(a ? 1.0 : -1.0) * lb * lc, la * (b ? 1.0 : -1.0) * lc, la * lb * (cc ? 1.0 : -1.0)};
        for (int i = 0; i < 3; ++i)
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
      const double det = J[0][0] * K[0][0] + J[0][1] * K[1][0] + J[0][2] * K[2][0];
      const double s = w1[ix] * w1[iy] * w1[iz] / det;
      double* g = G + (c * nq + q) * 6;
      g[0] = (K[0][0] * K[0][0] + K[0][1] * K[0][1] + K[0][2] * K[0][2]) * s;
      g[1] = (K[1][0] * K[0][0] + K[1][1] * K[0][1] + K[1][2] * K[0][2]) * s;
      g[2] = (K[2][0] * K[0][0] + K[2][1] * K[0][1] + K[2][2] * K[0][2]) * s;
      g[3] = (K[1][0] * K[1][0] + K[1][1] * K[1][1] + K[1][2] * K[1][2]) * s;
      g[4] = (K[2][0] * K[1][0] + K[2][1] * K[1][1] + K[2][2] * K[1][2]) * s;
      g[5] = (K[2][0] * K[2][0] + K[2][1] * K[2][1] + K[2][2] * K[2][2]) * s;
      if (detj)
        detj[c * nq + q] = det;
    }
  }
}

/* y = A x (zero fill included, src/laplacian.hpp:466); D[q*n+i]; G[c][q][6] */
void orc_apply(int n, const double* D, const int32_t* dofmap, const double* G, const double* kappa,
               const int8_t* bc, int64_t ncells, int64_t ndofs, const double* x, double* y)
{
  const int n2 = n * n, n3 = n2 * n;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < ndofs; ++i)
    y[i] = 0.0;
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < ncells; ++c)
  {
    double u[MAXN * MAXN * MAXN], fx[MAXN * MAXN * MAXN], fy[MAXN * MAXN * MAXN], fz[MAXN * MAXN * MAXN];
    const int32_t* dm = dofmap + c * n3;
    for (int a = 0; a < n3; ++a)
      u[a] = bc[dm[a]] ? 0.0 : x[dm[a]];
    const double* g = G + c * n3 * 6;
    const double kap = kappa[c];
    for (int ix = 0; ix < n; ++ix)
      for (int iy = 0; iy < n; ++iy)
        for (int iz = 0; iz < n; ++iz)
        {
          double vx = 0, vy = 0, vz = 0;
          for (int l = 0; l < n; ++l)
          {
            vx += D[ix * n + l] * u[l * n2 + iy * n + iz];
            vy += D[iy * n + l] * u[ix * n2 + l * n + iz];
            vz += D[iz * n + l] * u[ix * n2 + iy * n + l];
          }
          const int q = ix * n2 + iy * n + iz;
          const double* gq = g + q * 6;
          fx[q] = kap * (gq[0] * vx + gq[1] * vy + gq[2] * vz);
          fy[q] = kap * (gq[1] * vx + gq[3] * vy + gq[4] * vz);
          fz[q] = kap * (gq[2] * vx + gq[4] * vy + gq[5] * vz);
        }
    for (int ix = 0; ix < n; ++ix)
      for (int iy = 0; iy < n; ++iy)
        for (int iz = 0; iz < n; ++iz)
        {
          double vx = 0, vy = 0, vz = 0;
          for (int l = 0; l < n; ++l)
          {
            vx += D[l * n + ix] * fx[l * n2 + iy * n + iz];
            vy += D[l * n + iy] * fy[ix * n2 + l * n + iz];
            vz += D[l * n + iz] * fz[ix * n2 + iy * n + l];
          }
          const int a = ix * n2 + iy * n + iz;
          const int32_t d = dm[a];
          if (bc[d])
            y[d] = x[d];
          else
          {
            const double val = vx + vy + vz;
#pragma omp atomic
            y[d] += val;
          }
        }
  }
}

void orc_spmv(int64_t nrows, const int32_t* row_ptr, const int32_t* cols, const double* vals,
              const double* x, double* y)
{
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nrows; ++i)
  {
    double s = 0.0;
    for (int32_t j = row_ptr[i]; j < row_ptr[i + 1]; ++j)
      s += vals[j] * x[cols[j]];
    y[i] = s;
  }
}

void orc_axpy(int64_t n, double alpha, const double* x, const double* y, double* r)
{
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
    r[i] = x[i] * alpha + y[i];
}

double orc_dot(int64_t n, const double* a, const double* b)
{
  double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
  for (int64_t i = 0; i < n; ++i)
    s += a[i] * b[i];
  return s;
}
