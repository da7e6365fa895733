// 1-D Gauss-Lobatto-Legendre tables, tqli and the mesh-size fit (host code).
// Stand-in for the Basix calls of the reference: make_quadrature(gll) + create_element(P,
// interval, gll_warped).tabulate(1, ...) (src/laplacian.hpp:299-317, src/precompute.hpp:255-271)
// and compute_interpolation_operator (src/interpolate.hpp:118).  Nodes == quadrature points,
// ascending on [0,1].
#include "common.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace pmgx
{
static thread_local std::string g_last_error;

void set_error(const char* fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

// Legendre P_N(t) and derivatives by the three-term recurrence.
static void legendre(int N, double t, double& p, double& dp, double& ddp)
{
  double p0 = 1.0, p1 = t;
  if (N == 0)
  {
    p = 1.0, dp = 0.0, ddp = 0.0;
    return;
  }
  for (int k = 2; k <= N; ++k)
  {
    double pk = ((2.0 * k - 1.0) * t * p1 - (k - 1.0) * p0) / k;
    p0 = p1;
    p1 = pk;
  }
  p = p1;
  // (1-t^2) P_N' = N (P_{N-1} - t P_N)
  dp = N * (p0 - t * p1) / (1.0 - t * t);
  // (1-t^2) P_N'' = 2 t P_N' - N(N+1) P_N
  ddp = (2.0 * t * dp - N * (N + 1.0) * p1) / (1.0 - t * t);
}

void gll_points_weights(int n, std::vector<double>& x, std::vector<double>& w)
{
  if (n < 2)
    throw std::runtime_error("GLL rule needs at least 2 points");
  const int N = n - 1;
  std::vector<double> t(n);
  t[0] = -1.0;
  t[N] = 1.0;
  for (int i = 1; i < N; ++i)
  {
    // Chebyshev-Gauss-Lobatto initial guess, Newton on P_N'
    double ti = -std::cos(M_PI * i / N);
    for (int it = 0; it < 100; ++it)
    {
      double p, dp, ddp;
      legendre(N, ti, p, dp, ddp);
      double dt = dp / ddp;
      ti -= dt;
      if (std::fabs(dt) < 1e-16)
        break;
    }
    t[i] = ti;
  }
  for (int i = 0; i < n / 2; ++i)
  {
    double s = 0.5 * (t[n - 1 - i] - t[i]);
    t[i] = -s;
    t[n - 1 - i] = s;
  }
  if (n % 2 == 1)
    t[n / 2] = 0.0;
  x.resize(n);
  w.resize(n);
  for (int i = 0; i < n; ++i)
  {
    double p, dp, ddp;
    if (i == 0 || i == N)
    {
      p = (i == 0 && (N % 2)) ? -1.0 : 1.0;
    }
    else
      legendre(N, t[i], p, dp, ddp);
    x[i] = 0.5 * (t[i] + 1.0);
    w[i] = 0.5 * 2.0 / (N * (N + 1.0) * p * p);
  }
}

static std::vector<double> bary_weights(const std::vector<double>& x)
{
  const int n = (int)x.size();
  std::vector<double> b(n, 1.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      if (i != j)
        b[i] /= (x[i] - x[j]);
  return b;
}

void gll_deriv_matrix(const std::vector<double>& x, std::vector<double>& D)
{
  const int n = (int)x.size();
  std::vector<double> b = bary_weights(x);
  D.assign((size_t)n * n, 0.0);
  for (int q = 0; q < n; ++q)
  {
    double s = 0.0;
    for (int i = 0; i < n; ++i)
      if (i != q)
      {
        D[q * n + i] = (b[i] / b[q]) / (x[q] - x[i]);
        s += D[q * n + i];
      }
    D[q * n + q] = -s;
  }
}

void gll_interp_matrix(int pc, int pf, std::vector<double>& M)
{
  std::vector<double> xc, wc, xf, wf;
  gll_points_weights(pc + 1, xc, wc);
  gll_points_weights(pf + 1, xf, wf);
  std::vector<double> b = bary_weights(xc);
  const int nc = pc + 1, nf = pf + 1;
  M.assign((size_t)nf * nc, 0.0);
  for (int f = 0; f < nf; ++f)
  {
    int hit = -1;
    for (int c = 0; c < nc; ++c)
      if (std::fabs(xf[f] - xc[c]) < 1e-14)
        hit = c;
    if (hit >= 0)
    {
      M[f * nc + hit] = 1.0;
      continue;
    }
    double s = 0.0;
    for (int c = 0; c < nc; ++c)
    {
      M[f * nc + c] = b[c] / (xf[f] - xc[c]);
      s += M[f * nc + c];
    }
    for (int c = 0; c < nc; ++c)
      M[f * nc + c] /= s;
  }
}

// QL with implicit shifts on a symmetric tridiagonal matrix (role of tqli, src/cg.hpp:15-84).
static int tridiag_ql(double* d, double* e, int n)
{
  for (int l = 0; l < n; ++l)
  {
    int sweeps = 0;
    for (;;)
    {
      int m = l;
      for (; m < n - 1; ++m)
      {
        const double scale = std::fabs(d[m]) + std::fabs(d[m + 1]);
        if (std::fabs(e[m]) + scale == scale)
          break;
      }
      if (m == l)
        break;
      if (sweeps++ == 30)
        return -1;
      double shift = (d[l + 1] - d[l]) / (2.0 * e[l]);
      const double hyp = std::sqrt(shift * shift + 1.0);
      shift = d[m] - d[l] + e[l] / (shift + (shift >= 0 ? hyp : -hyp));
      double sn = 1.0, cs = 1.0, carry = 0.0;
      bool deflated = false;
      for (int i = m - 1; i >= l; --i)
      {
        const double f = sn * e[i];
        const double b = cs * e[i];
        const double rad = std::sqrt(f * f + shift * shift);
        e[i + 1] = rad;
        if (rad == 0.0)
        {
          d[i + 1] -= carry;
          e[m] = 0.0;
          deflated = true;
          break;
        }
        sn = f / rad;
        cs = shift / rad;
        shift = d[i + 1] - carry;
        const double t = (d[i] - shift) * sn + 2.0 * cs * b;
        carry = sn * t;
        d[i + 1] = shift + carry;
        shift = cs * t - b;
      }
      if (deflated)
        continue;
      d[l] -= carry;
      e[l] = shift;
      e[m] = 0.0;
    }
    e[l] = 0.0;
  }
  return 0;
}
} // namespace pmgx

extern "C"
{
const char* pmgx_last_error_string(void) { return pmgx::g_last_error.c_str(); }
int pmgx_version(void) { return 100; }

int pmgx_gll_tables(int degree, double* points_h, double* weights_h, double* dphi_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(degree >= 1 && degree <= PMGX_MAX_DEGREE, "Unsupported degree %d", degree);
  std::vector<double> x, w, D;
  pmgx::gll_points_weights(degree + 1, x, w);
  pmgx::gll_deriv_matrix(x, D);
  if (points_h)
    std::memcpy(points_h, x.data(), x.size() * sizeof(double));
  if (weights_h)
    std::memcpy(weights_h, w.data(), w.size() * sizeof(double));
  if (dphi_h)
    std::memcpy(dphi_h, D.data(), D.size() * sizeof(double));
  PMGX_API_END
}

int pmgx_gll_interp_1d(int degree_coarse, int degree_fine, double* interp_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(degree_coarse >= 1 && degree_fine <= PMGX_MAX_DEGREE && degree_coarse <= degree_fine,
               "Unsupported degrees %d -> %d", degree_coarse, degree_fine);
  std::vector<double> M;
  pmgx::gll_interp_matrix(degree_coarse, degree_fine, M);
  std::memcpy(interp_h, M.data(), M.size() * sizeof(double));
  PMGX_API_END
}

int pmgx_tqli(double* d_h, double* e_h, int n)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(n >= 1 && d_h && e_h, "tqli: bad arguments");
  if (pmgx::tridiag_ql(d_h, e_h, n) != 0)
  {
    pmgx::set_error("Eigenvalue estimate failed");
    return PMGX_ERR_NUMERIC;
  }
  PMGX_API_END
}

int pmgx_boxmesh_fit(long long ndofs_total, int order, int* nxyz_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(ndofs_total > 0 && order >= 1 && nxyz_h, "boxmesh_fit: bad arguments");
  // cube-root start, then +-5 search in each direction (examples/pmg/main.cpp:412-435)
  const double approx = (std::pow((double)ndofs_total, 1.0 / 3.0) - 1.0) / order;
  const long long n0 = std::max<long long>(1, (long long)approx);
  long long best[3] = {n0, n0, n0};
  if (n0 > 5)
  {
    auto count = [order](long long a, long long b, long long c)
    { return (a * order + 1) * (b * order + 1) * (c * order + 1); };
    long long best_misfit = std::llabs(count(n0, n0, n0) - ndofs_total);
    for (long long a = n0 - 5; a < n0 + 6; ++a)
      for (long long b = n0 - 5; b < n0 + 6; ++b)
        for (long long c = n0 - 5; c < n0 + 6; ++c)
        {
          const long long mis = std::llabs(count(a, b, c) - ndofs_total);
          if (mis < best_misfit)
          {
            best_misfit = mis;
            best[0] = a, best[1] = b, best[2] = c;
          }
        }
  }
  for (int i = 0; i < 3; ++i)
    nxyz_h[i] = (int)best[i];
  PMGX_API_END
}
}
