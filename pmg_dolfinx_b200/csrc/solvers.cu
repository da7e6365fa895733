// Host orchestration + fused vector kernels of the smoother, the Krylov solver, the coarse
// solver and the V-cycle.  Replaces acc::Chebyshev (src/chebyshev.hpp), acc::CGSolver
// (src/cg.hpp), CoarseSolverType (src/amg.hpp) and acc::MultigridPreconditioner (src/pmg.hpp).
//
// Fusion (north_star items 2 and 3): every Chebyshev iteration is one operator apply plus ONE
// vector pass (r -= q, z <- c1 z + c2 D^-1 r, x += z); every CG iteration is one apply plus
// three passes (p.y | x,r update + r.D^-1 r | p update) with alpha formed on the device, so
// the host blocks once per iteration (for the reference's convergence test, src/cg.hpp:206).
#include "common.hpp"
#include "operator.hpp"
#include "reduce.cuh"

#include <cmath>

namespace pmgx
{
namespace
{
constexpr int FT = RED_THREADS;

inline int fused_grid(pmgx_ctx* c, long long n)
{
  long long b = (n + FT * 4 - 1) / (FT * 4);
  b = std::min<long long>(b, c->max_red_blocks);
  return (int)std::max<long long>(b, 1);
}

// r = b - q ; z = (D^-1 r) * c0 ; [x += z] ; optional ||r||^2        (chebyshev.hpp:56-68,73)
template <bool NORM>
__global__ void __launch_bounds__(FT)
k_cheb_init(const double* __restrict__ b, const double* __restrict__ q, const double* __restrict__ dinv,
            double* __restrict__ r, double* __restrict__ z, double* __restrict__ x, double c0,
            bool add_x, long long n, double* partials, unsigned int* counter, double* out)
{
  const long long nth = (long long)gridDim.x * blockDim.x;
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth)
  {
    const double rv = q ? q[i] * (-1.0) + b[i] : b[i]; // q == nullptr: x is known to be zero
    const double zv = (rv * dinv[i]) * c0;
    r[i] = rv;
    z[i] = zv;
    if (add_x)
      x[i] = q ? zv * 1.0 + x[i] : zv;
    if (NORM)
      s = fma(rv, rv, s);
  }
  if (NORM)
  {
    double v[1] = {s};
    grid_reduce<1>(v, partials, counter, out);
  }
}

// r -= q ; z = z*c1 + c2*(D^-1 r) ; [x += z] ; optional ||r||^2     (chebyshev.hpp:76-83,73)
template <bool NORM>
__global__ void __launch_bounds__(FT)
k_cheb_step(const double* __restrict__ q, const double* __restrict__ dinv, double* __restrict__ r,
            double* __restrict__ z, double* __restrict__ x, double c1, double c2, bool add_x,
            long long n, double* partials, unsigned int* counter, double* out)
{
  const long long nth = (long long)gridDim.x * blockDim.x;
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth)
  {
    const double rv = q[i] * (-1.0) + r[i];
    const double zs = z[i] * c1;
    const double zv = (rv * dinv[i]) * c2 + zs;
    r[i] = rv;
    z[i] = zv;
    if (add_x)
      x[i] = zv * 1.0 + x[i];
    if (NORM)
      s = fma(rv, rv, s);
  }
  if (NORM)
  {
    double v[1] = {s};
    grid_reduce<1>(v, partials, counter, out);
  }
}

// r = b - y ; p = D^-1 r ; rnorm0 = p.r                               (cg.hpp:160-164)
__global__ void __launch_bounds__(FT)
k_cg_init(const double* __restrict__ b, const double* __restrict__ y, const double* __restrict__ dinv,
          double* __restrict__ r, double* __restrict__ p, long long n, double* partials,
          unsigned int* counter, double* out)
{
  const long long nth = (long long)gridDim.x * blockDim.x;
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth)
  {
    const double rv = y[i] * (-1.0) + b[i];
    const double pv = rv * dinv[i];
    r[i] = rv;
    p[i] = pv;
    s = fma(pv, rv, s);
  }
  double v[1] = {s};
  grid_reduce<1>(v, partials, counter, out);
}

// alpha = rnorm / (p.y) ; x += alpha p ; r -= alpha y ; y = D^-1 r ; out = r.y  (cg.hpp:182-195)
__global__ void __launch_bounds__(FT)
k_cg_update(double rnorm, const double* __restrict__ rnorm_dev, const double* __restrict__ pAp,
            const double* __restrict__ p, const double* __restrict__ dinv, double* __restrict__ x,
            double* __restrict__ r, double* __restrict__ y, long long n, double* partials,
            unsigned int* counter, double* out)
{
  if (rnorm_dev)
    rnorm = *rnorm_dev; // device-resident variant (no host round trip between iterations)
  const double alpha = rnorm / *pAp;
  const double nalpha = -alpha;
  const long long nth = (long long)gridDim.x * blockDim.x;
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth)
  {
    const double pv = p[i];
    x[i] = pv * alpha + x[i];
    const double rv = y[i] * nalpha + r[i];
    const double yv = rv * dinv[i];
    r[i] = rv;
    y[i] = yv;
    s = fma(rv, yv, s);
  }
  double v[1] = {s};
  grid_reduce<1>(v, partials, counter, out);
}

// p = (rn_new / rn_old) p + y with the ratio formed on the device                (cg.hpp:197,211)
__global__ void __launch_bounds__(FT)
k_cg_pupdate(const double* __restrict__ rn_new, const double* __restrict__ rn_old,
             const double* __restrict__ y, double* __restrict__ p, long long n)
{
  const double beta = *rn_new / *rn_old;
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth)
    p[i] = p[i] * beta + y[i];
}

void check(const char* w) { check_launch(w); }
} // namespace
} // namespace pmgx

// ------------------------------------------------------------------------- Chebyshev --
struct pmgx_cheb
{
  pmgx_ctx* ctx = nullptr;
  int n_owned = 0, n_ghost = 0;
  double eig_min = 0.0, eig_max = 1.0;
  int max_iter = 0;
  pmgx::DevBuf<double> z, q, r; // work vectors (owned + ghost), src/chebyshev.hpp:101-105
};

namespace pmgx
{
// x <- Chebyshev(A, x, b). hist: max_iter+1 residual norms or nullptr.
void cheb_solve(pmgx_cheb* s, pmgx_operator* A, double* x, const double* b, double* hist,
                bool x_is_zero = false)
{
  pmgx_ctx* c = s->ctx;
  const long long n = s->n_owned;
  const double lmax = s->eig_max; // only the upper bound is used (chebyshev.hpp:51)
  const double* dinv = A->diag_inv.p; // get_diag_inverse without the per-call copy (:53)
  const int grid = fused_grid(c, n);
  if (!x_is_zero)
    A->apply(x, s->q.p); // :56
  const double* q0 = x_is_zero ? nullptr : s->q.p; // A*0 = 0: skip the apply, r = b
  const bool first_add = s->max_iter >= 1;
  const double c0 = 4.0 / (3.0 * lmax);
  if (hist)
    k_cheb_init<true><<<grid, FT, 0, c->stream>>>(b, q0, dinv, s->r.p, s->z.p, x, c0, first_add, n,
                                                  c->d_partials, c->d_counter, c->d_scalars + 8);
  else
    k_cheb_init<false><<<grid, FT, 0, c->stream>>>(b, q0, dinv, s->r.p, s->z.p, x, c0, first_add, n,
                                                   c->d_partials, c->d_counter, c->d_scalars + 8);
  check("k_cheb_init");
  count_launch(c);
  if (hist)
  {
    vec::allreduce_scalars(c, 8, 1, false);
    hist[0] = std::sqrt(vec::read_scalar(c, 8));
  }
  for (int i = 1; i <= s->max_iter; ++i)
  {
    A->apply(s->z.p, s->q.p); // :76
    const double c1 = double(2 * i - 1) / double(2 * i + 3);
    const double c2 = double(8 * i + 4) / double(2 * i + 3) / lmax;
    const bool add = i < s->max_iter; // the x += z of iteration i+1 (:73)
    if (hist)
      k_cheb_step<true><<<grid, FT, 0, c->stream>>>(s->q.p, dinv, s->r.p, s->z.p, x, c1, c2, add, n,
                                                    c->d_partials, c->d_counter, c->d_scalars + 8);
    else
      k_cheb_step<false><<<grid, FT, 0, c->stream>>>(s->q.p, dinv, s->r.p, s->z.p, x, c1, c2, add, n,
                                                     c->d_partials, c->d_counter, c->d_scalars + 8);
    check("k_cheb_step");
    count_launch(c);
    if (hist)
    {
      vec::allreduce_scalars(c, 8, 1, false);
      hist[i] = std::sqrt(vec::read_scalar(c, 8));
    }
  }
}
} // namespace pmgx

// -------------------------------------------------------------------------------- CG --
struct pmgx_cg
{
  pmgx_ctx* ctx = nullptr;
  int n_owned = 0, n_ghost = 0;
  int max_iter = 0;
  double rtol = 0.0;
  bool store = false;
  pmgx::DevBuf<double> r, y, p; // src/cg.hpp:241-244
  std::vector<double> alphas, betas, residuals; // stored coefficients (:213-218)
  std::vector<double> history;                  // every iteration's r.M^-1 r
  double rnorm0 = 0.0;
};

namespace pmgx
{
int cg_solve(pmgx_cg* s, pmgx_operator* A, double* x, const double* b)
{
  pmgx_ctx* c = s->ctx;
  const long long n = s->n_owned;
  const double* dinv = A->diag_inv.p;
  const int grid = fused_grid(c, n);
  s->history.clear();
  A->apply(x, s->y.p);                                                           // :159
  k_cg_init<<<grid, FT, 0, c->stream>>>(b, s->y.p, dinv, s->r.p, s->p.p, n, c->d_partials,
                                        c->d_counter, c->d_scalars + 1);          // :160-164
  check("k_cg_init");
  count_launch(c);
  vec::allreduce_scalars(c, 1, 1, false);
  const double rnorm0 = vec::read_scalar(c, 1);
  s->rnorm0 = rnorm0;
  double rnorm = rnorm0;
  const double rtol2 = s->rtol * s->rtol;
  int k = 0;
  while (k < s->max_iter)
  {
    ++k;
    A->apply(s->p.p, s->y.p);                                                    // :179
    vec::dot_device(c, s->p.p, s->y.p, n, 2);                                    // :182
    k_cg_update<<<grid, FT, 0, c->stream>>>(rnorm, nullptr, c->d_scalars + 2, s->p.p, dinv, x, s->r.p, s->y.p, n,
                                            c->d_partials, c->d_counter, c->d_scalars + 3); // :186-195
    check("k_cg_update");
    count_launch(c);
    vec::allreduce_scalars(c, 3, 1, false);
    // one host round trip per iteration: p.y (for alpha) and the new r.M^-1 r
    PMGX_CUDA(cudaMemcpyAsync(c->h_scalars + 2, c->d_scalars + 2, 2 * sizeof(double),
                              cudaMemcpyDeviceToHost, c->stream));
    PMGX_CUDA(cudaStreamSynchronize(c->stream));
    const double alpha = rnorm / c->h_scalars[2];
    const double rnorm_new = c->h_scalars[3];
    const double beta = rnorm_new / rnorm;
    rnorm = rnorm_new;
    s->history.push_back(rnorm);
    if (rnorm / rnorm0 < rtol2)                                                  // :206
      break;
    vec::axpy(c, s->p.p, beta, s->p.p, s->y.p, n);                               // :211
    if (s->store)
    {
      s->alphas.push_back(alpha);
      s->betas.push_back(beta);
      s->residuals.push_back(rnorm);
    }
  }
  return k;
}
} // namespace pmgx

namespace pmgx
{
// Same recurrence as cg_solve, but alpha and beta never leave the device: the host only looks
// at the residual every `check_every` iterations, so an iteration costs launches and two
// all-reduces, no host round trip.  Used by the coarse solver, where iteration counts are not a
// parity quantity (the reference runs PETSc CG + BoomerAMG there, src/amg.hpp:33-47).
int cg_solve_device(pmgx_cg* s, pmgx_operator* A, double* x, const double* b, int check_every)
{
  pmgx_ctx* c = s->ctx;
  const long long n = s->n_owned;
  const double* dinv = A->diag_inv.p;
  const int grid = fused_grid(c, n);
  double* rn = c->d_scalars + 16; // rn[0], rn[1]: ping-pong r.M^-1 r
  double* pap = c->d_scalars + 18;
  A->apply(x, s->y.p);
  k_cg_init<<<grid, FT, 0, c->stream>>>(b, s->y.p, dinv, s->r.p, s->p.p, n, c->d_partials, c->d_counter, rn);
  check("k_cg_init");
  count_launch(c);
  vec::allreduce_scalars(c, 16, 1, false);
  const double rnorm0 = vec::read_scalar(c, 16);
  s->rnorm0 = rnorm0;
  s->history.clear();
  if (!(rnorm0 > 0.0))
    return 0;
  const double rtol2 = s->rtol * s->rtol;
  int k = 0;
  while (k < s->max_iter)
  {
    const int cur = k & 1, nxt = cur ^ 1;
    ++k;
    A->apply(s->p.p, s->y.p);
    vec::dot_device(c, s->p.p, s->y.p, n, 18);
    k_cg_update<<<grid, FT, 0, c->stream>>>(0.0, rn + cur, pap, s->p.p, dinv, x, s->r.p, s->y.p, n,
                                            c->d_partials, c->d_counter, rn + nxt);
    check("k_cg_update");
    vec::allreduce_scalars(c, 16 + nxt, 1, false);
    if (k % check_every == 0 || k == s->max_iter)
    {
      const double r = vec::read_scalar(c, 16 + nxt);
      s->history.push_back(r);
      if (r / rnorm0 < rtol2)
        break;
    }
    k_cg_pupdate<<<grid, FT, 0, c->stream>>>(rn + nxt, rn + cur, s->y.p, s->p.p, n);
    check("k_cg_pupdate");
    count_launch(c, 2);
  }
  return k;
}
} // namespace pmgx

// ------------------------------------------------------------------------ coarse solver --
struct pmgx_coarse
{
  pmgx_ctx* ctx = nullptr;
  pmgx_operator* A = nullptr;
  pmgx_cg* cg = nullptr;
};

// ----------------------------------------------------------------------------- V-cycle --
struct pmgx_vcycle
{
  pmgx_ctx* ctx = nullptr;
  int n_levels = 0;
  int flags = 0;
  std::vector<pmgx_operator*> ops;
  std::vector<pmgx_cheb*> smoothers;
  std::vector<pmgx_interp*> interps;
  std::vector<const int8_t*> bc;
  pmgx_coarse* coarse = nullptr;
  std::vector<pmgx::DevBuf<double>> u, r, b, du; // src/pmg.hpp:165-168
  std::vector<double> diagnostics;
};

namespace pmgx
{
namespace
{
// r = b - A u on level i; optionally its norm
double residual(pmgx_vcycle* v, int i, bool want_norm)
{
  pmgx_ctx* c = v->ctx;
  pmgx_operator* A = v->ops[i];
  A->apply(v->u[i].p, v->r[i].p);
  vec::axpy(c, v->r[i].p, -1.0, v->r[i].p, v->b[i].p, A->n_owned);
  if (!want_norm)
    return 0.0;
  return std::sqrt(vec::dot(c, v->r[i].p, v->r[i].p, A->n_owned));
}
} // namespace
} // namespace pmgx

extern "C"
{
// ---- Chebyshev
int pmgx_cheb_create(pmgx_ctx* ctx, int n_owned, int n_ghost, double eig_min, double eig_max, pmgx_cheb** out)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(ctx && out && n_owned >= 0 && n_ghost >= 0, "cheb_create: bad arguments");
  PMGX_REQUIRE(eig_max > 0.0, "cheb_create: eig_max must be positive");
  PMGX_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<pmgx_cheb> s(new pmgx_cheb());
  s->ctx = ctx;
  s->n_owned = n_owned;
  s->n_ghost = n_ghost;
  s->eig_min = eig_min;
  s->eig_max = eig_max;
  const size_t nt = (size_t)n_owned + n_ghost;
  s->z.alloc(nt);
  s->q.alloc(nt);
  s->r.alloc(nt);
  if (nt > 0)
  {
    PMGX_CUDA(cudaMemsetAsync(s->z.p, 0, nt * sizeof(double), ctx->stream));
    PMGX_CUDA(cudaMemsetAsync(s->q.p, 0, nt * sizeof(double), ctx->stream));
    PMGX_CUDA(cudaMemsetAsync(s->r.p, 0, nt * sizeof(double), ctx->stream));
  }
  *out = s.release();
  PMGX_API_END
}
int pmgx_cheb_set_max_iterations(pmgx_cheb* s, int max_iter)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(s && max_iter >= 0, "cheb_set_max_iterations: bad arguments");
  s->max_iter = max_iter;
  PMGX_API_END
}
int pmgx_cheb_solve(pmgx_cheb* s, pmgx_operator* A, double* x, const double* b, double* resid_hist_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(s && A && x && b, "cheb_solve: null argument");
  PMGX_REQUIRE(A->n_owned == s->n_owned && A->n_ghost == s->n_ghost, "Incompatible vector sizes");
  PMGX_CUDA(cudaSetDevice(s->ctx->device));
  pmgx::cheb_solve(s, A, x, b, resid_hist_h);
  PMGX_API_END
}
int pmgx_cheb_residual(pmgx_cheb* s, pmgx_operator* A, double* x, const double* b, double* rnorm_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(s && A && x && b && rnorm_h, "cheb_residual: null argument");
  PMGX_REQUIRE(A->n_owned == s->n_owned && A->n_ghost == s->n_ghost, "Incompatible vector sizes");
  PMGX_CUDA(cudaSetDevice(s->ctx->device));
  A->apply(x, s->q.p);                                                  // chebyshev.hpp:40
  pmgx::vec::axpy(s->ctx, s->r.p, -1.0, s->q.p, b, s->n_owned);         // :41
  *rnorm_h = std::sqrt(pmgx::vec::dot(s->ctx, s->r.p, s->r.p, s->n_owned));
  PMGX_API_END
}
int pmgx_cheb_destroy(pmgx_cheb* s)
{
  PMGX_API_BEGIN
  if (s)
  {
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    delete s;
  }
  PMGX_API_END
}

// ---- CG
int pmgx_cg_create(pmgx_ctx* ctx, int n_owned, int n_ghost, pmgx_cg** out)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(ctx && out && n_owned >= 0 && n_ghost >= 0, "cg_create: bad arguments");
  PMGX_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<pmgx_cg> s(new pmgx_cg());
  s->ctx = ctx;
  s->n_owned = n_owned;
  s->n_ghost = n_ghost;
  const size_t nt = (size_t)n_owned + n_ghost;
  s->r.alloc(nt);
  s->y.alloc(nt);
  s->p.alloc(nt);
  if (nt > 0)
  {
    PMGX_CUDA(cudaMemsetAsync(s->r.p, 0, nt * sizeof(double), ctx->stream));
    PMGX_CUDA(cudaMemsetAsync(s->y.p, 0, nt * sizeof(double), ctx->stream));
    PMGX_CUDA(cudaMemsetAsync(s->p.p, 0, nt * sizeof(double), ctx->stream));
  }
  *out = s.release();
  PMGX_API_END
}
int pmgx_cg_set_max_iterations(pmgx_cg* s, int max_iter)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(s && max_iter >= 0, "cg_set_max_iterations: bad arguments");
  s->max_iter = max_iter;
  s->alphas.reserve(max_iter);
  s->betas.reserve(max_iter);
  s->residuals.reserve(max_iter);
  PMGX_API_END
}
int pmgx_cg_set_tolerance(pmgx_cg* s, double rtol)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(s, "cg_set_tolerance: null solver");
  s->rtol = rtol;
  PMGX_API_END
}
int pmgx_cg_store_coefficients(pmgx_cg* s, int on)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(s, "cg_store_coefficients: null solver");
  s->store = on != 0;
  PMGX_API_END
}
int pmgx_cg_solve(pmgx_cg* s, pmgx_operator* A, double* x, const double* b, int* iters_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(s && A && x && b, "cg_solve: null argument");
  PMGX_REQUIRE(A->n_owned == s->n_owned && A->n_ghost == s->n_ghost, "Incompatible vector sizes");
  PMGX_CUDA(cudaSetDevice(s->ctx->device));
  const int k = pmgx::cg_solve(s, A, x, b);
  if (iters_h)
    *iters_h = k;
  PMGX_API_END
}
int pmgx_cg_num_coefficients(pmgx_cg* s) { return s ? (int)s->alphas.size() : -1; }
int pmgx_cg_get_coefficients(pmgx_cg* s, double* alphas_h, double* betas_h, double* residuals_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(s, "cg_get_coefficients: null solver");
  for (size_t i = 0; i < s->alphas.size(); ++i)
  {
    if (alphas_h)
      alphas_h[i] = s->alphas[i];
    if (betas_h)
      betas_h[i] = s->betas[i];
    if (residuals_h)
      residuals_h[i] = s->residuals[i];
  }
  PMGX_API_END
}
int pmgx_cg_get_history(pmgx_cg* s, double* rnorm0_h, double* rnorms_h, int* n_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(s, "cg_get_history: null solver");
  if (rnorm0_h)
    *rnorm0_h = s->rnorm0;
  if (n_h)
    *n_h = (int)s->history.size();
  if (rnorms_h)
    for (size_t i = 0; i < s->history.size(); ++i)
      rnorms_h[i] = s->history[i];
  PMGX_API_END
}
int pmgx_cg_compute_eigenvalues(pmgx_cg* s, double* eig_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(s && eig_h, "cg_compute_eigenvalues: null argument");
  const int ne = (int)s->alphas.size();
  if (ne < 2)
  {
    pmgx::set_error("Insufficient data to compute eigenvalues"); // src/cg.hpp:125
    return PMGX_ERR_NUMERIC;
  }
  // Lanczos tridiagonal from the CG coefficients (src/cg.hpp:127-135)
  std::vector<double> d(ne, 0.0), e(ne, 0.0);
  for (int i = 0; i < ne; ++i)
    d[i] = 1.0 / s->alphas[i];
  for (int i = 0; i < ne - 1; ++i)
  {
    d[i + 1] += s->betas[i] / s->alphas[i];
    e[i] = std::sqrt(s->betas[i]) / s->alphas[i];
  }
  const int rc = pmgx_tqli(d.data(), e.data(), ne);
  if (rc != PMGX_OK)
    return rc;
  std::sort(d.begin(), d.end());
  for (int i = 0; i < ne; ++i)
    eig_h[i] = d[i];
  PMGX_API_END
}
int pmgx_cg_destroy(pmgx_cg* s)
{
  PMGX_API_BEGIN
  if (s)
  {
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    delete s;
  }
  PMGX_API_END
}

// ---- coarse solver
int pmgx_coarse_create(pmgx_ctx* ctx, pmgx_operator* A, int max_iter, double rtol, pmgx_coarse** out)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(ctx && A && out && max_iter >= 0, "coarse_create: bad arguments");
  std::unique_ptr<pmgx_coarse> cs(new pmgx_coarse());
  cs->ctx = ctx;
  cs->A = A;
  int rc = pmgx_cg_create(ctx, A->n_owned, A->n_ghost, &cs->cg);
  if (rc != PMGX_OK)
    return rc;
  cs->cg->max_iter = max_iter;
  cs->cg->rtol = rtol;
  cs->cg->store = false;
  *out = cs.release();
  PMGX_API_END
}
int pmgx_coarse_solve(pmgx_coarse* cs, double* x, const double* b, int* iters_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(cs && x && b, "coarse_solve: null argument");
  PMGX_CUDA(cudaSetDevice(cs->ctx->device));
  const int k = pmgx::cg_solve_device(cs->cg, cs->A, x, b, 8);
  if (iters_h)
    *iters_h = k;
  PMGX_API_END
}
int pmgx_coarse_destroy(pmgx_coarse* cs)
{
  PMGX_API_BEGIN
  if (cs)
  {
    pmgx_cg_destroy(cs->cg);
    delete cs;
  }
  PMGX_API_END
}

// ---- V-cycle
int pmgx_vcycle_create(pmgx_ctx* ctx, int n_levels, pmgx_operator** ops, pmgx_cheb** smoothers,
                       pmgx_interp** interps, const int8_t** bc_markers, pmgx_coarse* coarse, int flags,
                       pmgx_vcycle** out)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(ctx && out && n_levels >= 1 && ops && smoothers && bc_markers, "vcycle_create: bad arguments");
  PMGX_REQUIRE(n_levels == 1 || interps, "vcycle_create: interpolators missing");
  PMGX_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<pmgx_vcycle> v(new pmgx_vcycle());
  v->ctx = ctx;
  v->n_levels = n_levels;
  v->flags = flags;
  v->coarse = coarse;
  v->u.resize(n_levels);
  v->r.resize(n_levels);
  v->b.resize(n_levels);
  v->du.resize(n_levels);
  for (int i = 0; i < n_levels; ++i)
  {
    PMGX_REQUIRE(ops[i] && smoothers[i] && bc_markers[i], "vcycle_create: null level %d", i);
    PMGX_REQUIRE(smoothers[i]->n_owned == ops[i]->n_owned, "Incompatible vector sizes");
    v->ops.push_back(ops[i]);
    v->smoothers.push_back(smoothers[i]);
    v->bc.push_back(bc_markers[i]);
    if (i < n_levels - 1)
    {
      PMGX_REQUIRE(interps[i], "vcycle_create: null interpolator %d", i);
      v->interps.push_back(interps[i]);
    }
    const size_t nt = (size_t)ops[i]->n_owned + ops[i]->n_ghost;
    for (auto* buf : {&v->u[i], &v->r[i], &v->b[i], &v->du[i]})
    {
      buf->alloc(nt);
      if (nt > 0)
        PMGX_CUDA(cudaMemsetAsync(buf->p, 0, nt * sizeof(double), ctx->stream));
    }
  }
  *out = v.release();
  PMGX_API_END
}

int pmgx_vcycle_apply(pmgx_vcycle* v, const double* b_in, double* u_inout, double* rnorm_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(v && b_in && u_inout, "vcycle_apply: null argument");
  pmgx_ctx* c = v->ctx;
  PMGX_CUDA(cudaSetDevice(c->device));
  const int nl = v->n_levels;
  const int top = nl - 1;
  const bool diag = (v->flags & PMGX_VC_DIAGNOSTICS) != 0;
  const bool literal_bc = (v->flags & PMGX_VC_LITERAL_REFERENCE_BC) != 0;
  const bool literal_seq = (v->flags & PMGX_VC_LITERAL_SEQUENCE) != 0;
  v->diagnostics.clear();
  namespace vec = pmgx::vec;
  for (int i = 0; i < top; ++i) // pmg.hpp:63-64
    vec::set(c, v->u[i].p, (long long)v->ops[i]->n_owned + v->ops[i]->n_ghost, 0.0);
  vec::copy(c, v->u[top].p, u_inout, v->ops[top]->n_owned); // :65
  vec::copy(c, v->b[top].p, b_in, v->ops[top]->n_owned);    // :68

  for (int i = top; i > 0; --i)
  {
    if (diag)
      v->diagnostics.push_back(pmgx::residual(v, i, true));                       // :76-80
    // below the top level u is zero on the way down: the smoother's first apply is skipped
    pmgx::cheb_solve(v->smoothers[i], v->ops[i], v->u[i].p, v->b[i].p, nullptr, !literal_seq && i < top); // :83
    double* rfine = v->r[i].p;
    if (literal_seq)
    {
      const double rn = pmgx::residual(v, i, diag);                               // :86-89
      if (diag)
        v->diagnostics.push_back(rn);
    }
    else
    {
      // the smoother's recurrence already holds r = b - A u of its final iterate
      // (src/chebyshev.hpp:73-77); the reference recomputes it with one more apply
      rfine = v->smoothers[i]->r.p;
      if (diag)
        v->diagnostics.push_back(std::sqrt(vec::dot(c, rfine, rfine, v->ops[i]->n_owned)));
    }
    int rc = pmgx_interp_restrict(v->interps[i - 1], rfine, v->b[i - 1].p);       // :92
    if (rc != PMGX_OK)
      return rc;
    if (!literal_bc && i - 1 > 0) // quirk Q9: keep Dirichlet rows of intermediate levels clean
      vec::mask_bc(c, v->b[i - 1].p, v->bc[i - 1], v->ops[i - 1]->n_owned);
  }
  vec::mask_bc(c, v->b[0].p, v->bc[0], v->ops[0]->n_owned);                        // :100-103
  if (v->coarse && nl > 1)
  {
    int rc = pmgx_coarse_solve(v->coarse, v->u[0].p, v->b[0].p, nullptr);         // :106-107
    if (rc != PMGX_OK)
      return rc;
  }
  else
    pmgx::cheb_solve(v->smoothers[0], v->ops[0], v->u[0].p, v->b[0].p, nullptr, !literal_seq && nl > 1); // :109
  if (diag)
    v->diagnostics.push_back(pmgx::residual(v, 0, true));                         // :114-117

  for (int i = 0; i < top; ++i)
  {
    int rc = pmgx_interp_prolong(v->interps[i], v->u[i].p, v->du[i + 1].p);       // :123
    if (rc != PMGX_OK)
      return rc;
    vec::axpy(c, v->u[i + 1].p, 1.0, v->u[i + 1].p, v->du[i + 1].p, v->ops[i + 1]->n_owned); // :129
    if (diag)
      v->diagnostics.push_back(pmgx::residual(v, i + 1, true));                   // :132-135
    pmgx::cheb_solve(v->smoothers[i + 1], v->ops[i + 1], v->u[i + 1].p, v->b[i + 1].p, nullptr); // :138
    if (diag && i + 1 < top)
      v->diagnostics.push_back(pmgx::residual(v, i + 1, true));                   // :141-144
  }
  if (rnorm_h || diag)
  {
    const double rn = pmgx::residual(v, top, true);                               // :141-149
    if (rnorm_h)
      *rnorm_h = rn;
    if (diag)
      v->diagnostics.push_back(rn);
  }
  vec::copy(c, u_inout, v->u[top].p, v->ops[top]->n_owned);                        // :154
  PMGX_API_END
}

int pmgx_vcycle_get_diagnostics(pmgx_vcycle* v, double* out_h, int cap, int* n_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(v, "vcycle_get_diagnostics: null handle");
  const int n = (int)v->diagnostics.size();
  if (n_h)
    *n_h = n;
  if (out_h)
    for (int i = 0; i < n && i < cap; ++i)
      out_h[i] = v->diagnostics[i];
  PMGX_API_END
}

int pmgx_vcycle_destroy(pmgx_vcycle* v)
{
  PMGX_API_BEGIN
  if (v)
  {
    cudaSetDevice(v->ctx->device);
    cudaStreamSynchronize(v->ctx->stream);
    delete v;
  }
  PMGX_API_END
}
}
