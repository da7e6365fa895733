// Host set-up of a smoothed-aggregation hierarchy on an assembled CSR matrix: the first half of the
// multilevel coarse solver that is to replace Jacobi-PCG behind CoarseSolverType::solve (the reference
// runs PETSc CG + BoomerAMG there, src/amg.hpp:33-47).  NOT YET USED BY THE V-CYCLE: this file only
// builds and exposes the level matrices (single rank: every column is owned); the device cycle --
// SpMV, the 4th-kind Chebyshev smoother and CSR transfer operators, all of which exist in libpmgx --
// and the distributed Galerkin product come next.  Sized and checked
// against the numpy prototype scripts/prototype_sa_amg.py (tests/test_amg_setup.py, CPU only).
//
// Algorithm per level: greedy (Vanek) aggregation on the graph of the free rows (rows that hold
// only their diagonal -- Dirichlet rows -- stay out of the hierarchy), tentative prolongator T
// (piecewise constant), lambda_max(D^-1 A) by power iteration (x1.1), P = (I - 4/(3 lmax) D^-1 A) T,
// A_c = P^T A P with a row-wise hash SpGEMM; recursion stops at min_coarse rows or max_levels.
#include "common.hpp"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace
{
struct Csr
{
  int n_rows = 0, n_cols = 0;
  std::vector<int32_t> ptr, cols;
  std::vector<double> vals;
  long long nnz() const { return (long long)cols.size(); }
};

// C = A * B, rows sorted by column
Csr spgemm(const Csr& A, const Csr& B)
{
  Csr C;
  C.n_rows = A.n_rows;
  C.n_cols = B.n_cols;
  C.ptr.assign((size_t)A.n_rows + 1, 0);
  std::vector<double> acc((size_t)B.n_cols, 0.0);
  std::vector<int32_t> mark((size_t)B.n_cols, -1), touched;
  for (int i = 0; i < A.n_rows; ++i)
  {
    touched.clear();
    for (int32_t ja = A.ptr[i]; ja < A.ptr[i + 1]; ++ja)
    {
      const int32_t k = A.cols[ja];
      const double a = A.vals[ja];
      for (int32_t jb = B.ptr[k]; jb < B.ptr[k + 1]; ++jb)
      {
        const int32_t c = B.cols[jb];
        if (mark[c] != i)
        {
          mark[c] = i;
          acc[c] = 0.0;
          touched.push_back(c);
        }
        acc[c] += a * B.vals[jb];
      }
    }
    std::sort(touched.begin(), touched.end());
    for (int32_t c : touched)
    {
      C.cols.push_back(c);
      C.vals.push_back(acc[c]);
    }
    C.ptr[i + 1] = (int32_t)C.cols.size();
  }
  return C;
}

Csr transpose(const Csr& A)
{
  Csr T;
  T.n_rows = A.n_cols;
  T.n_cols = A.n_rows;
  T.ptr.assign((size_t)A.n_cols + 1, 0);
  for (int32_t c : A.cols)
    ++T.ptr[c + 1];
  std::partial_sum(T.ptr.begin(), T.ptr.end(), T.ptr.begin());
  T.cols.resize(A.cols.size());
  T.vals.resize(A.vals.size());
  std::vector<int32_t> next(T.ptr.begin(), T.ptr.end() - 1);
  for (int i = 0; i < A.n_rows; ++i)
    for (int32_t j = A.ptr[i]; j < A.ptr[i + 1]; ++j)
    {
      const int32_t p = next[A.cols[j]]++;
      T.cols[p] = i;
      T.vals[p] = A.vals[j];
    }
  return T;
}

std::vector<double> diagonal(const Csr& A)
{
  std::vector<double> d((size_t)A.n_rows, 0.0);
  for (int i = 0; i < A.n_rows; ++i)
    for (int32_t j = A.ptr[i]; j < A.ptr[i + 1]; ++j)
      if (A.cols[j] == i)
        d[i] = A.vals[j];
  return d;
}

// lambda_max(D^-1 A) by power iteration from a fixed pseudo-random start (deterministic)
double lambda_max(const Csr& A, const std::vector<double>& d, int its)
{
  const int n = A.n_rows;
  std::vector<double> x((size_t)n), y((size_t)n);
  unsigned long long s = 0x9E3779B97F4A7C15ull;
  for (int i = 0; i < n; ++i)
  {
    s ^= s << 13, s ^= s >> 7, s ^= s << 17;
    x[i] = (double)(s >> 11) / 9007199254740992.0 - 0.5;
  }
  double lam = 1.0;
  for (int it = 0; it < its; ++it)
  {
    double nrm = 0.0;
    for (int i = 0; i < n; ++i)
    {
      double t = 0.0;
      for (int32_t j = A.ptr[i]; j < A.ptr[i + 1]; ++j)
        t += A.vals[j] * x[A.cols[j]];
      y[i] = t / d[i];
      nrm += y[i] * y[i];
    }
    lam = std::sqrt(nrm);
    if (!(lam > 0.0))
      return 1.0;
    for (int i = 0; i < n; ++i)
      x[i] = y[i] / lam;
  }
  return lam;
}

// Greedy aggregation on the free rows: pass 1 -- a node whose free neighbours are all unaggregated
// founds an aggregate with them; pass 2 -- leftovers join a neighbouring aggregate (or found one).
int aggregate(const Csr& A, const std::vector<char>& is_free, std::vector<int32_t>& agg)
{
  const int n = A.n_rows;
  agg.assign((size_t)n, -1);
  int na = 0;
  for (int i = 0; i < n; ++i)
  {
    if (!is_free[i])
      continue;
    bool ok = true;
    for (int32_t j = A.ptr[i]; j < A.ptr[i + 1] && ok; ++j)
      ok = !is_free[A.cols[j]] || agg[A.cols[j]] < 0;
    if (!ok)
      continue;
    for (int32_t j = A.ptr[i]; j < A.ptr[i + 1]; ++j)
      if (is_free[A.cols[j]])
        agg[A.cols[j]] = na;
    agg[i] = na++;
  }
  for (int i = 0; i < n; ++i)
  {
    if (!is_free[i] || agg[i] >= 0)
      continue;
    for (int32_t j = A.ptr[i]; j < A.ptr[i + 1] && agg[i] < 0; ++j)
      if (agg[A.cols[j]] >= 0)
        agg[i] = agg[A.cols[j]];
    if (agg[i] < 0)
      agg[i] = na++;
  }
  return na;
}

struct Level
{
  Csr A, P; // P: this level -> next coarser one (empty on the coarsest level)
  double lmax = 1.0;
};
} // namespace

struct pmgx_amg_hier
{
  std::vector<Level> levels;
};

extern "C"
{
int pmgx_amg_setup_h(int n_rows, const int32_t* row_ptr_h, const int32_t* cols_h, const double* values_h,
                     int min_coarse, int max_levels, pmgx_amg_hier** out)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(out && row_ptr_h && n_rows >= 0 && max_levels >= 1, "amg_setup: bad arguments");
  std::unique_ptr<pmgx_amg_hier> H(new pmgx_amg_hier());
  Csr A;
  A.n_rows = A.n_cols = n_rows;
  A.ptr.assign(row_ptr_h, row_ptr_h + n_rows + 1);
  const long long nnz = row_ptr_h[n_rows];
  PMGX_REQUIRE(nnz == 0 || (cols_h && values_h), "amg_setup: null matrix arrays");
  A.cols.assign(cols_h, cols_h + nnz);
  A.vals.assign(values_h, values_h + nnz);
  for (int32_t c : A.cols)
    PMGX_REQUIRE(c >= 0 && c < n_rows, "amg_setup: column out of range (single-rank matrices only)");
  while (true)
  {
    Level L;
    L.A = std::move(A);
    const Csr& M = L.A;
    const std::vector<double> d = diagonal(M);
    for (double v : d)
      PMGX_REQUIRE(v > 0.0, "amg_setup: non-positive diagonal entry");
    L.lmax = 1.1 * lambda_max(M, d, 15);
    // free rows: everything except rows holding only their diagonal (Dirichlet rows, src/csr.hpp:84-86)
    std::vector<char> is_free((size_t)M.n_rows, 0);
    int n_free = 0;
    for (int i = 0; i < M.n_rows; ++i)
    {
      is_free[i] = (M.ptr[i + 1] - M.ptr[i]) > 1;
      n_free += is_free[i];
    }
    const bool last = n_free <= min_coarse || (int)H->levels.size() + 1 >= max_levels;
    if (last)
    {
      H->levels.push_back(std::move(L));
      break;
    }
    std::vector<int32_t> agg;
    const int na = aggregate(M, is_free, agg);
    if (na == 0 || na >= n_free)
    {
      H->levels.push_back(std::move(L));
      break;
    }
    // tentative prolongator T, then P = T - omega D^-1 (A T)
    Csr T;
    T.n_rows = M.n_rows;
    T.n_cols = na;
    T.ptr.assign((size_t)M.n_rows + 1, 0);
    for (int i = 0; i < M.n_rows; ++i)
    {
      if (agg[i] >= 0)
      {
        T.cols.push_back(agg[i]);
        T.vals.push_back(1.0);
      }
      T.ptr[i + 1] = (int32_t)T.cols.size();
    }
    Csr AT = spgemm(M, T);
    const double omega = 4.0 / (3.0 * L.lmax);
    Csr P;
    P.n_rows = M.n_rows;
    P.n_cols = na;
    P.ptr.assign((size_t)M.n_rows + 1, 0);
    for (int i = 0; i < M.n_rows; ++i)
    {
      // merge row i of T (at most one entry) with -omega/d_i * row i of AT (sorted)
      const int32_t tc = agg[i];
      bool t_done = tc < 0;
      for (int32_t j = AT.ptr[i]; j < AT.ptr[i + 1]; ++j)
      {
        const int32_t c = AT.cols[j];
        double v = -omega / d[i] * AT.vals[j];
        if (!t_done && tc < c)
        {
          P.cols.push_back(tc);
          P.vals.push_back(1.0);
          t_done = true;
        }
        if (!t_done && tc == c)
        {
          v += 1.0;
          t_done = true;
        }
        P.cols.push_back(c);
        P.vals.push_back(v);
      }
      if (!t_done)
      {
        P.cols.push_back(tc);
        P.vals.push_back(1.0);
      }
      P.ptr[i + 1] = (int32_t)P.cols.size();
    }
    Csr AP = spgemm(M, P);
    A = spgemm(transpose(P), AP); // Galerkin product
    L.P = std::move(P);
    H->levels.push_back(std::move(L));
  }
  *out = H.release();
  PMGX_API_END
}

int pmgx_amg_num_levels(pmgx_amg_hier* h) { return h ? (int)h->levels.size() : -1; }

/* out_h[0] = rows of A, [1] = nnz(A), [2] = columns of P (0 on the coarsest level), [3] = nnz(P) */
int pmgx_amg_level_sizes(pmgx_amg_hier* h, int level, long long* out_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(h && out_h && level >= 0 && level < (int)h->levels.size(), "amg_level_sizes: bad arguments");
  const Level& L = h->levels[level];
  out_h[0] = L.A.n_rows;
  out_h[1] = L.A.nnz();
  out_h[2] = L.P.n_cols;
  out_h[3] = L.P.nnz();
  PMGX_API_END
}

int pmgx_amg_level_get(pmgx_amg_hier* h, int level, int32_t* a_ptr_h, int32_t* a_cols_h, double* a_vals_h,
                       int32_t* p_ptr_h, int32_t* p_cols_h, double* p_vals_h, double* lmax_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(h && level >= 0 && level < (int)h->levels.size(), "amg_level_get: bad arguments");
  const Level& L = h->levels[level];
  if (a_ptr_h)
    std::copy(L.A.ptr.begin(), L.A.ptr.end(), a_ptr_h);
  if (a_cols_h)
    std::copy(L.A.cols.begin(), L.A.cols.end(), a_cols_h);
  if (a_vals_h)
    std::copy(L.A.vals.begin(), L.A.vals.end(), a_vals_h);
  if (p_ptr_h && !L.P.ptr.empty())
    std::copy(L.P.ptr.begin(), L.P.ptr.end(), p_ptr_h);
  if (p_cols_h)
    std::copy(L.P.cols.begin(), L.P.cols.end(), p_cols_h);
  if (p_vals_h)
    std::copy(L.P.vals.begin(), L.P.vals.end(), p_vals_h);
  if (lmax_h)
    *lmax_h = L.lmax;
  PMGX_API_END
}

int pmgx_amg_destroy(pmgx_amg_hier* h)
{
  PMGX_API_BEGIN
  delete h;
  PMGX_API_END
}
}
