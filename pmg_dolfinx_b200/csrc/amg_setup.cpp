// Host set-up of the distributed smoothed-aggregation hierarchy behind CoarseSolverType::solve (the
// reference runs PETSc CG + BoomerAMG there, src/amg.hpp:33-47).  See amg.hpp for the distribution
// model.  Pure host code: the only collective is a fixed-size all-gather of bytes handed in by the
// caller (NCCL in amg.cu, a torch.distributed callback in the CPU tests), everything else is built
// on it -- this is set-up, run once.
//
// Per level: greedy (Vanek) aggregation on the rank-local graph of the free rows (rows that hold
// only their diagonal -- Dirichlet rows, src/csr.hpp:84-86 -- stay out of the hierarchy), tentative
// prolongator T (piecewise constant), lambda_max(D^-1 A) by distributed power iteration (x1.1),
// P = (I - 4/(3 lmax) D^-1 A) T with the full rows of A (exchange of the aggregate ids of the ghost dofs),
// exchange of the P rows of the interface dofs, A_c = P^T A [P; P_ghost] by row-wise SpGEMM with the partial
// rows of foreign aggregates returned to their owners, R = the owned rows of the global P^T, coarse halo plan
// from the ghost aggregates that A_c or P reference.  Recursion stops at min_coarse free rows (globally) or
// max_levels; the last level is gathered and inverted densely.
// Sized and checked against scripts/prototype_sa_amg.py (tests/test_amg_setup.py, tests/test_dist_cpu.py).
#include "common.hpp"
#include "amg.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>

namespace pmgx
{
namespace amg
{
namespace
{
constexpr long long DENSE_CAP = 4096; // largest coarsest level that is inverted densely

// ------------------------------------------------------------- collectives on the all-gather --
std::vector<std::vector<char>> allgather_var(const Comm& cm, const std::vector<char>& mine)
{
  std::vector<std::vector<char>> out((size_t)cm.nranks);
  if (cm.nranks == 1)
  {
    out[0] = mine;
    return out;
  }
  long long sz = (long long)mine.size();
  std::vector<long long> sizes((size_t)cm.nranks, 0);
  cm.allgather(&sz, sizeof(sz), sizes.data());
  const size_t mx = (size_t)std::max<long long>(1, *std::max_element(sizes.begin(), sizes.end()));
  std::vector<char> pad(mx, 0), all(mx * cm.nranks, 0);
  if (!mine.empty())
    std::memcpy(pad.data(), mine.data(), mine.size());
  cm.allgather(pad.data(), mx, all.data());
  for (int r = 0; r < cm.nranks; ++r)
    out[r].assign(all.begin() + (size_t)r * mx, all.begin() + (size_t)r * mx + (size_t)sizes[r]);
  return out;
}

double allreduce_sum(const Comm& cm, double v)
{
  if (cm.nranks == 1)
    return v;
  std::vector<double> all((size_t)cm.nranks, 0.0);
  cm.allgather(&v, sizeof(v), all.data());
  double s = 0.0;
  for (double a : all) // rank order: identical on every rank
    s += a;
  return s;
}

long long allreduce_sum(const Comm& cm, long long v)
{
  if (cm.nranks == 1)
    return v;
  std::vector<long long> all((size_t)cm.nranks, 0);
  cm.allgather(&v, sizeof(v), all.data());
  return std::accumulate(all.begin(), all.end(), 0ll);
}

template <typename T>
void put(std::vector<char>& b, const T* p, size_t n)
{
  const size_t o = b.size();
  b.resize(o + n * sizeof(T));
  if (n)
    std::memcpy(b.data() + o, p, n * sizeof(T));
}
template <typename T>
void put1(std::vector<char>& b, T v)
{
  put(b, &v, 1);
}
struct Reader
{
  const char* p;
  const char* e;
  template <typename T>
  void get(T* out, size_t n)
  {
    if (p + n * sizeof(T) > e)
      throw std::runtime_error("amg set-up: truncated message");
    if (n)
      std::memcpy(out, p, n * sizeof(T));
    p += n * sizeof(T);
  }
  template <typename T>
  T get1()
  {
    T v;
    get(&v, 1);
    return v;
  }
};

// message k goes to rank dests[k]; returns the messages addressed to this rank, sorted by source
std::vector<std::pair<int, std::vector<char>>> neighbor_exchange(const Comm& cm, const std::vector<int>& dests,
                                                                 const std::vector<std::vector<char>>& msgs)
{
  std::vector<char> blob;
  put1<int32_t>(blob, (int32_t)dests.size());
  for (size_t k = 0; k < dests.size(); ++k)
  {
    put1<int32_t>(blob, dests[k]);
    put1<long long>(blob, (long long)msgs[k].size());
  }
  for (size_t k = 0; k < dests.size(); ++k)
    put(blob, msgs[k].data(), msgs[k].size());
  auto all = allgather_var(cm, blob);
  std::vector<std::pair<int, std::vector<char>>> out;
  for (int r = 0; r < cm.nranks; ++r)
  {
    Reader rd{all[r].data(), all[r].data() + all[r].size()};
    const int nm = rd.get1<int32_t>();
    std::vector<int32_t> d((size_t)nm);
    std::vector<long long> sz((size_t)nm);
    for (int k = 0; k < nm; ++k)
    {
      d[k] = rd.get1<int32_t>();
      sz[k] = rd.get1<long long>();
    }
    for (int k = 0; k < nm; ++k)
    {
      if (d[k] == cm.rank)
      {
        std::vector<char> m((size_t)sz[k]);
        rd.get(m.data(), m.size());
        out.emplace_back(r, std::move(m));
      }
      else
        rd.p += sz[k];
    }
  }
  return out;
}

// x[n_owned + slot] <- owner's value (Vector::scatter_fwd on the host)
void halo_exchange(const Comm& cm, const Plan& pl, int n_owned, std::vector<double>& x)
{
  if (cm.nranks == 1)
    return;
  std::vector<std::vector<char>> msgs(pl.send_ranks.size());
  for (size_t k = 0; k < pl.send_ranks.size(); ++k)
    for (int t = pl.send_offsets[k]; t < pl.send_offsets[k + 1]; ++t)
      put1<double>(msgs[k], x[pl.send_idx[t]]);
  auto in = neighbor_exchange(cm, pl.send_ranks, msgs);
  for (auto& m : in)
  {
    const auto it = std::find(pl.recv_ranks.begin(), pl.recv_ranks.end(), m.first);
    if (it == pl.recv_ranks.end())
      continue;
    const size_t k = it - pl.recv_ranks.begin();
    Reader rd{m.second.data(), m.second.data() + m.second.size()};
    for (int t = pl.recv_offsets[k]; t < pl.recv_offsets[k + 1]; ++t)
      x[(size_t)n_owned + pl.recv_idx[t]] = rd.get1<double>();
  }
}

// ------------------------------------------------------------------------ sparse kernels --
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
double lambda_max(const Level& L, const std::vector<double>& d, const Comm& cm, int its)
{
  const Csr& A = L.A;
  const int n = L.n_owned;
  std::vector<double> x((size_t)n + L.n_ghost, 0.0), y((size_t)n);
  unsigned long long s = 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull * (unsigned long long)cm.rank;
  for (int i = 0; i < n; ++i)
  {
    s ^= s << 13, s ^= s >> 7, s ^= s << 17;
    x[i] = (double)(s >> 11) / 9007199254740992.0 - 0.5;
  }
  double lam = 1.0;
  for (int it = 0; it < its; ++it)
  {
    halo_exchange(cm, L.plan, n, x);
    double nrm = 0.0;
    for (int i = 0; i < n; ++i)
    {
      double t = 0.0;
      for (int32_t j = A.ptr[i]; j < A.ptr[i + 1]; ++j)
        t += A.vals[j] * x[A.cols[j]];
      y[i] = t / d[i];
      nrm += y[i] * y[i];
    }
    lam = std::sqrt(allreduce_sum(cm, nrm));
    if (!(lam > 0.0))
      return 1.0;
    for (int i = 0; i < n; ++i)
      x[i] = y[i] / lam;
  }
  return lam;
}

// Greedy aggregation on the rank-local graph of the free rows: pass 1 -- a node whose free
// neighbours are all unaggregated founds an aggregate with them; pass 2 -- leftovers join a
// neighbouring aggregate (or found one).  Ghost columns are ignored: aggregates never span ranks.
int aggregate(const Csr& A, int n_owned, const std::vector<char>& is_free, std::vector<int32_t>& agg)
{
  const int n = n_owned;
  agg.assign((size_t)n, -1);
  int na = 0;
  for (int i = 0; i < n; ++i)
  {
    if (!is_free[i])
      continue;
    bool ok = true;
    for (int32_t j = A.ptr[i]; j < A.ptr[i + 1] && ok; ++j)
    {
      const int32_t c = A.cols[j];
      ok = c >= n || !is_free[c] || agg[c] < 0;
    }
    if (!ok)
      continue;
    for (int32_t j = A.ptr[i]; j < A.ptr[i + 1]; ++j)
      if (A.cols[j] < n && is_free[A.cols[j]])
        agg[A.cols[j]] = na;
    agg[i] = na++;
  }
  for (int i = 0; i < n; ++i)
  {
    if (!is_free[i] || agg[i] >= 0)
      continue;
    for (int32_t j = A.ptr[i]; j < A.ptr[i + 1] && agg[i] < 0; ++j)
      if (A.cols[j] < n && agg[A.cols[j]] >= 0)
        agg[i] = agg[A.cols[j]];
    if (agg[i] < 0)
      agg[i] = na++;
  }
  return na;
}

// Cholesky factorisation in place (lower triangle); false if the matrix is not positive definite
bool cholesky(std::vector<double>& M, int n)
{
  for (int j = 0; j < n; ++j)
  {
    double* rj = &M[(size_t)j * n];
    double s = rj[j];
    for (int k = 0; k < j; ++k)
      s -= rj[k] * rj[k];
    if (!(s > 0.0))
      return false;
    const double ljj = std::sqrt(s);
    rj[j] = ljj;
    for (int i = j + 1; i < n; ++i)
    {
      double* ri = &M[(size_t)i * n];
      double t = ri[j];
      for (int k = 0; k < j; ++k)
        t -= ri[k] * rj[k];
      ri[j] = t / ljj;
    }
  }
  return true;
}

// gather the level on every rank, invert it, keep this rank's rows (columns in local layout)
void finish_coarsest(Level& L, const Comm& cm)
{
  std::vector<long long> counts((size_t)cm.nranks, 0), off((size_t)cm.nranks + 1, 0);
  long long mine = L.n_owned;
  if (cm.nranks == 1)
    counts[0] = mine;
  else
    cm.allgather(&mine, sizeof(mine), counts.data());
  for (int r = 0; r < cm.nranks; ++r)
    off[r + 1] = off[r] + counts[r];
  const long long ng = off[cm.nranks];
  L.n_global = ng;
  L.dense = false;
  if (ng == 0 || ng > DENSE_CAP)
    return;
  std::vector<char> blob;
  put1<long long>(blob, L.A.nnz());
  for (int i = 0; i < L.n_owned; ++i)
    for (int32_t j = L.A.ptr[i]; j < L.A.ptr[i + 1]; ++j)
    {
      const int32_t c = L.A.cols[j];
      const long long gc = c < L.n_owned ? off[cm.rank] + c : off[L.ghost_src[c - L.n_owned]] + L.ghost_rid[c - L.n_owned];
      put1<long long>(blob, off[cm.rank] + i);
      put1<long long>(blob, gc);
      put1<double>(blob, L.A.vals[j]);
    }
  auto all = allgather_var(cm, blob);
  const int n = (int)ng;
  std::vector<double> M((size_t)n * n, 0.0);
  for (auto& b : all)
  {
    Reader rd{b.data(), b.data() + b.size()};
    const long long nz = rd.get1<long long>();
    for (long long t = 0; t < nz; ++t)
    {
      const long long r = rd.get1<long long>(), c = rd.get1<long long>();
      const double v = rd.get1<double>();
      M[(size_t)r * n + c] += v;
    }
  }
  // symmetrise against rounding differences between the ranks' copies of the same entry
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j)
    {
      const double v = 0.5 * (M[(size_t)i * n + j] + M[(size_t)j * n + i]);
      M[(size_t)i * n + j] = M[(size_t)j * n + i] = v;
    }
  if (!cholesky(M, n))
    return; // not SPD (should not happen): the caller smooths on this level instead
  // rows [off[rank], off[rank] + n_owned) of M^-1: solve L L^T z = e_row
  L.inv_rows.assign((size_t)L.n_owned * n, 0.0);
  std::vector<int> perm((size_t)n); // local column -> global column
  {
    int p = 0;
    for (long long g = off[cm.rank]; g < off[cm.rank + 1]; ++g)
      perm[p++] = (int)g;
    for (int r = 0; r < cm.nranks; ++r)
      if (r != cm.rank)
        for (long long g = off[r]; g < off[r + 1]; ++g)
          perm[p++] = (int)g;
  }
  std::vector<double> z((size_t)n);
#pragma omp parallel for firstprivate(z) schedule(static)
  for (int row = 0; row < L.n_owned; ++row)
  {
    const int g = (int)off[cm.rank] + row;
    std::fill(z.begin(), z.end(), 0.0);
    for (int i = g; i < n; ++i) // forward: entries before g stay zero
    {
      double t = (i == g) ? 1.0 : 0.0;
      const double* ri = &M[(size_t)i * n];
      for (int k = g; k < i; ++k)
        t -= ri[k] * z[k];
      z[i] = t / ri[i];
    }
    for (int i = n - 1; i >= 0; --i) // backward with L^T
    {
      double t = z[i];
      for (int k = i + 1; k < n; ++k)
        t -= M[(size_t)k * n + i] * z[k];
      z[i] = t / M[(size_t)i * n + i];
    }
    double* out = &L.inv_rows[(size_t)row * n];
    for (int c = 0; c < n; ++c)
      out[c] = z[perm[c]];
  }
  // all-gather plan: every rank sends all of its owned entries to every other rank
  Plan& gp = L.gather_plan;
  gp = Plan();
  int slot = 0;
  for (int r = 0; r < cm.nranks; ++r)
  {
    if (r == cm.rank)
      continue;
    gp.send_ranks.push_back(r);
    for (int i = 0; i < L.n_owned; ++i)
      gp.send_idx.push_back(i);
    gp.send_offsets.push_back((int)gp.send_idx.size());
    gp.recv_ranks.push_back(r);
    for (long long t = 0; t < counts[r]; ++t)
      gp.recv_idx.push_back(slot++);
    gp.recv_offsets.push_back((int)gp.recv_idx.size());
  }
  L.dense = true;
}

void check_plan(const Plan& p, int n_owned, int n_ghost)
{
  PMGX_REQUIRE(p.send_offsets.size() == p.send_ranks.size() + 1 && p.recv_offsets.size() == p.recv_ranks.size() + 1,
               "amg_setup: malformed halo plan");
  PMGX_REQUIRE((size_t)p.send_offsets.back() == p.send_idx.size() && (size_t)p.recv_offsets.back() == p.recv_idx.size(),
               "amg_setup: halo plan offsets do not match the index lists");
  for (int32_t v : p.send_idx)
    PMGX_REQUIRE(v >= 0 && v < n_owned, "amg_setup: send index out of range");
  for (int32_t v : p.recv_idx)
    PMGX_REQUIRE(v >= 0 && v < n_ghost, "amg_setup: ghost slot out of range");
}
} // namespace


Csr drop_stored_zeros(const Csr& A)
{
  Csr B;
  B.n_rows = A.n_rows;
  B.n_cols = A.n_cols;
  B.ptr.assign((size_t)A.n_rows + 1, 0);
  for (int i = 0; i < A.n_rows; ++i)
  {
    int32_t kept = 0;
    for (int32_t j = A.ptr[i]; j < A.ptr[i + 1]; ++j)
      kept += (A.vals[j] != 0.0 || A.cols[j] == i) ? 1 : 0;
    B.ptr[(size_t)i + 1] = B.ptr[i] + kept;
  }
  B.cols.resize((size_t)B.ptr[A.n_rows]);
  B.vals.resize((size_t)B.ptr[A.n_rows]);
  for (int i = 0; i < A.n_rows; ++i)
  {
    int32_t o = B.ptr[i];
    for (int32_t j = A.ptr[i]; j < A.ptr[i + 1]; ++j)
      if (A.vals[j] != 0.0 || A.cols[j] == i)
      {
        B.cols[o] = A.cols[j];
        B.vals[o] = A.vals[j];
        ++o;
      }
  }
  return B;
}

void setup(Hierarchy& H, Csr A0, int n_owned, int n_ghost, const Plan& plan0, const Comm& cm, int min_coarse,
           int max_levels)
{
  H.levels.clear();
  check_plan(plan0, n_owned, n_ghost);
  Level L;
  L.A = std::move(A0);
  L.A.n_rows = n_owned;
  L.A.n_cols = n_owned + n_ghost;
  L.n_owned = n_owned;
  L.n_ghost = n_ghost;
  L.plan = plan0;
  for (int32_t c : L.A.cols)
    PMGX_REQUIRE(c >= 0 && c < n_owned + n_ghost, "amg_setup: column out of range");
  // who owns my ghosts, and under which index: every rank tells its neighbours its send lists
  L.ghost_src.assign((size_t)n_ghost, -1);
  L.ghost_rid.assign((size_t)n_ghost, -1);
  if (cm.nranks > 1)
  {
    std::vector<std::vector<char>> msgs(plan0.send_ranks.size());
    for (size_t k = 0; k < plan0.send_ranks.size(); ++k)
      put(msgs[k], plan0.send_idx.data() + plan0.send_offsets[k], (size_t)(plan0.send_offsets[k + 1] - plan0.send_offsets[k]));
    for (auto& m : neighbor_exchange(cm, plan0.send_ranks, msgs))
    {
      const auto it = std::find(plan0.recv_ranks.begin(), plan0.recv_ranks.end(), m.first);
      PMGX_REQUIRE(it != plan0.recv_ranks.end(), "amg_setup: rank %d sends to rank %d, which does not expect it", m.first,
                   cm.rank);
      const size_t k = it - plan0.recv_ranks.begin();
      const size_t cnt = (size_t)(plan0.recv_offsets[k + 1] - plan0.recv_offsets[k]);
      PMGX_REQUIRE(m.second.size() == cnt * sizeof(int32_t), "amg_setup: send / receive counts of ranks %d -> %d differ",
                   m.first, cm.rank);
      Reader rd{m.second.data(), m.second.data() + m.second.size()};
      for (size_t t = 0; t < cnt; ++t)
      {
        const int32_t slot = plan0.recv_idx[plan0.recv_offsets[k] + t];
        L.ghost_src[slot] = m.first;
        L.ghost_rid[slot] = rd.get1<int32_t>();
      }
    }
  }

  const long long repl_cap = getenv("PMGX_AMG_REPL_CAP") ? atoll(getenv("PMGX_AMG_REPL_CAP")) : 65536;
  while (true)
  {
    // ---- a level that is small on the whole machine is REPLICATED: every rank gathers the complete level
    // matrix and builds the rest of the hierarchy for itself (deterministic, hence identical everywhere).
    // Below this point a cycle needs no communication at all -- one all-gather of the restricted residual
    // replaces ~6 latency-bound halo exchanges per level and application.
    if (cm.nranks > 1 && !H.levels.empty())
    {
      std::vector<long long> counts((size_t)cm.nranks, 0), off((size_t)cm.nranks + 1, 0);
      long long mine = L.n_owned;
      cm.allgather(&mine, sizeof(mine), counts.data());
      for (int r = 0; r < cm.nranks; ++r)
        off[r + 1] = off[r] + counts[r];
      const long long ng = off[cm.nranks];
      if (ng > 0 && ng <= repl_cap)
      {
        auto canon = [&](int32_t c) -> long long
        { return c < L.n_owned ? off[cm.rank] + c : off[L.ghost_src[c - L.n_owned]] + L.ghost_rid[c - L.n_owned]; };
        std::vector<char> blob;
        put1<long long>(blob, L.A.nnz());
        for (int i = 0; i < L.n_owned; ++i)
          for (int32_t j = L.A.ptr[i]; j < L.A.ptr[i + 1]; ++j)
          {
            put1<int32_t>(blob, (int32_t)(off[cm.rank] + i));
            put1<int32_t>(blob, (int32_t)canon(L.A.cols[j]));
            put1<double>(blob, L.A.vals[j]);
          }
        auto all = allgather_var(cm, blob);
        std::vector<std::map<int32_t, double>> rows((size_t)ng);
        for (auto& bl : all)
        {
          Reader rd{bl.data(), bl.data() + bl.size()};
          const long long nz = rd.get1<long long>();
          for (long long t = 0; t < nz; ++t)
          {
            const int32_t r = rd.get1<int32_t>(), c = rd.get1<int32_t>();
            rows[r][c] += rd.get1<double>();
          }
        }
        Csr Ag;
        Ag.n_rows = Ag.n_cols = (int)ng;
        Ag.ptr.assign((size_t)ng + 1, 0);
        for (long long r = 0; r < ng; ++r)
        {
          for (auto& e : rows[r])
          {
            Ag.cols.push_back(e.first);
            Ag.vals.push_back(e.second);
          }
          Ag.ptr[r + 1] = (int32_t)Ag.cols.size();
        }
        // the parent's prolongator now addresses the canonical (rank-ordered) numbering of the gathered level
        Level& parent = H.levels.back();
        {
          std::vector<std::pair<int32_t, double>> row;
          for (int i = 0; i < parent.P.n_rows; ++i)
          {
            row.clear();
            for (int32_t j = parent.P.ptr[i]; j < parent.P.ptr[i + 1]; ++j)
              row.emplace_back((int32_t)canon(parent.P.cols[j]), parent.P.vals[j]);
            std::sort(row.begin(), row.end(), [](const std::pair<int32_t, double>& a, const std::pair<int32_t, double>& b)
                      { return a.first < b.first; });
            for (int32_t j = parent.P.ptr[i], t = 0; j < parent.P.ptr[i + 1]; ++j, ++t)
              parent.P.cols[j] = row[t].first, parent.P.vals[j] = row[t].second;
          }
          parent.P.n_cols = (int)ng;
        }
        Hierarchy sub;
        Comm self;
        setup(sub, std::move(Ag), (int)ng, 0, Plan(), self, min_coarse, std::max(1, max_levels - (int)H.levels.size()));
        Level& first = sub.levels.front();
        first.n_mine = L.n_owned;
        // all-gather plan of the restricted residual (every rank sends its owned entries to every other rank) and
        // the position of every canonical index in this rank's gather layout [mine | the others in rank order]
        int slot = 0;
        first.gather_perm.assign((size_t)ng, 0);
        for (int i = 0; i < L.n_owned; ++i)
          first.gather_perm[(size_t)(off[cm.rank] + i)] = i;
        for (int r = 0; r < cm.nranks; ++r)
        {
          if (r == cm.rank)
            continue;
          first.repl_plan.send_ranks.push_back(r);
          for (int i = 0; i < L.n_owned; ++i)
            first.repl_plan.send_idx.push_back(i);
          first.repl_plan.send_offsets.push_back((int)first.repl_plan.send_idx.size());
          first.repl_plan.recv_ranks.push_back(r);
          for (long long t = 0; t < counts[r]; ++t)
          {
            first.gather_perm[(size_t)(off[r] + t)] = L.n_owned + slot;
            first.repl_plan.recv_idx.push_back(slot++);
          }
          first.repl_plan.recv_offsets.push_back((int)first.repl_plan.recv_idx.size());
        }
        for (Level& S : sub.levels)
        {
          S.replicated = true;
          H.levels.push_back(std::move(S));
        }
        return;
      }
    }
    const Csr& M = L.A;
    const int n = L.n_owned;
    const std::vector<double> d = diagonal(M);
    for (double v : d)
      PMGX_REQUIRE(v > 0.0, "amg_setup: non-positive diagonal entry");
    L.lmax = 1.1 * lambda_max(L, d, cm, 15);
    // free rows: everything except rows holding only their diagonal (Dirichlet rows, src/csr.hpp:84-86)
    std::vector<char> is_free((size_t)n, 0);
    long long n_free = 0;
    for (int i = 0; i < n; ++i)
    {
      is_free[i] = (M.ptr[i + 1] - M.ptr[i]) > 1;
      n_free += is_free[i];
    }
    const long long n_free_g = allreduce_sum(cm, n_free);
    bool last = n_free_g <= min_coarse || (int)H.levels.size() + 1 >= max_levels;
    std::vector<int32_t> agg;
    int na = 0;
    if (!last)
    {
      na = aggregate(M, n, is_free, agg);
      const long long na_g = allreduce_sum(cm, (long long)na);
      last = na_g == 0 || na_g >= n_free_g;
    }
    if (last)
    {
      finish_coarsest(L, cm);
      H.levels.push_back(std::move(L));
      break;
    }
    // ---- smoothed prolongator P = (I - omega D^-1 A) T with the FULL rows of A: aggregates are rank-local,
    // but a row next to a partition interface also interpolates from the neighbour's aggregates (with the
    // rank-local part of A only, the interface behaves like unsmoothed aggregation: 18 instead of ~10 PCG
    // iterations on 8 ranks).  Coarse unknowns are named by (owner rank, owner-local aggregate id) on the wire.
    using Key = std::pair<int, int32_t>;
    std::map<Key, int32_t> gmap; // ghost aggregates seen so far -> provisional local id (>= na, in discovery order)
    std::vector<Key> gkeys;
    auto local_id = [&](int owner, int32_t id) -> int32_t
    {
      if (owner == cm.rank)
        return id;
      auto it = gmap.find({owner, id});
      if (it == gmap.end())
      {
        it = gmap.emplace(Key{owner, id}, (int32_t)(na + (int)gkeys.size())).first;
        gkeys.push_back({owner, id});
      }
      return it->second;
    };
    auto key_of = [&](int32_t c) -> Key { return c < na ? Key{cm.rank, c} : gkeys[(size_t)(c - na)]; };
    // aggregate of every ghost dof: the owners send agg[] of the dofs on their send lists
    std::vector<int32_t> gagg((size_t)L.n_ghost, -1);
    if (cm.nranks > 1)
    {
      std::vector<std::vector<char>> msgs(L.plan.send_ranks.size());
      for (size_t k = 0; k < L.plan.send_ranks.size(); ++k)
        for (int t = L.plan.send_offsets[k]; t < L.plan.send_offsets[k + 1]; ++t)
          put1<int32_t>(msgs[k], agg[L.plan.send_idx[t]]);
      for (auto& m : neighbor_exchange(cm, L.plan.send_ranks, msgs))
      {
        const auto it = std::find(L.plan.recv_ranks.begin(), L.plan.recv_ranks.end(), m.first);
        if (it == L.plan.recv_ranks.end())
          continue;
        const size_t k = it - L.plan.recv_ranks.begin();
        Reader rd{m.second.data(), m.second.data() + m.second.size()};
        for (int t = L.plan.recv_offsets[k]; t < L.plan.recv_offsets[k + 1]; ++t)
        {
          const int32_t a = rd.get1<int32_t>();
          gagg[L.plan.recv_idx[t]] = a < 0 ? -1 : local_id(m.first, a);
        }
      }
    }
    // T over owned + ghost rows (at most one entry per row), then P for the owned rows
    const double omega = 4.0 / (3.0 * L.lmax);
    Csr P;
    P.n_rows = n;
    P.ptr.assign((size_t)n + 1, 0);
    {
      std::vector<std::pair<int32_t, double>> row;
      for (int i = 0; i < n; ++i)
      {
        row.clear();
        if (agg[i] >= 0)
          row.emplace_back(agg[i], 1.0);
        if (is_free[i])
          for (int32_t j = M.ptr[i]; j < M.ptr[i + 1]; ++j)
          {
            const int32_t cj = M.cols[j];
            const int32_t a = cj < n ? agg[cj] : gagg[cj - n];
            if (a >= 0)
              row.emplace_back(a, -omega / d[i] * M.vals[j]);
          }
        std::sort(row.begin(), row.end(), [](const std::pair<int32_t, double>& x, const std::pair<int32_t, double>& y)
                  { return x.first < y.first; });
        for (size_t t = 0; t < row.size(); ++t)
        {
          if (!P.cols.empty() && (int32_t)P.cols.size() > P.ptr[i] && P.cols.back() == row[t].first)
            P.vals.back() += row[t].second;
          else
          {
            P.cols.push_back(row[t].first);
            P.vals.push_back(row[t].second);
          }
        }
        P.ptr[i + 1] = (int32_t)P.cols.size();
      }
    }
    // prolongator rows of my ghost dofs: the owners send the rows of the dofs on their send lists, every
    // column as (owner, id)
    struct SRow
    {
      std::vector<int32_t> cols;
      std::vector<double> vals;
    };
    std::vector<SRow> grow((size_t)L.n_ghost);
    if (cm.nranks > 1)
    {
      std::vector<std::vector<char>> msgs(L.plan.send_ranks.size());
      for (size_t k = 0; k < L.plan.send_ranks.size(); ++k)
        for (int t = L.plan.send_offsets[k]; t < L.plan.send_offsets[k + 1]; ++t)
        {
          const int32_t i = L.plan.send_idx[t];
          const int32_t len = P.ptr[i + 1] - P.ptr[i];
          put1<int32_t>(msgs[k], len);
          for (int32_t j = P.ptr[i]; j < P.ptr[i + 1]; ++j)
          {
            const Key kk = key_of(P.cols[j]);
            put1<int32_t>(msgs[k], kk.first);
            put1<int32_t>(msgs[k], kk.second);
            put1<double>(msgs[k], P.vals[j]);
          }
        }
      for (auto& m : neighbor_exchange(cm, L.plan.send_ranks, msgs))
      {
        const auto it = std::find(L.plan.recv_ranks.begin(), L.plan.recv_ranks.end(), m.first);
        if (it == L.plan.recv_ranks.end())
          continue;
        const size_t k = it - L.plan.recv_ranks.begin();
        Reader rd{m.second.data(), m.second.data() + m.second.size()};
        for (int t = L.plan.recv_offsets[k]; t < L.plan.recv_offsets[k + 1]; ++t)
        {
          SRow& g = grow[L.plan.recv_idx[t]];
          const int32_t len = rd.get1<int32_t>();
          std::vector<std::pair<int32_t, double>> row((size_t)len);
          for (int32_t e = 0; e < len; ++e)
          {
            const int o = rd.get1<int32_t>();
            const int32_t id = rd.get1<int32_t>();
            row[e] = {local_id(o, id), rd.get1<double>()};
          }
          std::sort(row.begin(), row.end(), [](const std::pair<int32_t, double>& x, const std::pair<int32_t, double>& y)
                    { return x.first < y.first; });
          for (auto& e : row)
          {
            g.cols.push_back(e.first);
            g.vals.push_back(e.second);
          }
        }
      }
    }
    // P over owned + ghost rows; A P for the owned rows; P^T (A P): rows < na are mine, the others are
    // partial rows of the neighbours' aggregates and go to their owners
    Csr Pext;
    Pext.n_rows = n + L.n_ghost;
    Pext.ptr = P.ptr;
    Pext.cols = P.cols;
    Pext.vals = P.vals;
    Pext.ptr.resize((size_t)n + L.n_ghost + 1);
    for (int g = 0; g < L.n_ghost; ++g)
    {
      Pext.cols.insert(Pext.cols.end(), grow[g].cols.begin(), grow[g].cols.end());
      Pext.vals.insert(Pext.vals.end(), grow[g].vals.begin(), grow[g].vals.end());
      Pext.ptr[(size_t)n + g + 1] = (int32_t)Pext.cols.size();
    }
    Pext.n_cols = na + (int)gkeys.size();
    P.n_cols = Pext.n_cols;
    Csr AP = spgemm(M, Pext);
    Csr Call = spgemm(transpose(P), AP); // (na + ghosts) x (na + ghosts)
    std::vector<std::map<int32_t, double>> crow((size_t)na); // my coarse rows, accumulated
    for (int I = 0; I < na; ++I)
      for (int32_t j = Call.ptr[I]; j < Call.ptr[I + 1]; ++j)
        crow[I][Call.cols[j]] += Call.vals[j];
    if (cm.nranks > 1)
    {
      std::map<int, std::vector<char>> out;
      for (int I = na; I < Call.n_rows; ++I)
      {
        const int32_t len = Call.ptr[I + 1] - Call.ptr[I];
        if (len == 0)
          continue;
        const Key ki = gkeys[(size_t)(I - na)];
        std::vector<char>& b = out[ki.first];
        put1<int32_t>(b, ki.second);
        put1<int32_t>(b, len);
        for (int32_t j = Call.ptr[I]; j < Call.ptr[I + 1]; ++j)
        {
          const Key kc = key_of(Call.cols[j]);
          put1<int32_t>(b, kc.first);
          put1<int32_t>(b, kc.second);
          put1<double>(b, Call.vals[j]);
        }
      }
      std::vector<int> dests;
      std::vector<std::vector<char>> msgs;
      for (auto& kv : out)
      {
        dests.push_back(kv.first);
        msgs.push_back(std::move(kv.second));
      }
      for (auto& m : neighbor_exchange(cm, dests, msgs))
      {
        Reader rd{m.second.data(), m.second.data() + m.second.size()};
        while (rd.p < rd.e)
        {
          const int32_t I = rd.get1<int32_t>();
          const int32_t len = rd.get1<int32_t>();
          PMGX_REQUIRE(I >= 0 && I < na, "amg_setup: partial coarse row %d of %d from rank %d", I, na, m.first);
          for (int32_t e = 0; e < len; ++e)
          {
            const int o = rd.get1<int32_t>();
            const int32_t id = rd.get1<int32_t>();
            crow[I][local_id(o, id)] += rd.get1<double>();
          }
        }
      }
    }
    // final coarse numbering: the ghost aggregates that A_c or P reference, sorted by (owner, id)
    const int n_prov = (int)gkeys.size();
    std::vector<char> used((size_t)n_prov, 0);
    for (int I = 0; I < na; ++I)
      for (auto& e : crow[I])
        if (e.first >= na)
          used[e.first - na] = 1;
    for (int32_t c : P.cols)
      if (c >= na)
        used[c - na] = 1;
    std::vector<int32_t> order;
    for (int g = 0; g < n_prov; ++g)
      if (used[g])
        order.push_back(g);
    std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return gkeys[a] < gkeys[b]; });
    std::vector<int32_t> remap((size_t)n_prov, -1);
    Level C;
    C.n_owned = na;
    C.n_ghost = (int)order.size();
    for (size_t t = 0; t < order.size(); ++t)
    {
      remap[order[t]] = (int32_t)t;
      C.ghost_src.push_back(gkeys[order[t]].first);
      C.ghost_rid.push_back(gkeys[order[t]].second);
    }
    auto final_col = [&](int32_t c) -> int32_t { return c < na ? c : (remap[c - na] < 0 ? -1 : na + remap[c - na]); };
    {
      Csr& Ac = C.A;
      Ac.n_rows = na;
      Ac.n_cols = na + C.n_ghost;
      Ac.ptr.assign((size_t)na + 1, 0);
      std::vector<std::pair<int32_t, double>> row;
      for (int I = 0; I < na; ++I)
      {
        row.clear();
        for (auto& e : crow[I])
          row.emplace_back(final_col(e.first), e.second);
        std::sort(row.begin(), row.end(), [](const std::pair<int32_t, double>& x, const std::pair<int32_t, double>& y)
                  { return x.first < y.first; });
        for (auto& e : row)
        {
          Ac.cols.push_back(e.first);
          Ac.vals.push_back(e.second);
        }
        Ac.ptr[I + 1] = (int32_t)Ac.cols.size();
      }
    }
    // P and the restriction R = (P over owned + ghost rows, my aggregates only)^T in the final numbering
    {
      std::vector<std::pair<int32_t, double>> row;
      for (int i = 0; i < n; ++i)
      {
        row.clear();
        for (int32_t j = P.ptr[i]; j < P.ptr[i + 1]; ++j)
          row.emplace_back(final_col(P.cols[j]), P.vals[j]);
        std::sort(row.begin(), row.end(), [](const std::pair<int32_t, double>& x, const std::pair<int32_t, double>& y)
                  { return x.first < y.first; });
        for (int32_t j = P.ptr[i], t = 0; j < P.ptr[i + 1]; ++j, ++t)
          P.cols[j] = row[t].first, P.vals[j] = row[t].second;
      }
      P.n_cols = na + C.n_ghost;
      Csr Pmine; // rows: owned + ghost fine dofs, columns: my aggregates
      Pmine.n_rows = n + L.n_ghost;
      Pmine.n_cols = na;
      Pmine.ptr.assign((size_t)Pmine.n_rows + 1, 0);
      for (int i = 0; i < Pmine.n_rows; ++i)
      {
        for (int32_t j = Pext.ptr[i]; j < Pext.ptr[i + 1]; ++j)
          if (Pext.cols[j] < na)
          {
            Pmine.cols.push_back(Pext.cols[j]);
            Pmine.vals.push_back(Pext.vals[j]);
          }
        Pmine.ptr[i + 1] = (int32_t)Pmine.cols.size();
      }
      L.R = transpose(Pmine);
    }
    // coarse halo plan: my ghosts are grouped by owner already; tell the owners what I need
    {
      Plan& cp = C.plan;
      std::vector<int> req_ranks;
      std::vector<std::vector<char>> req;
      for (int g = 0; g < C.n_ghost; ++g)
      {
        if (req_ranks.empty() || req_ranks.back() != C.ghost_src[g])
        {
          req_ranks.push_back(C.ghost_src[g]);
          req.emplace_back();
        }
        put1<int32_t>(req.back(), C.ghost_rid[g]);
      }
      std::map<int, std::vector<int32_t>> wanted; // requester -> my aggregate ids
      if (cm.nranks > 1)
        for (auto& m : neighbor_exchange(cm, req_ranks, req))
        {
          std::vector<int32_t> ids(m.second.size() / sizeof(int32_t));
          if (!ids.empty())
            std::memcpy(ids.data(), m.second.data(), ids.size() * sizeof(int32_t));
          wanted[m.first] = std::move(ids);
        }
      // symmetric neighbourhoods (the peer-memory halo pairs every send with a receive): a rank I
      // only send to / only receive from gets an empty segment in the other direction
      std::map<int, std::vector<int32_t>> needed; // owner -> ghost slots
      for (int g = 0; g < C.n_ghost; ++g)
        needed[C.ghost_src[g]].push_back(g);
      std::vector<int> nbrs;
      for (auto& kv : wanted)
        nbrs.push_back(kv.first);
      for (auto& kv : needed)
        nbrs.push_back(kv.first);
      std::sort(nbrs.begin(), nbrs.end());
      nbrs.erase(std::unique(nbrs.begin(), nbrs.end()), nbrs.end());
      for (int r : nbrs)
      {
        cp.send_ranks.push_back(r);
        for (int32_t id : wanted[r])
        {
          PMGX_REQUIRE(id >= 0 && id < na, "amg_setup: rank %d asks for aggregate %d of %d", r, id, na);
          cp.send_idx.push_back(id);
        }
        cp.send_offsets.push_back((int)cp.send_idx.size());
        cp.recv_ranks.push_back(r);
        for (int32_t g : needed[r])
          cp.recv_idx.push_back(g);
        cp.recv_offsets.push_back((int)cp.recv_idx.size());
      }
    }
    L.P = std::move(P);
    H.levels.push_back(std::move(L));
    L = std::move(C);
  }
}
} // namespace amg
} // namespace pmgx

struct pmgx_amg_hier
{
  pmgx::amg::Hierarchy H;
};

using pmgx::amg::Level;

extern "C"
{
int pmgx_amg_setup_dist_h(int rank, int nranks, int n_owned, int n_ghost, const int32_t* row_ptr_h,
                          const int32_t* cols_h, const double* values_h, int n_send_nbr, const int* send_ranks_h,
                          const int* send_offsets_h, const int32_t* send_idx_h, int n_recv_nbr,
                          const int* recv_ranks_h, const int* recv_offsets_h, const int32_t* recv_idx_h,
                          pmgx_allgather_fn allgather, void* user, int min_coarse, int max_levels,
                          pmgx_amg_hier** out)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(out && row_ptr_h && n_owned >= 0 && n_ghost >= 0 && max_levels >= 1, "amg_setup: bad arguments");
  PMGX_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "amg_setup: bad rank");
  PMGX_REQUIRE(nranks == 1 || allgather, "amg_setup: an all-gather callback is needed on more than one rank");
  std::unique_ptr<pmgx_amg_hier> H(new pmgx_amg_hier());
  pmgx::amg::Csr A;
  A.ptr.assign(row_ptr_h, row_ptr_h + n_owned + 1);
  const long long nnz = row_ptr_h[n_owned];
  PMGX_REQUIRE(nnz == 0 || (cols_h && values_h), "amg_setup: null matrix arrays");
  A.cols.assign(cols_h, cols_h + nnz);
  A.vals.assign(values_h, values_h + nnz);
  // rows sorted by column (owned block first): the set-up relies on it
  for (int i = 0; i < n_owned; ++i)
    for (int32_t j = A.ptr[i] + 1; j < A.ptr[i + 1]; ++j)
      PMGX_REQUIRE(A.cols[j - 1] < A.cols[j], "amg_setup: row %d is not sorted by column", i);
  pmgx::amg::Plan pl;
  if (n_send_nbr > 0)
  {
    pl.send_ranks.assign(send_ranks_h, send_ranks_h + n_send_nbr);
    pl.send_offsets.assign(send_offsets_h, send_offsets_h + n_send_nbr + 1);
    pl.send_idx.assign(send_idx_h, send_idx_h + send_offsets_h[n_send_nbr]);
  }
  if (n_recv_nbr > 0)
  {
    pl.recv_ranks.assign(recv_ranks_h, recv_ranks_h + n_recv_nbr);
    pl.recv_offsets.assign(recv_offsets_h, recv_offsets_h + n_recv_nbr + 1);
    pl.recv_idx.assign(recv_idx_h, recv_idx_h + recv_offsets_h[n_recv_nbr]);
  }
  pmgx::amg::Comm cm;
  cm.rank = rank;
  cm.nranks = nranks;
  cm.allgather = [allgather, user](const void* mine, size_t bytes, void* all)
  {
    if (allgather(user, mine, bytes, all) != 0)
      throw std::runtime_error("amg set-up: the all-gather callback failed");
  };
  pmgx::amg::setup(H->H, std::move(A), n_owned, n_ghost, pl, cm, min_coarse, max_levels);
  *out = H.release();
  PMGX_API_END
}

int pmgx_amg_setup_h(int n_rows, const int32_t* row_ptr_h, const int32_t* cols_h, const double* values_h,
                     int min_coarse, int max_levels, pmgx_amg_hier** out)
{
  return pmgx_amg_setup_dist_h(0, 1, n_rows, 0, row_ptr_h, cols_h, values_h, 0, nullptr, nullptr, nullptr, 0, nullptr,
                               nullptr, nullptr, nullptr, nullptr, min_coarse, max_levels, out);
}

int pmgx_csr_drop_zeros_h(int n_rows, const int32_t* ptr_h, const int32_t* cols_h, const double* vals_h,
                          int32_t* out_ptr_h, int32_t* out_cols_h, double* out_vals_h, long long* kept_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(n_rows >= 0 && ptr_h && kept_h && (ptr_h[n_rows] == 0 || (cols_h && vals_h)), "csr_drop_zeros: bad arguments");
  pmgx::amg::Csr A;
  A.n_rows = A.n_cols = n_rows;
  A.ptr.assign(ptr_h, ptr_h + n_rows + 1);
  A.cols.assign(cols_h, cols_h + ptr_h[n_rows]);
  A.vals.assign(vals_h, vals_h + ptr_h[n_rows]);
  const pmgx::amg::Csr B = pmgx::amg::drop_stored_zeros(A);
  *kept_h = B.nnz();
  if (out_ptr_h)
    std::copy(B.ptr.begin(), B.ptr.end(), out_ptr_h);
  if (out_cols_h)
    std::copy(B.cols.begin(), B.cols.end(), out_cols_h);
  if (out_vals_h)
    std::copy(B.vals.begin(), B.vals.end(), out_vals_h);
  PMGX_API_END
}

int pmgx_amg_num_levels(pmgx_amg_hier* h) { return h ? (int)h->H.levels.size() : -1; }

int pmgx_amg_level_sizes(pmgx_amg_hier* h, int level, long long* out_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(h && out_h && level >= 0 && level < (int)h->H.levels.size(), "amg_level_sizes: bad arguments");
  const Level& L = h->H.levels[level];
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
  PMGX_REQUIRE(h && level >= 0 && level < (int)h->H.levels.size(), "amg_level_get: bad arguments");
  const Level& L = h->H.levels[level];
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

int pmgx_amg_level_dist_sizes(pmgx_amg_hier* h, int level, long long* out_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(h && out_h && level >= 0 && level < (int)h->H.levels.size(), "amg_level_dist_sizes: bad arguments");
  const Level& L = h->H.levels[level];
  out_h[0] = L.n_owned;
  out_h[1] = L.n_ghost;
  out_h[2] = (long long)L.plan.send_ranks.size();
  out_h[3] = (long long)L.plan.send_idx.size();
  out_h[4] = (long long)L.plan.recv_ranks.size();
  out_h[5] = (long long)L.plan.recv_idx.size();
  out_h[6] = L.dense ? 1 : 0;
  out_h[7] = L.n_global;
  out_h[8] = L.R.nnz();
  out_h[9] = L.replicated ? 1 : 0;
  out_h[10] = L.n_mine;
  PMGX_API_END
}

int pmgx_amg_level_dist_get(pmgx_amg_hier* h, int level, int* ghost_src_h, int32_t* ghost_rid_h, int* send_ranks_h,
                            int* send_offsets_h, int32_t* send_idx_h, int* recv_ranks_h, int* recv_offsets_h,
                            int32_t* recv_idx_h, double* inv_rows_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(h && level >= 0 && level < (int)h->H.levels.size(), "amg_level_dist_get: bad arguments");
  const Level& L = h->H.levels[level];
  if (ghost_src_h)
    std::copy(L.ghost_src.begin(), L.ghost_src.end(), ghost_src_h);
  if (ghost_rid_h)
    std::copy(L.ghost_rid.begin(), L.ghost_rid.end(), ghost_rid_h);
  if (send_ranks_h)
    std::copy(L.plan.send_ranks.begin(), L.plan.send_ranks.end(), send_ranks_h);
  if (send_offsets_h)
    std::copy(L.plan.send_offsets.begin(), L.plan.send_offsets.end(), send_offsets_h);
  if (send_idx_h)
    std::copy(L.plan.send_idx.begin(), L.plan.send_idx.end(), send_idx_h);
  if (recv_ranks_h)
    std::copy(L.plan.recv_ranks.begin(), L.plan.recv_ranks.end(), recv_ranks_h);
  if (recv_offsets_h)
    std::copy(L.plan.recv_offsets.begin(), L.plan.recv_offsets.end(), recv_offsets_h);
  if (recv_idx_h)
    std::copy(L.plan.recv_idx.begin(), L.plan.recv_idx.end(), recv_idx_h);
  if (inv_rows_h)
    std::copy(L.inv_rows.begin(), L.inv_rows.end(), inv_rows_h);
  PMGX_API_END
}

int pmgx_amg_level_get_restriction(pmgx_amg_hier* h, int level, int32_t* r_ptr_h, int32_t* r_cols_h, double* r_vals_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(h && level >= 0 && level < (int)h->H.levels.size(), "amg_level_get_restriction: bad arguments");
  const Level& L = h->H.levels[level];
  if (r_ptr_h && !L.R.ptr.empty())
    std::copy(L.R.ptr.begin(), L.R.ptr.end(), r_ptr_h);
  if (r_cols_h)
    std::copy(L.R.cols.begin(), L.R.cols.end(), r_cols_h);
  if (r_vals_h)
    std::copy(L.R.vals.begin(), L.R.vals.end(), r_vals_h);
  PMGX_API_END
}

int pmgx_amg_destroy(pmgx_amg_hier* h)
{
  PMGX_API_BEGIN
  delete h;
  PMGX_API_END
}
}
