// Host-side structured box mesh with a one-cell ghost layer, tensor-product GLL dofmaps,
// Dirichlet markers and forward-scatter lists.  Pure host code (no CUDA calls).
//
// Harness-side stand-in for what the reference drivers obtain from DOLFINx:
//   mesh::create_box + ghost_layer_mesh      examples/pmg/main.cpp:442-451, src/mesh.hpp:16-98
//   compute_boundary_cells (lcells/bcells)   src/mesh.hpp:105-143
//   tp dofmaps per degree, IndexMap sizes    examples/pmg/main.cpp:83-87,199-213
//   exterior-facet Dirichlet marker          examples/pmg/main.cpp:122-124,173-185
//   Scatterer local/remote index lists       src/vector.hpp:89-95
//
// Partition: nx*ny*nz cells split into px*py*pz blocks (cell c belongs to block
// b iff split[b] <= c < split[b+1], split[i] = i*n/p); rank = (bx*py+by)*pz+bz.
// Every cell sharing a vertex with an owned cell is a ghost cell (src/mesh.hpp:25-46).
// Dof ownership (documented choice): the lowest rank owning a cell that contains the dof;
// for this block partition that is the block of the lowest cell containing the dof in each
// direction.  Local numbering: owned dofs lexicographic, then ghosts sorted by (owner, global id).
#include "common.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>

namespace
{
inline uint64_t splitmix64(uint64_t z)
{
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
inline double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

struct Axis
{
  int n = 0, p = 0, b = 0; // cells, blocks, my block
  int lo = 0, hi = 0;      // owned cells [lo, hi)
  int elo = 0, ehi = 0;    // local (owned + ghost) cells [elo, ehi)
  int split(int i) const { return (int)(((long long)i * n) / p); }
  int block_of_cell(int c) const
  {
    int bb = (int)(((long long)(c + 1) * p - 1) / n); // largest b with split(b) <= c
    while (split(bb) > c)
      --bb;
    while (bb + 1 <= p - 1 && split(bb + 1) <= c)
      ++bb;
    return bb;
  }
  void init(int n_, int p_, int b_)
  {
    n = n_, p = p_, b = b_;
    lo = split(b), hi = split(b + 1);
    elo = std::max(0, lo - 1), ehi = std::min(n, hi + 1);
  }
  // dof-grid helpers for degree P
  int lowcell(int g, int P) const { return g == 0 ? 0 : (g - 1) / P; }
  int owner_block(int g, int P) const { return block_of_cell(lowcell(g, P)); }
  int own_lo(int P) const { return lo == 0 ? 0 : P * lo + 1; } // owned dof range [own_lo, own_hi]
  int own_hi(int P) const { return P * hi; }
  int ext_lo(int P) const { return P * elo; }
  int ext_hi(int P) const { return P * ehi; }
};

struct Space
{
  int P = 0;
  long long n_owned = 0, n_ghost = 0, n_global = 0;
  std::vector<int32_t> ext_to_local;       // dense over the extended dof box
  std::vector<long long> ghost_gid;        // global id of each ghost, local order
  std::vector<int> ghost_owner;
  std::vector<int> send_ranks, send_offsets, recv_ranks, recv_offsets;
  std::vector<int32_t> send_idx, recv_idx;
};
} // namespace

struct pmgx_boxmesh
{
  Axis ax[3];
  int rank = 0;
  double perturb = 0.0;
  uint64_t seed = 0;
  long long n_cells = 0, n_owned_cells = 0, n_points = 0;
  std::vector<std::array<int, 3>> cells; // global (cx,cy,cz), owned first then ghost
  std::vector<int32_t> lcells, bcells;
  std::map<int, std::unique_ptr<Space>> spaces;

  int rank_of_block(int bx, int by, int bz) const { return (bx * ax[1].p + by) * ax[2].p + bz; }

  void vertex_coord(int vx, int vy, int vz, double* out) const
  {
    const int v[3] = {vx, vy, vz};
    bool interior = true;
    for (int d = 0; d < 3; ++d)
      interior = interior && v[d] > 0 && v[d] < ax[d].n;
    const long long gid = ((long long)vx * (ax[1].n + 1) + vy) * (ax[2].n + 1) + vz;
    for (int d = 0; d < 3; ++d)
    {
      const double h = 1.0 / ax[d].n;
      double x = (double)v[d] / (double)ax[d].n;
      if (interior && perturb > 0.0)
        x += (2.0 * u01(splitmix64(seed ^ (uint64_t)(gid * 3 + d))) - 1.0) * perturb * h;
      out[d] = x;
    }
  }

  Space& space(int P)
  {
    auto it = spaces.find(P);
    if (it != spaces.end())
      return *it->second;
    auto sp = std::make_unique<Space>();
    Space& s = *sp;
    s.P = P;
    const long long N[3] = {(long long)P * ax[0].n + 1, (long long)P * ax[1].n + 1,
                            (long long)P * ax[2].n + 1};
    s.n_global = N[0] * N[1] * N[2];
    int ol[3], oh[3], el[3], eh[3];
    for (int d = 0; d < 3; ++d)
      ol[d] = ax[d].own_lo(P), oh[d] = ax[d].own_hi(P), el[d] = ax[d].ext_lo(P), eh[d] = ax[d].ext_hi(P);
    const long long O[3] = {oh[0] - ol[0] + 1, oh[1] - ol[1] + 1, oh[2] - ol[2] + 1};
    const long long E[3] = {eh[0] - el[0] + 1, eh[1] - el[1] + 1, eh[2] - el[2] + 1};
    s.n_owned = O[0] * O[1] * O[2];
    s.ext_to_local.assign((size_t)(E[0] * E[1] * E[2]), -1);
    // per-axis owner block of every extended dof coordinate
    std::vector<int> ob[3];
    for (int d = 0; d < 3; ++d)
    {
      ob[d].resize(E[d]);
      for (int g = el[d]; g <= eh[d]; ++g)
        ob[d][g - el[d]] = ax[d].owner_block(g, P);
    }
    struct GhostRec
    {
      int owner;
      long long gid;
      long long ext;
    };
    std::vector<GhostRec> ghosts;
    for (long long ex = 0; ex < E[0]; ++ex)
      for (long long ey = 0; ey < E[1]; ++ey)
        for (long long ez = 0; ez < E[2]; ++ez)
        {
          const int gx = el[0] + (int)ex, gy = el[1] + (int)ey, gz = el[2] + (int)ez;
          const long long ext = (ex * E[1] + ey) * E[2] + ez;
          const bool owned = ob[0][ex] == ax[0].b && ob[1][ey] == ax[1].b && ob[2][ez] == ax[2].b;
          if (owned)
            s.ext_to_local[ext]
                = (int32_t)((((long long)(gx - ol[0])) * O[1] + (gy - ol[1])) * O[2] + (gz - ol[2]));
          else
            ghosts.push_back({rank_of_block(ob[0][ex], ob[1][ey], ob[2][ez]),
                              ((long long)gx * N[1] + gy) * N[2] + gz, ext});
        }
    std::sort(ghosts.begin(), ghosts.end(), [](const GhostRec& a, const GhostRec& b)
              { return a.owner != b.owner ? a.owner < b.owner : a.gid < b.gid; });
    s.n_ghost = (long long)ghosts.size();
    s.ghost_gid.resize(ghosts.size());
    s.ghost_owner.resize(ghosts.size());
    s.recv_offsets.assign(1, 0);
    for (size_t i = 0; i < ghosts.size(); ++i)
    {
      s.ext_to_local[ghosts[i].ext] = (int32_t)(s.n_owned + (long long)i);
      s.ghost_gid[i] = ghosts[i].gid;
      s.ghost_owner[i] = ghosts[i].owner;
      if (s.recv_ranks.empty() || s.recv_ranks.back() != ghosts[i].owner)
      {
        if (!s.recv_ranks.empty())
          s.recv_offsets.push_back((int)i);
        s.recv_ranks.push_back(ghosts[i].owner);
      }
      s.recv_idx.push_back((int32_t)i);
    }
    if (!s.recv_ranks.empty())
      s.recv_offsets.push_back((int)ghosts.size());
    // send lists: for every nearby block q, my owned dofs inside q's extended dof box,
    // in global-id order (what q's ghost ordering expects)
    s.send_offsets.assign(1, 0);
    // (blocks one cell wide make the block two steps up a neighbour too: its ghost layer
    // reaches the dofs on my upper face)
    for (int dbx = -2; dbx <= 2; ++dbx)
      for (int dby = -2; dby <= 2; ++dby)
        for (int dbz = -2; dbz <= 2; ++dbz)
        {
          if (dbx == 0 && dby == 0 && dbz == 0)
            continue;
          const int qb[3] = {ax[0].b + dbx, ax[1].b + dby, ax[2].b + dbz};
          bool ok = true;
          for (int d = 0; d < 3; ++d)
            ok = ok && qb[d] >= 0 && qb[d] < ax[d].p;
          if (!ok)
            continue;
          int rl[3], rh[3];
          bool nonempty = true;
          for (int d = 0; d < 3; ++d)
          {
            Axis q;
            q.init(ax[d].n, ax[d].p, qb[d]);
            rl[d] = std::max(ol[d], q.ext_lo(P));
            rh[d] = std::min(oh[d], q.ext_hi(P));
            nonempty = nonempty && rl[d] <= rh[d];
          }
          if (!nonempty)
            continue;
          const size_t before = s.send_idx.size();
          for (int gx = rl[0]; gx <= rh[0]; ++gx)
            for (int gy = rl[1]; gy <= rh[1]; ++gy)
              for (int gz = rl[2]; gz <= rh[2]; ++gz)
                s.send_idx.push_back(
                    (int32_t)((((long long)(gx - ol[0])) * O[1] + (gy - ol[1])) * O[2] + (gz - ol[2])));
          if (s.send_idx.size() > before)
          {
            s.send_ranks.push_back(rank_of_block(qb[0], qb[1], qb[2]));
            s.send_offsets.push_back((int)s.send_idx.size());
          }
        }
    // order send neighbours by rank (cosmetic; ranks iterate in increasing order already)
    spaces[P] = std::move(sp);
    return *spaces[P];
  }
};

extern "C"
{
int pmgx_boxmesh_create(int nx, int ny, int nz, int px, int py, int pz, int rank, double perturb,
                        uint64_t seed, pmgx_boxmesh** out)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(out, "boxmesh_create: null output");
  PMGX_REQUIRE(nx >= 1 && ny >= 1 && nz >= 1 && px >= 1 && py >= 1 && pz >= 1, "boxmesh_create: bad sizes");
  PMGX_REQUIRE(px <= nx && py <= ny && pz <= nz, "boxmesh_create: more blocks than cells");
  PMGX_REQUIRE(rank >= 0 && rank < px * py * pz, "boxmesh_create: bad rank");
  PMGX_REQUIRE(perturb >= 0.0 && perturb < 0.5, "boxmesh_create: perturb must be in [0, 0.5)");
  auto m = std::make_unique<pmgx_boxmesh>();
  m->rank = rank;
  m->perturb = perturb;
  m->seed = seed;
  const int bz = rank % pz, by = (rank / pz) % py, bx = rank / (pz * py);
  m->ax[0].init(nx, px, bx);
  m->ax[1].init(ny, py, by);
  m->ax[2].init(nz, pz, bz);
  const Axis* a = m->ax;
  // owned cells first (lexicographic), then ghost cells (lexicographic in the extended block)
  for (int cx = a[0].lo; cx < a[0].hi; ++cx)
    for (int cy = a[1].lo; cy < a[1].hi; ++cy)
      for (int cz = a[2].lo; cz < a[2].hi; ++cz)
        m->cells.push_back({cx, cy, cz});
  m->n_owned_cells = (long long)m->cells.size();
  for (int cx = a[0].elo; cx < a[0].ehi; ++cx)
    for (int cy = a[1].elo; cy < a[1].ehi; ++cy)
      for (int cz = a[2].elo; cz < a[2].ehi; ++cz)
      {
        const bool owned = cx >= a[0].lo && cx < a[0].hi && cy >= a[1].lo && cy < a[1].hi
                           && cz >= a[2].lo && cz < a[2].hi;
        if (!owned)
          m->cells.push_back({cx, cy, cz});
      }
  m->n_cells = (long long)m->cells.size();
  m->n_points = (long long)(a[0].ehi - a[0].elo + 1) * (a[1].ehi - a[1].elo + 1) * (a[2].ehi - a[2].elo + 1);
  // lcells: owned cells whose dofs are all owned; bcells: the rest + all ghost cells
  for (long long c = 0; c < m->n_cells; ++c)
  {
    bool boundary = c >= m->n_owned_cells;
    if (!boundary)
      for (int d = 0; d < 3; ++d)
        boundary = boundary || (a[d].lo > 0 && m->cells[c][d] == a[d].lo);
    (boundary ? m->bcells : m->lcells).push_back((int32_t)c);
  }
  *out = m.release();
  PMGX_API_END
}

int pmgx_boxmesh_destroy(pmgx_boxmesh* m)
{
  delete m;
  return PMGX_OK;
}

int pmgx_boxmesh_sizes(pmgx_boxmesh* m, long long* out_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(m && out_h, "boxmesh_sizes: null argument");
  out_h[0] = m->n_cells;
  out_h[1] = m->n_owned_cells;
  out_h[2] = m->n_points;
  out_h[3] = (long long)m->lcells.size();
  out_h[4] = (long long)m->bcells.size();
  PMGX_API_END
}

int pmgx_boxmesh_geometry(pmgx_boxmesh* m, double* xgeom_h, int32_t* geom_dofmap_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(m, "boxmesh_geometry: null mesh");
  const Axis* a = m->ax;
  const long long VY = a[1].ehi - a[1].elo + 1, VZ = a[2].ehi - a[2].elo + 1;
  if (xgeom_h)
  {
    for (int vx = a[0].elo; vx <= a[0].ehi; ++vx)
      for (int vy = a[1].elo; vy <= a[1].ehi; ++vy)
        for (int vz = a[2].elo; vz <= a[2].ehi; ++vz)
        {
          const long long lv = ((long long)(vx - a[0].elo) * VY + (vy - a[1].elo)) * VZ + (vz - a[2].elo);
          m->vertex_coord(vx, vy, vz, xgeom_h + 3 * lv);
        }
  }
  if (geom_dofmap_h)
  {
    for (long long c = 0; c < m->n_cells; ++c)
      for (int k = 0; k < 8; ++k)
      {
        const int vx = m->cells[c][0] + ((k >> 2) & 1), vy = m->cells[c][1] + ((k >> 1) & 1),
                  vz = m->cells[c][2] + (k & 1);
        geom_dofmap_h[c * 8 + k]
            = (int32_t)(((long long)(vx - a[0].elo) * VY + (vy - a[1].elo)) * VZ + (vz - a[2].elo));
      }
  }
  PMGX_API_END
}

int pmgx_boxmesh_cell_lists(pmgx_boxmesh* m, int32_t* lcells_h, int32_t* bcells_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(m, "boxmesh_cell_lists: null mesh");
  if (lcells_h && !m->lcells.empty())
    std::memcpy(lcells_h, m->lcells.data(), m->lcells.size() * sizeof(int32_t));
  if (bcells_h && !m->bcells.empty())
    std::memcpy(bcells_h, m->bcells.data(), m->bcells.size() * sizeof(int32_t));
  PMGX_API_END
}

int pmgx_boxmesh_space_sizes(pmgx_boxmesh* m, int degree, long long* out_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(m && out_h, "boxmesh_space_sizes: null argument");
  PMGX_REQUIRE(degree >= 1 && degree <= PMGX_MAX_DEGREE, "Unsupported degree %d", degree);
  Space& s = m->space(degree);
  out_h[0] = s.n_owned;
  out_h[1] = s.n_ghost;
  out_h[2] = (long long)s.send_ranks.size();
  out_h[3] = (long long)s.send_idx.size();
  out_h[4] = (long long)s.recv_ranks.size();
  out_h[5] = (long long)s.recv_idx.size();
  out_h[6] = s.n_global;
  PMGX_API_END
}

int pmgx_boxmesh_space(pmgx_boxmesh* m, int degree, int32_t* dofmap_h, int8_t* bc_h,
                       long long* l2g_h, double* coords_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(m, "boxmesh_space: null mesh");
  PMGX_REQUIRE(degree >= 1 && degree <= PMGX_MAX_DEGREE, "Unsupported degree %d", degree);
  const int P = degree, n = P + 1, n3 = n * n * n;
  Space& s = m->space(P);
  const Axis* a = m->ax;
  const long long N[3] = {(long long)P * a[0].n + 1, (long long)P * a[1].n + 1, (long long)P * a[2].n + 1};
  const int el[3] = {a[0].ext_lo(P), a[1].ext_lo(P), a[2].ext_lo(P)};
  const long long E1 = a[1].ext_hi(P) - el[1] + 1, E2 = a[2].ext_hi(P) - el[2] + 1;
  const long long E0 = a[0].ext_hi(P) - el[0] + 1;
  if (dofmap_h)
  {
#pragma omp parallel for schedule(static)
    for (long long c = 0; c < m->n_cells; ++c)
    {
      const int bx = m->cells[c][0] * P - el[0], by = m->cells[c][1] * P - el[1],
                bz = m->cells[c][2] * P - el[2];
      int32_t* row = dofmap_h + c * n3;
      for (int ix = 0; ix < n; ++ix)
        for (int iy = 0; iy < n; ++iy)
          for (int iz = 0; iz < n; ++iz)
            row[(ix * n + iy) * n + iz] = s.ext_to_local[((long long)(bx + ix) * E1 + (by + iy)) * E2 + (bz + iz)];
    }
  }
  if (bc_h || l2g_h)
  {
#pragma omp parallel for schedule(static)
    for (long long ex = 0; ex < E0; ++ex)
      for (long long ey = 0; ey < E1; ++ey)
        for (long long ez = 0; ez < E2; ++ez)
        {
          const long long gx = el[0] + ex, gy = el[1] + ey, gz = el[2] + ez;
          const int32_t l = s.ext_to_local[(ex * E1 + ey) * E2 + ez];
          if (bc_h)
            bc_h[l] = (gx == 0 || gx == N[0] - 1 || gy == 0 || gy == N[1] - 1 || gz == 0 || gz == N[2] - 1) ? 1 : 0;
          if (l2g_h)
            l2g_h[l] = (gx * N[1] + gy) * N[2] + gz;
        }
  }
  if (coords_h)
  {
    std::vector<double> xs, ws;
    pmgx::gll_points_weights(n, xs, ws);
#pragma omp parallel for schedule(static)
    for (long long c = 0; c < m->n_cells; ++c)
    {
      double v[8][3];
      for (int k = 0; k < 8; ++k)
        m->vertex_coord(m->cells[c][0] + ((k >> 2) & 1), m->cells[c][1] + ((k >> 1) & 1),
                        m->cells[c][2] + (k & 1), v[k]);
      const int bx = m->cells[c][0] * P - el[0], by = m->cells[c][1] * P - el[1],
                bz = m->cells[c][2] * P - el[2];
      for (int ix = 0; ix < n; ++ix)
        for (int iy = 0; iy < n; ++iy)
          for (int iz = 0; iz < n; ++iz)
          {
            const double xi[3] = {xs[ix], xs[iy], xs[iz]};
            double X[3] = {0, 0, 0};
            for (int k = 0; k < 8; ++k)
            {
              const double phi = (((k >> 2) & 1) ? xi[0] : 1 - xi[0]) * (((k >> 1) & 1) ? xi[1] : 1 - xi[1])
                                 * ((k & 1) ? xi[2] : 1 - xi[2]);
              for (int d = 0; d < 3; ++d)
                X[d] += phi * v[k][d];
            }
            const int32_t l = s.ext_to_local[((long long)(bx + ix) * E1 + (by + iy)) * E2 + (bz + iz)];
            // shared dofs get the same value (to rounding) from every cell; benign overwrite
            for (int d = 0; d < 3; ++d)
              coords_h[3ll * l + d] = X[d];
          }
    }
  }
  PMGX_API_END
}

int pmgx_boxmesh_halo_lists(pmgx_boxmesh* m, int degree, int* send_ranks_h, int* send_offsets_h,
                            int32_t* send_idx_h, int* recv_ranks_h, int* recv_offsets_h,
                            int32_t* recv_idx_h)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(m, "boxmesh_halo_lists: null mesh");
  PMGX_REQUIRE(degree >= 1 && degree <= PMGX_MAX_DEGREE, "Unsupported degree %d", degree);
  Space& s = m->space(degree);
  auto cp = [](auto* dst, const auto& v)
  {
    if (dst && !v.empty())
      std::memcpy(dst, v.data(), v.size() * sizeof(v[0]));
  };
  cp(send_ranks_h, s.send_ranks);
  cp(send_offsets_h, s.send_offsets);
  cp(send_idx_h, s.send_idx);
  cp(recv_ranks_h, s.recv_ranks);
  cp(recv_offsets_h, s.recv_offsets);
  cp(recv_idx_h, s.recv_idx);
  PMGX_API_END
}
}
