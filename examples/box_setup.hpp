// Box-mesh set-up shared by the example drivers: the part of the reference drivers that is
// DOLFINx (mesh::create_box, ghost_layer_mesh, FunctionSpace, dofmaps, locate_dofs_topological;
// examples/pmg/main.cpp:63-124,199-256) replaced by the pmgx_boxmesh_* helpers of the C ABI.
#pragma once
#include <pmgx/dolfinx_acc_compat.hpp>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>

namespace box
{
struct Space
{
  int degree = 0;
  long long n_global = 0;
  std::shared_ptr<pmgx::IndexMap> map;
  pmgx::DeviceArray<std::int32_t> dofmap;
  pmgx::DeviceArray<std::int8_t> bc;
  std::vector<double> coords; // owned + ghost, xyz
};

inline int env_int(const char* name, int dflt)
{
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

// The reference broadcasts nothing (MPI is its transport); here rank 0 creates the NCCL id and
// hands it to the other ranks through a file.
inline void exchange_nccl_id(int rank, int nranks, const std::string& idfile, std::vector<char>& id)
{
  id.assign(PMGX_NCCL_ID_BYTES, 0);
  if (nranks == 1)
    return;
  if (idfile.empty())
    throw std::runtime_error("multi-rank run needs --idfile");
  if (rank == 0)
  {
    pmgx::check(pmgx_nccl_unique_id(id.data()));
    std::ofstream f(idfile + ".tmp", std::ios::binary);
    f.write(id.data(), id.size());
    f.close();
    std::rename((idfile + ".tmp").c_str(), idfile.c_str());
    return;
  }
  for (int i = 0; i < 3000; ++i)
  {
    std::ifstream f(idfile, std::ios::binary);
    if (f && f.read(id.data(), id.size()))
      return;
    std::this_thread::sleep_for(std::chrono::milliseconds(20));
  }
  throw std::runtime_error("timed out waiting for the NCCL id file");
}

/// Mesh of one rank: geometry, kappa and the interior / boundary cell lists on host + device.
struct Mesh
{
  pmgx_boxmesh* h = nullptr;
  int nxyz[3] = {0, 0, 0};
  int n_cells = 0, n_points = 0;
  std::vector<int> lcells, bcells;
  pmgx::DeviceArray<double> xgeom, kappa;
  pmgx::DeviceArray<std::int32_t> geometry_dofmap;

  Mesh(long long ndofs_per_rank, int order, int rank, int nranks, double perturb, double kappa_value)
  {
    pmgx::check(pmgx_boxmesh_fit(ndofs_per_rank * nranks, order, nxyz)); // examples/pmg/main.cpp:412-435
    const int pg[4][3] = {{1, 1, 1}, {2, 1, 1}, {2, 2, 1}, {2, 2, 2}};
    const int* p = pg[nranks == 1 ? 0 : (nranks == 2 ? 1 : (nranks == 4 ? 2 : 3))];
    if (p[0] * p[1] * p[2] != nranks)
      throw std::runtime_error("WORLD_SIZE must be 1, 2, 4 or 8");
    pmgx::check(pmgx_boxmesh_create(nxyz[0], nxyz[1], nxyz[2], p[0], p[1], p[2], rank, perturb, 1234, &h));
    long long ms[5];
    pmgx::check(pmgx_boxmesh_sizes(h, ms));
    n_cells = (int)ms[0], n_points = (int)ms[2];
    std::vector<double> xgeom_h((size_t)n_points * 3);
    std::vector<std::int32_t> gdm_h((size_t)n_cells * 8);
    pmgx::check(pmgx_boxmesh_geometry(h, xgeom_h.data(), gdm_h.data()));
    lcells.resize((size_t)ms[3]);
    bcells.resize((size_t)ms[4]);
    pmgx::check(pmgx_boxmesh_cell_lists(h, lcells.data(), bcells.data()));
    xgeom.assign(xgeom_h.data(), xgeom_h.size());
    geometry_dofmap.assign(gdm_h.data(), gdm_h.size());
    std::vector<double> k((size_t)n_cells, kappa_value);
    kappa.assign(k.data(), k.size());
  }
  ~Mesh() { pmgx_boxmesh_destroy(h); }
  Mesh(const Mesh&) = delete;

  /// Function space of one degree: IndexMap (with forward-scatter lists), device dofmap, BC marker.
  void make_space(std::shared_ptr<const pmgx::Context> ctx, int degree, Space& s) const
  {
    s.degree = degree;
    long long ss[7];
    pmgx::check(pmgx_boxmesh_space_sizes(h, degree, ss));
    const int n_owned = (int)ss[0], n_ghost = (int)ss[1];
    s.n_global = ss[6];
    const int nd3 = (degree + 1) * (degree + 1) * (degree + 1);
    std::vector<std::int32_t> dm((size_t)n_cells * nd3);
    std::vector<std::int8_t> bc((size_t)n_owned + n_ghost);
    s.coords.resize(((size_t)n_owned + n_ghost) * 3);
    pmgx::check(pmgx_boxmesh_space(h, degree, dm.data(), bc.data(), nullptr, s.coords.data()));
    std::vector<int> sr((size_t)ss[2]), so((size_t)ss[2] + 1), rr((size_t)ss[4]), ro((size_t)ss[4] + 1);
    std::vector<std::int32_t> si((size_t)ss[3]), ri((size_t)ss[5]);
    pmgx::check(pmgx_boxmesh_halo_lists(h, degree, sr.data(), so.data(), si.data(), rr.data(), ro.data(), ri.data()));
    s.map = std::make_shared<pmgx::IndexMap>(ctx, n_owned, n_ghost, sr, so, si, rr, ro, ri);
    s.dofmap.assign(dm.data(), dm.size());
    s.bc.assign(bc.data(), bc.size());
  }
};
} // namespace box
