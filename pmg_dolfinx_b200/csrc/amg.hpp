// Smoothed-aggregation AMG hierarchy shared between the host set-up (amg_setup.cpp) and the device
// cycle (amg.cu).  Stands where the reference runs PETSc CG + BoomerAMG (src/amg.hpp:33-47).
//
// Distribution model (one rank per GPU, rows = owned dofs, columns = owned + ghost, exactly the
// layout of acc::MatrixOperator, src/csr.hpp:57-131):
//   * aggregates never span ranks, but the prolongator is smoothed with the full rows of A, so a row
//     next to a partition interface also interpolates from the neighbour's aggregates (P has ghost
//     columns; its transpose is stored as R over owned + ghost fine columns): prolongation needs a
//     forward halo update of the coarse vector, restriction one of the fine residual;
//   * the Galerkin product A_c = P^T A P needs the prolongator rows of the ghost dofs (one exchange of
//     sparse rows at set-up) and returns partial rows of the neighbours' aggregates to their owners
//     (a second one); A_c has the same owned-rows / ghost-columns layout with a halo plan of its own
//     (covering the ghost aggregates A_c or P reference), so the construction recurses;
//   * the coarsest level is gathered and inverted densely on every rank (each keeps its own rows).
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <vector>

namespace pmgx
{
namespace amg
{
struct Csr
{
  int n_rows = 0, n_cols = 0;
  std::vector<int32_t> ptr, cols;
  std::vector<double> vals;
  long long nnz() const { return (long long)cols.size(); }
};

// forward-scatter plan in the format of pmgx_halo_create (Scatterer lists, src/vector.hpp:89-95)
struct Plan
{
  std::vector<int> send_ranks, send_offsets{0}, recv_ranks, recv_offsets{0};
  std::vector<int32_t> send_idx; // owned local indices grouped by destination
  std::vector<int32_t> recv_idx; // ghost slots (0-based in the ghost block) grouped by source
};

// the only collective the set-up needs: fixed-size all-gather of bytes (NCCL in production, a
// torch.distributed/gloo callback in the CPU tests)
struct Comm
{
  int rank = 0, nranks = 1;
  std::function<void(const void* mine, size_t bytes, void* all)> allgather;
};

struct Level
{
  int n_owned = 0, n_ghost = 0;
  Csr A;    // n_owned rows; columns < n_owned are owned, the rest address ghosts; rows sorted by column
  Plan plan;
  Csr P;    // n_owned x (owned + ghost dofs of the next level); empty on the coarsest level
  Csr R;    // (owned dofs of the next level) x (n_owned + n_ghost): the rows of the global P^T this rank owns
  double lmax = 1.0; // 1.1 * lambda_max(D^-1 A), power iteration
  std::vector<int> ghost_src;      // per ghost: owning rank
  std::vector<int32_t> ghost_rid;  // per ghost: index on the owning rank
  // replicated levels (small on the whole machine): the complete level lives on every rank, n_owned = its global
  // size, no ghosts, no halo.  The FIRST replicated level is entered through an all-gather of the restricted
  // residual: n_mine entries of it come from this rank's restriction, repl_plan brings the others,
  // gather_perm[g] = position of canonical index g in the gather layout [mine | others in rank order].  The
  // parent level's P addresses the canonical numbering directly.
  bool replicated = false;
  int n_mine = 0;
  std::vector<int32_t> gather_perm;
  Plan repl_plan;
  // coarsest level only: rows of the dense inverse of the gathered matrix, columns in this rank's
  // vector layout [owned | all other ranks' entries in rank order] (= gather_plan's ghost block)
  bool dense = false;
  long long n_global = 0;
  std::vector<double> inv_rows; // n_owned x n_global
  Plan gather_plan;
};

struct Hierarchy
{
  std::vector<Level> levels;
};

// A without its stored zeros (the diagonal entry of a row is kept whatever its value).  The device assembly keeps
// the full pattern of the element matrices (27 entries per row at P1) of which 7 (affine boxes) to 19 are non-zero:
// the set-up wants that pattern (its aggregates follow it), the SpMVs of the solve do not.
Csr drop_stored_zeros(const Csr& A);

// collective; A0's columns >= n_owned address the ghosts of plan0
void setup(Hierarchy& H, Csr A0, int n_owned, int n_ghost, const Plan& plan0, const Comm& comm, int min_coarse,
           int max_levels);
} // namespace amg
} // namespace pmgx
