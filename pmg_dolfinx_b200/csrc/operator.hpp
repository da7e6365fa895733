// Internal operator / halo types behind the opaque C handles.
#pragma once
#include "common.hpp"

// Forward-scatter plan (role of dolfinx::common::Scatterer + the pack/unpack buffers of
// acc::Vector, src/vector.hpp:83-95,186-238).
struct pmgx_halo
{
  pmgx_ctx* ctx = nullptr;
  int n_owned = 0, n_ghost = 0;
  std::vector<int> send_ranks, send_offsets, recv_ranks, recv_offsets;
  pmgx::DevBuf<int32_t> send_idx, recv_idx;
  pmgx::DevBuf<double> send_buf, recv_buf;
  cudaEvent_t ev_ready = nullptr; // compute stream -> comm stream (x is ready to pack)
  cudaEvent_t ev_done = nullptr;  // comm stream -> compute stream (ghosts are in place)
  bool in_flight = false;
  // NVLink peer-memory path (halo.cu, p2p.cu): packed values are stored straight into the
  // destination rank's receive buffer; an epoch flag per source rank signals completion
  bool p2p = false;
  double* xbuf = nullptr;                        // IPC-shared: [2][n_recv] doubles, then one uint64 flag per source
  pmgx::DevBuf<double*> d_peer_dst;              // per destination: its xbuf + the offset of my segment there
  pmgx::DevBuf<long long> d_peer_stride;         // per destination: its n_recv (distance between its two buffers)
  pmgx::DevBuf<unsigned long long*> d_peer_flag; // per destination: my flag in its xbuf
  pmgx::DevBuf<int> d_send_offsets;              // n_send_nbr + 1
  pmgx::DevBuf<unsigned int> d_ticket;
  std::vector<void*> mapped;
  pmgx::DevBuf<unsigned long long> d_epoch;      // completed exchanges (device-resident: launches carry no per-call state)
  int n_send() const { return send_offsets.empty() ? 0 : send_offsets.back(); }
  int n_recv() const { return recv_offsets.empty() ? 0 : recv_offsets.back(); }
};

namespace pmgx
{
// sub (may be null): the owners send x - sub instead of x (ghost block of x receives it)
void halo_fwd_begin(pmgx_halo* h, double* x, const double* sub = nullptr);
void halo_fwd_end(pmgx_halo* h, double* x);
void halo_setup_p2p(pmgx_halo* h);
cudaStream_t halo_stream(pmgx_halo* h, pmgx_ctx* c);
} // namespace pmgx

// Operator concept of the reference (operator()(in,out) + get_diag_inverse,
// src/chebyshev.hpp:53-56, src/cg.hpp:154-159) as a small polymorphic base.
struct pmgx_operator
{
  enum Kind
  {
    LAPLACIAN = 1,
    CSR = 2
  };
  pmgx_ctx* ctx = nullptr;
  Kind kind = LAPLACIAN;
  int n_owned = 0, n_ghost = 0;
  pmgx_halo* halo = nullptr;          // borrowed
  pmgx::DevBuf<double> diag_inv;      // n_owned
  virtual ~pmgx_operator() {}
  // y = A x (zero fill + halo update of x included)
  virtual void apply(double* x, double* y) = 0;
};

struct pmgx_interp;
namespace pmgx
{
// interpolate / reverse_interpolate (src/interpolate.hpp:185-303) with the fusions the V-cycle uses:
// add: fine += P coarse on owned fine dofs; sub: restrict (fine - sub) with sub taken on owned dofs
void interp_prolong(pmgx_interp* it, double* coarse, double* fine, bool add);
void interp_restrict(pmgx_interp* it, double* fine, const double* sub, double* coarse);
} // namespace pmgx
