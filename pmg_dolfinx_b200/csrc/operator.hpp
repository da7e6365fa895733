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
  // small exchanges (AMG levels): pack, wait and unpack run on the compute stream itself -- no comm-stream
  // hop, no events; set by the owner of the plan, peer-memory path only
  bool single_stream = false;
  double* pending_x = nullptr; // single-stream exchange begun, wait + unpack outstanding
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

namespace pmgx
{
// Chebyshev update fused into a ROW-COMPLETE operator apply (north_star item 2: "the Chebyshev recurrence
// fuses axpy, Jacobi scaling and the residual update into the operator apply, so each smoothing step makes
// one HBM pass").  q = A in is never stored: as soon as row i of it is complete the smoother's vector update
// of entry i (src/chebyshev.hpp:56-68,73-83) is applied.  Possible for gather-type operators (CSR: a row is
// owned by one lane group); the matrix-free operator scatter-adds into shared dofs, so its q is complete
// only after the whole grid has finished and its update stays a separate pass (solvers.cu).
struct ChebEp
{
  enum Mode
  {
    INIT = 0,  // in = x:  r = b - q ; z_out = c0 D^-1 r                                (:56-68)
    STEP = 1,  // in = z:  r -= q ; z_out = c1 z + c2 D^-1 r ; x = z_out + (x [+ z])     (:76-83,73)
    LAST = 2,  // in = z:  r -= q                                                        (:77)
    XONLY = 3  // in = z:  x = (c1 z + c2 D^-1 (r - q)) + (x [+ z]); r and z are dead
  };
  int mode = INIT;
  const double* b = nullptr;
  double* r = nullptr;
  double* z_out = nullptr;
  double* x = nullptr;
  const double* dinv = nullptr;
  double c0 = 0.0, c1 = 0.0, c2 = 0.0;
  int defer = 0;          // the x += z of the previous pass was postponed to this one (1), and x was zero (2)
  double* scratch = nullptr; // n_owned doubles: partial rows that wait for ghost columns
};
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
  // fused apply + Chebyshev update (see ChebEp); false: not supported, the caller runs apply + a pass
  virtual bool supports_cheb_fusion() const { return false; }
  virtual bool apply_cheb(double* in, const pmgx::ChebEp& e)
  {
    (void)in;
    (void)e;
    return false;
  }
};

struct pmgx_interp;
namespace pmgx
{
// interpolate / reverse_interpolate (src/interpolate.hpp:185-303) with the fusions the V-cycle uses:
// add: fine += P coarse on owned fine dofs; sub: restrict (fine - sub) with sub taken on owned dofs
void interp_prolong(pmgx_interp* it, double* coarse, double* fine, bool add);
void interp_restrict(pmgx_interp* it, double* fine, const double* sub, double* coarse);
} // namespace pmgx
