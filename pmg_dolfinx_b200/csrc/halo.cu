// Ghost-layer halo update: owner -> ghost forward scatter (and ghost -> owner reverse add)
// with grouped ncclSend/ncclRecv over NVLink on a dedicated comm stream, so the exchange
// overlaps the interior-cell kernel running on the compute stream.
// Replaces pack/unpack/unpack_add (src/vector.hpp:24-55) and Vector::scatter_fwd_begin/end,
// scatter_rev_begin/end (src/vector.hpp:186-294), which block the host twice per exchange
// and hand device pointers to GPU-aware MPI.
#include "common.hpp"
#include "operator.hpp"

namespace pmgx
{
namespace
{
__global__ void k_pack(int n, const int32_t* __restrict__ idx, const double* __restrict__ in,
                       double* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = in[idx[i]];
}
__global__ void k_pack_sub(int n, const int32_t* __restrict__ idx, const double* __restrict__ in,
                           const double* __restrict__ sub, double* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = sub[idx[i]] * (-1.0) + in[idx[i]];
}
__global__ void k_unpack(int n, const int32_t* __restrict__ idx, const double* __restrict__ in,
                         double* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    out[idx[i]] = in[i];
}
__global__ void k_unpack_add(int n, const int32_t* __restrict__ idx, const double* __restrict__ in,
                             double* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    atomicAdd(&out[idx[i]], in[i]);
}
constexpr int PT = 256;
} // namespace

void halo_fwd_begin(pmgx_halo* h, double* x, const double* sub)
{
  pmgx_ctx* c = h->ctx;
  const int ns = h->n_send(), nr = h->n_recv();
  if (ns == 0 && nr == 0)
    return;
  PMGX_CUDA(cudaEventRecord(h->ev_ready, c->stream));
  PMGX_CUDA(cudaStreamWaitEvent(c->comm_stream, h->ev_ready, 0));
  if (ns > 0)
  {
    if (sub)
      k_pack_sub<<<(ns + PT - 1) / PT, PT, 0, c->comm_stream>>>(ns, h->send_idx.p, x, sub, h->send_buf.p);
    else
      k_pack<<<(ns + PT - 1) / PT, PT, 0, c->comm_stream>>>(ns, h->send_idx.p, x, h->send_buf.p);
    check_launch("k_pack");
    count_launch(c);
  }
  PMGX_REQUIRE(c->comm != nullptr, "halo exchange needs a multi-rank context");
  PMGX_NCCL(ncclGroupStart());
  for (size_t i = 0; i < h->recv_ranks.size(); ++i)
    PMGX_NCCL(ncclRecv(h->recv_buf.p + h->recv_offsets[i], h->recv_offsets[i + 1] - h->recv_offsets[i],
                       ncclDouble, h->recv_ranks[i], c->comm, c->comm_stream));
  for (size_t i = 0; i < h->send_ranks.size(); ++i)
    PMGX_NCCL(ncclSend(h->send_buf.p + h->send_offsets[i], h->send_offsets[i + 1] - h->send_offsets[i],
                       ncclDouble, h->send_ranks[i], c->comm, c->comm_stream));
  PMGX_NCCL(ncclGroupEnd());
  if (nr > 0)
  {
    // interior cells never read the ghost block, so the unpack can run on the comm stream
    k_unpack<<<(nr + PT - 1) / PT, PT, 0, c->comm_stream>>>(nr, h->recv_idx.p, h->recv_buf.p,
                                                            x + h->n_owned);
    check_launch("k_unpack");
    count_launch(c);
  }
  PMGX_CUDA(cudaEventRecord(h->ev_done, c->comm_stream));
  h->in_flight = true;
}

void halo_fwd_end(pmgx_halo* h, double* x)
{
  (void)x;
  if (!h->in_flight)
    return;
  PMGX_CUDA(cudaStreamWaitEvent(h->ctx->stream, h->ev_done, 0));
  h->in_flight = false;
}

static void halo_rev(pmgx_halo* h, double* x)
{
  pmgx_ctx* c = h->ctx;
  const int ns = h->n_send(), nr = h->n_recv();
  if (ns == 0 && nr == 0)
    return;
  PMGX_REQUIRE(c->comm != nullptr, "halo exchange needs a multi-rank context");
  PMGX_CUDA(cudaEventRecord(h->ev_ready, c->stream));
  PMGX_CUDA(cudaStreamWaitEvent(c->comm_stream, h->ev_ready, 0));
  if (nr > 0)
  {
    k_pack<<<(nr + PT - 1) / PT, PT, 0, c->comm_stream>>>(nr, h->recv_idx.p, x + h->n_owned,
                                                          h->recv_buf.p);
    check_launch("k_pack(rev)");
    count_launch(c);
  }
  PMGX_NCCL(ncclGroupStart());
  for (size_t i = 0; i < h->send_ranks.size(); ++i)
    PMGX_NCCL(ncclRecv(h->send_buf.p + h->send_offsets[i], h->send_offsets[i + 1] - h->send_offsets[i],
                       ncclDouble, h->send_ranks[i], c->comm, c->comm_stream));
  for (size_t i = 0; i < h->recv_ranks.size(); ++i)
    PMGX_NCCL(ncclSend(h->recv_buf.p + h->recv_offsets[i], h->recv_offsets[i + 1] - h->recv_offsets[i],
                       ncclDouble, h->recv_ranks[i], c->comm, c->comm_stream));
  PMGX_NCCL(ncclGroupEnd());
  if (ns > 0)
  {
    k_unpack_add<<<(ns + PT - 1) / PT, PT, 0, c->comm_stream>>>(ns, h->send_idx.p, h->send_buf.p, x);
    check_launch("k_unpack_add");
    count_launch(c);
  }
  PMGX_CUDA(cudaEventRecord(h->ev_done, c->comm_stream));
  PMGX_CUDA(cudaStreamWaitEvent(c->stream, h->ev_done, 0));
}
} // namespace pmgx

extern "C"
{
int pmgx_halo_create(pmgx_ctx* ctx, int n_owned, int n_ghost, int n_send_nbr,
                     const int* send_ranks_h, const int* send_offsets_h, const int32_t* send_idx_h,
                     int n_recv_nbr, const int* recv_ranks_h, const int* recv_offsets_h,
                     const int32_t* recv_idx_h, pmgx_halo** out)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(ctx && out, "halo_create: null ctx/out");
  PMGX_REQUIRE(n_owned >= 0 && n_ghost >= 0 && n_send_nbr >= 0 && n_recv_nbr >= 0, "halo_create: bad sizes");
  PMGX_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<pmgx_halo> h(new pmgx_halo());
  h->ctx = ctx;
  h->n_owned = n_owned;
  h->n_ghost = n_ghost;
  h->send_offsets.assign(1, 0);
  h->recv_offsets.assign(1, 0);
  for (int i = 0; i < n_send_nbr; ++i)
  {
    PMGX_REQUIRE(send_ranks_h[i] >= 0 && send_ranks_h[i] < ctx->nranks && send_ranks_h[i] != ctx->rank,
                 "halo_create: bad destination rank %d", send_ranks_h[i]);
    PMGX_REQUIRE(send_offsets_h[i + 1] >= send_offsets_h[i], "halo_create: send offsets not monotone");
    h->send_ranks.push_back(send_ranks_h[i]);
    h->send_offsets.push_back(send_offsets_h[i + 1]);
  }
  for (int i = 0; i < n_recv_nbr; ++i)
  {
    PMGX_REQUIRE(recv_ranks_h[i] >= 0 && recv_ranks_h[i] < ctx->nranks && recv_ranks_h[i] != ctx->rank,
                 "halo_create: bad source rank %d", recv_ranks_h[i]);
    PMGX_REQUIRE(recv_offsets_h[i + 1] >= recv_offsets_h[i], "halo_create: recv offsets not monotone");
    h->recv_ranks.push_back(recv_ranks_h[i]);
    h->recv_offsets.push_back(recv_offsets_h[i + 1]);
  }
  PMGX_REQUIRE(n_send_nbr == 0 || send_offsets_h[0] == 0, "halo_create: send offsets must start at 0");
  PMGX_REQUIRE(n_recv_nbr == 0 || recv_offsets_h[0] == 0, "halo_create: recv offsets must start at 0");
  const int ns = h->n_send(), nr = h->n_recv();
  for (int i = 0; i < ns; ++i)
    PMGX_REQUIRE(send_idx_h[i] >= 0 && send_idx_h[i] < n_owned, "halo_create: send index out of range");
  for (int i = 0; i < nr; ++i)
    PMGX_REQUIRE(recv_idx_h[i] >= 0 && recv_idx_h[i] < n_ghost, "halo_create: ghost slot out of range");
  h->send_idx.upload(send_idx_h, ns, ctx->stream);
  h->recv_idx.upload(recv_idx_h, nr, ctx->stream);
  h->send_buf.alloc(ns);
  h->recv_buf.alloc(nr);
  PMGX_CUDA(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
  PMGX_CUDA(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
  *out = h.release();
  PMGX_API_END
}

int pmgx_halo_destroy(pmgx_halo* h)
{
  PMGX_API_BEGIN
  if (h)
  {
    cudaSetDevice(h->ctx->device);
    cudaStreamSynchronize(h->ctx->comm_stream);
    cudaStreamSynchronize(h->ctx->stream);
    if (h->ev_ready)
      cudaEventDestroy(h->ev_ready);
    if (h->ev_done)
      cudaEventDestroy(h->ev_done);
    delete h;
  }
  PMGX_API_END
}

int pmgx_halo_fwd_begin(pmgx_halo* h, double* x)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(h && x, "halo_fwd_begin: null argument");
  PMGX_CUDA(cudaSetDevice(h->ctx->device));
  pmgx::halo_fwd_begin(h, x);
  PMGX_API_END
}

int pmgx_halo_fwd_end(pmgx_halo* h, double* x)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(h && x, "halo_fwd_end: null argument");
  PMGX_CUDA(cudaSetDevice(h->ctx->device));
  pmgx::halo_fwd_end(h, x);
  PMGX_API_END
}

int pmgx_halo_rev(pmgx_halo* h, double* x)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(h && x, "halo_rev: null argument");
  PMGX_CUDA(cudaSetDevice(h->ctx->device));
  pmgx::halo_rev(h, x);
  PMGX_API_END
}

int pmgx_pack(pmgx_ctx* c, int n, const int32_t* idx, const double* in, double* out)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(c && n >= 0, "pack: bad arguments");
  if (n > 0)
  {
    PMGX_CUDA(cudaSetDevice(c->device));
    pmgx::k_pack<<<(n + pmgx::PT - 1) / pmgx::PT, pmgx::PT, 0, c->stream>>>(n, idx, in, out);
    pmgx::check_launch("k_pack");
    pmgx::count_launch(c);
  }
  PMGX_API_END
}

int pmgx_unpack(pmgx_ctx* c, int n, const int32_t* idx, const double* in, double* out)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(c && n >= 0, "unpack: bad arguments");
  if (n > 0)
  {
    PMGX_CUDA(cudaSetDevice(c->device));
    pmgx::k_unpack<<<(n + pmgx::PT - 1) / pmgx::PT, pmgx::PT, 0, c->stream>>>(n, idx, in, out);
    pmgx::check_launch("k_unpack");
    pmgx::count_launch(c);
  }
  PMGX_API_END
}

int pmgx_unpack_add(pmgx_ctx* c, int n, const int32_t* idx, const double* in, double* out)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(c && n >= 0, "unpack_add: bad arguments");
  if (n > 0)
  {
    PMGX_CUDA(cudaSetDevice(c->device));
    pmgx::k_unpack_add<<<(n + pmgx::PT - 1) / pmgx::PT, pmgx::PT, 0, c->stream>>>(n, idx, in, out);
    pmgx::check_launch("k_unpack_add");
    pmgx::count_launch(c);
  }
  PMGX_API_END
}
}
