// Ghost-layer halo update: owner -> ghost forward scatter (and ghost -> owner reverse add)
// with grouped ncclSend/ncclRecv over NVLink on a dedicated comm stream, so the exchange
// overlaps the interior-cell kernel running on the compute stream.
// Replaces pack/unpack/unpack_add (src/vector.hpp:24-55) and Vector::scatter_fwd_begin/end,
// scatter_rev_begin/end (src/vector.hpp:186-294), which block the host twice per exchange
// and hand device pointers to GPU-aware MPI.
#include "common.hpp"
#include "operator.hpp"
#include "reduce.cuh"

#include <algorithm>
#include <cstring>

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

// Pack + send in one kernel: entry i of the send list is stored straight into the destination
// rank's receive buffer over NVLink; the last CTA to finish releases this rank's epoch flag on
// every destination (system-scope fence before the ticket, so all CTAs' stores are ordered first).
__global__ void __launch_bounds__(PT)
k_pack_p2p(int ns, int n_nbr, const int* __restrict__ send_offsets, const int32_t* __restrict__ idx,
           const double* __restrict__ x, const double* __restrict__ sub, double* const* __restrict__ peer_dst,
           const long long* __restrict__ peer_stride, unsigned long long* const* __restrict__ peer_flag,
           unsigned long long* __restrict__ epoch_ptr, unsigned int* ticket)
{
  __shared__ bool is_last;
  const unsigned long long epoch = *epoch_ptr + 1; // stored back by the last CTA, after every CTA's ticket
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ns)
  {
    int nb = 0;
    while (i >= send_offsets[nb + 1])
      ++nb;
    const int32_t d = idx[i];
    const double v = sub ? sub[d] * (-1.0) + x[d] : x[d];
    peer_dst[nb][(long long)(epoch & 1ull) * peer_stride[nb] + (i - send_offsets[nb])] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0)
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last)
    return;
  __threadfence_system();
  if ((int)threadIdx.x < n_nbr)
    st_release_sys_u64(peer_flag[threadIdx.x], epoch);
  if (threadIdx.x == 0)
  {
    *ticket = 0u;
    *epoch_ptr = epoch;
  }
}

// One small CTA that waits until every source rank has released this epoch; the unpack kernel
// behind it in the stream then only runs once the data is here (no spinning CTAs holding SMs
// that the interior-cell kernel wants).
__global__ void k_wait_p2p(int n_nbr, const unsigned long long* __restrict__ flags,
                           const unsigned long long* __restrict__ epoch_ptr, unsigned long long timeout_ns)
{
  const unsigned long long epoch = *epoch_ptr; // the pack kernel before this one has advanced it
  for (int t = threadIdx.x; t < n_nbr; t += blockDim.x)
    wait_epoch(flags + t, epoch, timeout_ns); // bounded in wall time, see reduce.cuh
}

// wait + unpack in one kernel for SMALL exchanges on the compute stream (single_stream halos): every CTA
// waits for the sources' flags itself, then unpacks its part
__global__ void __launch_bounds__(PT)
k_wait_unpack_p2p(int n_nbr, const unsigned long long* __restrict__ flags, const unsigned long long* __restrict__ epoch_ptr,
                  unsigned long long timeout_ns, int n, const int32_t* __restrict__ idx, const double* __restrict__ bufs,
                  double* __restrict__ out)
{
  const unsigned long long epoch = *epoch_ptr;
  for (int s = threadIdx.x; s < n_nbr; s += blockDim.x)
    wait_epoch(flags + s, epoch, timeout_ns);
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double* in = bufs + (epoch & 1ull) * (size_t)n;
  if (i < n)
    out[idx[i]] = __ldcg(in + i);
}

__global__ void k_unpack_cg(int n, const int32_t* __restrict__ idx, const double* __restrict__ bufs,
                            const unsigned long long* __restrict__ epoch_ptr, double* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double* in = bufs + (*epoch_ptr & 1ull) * (size_t)n; // receive buffer of this epoch's parity
  if (i < n)
    out[idx[i]] = __ldcg(in + i); // written by a remote GPU: bypass L1
}
} // namespace

// Stream on which work that needs the ghost values can be enqueued right behind the exchange
// (the comm stream while an exchange is in flight, else the compute stream).  Boundary-cell
// kernels launched there start as soon as the ghosts are in place and fill the tail of the
// interior-cell kernel instead of waiting for it; halo_fwd_end joins the two streams.
cudaStream_t halo_stream(pmgx_halo* h, pmgx_ctx* c)
{
  return (h && h->in_flight) ? c->comm_stream : c->stream;
}

static void halo_fwd_begin_p2p(pmgx_halo* h, double* x, const double* sub)
{
  pmgx_ctx* c = h->ctx;
  const int ns = h->n_send(), nr = h->n_recv();
  // single_stream (small exchanges of the AMG levels): everything on the compute stream, no event hops --
  // the pack goes out here, the caller's owned-column work follows it, wait + unpack run in halo_fwd_end
  cudaStream_t cst = h->single_stream ? c->stream : c->comm_stream;
  // a neighbour with an EMPTY segment still takes part in the flag handshake (the coarse levels of
  // the AMG hierarchy pair every send with a receive, amg_setup.cpp): the pack kernel runs whenever
  // there is a destination, the wait whenever there is a source
  if (!h->send_ranks.empty())
  {
    k_pack_p2p<<<std::max((ns + PT - 1) / PT, 1), PT, 0, cst>>>(ns, (int)h->send_ranks.size(), h->d_send_offsets.p,
                                                             h->send_idx.p, x, sub, h->d_peer_dst.p, h->d_peer_stride.p,
                                                             h->d_peer_flag.p, h->d_epoch.p, h->d_ticket.p);
    check_launch("k_pack_p2p");
    count_launch(c);
  }
  if (h->single_stream)
  {
    h->pending_x = x;
    return;
  }
  if (!h->recv_ranks.empty())
  {
    const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(h->xbuf + 2 * (size_t)std::max(nr, 1));
    k_wait_p2p<<<1, 32, 0, c->comm_stream>>>((int)h->recv_ranks.size(), flags, h->d_epoch.p,
                                             c->p2p_timeout_ns);
    count_launch(c);
    if (nr > 0)
    {
      k_unpack_cg<<<(nr + PT - 1) / PT, PT, 0, c->comm_stream>>>(nr, h->recv_idx.p, h->xbuf, h->d_epoch.p, x + h->n_owned);
      count_launch(c);
    }
    check_launch("k_unpack(p2p)");
  }
}

void halo_fwd_begin(pmgx_halo* h, double* x, const double* sub)
{
  pmgx_ctx* c = h->ctx;
  const int ns = h->n_send(), nr = h->n_recv();
  if (h->send_ranks.empty() && h->recv_ranks.empty())
    return;
  if (!h->p2p && ns == 0 && nr == 0)
    return; // NCCL path: empty segments are skipped on both sides
  if (h->p2p && h->single_stream)
  {
    halo_fwd_begin_p2p(h, x, sub);
    return;
  }
  PMGX_CUDA(cudaEventRecord(h->ev_ready, c->stream));
  PMGX_CUDA(cudaStreamWaitEvent(c->comm_stream, h->ev_ready, 0));
  if (h->p2p)
  {
    halo_fwd_begin_p2p(h, x, sub);
    h->in_flight = true;
    return;
  }
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
    if (h->recv_offsets[i + 1] > h->recv_offsets[i])
      PMGX_NCCL(ncclRecv(h->recv_buf.p + h->recv_offsets[i], h->recv_offsets[i + 1] - h->recv_offsets[i],
                         ncclDouble, h->recv_ranks[i], c->comm, c->comm_stream));
  for (size_t i = 0; i < h->send_ranks.size(); ++i)
    if (h->send_offsets[i + 1] > h->send_offsets[i])
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
  h->in_flight = true;
}

void halo_fwd_end(pmgx_halo* h, double* x)
{
  (void)x;
  if (h->pending_x)
  {
    // single-stream exchange: wait for the sources and unpack, right here on the compute stream
    pmgx_ctx* c = h->ctx;
    const int nr = h->n_recv();
    if (!h->recv_ranks.empty())
    {
      const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(h->xbuf + 2 * (size_t)std::max(nr, 1));
      k_wait_unpack_p2p<<<std::max((nr + PT - 1) / PT, 1), PT, 0, c->stream>>>(
          (int)h->recv_ranks.size(), flags, h->d_epoch.p, c->p2p_timeout_ns, nr, h->recv_idx.p, h->xbuf, h->pending_x + h->n_owned);
      check_launch("k_wait_unpack_p2p");
      count_launch(c);
    }
    h->pending_x = nullptr;
    return;
  }
  if (!h->in_flight)
    return;
  // everything enqueued on the comm stream so far -- the exchange and any boundary-cell work the
  // caller put behind it (halo_stream) -- must finish before the compute stream goes on
  PMGX_CUDA(cudaEventRecord(h->ev_done, h->ctx->comm_stream));
  PMGX_CUDA(cudaStreamWaitEvent(h->ctx->stream, h->ev_done, 0));
  h->in_flight = false;
}

// Maps the receive buffers of the neighbours (CUDA IPC) so that k_pack_p2p can store into them.
// Every rank publishes: the IPC handle of its xbuf, its n_recv, and for every possible source
// rank the offset of that source's segment in the receive buffer and its flag index (-1: none).
void halo_setup_p2p(pmgx_halo* h)
{
  pmgx_ctx* c = h->ctx;
  const int R = c->nranks, nr = h->n_recv();
  // the flag handshake needs a symmetric neighbourhood (true for ghost-layer halos)
  std::vector<int> a(h->send_ranks), b(h->recv_ranks);
  std::sort(a.begin(), a.end());
  std::sort(b.begin(), b.end());
  bool ok = a == b;
  const size_t bytes = (2 * (size_t)std::max(nr, 1) + std::max<size_t>(h->recv_ranks.size(), 1)) * sizeof(double);
  cudaIpcMemHandle_t hd;
  std::memset(&hd, 0, sizeof(hd));
  if (ok)
  {
    ok = cudaMalloc(&h->xbuf, bytes) == cudaSuccess && cudaMemsetAsync(h->xbuf, 0, bytes, h->ctx->stream) == cudaSuccess
         && cudaStreamSynchronize(h->ctx->stream) == cudaSuccess
         && cudaIpcGetMemHandle(&hd, h->xbuf) == cudaSuccess;
    cudaGetLastError();
  }
  const size_t rec = sizeof(hd) + sizeof(long long) * (1 + 2 * (size_t)R);
  std::vector<char> mine(rec, 0), all;
  std::memcpy(mine.data(), &hd, sizeof(hd));
  long long* tab = reinterpret_cast<long long*>(mine.data() + sizeof(hd));
  tab[0] = nr;
  for (int r = 0; r < R; ++r)
    tab[1 + r] = -1, tab[1 + R + r] = -1;
  for (size_t i = 0; i < h->recv_ranks.size(); ++i)
  {
    tab[1 + h->recv_ranks[i]] = h->recv_offsets[i];
    tab[1 + R + h->recv_ranks[i]] = (long long)i;
  }
  p2p::allgather_bytes(c, mine.data(), rec, all);
  ok = p2p::all_agree(c, ok);
  const size_t nn = h->send_ranks.size();
  std::vector<double*> dst(nn, nullptr);
  std::vector<long long> stride(nn, 0);
  std::vector<unsigned long long*> flag(nn, nullptr);
  for (size_t i = 0; i < nn && ok; ++i)
  {
    const int p = h->send_ranks[i];
    const char* recp = all.data() + (size_t)p * rec;
    cudaIpcMemHandle_t hp;
    std::memcpy(&hp, recp, sizeof(hp));
    const long long* tp = reinterpret_cast<const long long*>(recp + sizeof(hp));
    const long long p_nr = tp[0], off = tp[1 + c->rank], fidx = tp[1 + R + c->rank];
    const int seg = h->send_offsets[i + 1] - h->send_offsets[i];
    if (off < 0 || fidx < 0 || off + seg > p_nr)
    {
      ok = false;
      break;
    }
    void* base = nullptr;
    if (cudaIpcOpenMemHandle(&base, hp, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
    {
      cudaGetLastError();
      ok = false;
      break;
    }
    h->mapped.push_back(base);
    dst[i] = static_cast<double*>(base) + off;
    stride[i] = p_nr;
    flag[i] = reinterpret_cast<unsigned long long*>(static_cast<double*>(base) + 2 * (size_t)std::max<long long>(p_nr, 1)) + fidx;
  }
  ok = p2p::all_agree(c, ok);
  if (!ok)
  {
    for (void* p : h->mapped)
      cudaIpcCloseMemHandle(p);
    h->mapped.clear();
    if (h->xbuf)
      cudaFree(h->xbuf);
    h->xbuf = nullptr;
    h->p2p = false;
    cudaGetLastError();
    return;
  }
  h->d_peer_dst.upload(dst.data(), nn, c->stream);
  h->d_peer_stride.upload(stride.data(), nn, c->stream);
  h->d_peer_flag.upload(flag.data(), nn, c->stream);
  h->d_send_offsets.upload(h->send_offsets.data(), h->send_offsets.size(), c->stream);
  h->d_ticket.alloc(1);
  h->d_epoch.alloc(1);
  PMGX_CUDA(cudaMemsetAsync(h->d_ticket.p, 0, sizeof(unsigned int), c->stream));
  PMGX_CUDA(cudaMemsetAsync(h->d_epoch.p, 0, sizeof(unsigned long long), c->stream));
  PMGX_CUDA(cudaStreamSynchronize(c->stream));
  h->p2p = true;
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
  if (ctx->p2p)
    pmgx::halo_setup_p2p(h.get()); // collective over all ranks (every rank creates its halos in the same order)
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
    for (void* p : h->mapped)
      cudaIpcCloseMemHandle(p);
    if (h->xbuf)
      cudaFree(h->xbuf);
    delete h;
  }
  PMGX_API_END
}

int pmgx_halo_uses_p2p(pmgx_halo* h) { return h && h->p2p ? 1 : 0; }

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
