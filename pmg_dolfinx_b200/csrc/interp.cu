// p-level prolongation / restriction as batched element-local sum-factorised kernels
// (north_star item 4).  Replaces interpolate_Q1Q2 / interpolate_Q2Q1 and Interpolator<T>
// (src/interpolate.hpp:21-87,93-329): one thread block handles a batch of cells, the local
// operator M = M1d (x) M1d (x) M1d is applied direction by direction in shared memory instead
// of a per-cell CSR walk by a single thread; cell lists live on the device (no per-call
// allocation, quirk Q10) and the fine-dof multiplicity is computed on the device.
#include "common.hpp"
#include "operator.hpp"

namespace pmgx
{
namespace
{
constexpr int IT = 256;

__global__ void k_count_mult(const int32_t* __restrict__ dm, long long total, double* __restrict__ mult)
{
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += nth)
    atomicAdd(&mult[dm[t]], 1.0);
}
__global__ void k_invert_mult(double* __restrict__ m, int n)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    m[i] = m[i] > 0.0 ? 1.0 / m[i] : 0.0;
}

// One tensor direction: out[a0][a1][a2] with the axis `dir` contracted against M.
//   expand (T=false): out index o in [0,no), in index i in [0,ni): out = sum_i M[o*ni+i] in
//   reduce (T=true) : out index o in [0,no), in index i in [0,ni): out = sum_i M[i*no+o] in
// dims: sizes of the three axes of `in`; axis `dir` has size ni and becomes no in `out`.
template <bool T>
__device__ __forceinline__ void contract_axis(const double* __restrict__ in, double* __restrict__ out,
                                              const double* __restrict__ M, int d0, int d1, int d2,
                                              int dir, int no, int tid, int nthreads)
{
  const int ni = dir == 0 ? d0 : (dir == 1 ? d1 : d2);
  const int o0 = dir == 0 ? no : d0, o1 = dir == 1 ? no : d1, o2 = dir == 2 ? no : d2;
  const int total = o0 * o1 * o2;
  const int istride = dir == 0 ? d1 * d2 : (dir == 1 ? d2 : 1);
  for (int t = tid; t < total; t += nthreads)
  {
    const int a0 = t / (o1 * o2), a1 = (t / o2) % o1, a2 = t % o2;
    const int o = dir == 0 ? a0 : (dir == 1 ? a1 : a2);
    const int base = (dir == 0 ? 0 : a0 * d1 * d2) + (dir == 1 ? 0 : a1 * d2) + (dir == 2 ? 0 : a2);
    double s = 0.0;
    for (int i = 0; i < ni; ++i)
      s = fma(T ? M[i * no + o] : M[o * ni + i], in[base + i * istride], s);
    out[t] = s;
  }
}

// One warp per cell (grid-stride over the cell list): the three 1-D passes run warp-synchronously
// on a private shared-memory slice, so there is no block barrier and a CTA keeps several cells
// in flight.  Prolongation: fine[dofs_f[j]] = sum_k M[j,k] coarse[dofs_c[k]] (overwrite; :21-45)
__global__ void __launch_bounds__(IT)
k_prolong(int nc, int nf, const double* __restrict__ M1, const int32_t* __restrict__ cells,
          int first, int count, const int32_t* __restrict__ dm_c, const int32_t* __restrict__ dm_f,
          const double* __restrict__ xc, double* __restrict__ xf)
{
  extern __shared__ double sm[];
  const int nc3 = nc * nc * nc, nf3 = nf * nf * nf;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  double* sM = sm;                            // nf*nc
  double* b0 = sM + nf * nc + warp * 2 * nf3; // nf3 per warp
  double* b1 = b0 + nf3;
  for (int t = threadIdx.x; t < nf * nc; t += blockDim.x)
    sM[t] = M1[t];
  __syncthreads();
  for (long long ci = (long long)blockIdx.x * wpb + warp; ci < count; ci += (long long)gridDim.x * wpb)
  {
    const long long cell = cells[first + ci];
    for (int t = lane; t < nc3; t += 32)
      b0[t] = xc[dm_c[cell * nc3 + t]];
    __syncwarp();
    contract_axis<false>(b0, b1, sM, nc, nc, nc, 2, nf, lane, 32); // [nc][nc][nf]
    __syncwarp();
    contract_axis<false>(b1, b0, sM, nc, nc, nf, 1, nf, lane, 32); // [nc][nf][nf]
    __syncwarp();
    contract_axis<false>(b0, b1, sM, nc, nf, nf, 0, nf, lane, 32); // [nf][nf][nf]
    __syncwarp();
    for (int t = lane; t < nf3; t += 32)
      xf[dm_f[cell * nf3 + t]] = b1[t];
    __syncwarp();
  }
}

// Restriction: coarse[dofs_c[j]] += sum_k M[k,j] fine[d_k] / mult[d_k] (atomics; :60-87)
__global__ void __launch_bounds__(IT)
k_restrict(int nc, int nf, const double* __restrict__ M1, const int32_t* __restrict__ cells,
           int first, int count, const int32_t* __restrict__ dm_c, const int32_t* __restrict__ dm_f,
           const double* __restrict__ xf, const double* __restrict__ inv_mult, double* __restrict__ xc)
{
  extern __shared__ double sm[];
  const int nc3 = nc * nc * nc, nf3 = nf * nf * nf;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  double* sM = sm;
  double* b0 = sM + nf * nc + warp * 2 * nf3;
  double* b1 = b0 + nf3;
  for (int t = threadIdx.x; t < nf * nc; t += blockDim.x)
    sM[t] = M1[t];
  __syncthreads();
  for (long long ci = (long long)blockIdx.x * wpb + warp; ci < count; ci += (long long)gridDim.x * wpb)
  {
    const long long cell = cells[first + ci];
    for (int t = lane; t < nf3; t += 32)
    {
      const int32_t d = dm_f[cell * nf3 + t];
      b0[t] = xf[d] * inv_mult[d];
    }
    __syncwarp();
    contract_axis<true>(b0, b1, sM, nf, nf, nf, 0, nc, lane, 32); // [nc][nf][nf]
    __syncwarp();
    contract_axis<true>(b1, b0, sM, nc, nf, nf, 1, nc, lane, 32); // [nc][nc][nf]
    __syncwarp();
    contract_axis<true>(b0, b1, sM, nc, nc, nf, 2, nc, lane, 32); // [nc][nc][nc]
    __syncwarp();
    for (int t = lane; t < nc3; t += 32)
      atomicAdd(&xc[dm_c[cell * nc3 + t]], b1[t]);
    __syncwarp();
  }
}
} // namespace
} // namespace pmgx

struct pmgx_interp
{
  pmgx_ctx* ctx = nullptr;
  int pc = 0, pf = 0, n_cells = 0, n_l = 0, n_b = 0;
  int n_coarse_total = 0, n_fine_total = 0;
  const int32_t* dm_c = nullptr; // borrowed device dofmaps (src/interpolate.hpp:322-323)
  const int32_t* dm_f = nullptr;
  pmgx_halo* halo_c = nullptr;
  pmgx_halo* halo_f = nullptr;
  pmgx::DevBuf<int32_t> cells;   // lcells then bcells
  pmgx::DevBuf<double> M1;       // [nf][nc]
  pmgx::DevBuf<double> inv_mult; // 1 / multiplicity of fine dofs
  // warps per block: as many as fit ~40 KB of private slices (2 nf^3 doubles each), at most 8
  int wpb() const
  {
    const int nf = pf + 1;
    return std::max(1, std::min(8, (int)(40960 / (2 * nf * nf * nf * sizeof(double)))));
  }
  int tpb() const { return 32 * wpb(); }
  size_t smem() const
  {
    const int nc = pc + 1, nf = pf + 1;
    return (size_t)(nf * nc + wpb() * 2 * nf * nf * nf) * sizeof(double);
  }
  int grid(int count) const
  {
    const int blocks = (count + wpb() - 1) / wpb();
    return std::max(1, std::min(blocks, ctx->num_sms * 16));
  }
};

extern "C"
{
int pmgx_interp_create(pmgx_ctx* ctx, int degree_coarse, int degree_fine, int n_cells,
                       const int32_t* dofmap_coarse, const int32_t* dofmap_fine, int n_coarse_total,
                       int n_fine_total, const int32_t* lcells_h, int n_lcells,
                       const int32_t* bcells_h, int n_bcells, pmgx_halo* halo_c, pmgx_halo* halo_f,
                       pmgx_interp** out)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(ctx && out, "interp_create: null ctx/out");
  PMGX_REQUIRE(degree_coarse >= 1 && degree_fine <= PMGX_MAX_DEGREE && degree_coarse <= degree_fine,
               "Unsupported degrees %d -> %d", degree_coarse, degree_fine);
  PMGX_REQUIRE(n_cells >= 0 && n_lcells >= 0 && n_bcells >= 0 && n_lcells + n_bcells <= n_cells,
               "interp_create: inconsistent cell counts");
  PMGX_REQUIRE(n_cells == 0 || (dofmap_coarse && dofmap_fine), "interp_create: null dofmap");
  PMGX_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<pmgx_interp> it(new pmgx_interp());
  it->ctx = ctx;
  it->pc = degree_coarse;
  it->pf = degree_fine;
  it->n_cells = n_cells;
  it->n_l = n_lcells;
  it->n_b = n_bcells;
  it->n_coarse_total = n_coarse_total;
  it->n_fine_total = n_fine_total;
  it->dm_c = dofmap_coarse;
  it->dm_f = dofmap_fine;
  it->halo_c = halo_c;
  it->halo_f = halo_f;
  std::vector<int32_t> cl((size_t)n_lcells + n_bcells);
  for (int i = 0; i < n_lcells; ++i)
    cl[i] = lcells_h[i];
  for (int i = 0; i < n_bcells; ++i)
    cl[n_lcells + i] = bcells_h[i];
  for (int32_t v : cl)
    PMGX_REQUIRE(v >= 0 && v < n_cells, "interp_create: cell index out of range");
  it->cells.upload(cl.data(), cl.size(), ctx->stream);
  std::vector<double> M;
  pmgx::gll_interp_matrix(degree_coarse, degree_fine, M);
  for (double& v : M) // the reference drops |v| <= 1e-12 when compressing (interpolate.hpp:120-128)
    if (std::fabs(v) <= 1e-12)
      v = 0.0;
  it->M1.upload(M.data(), M.size(), ctx->stream);
  // multiplicity over the whole fine dofmap (all local + ghost cells, :172-178)
  it->inv_mult.alloc((size_t)n_fine_total);
  if (n_fine_total > 0)
  {
    PMGX_CUDA(cudaMemsetAsync(it->inv_mult.p, 0, (size_t)n_fine_total * sizeof(double), ctx->stream));
    const int nf = degree_fine + 1;
    const long long total = (long long)n_cells * nf * nf * nf;
    if (total > 0)
    {
      const int g = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, ctx->num_sms * 32));
      pmgx::k_count_mult<<<g, 256, 0, ctx->stream>>>(dofmap_fine, total, it->inv_mult.p);
      pmgx::check_launch("k_count_mult");
    }
    pmgx::k_invert_mult<<<(n_fine_total + 255) / 256, 256, 0, ctx->stream>>>(it->inv_mult.p, n_fine_total);
    pmgx::check_launch("k_invert_mult");
    pmgx::count_launch(ctx, 2);
  }
  PMGX_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = it.release();
  PMGX_API_END
}

int pmgx_interp_prolong(pmgx_interp* it, double* coarse, double* fine)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(it && coarse && fine, "interp_prolong: null argument");
  pmgx_ctx* c = it->ctx;
  PMGX_CUDA(cudaSetDevice(c->device));
  const int nc = it->pc + 1, nf = it->pf + 1;
  if (it->halo_c)
    pmgx::halo_fwd_begin(it->halo_c, coarse);                                        // :202
  if (it->n_l > 0)
  {
    pmgx::k_prolong<<<it->grid(it->n_l), it->tpb(), it->smem(), c->stream>>>(
        nc, nf, it->M1.p, it->cells.p, 0, it->n_l, it->dm_c, it->dm_f, coarse, fine); // :208
    pmgx::check_launch("k_prolong");
    pmgx::count_launch(c);
  }
  if (it->halo_c)
    pmgx::halo_fwd_end(it->halo_c, coarse);                                          // :217
  if (it->n_b > 0)
  {
    pmgx::k_prolong<<<it->grid(it->n_b), it->tpb(), it->smem(), c->stream>>>(
        nc, nf, it->M1.p, it->cells.p, it->n_l, it->n_b, it->dm_c, it->dm_f, coarse, fine); // :227
    pmgx::check_launch("k_prolong");
    pmgx::count_launch(c);
  }
  PMGX_API_END
}

int pmgx_interp_restrict(pmgx_interp* it, double* fine, double* coarse)
{
  PMGX_API_BEGIN
  PMGX_REQUIRE(it && coarse && fine, "interp_restrict: null argument");
  pmgx_ctx* c = it->ctx;
  PMGX_CUDA(cudaSetDevice(c->device));
  const int nc = it->pc + 1, nf = it->pf + 1;
  if (it->halo_f)
    pmgx::halo_fwd_begin(it->halo_f, fine);                                          // :264
  PMGX_CUDA(cudaMemsetAsync(coarse, 0, (size_t)it->n_coarse_total * sizeof(double), c->stream)); // :270
  if (it->n_l > 0)
  {
    pmgx::k_restrict<<<it->grid(it->n_l), it->tpb(), it->smem(), c->stream>>>(
        nc, nf, it->M1.p, it->cells.p, 0, it->n_l, it->dm_c, it->dm_f, fine, it->inv_mult.p, coarse);
    pmgx::check_launch("k_restrict");
    pmgx::count_launch(c);
  }
  if (it->halo_f)
    pmgx::halo_fwd_end(it->halo_f, fine);                                            // :281
  if (it->n_b > 0)
  {
    pmgx::k_restrict<<<it->grid(it->n_b), it->tpb(), it->smem(), c->stream>>>(
        nc, nf, it->M1.p, it->cells.p, it->n_l, it->n_b, it->dm_c, it->dm_f, fine, it->inv_mult.p, coarse);
    pmgx::check_launch("k_restrict");
    pmgx::count_launch(c);
  }
  PMGX_API_END
}

int pmgx_interp_destroy(pmgx_interp* it)
{
  PMGX_API_BEGIN
  if (it)
  {
    cudaSetDevice(it->ctx->device);
    cudaStreamSynchronize(it->ctx->stream);
    delete it;
  }
  PMGX_API_END
}
}
