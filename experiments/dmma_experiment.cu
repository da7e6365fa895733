// FP64 DMMA experiment (north_star item 1): is the FP64 tensor-core path
// (mma.sync m8n8k4 f64) worth using for the 1-D contractions of the sum-factorised element kernel at
// high degree, where the contraction is densest (P7: n = 8 = the m8n8k4 tile)?
//
// Measured here, on the device, with CUDA events:
//   1. peak issue rate of the FP64 FMA pipe (independent DFMA chains)
//   2. peak issue rate of DMMA m8n8k4 (independent accumulator tiles)
//   3. both interleaved in one kernel (do they share a pipe?)
//   4. the z contraction of a P7 element, gz[(i,j),l] = sum_m U[(i,j),m] D[l][m], batched over cells that are
//      resident in registers (as in the apply kernel): (a) register FMAs with U rows per thread,
//      (b) DMMA with the fragments loaded in the m8n8k4 layout (8 row tiles x 2 k-steps per cell)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o dmma_experiment dmma_experiment.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                                      \
  do                                                                                               \
  {                                                                                                \
    cudaError_t e = (x);                                                                           \
    if (e != cudaSuccess)                                                                          \
    {                                                                                              \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);               \
      exit(1);                                                                                     \
    }                                                                                              \
  } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b, double c0, double c1)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1)
               : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

constexpr int CH = 8; // independent chains per thread

__global__ void k_dfma_peak(double* out, int iters, double a, double b)
{
  double acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c)
    acc[c] = threadIdx.x * 1e-9 + c;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int c = 0; c < CH; ++c)
      acc[c] = fma(acc[c], a, b);
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c)
    s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma_peak(double* out, int iters, double a, double b)
{
  double c0[CH / 2], c1[CH / 2];
#pragma unroll
  for (int c = 0; c < CH / 2; ++c)
    c0[c] = threadIdx.x * 1e-9 + c, c1[c] = c;
  const double fa = a + threadIdx.x * 1e-12, fb = b;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int c = 0; c < CH / 2; ++c)
      dmma884(c0[c], c1[c], fa, fb, c0[c], c1[c]);
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH / 2; ++c)
    s += c0[c] + c1[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mixed_peak(double* out, int iters, double a, double b)
{
  double acc[CH], c0[CH / 2], c1[CH / 2];
#pragma unroll
  for (int c = 0; c < CH; ++c)
    acc[c] = threadIdx.x * 1e-9 + c;
#pragma unroll
  for (int c = 0; c < CH / 2; ++c)
    c0[c] = threadIdx.x * 1e-9 + c, c1[c] = c;
  const double fa = a + threadIdx.x * 1e-12, fb = b;
  for (int it = 0; it < iters; ++it)
  {
#pragma unroll
    for (int c = 0; c < CH / 2; ++c)
      dmma884(c0[c], c1[c], fa, fb, c0[c], c1[c]);
#pragma unroll
    for (int c = 0; c < CH; ++c)
      acc[c] = fma(acc[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c)
    s += acc[c];
#pragma unroll
  for (int c = 0; c < CH / 2; ++c)
    s += c0[c] + c1[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- the z contraction of a P7 element held in registers, repeated `reps` times per cell (stand-in for the
// six contractions of an apply).  D in shared memory.
constexpr int N = 8;

// (a) FMA: thread = one (i,j) row of the cell: 8 inputs, 8 outputs, 64 DFMA; a warp holds half a cell
__global__ void __launch_bounds__(256) k_zcontract_fma(const double* __restrict__ U, double* __restrict__ out,
                                                       const double* __restrict__ Dg, int ncells, int reps)
{
  __shared__ double D[N * N];
  if (threadIdx.x < N * N)
    D[threadIdx.x] = Dg[threadIdx.x];
  __syncthreads();
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; // global (cell, ij) row
  if (row >= (long long)ncells * N * N)
    return;
  double u[N], g[N];
#pragma unroll
  for (int m = 0; m < N; ++m)
    u[m] = U[row * N + m];
  for (int r = 0; r < reps; ++r)
  {
#pragma unroll
    for (int l = 0; l < N; ++l)
    {
      double s = 0.0;
#pragma unroll
      for (int m = 0; m < N; ++m)
        s = fma(D[l * N + m], u[m], s);
      g[l] = s;
    }
#pragma unroll
    for (int m = 0; m < N; ++m)
      u[m] = g[m] * 0.5; // feed back: keeps the chain live
  }
#pragma unroll
  for (int m = 0; m < N; ++m)
    out[row * N + m] = u[m];
}

// (b) DMMA: a warp owns 8 rows x 8 columns tiles; per row tile two k-steps.  A fragment: thread holds
// U[row = lane/4][k = lane%4 (+4)], B fragment: D^T[k = lane%4 (+4)][col = lane/4], C: [row = lane/4][2*(lane%4)+{0,1}].
// Feeding the result back as the next A operand needs the C -> A re-layout (4 shuffles of doubles per tile).
__global__ void __launch_bounds__(256) k_zcontract_dmma(const double* __restrict__ U, double* __restrict__ out,
                                                        const double* __restrict__ Dg, int ncells, int reps)
{
  const int lane = threadIdx.x & 31;
  const long long tile = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; // 8-row tile index
  if (tile >= (long long)ncells * N)
    return;
  const int r = lane >> 2, q = lane & 3;
  const double b0 = Dg[r * N + q], b1 = Dg[r * N + 4 + q]; // D^T[k][col] = D[col][k]
  const double* base = U + tile * (8 * N);
  double a0 = base[r * N + q], a1 = base[r * N + 4 + q];
  double c0 = 0.0, c1 = 0.0;
  for (int it = 0; it < reps; ++it)
  {
    c0 = 0.0, c1 = 0.0;
    dmma884(c0, c1, a0, b0, c0, c1);
    dmma884(c0, c1, a1, b1, c0, c1);
    // C[row r][cols 2q, 2q+1] -> A[row r][cols q and 4+q]: column c lives in lane (r*4 + c/2), element c%2
    const double h0 = c0 * 0.5, h1 = c1 * 0.5;
    const int s0 = (r << 2) + (q >> 1), s1 = (r << 2) + ((4 + q) >> 1);
    const double x0 = __shfl_sync(0xffffffffu, h0, s0), x1 = __shfl_sync(0xffffffffu, h1, s0);
    const double y0 = __shfl_sync(0xffffffffu, h0, s1), y1 = __shfl_sync(0xffffffffu, h1, s1);
    a0 = (q & 1) ? x1 : x0;
    a1 = (q & 1) ? y1 : y0;
  }
  double* o = out + tile * (8 * N);
  o[r * N + q] = a0;
  o[r * N + 4 + q] = a1;
}

// latency / per-warp pipelining: ONE warp, NCH independent accumulator chains, `iters` dependent steps each
template <int NCH>
__global__ void k_dmma_latency(double* out, long long* cycles, int iters, double a, double b)
{
  double c0[NCH], c1[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c)
    c0[c] = threadIdx.x * 1e-9 + c, c1[c] = c;
  const double fa = a + threadIdx.x * 1e-12, fb = b;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      dmma884(c0[c], c1[c], fa, fb, c0[c], c1[c]);
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < NCH; ++c)
    s += c0[c] + c1[c];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0)
    *cycles = t1 - t0;
}

template <int NCH>
__global__ void k_dfma_latency(double* out, long long* cycles, int iters, double a, double b)
{
  double acc[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c)
    acc[c] = threadIdx.x * 1e-9 + c;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      acc[c] = fma(acc[c], a, b);
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < NCH; ++c)
    s += acc[c];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0)
    *cycles = t1 - t0;
}

template <typename F>
float time_ms(F f, int reps = 5)
{
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  f();
  f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i)
    f();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

int main(int argc, char** argv)
{
  const int mode = argc > 1 ? atoi(argv[1]) : 0; // 0 all; 1..5 single kernel; 9: every kernel exactly once (for ncu)
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  printf("device %s, %d SMs\n", p.name, sms);
  const int blocks = sms * 8, threads = 256, iters = 4096;
  double* out;
  CK(cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)));
  const double nthreads = (double)blocks * threads;
  if (mode == 0 || mode == 1)
  {
    float ms = time_ms([&] { k_dfma_peak<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf("DFMA peak      : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * CH * iters * nthreads / ms / 1e9);
  }
  if (mode == 0 || mode == 2)
  {
    float ms = time_ms([&] { k_dmma_peak<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    // one m8n8k4 = 256 FMA per warp = 8 FMA per thread
    printf("DMMA m8n8k4    : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * 8 * (CH / 2) * iters * nthreads / ms / 1e9);
  }
  if (mode == 0 || mode == 3)
  {
    float ms = time_ms([&] { k_mixed_peak<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf("DFMA + DMMA    : %8.3f ms  %7.2f TFLOP/s (sum of both)\n", ms,
           2.0 * (CH + 8 * (CH / 2)) * iters * nthreads / ms / 1e9);
  }
  if (mode == 0 || mode == 6)
  {
    long long* cyc;
    CK(cudaMalloc(&cyc, sizeof(long long)));
    long long h = 0;
    const int it = 2048;
#define LAT(K, N)                                                                                  \
  K<N><<<1, 32>>>(out, cyc, it, 1.0000001, 1e-9);                                                  \
  CK(cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost));                                      \
  printf("  %-16s one warp, %d independent chain(s): %7.1f cycles per step (%.1f per instruction)\n", #K, N,       \
         (double)h / it, (double)h / it / N);
    printf("latency / per-warp pipelining (clock64 around a dependent loop):\n");
    LAT(k_dfma_latency, 1)
    LAT(k_dfma_latency, 4)
    LAT(k_dmma_latency, 1)
    LAT(k_dmma_latency, 2)
    LAT(k_dmma_latency, 4)
    LAT(k_dmma_latency, 8)
#undef LAT
  }
  // z contraction of P7 elements
  const int ncells = sms * 2048, reps = 64;
  const size_t nU = (size_t)ncells * N * N * N;
  double *U, *V, *D;
  CK(cudaMalloc(&U, nU * sizeof(double)));
  CK(cudaMalloc(&V, nU * sizeof(double)));
  CK(cudaMalloc(&D, N * N * sizeof(double)));
  {
    double* h = (double*)malloc(nU * sizeof(double));
    for (size_t i = 0; i < nU; ++i)
      h[i] = (double)((i * 2654435761u) % 1000) / 1000.0 - 0.5;
    CK(cudaMemcpy(U, h, nU * sizeof(double), cudaMemcpyHostToDevice));
    double hd[N * N];
    for (int i = 0; i < N * N; ++i)
      hd[i] = ((i * 7) % 11 - 5) / 8.0;
    CK(cudaMemcpy(D, hd, sizeof(hd), cudaMemcpyHostToDevice));
    free(h);
  }
  const double flops = 2.0 * ncells * (double)N * N * N * N * reps;
  if (mode == 0 || mode == 4)
  {
    const long long rows = (long long)ncells * N * N;
    float ms = time_ms([&] { k_zcontract_fma<<<(int)((rows + 255) / 256), 256>>>(U, V, D, ncells, reps); });
    printf("z-contraction n=8, FMA  (%d reps per load): %8.3f ms  %7.2f TFLOP/s\n", reps, ms, flops / ms / 1e9);
  }
  if (mode == 0 || mode == 5)
  {
    const long long tiles = (long long)ncells * N;
    float ms = time_ms([&] { k_zcontract_dmma<<<(int)((tiles * 32 + 255) / 256), 256>>>(U, V, D, ncells, reps); });
    printf("z-contraction n=8, DMMA (%d reps per load): %8.3f ms  %7.2f TFLOP/s\n", reps, ms, flops / ms / 1e9);
  }
  if (mode == 9)
  {
    // one launch of each kernel, no warm-up: ncu --set full captures all five in one invocation
    const long long rows = (long long)ncells * N * N, tiles = (long long)ncells * N;
    k_dfma_peak<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    k_dmma_peak<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    k_mixed_peak<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    k_zcontract_fma<<<(int)((rows + 255) / 256), 256>>>(U, V, D, ncells, reps);
    k_zcontract_dmma<<<(int)((tiles * 32 + 255) / 256), 256>>>(U, V, D, ncells, reps);
    CK(cudaDeviceSynchronize());
    printf("mode 9: five kernels launched once\n");
  }
  // the two contraction kernels compute the same thing
  if (mode == 0)
  {
    double *hv = (double*)malloc(nU * sizeof(double)), *hw = (double*)malloc(nU * sizeof(double));
    const long long rows = (long long)ncells * N * N, tiles = (long long)ncells * N;
    k_zcontract_fma<<<(int)((rows + 255) / 256), 256>>>(U, V, D, ncells, 3);
    CK(cudaMemcpy(hv, V, nU * sizeof(double), cudaMemcpyDeviceToHost));
    k_zcontract_dmma<<<(int)((tiles * 32 + 255) / 256), 256>>>(U, V, D, ncells, 3);
    CK(cudaMemcpy(hw, V, nU * sizeof(double), cudaMemcpyDeviceToHost));
    double err = 0, nrm = 0;
    for (size_t i = 0; i < nU; ++i)
      err = fmax(err, fabs(hv[i] - hw[i])), nrm = fmax(nrm, fabs(hv[i]));
    printf("max |FMA - DMMA| = %.3e (max |value| %.3e)\n", err, nrm);
  }
  return 0;
}
