// mb_reduce.cu -- standalone micro-benchmark of reduction-kernel variants (tuning
// aid, not product).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false
//   -o build/mb_reduce tools/mb_reduce.cu ;  build/mb_reduce [log2n]
// Variants: final-pass style (none / device / host fence+flag / host 16B store),
// grid cap, unroll, software pipelining, NIN = 1 or 2.  CUDA-event timing over
// REPS launches rotating over buffer sets larger than L2.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess)                                                            \
    {                                                                                 \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));      \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

constexpr int kBlock = 256;

__device__ __forceinline__ void ldg4(const double* p, double (&v)[4])
{
  asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
               : "l"(p)
               : "memory");
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  return v;
}

template <int BLOCK>
__device__ __forceinline__ double block_sum(double v, double* smem)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0)
  {
    r = (lane < BLOCK / 32) ? smem[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

struct Out
{
  double* partials;
  unsigned int* counter;
  double* d_res;
  double* h_res;                       // mapped host
  volatile unsigned long long* h_flag; // mapped host
  unsigned long long seq;
};

// FIN: 0 none, 1 device result, 2 host store + fence.sys + flag, 3 host single 16B store
template <int BLOCK, int FIN>
__device__ __forceinline__ void finish(double v, const Out& o, double* smem)
{
  __shared__ bool s_last;
  if (threadIdx.x == 0)
  {
    o.partials[blockIdx.x] = v;
    if (FIN >= 4)
    {
      unsigned int t;
      asm volatile("atom.add.acq_rel.gpu.u32 %0, [%1], 1;" : "=r"(t) : "l"(o.counter) : "memory");
      s_last = (t == gridDim.x - 1);
    }
    else if (FIN != 0)
    {
      __threadfence();
      const unsigned int t = atomicAdd(o.counter, 1u);
      s_last               = (t == gridDim.x - 1);
    }
  }
  if (FIN == 0) return;
  __syncthreads();
  if (!s_last) return;
  if (FIN < 4) __threadfence();
  double a = 0.0;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += BLOCK) a += __ldcg(o.partials + i);
  a = block_sum<BLOCK>(a, smem);
  if (threadIdx.x == 0)
  {
    *o.d_res   = a;
    *o.counter = 0u;
    if (FIN == 2)
    {
      *o.h_res = a;
      __threadfence_system();
      *o.h_flag = o.seq;
    }
    if (FIN == 3 || FIN == 5)
    {
      asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(o.h_res), "d"(a), "d"(__longlong_as_double((long long)o.seq))
                   : "memory");
    }
  }
}

// PIPE 0: load U tiles, consume, loop.  PIPE 1: prefetch next tile set before consuming.
template <int BLOCK, int U, int NIN, int FIN, int PIPE>
__global__ void __launch_bounds__(BLOCK) k_red(const double* x, const double* y, int64_t n, Out o)
{
  __shared__ double smem[BLOCK / 32];
  constexpr int W        = 4;
  constexpr int64_t TILE = (int64_t)BLOCK * W * U;
  constexpr int64_t STEP = (int64_t)BLOCK * W;
  const int64_t nfull    = n / TILE;
  double acc[W]          = {0, 0, 0, 0};
  if (PIPE == 0)
  {
    for (int64_t t = blockIdx.x; t < nfull; t += gridDim.x)
    {
      const int64_t base = t * TILE + (int64_t)threadIdx.x * W;
      double a[U][W], b[U][W];
#pragma unroll
      for (int u = 0; u < U; u++)
      {
        ldg4(x + base + u * STEP, a[u]);
        if (NIN == 2) ldg4(y + base + u * STEP, b[u]);
      }
#pragma unroll
      for (int u = 0; u < U; u++)
#pragma unroll
        for (int w = 0; w < W; w++) acc[w] += (NIN == 2) ? a[u][w] * b[u][w] : fabs(a[u][w]);
    }
  }
  else
  {
    int64_t t = blockIdx.x;
    double a[U][W], b[U][W];
    if (t < nfull)
    {
      const int64_t base = t * TILE + (int64_t)threadIdx.x * W;
#pragma unroll
      for (int u = 0; u < U; u++)
      {
        ldg4(x + base + u * STEP, a[u]);
        if (NIN == 2) ldg4(y + base + u * STEP, b[u]);
      }
    }
    while (t < nfull)
    {
      const int64_t tn = t + gridDim.x;
      double a2[U][W], b2[U][W];
      if (tn < nfull)
      {
        const int64_t base = tn * TILE + (int64_t)threadIdx.x * W;
#pragma unroll
        for (int u = 0; u < U; u++)
        {
          ldg4(x + base + u * STEP, a2[u]);
          if (NIN == 2) ldg4(y + base + u * STEP, b2[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++)
#pragma unroll
        for (int w = 0; w < W; w++) acc[w] += (NIN == 2) ? a[u][w] * b[u][w] : fabs(a[u][w]);
      if (tn < nfull)
      {
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
          for (int w = 0; w < W; w++)
          {
            a[u][w] = a2[u][w];
            if (NIN == 2) b[u][w] = b2[u][w];
          }
      }
      t = tn;
    }
  }
  double v = ((acc[0] + acc[1]) + acc[2]) + acc[3];
  v        = block_sum<BLOCK>(v, smem);
  finish<BLOCK, FIN>(v, o, smem);
}

// streaming copy / write kernels for context
template <int U>
__global__ void __launch_bounds__(kBlock) k_copy(const double* x, double* z, int64_t n)
{
  constexpr int W        = 4;
  constexpr int64_t TILE = (int64_t)kBlock * W * U;
  constexpr int64_t STEP = (int64_t)kBlock * W;
  const int64_t nfull    = n / TILE;
  for (int64_t t = blockIdx.x; t < nfull; t += gridDim.x)
  {
    const int64_t base = t * TILE + (int64_t)threadIdx.x * W;
    double a[U][W];
#pragma unroll
    for (int u = 0; u < U; u++) ldg4(x + base + u * STEP, a[u]);
#pragma unroll
    for (int u = 0; u < U; u++)
      asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(z + base + u * STEP), "d"(a[u][0]), "d"(a[u][1]),
                   "d"(a[u][2]), "d"(a[u][3])
                   : "memory");
  }
}

template <int U, int PDL>
__global__ void __launch_bounds__(kBlock) k_copy_pdl(const double* x, double* z, int64_t n)
{
  if (PDL)
  {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  constexpr int W        = 4;
  constexpr int64_t TILE = (int64_t)kBlock * W * U;
  constexpr int64_t STEP = (int64_t)kBlock * W;
  const int64_t nfull    = n / TILE;
  for (int64_t t = blockIdx.x; t < nfull; t += gridDim.x)
  {
    const int64_t base = t * TILE + (int64_t)threadIdx.x * W;
    double a[U][W];
#pragma unroll
    for (int u = 0; u < U; u++) ldg4(x + base + u * STEP, a[u]);
#pragma unroll
    for (int u = 0; u < U; u++)
      asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(z + base + u * STEP), "d"(a[u][0]), "d"(a[u][1]),
                   "d"(a[u][2]), "d"(a[u][3])
                   : "memory");
  }
}

template <int U, int PDL>
static void launch_copy_pdl(cudaStream_t st, int grid, const double* x, double* z, int64_t n)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim            = dim3(grid);
  cfg.blockDim           = dim3(kBlock);
  cfg.stream             = st;
  cudaLaunchAttribute at[1];
  at[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = PDL;
  cfg.attrs                                        = at;
  cfg.numAttrs                                     = PDL ? 1 : 0;
  CK(cudaLaunchKernelEx(&cfg, k_copy_pdl<U, PDL>, x, z, n));
}

struct Bench
{
  std::vector<double*> bufs;
  int64_t n;
  Out o;
  double* h_res;
  cudaEvent_t e0, e1;
  int reps = 40;
};

template <class F>
static double time_us(Bench& B, F launch)
{
  for (int r = 0; r < 4; r++) launch(r);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(B.e0));
  for (int r = 0; r < B.reps; r++) launch(r);
  CK(cudaEventRecord(B.e1));
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  float ms;
  CK(cudaEventElapsedTime(&ms, B.e0, B.e1));
  return ms * 1e3 / B.reps;
}

template <int BLOCK, int U, int NIN, int FIN, int PIPE>
static void run(Bench& B, int grid_cap)
{
  const int64_t tiles = B.n / ((int64_t)BLOCK * 4 * U);
  const int grid      = (int)((tiles < grid_cap) ? tiles : grid_cap);
  const int nb        = (int)B.bufs.size();
  double us           = time_us(B, [&](int r) {
    B.o.seq++;
    k_red<BLOCK, U, NIN, FIN, PIPE><<<grid, BLOCK>>>(B.bufs[(2 * r) % nb], B.bufs[(2 * r + 1) % nb], B.n, B.o);
  });
  const double bytes = 8.0 * NIN * B.n;
  printf("red  BLOCK=%d U=%d NIN=%d FIN=%d PIPE=%d grid=%5d : %8.2f us  %7.1f GB/s\n", BLOCK, U, NIN, FIN, PIPE, grid,
         us, bytes / us / 1e3);
  fflush(stdout);
}

template <int BLOCK, int U, int NIN, int PIPE>
static void run_fins(Bench& B, int cap)
{
  run<BLOCK, U, NIN, 0, PIPE>(B, cap);
  run<BLOCK, U, NIN, 1, PIPE>(B, cap);
  run<BLOCK, U, NIN, 2, PIPE>(B, cap);
  run<BLOCK, U, NIN, 3, PIPE>(B, cap);
  run<BLOCK, U, NIN, 4, PIPE>(B, cap);
  run<BLOCK, U, NIN, 5, PIPE>(B, cap);
}

int main(int argc, char** argv)
{
  const int log2n = (argc > 1) ? atoi(argv[1]) : 24;
  Bench B;
  B.n = (int64_t)1 << log2n;
  int nb = (int)(((int64_t)3 << 30) / (8 * B.n)); // 3 GiB of buffers
  if (nb < 4) nb = 4;
  if (nb > 24) nb = 24;
  for (int i = 0; i < nb; i++)
  {
    double* p;
    CK(cudaMalloc(&p, 8 * B.n));
    CK(cudaMemset(p, 0, 8 * B.n));
    B.bufs.push_back(p);
  }
  CK(cudaMalloc(&B.o.partials, 8 * 65536));
  CK(cudaMalloc(&B.o.counter, 4));
  CK(cudaMemset(B.o.counter, 0, 4));
  CK(cudaMalloc(&B.o.d_res, 8));
  CK(cudaHostAlloc(&B.h_res, 64, cudaHostAllocMapped));
  CK(cudaHostGetDevicePointer((void**)&B.o.h_res, B.h_res, 0));
  B.o.h_flag = (volatile unsigned long long*)(B.o.h_res + 2);
  B.o.seq    = 0;
  CK(cudaEventCreate(&B.e0));
  CK(cudaEventCreate(&B.e1));
  printf("n = 2^%d, %d buffers\n", log2n, nb);

  // context: copy kernels and cudaMemcpy D2D
  {
    const int64_t tiles = B.n / (kBlock * 4 * 4);
    double us = time_us(B, [&](int r) { k_copy<4><<<(int)tiles, kBlock>>>(B.bufs[(2 * r) % nb], B.bufs[(2 * r + 1) % nb], B.n); });
    printf("copy U=4 one-tile-per-CTA: %8.2f us %7.1f GB/s\n", us, 16.0 * B.n / us / 1e3);
    us = time_us(B, [&](int r) { CK(cudaMemcpyAsync(B.bufs[(2 * r + 1) % nb], B.bufs[(2 * r) % nb], 8 * B.n, cudaMemcpyDeviceToDevice)); });
    printf("cudaMemcpy D2D           : %8.2f us %7.1f GB/s\n", us, 16.0 * B.n / us / 1e3);
  }

  // PDL: chain of dependent copies z_{k+1} = z_k, with/without programmatic dependent launch
  for (int lg : {16, 18, 20, 22, 24})
  {
    if (lg > log2n) continue;
    const int64_t nn = (int64_t)1 << lg;
    const int U = 4;
    int64_t tiles = nn / (kBlock * 4 * U);
    if (tiles < 1) tiles = 1;
    int64_t saved = B.n;
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    for (int pass = 0; pass < 2; pass++)
    {
      cudaStream_t s = pass ? st : (cudaStream_t)0;
      double us0 = time_us(B, [&](int r) { launch_copy_pdl<4, 0>(s, (int)tiles, B.bufs[r % nb], B.bufs[(r + 1) % nb], nn); });
      double us1 = time_us(B, [&](int r) { launch_copy_pdl<4, 1>(s, (int)tiles, B.bufs[r % nb], B.bufs[(r + 1) % nb], nn); });
      printf("chain copy n=2^%d stream=%s : plain %7.2f us   PDL %7.2f us\n", lg, pass ? "user" : "legacy0", us0, us1);
    }
    CK(cudaStreamDestroy(st));
    B.n = saved;
  }

  const int caps[] = {148 * 2, 148 * 3, 148 * 4, 148 * 6, 148 * 8};
  printf("--- NIN=1 (max-norm / l1 shape), U=4, no pipe, FIN sweep x grid cap\n");
  for (int cap : caps) run_fins<256, 4, 1, 0>(B, cap);
  printf("--- NIN=1 BLOCK=512 U=4\n");
  for (int cap : {148, 148 * 2}) run_fins<512, 4, 1, 0>(B, cap);
  printf("--- NIN=2 (dot shape) U=4\n");
  for (int cap : caps) run_fins<256, 4, 2, 0>(B, cap);
  return 0;
}
