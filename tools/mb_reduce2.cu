// mb_reduce2.cu -- round-2 micro-benchmark of single-output reduction designs (tuning aid, not
// product).  Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo
//   -o build/mb_reduce2 tools/mb_reduce2.cu ;  build/mb_reduce2 [log2n ...]
//
// What it separates (VERDICT r01 "Next" 3 i-iv):
//   * threads per SM (1024 vs 2048), loads in flight per thread, software pipelining,
//     tile partition (round-robin vs one contiguous range per CTA), TMA bulk ring;
//   * the epilogue: none / ticket + last-CTA pass / tagged partials + polling finisher CTA;
//   * where an API call's time goes: host launch -> first CTA (%globaltimer), streaming,
//     final pass, publication to pinned memory, host poll.
// Every configuration prints: kernel-only (back-to-back async launches, CUDA events) and API
// (launch + spin on the pinned result, host clock, mean and min) times.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess)                                                       \
    {                                                                            \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(1);                                                                   \
    }                                                                            \
  } while (0)

__device__ __forceinline__ void ldg4(const double* p, double (&v)[4])
{
  asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
               : "l"(p)
               : "memory");
}
__device__ __forceinline__ unsigned long long gtime()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
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
  return r; // valid in thread 0
}

struct Out
{
  double* partials;             // [grid]                (ticket scheme)
  unsigned long long* tagged;   // [grid][2] tagged words (polling scheme)
  unsigned int* counter;
  double* d_res;
  unsigned long long* h_words;  // mapped pinned: [0..1] tagged pair, [2..3] {value, seq} 16-byte pair
  unsigned long long* stamps;   // device: [0] min CTA start, [1] finisher has all partials, [2] published; or NULL
  unsigned int seq;
};

__device__ __forceinline__ void publish_tagged(const Out& o, double a)
{
  const unsigned long long bits = (unsigned long long)__double_as_longlong(a);
  const unsigned long long tag  = (unsigned long long)o.seq << 32;
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(o.h_words), "l"(tag | (bits & 0xffffffffull)) : "memory");
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(o.h_words + 1), "l"(tag | (bits >> 32)) : "memory");
}
__device__ __forceinline__ void publish_pair16(const Out& o, double a)
{
  asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(o.h_words + 2), "d"(a),
               "d"(__longlong_as_double((long long)o.seq))
               : "memory");
}

// FIN 0: partial only.  1: ticket (acq_rel) + last-CTA pass + 16-byte pair (round-1 product).
// 2: tagged partials, CTA 0 polls them (no atomic, no fence), tagged publication.
// 3: ticket + last-CTA pass + tagged publication.
template <int BLOCK, int FIN>
__device__ __forceinline__ void finish(double v /* thread 0 */, const Out& o, double* smem)
{
  if (FIN == 0)
  {
    if (threadIdx.x == 0) o.partials[blockIdx.x] = v;
    return;
  }
  if (FIN == 1 || FIN == 3)
  {
    __shared__ bool s_last;
    if (threadIdx.x == 0)
    {
      o.partials[blockIdx.x] = v;
      unsigned int t;
      asm volatile("atom.add.acq_rel.gpu.u32 %0, [%1], 1;" : "=r"(t) : "l"(o.counter) : "memory");
      s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    if (o.stamps && threadIdx.x == 0) o.stamps[1] = gtime();
    double a = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += BLOCK) a += __ldcg(o.partials + i);
    a = block_sum<BLOCK>(a, smem);
    if (threadIdx.x == 0)
    {
      *o.counter = 0u;
      *o.d_res   = a;
      if (FIN == 1) publish_pair16(o, a);
      else publish_tagged(o, a);
      if (o.stamps) o.stamps[2] = gtime();
    }
    return;
  }
  if (FIN == 2)
  {
    if (threadIdx.x == 0)
    {
      const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
      const unsigned long long tag  = (unsigned long long)o.seq << 32;
      unsigned long long* dst       = o.tagged + 2 * (size_t)blockIdx.x;
      asm volatile("st.volatile.global.v2.u64 [%0], {%1,%2};" ::"l"(dst), "l"(tag | (bits & 0xffffffffull)),
                   "l"(tag | (bits >> 32))
                   : "memory");
    }
    if (blockIdx.x != 0) return;
    double a = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += BLOCK)
    {
      const unsigned long long* src = o.tagged + 2 * (size_t)i;
      unsigned long long w0, w1;
      for (;;)
      {
        asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
        if ((unsigned int)(w0 >> 32) == o.seq && (unsigned int)(w1 >> 32) == o.seq) break;
      }
      a += __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull)));
    }
    if (o.stamps)
    {
      __syncthreads();
      if (threadIdx.x == 0) o.stamps[1] = gtime();
    }
    a = block_sum<BLOCK>(a, smem);
    if (threadIdx.x == 0)
    {
      *o.d_res = a;
      publish_tagged(o, a);
      if (o.stamps) o.stamps[2] = gtime();
    }
  }
}

// PART 0: tiles round-robin over CTAs; 1: one contiguous tile range per CTA.
// PIPE 1: next tile's loads are issued before the current tile is folded.
template <int BLOCK, int MINB, int U, int NIN, int PART, int PIPE, int FIN, int PDL = 0>
__global__ void __launch_bounds__(BLOCK, MINB) k_red(const double* x, const double* y, const double* z, int64_t n, Out o)
{
  __shared__ double smem[BLOCK / 32];
  if (PDL)
  {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  constexpr int W        = 4;
  constexpr int64_t TILE = (int64_t)BLOCK * W * U;
  constexpr int64_t STEP = (int64_t)BLOCK * W;
  const int64_t nfull    = n / TILE;
  if (o.stamps && threadIdx.x == 0) atomicMin(o.stamps, gtime());
  int64_t t0, t1, dt;
  if (PART == 0) { t0 = blockIdx.x; t1 = nfull; dt = gridDim.x; }
  else
  {
    t0 = nfull * blockIdx.x / gridDim.x;
    t1 = nfull * (blockIdx.x + 1) / gridDim.x;
    dt = 1;
  }
  double acc[W] = {0, 0, 0, 0};
  auto load = [&](int64_t t, double (&a)[U][W], double (&b)[U][W], double (&c)[U][W]) {
    const int64_t base = t * TILE + (int64_t)threadIdx.x * W;
#pragma unroll
    for (int u = 0; u < U; u++)
    {
      ldg4(x + base + u * STEP, a[u]);
      if (NIN >= 2) ldg4(y + base + u * STEP, b[u]);
      if (NIN >= 3) ldg4(z + base + u * STEP, c[u]);
    }
  };
  auto fold = [&](double (&a)[U][W], double (&b)[U][W], double (&c)[U][W]) {
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int w = 0; w < W; w++)
      {
        double term;
        if (NIN == 1) term = fabs(a[u][w]);
        else if (NIN == 2) term = a[u][w] * b[u][w];
        else
        {
          const double p = a[u][w] * b[u][w];
          term           = (c[u][w] > 0.0) ? p * p : 0.0;
        }
        acc[w] += term;
      }
  };
  if (PIPE == 0)
  {
    for (int64_t t = t0; t < t1; t += dt)
    {
      double a[U][W], b[U][W], c[U][W];
      load(t, a, b, c);
      fold(a, b, c);
    }
  }
  else
  {
    double a[U][W], b[U][W], c[U][W];
    int64_t t = t0;
    if (t < t1) load(t, a, b, c);
    while (t < t1)
    {
      const int64_t tn = t + dt;
      double a2[U][W], b2[U][W], c2[U][W];
      if (tn < t1) load(tn, a2, b2, c2);
      fold(a, b, c);
      if (tn < t1)
      {
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
          for (int w = 0; w < W; w++)
          {
            a[u][w] = a2[u][w];
            if (NIN >= 2) b[u][w] = b2[u][w];
            if (NIN >= 3) c[u][w] = c2[u][w];
          }
      }
      t = tn;
    }
  }
  double v = ((acc[0] + acc[1]) + acc[2]) + acc[3];
  v        = block_sum<BLOCK>(v, smem);
  __syncthreads();
  finish<BLOCK, FIN>(v, o, smem);
}

// ---------------------------------------------------------------- TMA bulk ring
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  uint32_t ok = 0;
  for (;;)
  {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (ok) return;
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// CW consumer warps + 1 producer warp; a stage holds CHUNK doubles of each of the NIN operands.
template <int CW, int STAGES, int CHUNK, int NIN, int FIN>
__global__ void __launch_bounds__(CW * 32 + 32) k_red_tma(const double* x, const double* y, const double* z, int64_t n, Out o)
{
  constexpr int BLOCK = CW * 32 + 32;
  constexpr int CT    = CW * 32;
  extern __shared__ __align__(128) unsigned char raw[];
  double* buf         = reinterpret_cast<double*>(raw);
  __shared__ __align__(8) unsigned long long bars[2 * STAGES];
  __shared__ double smem[BLOCK / 32];
  const uint32_t full0 = smem_addr(bars), empty0 = smem_addr(bars + STAGES);
  if (o.stamps && threadIdx.x == 0) atomicMin(o.stamps, gtime());
  if (threadIdx.x == 0)
  {
    for (int s = 0; s < STAGES; s++)
    {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, CW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int64_t nch = n / CHUNK;
  const int64_t c0 = blockIdx.x, dc = gridDim.x;
  const int64_t mine = (nch > c0) ? (nch - c0 + dc - 1) / dc : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double acc[4] = {0, 0, 0, 0};
  if (warp == CW)
  {
    if (lane == 0)
    {
      for (int64_t i = 0; i < mine; i++)
      {
        const int s = (int)(i % STAGES);
        if (i >= STAGES) mbar_wait(empty0 + 8 * s, (uint32_t)(((i / STAGES) - 1) & 1));
        const int64_t off = (c0 + i * dc) * CHUNK;
        mbar_expect_tx(full0 + 8 * s, NIN * CHUNK * 8);
        const uint32_t dst = smem_addr(buf + (size_t)s * NIN * CHUNK);
        bulk_g2s(dst, x + off, CHUNK * 8, full0 + 8 * s);
        if (NIN >= 2) bulk_g2s(dst + CHUNK * 8, y + off, CHUNK * 8, full0 + 8 * s);
        if (NIN >= 3) bulk_g2s(dst + 2 * CHUNK * 8, z + off, CHUNK * 8, full0 + 8 * s);
      }
    }
  }
  else
  {
    constexpr int PER = CHUNK / (CT * 2); // double2 per thread per stage
    static_assert(PER >= 1 && CHUNK % (CT * 2) == 0, "chunk must split into double2 per consumer thread");
    for (int64_t i = 0; i < mine; i++)
    {
      const int s = (int)(i % STAGES);
      mbar_wait(full0 + 8 * s, (uint32_t)((i / STAGES) & 1));
      const double2* bx = reinterpret_cast<const double2*>(buf + (size_t)s * NIN * CHUNK);
      const double2* by = bx + CHUNK / 2;
      const double2* bz = by + CHUNK / 2;
#pragma unroll
      for (int k = 0; k < PER; k++)
      {
        const double2 a = bx[k * CT + threadIdx.x];
        double t0, t1;
        if (NIN == 1) { t0 = fabs(a.x); t1 = fabs(a.y); }
        else
        {
          const double2 b = by[k * CT + threadIdx.x];
          if (NIN == 2) { t0 = a.x * b.x; t1 = a.y * b.y; }
          else
          {
            const double2 c = bz[k * CT + threadIdx.x];
            const double p0 = a.x * b.x, p1 = a.y * b.y;
            t0 = (c.x > 0.0) ? p0 * p0 : 0.0;
            t1 = (c.y > 0.0) ? p1 * p1 : 0.0;
          }
        }
        acc[(2 * k) & 3] += t0;
        acc[(2 * k + 1) & 3] += t1;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * s);
    }
  }
  double v = ((acc[0] + acc[1]) + acc[2]) + acc[3];
  v        = block_sum<BLOCK>(v, smem);
  __syncthreads();
  finish<BLOCK, FIN>(v, o, smem);
}

// ------------------------------------------------------------------------ host
struct Bench
{
  std::vector<double*> bufs;
  int64_t n = 0;
  Out o;
  unsigned long long* h_words = nullptr;
  unsigned long long* d_stamps = nullptr;
  cudaEvent_t e0, e1;
  int reps = 30;
};

static double now_us()
{
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// wait until the kernel of sequence seq has published (fin: 1 = 16-byte pair, 2/3 = tagged words)
static void host_wait(Bench& B, int fin, unsigned int seq)
{
  volatile unsigned long long* w = B.h_words;
  if (fin == 1)
  {
    while (w[3] != (unsigned long long)seq) {}
  }
  else
  {
    while ((unsigned int)(w[0] >> 32) != seq || (unsigned int)(w[1] >> 32) != seq) {}
  }
}

template <class L>
static void measure(Bench& B, const char* name, int grid, int fin, double bytes, L launch)
{
  const int nb = (int)B.bufs.size();
  B.o.stamps   = nullptr;
  for (int r = 0; r < 4; r++) { B.o.seq++; launch(r); }
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  CK(cudaEventRecord(B.e0));
  for (int r = 0; r < B.reps; r++) { B.o.seq++; launch(r); }
  CK(cudaEventRecord(B.e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, B.e0, B.e1));
  const double k_us = ms * 1e3 / B.reps;
  double api_mean = 0, api_min = 1e30, dev_mean = 0, fin_mean = 0, pub_mean = 0;
  if (fin != 0)
  {
    std::vector<double> v;
    for (int r = 0; r < B.reps; r++)
    {
      B.o.seq++;
      const double t0 = now_us();
      launch(r);
      host_wait(B, fin, B.o.seq);
      v.push_back(now_us() - t0);
    }
    CK(cudaDeviceSynchronize());
    for (double d : v) { api_mean += d; api_min = std::min(api_min, d); }
    api_mean /= v.size();
    // stamped launches (not timed): where the device-side time goes
    B.o.stamps = B.d_stamps;
    const int ns = 8;
    for (int r = 0; r < ns; r++)
    {
      unsigned long long init[3] = {~0ull, 0, 0}, got[3];
      CK(cudaMemcpy(B.d_stamps, init, sizeof(init), cudaMemcpyHostToDevice));
      B.o.seq++;
      launch(r);
      host_wait(B, fin, B.o.seq);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(got, B.d_stamps, sizeof(got), cudaMemcpyDeviceToHost));
      dev_mean += (double)(got[2] - got[0]) * 1e-3;
      fin_mean += (double)(got[2] - got[1]) * 1e-3;
      pub_mean += (double)(got[1] - got[0]) * 1e-3;
    }
    dev_mean /= ns; fin_mean /= ns; pub_mean /= ns;
    B.o.stamps = nullptr;
  }
  (void)nb;
  printf("%-44s grid=%5d | kernel %7.2f us %7.1f GB/s", name, grid, k_us, bytes / k_us / 1e3);
  if (fin != 0)
    printf(" | api %7.2f (min %6.2f) us %7.1f GB/s | dev start->pub %6.2f  (stream %6.2f + final %5.2f)  host+launch %5.2f",
           api_mean, api_min, bytes / api_mean / 1e3, dev_mean, pub_mean, fin_mean, api_mean - dev_mean);
  printf("\n");
  fflush(stdout);
}

template <int BLOCK, int MINB, int U, int NIN, int PART, int PIPE, int FIN>
static void run(Bench& B, int cap)
{
  const int64_t tiles = std::max<int64_t>(1, B.n / ((int64_t)BLOCK * 4 * U));
  const int grid      = (int)std::min<int64_t>(tiles, cap);
  const int nb        = (int)B.bufs.size();
  char name[128];
  snprintf(name, sizeof(name), "ldg B=%d minb=%d U=%d NIN=%d part=%s pipe=%d fin=%d", BLOCK, MINB, U, NIN, PART ? "blk" : "rr",
           PIPE, FIN);
  measure(B, name, grid, FIN, 8.0 * NIN * B.n, [&](int r) {
    k_red<BLOCK, MINB, U, NIN, PART, PIPE, FIN><<<grid, BLOCK>>>(B.bufs[(3 * r) % nb], B.bufs[(3 * r + 1) % nb],
                                                                   B.bufs[(3 * r + 2) % nb], B.n, B.o);
  });
}

// PDL 1: griddepcontrol in the kernel, plain launch.  2: + programmatic-stream-serialization launch attribute
template <int BLOCK, int MINB, int U, int NIN, int FIN, int PDL>
static void run_pdl(Bench& B, int cap)
{
  const int64_t tiles = std::max<int64_t>(1, B.n / ((int64_t)BLOCK * 4 * U));
  const int grid      = (int)std::min<int64_t>(tiles, cap);
  const int nb        = (int)B.bufs.size();
  char name[128];
  snprintf(name, sizeof(name), "ldg B=%d minb=%d U=%d NIN=%d fin=%d PDL=%d%s", BLOCK, MINB, U, NIN, FIN, PDL,
           PDL == 2 ? " (+launch attr)" : PDL == 1 ? " (griddepcontrol only)" : "");
  measure(B, name, grid, FIN, 8.0 * NIN * B.n, [&](int r) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim            = dim3(grid);
    cfg.blockDim           = dim3(BLOCK);
    cfg.stream             = 0;
    cudaLaunchAttribute at[1];
    at[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs                                        = at;
    cfg.numAttrs                                     = (PDL == 2) ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, k_red<BLOCK, MINB, U, NIN, 0, 0, FIN, (PDL > 0)>, (const double*)B.bufs[(3 * r) % nb],
                          (const double*)B.bufs[(3 * r + 1) % nb], (const double*)B.bufs[(3 * r + 2) % nb], B.n, B.o));
  });
}

template <int CW, int STAGES, int CHUNK, int NIN, int FIN>
static void run_tma(Bench& B, int cap)
{
  const int64_t chunks = std::max<int64_t>(1, B.n / CHUNK);
  const int grid       = (int)std::min<int64_t>(chunks, cap);
  const int nb         = (int)B.bufs.size();
  const size_t smem    = (size_t)STAGES * NIN * CHUNK * 8;
  if (smem > 220 * 1024) return;
  CK(cudaFuncSetAttribute(k_red_tma<CW, STAGES, CHUNK, NIN, FIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  char name[128];
  snprintf(name, sizeof(name), "tma cw=%d stages=%d chunk=%d NIN=%d fin=%d (%zu KB)", CW, STAGES, CHUNK, NIN, FIN, smem >> 10);
  measure(B, name, grid, FIN, 8.0 * NIN * B.n, [&](int r) {
    k_red_tma<CW, STAGES, CHUNK, NIN, FIN><<<grid, CW * 32 + 32, smem>>>(B.bufs[(3 * r) % nb], B.bufs[(3 * r + 1) % nb],
                                                                          B.bufs[(3 * r + 2) % nb], B.n, B.o);
  });
}

template <int NIN>
static void sweep(Bench& B)
{
  const int S = 148;
  printf("---- NIN=%d  n=2^%d\n", NIN, (int)__builtin_ctzll((unsigned long long)B.n));
  // launch path: plain <<<>>> vs cudaLaunchKernelEx, griddepcontrol, PDL launch attribute
  run_pdl<512, 2, 4, NIN, 2, 0>(B, 2 * S);
  run_pdl<512, 2, 4, NIN, 2, 1>(B, 2 * S);
  run_pdl<512, 2, 4, NIN, 2, 2>(B, 2 * S);
  // same buffers every launch (what a per-op timing loop does) instead of rotating over 3 GiB
  {
    std::vector<double*> keep = B.bufs;
    B.bufs.resize(3);
    run<512, 2, 4, NIN, 0, 0, 2>(B, 2 * S);
    B.bufs = keep;
  }
  // round-1 product shape and its epilogue variants
  run<512, 2, 4, NIN, 0, 0, 0>(B, 2 * S);
  run<512, 2, 4, NIN, 0, 0, 1>(B, 2 * S);
  run<512, 2, 4, NIN, 0, 0, 3>(B, 2 * S);
  run<512, 2, 4, NIN, 0, 0, 2>(B, 2 * S);
  // contiguous range per CTA
  run<512, 2, 4, NIN, 1, 0, 0>(B, 2 * S);
  run<512, 2, 4, NIN, 1, 0, 2>(B, 2 * S);
  // software pipelined (register double buffer)
  run<512, 2, 2, NIN, 0, 1, 0>(B, 2 * S);
  run<512, 2, 2, NIN, 0, 1, 2>(B, 2 * S);
  run<256, 4, 4, NIN, 0, 1, 2>(B, 4 * S);
  // 2048 threads per SM
  run<512, 4, 2, NIN, 0, 0, 0>(B, 4 * S);
  run<512, 4, 2, NIN, 0, 0, 2>(B, 4 * S);
  run<256, 8, 2, NIN, 0, 0, 0>(B, 8 * S);
  run<256, 8, 2, NIN, 0, 0, 2>(B, 8 * S);
  run<1024, 2, 2, NIN, 0, 0, 2>(B, 2 * S);
  run<1024, 2, 2, NIN, 1, 0, 2>(B, 2 * S);
  run<512, 4, 1, NIN, 0, 1, 2>(B, 4 * S);
  // one CTA per SM
  run<1024, 1, 4, NIN, 0, 0, 2>(B, S);
  run<512, 1, 4, NIN, 0, 0, 2>(B, S);
  // TMA bulk ring
  run_tma<8, 6, 2048, NIN, 0>(B, S);
  run_tma<8, 6, 2048, NIN, 2>(B, S);
  run_tma<8, 4, 2048, NIN, 2>(B, 2 * S);
  run_tma<8, 8, 1024, NIN, 2>(B, 2 * S);
  run_tma<4, 8, 1024, NIN, 2>(B, 2 * S);
  run_tma<8, 3, 4096, NIN, 2>(B, S);
}

int main(int argc, char** argv)
{
  std::vector<int> sizes;
  for (int i = 1; i < argc; i++) sizes.push_back(atoi(argv[i]));
  if (sizes.empty()) sizes = {24};
  Bench B;
  CK(cudaMalloc(&B.o.partials, 8 * 65536));
  CK(cudaMalloc(&B.o.tagged, 16 * 65536));
  CK(cudaMemset(B.o.tagged, 0, 16 * 65536));
  CK(cudaMalloc(&B.o.counter, 4));
  CK(cudaMemset(B.o.counter, 0, 4));
  CK(cudaMalloc(&B.o.d_res, 8));
  CK(cudaMalloc(&B.d_stamps, 64));
  CK(cudaHostAlloc(&B.h_words, 64, cudaHostAllocMapped));
  memset(B.h_words, 0, 64);
  CK(cudaHostGetDevicePointer((void**)&B.o.h_words, B.h_words, 0));
  B.o.seq    = 0;
  B.o.stamps = nullptr;
  CK(cudaEventCreate(&B.e0));
  CK(cudaEventCreate(&B.e1));
  for (int lg : sizes)
  {
    B.n    = (int64_t)1 << lg;
    int nb = (int)(((int64_t)3 << 30) / (8 * B.n)); // 3 GiB of buffers rotate (> L2 for n >= 2^22)
    nb     = std::max(6, std::min(nb, 24));
    for (double* p : B.bufs) CK(cudaFree(p));
    B.bufs.clear();
    for (int i = 0; i < nb; i++)
    {
      double* p;
      CK(cudaMalloc(&p, 8 * B.n));
      CK(cudaMemset(p, 0, 8 * B.n));
      B.bufs.push_back(p);
    }
    sweep<1>(B);
    sweep<2>(B);
    if (lg == 24) sweep<3>(B);
  }
  return 0;
}
