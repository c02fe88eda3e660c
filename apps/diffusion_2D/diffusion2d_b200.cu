/* diffusion2d_b200.cu -- re-host of the reference's benchmarks/diffusion_2D
 * (ARKODE DIRK + PCG + Jacobi) on NVECTOR_B200, without MPI.
 *
 * What the reference does per right-hand-side evaluation (mpi_gpu/diffusion.cpp:341-397):
 *   pack_buffers kernel -> 4x MPI_Irecv/Isend of device buffers (CUDA-aware MPI) ->
 *   interior kernel -> MPI_Wait -> boundary kernel, and 6 sin/cos per grid point
 *   for the forcing term (:38-57).
 * Here it is ONE kernel (k_diffusion_rhs):
 *   - 1-D strips in y; each CTA marches R rows of a 1024-point x-tile keeping the
 *     three stencil rows in registers (each u value is read from HBM once),
 *     west/east neighbours by warp shuffle;
 *   - the CTAs that own the strip's first / last row first PUSH that row into the
 *     neighbour rank's halo buffer over NVLink peer memory (symmetric allocation,
 *     b200vec_comm_peer_alloc) and bump the neighbour's arrival counter, then
 *     compute their other rows, and only at the very end wait for the
 *     neighbour's row and finish the halo-dependent row: exchange and stencil
 *     overlap inside the kernel, no pack kernel, no host-side wait;
 *   - the forcing b(t,x,y) is separable: per-x and per-y factor tables (built
 *     once with the host libm, exactly the factors of mpi_serial/diffusion.cpp:60-90)
 *     and two per-call time factors give the reference's expression with ~10
 *     flops per point instead of 6 transcendental calls.
 * Everything else -- vectors, Krylov dot products, norms, linear combinations --
 * is NVECTOR_B200 called by the unmodified ARKODE / SUNLinSol_PCG.
 */
#include <arkode/arkode_arkstep.h>
#include <cuda_runtime.h>
#include <sunlinsol/sunlinsol_pcg.h>
#include <sunlinsol/sunlinsol_spgmr.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "diffusion2d_b200.h"
#include "nvector_b200.h"

#define PI_ 3.141592653589793238462643383279502884197169

namespace {

constexpr int kThreads = 256;

struct RhsArgs
{
  const double* u;
  double* f;
  int64_t nx, ny, ny_loc, js;
  double cx, cy, cc;
  double stct, c2t; /* sin(pi t) cos(pi t), cos^2(pi t) */
  int forcing;
  const double *dcx, *ssx;      /* per x: bx (cos^2 - sin^2), sin^2                    */
  const double *ssy, *dcy;      /* per local y: sin^2, by (cos^2 - sin^2)          */
  /* halo exchange over peer memory */
  int hasS, hasN;
  double* sendS;       /* S neighbour's "row from the north" buffer (this call's parity) */
  double* sendN;       /* N neighbour's "row from the south" buffer                      */
  const double* recvS; /* own buffers                                                     */
  const double* recvN;
  unsigned long long* ctrS_remote; /* S neighbour's arrivals-from-north counter */
  unsigned long long* ctrN_remote; /* N neighbour's arrivals-from-south counter */
  const unsigned long long* ctrS_local;
  const unsigned long long* ctrN_local;
  unsigned long long expected; /* call number x x-tiles: counters are cumulative */
};

template <int W>
__device__ __forceinline__ void ld_row(const double* p, double (&v)[W])
{
  if (W == 4)
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
  else v[0] = *p;
}
/* halo rows were written by another GPU: read them at L2 */
template <int W>
__device__ __forceinline__ void ld_halo(const double* p, double (&v)[W])
{
  if (W == 4)
    asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
  else asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v[0]) : "l"(p) : "memory");
}
template <int W>
__device__ __forceinline__ void st_row(double* p, const double (&v)[W])
{
  if (W == 4)
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
  else *p = v[0];
}

__device__ __forceinline__ void wait_counter(const unsigned long long* ctr, unsigned long long expected)
{
  if (threadIdx.x == 0)
  {
    unsigned long long v, t0 = 0;
    unsigned int spins = 0;
    for (;;)
    {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
      if (v >= expected) break;
      if ((++spins & 0xfffu) == 0)
      {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 30000000000ull) asm volatile("trap;"); /* dead neighbour: fail loudly, do not hang */
      }
    }
  }
  __syncthreads();
}

/* one row of this thread's W points in flight: the values and, in lane 0 / lane 31
   only, the west / east neighbour of the thread's span (the other lanes get theirs by
   shuffle).  Loaded together so that no stencil row waits on a dependent load. */
template <int W>
struct Row
{
  double v[W];
  double edge;
};

template <int W>
__device__ __forceinline__ void load_row(const RhsArgs& a, int64_t j, int64_t xo, int64_t x0, int lane, Row<W>& r)
{
  const double* row = a.u + j * a.nx;
  ld_row<W>(row + xo, r.v);
  if (lane == 0 && x0 > 0 && x0 < a.nx) r.edge = row[x0 - 1];
  if (lane == 31 && x0 + W < a.nx) r.edge = row[x0 + W];
}

/* one stencil row: s/c/n = south/centre/north values of this thread's W points, w/e the
   west/east neighbours of the span; sy, dy = the row's forcing factors sin^2(pi y),
   by (cos^2 - sin^2)(pi y).  Arithmetic in the reference's order (-fmad=false). */
template <int W>
__device__ __forceinline__ void stencil_store(const RhsArgs& a, int64_t j, int64_t x0, const double (&s)[W],
                                              const double (&c)[W], const double (&n)[W], double w, double e, double sy,
                                              double dy, const double (&tdcx)[W], const double (&tssx)[W])
{
  const int64_t jg  = a.js + j;
  const bool ybound = (jg == 0) || (jg == a.ny - 1);
  double r[W];
#pragma unroll
  for (int k = 0; k < W; k++)
  {
    const int64_t i = x0 + k;
    const double uw = (k == 0) ? w : c[k - 1];
    const double ue = (k == W - 1) ? e : c[k + 1];
    double v        = 0.0;
    if (!ybound && i > 0 && i < a.nx - 1)
    {
      /* mpi_serial/diffusion.cpp:98-101 / mpi_gpu/diffusion.cpp:83 */
      v = a.cc * c[k] + a.cx * (uw + ue) + a.cy * (s[k] + n[k]);
      if (a.forcing)
      {
        /* -2 pi sin^2x sin^2y sin t cos t - bx (cos^2x - sin^2x) sin^2y cos^2t
           - by (cos^2y - sin^2y) sin^2x cos^2t, left to right as the reference writes it
           (mpi_serial/diffusion.cpp:82-85) */
        const double b = (-2.0 * PI_) * tssx[k] * sy * a.stct - tdcx[k] * sy * a.c2t - dy * tssx[k] * a.c2t;
        v += b;
      }
    }
    r[k] = v;
  }
  st_row<W>(a.f + j * a.nx + x0, r);
}

/* the same row for the TMA sweep: the x-interior / y-boundary tests are loop invariants there, so
   they arrive as a bit mask (bit k = point k is an interior node) and the row is computed without
   branches; a masked-off point stores 0 exactly like stencil_store.  Same operation order. */
template <int W>
__device__ __forceinline__ void stencil_store_masked(const RhsArgs& a, double* out, const double (&s)[W],
                                                     const double (&c)[W], const double (&n)[W], double w, double e,
                                                     double sy, double dy, const double (&tdcx)[W],
                                                     const double (&tssx)[W], unsigned mask)
{
  double r[W];
#pragma unroll
  for (int k = 0; k < W; k++)
  {
    const double uw = (k == 0) ? w : c[k - 1];
    const double ue = (k == W - 1) ? e : c[k + 1];
    r[k]            = a.cc * c[k] + a.cx * (uw + ue) + a.cy * (s[k] + n[k]);
  }
  if (a.forcing)
  {
#pragma unroll
    for (int k = 0; k < W; k++)
    {
      const double b = (-2.0 * PI_) * tssx[k] * sy * a.stct - tdcx[k] * sy * a.c2t - dy * tssx[k] * a.c2t;
      r[k] += b;
    }
  }
#pragma unroll
  for (int k = 0; k < W; k++) r[k] = ((mask >> k) & 1u) ? r[k] : 0.0;
  st_row<W>(out, r);
}

/* register path: west / east neighbours by warp shuffle, warp edges prefetched */
template <int W>
__device__ __forceinline__ void compute_row(const RhsArgs& a, int64_t j, int64_t x0, bool act, const double (&s)[W],
                                            const Row<W>& c, const double (&n)[W], double sy, double dy,
                                            const double (&tdcx)[W], const double (&tssx)[W])
{
  const int lane = threadIdx.x & 31;
  double w = __shfl_up_sync(0xffffffffu, c.v[W - 1], 1);
  double e = __shfl_down_sync(0xffffffffu, c.v[0], 1);
  if (!act) return;
  if (lane == 0) w = c.edge;
  if (lane == 31) e = c.edge;
  stencil_store<W>(a, j, x0, s, c.v, n, w, e, sy, dy, tdcx, tssx);
}

constexpr int kMaxRows = 512; /* rows per CTA whose y factors are staged in shared memory */

/* LA = rows of look-ahead: the ring keeps rows j-1 .. j+LA of the thread's span in
   registers, i.e. LA x 32 B per thread are in flight while row j is computed. */
template <int W, int LA, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_diffusion_rhs(const __grid_constant__ RhsArgs a, int R)
{
  constexpr int NR = LA + 2;
  __shared__ double s_sy[kMaxRows], s_dy[kMaxRows];
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int lane   = threadIdx.x & 31;
  const int64_t x0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * W;
  const bool act   = x0 < a.nx;
  /* row blocks: the two that own the strip's first and last rows are scheduled
     FIRST so that their pushes leave at kernel start */
  const int nyb = gridDim.y;
  int yb        = blockIdx.y;
  if (nyb > 2) yb = (blockIdx.y == 0) ? 0 : (blockIdx.y == 1) ? nyb - 1 : blockIdx.y - 1;
  const int64_t jb = (int64_t)yb * R;
  const int64_t je = (jb + R < a.ny_loc) ? jb + R : a.ny_loc;
  const bool ownS  = (jb == 0) && a.hasS;
  const bool ownN  = (je == a.ny_loc) && a.hasN;
  const int64_t xo = act ? x0 : 0; /* inactive threads read a valid address, results unused */

  /* ---- 1. push the boundary rows to the neighbours */
  if (ownS || ownN)
  {
    double v[W];
    if (ownS && act)
    {
      ld_row<W>(a.u + xo, v);
      st_row<W>(a.sendS + xo, v);
    }
    if (ownN && act)
    {
      ld_row<W>(a.u + (a.ny_loc - 1) * a.nx + xo, v);
      st_row<W>(a.sendN + xo, v);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
    {
      if (ownS) atomicAdd_system(a.ctrS_remote, 1ull);
      if (ownN) atomicAdd_system(a.ctrN_remote, 1ull);
    }
  }

  double tdcx[W], tssx[W];
#pragma unroll
  for (int k = 0; k < W; k++) tdcx[k] = tssx[k] = 0.0;
  if (a.forcing)
  {
    if (act)
    {
      ld_row<W>(a.dcx + xo, tdcx);
      ld_row<W>(a.ssx + xo, tssx);
    }
    for (int64_t j = jb + threadIdx.x; j < je; j += kThreads)
    {
      s_sy[j - jb] = a.ssy[j];
      s_dy[j - jb] = a.dcy[j];
    }
  }
  else
    for (int j = threadIdx.x; j < kMaxRows; j += kThreads) s_sy[j] = s_dy[j] = 0.0;
  __syncthreads();

  /* ---- 2. march the rows that need no halo: [m0, m1) */
  const int64_t m0 = jb + (ownS ? 1 : 0);
  const int64_t m1 = je - (ownN ? 1 : 0);
  if (m0 < m1)
  {
    /* highest row any stencil of [m0, m1) touches (a global-boundary row needs no
       neighbours: its result is 0) */
    const int64_t top = (m1 < a.ny_loc) ? m1 : a.ny_loc - 1;
    Row<W> ring[NR];
#pragma unroll
    for (int s = 0; s < NR; s++)
    {
#pragma unroll
      for (int k = 0; k < W; k++) ring[s].v[k] = 0.0;
      ring[s].edge = 0.0;
    }
#pragma unroll
    for (int s = 0; s < NR; s++)
    {
      const int64_t row = m0 - 1 + s;
      if (row >= 0 && row <= top) load_row<W>(a, row, xo, x0, lane, ring[s]);
    }
    for (int64_t jj = m0; jj < m1; jj += NR)
    {
#pragma unroll
      for (int k = 0; k < NR; k++)
      {
        const int64_t j = jj + k;
        if (j < m1)
        {
          compute_row<W>(a, j, x0, act, ring[k].v, ring[(k + 1) % NR], ring[(k + 2) % NR].v, s_sy[j - jb],
                         s_dy[j - jb], tdcx, tssx);
          const int64_t nr = j + NR - 1; /* the row that replaces row j-1 in the ring */
          if (nr <= top) load_row<W>(a, nr, xo, x0, lane, ring[k]);
        }
      }
    }
  }

  /* ---- 3. the halo-dependent rows, last */
  if (ownN)
  {
    const int64_t j = a.ny_loc - 1;
    Row<W> s, c, n;
    s.edge = c.edge = n.edge = 0.0;
    load_row<W>(a, j - 1, xo, x0, lane, s);
    load_row<W>(a, j, xo, x0, lane, c);
    wait_counter(a.ctrN_local, a.expected);
    ld_halo<W>(a.recvN + xo, n.v);
    compute_row<W>(a, j, x0, act, s.v, c, n.v, s_sy[j - jb], s_dy[j - jb], tdcx, tssx);
  }
  if (ownS)
  {
    Row<W> s, c, n;
    s.edge = c.edge = n.edge = 0.0;
    load_row<W>(a, 0, xo, x0, lane, c);
    load_row<W>(a, 1, xo, x0, lane, n);
    wait_counter(a.ctrS_local, a.expected);
    ld_halo<W>(a.recvS + xo, s.v);
    compute_row<W>(a, 0, x0, act, s.v, c, n.v, s_sy[0], s_dy[0], tdcx, tssx);
  }
}

/* ------------------------------------------------------------------------------------
 * TMA path (nx % 4 == 0, 16-byte aligned u): the rows of the CTA's x-tile stream through a
 * ring of kStages shared-memory stages filled by 1-D bulk copies (cp.async.bulk, one
 * elected producer thread, mbarrier complete_tx), so the bytes in flight per SM are set by
 * shared memory (2 CTAs x 11 rows x 8 KB = 180 KB) and not by registers -- the register
 * march above tops out at ~64 KB in flight per SM and 4.2 TB/s (profiles/r01_rhs_*).
 * 8 consumer warps keep the south and centre rows in registers, read the new north row and
 * the two x-neighbours of their span from the ring, and release a stage (empty mbarrier)
 * once it has served as centre row.  The halo-dependent first / last row of the strip is
 * finished afterwards by the register path, exactly as in k_diffusion_rhs.
 * ------------------------------------------------------------------------------------ */
constexpr int kTX        = kThreads * 4; /* x-points per CTA tile */
constexpr int kStagesMax = 12; /* the mbarrier block at the head of shared memory is sized for this */
constexpr int kRowDbl    = kTX + 4; /* a stage holds x0-2 .. x0+kTX+1 (16-byte aligned both ends) */
constexpr int kTmaRows   = 256;     /* max rows per CTA (y factors staged in shared memory) */
constexpr int kTmaThreads = kThreads + 32;
constexpr size_t tma_smem(int stages) { return 256 + (size_t)stages * kRowDbl * sizeof(double); }

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
  uint32_t ok          = 0, spins = 0;
  unsigned long long t0 = 0;
  for (;;)
  {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (ok) return;
    if ((++spins & 0x3ffu) == 0)
    {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20000000000ull) asm volatile("trap;"); /* a lost copy must not hang the GPU */
    }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void lds4(const double* p, double (&v)[4])
{
  const uint32_t a = smem_addr(p);
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(a) : "memory");
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[2]), "=d"(v[3]) : "r"(a + 16) : "memory");
}

template <int kStages, int MINB>
__global__ void __launch_bounds__(kTmaThreads, MINB) k_diffusion_rhs_tma(const __grid_constant__ RhsArgs a, int R)
{
  static_assert(kStages <= kStagesMax, "mbarrier block");
  constexpr int W = 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double s_sy[kTmaRows], s_dy[kTmaRows];
  const uint32_t bar_full  = smem_addr(smem_raw);                 /* kStages x 8 B */
  const uint32_t bar_empty = bar_full + 8 * kStagesMax;           /* kStages x 8 B */
  double* const ring       = (double*)(smem_raw + 256);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int tid       = threadIdx.x;
  const int lane      = tid & 31;
  const bool consumer = tid < kThreads;
  const int64_t xt    = (int64_t)blockIdx.x * kTX; /* first x of the tile */
  const int64_t x0    = consumer ? xt + (int64_t)tid * W : a.nx;
  const bool act      = x0 < a.nx;
  const int nyb       = gridDim.y;
  int yb              = blockIdx.y;
  if (nyb > 2) yb = (blockIdx.y == 0) ? 0 : (blockIdx.y == 1) ? nyb - 1 : blockIdx.y - 1;
  const int64_t jb = (int64_t)yb * R;
  const int64_t je = (jb + R < a.ny_loc) ? jb + R : a.ny_loc;
  const bool ownS  = (jb == 0) && a.hasS;
  const bool ownN  = (je == a.ny_loc) && a.hasN;
  const int64_t xo = act ? x0 : 0;

  if (tid == 0)
  {
    for (int s = 0; s < kStages; s++)
    {
      mbar_init(bar_full + 8 * s, 1);               /* the producer's expect_tx arrive */
      mbar_init(bar_empty + 8 * s, kThreads / 32);  /* one arrive per consumer warp    */
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }

  /* ---- 1. push the boundary rows to the neighbours */
  if (ownS || ownN)
  {
    double v[W];
    if (ownS && act)
    {
      ld_row<W>(a.u + xo, v);
      st_row<W>(a.sendS + xo, v);
    }
    if (ownN && act)
    {
      ld_row<W>(a.u + (a.ny_loc - 1) * a.nx + xo, v);
      st_row<W>(a.sendN + xo, v);
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0)
    {
      if (ownS) atomicAdd_system(a.ctrS_remote, 1ull);
      if (ownN) atomicAdd_system(a.ctrN_remote, 1ull);
    }
  }

  double tdcx[W], tssx[W];
#pragma unroll
  for (int k = 0; k < W; k++) tdcx[k] = tssx[k] = 0.0;
  if (a.forcing)
  {
    if (act)
    {
      ld_row<W>(a.dcx + xo, tdcx);
      ld_row<W>(a.ssx + xo, tssx);
    }
    for (int64_t j = jb + tid; j < je; j += kTmaThreads)
    {
      s_sy[j - jb] = a.ssy[j];
      s_dy[j - jb] = a.dcy[j];
    }
  }
  else
    for (int j = tid; j < kTmaRows; j += kTmaThreads) s_sy[j] = s_dy[j] = 0.0;
  __syncthreads(); /* barriers initialised, y factors staged */

  /* ---- 2. the rows that need no halo: [m0, m1); ring rows lo .. hi */
  const int64_t m0 = jb + (ownS ? 1 : 0);
  const int64_t m1 = je - (ownN ? 1 : 0);
  if (m0 < m1)
  {
    const int64_t lo = (m0 > 0) ? m0 - 1 : 0;
    const int64_t hi = (m1 < a.ny_loc) ? m1 : a.ny_loc - 1;
    if (!consumer)
    {
      if (lane == 0)
      {
        /* the tile's span of a row, clipped to the mesh: [xs, xe) */
        const int64_t xs     = (xt >= 2) ? xt - 2 : 0;
        const int64_t xe     = (xt + kTX + 2 <= a.nx) ? xt + kTX + 2 : a.nx;
        const uint32_t bytes = (uint32_t)((xe - xs) * sizeof(double));
        const uint32_t doff  = (uint32_t)((xs - (xt - 2)) * sizeof(double));
        asm volatile("fence.proxy.async;" ::: "memory");
        int st = 0;             /* stage of row r and the parity of its previous use */
        uint32_t par = 1;       /* flips each time the ring wraps; first pass: nothing to wait for */
        bool wrapped = false;
        const double* src = a.u + lo * a.nx + xs;
        for (int64_t r = lo; r <= hi; r++, src += a.nx)
        {
          if (wrapped) mbar_wait(bar_empty + 8 * st, par);
          mbar_expect_tx(bar_full + 8 * st, bytes);
          bulk_g2s(smem_addr(ring + (size_t)st * kRowDbl) + doff, src, bytes, bar_full + 8 * st);
          if (++st == kStages)
          {
            st      = 0;
            par ^= 1u;
            wrapped = true;
          }
        }
      }
    }
    else
    {
      const int px = 2 + tid * W; /* this thread's first point inside a stage */
      double s[W], c[W], n[W];
#pragma unroll
      for (int k = 0; k < W; k++) s[k] = c[k] = n[k] = 0.0;
      int st = 0;       /* stage and parity of the next row to take from the ring */
      uint32_t par = 0;
      auto next_row = [&]() -> const double* {
        mbar_wait(bar_full + 8 * st, par);
        const double* p = ring + (size_t)st * kRowDbl;
        if (++st == kStages)
        {
          st = 0;
          par ^= 1u;
        }
        return p;
      };
      int rel = 0; /* stage of the oldest row not yet released */
      auto release = [&]() {
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * rel);
        if (++rel == kStages) rel = 0;
      };
      if (m0 > lo) /* south of the first row */
      {
        lds4(next_row() + px, s);
        release();
      }
      const double* pc = next_row();
      lds4(pc + px, c);
      /* loop invariants of the row sweep: which of the thread's W points are x-interior, the two
         (CTA-relative) rows that lie on the global y boundary, and a running output pointer */
      unsigned inmask = 0;
#pragma unroll
      for (int k = 0; k < W; k++)
        if (act && x0 + k > 0 && x0 + k < a.nx - 1) inmask |= 1u << k;
      const int64_t g0 = a.js + jb; /* global index of the CTA's first row */
      const int yb0    = (g0 <= 0 && -g0 < kTmaRows) ? (int)(-g0) : -1;
      const int yb1    = (a.ny - 1 - g0 >= 0 && a.ny - 1 - g0 < kTmaRows) ? (int)(a.ny - 1 - g0) : -1;
      int jr           = (int)(m0 - jb);
      const int jr_end = (int)(m1 - jb), jr_hi = (int)(hi - jb);
      double* fp       = a.f + m0 * a.nx + xo;
      /* one row with the roles of the three register rows given by the caller, so that the sweep
         below rotates names instead of moving 8 doubles per row */
      auto row = [&](const double(&S)[W], const double(&C)[W], double(&N)[W]) {
        const double* pn = pc;
        if (jr < jr_hi)
        {
          pn = next_row();
          lds4(pn + px, N);
        }
        if (act)
        {
          const double w = pc[px - 1], e = pc[px + W];
          const unsigned m = (jr == yb0 || jr == yb1) ? 0u : inmask;
          stencil_store_masked<W>(a, fp, S, C, N, w, e, s_sy[jr], s_dy[jr], tdcx, tssx, m);
        }
        release(); /* the row served as centre: free its stage */
        fp += a.nx;
        pc = pn;
        jr++;
      };
      while (jr + 3 <= jr_end)
      {
        row(s, c, n);
        row(c, n, s);
        row(n, s, c);
      }
      if (jr < jr_end)
      {
        row(s, c, n);
        if (jr < jr_end) row(c, n, s);
      }
    }
  }

  /* ---- 3. the halo-dependent rows, last (register path) */
  if (ownN)
  {
    const int64_t j = a.ny_loc - 1;
    Row<W> s, c, n;
    s.edge = c.edge = n.edge = 0.0;
    load_row<W>(a, j - 1, xo, x0, lane, s);
    load_row<W>(a, j, xo, x0, lane, c);
    wait_counter(a.ctrN_local, a.expected);
    ld_halo<W>(a.recvN + xo, n.v);
    compute_row<W>(a, j, x0, act, s.v, c, n.v, s_sy[j - jb], s_dy[j - jb], tdcx, tssx);
  }
  if (ownS)
  {
    Row<W> s, c, n;
    s.edge = c.edge = n.edge = 0.0;
    load_row<W>(a, 0, xo, x0, lane, c);
    load_row<W>(a, 1, xo, x0, lane, n);
    wait_counter(a.ctrS_local, a.expected);
    ld_halo<W>(a.recvS + xo, s.v);
    compute_row<W>(a, 0, x0, act, s.v, c, n.v, s_sy[0], s_dy[0], tdcx, tssx);
  }
}

/* u = sin^2(pi x) sin^2(pi y) cos^2(pi t) inside, 0 on the boundary
   (mpi_serial/solution.cpp:25-62) */
__global__ void __launch_bounds__(kThreads) k_solution(double* u, int64_t nx, int64_t ny, int64_t ny_loc, int64_t js,
                                                       const double* ssx, const double* ssy, double c2t)
{
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int64_t total = nx * ny_loc;
  for (int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x; c < total; c += (int64_t)gridDim.x * kThreads)
  {
    const int64_t j = c / nx, i = c - j * nx, jg = js + j;
    const bool inside = (i > 0 && i < nx - 1 && jg > 0 && jg < ny - 1);
    u[c]              = inside ? ssx[i] * ssy[j] * c2t : 0.0;
  }
}

} // namespace

namespace {
/* look-ahead / residency variants of the RHS kernel (B200_DIFFUSION_VARIANT picks one; tuning) */
typedef void (*rhs_kernel_t)(const RhsArgs, int);
struct RhsVariant
{
  const char* name;
  rhs_kernel_t k4, k1;
};
const RhsVariant kVariants[] = {
  {"la4x2", k_diffusion_rhs<4, 4, 2>, k_diffusion_rhs<1, 4, 2>}, /* default */
  {"la2x4", k_diffusion_rhs<4, 2, 4>, k_diffusion_rhs<1, 2, 4>},
  {"la3x3", k_diffusion_rhs<4, 3, 3>, k_diffusion_rhs<1, 3, 3>},
  {"la6x2", k_diffusion_rhs<4, 6, 2>, k_diffusion_rhs<1, 6, 2>},
  {"la8x1", k_diffusion_rhs<4, 8, 1>, k_diffusion_rhs<1, 8, 1>},
};
} // namespace

struct b200_diffusion2d_plan_s
{
  rhs_kernel_t kernel = nullptr; /* register-march kernel (any nx, any alignment) */
  bool use_tma        = false;   /* TMA ring kernel when nx % 4 == 0 and u is 16-byte aligned */
  rhs_kernel_t tma_kernel = nullptr; /* the ring instantiation in use and its dynamic shared memory */
  size_t tma_smem_bytes   = 0;
  b200vec_ctx ctx     = nullptr;
  b200_diffusion2d_opts o;
  int rank = 0, np = 1;
  int64_t nx = 0, ny = 0, ny_loc = 0, js = 0, nodes = 0, nodes_loc = 0;
  double dx = 0, dy = 0;
  int hasS = 0, hasN = 0;
  double* d_tables = nullptr; /* dcx | ssx (nx each) | ssy | dcy (ny_loc each) */
  /* symmetric halo region: [16 u64 counters][Srecv 2 x nx][Nrecv 2 x nx] */
  void* peers[8]          = {nullptr};
  bool have_peers         = false;
  unsigned long long seq  = 0;
  int tiles_x             = 1;
  int W                   = 4;
  N_Vector diag           = nullptr; /* Jacobi */
  /* timing */
  bool time_rhs = false;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  double rhs_ms = 0;
  long rhs_calls = 0;
};

namespace {

inline unsigned long long* ctr_of(void* base, int which) { return (unsigned long long*)base + 8 * which; } /* 0 fromS, 1 fromN */
inline double* srecv_of(void* base, int64_t nx, int par) { return (double*)((unsigned long long*)base + 16) + (size_t)par * nx; }
inline double* nrecv_of(void* base, int64_t nx, int par) { return (double*)((unsigned long long*)base + 16) + (size_t)(2 + par) * nx; }

int fail(const char* what)
{
  fprintf(stderr, "[diffusion2d_b200] ERROR: %s (%s)\n", what, b200vec_last_error());
  return -1;
}

} // namespace

extern "C" {

void b200_diffusion2d_default_opts(b200_diffusion2d_opts* o)
{
  memset(o, 0, sizeof(*o));
  o->nx = o->ny = 32;
  o->xu = o->yu = 1.0;
  o->kx = o->ky = 1.0;
  o->tf         = 1.0;
  o->forcing    = 1;
  o->rtol       = 1e-5;
  o->atol       = 1e-10;
  o->order      = 3;
  o->linear     = 1;
  o->ls_gmres   = 0;
  o->preconditioning = 1;
  o->liniters   = 20;
  o->msbp       = 0;
  o->epslin     = 0.0;
  o->maxsteps   = 0;
  strcpy(o->controller, "I");
  o->output       = 1;
  o->nout         = 20;
  o->fused_ops    = 1;
  o->rows_per_cta = 0;
  o->fused_ewt    = 1;
}

int b200_diffusion2d_plan_create(b200vec_ctx ctx, const b200_diffusion2d_opts* opts, b200_diffusion2d_plan* out)
{
  if (!ctx || !opts || !out) return -1;
  auto* p = new b200_diffusion2d_plan_s();
  p->ctx  = ctx;
  p->o    = *opts;
  p->rank = b200vec_comm_rank(ctx);
  p->np   = b200vec_comm_size(ctx);
  p->nx   = opts->nx;
  p->ny   = opts->ny;
  /* y-extents of this strip: UserData::setup, diffusion_2D.cpp:323-329 with npx = 1, npy = np */
  const int64_t qy = p->ny / p->np, ry = p->ny % p->np;
  p->js            = qy * p->rank + (p->rank < ry ? p->rank : ry);
  p->ny_loc        = qy + (p->rank < ry ? 1 : 0);
  p->nodes         = p->nx * p->ny;
  p->nodes_loc     = p->nx * p->ny_loc;
  p->dx            = opts->xu / (double)(p->nx - 1);
  p->dy            = opts->yu / (double)(p->ny - 1);
  p->hasS          = (p->js != 0);
  p->hasN          = (p->js + p->ny_loc != p->ny);
  if (p->np > 1 && p->ny_loc < 2)
  {
    delete p;
    return fail("every rank needs at least 2 mesh rows");
  }
  p->W       = (p->nx % 4 == 0) ? 4 : 1;
  p->tiles_x = (int)((p->nx + (int64_t)kThreads * p->W - 1) / ((int64_t)kThreads * p->W));
  /* rows per CTA: by default ONE wave of equal row blocks -- (resident CTAs per SM x SMs)
     CTAs, every CTA marching the same number of rows, so there is no tail wave and only
     2 halo rows are re-read per R rows; rows_per_cta > 0 overrides (tuning) */
  {
    int dev = 0, sms = 148, occ = 2;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const RhsVariant* var = &kVariants[0];
    if (const char* want = getenv("B200_DIFFUSION_VARIANT"))
      for (const RhsVariant& v : kVariants)
        if (!strcmp(v.name, want)) var = &v;
    p->kernel = (p->W == 4) ? var->k4 : var->k1;
    const char* want = getenv("B200_DIFFUSION_VARIANT");
    p->use_tma       = (p->W == 4) && (!want || !strcmp(want, "tma"));
    if (p->use_tma)
    {
      /* ring geometry: stages x resident CTAs per SM (B200_DIFFUSION_TMA = "12x2" | "8x3" | "6x3"; tuning) */
      struct { const char* name; rhs_kernel_t k; int stages; } rings[] = {
        {"12x2", k_diffusion_rhs_tma<12, 2>, 12}, {"8x3", k_diffusion_rhs_tma<8, 3>, 8}, {"6x3", k_diffusion_rhs_tma<6, 3>, 6},
        {"8x2", k_diffusion_rhs_tma<8, 2>, 8}};
      int pick = 0;
      if (const char* r = getenv("B200_DIFFUSION_TMA"))
        for (int i = 0; i < 4; i++)
          if (!strcmp(rings[i].name, r)) pick = i;
      p->tma_kernel     = rings[pick].k;
      p->tma_smem_bytes = tma_smem(rings[pick].stages);
      if (cudaFuncSetAttribute(p->tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->tma_smem_bytes) !=
          cudaSuccess)
      {
        cudaGetLastError();
        p->use_tma = false;
      }
    }
    if (p->use_tma) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, p->tma_kernel, kTmaThreads, p->tma_smem_bytes);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, p->kernel, kThreads, 0);
    if (occ < 1) occ = 1;
    int64_t R = p->o.rows_per_cta;
    if (R <= 0)
    {
      int64_t blocks_y = ((int64_t)sms * occ) / p->tiles_x;
      if (blocks_y < 1) blocks_y = 1;
      R = (p->ny_loc + blocks_y - 1) / blocks_y;
    }
    if (R < 2) R = 2;
    if (R > (p->use_tma ? kTmaRows : kMaxRows)) R = p->use_tma ? kTmaRows : kMaxRows;
    p->o.rows_per_cta = (int)R;
  }

  /* factor tables with the HOST libm: the same factors the reference's CPU
     code evaluates per point (mpi_serial/diffusion.cpp:60-90) */
  const int64_t nx = p->nx, nyl = p->ny_loc;
  std::vector<double> t(2 * nx + 2 * nyl);
  const double bx = opts->kx * 2.0 * PI_ * PI_, by = opts->ky * 2.0 * PI_ * PI_;
  for (int64_t i = 0; i < nx; i++)
  {
    const double x   = (double)i * p->dx;
    const double ssx = sin(PI_ * x) * sin(PI_ * x), csx = cos(PI_ * x) * cos(PI_ * x);
    t[i]             = bx * (csx - ssx);
    t[nx + i]        = ssx;
  }
  for (int64_t j = 0; j < nyl; j++)
  {
    const double y   = (double)(p->js + j) * p->dy;
    const double ssy = sin(PI_ * y) * sin(PI_ * y), csy = cos(PI_ * y) * cos(PI_ * y);
    t[2 * nx + j]       = ssy;
    t[2 * nx + nyl + j] = by * (csy - ssy);
  }
  void* d = nullptr;
  if (b200vec_malloc_device(ctx, t.size() * sizeof(double), &d)) { delete p; return fail("table allocation"); }
  p->d_tables = (double*)d;
  if (b200vec_copy_h2d(ctx, d, t.data(), t.size() * sizeof(double), 1)) { delete p; return fail("table upload"); }

  if (p->np > 1)
  {
    const size_t bytes = 16 * sizeof(unsigned long long) + (size_t)4 * nx * sizeof(double);
    if (b200vec_comm_peer_alloc(ctx, bytes, p->peers)) { delete p; return fail("peer allocation for the halo buffers"); }
    p->have_peers = true;
  }
  p->time_rhs = getenv("B200_DIFFUSION_TIME_RHS") != nullptr;
  if (p->time_rhs)
  {
    cudaEventCreate(&p->e0);
    cudaEventCreate(&p->e1);
  }
  *out = p;
  return 0;
}

int64_t b200_diffusion2d_plan_local_nodes(b200_diffusion2d_plan p) { return p ? p->nodes_loc : -1; }

void b200_diffusion2d_plan_destroy(b200_diffusion2d_plan p)
{
  if (!p) return;
  b200vec_ctx_sync(p->ctx);
  if (p->have_peers) b200vec_comm_peer_free(p->ctx, p->peers);
  if (p->d_tables) b200vec_free_device(p->ctx, p->d_tables, (size_t)(2 * p->nx + 2 * p->ny_loc) * sizeof(double));
  if (p->e0) cudaEventDestroy(p->e0);
  if (p->e1) cudaEventDestroy(p->e1);
  delete p;
}

int b200_diffusion2d_rhs(b200_diffusion2d_plan p, double t, const double* u, double* f)
{
  RhsArgs a;
  memset(&a, 0, sizeof(a));
  a.u = u;
  a.f = f;
  a.nx = p->nx; a.ny = p->ny; a.ny_loc = p->ny_loc; a.js = p->js;
  /* mpi_gpu/diffusion.cpp:77-79 */
  a.cx = p->o.kx / (p->dx * p->dx);
  a.cy = p->o.ky / (p->dy * p->dy);
  a.cc = -2.0 * (a.cx + a.cy);
  a.stct    = sin(PI_ * t) * cos(PI_ * t);
  a.c2t     = cos(PI_ * t) * cos(PI_ * t);
  a.forcing = p->o.forcing;
  a.dcx = p->d_tables;
  a.ssx = p->d_tables + p->nx;
  a.ssy = p->d_tables + 2 * p->nx;
  a.dcy = p->d_tables + 2 * p->nx + p->ny_loc;
  a.hasS = p->hasS;
  a.hasN = p->hasN;
  const unsigned long long seq = ++p->seq;
  const int par                = (int)(seq & 1ull);
  if (p->np > 1)
  {
    void* me = p->peers[p->rank];
    if (p->hasS)
    {
      void* nb      = p->peers[p->rank - 1];
      a.sendS       = nrecv_of(nb, p->nx, par); /* my first row is its "row from the north" */
      a.ctrS_remote = ctr_of(nb, 1);
      a.recvS       = srecv_of(me, p->nx, par);
      a.ctrS_local  = ctr_of(me, 0);
    }
    if (p->hasN)
    {
      void* nb      = p->peers[p->rank + 1];
      a.sendN       = srecv_of(nb, p->nx, par);
      a.ctrN_remote = ctr_of(nb, 0);
      a.recvN       = nrecv_of(me, p->nx, par);
      a.ctrN_local  = ctr_of(me, 1);
    }
    a.expected = seq * (unsigned long long)p->tiles_x;
  }
  const int R = p->o.rows_per_cta;
  dim3 grid((unsigned)p->tiles_x, (unsigned)((p->ny_loc + R - 1) / R));
  cudaStream_t s = (cudaStream_t)b200vec_ctx_get_stream(p->ctx);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim            = grid;
  cfg.blockDim           = dim3(kThreads);
  cfg.stream             = s;
  cudaLaunchAttribute at[1];
  at[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs                                        = at;
  cfg.numAttrs                                     = 1;
  if (p->time_rhs) cudaEventRecord(p->e0, s);
  cudaError_t e;
  if (p->use_tma && ((uintptr_t)u % 16 == 0) && ((uintptr_t)f % 32 == 0))
  {
    cfg.blockDim         = dim3(kTmaThreads);
    cfg.dynamicSmemBytes = p->tma_smem_bytes;
    e                    = cudaLaunchKernelEx(&cfg, p->tma_kernel, a, R);
  }
  else e = cudaLaunchKernelEx(&cfg, p->kernel, a, R);
  if (e != cudaSuccess)
  {
    fprintf(stderr, "[diffusion2d_b200] RHS launch failed: %s\n", cudaGetErrorString(e));
    return -1;
  }
  if (p->time_rhs)
  {
    cudaEventRecord(p->e1, s);
    cudaEventSynchronize(p->e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, p->e0, p->e1);
    p->rhs_ms += ms;
  }
  p->rhs_calls++;
  return 0;
}

int b200_diffusion2d_solution(b200_diffusion2d_plan p, double t, double* u)
{
  const double c2t = cos(PI_ * t) * cos(PI_ * t);
  int64_t blocks   = (p->nodes_loc + kThreads - 1) / kThreads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  cudaStream_t s = (cudaStream_t)b200vec_ctx_get_stream(p->ctx);
  k_solution<<<(unsigned)blocks, kThreads, 0, s>>>(u, p->nx, p->ny, p->ny_loc, p->js, p->d_tables + p->nx,
                                                   p->d_tables + 2 * p->nx, c2t);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

} /* extern "C" */

/* ---- callbacks handed to the (unmodified) integrator ------------------------ */
namespace {

int rhs_cb(sunrealtype t, N_Vector u, N_Vector f, void* user_data)
{
  auto* p = (b200_diffusion2d_plan)user_data;
  return b200_diffusion2d_rhs(p, t, N_VGetDeviceArrayPointer_B200(u), N_VGetDeviceArrayPointer_B200(f));
}

/* preconditioner_jacobi.cpp:25-61 */
/* ARKEwtFn: ewt = 1 / (rtol |y| + atol) in one kernel, the bits of arkEwtSetSS (arkode.c:2935-2947) */
int ewt_cb(N_Vector y, N_Vector ewt, void* user_data)
{
  auto* p = (b200_diffusion2d_plan)user_data;
  return N_VEwtSet_B200(p->o.rtol, p->o.atol, nullptr, p->o.atol == 0.0, y, ewt);
}

int psetup_cb(sunrealtype, N_Vector, N_Vector, sunbooleantype, sunbooleantype*, sunrealtype gamma, void* user_data)
{
  auto* p         = (b200_diffusion2d_plan)user_data;
  const double cx = p->o.kx / (p->dx * p->dx), cy = p->o.ky / (p->dy * p->dy);
  const double cc = -2.0 * (cx + cy);
  N_VConst(1.0 / (1.0 - gamma * cc), p->diag);
  return 0;
}
int psolve_cb(sunrealtype, N_Vector, N_Vector, N_Vector r, N_Vector z, sunrealtype, sunrealtype, int, void* user_data)
{
  auto* p = (b200_diffusion2d_plan)user_data;
  N_VProd(p->diag, r, z);
  return 0;
}

/* UserOutput::write, diffusion_2D.cpp:778-846 */
void write_row(b200_diffusion2d_plan p, double t, N_Vector u, N_Vector err, bool table, double* urms_out, double* max_out)
{
  double mx = 0.0;
  if (err)
  {
    b200_diffusion2d_solution(p, t, N_VGetDeviceArrayPointer_B200(err));
    N_VLinearSum(1.0, u, -1.0, err, err);
    N_VAbs(err, err);
    mx = N_VMaxNorm(err);
  }
  const double urms = sqrt(N_VDotProd(u, u) / (double)p->nx / (double)p->ny);
  if (table && p->rank == 0)
  {
    if (err) printf("%22.15e%25.15e%25.15e\n", t, urms, mx);
    else printf("%22.15e%25.15e\n", t, urms);
  }
  *urms_out = urms;
  *max_out  = mx;
}

#define CHK(call, what)                                                        \
  do {                                                                         \
    int flag_ = (call);                                                        \
    if (flag_ < 0)                                                             \
    {                                                                          \
      fprintf(stderr, "[diffusion2d_b200] %s failed with flag %d\n", what, flag_); \
      return -1;                                                               \
    }                                                                          \
  } while (0)

double now_s()
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

extern "C" int b200_diffusion2d_run(b200vec_ctx ctx, const b200_diffusion2d_opts* opts, b200_diffusion2d_stats* st)
{
  if (!ctx || !opts || !st) return -1;
  memset(st, 0, sizeof(*st));
  const double t_setup0 = now_s();
  b200_diffusion2d_plan p = nullptr;
  if (b200_diffusion2d_plan_create(ctx, opts, &p)) return -1;
  const bool table = opts->output > 0;

  SUNContext sunctx = nullptr;
  CHK(SUNContext_Create(SUN_COMM_NULL, &sunctx), "SUNContext_Create");

  N_Vector u = N_VNewWithCtx_B200(p->nodes_loc, B200_MEM_DEVICE, ctx, sunctx);
  if (!u) return fail("N_VNewWithCtx_B200");
  if (p->np > 1 && N_VMakeDistributed_B200(u, p->nodes)) return fail("N_VMakeDistributed_B200");
  if (opts->fused_ops) N_VEnableFusedOps_B200(u, SUNTRUE);

  /* initial condition; error vector only with forcing (main_arkode.cpp:171-180) */
  b200_diffusion2d_solution(p, 0.0, N_VGetDeviceArrayPointer_B200(u));
  N_Vector err = opts->forcing ? N_VClone(u) : nullptr;

  const int prectype = opts->preconditioning ? SUN_PREC_RIGHT : SUN_PREC_NONE; /* main_arkode.cpp:207 */
  SUNLinearSolver LS = opts->ls_gmres ? SUNLinSol_SPGMR(u, prectype, opts->liniters, sunctx)
                                      : SUNLinSol_PCG(u, prectype, opts->liniters, sunctx);
  if (!LS) return fail("SUNLinSol constructor");
  if (opts->preconditioning) p->diag = N_VClone(u);

  void* mem = ARKStepCreate(nullptr, rhs_cb, 0.0, u, sunctx);
  if (!mem) return fail("ARKStepCreate");
  CHK(ARKodeSStolerances(mem, opts->rtol, opts->atol), "ARKodeSStolerances");
  CHK(ARKodeSetUserData(mem, p), "ARKodeSetUserData");
  /* DIRK (implicit): the built-in error-weight routine would be arkEwtSetSS; the registered function
     computes the same vector with one kernel instead of five */
  if (opts->fused_ewt) CHK(ARKodeWFtolerances(mem, ewt_cb), "ARKodeWFtolerances");
  CHK(ARKodeSetLinearSolver(mem, LS, nullptr), "ARKodeSetLinearSolver");
  if (opts->preconditioning)
  {
    CHK(ARKodeSetPreconditioner(mem, psetup_cb, psolve_cb), "ARKodeSetPreconditioner");
    CHK(ARKodeSetLSetupFrequency(mem, opts->msbp), "ARKodeSetLSetupFrequency");
  }
  CHK(ARKodeSetEpsLin(mem, opts->epslin), "ARKodeSetEpsLin");
  CHK(ARKodeSetOrder(mem, opts->order), "ARKodeSetOrder");
  if (opts->linear) CHK(ARKodeSetLinear(mem, 0), "ARKodeSetLinear");
  CHK(ARKodeSetAdaptControllerByName(mem, opts->controller), "ARKodeSetAdaptControllerByName");
  CHK(ARKodeSetMaxNumSteps(mem, opts->maxsteps), "ARKodeSetMaxNumSteps");
  CHK(ARKodeSetStopTime(mem, opts->tf), "ARKodeSetStopTime");

  if (table && p->rank == 0)
  {
    printf("\n");
    if (err)
    {
      printf("          t                     ||u||_rms                max error      \n");
      printf(" -----------------------------------------------------------------------\n");
    }
    else
    {
      printf("          t                     ||u||_rms      \n");
      printf(" ----------------------------------------------\n");
    }
  }
  double t = 0.0, urms = 0.0, mx = 0.0;
  write_row(p, t, u, err, table, &urms, &mx);
  b200vec_ctx_sync(ctx);
  st->setup_seconds = now_s() - t_setup0;

  const double dTout = opts->tf / opts->nout;
  double tout        = dTout;
  double evolve      = 0.0;
  int rc             = 0;
  for (int iout = 0; iout < opts->nout; iout++)
  {
    const double t0 = now_s();
    const int flag  = ARKodeEvolve(mem, tout, u, &t, ARK_NORMAL);
    b200vec_ctx_sync(ctx);
    evolve += now_s() - t0;
    if (flag < 0)
    {
      fprintf(stderr, "[diffusion2d_b200] ARKodeEvolve failed with flag %d\n", flag);
      rc = -1;
      break;
    }
    write_row(p, t, u, err, table, &urms, &mx);
    tout += dTout;
    tout = (tout > opts->tf) ? opts->tf : tout;
  }
  if (table && p->rank == 0)
  {
    printf(err ? " -----------------------------------------------------------------------\n\n"
               : " ----------------------------------------------\n\n");
    printf("Final integrator statistics:\n");
    ARKodePrintAllStats(mem, stdout, SUN_OUTPUTFORMAT_TABLE);
    fflush(stdout);
  }
  ARKodeGetNumSteps(mem, &st->nst);
  ARKodeGetNumStepAttempts(mem, &st->nst_a);
  ARKodeGetNumErrTestFails(mem, &st->netf);
  ARKodeGetNumRhsEvals(mem, 0, &st->nfe);
  ARKodeGetNumRhsEvals(mem, 1, &st->nfi);
  ARKodeGetNumNonlinSolvIters(mem, &st->nni);
  ARKodeGetNumNonlinSolvConvFails(mem, &st->ncfn);
  ARKodeGetNumLinSolvSetups(mem, &st->nsetups);
  ARKodeGetNumLinIters(mem, &st->nli);
  ARKodeGetNumLinConvFails(mem, &st->nlcf);
  if (opts->preconditioning)
  {
    ARKodeGetNumPrecEvals(mem, &st->npe);
    ARKodeGetNumPrecSolves(mem, &st->nps);
  }
  ARKodeGetNumJtimesEvals(mem, &st->njv);
  ARKodeGetNumLinRhsEvals(mem, &st->nfeLS);
  st->t_final        = t;
  st->urms           = urms;
  st->max_err        = mx;
  st->evolve_seconds = evolve;
  st->rhs_seconds    = p->rhs_ms * 1e-3;
  st->rhs_calls      = p->rhs_calls;
  st->nodes          = p->nodes;
  st->nodes_loc      = p->nodes_loc;
  st->nranks         = p->np;

  ARKodeFree(&mem);
  SUNLinSolFree(LS);
  if (p->diag) N_VDestroy(p->diag);
  if (err) N_VDestroy(err);
  N_VDestroy(u);
  b200_diffusion2d_plan_destroy(p);
  SUNContext_Free(&sunctx);
  return rc;
}
