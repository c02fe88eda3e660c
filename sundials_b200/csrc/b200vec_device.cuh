/* b200vec_device.cuh -- device-side building blocks for the sm_100a kernels.
 *
 *  - wide global loads/stores: 256-bit (LDG.E.256 / STG.E.256, new on sm_100),
 *    128-bit and 64-bit, all with L1::no_allocate (streaming data has no reuse
 *    inside a kernel; default L2 policy is kept on purpose -- inside an integrator
 *    the output of one op is the input of the next and the 126 MB L2 holds it).
 *  - deterministic warp / block combiners for the two-stage reductions.
 */
#ifndef B200VEC_DEVICE_CUH
#define B200VEC_DEVICE_CUH

#include <cfloat>
#include <cstdint>

#include "b200vec_internal.h"

namespace b200 {

/* ---- W-wide loads and stores (W doubles = 8W bytes, address 8W-aligned) ---- */
template <int W>
__device__ __forceinline__ void ldg(const double* p, double (&v)[W]);
template <int W>
__device__ __forceinline__ void stg(double* p, const double (&v)[W]);

template <>
__device__ __forceinline__ void ldg<1>(const double* p, double (&v)[1])
{
  asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v[0]) : "l"(p) : "memory");
}
template <>
__device__ __forceinline__ void ldg<2>(const double* p, double (&v)[2])
{
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
               : "=d"(v[0]), "=d"(v[1])
               : "l"(p)
               : "memory");
}
template <>
__device__ __forceinline__ void ldg<4>(const double* p, double (&v)[4])
{
  asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
               : "l"(p)
               : "memory");
}
template <>
__device__ __forceinline__ void stg<1>(double* p, const double (&v)[1])
{
  asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(v[0]) : "memory");
}
template <>
__device__ __forceinline__ void stg<2>(double* p, const double (&v)[2])
{
  asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v[0]), "d"(v[1]) : "memory");
}
template <>
__device__ __forceinline__ void stg<4>(double* p, const double (&v)[4])
{
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]),
               "d"(v[3])
               : "memory");
}

/* ---- programmatic dependent launch: first statement(s) of every kernel.
 * wait   = all prerequisite grids on the stream have completed and flushed;
 * launch = the next kernel on the stream may start becoming resident. ---- */
__device__ __forceinline__ void pdl_prologue()
{
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

/* ---- combiners.  The predicates mirror nvector_serial.c:635,719 (strict
 * comparisons, so NaNs never win) rather than fmax/fmin. ---- */
struct CombSum
{
  static constexpr int op = B200VEC_SUM;
  static __device__ __forceinline__ double identity() { return 0.0; }
  static __device__ __forceinline__ double apply(double a, double b) { return a + b; }
  static constexpr bool order_sensitive = true;
};
struct CombMax
{
  static constexpr int op = B200VEC_MAX;
  static __device__ __forceinline__ double identity() { return 0.0; } /* max |x| starts at 0 (serial:627) */
  static __device__ __forceinline__ double apply(double a, double b) { return (b > a) ? b : a; }
  static constexpr bool order_sensitive = false;
};
struct CombMin
{
  static constexpr int op = B200VEC_MIN;
  static __device__ __forceinline__ double identity() { return DBL_MAX; }
  static __device__ __forceinline__ double apply(double a, double b) { return (b < a) ? b : a; }
  static constexpr bool order_sensitive = false;
};

template <class C>
__device__ __forceinline__ double warp_combine(double v)
{
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
  {
    double o = __shfl_down_sync(0xffffffffu, v, off);
    v        = C::apply(v, o);
  }
  return v; /* valid in lane 0 */
}

/* block-wide combine of one value per thread; result valid in thread 0.
 * Fixed tree: lanes by shuffle, then the BLOCK/32 warp leaders by shuffle in warp 0. */
template <class C, int BLOCK = kBlock>
__device__ __forceinline__ double block_combine(double v, double* smem /* >= BLOCK/32 doubles */)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_combine<C>(v);
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  double r = C::identity();
  if (warp == 0)
  {
    r = (lane < BLOCK / 32) ? smem[lane] : C::identity();
    r = warp_combine<C>(r);
  }
  __syncthreads();
  return r;
}

/* ---- cross-rank combine over NVLink peer memory (warp-collective).
 * v: this rank's value, the same in all 32 lanes.  Lane r < nranks posts it into
 * rank r's mailbox (two tagged 8-byte stores: single-copy atomic, no fence) and
 * then polls the own mailbox for source r; the values are folded in RANK ORDER
 * 0,1,.. so every rank computes the bit-identical result (integrators must
 * branch identically everywhere) and it is independent of arrival order.
 * Replaces the MPI_Allreduce of nvector_manyvector.c:815-1793 / ncclAllReduce.
 * A peer that never arrives (dead rank) traps the kernel after ~30 s -- a loud
 * CUDA error on the host -- instead of hanging the GPU. */
template <class C>
__device__ __forceinline__ double xrank_combine_warp(double v, int slot, const XArgs& x)
{
  const int lane               = threadIdx.x & 31;
  const unsigned int par       = x.seq & 1u;
  const unsigned long long tag = (unsigned long long)x.seq << 32;
  double vq                    = C::identity();
  unsigned long long prof_t0   = 0;
  if (x.prof && lane == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_t0));
  /* mailbox of peer `lane` and the own one, picked by compile-time indices: a run-time index into
     the by-value XArgs would make the compiler copy the whole struct into LOCAL memory at the top of
     every kernel that can reach this function (an 88-byte stack frame per thread: ~13 MB of stores
     for a 296 x 512 grid, 3-4 us on a 20 us reduction -- measured, profiles/r02_reduce_attribution.md) */
  unsigned long long* mb_peer = nullptr;
  unsigned long long* mb_own  = nullptr;
#pragma unroll
  for (int r = 0; r < kMaxPeers; r++)
  {
    if (lane == r) mb_peer = x.mbox[r];
    if (x.rank == r) mb_own = x.mbox[r];
  }
  if (lane < x.nranks)
  {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    unsigned long long* dst = mb_peer + ((size_t)(par * kMaxPeers + x.rank) * kMaxOut + slot) * 2;
    const unsigned long long w0 = tag | (bits & 0xffffffffull), w1 = tag | (bits >> 32);
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst), "l"(w0) : "memory");
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst + 1), "l"(w1) : "memory");

    const unsigned long long* src = mb_own + ((size_t)(par * kMaxPeers + lane) * kMaxOut + slot) * 2;
    unsigned long long a, b, t0 = 0;
    unsigned int spins = 0;
    for (;;)
    {
      asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(a) : "l"(src) : "memory");
      asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(b) : "l"(src + 1) : "memory");
      if ((a >> 32) == x.seq && (b >> 32) == x.seq) break;
      if ((++spins & 0xfffu) == 0)
      {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 30000000000ull) asm volatile("trap;");
      }
    }
    vq = __longlong_as_double((long long)((b << 32) | (a & 0xffffffffull)));
  }
  double acc = __shfl_sync(0xffffffffu, vq, 0);
  for (int q = 1; q < x.nranks; q++) acc = C::apply(acc, __shfl_sync(0xffffffffu, vq, q));
  if (x.prof && lane == 0)
  {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    atomicAdd(x.prof, t1 - prof_t0);
    atomicAdd(x.prof + 1, 1ull);
  }
  return acc;
}

} // namespace b200
#endif
