/* b200vec_reduce.cu -- local reduction kernels for sm_100a.
 *
 * Two-stage, single-launch, deterministic:
 *   stage 1  every thread folds its grid-stride share sequentially into W
 *            register accumulators (wide loads, U tiles in flight), folds those
 *            in fixed order, then warp shuffle tree -> warp leaders -> CTA value;
 *   stage 2  single-output kernels: every CTA stores its value as a TAGGED pair
 *            (two 8-byte words, each (sequence << 32) | 32 value bits) into its slot
 *            of the context's partials array; CTA 0 polls the slots until they carry
 *            this launch's tag and folds them in FIXED INDEX ORDER (thread t takes
 *            t, t+512, ...; then the same tree) -- no ticket atomic, no fence, the
 *            finishing CTA is known in advance, and the result does not depend on
 *            scheduling: run-to-run bitwise reproducible.  Multi-output kernels
 *            (up to 24 outputs) keep a last-block-done ticket (one acq_rel atomic)
 *            and let warp w of the last CTA fold outputs w, w + 8, ...
 *            The value goes to the context's device result slot and -- when the host
 *            wants the scalar -- as two tagged 8-byte words to pinned, device-mapped
 *            host memory that the host polls: a scalar-returning N_Vector op needs no
 *            memcpy, no stream sync, no system fence and no reliance on 16-byte store
 *            atomicity (the reference needs H2D + D2H + sync, nvector_cuda.cu:2277-2411,
 *            and atomics on doubles, sundials_cuda_kernels.cuh:417-424, which are
 *            order-nondeterministic).  Measured (profiles/r02_reduce_attribution.md): 128 MiB
 *            stream in 20.6-21.0 us device time = the measured HBM copy peak; the API call
 *            adds 7.5-9 us of launch + PCIe that no kernel design removes.
 *
 * Exact-order path: for n <= exact_threshold (default 1024) one CTA stages the
 * per-element terms in shared memory and thread 0 adds them strictly
 * left-to-right -- bit-identical to nvector_serial.c's loops, which is what
 * keeps integrator step/iteration counts identical to the CPU reference on the
 * small regression problems (cvDiurnal_kry N=200, ark_heat2D N=1024).
 *
 * Fused forms built on the same epilogue: k_lincomb_sqnorm (linear combination + squared
 * norm of its result), k_reduce<RAxpyDot> (update + next projection), k_reduce<REwt> (the
 * integrators' error-weight vector), and the chained Gram-Schmidt sweeps that launch a
 * whole column back to back with device-resident coefficients and ONE host wait.
 *
 * Arithmetic per term follows serial exactly (e.g. prodi = x*w; sum += prodi*prodi,
 * serial:664-665 -- NOT x*w*x*w as VectorKernels.cuh:167 computes).
 * HBM bytes per element: dot/wsqrsum 16, masked 24, maxnorm/min/l1 8,
 * invtest/minquotient 16, constrmask 24, multi-dot 8(nv+1).
 */
#include <cstring>

#include "b200vec_device.cuh"

namespace b200 {

/* --------------------------------------------------------------- policies
 * term(x0,x1,x2, outv, store) -> contribution; outv/store only if HAS_OUT  */
struct RDot
{
  using Comb = CombSum;
  static constexpr int NIN = 2;
  static constexpr bool HAS_OUT = false;
  static constexpr bool NAN_HEAD = false;
  static constexpr int MAXU = 4;
  __device__ double term(double x, double y, double, double&, bool&) const { return x * y; }
};
/* x . x through ONE operand stream: Gram-Schmidt asks for N_VDotProd(v, v) once (classical) or
   twice (modified) per column; reading v through both operand slots costs 16 B/elt of L2 traffic
   for 8 B/elt of data.  Same products, same tile order: the same bits as RDot with y == x. */
struct RSqr
{
  using Comb = CombSum;
  static constexpr int NIN = 1;
  static constexpr bool HAS_OUT = false;
  static constexpr bool NAN_HEAD = false;
  static constexpr int MAXU = 4;
  __device__ double term(double x, double, double, double&, bool&) const { return x * x; }
};
struct RMaxNorm
{
  using Comb = CombMax;
  static constexpr int NIN = 1;
  static constexpr bool HAS_OUT = false;
  static constexpr bool NAN_HEAD = false;
  static constexpr int MAXU = 4;
  __device__ double term(double x, double, double, double&, bool&) const { return fabs(x); }
};
struct RMin
{
  using Comb = CombMin;
  static constexpr int NIN = 1;
  static constexpr bool HAS_OUT = false;
  static constexpr bool NAN_HEAD = true;
  static constexpr int MAXU = 4;
  __device__ double term(double x, double, double, double&, bool&) const { return x; }
};
struct RL1
{
  using Comb = CombSum;
  static constexpr int NIN = 1;
  static constexpr bool HAS_OUT = false;
  static constexpr bool NAN_HEAD = false;
  static constexpr int MAXU = 4;
  __device__ double term(double x, double, double, double&, bool&) const { return fabs(x); }
};
struct RWSqr
{
  using Comb = CombSum;
  static constexpr int NIN = 2;
  static constexpr bool HAS_OUT = false;
  static constexpr bool NAN_HEAD = false;
  static constexpr int MAXU = 4;
  __device__ double term(double x, double w, double, double&, bool&) const
  {
    const double p = x * w;
    return p * p;
  }
};
struct RWSqrMask
{
  using Comb = CombSum;
  static constexpr int NIN = 3;
  static constexpr bool HAS_OUT = false;
  static constexpr bool NAN_HEAD = false;
  static constexpr int MAXU = 4;
  __device__ double term(double x, double w, double id, double&, bool&) const
  {
    const double p = x * w;
    return (id > 0.0) ? p * p : 0.0; /* adding +0.0 to a non-negative sum is exact */
  }
};
/* flag reductions: value 1.0 = "ok", combined with min */
struct RInvTest
{
  using Comb = CombMin;
  static constexpr int NIN = 1;
  static constexpr bool HAS_OUT = true;
  static constexpr bool NAN_HEAD = false;
  static constexpr int MAXU = 4;
  __device__ double term(double x, double, double, double& outv, bool& store) const
  {
    store = (x != 0.0);
    outv  = 1.0 / x;
    return store ? 1.0 : 0.0;
  }
};
struct RConstrMask
{
  using Comb = CombMin;
  static constexpr int NIN = 2;
  static constexpr bool HAS_OUT = true;
  static constexpr bool NAN_HEAD = false;
  static constexpr int MAXU = 2 /* fp64 division / predicate chain: two tiles in flight fit 64 registers */;
  __device__ double term(double c, double x, double, double& outv, bool& store) const
  {
    store          = true;
    const double s = x * c;
    const double ac = fabs(c);
    const bool viol = (c != 0.0) && ((ac > 1.5 && s <= 0.0) || (ac > 0.5 && s < 0.0));
    outv            = viol ? 1.0 : 0.0;
    return viol ? 0.0 : 1.0;
  }
};
struct RMinQuot
{
  using Comb = CombMin;
  static constexpr int NIN = 2;
  static constexpr bool HAS_OUT = false;
  static constexpr bool NAN_HEAD = false;
  static constexpr int MAXU = 2 /* fp64 division / predicate chain: two tiles in flight fit 64 registers */;
  __device__ double term(double num, double den, double, double&, bool&) const
  {
    return (den == 0.0) ? DBL_MAX : num / den;
  }
};

/* z <- (a * x) + z and the contribution w * z_new of the UPDATED element: one step of a modified
   Gram-Schmidt sweep (N_VLinearSum(1, v_k, -h_i, v_i, v_k) followed by N_VDotProd(v_{i+1}, v_k),
   sundials_iterative.c:62-67) as one pass: 32 B/elt instead of 24 + 16.  Operands: p0 = x, p1 = z
   (also the output), p2 = w.  The update is serial's Vaxpy form (serial:1734-1760). */
struct RAxpyDot
{
  using Comb = CombSum;
  static constexpr int NIN = 3;
  static constexpr bool HAS_OUT = true;
  static constexpr bool NAN_HEAD = false;
  static constexpr int MAXU = 2;
  double a;
  const double* a_dev; /* NULL, or: a = -(*a_dev), read on the device when the kernel starts -- the
                          projection a preceding kernel of the same sweep left in a result slot, so the
                          host need not see it before launching this kernel */
  __device__ double term(double x, double z, double w, double& outv, bool& store) const
  {
    store = true;
    outv  = (a * x) + z;
    return w * outv;
  }
};

/* hook run by every reduction kernel after griddepcontrol.wait: policies with device-resident
   parameters load them here (the kernel's by-value copy of the policy is thread-private) */
template <class R>
__device__ __forceinline__ void prepare_policy(R&)
{}
__device__ __forceinline__ void prepare_policy(RAxpyDot& r)
{
  if (r.a_dev) r.a = -__ldcg(r.a_dev);
}

/* error-weight vector of the integrators in ONE pass: w_i = 1 / (rtol |y_i| + atol_i), and the minimum
   of the denominators (the integrators refuse a non-positive one when an absolute tolerance is zero).
   Replaces N_VAbs + N_VScale + N_VAddConst (or N_VLinearSum) + [N_VMin] + N_VInv of cvEwtSetSS/SV
   (src/cvode/cvode.c:4794-4860), arkEwtSetSS/SV (src/arkode/arkode.c:2935-2975), IDAEwtSetSS/SV -- the
   op sequence src/cvode/cvode_fused_gpu.cpp:62 (cvEwtSetSS_kernel) fuses for nvector_cuda -- with the
   same operation order, hence the same bits: 16 B/elt (24 with a vector atol) instead of 64 + 8.
   p0 = y, p1 = vector atol (VEC), out = w. */
template <bool VEC>
struct REwt
{
  using Comb = CombMin;
  static constexpr int NIN = VEC ? 2 : 1;
  static constexpr bool HAS_OUT = true;
  static constexpr bool NAN_HEAD = false;
  static constexpr int MAXU = VEC ? 2 : 4; /* measured: 43.2 vs 46.0 us (scalar atol), 60.3 vs 71.2 us (vector atol) at 2^24 */
  double rtol, atol;
  __device__ double term(double y, double av, double, double& outv, bool& store) const
  {
    const double t = (rtol * fabs(y)) + (VEC ? av : atol);
    store          = true;
    outv           = 1.0 / t;
    return t;
  }
};

struct RedPtrs
{
  const double* p0;
  const double* p1;
  const double* p2;
  double* out;
};

/* where a finished reduction publishes its value(s): the context's device slots and -- only when
   the host asked for the scalar(s), h_words != NULL -- two tagged 8-byte words per slot in pinned,
   device-mapped host memory: word = (tag << 32) | 32 bits of the value.  Each 8-byte store is
   single-copy atomic all the way across PCIe and carries its own tag, so there is no fence, no
   separate flag and no reliance on 16-byte store atomicity; the host polls the words instead of
   paying a cudaStreamSynchronize round trip. */
struct ResOut
{
  double* d_res;
  unsigned long long* h_words; /* words of slot 0 of this launch, or NULL: no host publication */
  unsigned int tag;            /* low 32 bits of the context's reduction sequence number: tags the
                                  CTA partials of THIS launch                                     */
  unsigned int htag;           /* tag of the pinned words: == tag, except inside a chained sweep
                                  (several kernels, one host wait) where all launches share one    */
};

__device__ __forceinline__ void store_tagged(unsigned long long* dst, unsigned int tag, double v)
{
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const unsigned long long t    = (unsigned long long)tag << 32;
  asm volatile("st.volatile.global.v2.u64 [%0], {%1,%2};" ::"l"(dst), "l"(t | (bits & 0xffffffffull)), "l"(t | (bits >> 32))
               : "memory");
}

__device__ __forceinline__ void publish_slot(const ResOut& o, int slot, double v)
{
  o.d_res[slot] = v;
  if (o.h_words) store_tagged(o.h_words + 2 * slot, o.htag, v);
}

/* ticket of the last-block-done scheme (multi-output kernels): ONE acq_rel atomic releases this
   CTA's partials and acquires everybody else's */
__device__ __forceinline__ bool take_ticket(unsigned int* counter)
{
  unsigned int t;
  asm volatile("atom.add.acq_rel.gpu.u32 %0, [%1], 1;" : "=r"(t) : "l"(counter) : "memory");
  return t == gridDim.x - 1;
}

__device__ __forceinline__ unsigned long long global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

/* local value of this rank (valid in thread 0 of the calling CTA) -> optional
   cross-rank combine by warp 0 -> publication by thread 0.  Called by all
   threads of ONE CTA (the finishing one, or the only one). */
template <class C>
__device__ __forceinline__ void combine_and_publish(double a, const ResOut& o, const XArgs& x)
{
  if (x.nranks > 1)
  {
    if (threadIdx.x >= 32) return;
    a = __shfl_sync(0xffffffffu, a, 0);
    a = xrank_combine_warp<C>(a, 0, x);
  }
  if (threadIdx.x == 0)
  {
    publish_slot(o, 0, a);
    if (x.prof) x.prof[7] = global_ns();
  }
}

/* N_VMin starts from x[0] (serial:715-720): a NaN there is returned, a NaN anywhere else never
   wins a strict comparison.  Mirrored here on the local block (on a distributed vector the fold
   over ranks starts from rank 0's value, so the global result is serial's on the whole vector). */
template <bool NANHEAD>
__device__ __forceinline__ double nan_head(double a, const double* p0)
{
  if (NANHEAD)
  {
    const double h = *p0;
    if (h != h) return h;
  }
  return a;
}

/* stage 2.  Every CTA has its value in thread 0 and stores it as a TAGGED pair into its slot of
   the context's partials array; CTA 0 then polls the slots of all CTAs (tag == this launch's
   sequence number) and folds them in fixed index order -- no ticket atomic, no fence, and the
   finishing CTA is known in advance.  CTA 0 never waits for a CTA that has not been scheduled
   while it holds resources that CTA needs beyond one CTA slot: every other CTA runs to completion
   without waiting for anybody. */
template <class C, int BLOCK, bool NANHEAD>
__device__ __forceinline__ void finish_block(double v, unsigned long long* tagged, const ResOut& o, const XArgs& x,
                                             double* smem, const double* p0)
{
  if (gridDim.x == 1)
  {
    if (threadIdx.x == 0) v = nan_head<NANHEAD>(v, p0);
    combine_and_publish<C>(v, o, x);
    return;
  }
  if (threadIdx.x == 0 && blockIdx.x != 0) store_tagged(tagged + 2 * (size_t)blockIdx.x, o.tag, v);
  if (blockIdx.x != 0) return;
  double a = C::identity();
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += BLOCK)
  {
    if (i == 0)
    {
      a = C::apply(a, v); /* own value: thread 0 */
      continue;
    }
    const unsigned long long* src = tagged + 2 * (size_t)i;
    unsigned long long w0, w1, t0 = 0;
    unsigned int spins = 0;
    for (;;)
    {
      asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
      if ((unsigned int)(w0 >> 32) == o.tag && (unsigned int)(w1 >> 32) == o.tag) break;
      if ((++spins & 0xfffu) == 0)
      { /* a CTA of this grid that never reports is a lost launch: fail loudly, do not hang */
        const unsigned long long now = global_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 30000000000ull) asm volatile("trap;");
      }
    }
    a = C::apply(a, __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull))));
  }
  a = block_combine<C, BLOCK>(a, smem);
  if (threadIdx.x == 0) a = nan_head<NANHEAD>(a, p0);
  combine_and_publish<C>(a, o, x);
}

/* one CTA tile (TILE doubles of every input operand) -> L2, issued by 4 threads of the CTA as
   4 bulk prefetches per operand: the DRAM fetch of a later tile is in flight while the CTA
   folds the current one, independently of how many wide loads ptxas keeps outstanding */
template <int NIN, int64_t TILE>
__device__ __forceinline__ void prefetch_tile_l2(const RedPtrs& p, int64_t tile)
{
  static_assert((TILE / 4 * 8) % 16 == 0, "bulk prefetch size is a multiple of 16 bytes");
  if ((threadIdx.x & (kRBlock / 4 - 1)) != 0) return;
  const int64_t off      = tile * TILE + (int64_t)(threadIdx.x / (kRBlock / 4)) * (TILE / 4);
  constexpr uint32_t by  = (uint32_t)(TILE / 4 * 8);
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.p0 + off), "r"(by) : "memory");
  if (NIN >= 2) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.p1 + off), "r"(by) : "memory");
  if (NIN >= 3) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.p2 + off), "r"(by) : "memory");
}

template <int W, int U, class R>
__global__ void __launch_bounds__(kRBlock, 2)
  k_reduce(R r, RedPtrs p, int64_t n, unsigned long long* tagged, ResOut o, const __grid_constant__ XArgs x, int pf)
{
  using C = typename R::Comb;
  __shared__ double smem[kRBlock / 32];
  constexpr int64_t TILE = (int64_t)kRBlock * W * U;
  constexpr int64_t STEP = (int64_t)kRBlock * W;
  const int64_t nfull    = n / TILE;
  pdl_prologue();
  prepare_policy(r);
  if (x.prof && threadIdx.x == 0) atomicMin(x.prof + 6, global_ns());
  if (W < 2) pf = 0; /* bulk prefetch needs 16-byte aligned addresses */

  double acc[W];
#pragma unroll
  for (int w = 0; w < W; w++) acc[w] = C::identity();

  for (int d = 1; d < pf; d++)
  {
    const int64_t tn = (int64_t)blockIdx.x + (int64_t)d * gridDim.x;
    if (tn < nfull) prefetch_tile_l2<R::NIN, TILE>(p, tn);
  }
  for (int64_t t = blockIdx.x; t < nfull; t += gridDim.x)
  {
    if (pf > 0)
    {
      const int64_t tn = t + (int64_t)pf * gridDim.x;
      if (tn < nfull) prefetch_tile_l2<R::NIN, TILE>(p, tn);
    }
    const int64_t base = t * TILE + (int64_t)threadIdx.x * W;
    double a[U][W], b[U][W], c[U][W];
#pragma unroll
    for (int u = 0; u < U; u++)
    {
      ldg<W>(p.p0 + base + u * STEP, a[u]);
      if (R::NIN >= 2) ldg<W>(p.p1 + base + u * STEP, b[u]);
      if (R::NIN >= 3) ldg<W>(p.p2 + base + u * STEP, c[u]);
    }
#pragma unroll
    for (int u = 0; u < U; u++)
    {
      double o[W];
      bool st[W];
      bool all = true;
#pragma unroll
      for (int w = 0; w < W; w++)
      {
        st[w]          = false;
        const double v = r.term(a[u][w], R::NIN >= 2 ? b[u][w] : 0.0, R::NIN >= 3 ? c[u][w] : 0.0, o[w], st[w]);
        acc[w]         = C::apply(acc[w], v);
        all            = all && st[w];
      }
      if (R::HAS_OUT)
      {
        double* q = p.out + base + u * STEP;
        if (all) stg<W>(q, o);
        else
        {
#pragma unroll
          for (int w = 0; w < W; w++)
            if (st[w]) q[w] = o[w];
        }
      }
    }
  }

  const int64_t tail0 = nfull * TILE;
  if (tail0 < n && blockIdx.x == (unsigned)(nfull % gridDim.x))
  {
    for (int64_t i = tail0 + threadIdx.x; i < n; i += kRBlock)
    {
      double o;
      bool st        = false;
      const double v = r.term(p.p0[i], R::NIN >= 2 ? p.p1[i] : 0.0, R::NIN >= 3 ? p.p2[i] : 0.0, o, st);
      acc[0]         = C::apply(acc[0], v);
      if (R::HAS_OUT && st) p.out[i] = o;
    }
  }

  double v = acc[0];
#pragma unroll
  for (int w = 1; w < W; w++) v = C::apply(v, acc[w]);
  v = block_combine<C, kRBlock>(v, smem);
  finish_block<C, kRBlock, R::NAN_HEAD>(v, tagged, o, x, smem, p.p0);
}

/* exact-order path: one CTA, terms staged in shared memory, thread 0 folds
   them left-to-right exactly as the serial loop does */
template <class R>
__global__ void __launch_bounds__(kBlock) k_reduce_exact(R r, RedPtrs p, int n, ResOut o, const __grid_constant__ XArgs x)
{
  using C = typename R::Comb;
  __shared__ double buf[kExactMaxElems];
  pdl_prologue();
  prepare_policy(r);
  if (x.prof && threadIdx.x == 0) atomicMin(x.prof + 6, global_ns());
  for (int i = threadIdx.x; i < n; i += kBlock)
  {
    double o;
    bool st = false;
    buf[i]  = r.term(p.p0[i], R::NIN >= 2 ? p.p1[i] : 0.0, R::NIN >= 3 ? p.p2[i] : 0.0, o, st);
    if (R::HAS_OUT && st) p.out[i] = o;
  }
  __syncthreads();
  double a = C::identity();
  if (threadIdx.x == 0)
  {
#pragma unroll 8
    for (int i = 0; i < n; i++) a = C::apply(a, buf[i]);
    if (n > 0) a = nan_head<R::NAN_HEAD>(a, p.p0);
  }
  combine_and_publish<C>(a, o, x);
}

/* next sequence number + where the kernel publishes; to_host = the caller wants
   the scalar(s) on the host (otherwise they stay in the device slots, e.g. for
   the allreduce that follows on a distributed vector) */
static ResOut next_out(b200vec_ctx ctx, int slot0, bool to_host)
{
  ResOut o;
  ++ctx->seq;
  o.d_res   = ctx->d_result + slot0;
  o.h_words = to_host ? (unsigned long long*)(ctx->h_result_dev + kMaxRows) + 2 * slot0 : nullptr;
  o.tag     = (unsigned int)ctx->seq;
  o.htag    = o.tag;
  return o;
}

/* wait for the pinned words of result slots [0, count) to carry the tag of the reduction launched
   last, then decode them.  The final pass stores them straight into pinned host memory: polling
   costs ~1 us after the kernel's last store, a cudaStreamSynchronize round trip several times that.
   Bounded spin, then fall back to the sync (which also surfaces asynchronous errors). */
static int finish_host(b200vec_ctx ctx, int slot0, int count, double* result_host, long long want_tag = -1)
{
  if (!result_host) return B200VEC_OK;
  volatile unsigned long long* w = (volatile unsigned long long*)(ctx->h_result + kMaxRows) + 2 * slot0;
  const unsigned int want        = (want_tag < 0) ? (unsigned int)ctx->seq : (unsigned int)want_tag;
  bool seen                      = false;
  if (ctx->tune.spin_wait)
  {
    for (long spins = 0; spins < 20000000L && !seen; spins++)
    {
      seen = true;
      for (int i = 2 * count - 1; i >= 0; i--)
        if ((unsigned int)(w[i] >> 32) != want)
        {
          seen = false;
          break;
        }
      if (seen) break;
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
      if ((spins & 0xffff) == 0xffff && cudaStreamQuery(ctx->stream) != cudaErrorNotReady) break;
    }
  }
  if (!seen)
  {
    int rc = check_cuda(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize(reduction)");
    if (rc) return rc;
    for (int i = 0; i < 2 * count; i++)
      if ((unsigned int)(w[i] >> 32) != want)
        return set_error(B200VEC_ERR_CUDA, "reduction finished without publishing result slot %d", i / 2);
  }
  __atomic_thread_fence(__ATOMIC_ACQUIRE);
  for (int i = 0; i < count; i++)
  {
    const unsigned long long bits = (w[2 * i + 1] << 32) | (w[2 * i] & 0xffffffffull);
    memcpy(result_host + i, &bits, sizeof(double));
  }
  return B200VEC_OK;
}

int finish_reduction(b200vec_ctx ctx, int count, double* result_host) { return finish_host(ctx, 0, count, result_host); }

/* global reduction without the peer transport: allreduce the device slots with
   NCCL on the same stream, then one fetch (nvector_manyvector.c:815 pattern) */
static int finish_global_nccl(b200vec_ctx ctx, int count, int op, double* result_host)
{
  int rc = b200vec_allreduce(ctx, count, op);
  if (!rc && result_host) rc = b200vec_result_fetch(ctx, count, result_host);
  return rc;
}

/* grid of a single-output reduction: every CTA adds a slot to the finishing CTA's polling pass and
   small vectors are latency-bound, so below 2^20 elements one CTA per SM at most */
static int reduce_grid_cap(b200vec_ctx ctx, int64_t n)
{
  int64_t cap = ctx->tune.max_blocks;
  if (cap == kMaxBlocksDef && n <= ((int64_t)1 << 20)) cap = kSMs;
  return (int)cap;
}

/* launch ONE single-output reduction kernel for policy R (exact-order or tree form), result into
   slot `slot0`; no host wait */
template <class R>
static int launch_reduce_kernel(b200vec_ctx ctx, const char* name, R r, RedPtrs p, int64_t n, const ResOut& out,
                                const XArgs& xa)
{
  if (n <= ctx->tune.exact_threshold) { launch_k(ctx, k_reduce_exact<R>, dim3(1), dim3(kBlock), r, p, (int)n, out, xa); }
  else
  {
    int wmax = align_width(p.p0);
    wmax     = min(wmax, align_width(p.p1));
    wmax     = min(wmax, align_width(p.p2));
    wmax     = min(wmax, align_width(p.out));
    MapCfg c = pick_map_cfg(ctx, n, wmax, true, kRBlock);
    if (c.U > R::MAXU)
    { /* keep the instantiation within 64 registers (2 CTAs x 512 threads per SM) */
      c.U           = R::MAXU;
      int64_t tiles = n / ((int64_t)kRBlock * c.W * c.U);
      if (tiles < 1) tiles = 1;
      c.grid = (int)((tiles < ctx->tune.max_blocks) ? tiles : ctx->tune.max_blocks);
    }
    c.grid = min(c.grid, reduce_grid_cap(ctx, n));
    if (R::HAS_OUT && ctx->tune.max_blocks == kMaxBlocksDef && n > ((int64_t)1 << 20))
    { /* InvTest / ConstrMask / AxpyDot also WRITE a vector: like the streaming kernels they run best
         with one tile per CTA (the block scheduler back-fills SMs as CTAs retire; InvTest 48.2 -> 42.4 us
         at 2^24, profiles/r02_reduce_attribution.md); CTA 0 then polls up to kMaxPartialBlocks slots */
      int64_t tiles = n / ((int64_t)kRBlock * c.W * c.U);
      if (tiles < 1) tiles = 1;
      c.grid = (int)((tiles < kMaxPartialBlocks) ? tiles : kMaxPartialBlocks);
    }
#define B200_RED_CASE(WW, UU)    \
  if (c.W == WW && c.U == UU)    \
  launch_k(ctx, k_reduce<WW, UU, R>, dim3(c.grid), dim3(kRBlock), r, p, n, ctx->d_tagged, out, xa, \
           (int)ctx->tune.l2_prefetch)
    B200_RED_CASE(4, 4);
    else B200_RED_CASE(4, 2);
    else B200_RED_CASE(4, 1);
    else B200_RED_CASE(2, 4);
    else B200_RED_CASE(2, 2);
    else B200_RED_CASE(2, 1);
    else B200_RED_CASE(1, 4);
    else B200_RED_CASE(1, 2);
    else B200_RED_CASE(1, 1);
#undef B200_RED_CASE
  }
  return check_launch(ctx, name);
}

template <class R>
static int launch_reduce(b200vec_ctx ctx, const char* name, R r, RedPtrs p, int64_t n, double empty_value,
                         double* result_host)
{
  using C = typename R::Comb;
  DeviceGuard g(ctx->device);
  cudaStream_t s = ctx->stream;
  XArgs xa;
  const int scope = take_scope(ctx, &xa); /* 0 local, 1 global over peer memory, 2 global over NCCL */
  if (n == 0 && scope == 0)
  {
    /* nothing to read: publish the identity (N_VMin on an empty vector is
       undefined in the reference, serial:715; we return DBL_MAX) */
    ctx->h_result[0] = empty_value;
    int rc = check_cuda(cudaMemcpyAsync(ctx->d_result, ctx->h_result, sizeof(double), cudaMemcpyHostToDevice, s),
                        "cudaMemcpyAsync(empty reduction)");
    if (rc) return rc;
    rc = check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize");
    if (rc) return rc;
    if (result_host) *result_host = empty_value;
    return B200VEC_OK;
  }
  /* an empty local block of a distributed vector still takes part: the exact
     kernel with n == 0 contributes the identity */
  const bool to_host = (result_host != nullptr) && scope != 2;
  const ResOut out   = next_out(ctx, 0, to_host);
  int rc             = launch_reduce_kernel(ctx, name, r, p, n, out, xa);
  if (rc) return rc;
  if (scope == 2) return finish_global_nccl(ctx, 1, C::op, result_host);
  return finish_host(ctx, 0, 1, result_host);
}

/* ------------------------------------------------------ multi-output family
 * MODE 0: d_j = sum x * A_j            (DotProdMulti; shared = x)
 * MODE 1: d_j = sum (A_j * B_j)^2      (WrmsNormVectorArray)
 * MODE 2: same, masked by shared > 0   (WrmsNormMaskVectorArray; shared = id)
 * Up to kMaxOut (MODE 0) / kMaxPair (MODE 1, 2) outputs per launch; the shared operand is read
 * once per element and every A_j/B_j exactly once. */
struct MultiArgs
{
  const double* shared;
  const double* A[kMaxOut];
  const double* B[kMaxPair];
  int nout;
  int self_j; /* MODE 0: index j with A[j] == shared (classical Gram-Schmidt puts x itself into Y,
                 sundials_iterative.c:135), -1 if none: that operand is not loaded a second time */
};

template <int MODE>
__device__ __forceinline__ double multi_term(double sh, double a, double b)
{
  if (MODE == 0) return sh * a;
  const double p = a * b;
  if (MODE == 1) return p * p;
  return (sh > 0.0) ? p * p : 0.0;
}

/* s_fin[0..nout): this rank's values (visible to the whole CTA).  Warp w folds outputs w, w + 8, ..
   across ranks, then threads j < nout publish slot j (device slot + tagged pinned words).
   Called by ALL threads of one CTA. */
__device__ __forceinline__ void multi_combine_and_publish(double* s_fin, int nout, const ResOut& o, const XArgs& x)
{
  if (x.nranks > 1)
  {
    const int warp = threadIdx.x >> 5;
    for (int j = warp; j < nout; j += kBlock / 32)
    {
      const double v = xrank_combine_warp<CombSum>(s_fin[j], j, x);
      __syncwarp();
      if ((threadIdx.x & 31) == 0) s_fin[j] = v;
    }
    __syncthreads();
  }
  if (threadIdx.x < nout) publish_slot(o, threadIdx.x, s_fin[threadIdx.x]);
  if (x.prof && threadIdx.x == 0) x.prof[7] = global_ns();
}

/* NO = outputs compiled in (2, 4, 8, 16 or 24; the launch's nout <= NO), U = tiles in flight per
   thread.  Small output counts are what a GMRES cycle issues (classical Gram-Schmidt at column
   k = 1..maxl is a (k+1)-wide multi-dot): with one tile per thread they would keep only
   (1 + nout) x 32 B per thread outstanding on a few hundred resident threads -- too little to
   cover HBM latency -- so the 2-output bucket unrolls two tiles and all small buckets, having
   fewer accumulators and operands in registers, run more CTAs per SM (the grid is sized from
   the occupancy of the instantiation).  Beyond 8 outputs the operands are loaded in batches of 8
   while the shared operand's tile stays in registers: x is still read once for all of them. */
template <int W, int MODE, int NO, int U>
__global__ void __launch_bounds__(kBlock) k_reduce_multi(const __grid_constant__ MultiArgs m, int64_t n,
                                                         double* partials, unsigned int* counter, ResOut o, const __grid_constant__ XArgs x)
{
  static_assert(NO <= kMaxOut && (MODE == 0 || NO <= kMaxPair), "outputs per launch");
  constexpr int NB = (NO < 8) ? NO : 8; /* operands whose loads are in flight together */
  static_assert(NO % NB == 0, "whole batches");
  __shared__ double s_fin[kMaxOut];
  __shared__ bool s_last;
  constexpr int64_t STEP = (int64_t)kBlock * W;
  constexpr int64_t TILE = STEP * U;
  const int64_t nfull    = n / TILE;
  const int nout         = m.nout;
  const int self_j       = m.self_j;

  double acc[NO];
#pragma unroll
  for (int j = 0; j < NO; j++) acc[j] = 0.0;
  pdl_prologue();
  if (x.prof && threadIdx.x == 0) atomicMin(x.prof + 6, global_ns());

  for (int64_t t = blockIdx.x; t < nfull; t += gridDim.x)
  {
    const int64_t base = t * TILE + (int64_t)threadIdx.x * W;
    double sh[U][W];
#pragma unroll
    for (int u = 0; u < U; u++)
      if (MODE != 1) ldg<W>(m.shared + base + u * STEP, sh[u]);
#pragma unroll
    for (int j0 = 0; j0 < NO; j0 += NB)
    {
      double a[U][NB][W], b[U][NB][W];
#pragma unroll
      for (int u = 0; u < U; u++)
      {
        const int64_t off = base + u * STEP;
#pragma unroll
        for (int jj = 0; jj < NB; jj++)
          if (j0 + jj < nout)
          {
            /* every load of the batch is issued before any value is used; the self operand is
               taken from sh at fold time (its a[u][jj] stays unloaded and unselected) */
            if (!(MODE == 0 && j0 + jj == self_j)) ldg<W>(m.A[j0 + jj] + off, a[u][jj]);
            if (MODE != 0) ldg<W>(m.B[(j0 + jj) % kMaxPair] + off, b[u][jj]);
          }
      }
#pragma unroll
      for (int u = 0; u < U; u++)
      {
#pragma unroll
        for (int jj = 0; jj < NB; jj++)
          if (j0 + jj < nout)
          {
            const bool self = (MODE == 0 && j0 + jj == self_j);
#pragma unroll
            for (int w = 0; w < W; w++)
              acc[j0 + jj] += multi_term<MODE>(MODE != 1 ? sh[u][w] : 0.0, self ? sh[u][w] : a[u][jj][w],
                                               MODE != 0 ? b[u][jj][w] : 0.0);
          }
      }
    }
  }

  const int64_t tail0 = nfull * TILE;
  if (tail0 < n && blockIdx.x == (unsigned)(nfull % gridDim.x))
  {
    for (int64_t i = tail0 + threadIdx.x; i < n; i += kBlock)
    {
      const double sh = (MODE != 1) ? m.shared[i] : 0.0;
#pragma unroll
      for (int j = 0; j < NO; j++)
        if (j < nout) acc[j] += multi_term<MODE>(sh, m.A[j][i], MODE != 0 ? m.B[j % kMaxPair][i] : 0.0);
    }
  }

  /* CTA value of every output -> partials[j][block].  All outputs share ONE barrier: every warp
     folds its lanes per output by shuffle, lane 0 parks the NO warp values in shared memory, and
     after the barrier thread j folds output j's kBlock/32 warp values in the fixed tree
     ((w0+w4)+(w2+w6))+((w1+w5)+(w3+w7)) -- the order block_combine uses, so the bits are the same. */
  static_assert(kBlock / 32 == 8, "warp-value tree below is written for 8 warps");
  __shared__ double s_w[kMaxOut][kBlock / 32];
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NO; j++)
      if (j < nout)
      {
        const double v = warp_combine<CombSum>(acc[j]);
        if (lane == 0) s_w[j][warp] = v;
      }
    __syncthreads();
    if (threadIdx.x < nout)
    {
      const double* w = s_w[threadIdx.x];
      const double v  = ((w[0] + w[4]) + (w[2] + w[6])) + ((w[1] + w[5]) + (w[3] + w[7]));
      if (gridDim.x == 1) s_fin[threadIdx.x] = v;
      else partials[(size_t)threadIdx.x * kMaxPartialBlocks + blockIdx.x] = v;
    }
  }
  if (gridDim.x == 1)
  {
    __syncthreads();
    multi_combine_and_publish(s_fin, nout, o, x);
    return;
  }

  /* the partials were written by threads 0..nout-1: make them visible before thread 0 takes the
     ticket (release) */
  if (threadIdx.x < nout) __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = take_ticket(counter);
  __syncthreads();
  if (!s_last) return;
  /* last CTA: warp w folds the partials of outputs w, w + 8, .. (lane-strided, then the shuffle
     tree) -- all outputs at once instead of one block-wide pass per output; fixed order, so
     run-to-run identical */
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = warp; j < nout; j += kBlock / 32)
    {
      const double* row = partials + (size_t)j * kMaxPartialBlocks;
      double a          = 0.0;
      for (unsigned int i = lane; i < gridDim.x; i += 32) a += __ldcg(row + i);
      a = warp_combine<CombSum>(a);
      if (lane == 0) s_fin[j] = a;
    }
  }
  if (threadIdx.x == 0) *counter = 0u;
  __syncthreads();
  multi_combine_and_publish(s_fin, nout, o, x);
}

/* exact-order multi: n * nout <= kExactMaxElems; thread j folds column j */
template <int MODE>
__global__ void __launch_bounds__(kBlock)
  k_reduce_multi_exact(const __grid_constant__ MultiArgs m, int n, ResOut o, const __grid_constant__ XArgs x)
{
  __shared__ double buf[kExactMaxElems];
  __shared__ double s_fin[kMaxOut];
  const int nout = m.nout;
  pdl_prologue();
  for (int i = threadIdx.x; i < n; i += kBlock)
  {
    const double sh = (MODE != 1) ? m.shared[i] : 0.0;
    for (int j = 0; j < nout; j++)
      buf[j * n + i] = multi_term<MODE>(sh, m.A[j][i], MODE != 0 ? m.B[j % kMaxPair][i] : 0.0);
  }
  __syncthreads();
  if (threadIdx.x < nout)
  {
    const double* col = buf + threadIdx.x * n;
    double a          = 0.0;
#pragma unroll 8
    for (int i = 0; i < n; i++) a += col[i];
    s_fin[threadIdx.x] = a;
  }
  __syncthreads();
  multi_combine_and_publish(s_fin, nout, o, x);
}

/* CTAs of one instantiation that fit an SM (registers decide: 8 outputs x 2 operand arrays in
   256-bit form take ~170, the 2-output bucket ~70), asked once from the runtime */
template <int W, int MODE, int NO, int U>
static int multi_resident_ctas()
{
  static int nb = 0;
  if (nb == 0)
  {
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_reduce_multi<W, MODE, NO, U>, kBlock, 0) != cudaSuccess || v < 1)
    {
      (void)cudaGetLastError();
      v = 1;
    }
    nb = (v > 4) ? 4 : v; /* 1024 threads per SM are plenty; keeps the partials rows short */
  }
  return nb;
}

/* persistent-style launch: exactly the resident CTA count of the instantiation (measured best for
   these register-heavy kernels), fewer when the vector has fewer tiles */
template <int W, int MODE, int NO, int U>
static int launch_multi_cfg(b200vec_ctx ctx, const char* name, const MultiArgs& m, int64_t n, const ResOut& out,
                            const XArgs& xa)
{
  int64_t tiles = n / ((int64_t)kBlock * W * U);
  if (tiles < 1) tiles = 1;
  int64_t cap = (int64_t)kSMs * multi_resident_ctas<W, MODE, NO, U>();
  if (ctx->tune.max_blocks != kMaxBlocksDef && cap > ctx->tune.max_blocks) cap = ctx->tune.max_blocks; /* user override */
  if (cap > kMaxPartialBlocks) cap = kMaxPartialBlocks;
  const int grid = (int)((tiles < cap) ? tiles : cap);
  launch_k(ctx, k_reduce_multi<W, MODE, NO, U>, dim3(grid), dim3(kBlock), m, n, ctx->d_partials, ctx->d_count, out, xa);
  return check_launch(ctx, name);
}

template <int MODE>
static int launch_multi_group(b200vec_ctx ctx, const char* name, const MultiArgs& m, int64_t n, int slot0,
                              bool to_host, const XArgs& xa, long long htag = -1)
{
  ResOut out = next_out(ctx, slot0, to_host);
  if (htag >= 0) out.htag = (unsigned int)htag;
  if (n <= ctx->tune.exact_threshold && n * m.nout <= kExactMaxElems)
  {
    launch_k(ctx, k_reduce_multi_exact<MODE>, dim3(1), dim3(kBlock), m, (int)n, out, xa);
    return check_launch(ctx, name);
  }
  int wmax = align_width(m.shared);
  for (int j = 0; j < m.nout; j++)
  {
    wmax = min(wmax, align_width(m.A[j]));
    if (MODE != 0) wmax = min(wmax, align_width(m.B[j]));
  }
  int W = wmax;
  if (ctx->tune.vec_width > 0 && ctx->tune.vec_width < W) W = (int)ctx->tune.vec_width;
  const int bucket = (m.nout <= 2) ? 2 : (m.nout <= 4) ? 4 : (m.nout <= 8) ? 8 : (m.nout <= 16) ? 16 : 24;
#define B200_MULTI_CASE(WW, NN, UU)                                                                        \
  if (W == WW && bucket == NN)                                                                             \
  return launch_multi_cfg<WW, MODE, NN, UU>(ctx, name, m, n, out, xa)
  B200_MULTI_CASE(4, 2, 2);
  B200_MULTI_CASE(4, 4, 1);
  B200_MULTI_CASE(4, 8, 1);
  B200_MULTI_CASE(2, 2, 2);
  B200_MULTI_CASE(2, 4, 1);
  B200_MULTI_CASE(2, 8, 1);
  B200_MULTI_CASE(1, 2, 2);
  B200_MULTI_CASE(1, 4, 1);
  B200_MULTI_CASE(1, 8, 1);
  if constexpr (MODE == 0)
  {
    B200_MULTI_CASE(4, 16, 1);
    B200_MULTI_CASE(4, 24, 1);
    B200_MULTI_CASE(2, 16, 1);
    B200_MULTI_CASE(2, 24, 1);
    B200_MULTI_CASE(1, 16, 1);
    B200_MULTI_CASE(1, 24, 1);
  }
#undef B200_MULTI_CASE
  return set_error(B200VEC_ERR_ARG, "%s: no kernel for width %d", name, W);
}

/* nout outputs in groups of <= kMaxOut / kMaxPair (fewer on the exact path so the
   staged terms fit shared memory); every group's results land in slots [j0, j0 + group) and the
   host collects all of them after the last launch */
template <int MODE>
static int launch_multi(b200vec_ctx ctx, const char* name, int nout, const double* shared, const double* const* A,
                        const double* const* B, int64_t n, double* result_host)
{
  if (nout > kMaxRows)
  {
    ctx->scope_global = false;
    return set_error(B200VEC_ERR_ARG, "%s: at most %d outputs per call", name, kMaxRows);
  }
  DeviceGuard g(ctx->device);
  XArgs xa;
  const int scope = take_scope(ctx, &xa);
  if (n == 0 && scope == 0)
  {
    for (int j = 0; j < nout; j++) ctx->h_result[j] = 0.0;
    int rc = check_cuda(cudaMemcpyAsync(ctx->d_result, ctx->h_result, sizeof(double) * nout, cudaMemcpyHostToDevice,
                                        ctx->stream),
                        "cudaMemcpyAsync(empty reduction)");
    if (!rc) rc = check_cuda(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
    if (!rc && result_host)
      for (int j = 0; j < nout; j++) result_host[j] = 0.0;
    return rc;
  }
  int group = (MODE == 0) ? kMaxOut : kMaxPair;
  if (n > 0 && n <= ctx->tune.exact_threshold)
  {
    int fit = (int)(kExactMaxElems / n);
    if (fit < 1) fit = 1;
    if (fit < group) group = fit;
  }
  const bool to_host = (result_host != nullptr) && scope != 2;
  for (int j0 = 0; j0 < nout; j0 += group)
  {
    MultiArgs m;
    m.shared = shared;
    m.nout   = (nout - j0 < group) ? nout - j0 : group;
    m.self_j = -1;
    for (int j = 0; j < kMaxOut; j++)
    {
      m.A[j] = (j < m.nout) ? A[j0 + j] : nullptr;
      if (j < kMaxPair) m.B[j] = (j < m.nout && B) ? B[j0 + j] : nullptr;
      if (MODE == 0 && m.self_j < 0 && j < m.nout && shared && m.A[j] == shared) m.self_j = j;
    }
    if (j0 > 0 && scope == 1) next_xargs(ctx, &xa); /* every group is its own collective */
    int rc = launch_multi_group<MODE>(ctx, name, m, n, j0, to_host, xa);
    if (rc) return rc;
    /* each launch publishes under its own tag: collect this group's slots before the next launch
       re-tags (the host would otherwise see stale tags on the earlier slots) */
    if (to_host)
    {
      rc = finish_host(ctx, j0, m.nout, result_host + j0);
      if (rc) return rc;
    }
  }
  if (scope == 2) return finish_global_nccl(ctx, nout, B200VEC_SUM, result_host);
  return B200VEC_OK;
}

/* ------------------------------------------------- fused linear combination + squared norm
 * z = sum_i c_i X_i (z may alias X[0]) and, in the same pass, sum_k z_k^2 of the values just
 * written.  This is the second half of a classical Gram-Schmidt step (sundials_iterative.c:137-152:
 * N_VLinearCombination followed by N_VDotProd(v[k], v[k])) in ONE kernel and ONE host round trip:
 * v[k] is not read back (8 B/elt saved) and the third launch + sync disappears.  The combination is
 * k_lincomb_rows' arithmetic (register accumulator in serial's term order: bit-identical z); the
 * norm uses the single-output reductions' epilogue (tagged partials, CTA 0 folds in fixed order).
 * Bytes per element: 8 (nterms + 1). */
constexpr int kLcnMaxTerms = 32;
constexpr int kLcnBatch    = 4;

struct LcNormArgs
{
  const double* X[kLcnMaxTerms];
  double c[kLcnMaxTerms];
  const double* c_dev; /* NULL, or device-resident coefficients: c_0 = 1, c_j = -c_dev[j-1] (the projections
                          a preceding multi-dot left in the result slots: classical Gram-Schmidt's
                          stemp[i+1] = -stemp[i], sundials_iterative.c:135) */
  double* z;
  int nterms;
};

template <int W>
__global__ void __launch_bounds__(kBlock)
  k_lincomb_sqnorm(const __grid_constant__ LcNormArgs a, int64_t n, unsigned long long* tagged, ResOut o, const __grid_constant__ XArgs x)
{
  __shared__ double s_c[kLcnMaxTerms];
  __shared__ const double* s_x[kLcnMaxTerms];
  __shared__ double smem[kBlock / 32];
  const int nterms = a.nterms;
  if (threadIdx.x < nterms) s_x[threadIdx.x] = a.X[threadIdx.x];
  pdl_prologue(); /* device-resident coefficients are valid only after the wait */
  if (threadIdx.x < nterms)
    s_c[threadIdx.x] = a.c_dev ? (threadIdx.x == 0 ? 1.0 : -__ldcg(a.c_dev + threadIdx.x - 1)) : a.c[threadIdx.x];
  if (x.prof && threadIdx.x == 0) atomicMin(x.prof + 6, global_ns());
  __syncthreads();
  double* z = a.z;

  constexpr int64_t TILE = (int64_t)kBlock * W;
  const int64_t nfull    = n / TILE;
  double sq[W];
#pragma unroll
  for (int w = 0; w < W; w++) sq[w] = 0.0;
  for (int64_t t = blockIdx.x; t < nfull; t += gridDim.x)
  {
    const int64_t base = t * TILE + (int64_t)threadIdx.x * W;
    double acc[W];
    for (int i0 = 0; i0 < nterms; i0 += kLcnBatch)
    {
      double v[kLcnBatch][W];
#pragma unroll
      for (int k = 0; k < kLcnBatch; k++)
        if (i0 + k < nterms) ldg<W>(s_x[i0 + k] + base, v[k]);
#pragma unroll
      for (int k = 0; k < kLcnBatch; k++)
        if (i0 + k < nterms)
        {
          const double ck = s_c[i0 + k];
#pragma unroll
          for (int w = 0; w < W; w++)
          {
            const double pr = ck * v[k][w];
            acc[w]          = (i0 + k == 0) ? pr : acc[w] + pr;
          }
        }
    }
    stg<W>(z + base, acc);
#pragma unroll
    for (int w = 0; w < W; w++) sq[w] += acc[w] * acc[w];
  }
  const int64_t tail0 = nfull * TILE;
  if (tail0 < n && blockIdx.x == (unsigned)(nfull % gridDim.x))
  {
    for (int64_t i = tail0 + threadIdx.x; i < n; i += kBlock)
    {
      double acc = s_c[0] * s_x[0][i];
      for (int k = 1; k < nterms; k++) acc += s_c[k] * s_x[k][i];
      z[i] = acc;
      sq[0] += acc * acc;
    }
  }
  double v = sq[0];
#pragma unroll
  for (int w = 1; w < W; w++) v += sq[w];
  v = block_combine<CombSum, kBlock>(v, smem);
  finish_block<CombSum, kBlock, false>(v, tagged, o, x, smem, nullptr);
}

/* exact-order form (n <= exact_threshold): z as above, then thread 0 adds z_0^2, z_1^2, ...
   strictly left to right -- the bits of serial's N_VLinearCombination + N_VDotProd */
__global__ void __launch_bounds__(kBlock) k_lincomb_sqnorm_exact(const __grid_constant__ LcNormArgs a, int n, ResOut o, const __grid_constant__ XArgs x)
{
  __shared__ double buf[kExactMaxElems];
  __shared__ double s_c[kLcnMaxTerms];
  pdl_prologue();
  if (threadIdx.x < a.nterms)
    s_c[threadIdx.x] = a.c_dev ? (threadIdx.x == 0 ? 1.0 : -__ldcg(a.c_dev + threadIdx.x - 1)) : a.c[threadIdx.x];
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += kBlock)
  {
    double acc = s_c[0] * a.X[0][i];
    for (int k = 1; k < a.nterms; k++) acc += s_c[k] * a.X[k][i];
    a.z[i] = acc;
    buf[i] = acc * acc;
  }
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
  {
#pragma unroll 8
    for (int i = 0; i < n; i++) s += buf[i];
  }
  combine_and_publish<CombSum>(s, o, x);
}

/* launch k_lincomb_sqnorm (exact-order or tree form); c_dev != NULL: device-resident coefficients */
static void launch_lincomb_sqnorm(b200vec_ctx ctx, int nvec, const double* c, const double* c_dev, const double* const* X,
                                  double* z, int64_t n, const ResOut& out, const XArgs& xa)
{
  LcNormArgs a;
  int wmax = align_width(z);
  for (int i = 0; i < nvec; i++)
  {
    a.X[i] = X[i];
    a.c[i] = c ? c[i] : 0.0;
    wmax   = min(wmax, align_width(X[i]));
  }
  a.c_dev  = c_dev;
  a.z      = z;
  a.nterms = nvec;
  if (n <= ctx->tune.exact_threshold) launch_k(ctx, k_lincomb_sqnorm_exact, dim3(1), dim3(kBlock), a, (int)n, out, xa);
  else
  {
    int W = wmax;
    if (ctx->tune.vec_width > 0 && ctx->tune.vec_width < W) W = (int)ctx->tune.vec_width;
    int64_t tiles = n / ((int64_t)kBlock * W);
    if (tiles < 1) tiles = 1;
    int64_t cap = (int64_t)kSMs * 4; /* 4 CTAs of 256 threads per SM (64 registers): one resident wave */
    if (ctx->tune.max_blocks != kMaxBlocksDef && cap > ctx->tune.max_blocks) cap = ctx->tune.max_blocks;
    const int grid = (int)((tiles < cap) ? tiles : cap);
    if (W == 4) launch_k(ctx, k_lincomb_sqnorm<4>, dim3(grid), dim3(kBlock), a, n, ctx->d_tagged, out, xa);
    else if (W == 2) launch_k(ctx, k_lincomb_sqnorm<2>, dim3(grid), dim3(kBlock), a, n, ctx->d_tagged, out, xa);
    else launch_k(ctx, k_lincomb_sqnorm<1>, dim3(grid), dim3(kBlock), a, n, ctx->d_tagged, out, xa);
  }
}

/* ------------------------------------------------------------------ chained Gram-Schmidt sweeps
 * All kernels of one orthogonalisation column are launched back to back: each one reads the
 * coefficient(s) it needs from the DEVICE result slot(s) its predecessor wrote (stream order +
 * griddepcontrol.wait make them visible), so the host does not sit between two kernels.  Every
 * kernel publishes its scalar(s) to its own pinned slot under ONE shared host tag; the host waits
 * once, at the end, for all of them.  One host round trip per column instead of k + 1 (modified) or
 * 2 (classical), and programmatic dependent launch overlaps each kernel's launch ramp with its
 * predecessor's tail. */
struct Chain
{
  b200vec_ctx ctx;
  int scope;          /* 0 local, 1 peer-memory exchange inside each kernel, 2 ncclAllReduce after each kernel */
  XArgs xa;
  bool first = true;
  long long htag = -1;
  bool to_host;

  Chain(b200vec_ctx c, bool want_host) : ctx(c)
  {
    scope   = take_scope(c, &xa);
    to_host = want_host && scope != 2;
  }
  /* bookkeeping before ANY launch of the chain; returns the chain's host tag (the sequence number of
     its first launch).  For launchers that take their own sequence number (the multi-dot groups). */
  long long advance()
  {
    if (!first && scope == 1) next_xargs(ctx, &xa); /* every kernel is its own collective */
    if (first) htag = (long long)(unsigned int)(ctx->seq + 1);
    first = false;
    return htag;
  }
  /* ResOut of the next kernel of the chain: own partial tag, shared host tag, result in `slot` */
  ResOut next(int slot)
  {
    const long long t = advance();
    ResOut out        = next_out(ctx, slot, to_host);
    out.htag          = (unsigned int)t;
    return out;
  }
  /* NCCL transport: fold the slots the kernel just wrote before the next kernel reads them */
  int after(int slot, int count)
  {
    if (scope != 2) return B200VEC_OK;
    return b200vec_allreduce_buffer(ctx, ctx->d_result + slot, count, B200VEC_SUM);
  }
  int finish(int count, double* res)
  {
    if (scope == 2) return b200vec_result_fetch(ctx, count, res);
    return finish_host(ctx, 0, count, res, htag);
  }
};

} // namespace b200

using namespace b200;

/* an argument error must not leave the one-shot global scope armed for the next (local) reduction */
#define B200_RARGS(cond)                                                               \
  B200_CHECK_CTX(ctx);                                                                 \
  if (n < 0 || (n > 0 && !(cond)))                                                     \
  {                                                                                    \
    ctx->scope_global = false;                                                         \
    return set_error(B200VEC_ERR_ARG, "%s: bad argument", __func__);                   \
  }                                                                                    \
  (void)0

extern "C" {

int b200vec_dot_prod(b200vec_ctx ctx, const double* x, const double* y, int64_t n, double* result_host)
{
  B200_RARGS(x && y);
  if (x == y) return launch_reduce(ctx, "dot_prod(self)", RSqr{}, RedPtrs{x, nullptr, nullptr, nullptr}, n, 0.0, result_host);
  return launch_reduce(ctx, "dot_prod", RDot{}, RedPtrs{x, y, nullptr, nullptr}, n, 0.0, result_host);
}

int b200vec_max_norm(b200vec_ctx ctx, const double* x, int64_t n, double* result_host)
{
  B200_RARGS(x);
  return launch_reduce(ctx, "max_norm", RMaxNorm{}, RedPtrs{x, nullptr, nullptr, nullptr}, n, 0.0, result_host);
}

int b200vec_min(b200vec_ctx ctx, const double* x, int64_t n, double* result_host)
{
  B200_RARGS(x);
  return launch_reduce(ctx, "min", RMin{}, RedPtrs{x, nullptr, nullptr, nullptr}, n, DBL_MAX, result_host);
}

int b200vec_l1_norm(b200vec_ctx ctx, const double* x, int64_t n, double* result_host)
{
  B200_RARGS(x);
  return launch_reduce(ctx, "l1_norm", RL1{}, RedPtrs{x, nullptr, nullptr, nullptr}, n, 0.0, result_host);
}

int b200vec_wsqr_sum(b200vec_ctx ctx, const double* x, const double* w, int64_t n, double* result_host)
{
  B200_RARGS(x && w);
  return launch_reduce(ctx, "wsqr_sum", RWSqr{}, RedPtrs{x, w, nullptr, nullptr}, n, 0.0, result_host);
}

int b200vec_wsqr_sum_mask(b200vec_ctx ctx, const double* x, const double* w, const double* id, int64_t n,
                          double* result_host)
{
  B200_RARGS(x && w && id);
  return launch_reduce(ctx, "wsqr_sum_mask", RWSqrMask{}, RedPtrs{x, w, id, nullptr}, n, 0.0, result_host);
}

int b200vec_inv_test(b200vec_ctx ctx, const double* x, double* z, int64_t n, double* result_host)
{
  B200_RARGS(x && z);
  return launch_reduce(ctx, "inv_test", RInvTest{}, RedPtrs{x, nullptr, nullptr, z}, n, 1.0, result_host);
}

int b200vec_constr_mask(b200vec_ctx ctx, const double* c, const double* x, double* m, int64_t n, double* result_host)
{
  B200_RARGS(c && x && m);
  return launch_reduce(ctx, "constr_mask", RConstrMask{}, RedPtrs{c, x, nullptr, m}, n, 1.0, result_host);
}

int b200vec_min_quotient(b200vec_ctx ctx, const double* num, const double* denom, int64_t n, double* result_host)
{
  B200_RARGS(num && denom);
  return launch_reduce(ctx, "min_quotient", RMinQuot{}, RedPtrs{num, denom, nullptr, nullptr}, n, DBL_MAX,
                       result_host);
}

int b200vec_ewt_set(b200vec_ctx ctx, double rtol, double atol, const double* atol_vec, const double* y, double* w,
                    int64_t n, double* min_denominator_host)
{
  B200_RARGS(y && w);
  if (atol_vec)
    return launch_reduce(ctx, "ewt_set(vector atol)", REwt<true>{rtol, 0.0}, RedPtrs{y, atol_vec, nullptr, w}, n, DBL_MAX,
                         min_denominator_host);
  return launch_reduce(ctx, "ewt_set", REwt<false>{rtol, atol}, RedPtrs{y, nullptr, nullptr, w}, n, DBL_MAX,
                       min_denominator_host);
}

int b200vec_axpy_dot(b200vec_ctx ctx, double a, const double* x, double* z, const double* w, int64_t n,
                     double* result_host)
{
  B200_RARGS(x && z && w);
  return launch_reduce(ctx, "axpy_dot", RAxpyDot{a, nullptr}, RedPtrs{x, z, w, z}, n, 0.0, result_host);
}

int b200vec_dot_prod_multi(b200vec_ctx ctx, int nvec, const double* x, const double* const* Y, int64_t n,
                           double* result_host)
{
  B200_CHECK_CTX(ctx);
  if (nvec < 1 || n < 0 || !Y || (n > 0 && !x))
  {
    ctx->scope_global = false;
    return set_error(B200VEC_ERR_ARG, "%s: bad argument", __func__);
  }
  /* nvec == 1 is N_VDotProd in the reference (serial:1007-1012): same kernel family */
  if (nvec == 1) return b200vec_dot_prod(ctx, x, Y[0], n, result_host);
  return launch_multi<0>(ctx, "dot_prod_multi", nvec, x, Y, nullptr, n, result_host);
}

int b200vec_wsqr_sum_vector_array(b200vec_ctx ctx, int nvec, const double* const* X, const double* const* W,
                                  const double* id, int64_t n, double* result_host)
{
  B200_CHECK_CTX(ctx);
  if (nvec < 1 || n < 0 || !X || !W)
  {
    ctx->scope_global = false;
    return set_error(B200VEC_ERR_ARG, "%s: bad argument", __func__);
  }
  if (nvec == 1)
    return id ? b200vec_wsqr_sum_mask(ctx, X[0], W[0], id, n, result_host)
              : b200vec_wsqr_sum(ctx, X[0], W[0], n, result_host);
  if (id) return launch_multi<2>(ctx, "wsqr_sum_mask_vector_array", nvec, id, X, W, n, result_host);
  return launch_multi<1>(ctx, "wsqr_sum_vector_array", nvec, nullptr, X, W, n, result_host);
}

/* z = sum c_j X_j and result = sum z^2 in one pass (see k_lincomb_sqnorm).  Falls back to the two
   separate operations where the reference's N_VLinearCombination would pick a special algebraic
   form (nvec == 1, and the +-1 / a == +-b forms of N_VLinearSum for nvec == 2 other than the
   in-place axpy), so z is always bit-identical to serial:871-942. */
int b200vec_linear_combination_sqnorm(b200vec_ctx ctx, int nvec, const double* c, const double* const* X, double* z,
                                      int64_t n, double* result_host)
{
  B200_CHECK_CTX(ctx);
  if (nvec < 1 || n < 0 || !c || !X || (n > 0 && !z))
  {
    ctx->scope_global = false;
    return set_error(B200VEC_ERR_ARG, "%s: bad argument", __func__);
  }
  bool fused = nvec >= 3 && nvec <= kLcnMaxTerms;
  if (nvec == 2)
  {
    const double a = c[0], b = c[1];
    const bool axpy    = (a == 1.0 && X[0] == z);                      /* (b*x1) + z == (1*z) + b*x1 */
    const bool special = (a == 1.0 || a == -1.0 || b == 1.0 || b == -1.0 || a == b || a == -b);
    fused              = axpy ? (b != 1.0 && b != -1.0) : !special;
  }
  if (!fused)
  {
    const bool global = ctx->scope_global; /* the scope belongs to the reduction, not to the streaming op */
    ctx->scope_global = false;
    int rc            = b200vec_linear_combination(ctx, nvec, c, X, z, n);
    if (rc) return rc;
    ctx->scope_global = global;
    return b200vec_dot_prod(ctx, z, z, n, result_host);
  }
  DeviceGuard g(ctx->device);
  XArgs xa;
  const int scope = take_scope(ctx, &xa);
  if (n == 0 && scope == 0)
  {
    if (result_host) *result_host = 0.0;
    return B200VEC_OK;
  }
  const bool to_host = (result_host != nullptr) && scope != 2;
  const ResOut out   = next_out(ctx, 0, to_host);
  launch_lincomb_sqnorm(ctx, nvec, c, nullptr, X, z, n, out, xa);
  int rc = check_launch(ctx, "linear_combination_sqnorm");
  if (rc) return rc;
  if (scope == 2) return finish_global_nccl(ctx, 1, B200VEC_SUM, result_host);
  return finish_host(ctx, 0, 1, result_host);
}

/* One column of MODIFIED Gram-Schmidt (sundials_iterative.c:45-80) as a chain of nproj + 1 kernels and
   ONE host wait:  res[0] = v_k . v_k,  h_0 = V_0 . v_k          (2-wide multi-dot, v_k read once)
                   v_k <- v_k - h_i V_i,  h_{i+1} = V_{i+1} . v_k  (k_reduce<RAxpyDot>, h_i read on the device)
                   v_k <- v_k - h_{last} V_last,  ||v_k||^2        (k_lincomb_sqnorm, device coefficient)
   h_host[nproj]; norms_host[0] = v_k . v_k before, norms_host[1] = after. */
int b200vec_mgs_sweep(b200vec_ctx ctx, int nproj, double* vk, const double* const* V, int64_t n, double* h_host,
                      double* norms_host)
{
  B200_CHECK_CTX(ctx);
  if (nproj < 1 || nproj + 2 > kMaxRows || n < 0 || !V || !h_host || !norms_host || (n > 0 && !vk))
  {
    ctx->scope_global = false;
    return set_error(B200VEC_ERR_ARG, "%s: bad argument", __func__);
  }
  DeviceGuard g(ctx->device);
  Chain ch(ctx, true);
  if (n == 0 && ch.scope == 0)
  {
    for (int i = 0; i < nproj; i++) h_host[i] = 0.0;
    norms_host[0] = norms_host[1] = 0.0;
    return B200VEC_OK;
  }
  int rc;
  {
    MultiArgs m;
    m.shared = vk;
    m.nout   = 2;
    m.self_j = 0;
    for (int j = 0; j < kMaxOut; j++) m.A[j] = nullptr;
    for (int j = 0; j < kMaxPair; j++) m.B[j] = nullptr;
    m.A[0] = vk;
    m.A[1] = V[0];
    const long long t = ch.advance();
    rc = launch_multi_group<0>(ctx, "mgs_sweep(norm, h0)", m, n, 0, ch.to_host, ch.xa, t);
    if (!rc) rc = ch.after(0, 2);
    if (rc) return rc;
  }
  for (int i = 0; i + 1 < nproj; i++)
  {
    const ResOut out = ch.next(2 + i);
    rc = launch_reduce_kernel(ctx, "mgs_sweep(axpy_dot)", RAxpyDot{0.0, ctx->d_result + 1 + i},
                              RedPtrs{V[i], vk, V[i + 1], vk}, n, out, ch.xa);
    if (!rc) rc = ch.after(2 + i, 1);
    if (rc) return rc;
  }
  {
    const double* X2[2] = {vk, V[nproj - 1]};
    const ResOut out    = ch.next(nproj + 1);
    launch_lincomb_sqnorm(ctx, 2, nullptr, ctx->d_result + nproj, X2, vk, n, out, ch.xa);
    rc = check_launch(ctx, "mgs_sweep(last update, norm)");
    if (!rc) rc = ch.after(nproj + 1, 1);
    if (rc) return rc;
  }
  double res[kMaxRows];
  rc = ch.finish(nproj + 2, res);
  if (rc) return rc;
  norms_host[0] = res[0];
  for (int i = 0; i < nproj; i++) h_host[i] = res[1 + i];
  norms_host[1] = res[nproj + 1];
  return B200VEC_OK;
}

/* One column of CLASSICAL Gram-Schmidt (sundials_iterative.c:130-146) as a chain of 2 kernels and ONE host wait:
     dots_j = x . Ydots_j, j < nvec          (multi-dot; x itself may be among the Ydots and is then read once)
     z <- Xcomb_0 - sum_j dots_{j-1} Xcomb_j, ||z||^2   (k_lincomb_sqnorm with device coefficients c_0 = 1,
                                                         c_j = -dots_{j-1}; nvec terms, z may alias Xcomb_0)
   dots_host[nvec]; *sqnorm_host = z . z. */
int b200vec_cgs_step(b200vec_ctx ctx, int nvec, const double* x, const double* const* Ydots, const double* const* Xcomb,
                     double* z, int64_t n, double* dots_host, double* sqnorm_host)
{
  B200_CHECK_CTX(ctx);
  if (nvec < 2 || nvec > kMaxOut || nvec > kLcnMaxTerms || nvec + 1 > kMaxRows || n < 0 || !Ydots || !Xcomb || !dots_host ||
      !sqnorm_host || (n > 0 && (!x || !z)))
  {
    ctx->scope_global = false;
    return set_error(B200VEC_ERR_ARG, "%s: bad argument", __func__);
  }
  DeviceGuard g(ctx->device);
  Chain ch(ctx, true);
  if (n == 0 && ch.scope == 0)
  {
    for (int i = 0; i < nvec; i++) dots_host[i] = 0.0;
    *sqnorm_host = 0.0;
    return B200VEC_OK;
  }
  int rc;
  /* the multi-dot, in groups only on the exact-order path (the staged terms of a group must fit shared
     memory: bit-identical results for n <= exact_threshold whatever nvec is) */
  int group = kMaxOut;
  if (n > 0 && n <= ctx->tune.exact_threshold)
  {
    int fit = (int)(kExactMaxElems / n);
    if (fit < 1) fit = 1;
    if (fit < group) group = fit;
  }
  for (int j0 = 0; j0 < nvec; j0 += group)
  {
    MultiArgs m;
    m.shared = x;
    m.nout   = (nvec - j0 < group) ? nvec - j0 : group;
    m.self_j = -1;
    for (int j = 0; j < kMaxOut; j++)
    {
      m.A[j] = (j < m.nout) ? Ydots[j0 + j] : nullptr;
      if (m.self_j < 0 && j < m.nout && m.A[j] == x) m.self_j = j;
    }
    for (int j = 0; j < kMaxPair; j++) m.B[j] = nullptr;
    const long long t = ch.advance();
    rc = launch_multi_group<0>(ctx, "cgs_step(dots)", m, n, j0, ch.to_host, ch.xa, t);
    if (!rc) rc = ch.after(j0, m.nout);
    if (rc) return rc;
  }
  {
    const ResOut out = ch.next(nvec);
    launch_lincomb_sqnorm(ctx, nvec, nullptr, ctx->d_result, Xcomb, z, n, out, ch.xa);
    rc = check_launch(ctx, "cgs_step(combination, norm)");
    if (!rc) rc = ch.after(nvec, 1);
    if (rc) return rc;
  }
  double res[kMaxRows];
  rc = ch.finish(nvec + 1, res);
  if (rc) return rc;
  for (int i = 0; i < nvec; i++) dots_host[i] = res[i];
  *sqnorm_host = res[nvec];
  return B200VEC_OK;
}

} /* extern "C" */
