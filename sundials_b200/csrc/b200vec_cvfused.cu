/* b200vec_cvfused.cu -- the integrator-level fused streaming kernels (SURVEY.md §8 row f-N3).
 *
 * The reference has a second plugin boundary next to the vector's ops table: libsundials_cvode links one
 * of libsundials_cvode_fused_{stubs,cuda,hip} (src/cvode/CMakeLists.txt:43-75), which all export the same
 * seven C functions (src/cvode/cvode_impl.h:639-672).  The stubs spell each of them as a sequence of
 * N_V* calls (src/cvode/cvode_fused_stubs.c:38-162: 2-10 ops, 6-28 array passes), the CUDA library as one
 * kernel each (src/cvode/cvode_fused_gpu.cpp:62-420).  This file holds the sm_100a kernels behind
 * libsundials_cvode_fused_b200.so (cvode_fused_b200.c): one launch per function, every operand read
 * once, every result written once, and -- unlike the reference's CUDA kernels, which are compiled with
 * FMA contraction and reorder two of the sequences -- the arithmetic is the STUBS' op sequence element by
 * element (including the case analysis of N_VLinearSum, nvector_serial.c:397-477), so the results are
 * bit-identical to the unfused path on nvector_serial.  Compiled with -fmad=false.
 *
 * One kernel template, k_mapn<W,U,NIN,NOUT,F>: the tile walk of k_map (b200vec_stream.cu) with up to
 * four inputs and four outputs.  An output may alias an input (the integrators update M, y, tempv in
 * place): every thread loads all its operands of a tile before it stores any.
 */
#include <cuda_runtime.h>

#include <cstdint>

#include "b200vec.h"
#include "b200vec_device.cuh"
#include "b200vec_internal.h"

namespace b200 {

constexpr int kMapnMax = 4;

struct MapnPtrs
{
  const double* in[kMapnMax];
  double* out[kMapnMax];
};

/* ------------------------------------------------------------ functors: in[] -> out[] for one element */

/* cvEwtSetSS / cvEwtSetSV (stubs:38-72): tempv = (rtol*|y|) + atol, weight = 1/tempv.  NOUT = 1 stops
   after tempv (the atolmin0 path tests min(tempv) before anything is inverted, as the stubs do). */
template <bool VEC>
struct FCvEwt
{
  double rtol, atol;
  __device__ void operator()(const double* in, double* out, int nout) const
  {
    const double t = (rtol * fabs(in[0])) + (VEC ? in[1] : atol);
    out[0]         = t;
    if (nout > 1) out[1] = 1.0 / t;
  }
};

/* cvCheckConstraints_fused (stubs:80-89): in = c, ewt, y, mm; out = tmp */
struct FCvConstraints
{
  __device__ void operator()(const double* in, double* out, int) const
  {
    double t = (fabs(in[0]) >= 1.5) ? 1.0 : 0.0; /* N_VCompare(1.5, c, tmp)                        */
    t        = t * in[0];                        /* N_VProd(tmp, c, tmp)                           */
    t        = t / in[1];                        /* N_VDiv(tmp, ewt, tmp)                          */
    t        = (-0.1 * t) + in[2];               /* N_VLinearSum(1, y, -0.1, tmp, tmp): VLin1 form */
    out[0]   = t * in[3];                        /* N_VProd(tmp, mm, tmp)                          */
  }
};

/* cvNlsResid_fused (stubs:97-104): in = zn1, ycor, ftemp; out = res */
struct FCvNlsResid
{
  double rl1, ngamma;
  __device__ void operator()(const double* in, double* out, int) const
  {
    const double t = (rl1 * in[0]) + in[1]; /* every N_VLinearSum form with b == 1 gives these bits */
    out[0]         = (ngamma * in[2]) + t;  /* Vaxpy: res += ngamma * ftemp                         */
  }
};

/* cvDiagSetup_formY (stubs:112-119): in = fpred, zn1, ypred; out = ftemp, y */
struct FCvDiagFormY
{
  double h, r;
  __device__ void operator()(const double* in, double* out, int) const
  {
    const double f = (h * in[0]) - in[1]; /* VLin2 (b == -1); h == +-1 give the same bits */
    out[0]         = f;
    out[1]         = (r * f) + in[2];
  }
};

/* cvDiagSetup_buildM (stubs:128-147): in = ftemp, fpred, ewt, M; out = bit, bitcomp, y, M.
   form = which branch of N_VLinearSum(FRACT, ftemp, -h, M, M) the scalars select (serial:397-477):
   0 general / axpy / VLin2 (all the same bits as (a x) + (b y)), 1 a == b, 2 a == -b. */
struct FCvDiagBuildM
{
  double uround, h;
  int form;
  __device__ void operator()(const double* in, double* out, int) const
  {
    const double fract = 0.1; /* FRACT of the stubs (cvode_fused_stubs.c:28) */
    const double ft    = in[0];
    double M           = in[3] - in[1]; /* N_VLinearSum(1, M, -1, fpred, M): M -= fpred */
    if (form == 1) M = fract * (ft + M);
    else if (form == 2) M = fract * (ft - M);
    else M = (fract * ft) + ((-h) * M);
    double y         = ft * in[2];                       /* N_VProd(ftemp, ewt, y)          */
    const double bit = (fabs(y) >= uround) ? 1.0 : 0.0;  /* N_VCompare(uround, y, bit)      */
    const double bc  = bit + (-1.0);                     /* N_VAddConst(bit, -1, bitcomp)   */
    y                = ft * bit;                         /* N_VProd(ftemp, bit, y)          */
    y                = (fract * y) - bc;                 /* N_VLinearSum(FRACT, y, -1, bitcomp, y): VLin2 */
    M                = M / y;                            /* N_VDiv(M, y, M)                 */
    M                = M * bit;                          /* N_VProd(M, bit, M)              */
    M                = M - bc;                           /* N_VLinearSum(1, M, -1, bitcomp, M): M -= bitcomp */
    out[0]           = bit;
    out[1]           = bc;
    out[2]           = y;
    out[3]           = M;
  }
};

/* cvDiagSolve_updateM (stubs:154-161): in = M; out = M */
struct FCvDiagUpdateM
{
  double r;
  __device__ void operator()(const double* in, double* out, int) const
  {
    double m = 1.0 / in[0];
    m        = m + (-1.0);
    m        = r * m;
    out[0]   = m + 1.0;
  }
};

/* --------------------------------------------------------------- kernel */
template <int W, int U, int NIN, int NOUT, class F>
__global__ void __launch_bounds__(kBlock) k_mapn(F f, const __grid_constant__ MapnPtrs p, int64_t n)
{
  constexpr int64_t TILE = (int64_t)kBlock * W * U;
  constexpr int64_t STEP = (int64_t)kBlock * W;
  const int64_t nfull    = n / TILE;
  pdl_prologue();

  for (int64_t t = blockIdx.x; t < nfull; t += gridDim.x)
  {
    const int64_t base = t * TILE + (int64_t)threadIdx.x * W;
    double a[U][NIN][W];
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int i = 0; i < NIN; i++) ldg<W>(p.in[i] + base + u * STEP, a[u][i]);
#pragma unroll
    for (int u = 0; u < U; u++)
    {
      double r[NOUT][W];
#pragma unroll
      for (int w = 0; w < W; w++)
      {
        double in[NIN], out[NOUT];
#pragma unroll
        for (int i = 0; i < NIN; i++) in[i] = a[u][i][w];
        f(in, out, NOUT);
#pragma unroll
        for (int o = 0; o < NOUT; o++) r[o][w] = out[o];
      }
#pragma unroll
      for (int o = 0; o < NOUT; o++) stg<W>(p.out[o] + base + u * STEP, r[o]);
    }
  }

  /* ragged tail (< TILE elements): scalar, by the least-loaded CTA */
  const int64_t tail0 = nfull * TILE;
  if (tail0 < n && blockIdx.x == (unsigned)(nfull % gridDim.x))
  {
    for (int64_t i = tail0 + threadIdx.x; i < n; i += kBlock)
    {
      double in[NIN], out[NOUT];
#pragma unroll
      for (int k = 0; k < NIN; k++) in[k] = p.in[k][i];
      f(in, out, NOUT);
#pragma unroll
      for (int o = 0; o < NOUT; o++) p.out[o][i] = out[o];
    }
  }
}

/* widest access every operand's alignment allows; two tiles in flight per thread with one or two inputs,
   one with three or four (64-128 B of loads in flight per thread, as k_map's two operands at U = 2) */
template <int NIN, int NOUT, class F>
static int launch_mapn(b200vec_ctx ctx, const char* name, F f, const MapnPtrs& p, int64_t n)
{
  if (n == 0) return B200VEC_OK;
  int wmax = 4;
  for (int i = 0; i < NIN; i++) wmax = min(wmax, align_width(p.in[i]));
  for (int o = 0; o < NOUT; o++) wmax = min(wmax, align_width(p.out[o]));
  MapCfg c = pick_map_cfg(ctx, n, wmax, false);
  constexpr int UMAX = (NIN <= 2) ? 2 : 1;
  if (c.U > UMAX)
  {
    c.U           = UMAX;
    int64_t tiles = n / ((int64_t)kBlock * c.W * c.U);
    int64_t cap   = ctx->tune.stream_max_blocks > 0 ? ctx->tune.stream_max_blocks : 0x7fffffff;
    if (tiles < 1) tiles = 1;
    c.grid = (int)(tiles < cap ? tiles : cap);
  }
  DeviceGuard g(ctx->device);
#define B200_MAPN_CASE(WW, UU) \
  if (c.W == WW && c.U == UU) launch_k(ctx, k_mapn<WW, UU, NIN, NOUT, F>, dim3(c.grid), dim3(kBlock), f, p, n)
  if (UMAX == 2)
  {
    B200_MAPN_CASE(4, 2);
    else B200_MAPN_CASE(4, 1);
    else B200_MAPN_CASE(2, 2);
    else B200_MAPN_CASE(2, 1);
    else B200_MAPN_CASE(1, 2);
    else B200_MAPN_CASE(1, 1);
  }
  else
  {
    B200_MAPN_CASE(4, 1);
    else B200_MAPN_CASE(2, 1);
    else B200_MAPN_CASE(1, 1);
  }
#undef B200_MAPN_CASE
  return check_launch(ctx, name);
}

} // namespace b200

using namespace b200;

#define B200_CVARGS(ok)                                                                        \
  B200_CHECK_CTX(ctx);                                                                         \
  if (n < 0 || (n > 0 && !(ok))) return set_error(B200VEC_ERR_ARG, "%s: bad argument", __func__)

extern "C" {

int b200vec_cv_ewt(b200vec_ctx ctx, double rtol, double atol, const double* atol_vec, const double* y, double* tempv,
                   double* weight, int64_t n)
{
  B200_CVARGS(y && tempv);
  MapnPtrs p = {{y, atol_vec, nullptr, nullptr}, {tempv, weight, nullptr, nullptr}};
  if (atol_vec)
  {
    if (weight) return launch_mapn<2, 2>(ctx, "cv_ewt(vector atol)", FCvEwt<true>{rtol, 0.0}, p, n);
    return launch_mapn<2, 1>(ctx, "cv_ewt_denominators(vector atol)", FCvEwt<true>{rtol, 0.0}, p, n);
  }
  if (weight) return launch_mapn<1, 2>(ctx, "cv_ewt", FCvEwt<false>{rtol, atol}, p, n);
  return launch_mapn<1, 1>(ctx, "cv_ewt_denominators", FCvEwt<false>{rtol, atol}, p, n);
}

int b200vec_cv_constraints(b200vec_ctx ctx, const double* c, const double* ewt, const double* y, const double* mm,
                           double* tmp, int64_t n)
{
  B200_CVARGS(c && ewt && y && mm && tmp);
  MapnPtrs p = {{c, ewt, y, mm}, {tmp, nullptr, nullptr, nullptr}};
  return launch_mapn<4, 1>(ctx, "cv_constraints", FCvConstraints{}, p, n);
}

int b200vec_cv_nls_resid(b200vec_ctx ctx, double rl1, double ngamma, const double* zn1, const double* ycor,
                         const double* ftemp, double* res, int64_t n)
{
  B200_CVARGS(zn1 && ycor && ftemp && res);
  MapnPtrs p = {{zn1, ycor, ftemp, nullptr}, {res, nullptr, nullptr, nullptr}};
  return launch_mapn<3, 1>(ctx, "cv_nls_resid", FCvNlsResid{rl1, ngamma}, p, n);
}

int b200vec_cv_diag_form_y(b200vec_ctx ctx, double h, double r, const double* fpred, const double* zn1,
                           const double* ypred, double* ftemp, double* y, int64_t n)
{
  B200_CVARGS(fpred && zn1 && ypred && ftemp && y);
  MapnPtrs p = {{fpred, zn1, ypred, nullptr}, {ftemp, y, nullptr, nullptr}};
  return launch_mapn<3, 2>(ctx, "cv_diag_form_y", FCvDiagFormY{h, r}, p, n);
}

int b200vec_cv_diag_build_m(b200vec_ctx ctx, double uround, double h, const double* ftemp, const double* fpred,
                            const double* ewt, double* bit, double* bitcomp, double* y, double* M, int64_t n)
{
  B200_CVARGS(ftemp && fpred && ewt && bit && bitcomp && y && M);
  const double a = 0.1, b = -h;
  /* the branch N_VLinearSum(FRACT, ftemp, -h, M, M) takes (serial:397-477, tested in that order): b == 1
     (axpy), b == -1 (VLin2) and the general form share their bits; a == b and a == -b do not */
  const int form = (b == 1.0 || b == -1.0) ? 0 : (a == b) ? 1 : (a == -b) ? 2 : 0;
  MapnPtrs p     = {{ftemp, fpred, ewt, M}, {bit, bitcomp, y, M}};
  return launch_mapn<4, 4>(ctx, "cv_diag_build_m", FCvDiagBuildM{uround, h, form}, p, n);
}

int b200vec_cv_diag_update_m(b200vec_ctx ctx, double r, double* M, int64_t n)
{
  B200_CVARGS(M);
  MapnPtrs p = {{M, nullptr, nullptr, nullptr}, {M, nullptr, nullptr, nullptr}};
  return launch_mapn<1, 1>(ctx, "cv_diag_update_m", FCvDiagUpdateM{r}, p, n);
}

} /* extern "C" */
