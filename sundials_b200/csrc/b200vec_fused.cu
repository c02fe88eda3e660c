/* b200vec_fused.cu -- fused multi-vector and vector-array kernels for sm_100a.
 *
 * All three kernel families are "row" kernels: blockIdx.y selects one output
 * row (one vector of a vector array), blockIdx.x walks that row's tiles
 * grid-stride.  Pointer tables and coefficients travel as by-value kernel
 * parameters (no H2D copy, no sync -- the reference packs them into a pinned
 * buffer, copies it and synchronises on EVERY fused call,
 * src/nvector/cuda/nvector_cuda.cu:2413-2604) and are staged into shared memory
 * once per CTA.
 *
 *  k_lincomb_rows   Z_r = sum_i c_i X[i][r]      each X read once, register
 *                   accumulator, z written once (the reference kernel
 *                   re-reads and re-writes z in global memory for every term,
 *                   VectorArrayKernels.cuh:42-51).  Accumulation order i = 0,1,..
 *                   is exactly the pass order of nvector_serial.c:871-942 ->
 *                   bit-identical.  Serves N_VLinearCombination (1 row) and
 *                   N_VLinearCombinationVectorArray.
 *  k_scaleadd_rows  Z[j][r] = a_j X_r + Y[j][r]  X_r read once for all j.
 *                   Serves N_VScaleAddMulti (1 row) and
 *                   N_VScaleAddMultiVectorArray.
 *  k_map_rows       per-row elementwise map with the k_map functors: serves
 *                   N_VLinearSumVectorArray (all 12 forms), N_VScaleVectorArray,
 *                   N_VConstVectorArray in ONE launch for all rows.
 *
 * Bytes per element: LinearCombination 8(nv+1); ScaleAddMulti 8(2nv+1);
 * LinearSumVA 24nv; ScaleVA 16nv; ConstVA 8nv; ScaleAddMultiVA 8(nv+2 nv ns);
 * LinearCombinationVA 8(nv ns + nv).
 */
#include "b200vec_device.cuh"

namespace b200 {

constexpr int kLcMaxTerms = 32; /* GMRES with maxl = 20 updates the solution with a 21-term combination
                                   (sunlinsol_spgmr.c:790,866): one launch, z written once */
constexpr int kLcMaxRows  = 16;
constexpr int kLcMaxPtrs  = 256; /* terms x rows per launch: the parameter struct stays under 4 KB */
constexpr int kTermBatch  = 4; /* loads of this many terms are in flight together */

struct LinCombArgs
{
  const double* X[kLcMaxPtrs]; /* X[i * nrows + r] */
  double* Z[kLcMaxRows];
  double c[kLcMaxTerms];
  int nterms;
  int nrows;
};

template <int W>
__global__ void __launch_bounds__(kBlock) k_lincomb_rows(const __grid_constant__ LinCombArgs a, int64_t n)
{
  __shared__ double s_c[kLcMaxTerms];
  __shared__ const double* s_x[kLcMaxTerms];
  const int row    = blockIdx.y;
  const int nterms = a.nterms;
  if (threadIdx.x < nterms)
  {
    s_c[threadIdx.x] = a.c[threadIdx.x];
    s_x[threadIdx.x] = a.X[threadIdx.x * a.nrows + row];
  }
  pdl_prologue(); /* parameters are staged before the wait on the previous kernel */
  __syncthreads();
  double* z = a.Z[row];

  constexpr int64_t TILE = (int64_t)kBlock * W;
  const int64_t nfull    = n / TILE;
  for (int64_t t = blockIdx.x; t < nfull; t += gridDim.x)
  {
    const int64_t base = t * TILE + (int64_t)threadIdx.x * W;
    double acc[W];
    for (int i0 = 0; i0 < nterms; i0 += kTermBatch)
    {
      double v[kTermBatch][W];
#pragma unroll
      for (int k = 0; k < kTermBatch; k++)
        if (i0 + k < nterms) ldg<W>(s_x[i0 + k] + base, v[k]);
#pragma unroll
      for (int k = 0; k < kTermBatch; k++)
        if (i0 + k < nterms)
        {
          const double ck = s_c[i0 + k];
#pragma unroll
          for (int w = 0; w < W; w++)
          {
            const double p = ck * v[k][w];
            acc[w]         = (i0 + k == 0) ? p : acc[w] + p;
          }
        }
    }
    stg<W>(z + base, acc);
  }

  const int64_t tail0 = nfull * TILE;
  if (tail0 < n && blockIdx.x == (unsigned)(nfull % gridDim.x))
  {
    for (int64_t i = tail0 + threadIdx.x; i < n; i += kBlock)
    {
      double acc = s_c[0] * s_x[0][i];
      for (int k = 1; k < nterms; k++) acc += s_c[k] * s_x[k][i];
      z[i] = acc;
    }
  }
}

constexpr int kSamMaxSums = 16;
constexpr int kSamMaxPtrs = 160;
constexpr int kSamMaxRows = 32;

struct ScaleAddArgs
{
  const double* X[kSamMaxRows];
  const double* Y[kSamMaxPtrs]; /* Y[j * nrows + r] */
  double* Z[kSamMaxPtrs];
  double a[kSamMaxSums];
  int nsum;
  int nrows;
};

template <int W>
__global__ void __launch_bounds__(kBlock) k_scaleadd_rows(const __grid_constant__ ScaleAddArgs a, int64_t n)
{
  __shared__ double s_a[kSamMaxSums];
  __shared__ const double* s_y[kSamMaxSums];
  __shared__ double* s_z[kSamMaxSums];
  const int row  = blockIdx.y;
  const int nsum = a.nsum;
  if (threadIdx.x < nsum)
  {
    s_a[threadIdx.x] = a.a[threadIdx.x];
    s_y[threadIdx.x] = a.Y[threadIdx.x * a.nrows + row];
    s_z[threadIdx.x] = a.Z[threadIdx.x * a.nrows + row];
  }
  pdl_prologue();
  __syncthreads();
  const double* x = a.X[row];

  constexpr int64_t TILE = (int64_t)kBlock * W;
  const int64_t nfull    = n / TILE;
  for (int64_t t = blockIdx.x; t < nfull; t += gridDim.x)
  {
    const int64_t base = t * TILE + (int64_t)threadIdx.x * W;
    double xv[W];
    ldg<W>(x + base, xv);
    for (int j0 = 0; j0 < nsum; j0 += kTermBatch)
    {
      double yv[kTermBatch][W];
#pragma unroll
      for (int k = 0; k < kTermBatch; k++)
        if (j0 + k < nsum) ldg<W>(s_y[j0 + k] + base, yv[k]);
#pragma unroll
      for (int k = 0; k < kTermBatch; k++)
        if (j0 + k < nsum)
        {
          const double ak = s_a[j0 + k];
          double r[W];
#pragma unroll
          for (int w = 0; w < W; w++) r[w] = ak * xv[w] + yv[k][w];
          stg<W>(s_z[j0 + k] + base, r);
        }
    }
  }

  const int64_t tail0 = nfull * TILE;
  if (tail0 < n && blockIdx.x == (unsigned)(nfull % gridDim.x))
  {
    for (int64_t i = tail0 + threadIdx.x; i < n; i += kBlock)
    {
      const double xi = x[i];
      for (int j = 0; j < nsum; j++) s_z[j][i] = s_a[j] * xi + s_y[j][i];
    }
  }
}

/* ---- per-row elementwise maps (vector-array forms of the k_map functors) ---- */
constexpr int kMapMaxRows = 64;

struct MapRowsArgs
{
  const double* p0[kMapMaxRows];
  const double* p1[kMapMaxRows];
  double* out[kMapMaxRows];
  double s[kMapMaxRows]; /* per-row scalar (ScaleVectorArray) */
};

enum RowForm
{
  RF_SUM,       /* p0 + p1           */
  RF_DIFF,      /* p0 - p1           */
  RF_LIN1,      /* a*p0 + p1         */
  RF_LIN2,      /* a*p0 - p1         */
  RF_SCALESUM,  /* a*(p0 + p1)       */
  RF_SCALEDIFF, /* a*(p0 - p1)       */
  RF_GENERAL,   /* a*p0 + b*p1       */
  RF_ROWSCALE,  /* s[row] * p0       */
  RF_CONST      /* a                 */
};

template <int FORM>
__device__ __forceinline__ double row_apply(double x, double y, double a, double b, double s)
{
  switch (FORM)
  {
  case RF_SUM: return x + y;
  case RF_DIFF: return x - y;
  case RF_LIN1: return (a * x) + y;
  case RF_LIN2: return (a * x) - y;
  case RF_SCALESUM: return a * (x + y);
  case RF_SCALEDIFF: return a * (x - y);
  case RF_GENERAL: return (a * x) + (b * y);
  case RF_ROWSCALE: return s * x;
  default: return a;
  }
}

template <int W, int U, int FORM>
__global__ void __launch_bounds__(kBlock)
  k_map_rows(const __grid_constant__ MapRowsArgs m, double a, double b, int64_t n)
{
  constexpr int NIN      = (FORM == RF_CONST) ? 0 : (FORM == RF_ROWSCALE) ? 1 : 2;
  constexpr int64_t TILE = (int64_t)kBlock * W * U;
  constexpr int64_t STEP = (int64_t)kBlock * W;
  const int row          = blockIdx.y;
  const double* p0       = m.p0[row];
  const double* p1       = m.p1[row];
  double* out            = m.out[row];
  const double s         = m.s[row];
  const int64_t nfull    = n / TILE;
  pdl_prologue();

  for (int64_t t = blockIdx.x; t < nfull; t += gridDim.x)
  {
    const int64_t base = t * TILE + (int64_t)threadIdx.x * W;
    double x[U][W], y[U][W];
#pragma unroll
    for (int u = 0; u < U; u++)
    {
      if (NIN >= 1) ldg<W>(p0 + base + u * STEP, x[u]);
      if (NIN >= 2) ldg<W>(p1 + base + u * STEP, y[u]);
    }
#pragma unroll
    for (int u = 0; u < U; u++)
    {
      double r[W];
#pragma unroll
      for (int w = 0; w < W; w++) r[w] = row_apply<FORM>(NIN >= 1 ? x[u][w] : 0.0, NIN >= 2 ? y[u][w] : 0.0, a, b, s);
      stg<W>(out + base + u * STEP, r);
    }
  }
  const int64_t tail0 = nfull * TILE;
  if (tail0 < n && blockIdx.x == (unsigned)(nfull % gridDim.x))
  {
    for (int64_t i = tail0 + threadIdx.x; i < n; i += kBlock)
      out[i] = row_apply<FORM>(NIN >= 1 ? p0[i] : 0.0, NIN >= 2 ? p1[i] : 0.0, a, b, s);
  }
}

/* ------------------------------------------------------------- launchers */

/* grid.x for a row kernel: spread max_blocks CTAs over the rows */
static int rows_grid_x(b200vec_ctx ctx, int64_t tiles, int nrows)
{
  if (tiles < 1) tiles = 1;
  if (ctx->tune.stream_max_blocks <= 0) return (int)((tiles < 0x7fffffff) ? tiles : 0x7fffffff);
  int64_t per_row = ctx->tune.stream_max_blocks / nrows;
  if (per_row < 1) per_row = 1;
  return (int)((tiles < per_row) ? tiles : per_row);
}

static int tuned_width(b200vec_ctx ctx, int wmax)
{
  if (ctx->tune.vec_width > 0 && ctx->tune.vec_width < wmax) return (int)ctx->tune.vec_width;
  return wmax;
}

/* X[i*ldx + r], i < nterms, r < nrows; z-accumulation across term chunks keeps
   the serial order: chunk k>0 starts from 1.0*z (exact) and continues adding */
static int lincomb_rows(b200vec_ctx ctx, int nterms, int nrows, const double* c, const double* const* X, int ldx,
                        double* const* Z, int64_t n)
{
  if (n == 0) return B200VEC_OK;
  DeviceGuard g(ctx->device);
  /* rows per launch: as many as the pointer table holds for this term count */
  const int nt_launch = (nterms < kLcMaxTerms) ? nterms : kLcMaxTerms;
  int rows_per        = kLcMaxPtrs / nt_launch;
  if (rows_per > kLcMaxRows) rows_per = kLcMaxRows;
  for (int r0 = 0; r0 < nrows; r0 += rows_per)
  {
    const int nr = (nrows - r0 < rows_per) ? nrows - r0 : rows_per;
    int i0       = 0;
    bool first   = true;
    while (i0 < nterms)
    {
      LinCombArgs a;
      int nt = 0;
      int wmax = 4;
      if (!first)
      { /* continue accumulating on top of what the previous chunk wrote */
        a.c[0] = 1.0;
        for (int r = 0; r < nr; r++) a.X[0 * nr + r] = Z[r0 + r];
        nt = 1;
      }
      while (nt < kLcMaxTerms && i0 < nterms)
      {
        a.c[nt] = c[i0];
        for (int r = 0; r < nr; r++) a.X[nt * nr + r] = X[(size_t)i0 * ldx + r0 + r];
        nt++;
        i0++;
      }
      for (int r = 0; r < nr; r++)
      {
        a.Z[r] = Z[r0 + r];
        wmax   = min(wmax, align_width(a.Z[r]));
      }
      for (int k = 0; k < nt * nr; k++) wmax = min(wmax, align_width(a.X[k]));
      a.nterms = nt;
      a.nrows  = nr;
      const int W = tuned_width(ctx, wmax);
      dim3 grid(rows_grid_x(ctx, n / ((int64_t)kBlock * W), nr), nr);
      if (W == 4) launch_k(ctx, k_lincomb_rows<4>, grid, dim3(kBlock), a, n);
      else if (W == 2) launch_k(ctx, k_lincomb_rows<2>, grid, dim3(kBlock), a, n);
      else launch_k(ctx, k_lincomb_rows<1>, grid, dim3(kBlock), a, n);
      int rc = check_launch(ctx, "linear_combination");
      if (rc) return rc;
      first = false;
    }
  }
  return B200VEC_OK;
}

/* Y/Z[j*ld + r] */
static int scaleadd_rows(b200vec_ctx ctx, int nsum, int nrows, const double* a_host, const double* const* X,
                         const double* const* Y, double* const* Z, int ld, int64_t n)
{
  if (n == 0) return B200VEC_OK;
  DeviceGuard g(ctx->device);
  for (int j0 = 0; j0 < nsum; j0 += kSamMaxSums)
  {
    const int ns = (nsum - j0 < kSamMaxSums) ? nsum - j0 : kSamMaxSums;
    int rows_per = kSamMaxPtrs / ns;
    if (rows_per > kSamMaxRows) rows_per = kSamMaxRows;
    for (int r0 = 0; r0 < nrows; r0 += rows_per)
    {
      const int nr = (nrows - r0 < rows_per) ? nrows - r0 : rows_per;
      ScaleAddArgs a;
      int wmax = 4;
      for (int r = 0; r < nr; r++)
      {
        a.X[r] = X[r0 + r];
        wmax   = min(wmax, align_width(a.X[r]));
      }
      for (int j = 0; j < ns; j++)
      {
        a.a[j] = a_host[j0 + j];
        for (int r = 0; r < nr; r++)
        {
          a.Y[j * nr + r] = Y[(size_t)(j0 + j) * ld + r0 + r];
          a.Z[j * nr + r] = Z[(size_t)(j0 + j) * ld + r0 + r];
          wmax            = min(wmax, min(align_width(a.Y[j * nr + r]), align_width(a.Z[j * nr + r])));
        }
      }
      a.nsum  = ns;
      a.nrows = nr;
      const int W = tuned_width(ctx, wmax);
      dim3 grid(rows_grid_x(ctx, n / ((int64_t)kBlock * W), nr), nr);
      if (W == 4) launch_k(ctx, k_scaleadd_rows<4>, grid, dim3(kBlock), a, n);
      else if (W == 2) launch_k(ctx, k_scaleadd_rows<2>, grid, dim3(kBlock), a, n);
      else launch_k(ctx, k_scaleadd_rows<1>, grid, dim3(kBlock), a, n);
      int rc = check_launch(ctx, "scale_add_multi");
      if (rc) return rc;
    }
  }
  return B200VEC_OK;
}

template <int FORM>
static int map_rows(b200vec_ctx ctx, const char* name, int nrows, const double* const* P0, const double* const* P1,
                    double* const* OUT, const double* s_host, double a, double b, int64_t n)
{
  if (n == 0) return B200VEC_OK;
  DeviceGuard g(ctx->device);
  for (int r0 = 0; r0 < nrows; r0 += kMapMaxRows)
  {
    const int nr = (nrows - r0 < kMapMaxRows) ? nrows - r0 : kMapMaxRows;
    MapRowsArgs m;
    int wmax = 4;
    for (int r = 0; r < nr; r++)
    {
      m.p0[r]  = P0 ? P0[r0 + r] : nullptr;
      m.p1[r]  = P1 ? P1[r0 + r] : nullptr;
      m.out[r] = OUT[r0 + r];
      m.s[r]   = s_host ? s_host[r0 + r] : 0.0;
      wmax     = min(wmax, min(align_width(m.p0[r]), min(align_width(m.p1[r]), align_width(m.out[r]))));
    }
    MapCfg c = pick_map_cfg(ctx, n, wmax, false);
    const int U = (c.U >= 4) ? 4 : 1;
    dim3 grid(rows_grid_x(ctx, n / ((int64_t)kBlock * c.W * U), nr), nr);
#define B200_ROWS_CASE(WW, UU) \
  if (c.W == WW && U == UU) launch_k(ctx, k_map_rows<WW, UU, FORM>, grid, dim3(kBlock), m, a, b, n)
    B200_ROWS_CASE(4, 4);
    else B200_ROWS_CASE(4, 1);
    else B200_ROWS_CASE(2, 4);
    else B200_ROWS_CASE(2, 1);
    else B200_ROWS_CASE(1, 4);
    else B200_ROWS_CASE(1, 1);
#undef B200_ROWS_CASE
    int rc = check_launch(ctx, name);
    if (rc) return rc;
  }
  return B200VEC_OK;
}

/* array form of Vaxpy (serial:1902-1940): ACC_r <- s*SRC_r + ACC_r */
static int axpy_rows(b200vec_ctx ctx, int nrows, double s, const double* const* SRC, double* const* ACC, int64_t n)
{
  const double* const* ACCc = (const double* const*)ACC;
  if (s == 1.0) return map_rows<RF_SUM>(ctx, "axpy_rows(+1)", nrows, ACCc, SRC, ACC, nullptr, 0, 0, n);
  if (s == -1.0) return map_rows<RF_DIFF>(ctx, "axpy_rows(-1)", nrows, ACCc, SRC, ACC, nullptr, 0, 0, n);
  return map_rows<RF_LIN1>(ctx, "axpy_rows", nrows, SRC, ACCc, ACC, nullptr, s, 0, n);
}

} // namespace b200

using namespace b200;

extern "C" {

int b200vec_linear_combination(b200vec_ctx ctx, int nvec, const double* c, const double* const* X, double* z,
                               int64_t n)
{
  B200_CHECK_CTX(ctx);
  if (nvec < 1 || n < 0 || !c || !X || (n > 0 && !z)) return set_error(B200VEC_ERR_ARG, "%s: bad argument", __func__);
  if (nvec == 1) return scale_dispatch(ctx, c[0], X[0], z, n);                                  /* serial:885-890 */
  if (nvec == 2) return linear_sum_dispatch(ctx, c[0], X[0], c[1], X[1], z, z == X[0], z == X[1], n); /* serial:893-898 */
  /* serial:907-941: the three variants (X[0]==z with c0==1, X[0]==z, general)
     all evaluate ((c0*x0 + c1*x1) + c2*x2) + ... per element (1.0*z == z) */
  double* Z[1] = {z};
  return lincomb_rows(ctx, nvec, 1, c, X, 1, Z, n);
}

int b200vec_scale_add_multi(b200vec_ctx ctx, int nvec, const double* a, const double* x, const double* const* Y,
                            double* const* Z, int64_t n)
{
  B200_CHECK_CTX(ctx);
  if (nvec < 1 || n < 0 || !a || !Y || !Z || (n > 0 && !x))
    return set_error(B200VEC_ERR_ARG, "%s: bad argument", __func__);
  if (nvec == 1) /* serial:958-963 */
    return linear_sum_dispatch(ctx, a[0], x, 1.0, Y[0], Z[0], Z[0] == x, Z[0] == Y[0], n);
  const double* X[1] = {x};
  return scaleadd_rows(ctx, nvec, 1, a, X, Y, Z, 1, n);
}

int b200vec_linear_sum_vector_array(b200vec_ctx ctx, int nvec, double a, const double* const* X, double b,
                                    const double* const* Y, double* const* Z, int z_is_x, int z_is_y, int64_t n)
{
  B200_CHECK_CTX(ctx);
  if (nvec < 1 || n < 0 || !X || !Y || !Z) return set_error(B200VEC_ERR_ARG, "%s: bad argument", __func__);
  if (nvec == 1) /* serial:1054-1059: handle identities of the single vectors */
    return linear_sum_dispatch(ctx, a, X[0], b, Y[0], Z[0], Z[0] == X[0], Z[0] == Y[0], n);
  /* serial:1062-1147 with ARRAY identities */
  if (b == 1.0 && z_is_y) return axpy_rows(ctx, nvec, a, X, Z, n);
  if (a == 1.0 && z_is_x) return axpy_rows(ctx, nvec, b, Y, Z, n);
  if (a == 1.0 && b == 1.0) return map_rows<RF_SUM>(ctx, "linear_sum_va(sum)", nvec, X, Y, Z, nullptr, 0, 0, n);
  if (a == 1.0 && b == -1.0) return map_rows<RF_DIFF>(ctx, "linear_sum_va(diff)", nvec, X, Y, Z, nullptr, 0, 0, n);
  if (a == -1.0 && b == 1.0) return map_rows<RF_DIFF>(ctx, "linear_sum_va(diff)", nvec, Y, X, Z, nullptr, 0, 0, n);
  if (a == 1.0) return map_rows<RF_LIN1>(ctx, "linear_sum_va(lin1)", nvec, Y, X, Z, nullptr, b, 0, n);
  if (b == 1.0) return map_rows<RF_LIN1>(ctx, "linear_sum_va(lin1)", nvec, X, Y, Z, nullptr, a, 0, n);
  if (a == -1.0) return map_rows<RF_LIN2>(ctx, "linear_sum_va(lin2)", nvec, Y, X, Z, nullptr, b, 0, n);
  if (b == -1.0) return map_rows<RF_LIN2>(ctx, "linear_sum_va(lin2)", nvec, X, Y, Z, nullptr, a, 0, n);
  if (a == b) return map_rows<RF_SCALESUM>(ctx, "linear_sum_va(scalesum)", nvec, X, Y, Z, nullptr, a, 0, n);
  if (a == -b) return map_rows<RF_SCALEDIFF>(ctx, "linear_sum_va(scalediff)", nvec, X, Y, Z, nullptr, a, 0, n);
  return map_rows<RF_GENERAL>(ctx, "linear_sum_va(general)", nvec, X, Y, Z, nullptr, a, b, n);
}

int b200vec_scale_vector_array(b200vec_ctx ctx, int nvec, const double* c, const double* const* X, double* const* Z,
                               int64_t n)
{
  B200_CHECK_CTX(ctx);
  if (nvec < 1 || n < 0 || !c || !X || !Z) return set_error(B200VEC_ERR_ARG, "%s: bad argument", __func__);
  if (nvec == 1) return scale_dispatch(ctx, c[0], X[0], Z[0], n); /* serial:1166-1171 */
  /* serial:1179-1197: always c_i * x (no +-1 shortcuts for nvec > 1) */
  return map_rows<RF_ROWSCALE>(ctx, "scale_vector_array", nvec, X, nullptr, Z, c, 0, 0, n);
}

int b200vec_const_vector_array(b200vec_ctx ctx, int nvec, double c, double* const* Z, int64_t n)
{
  B200_CHECK_CTX(ctx);
  if (nvec < 1 || n < 0 || !Z) return set_error(B200VEC_ERR_ARG, "%s: bad argument", __func__);
  return map_rows<RF_CONST>(ctx, "const_vector_array", nvec, nullptr, nullptr, Z, nullptr, c, 0, n);
}

int b200vec_scale_add_multi_vector_array(b200vec_ctx ctx, int nvec, int nsum, const double* a,
                                         const double* const* X, const double* const* Y, double* const* Z,
                                         int y_is_z, int64_t n)
{
  B200_CHECK_CTX(ctx);
  if (nvec < 1 || nsum < 1 || n < 0 || !a || !X || !Y || !Z)
    return set_error(B200VEC_ERR_ARG, "%s: bad argument", __func__);
  if (nvec == 1)
  {
    if (nsum == 1) /* serial:1332-1337 */
      return linear_sum_dispatch(ctx, a[0], X[0], 1.0, Y[0], Z[0], Z[0] == X[0], Z[0] == Y[0], n);
    /* serial:1340-1351: N_VScaleAddMulti on gathered handle arrays */
    return b200vec_scale_add_multi(ctx, nsum, a, X[0], Y, Z, n);
  }
  if (nsum == 1) /* serial:1364-1368; row arrays Y[0], Z[0] alias iff the caller's did */
    return b200vec_linear_sum_vector_array(ctx, nvec, a[0], X, 1.0, Y, Z, 0, y_is_z, n);
  /* serial:1380-1406: in-place and out-of-place evaluate a_j*x + y identically */
  return scaleadd_rows(ctx, nsum, nvec, a, X, Y, Z, nvec, n);
}

int b200vec_linear_combination_vector_array(b200vec_ctx ctx, int nvec, int nsum, const double* c,
                                            const double* const* X, double* const* Z, int x0_is_z, int64_t n)
{
  B200_CHECK_CTX(ctx);
  if (nvec < 1 || nsum < 1 || n < 0 || !c || !X || !Z) return set_error(B200VEC_ERR_ARG, "%s: bad argument", __func__);
  if (nvec == 1)
  {
    if (nsum == 1) return scale_dispatch(ctx, c[0], X[0], Z[0], n);                               /* serial:1434-1439 */
    if (nsum == 2)                                                                                /* serial:1442-1447 */
      return linear_sum_dispatch(ctx, c[0], X[0], c[1], X[1], Z[0], Z[0] == X[0], Z[0] == X[1], n);
    return b200vec_linear_combination(ctx, nsum, c, X, Z[0], n);                                  /* serial:1450-1459 */
  }
  if (nsum == 1)
  { /* serial:1467-1477: ScaleVectorArray with c[0] replicated */
    double ctmp[kMapMaxRows];
    for (int r0 = 0; r0 < nvec; r0 += kMapMaxRows)
    {
      const int nr = (nvec - r0 < kMapMaxRows) ? nvec - r0 : kMapMaxRows;
      for (int r = 0; r < nr; r++) ctmp[r] = c[0];
      int rc = b200vec_scale_vector_array(ctx, nr, ctmp, X + r0, Z + r0, n);
      if (rc) return rc;
    }
    return B200VEC_OK;
  }
  if (nsum == 2) /* serial:1481-1485: Z == X[0] array identity decides the axpy form */
    return b200vec_linear_sum_vector_array(ctx, nvec, c[0], X, c[1], X + nvec, Z, x0_is_z, 0, n);
  return lincomb_rows(ctx, nsum, nvec, c, X, nvec, Z, n);
}

} /* extern "C" */
