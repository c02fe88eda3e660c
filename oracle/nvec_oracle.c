/* nvec_oracle.c -- TEST INFRASTRUCTURE ONLY (parity oracle), never shipped.
 *
 * CPU restatement of the arithmetic of the reference's nvector_serial
 * (SUNDIALS 7.5.0, /root/reference/src/nvector/serial/nvector_serial.c; cited
 * below as "serial:<line>").  See nvec_oracle.h for scope and how this file is
 * pinned against the real reference.  Compile with -ffp-contract=off: the
 * reference CPU library contains no fused multiply-adds.
 *
 * What matters for bit-parity is (1) which algebraic *form* the reference picks
 * for given scalars / aliasing and (2) the order of floating-point operations
 * inside that form.  The form selection is written as small classifiers
 * (lsum_form, ...) so the tests can also query it directly.
 */
#include "nvec_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>

#define ORC_SUCCESS 0
#define ORC_ERR_ARG (-1)

/* sundials_math.h:83  SUNRsqrt(x) = x <= 0 ? 0 : sqrt(x) */
static double guarded_sqrt(double v) { return (v <= 0.0) ? 0.0 : sqrt(v); }

/* ---------------------------------------------------------------------------
 * N_VLinearSum   serial:387-480 and helpers serial:1616-1760
 * ------------------------------------------------------------------------- */
enum lsum_form
{
  LS_AXPY_INTO_Y, /* y <- a x + y  (b == 1, z is y)           serial:397-401 */
  LS_AXPY_INTO_X, /* x <- b y + x  (a == 1, z is x)           serial:403-407 */
  LS_SUM,         /* z = x + y                                serial:411-415 */
  LS_DIFF_YX,     /* z = y - x   (a == -1, b == 1)            serial:419-425 */
  LS_DIFF_XY,     /* z = x - y   (a == 1, b == -1)            serial:419-425 */
  LS_LIN1_BY_X,   /* z = b y + x (a == 1)                     serial:430-437 */
  LS_LIN1_AX_Y,   /* z = a x + y (b == 1)                     serial:430-437 */
  LS_LIN2_BY_X,   /* z = b y - x (a == -1)                    serial:441-448 */
  LS_LIN2_AX_Y,   /* z = a x - y (b == -1)                    serial:441-448 */
  LS_SCALE_SUM,   /* z = a (x + y) (a == b)                   serial:453-457 */
  LS_SCALE_DIFF,  /* z = a (x - y) (a == -b)                  serial:461-465 */
  LS_GENERAL      /* z = a x + b y                            serial:472-477 */
};

/* z_is_y / z_is_x stand for the reference's N_Vector handle comparisons */
static enum lsum_form lsum_classify(double a, double b, int z_is_x, int z_is_y)
{
  if (b == 1.0 && z_is_y) return LS_AXPY_INTO_Y;
  if (a == 1.0 && z_is_x) return LS_AXPY_INTO_X;
  if (a == 1.0 && b == 1.0) return LS_SUM;
  if (a == 1.0 && b == -1.0) return LS_DIFF_XY;
  if (a == -1.0 && b == 1.0) return LS_DIFF_YX;
  if (a == 1.0) return LS_LIN1_BY_X;
  if (b == 1.0) return LS_LIN1_AX_Y;
  if (a == -1.0) return LS_LIN2_BY_X;
  if (b == -1.0) return LS_LIN2_AX_Y;
  if (a == b) return LS_SCALE_SUM;
  if (a == -b) return LS_SCALE_DIFF;
  return LS_GENERAL;
}

/* serial:1734-1760  (Vaxpy): three sub-forms depending on the scalar */
static void axpy_inplace(double s, const double* src, double* acc, orc_index n)
{
  orc_index i;
  if (s == 1.0)
  {
    for (i = 0; i < n; i++) acc[i] += src[i];
  }
  else if (s == -1.0)
  {
    for (i = 0; i < n; i++) acc[i] -= src[i];
  }
  else
  {
    for (i = 0; i < n; i++) acc[i] += s * src[i];
  }
}

static void lsum_apply(enum lsum_form f, double a, const double* x, double b, const double* y,
                       double* z, orc_index n)
{
  orc_index i;
  switch (f)
  {
  case LS_AXPY_INTO_Y: axpy_inplace(a, x, z, n); break;
  case LS_AXPY_INTO_X: axpy_inplace(b, y, z, n); break;
  case LS_SUM: /* serial:1628 */
    for (i = 0; i < n; i++) z[i] = x[i] + y[i];
    break;
  case LS_DIFF_XY: /* serial:1645 with (v2,v1) = (x,y) */
    for (i = 0; i < n; i++) z[i] = x[i] - y[i];
    break;
  case LS_DIFF_YX: /* serial:1645 with (v2,v1) = (y,x) */
    for (i = 0; i < n; i++) z[i] = y[i] - x[i];
    break;
  case LS_LIN1_BY_X: /* serial:1712 */
    for (i = 0; i < n; i++) z[i] = (b * y[i]) + x[i];
    break;
  case LS_LIN1_AX_Y:
    for (i = 0; i < n; i++) z[i] = (a * x[i]) + y[i];
    break;
  case LS_LIN2_BY_X: /* serial:1729 */
    for (i = 0; i < n; i++) z[i] = (b * y[i]) - x[i];
    break;
  case LS_LIN2_AX_Y:
    for (i = 0; i < n; i++) z[i] = (a * x[i]) - y[i];
    break;
  case LS_SCALE_SUM: /* serial:1678 */
    for (i = 0; i < n; i++) z[i] = a * (x[i] + y[i]);
    break;
  case LS_SCALE_DIFF: /* serial:1695 */
    for (i = 0; i < n; i++) z[i] = a * (x[i] - y[i]);
    break;
  case LS_GENERAL: /* serial:477 */
    for (i = 0; i < n; i++) z[i] = (a * x[i]) + (b * y[i]);
    break;
  }
}

void orc_linear_sum(double a, const double* x, double b, const double* y, double* z, orc_index n)
{
  lsum_apply(lsum_classify(a, b, z == x, z == y), a, x, b, y, z, n);
}

/* exported for tests: which form would the reference take? */
int orc_linear_sum_form(double a, double b, int z_is_x, int z_is_y)
{
  return (int)lsum_classify(a, b, z_is_x, z_is_y);
}

/* ---------------------------------------------------------------------------
 * simple streaming ops
 * ------------------------------------------------------------------------- */
void orc_const(double c, double* z, orc_index n) /* serial:482-495 */
{
  for (orc_index i = 0; i < n; i++) z[i] = c;
}

void orc_prod(const double* x, const double* y, double* z, orc_index n) /* serial:497-512 */
{
  for (orc_index i = 0; i < n; i++) z[i] = x[i] * y[i];
}

void orc_div(const double* x, const double* y, double* z, orc_index n) /* serial:514-529 */
{
  for (orc_index i = 0; i < n; i++) z[i] = x[i] / y[i];
}

/* serial:531-555: in place -> x *= c (serial:1772); c==1 copy; c==-1 negate */
void orc_scale(double c, const double* x, double* z, orc_index n)
{
  orc_index i;
  if (z == x)
  {
    for (i = 0; i < n; i++) z[i] *= c;
  }
  else if (c == 1.0)
  {
    for (i = 0; i < n; i++) z[i] = x[i];
  }
  else if (c == -1.0)
  {
    for (i = 0; i < n; i++) z[i] = -x[i];
  }
  else
  {
    for (i = 0; i < n; i++) z[i] = c * x[i];
  }
}

void orc_abs(const double* x, double* z, orc_index n) /* serial:557-571 */
{
  for (orc_index i = 0; i < n; i++) z[i] = fabs(x[i]);
}

void orc_inv(const double* x, double* z, orc_index n) /* serial:573-587 */
{
  for (orc_index i = 0; i < n; i++) z[i] = 1.0 / x[i];
}

void orc_add_const(const double* x, double b, double* z, orc_index n) /* serial:589-603 */
{
  for (orc_index i = 0; i < n; i++) z[i] = x[i] + b;
}

void orc_compare(double c, const double* x, double* z, orc_index n) /* serial:762-776 */
{
  for (orc_index i = 0; i < n; i++) z[i] = (fabs(x[i]) >= c) ? 1.0 : 0.0;
}

/* ---------------------------------------------------------------------------
 * reductions (strictly left-to-right accumulation, as the reference)
 * ------------------------------------------------------------------------- */
double orc_dot_prod(const double* x, const double* y, orc_index n) /* serial:605-620 */
{
  double acc = 0.0;
  for (orc_index i = 0; i < n; i++) acc += x[i] * y[i];
  return acc;
}

double orc_max_norm(const double* x, orc_index n) /* serial:622-639 */
{
  double best = 0.0;
  for (orc_index i = 0; i < n; i++)
  {
    double m = fabs(x[i]);
    if (m > best) best = m; /* strict '>' : a NaN never replaces the max */
  }
  return best;
}

double orc_wsqr_sum(const double* x, const double* w, orc_index n) /* serial:650-669 */
{
  double acc = 0.0;
  for (orc_index i = 0; i < n; i++)
  {
    double p = x[i] * w[i];
    acc += p * p;
  }
  return acc;
}

double orc_wsqr_sum_mask(const double* x, const double* w, const double* id,
                         orc_index n) /* serial:680-703 */
{
  double acc = 0.0;
  for (orc_index i = 0; i < n; i++)
  {
    if (id[i] > 0.0)
    {
      double p = x[i] * w[i];
      acc += p * p;
    }
  }
  return acc;
}

double orc_wrms_norm(const double* x, const double* w, orc_index n) /* serial:641-648 */
{
  return guarded_sqrt(orc_wsqr_sum(x, w, n) / (double)n);
}

double orc_wrms_norm_mask(const double* x, const double* w, const double* id,
                          orc_index n) /* serial:671-678 */
{
  return guarded_sqrt(orc_wsqr_sum_mask(x, w, id, n) / (double)n);
}

double orc_min(const double* x, orc_index n) /* serial:705-723 (undefined for n == 0) */
{
  double best = x[0];
  for (orc_index i = 1; i < n; i++)
  {
    if (x[i] < best) best = x[i];
  }
  return best;
}

double orc_wl2_norm(const double* x, const double* w, orc_index n) /* serial:725-744 */
{
  return guarded_sqrt(orc_wsqr_sum(x, w, n));
}

double orc_l1_norm(const double* x, orc_index n) /* serial:746-760 */
{
  double acc = 0.0;
  for (orc_index i = 0; i < n; i++) acc += fabs(x[i]);
  return acc;
}

/* serial:778-798: z is left untouched where x == 0; returns 1 iff no zero */
int orc_inv_test(const double* x, double* z, orc_index n)
{
  int all_nonzero = 1;
  for (orc_index i = 0; i < n; i++)
  {
    if (x[i] == 0.0) all_nonzero = 0;
    else z[i] = 1.0 / x[i];
  }
  return all_nonzero;
}

/* serial:800-831.  Constraint codes: |c|>1.5 -> x*c must be > 0; |c|>0.5 ->
 * x*c must be >= 0; c == 0 -> unconstrained (m = 0).  Returns 1 iff no violation. */
int orc_constr_mask(const double* c, const double* x, double* m, orc_index n)
{
  int violated = 0;
  for (orc_index i = 0; i < n; i++)
  {
    m[i] = 0.0;
    if (c[i] == 0.0) continue;
    double s  = x[i] * c[i];
    double ac = fabs(c[i]);
    if ((ac > 1.5 && s <= 0.0) || (ac > 0.5 && s < 0.0))
    {
      m[i]     = 1.0;
      violated = 1;
    }
  }
  return !violated;
}

/* serial:833-863.  min over denom != 0 of num/denom; DBL_MAX (SUN_BIG_REAL) when
 * every denominator is zero.  First hit initialises, later hits use
 * SUNMIN(min, q) = (min < q ? min : q). */
double orc_min_quotient(const double* num, const double* denom, orc_index n)
{
  int seen    = 0;
  double best = DBL_MAX;
  for (orc_index i = 0; i < n; i++)
  {
    if (denom[i] == 0.0) continue;
    double q = num[i] / denom[i];
    if (!seen)
    {
      best = q;
      seen = 1;
    }
    else { best = (best < q) ? best : q; }
  }
  return best;
}

/* ---------------------------------------------------------------------------
 * fused ops
 * ------------------------------------------------------------------------- */

/* serial:871-942.  nvec==1 -> Scale, nvec==2 -> LinearSum (with all its forms);
 * otherwise accumulate in j order, one full pass per term. The three in-place
 * variants (serial:907-941) only differ in how the first term is produced. */
int orc_linear_combination(int nvec, const double* c, double* const* X, double* z, orc_index n)
{
  if (nvec < 1) return ORC_ERR_ARG;
  if (nvec == 1)
  {
    orc_scale(c[0], X[0], z, n);
    return ORC_SUCCESS;
  }
  if (nvec == 2)
  {
    orc_linear_sum(c[0], X[0], c[1], X[1], z, n);
    return ORC_SUCCESS;
  }
  orc_index k;
  if (X[0] == z)
  {
    if (c[0] != 1.0)
    {
      for (k = 0; k < n; k++) z[k] *= c[0]; /* serial:922 */
    }
  }
  else
  {
    const double* x0 = X[0];
    for (k = 0; k < n; k++) z[k] = c[0] * x0[k]; /* serial:935 */
  }
  for (int j = 1; j < nvec; j++)
  {
    const double* xj = X[j];
    for (k = 0; k < n; k++) z[k] += c[j] * xj[k]; /* serial:912,926,939 */
  }
  return ORC_SUCCESS;
}

/* serial:944-992 */
int orc_scale_add_multi(int nvec, const double* a, const double* x, double* const* Y,
                        double* const* Z, orc_index n)
{
  if (nvec < 1) return ORC_ERR_ARG;
  if (nvec == 1)
  {
    orc_linear_sum(a[0], x, 1.0, Y[0], Z[0], n); /* serial:960 */
    return ORC_SUCCESS;
  }
  for (int j = 0; j < nvec; j++)
  {
    const double* y = Y[j];
    double* z       = Z[j];
    if (Y == Z)
    {
      for (orc_index k = 0; k < n; k++) z[k] += a[j] * x[k]; /* serial:977 */
    }
    else
    {
      for (orc_index k = 0; k < n; k++) z[k] = a[j] * x[k] + y[k]; /* serial:989 */
    }
  }
  return ORC_SUCCESS;
}

/* serial:994-1027 */
int orc_dot_prod_multi(int nvec, const double* x, double* const* Y, double* dotprods, orc_index n)
{
  if (nvec < 1) return ORC_ERR_ARG;
  for (int j = 0; j < nvec; j++) dotprods[j] = orc_dot_prod(x, Y[j], n);
  return ORC_SUCCESS;
}

/* ---------------------------------------------------------------------------
 * vector-array ops
 * ------------------------------------------------------------------------- */

/* serial:1035-1151.  nvec == 1 delegates to N_VLinearSum (handle aliasing);
 * nvec > 1 uses the same scalar case analysis but with ARRAY identity (Z == Y,
 * Z == X as N_Vector* arrays) deciding the axpy forms. */
int orc_linear_sum_vector_array(int nvec, double a, double* const* X, double b, double* const* Y,
                                double* const* Z, orc_index n)
{
  if (nvec < 1) return ORC_ERR_ARG;
  if (nvec == 1)
  {
    orc_linear_sum(a, X[0], b, Y[0], Z[0], n);
    return ORC_SUCCESS;
  }
  enum lsum_form f = lsum_classify(a, b, Z == X, Z == Y);
  for (int i = 0; i < nvec; i++) lsum_apply(f, a, X[i], b, Y[i], Z[i], n);
  return ORC_SUCCESS;
}

/* serial:1153-1199 */
int orc_scale_vector_array(int nvec, const double* c, double* const* X, double* const* Z, orc_index n)
{
  if (nvec < 1) return ORC_ERR_ARG;
  if (nvec == 1)
  {
    orc_scale(c[0], X[0], Z[0], n);
    return ORC_SUCCESS;
  }
  for (int i = 0; i < nvec; i++)
  {
    const double* x = X[i];
    double* z       = Z[i];
    if (X == Z)
    {
      for (orc_index k = 0; k < n; k++) z[k] *= c[i]; /* serial:1184 */
    }
    else
    {
      for (orc_index k = 0; k < n; k++) z[k] = c[i] * x[k]; /* serial:1196 (no +-1 shortcuts) */
    }
  }
  return ORC_SUCCESS;
}

/* serial:1201-1230 */
int orc_const_vector_array(int nvec, double c, double* const* Z, orc_index n)
{
  if (nvec < 1) return ORC_ERR_ARG;
  for (int i = 0; i < nvec; i++) orc_const(c, Z[i], n);
  return ORC_SUCCESS;
}

/* serial:1232-1266 */
int orc_wrms_norm_vector_array(int nvec, double* const* X, double* const* W, double* nrm, orc_index n)
{
  if (nvec < 1) return ORC_ERR_ARG;
  for (int i = 0; i < nvec; i++) nrm[i] = orc_wrms_norm(X[i], W[i], n);
  return ORC_SUCCESS;
}

/* serial:1268-1307 */
int orc_wrms_norm_mask_vector_array(int nvec, double* const* X, double* const* W, const double* id,
                                    double* nrm, orc_index n)
{
  if (nvec < 1) return ORC_ERR_ARG;
  for (int i = 0; i < nvec; i++) nrm[i] = orc_wrms_norm_mask(X[i], W[i], id, n);
  return ORC_SUCCESS;
}

/* serial:1309-1408.  Y, Z flattened as P[j*nvec + i]; "Y == Z" in the reference
 * (identity of the N_Vector** arrays) is expressed by passing the same pointer. */
int orc_scale_add_multi_vector_array(int nvec, int nsum, const double* a, double* const* X,
                                     double* const* Y, double* const* Z, orc_index n)
{
  if (nvec < 1 || nsum < 1) return ORC_ERR_ARG;
  if (nvec == 1)
  {
    if (nsum == 1)
    {
      orc_linear_sum(a[0], X[0], 1.0, Y[0], Z[0], n); /* serial:1334 */
      return ORC_SUCCESS;
    }
    /* serial:1340-1351: gathers fresh YY/ZZ handle arrays, so the delegated
       N_VScaleAddMulti never sees Y == Z array identity */
    double** yy = (double**)malloc((size_t)nsum * sizeof(double*));
    double** zz = (double**)malloc((size_t)nsum * sizeof(double*));
    for (int j = 0; j < nsum; j++)
    {
      yy[j] = Y[j * nvec];
      zz[j] = Z[j * nvec];
    }
    int rc = orc_scale_add_multi(nsum, a, X[0], yy, zz, n);
    free(yy);
    free(zz);
    return rc;
  }
  if (nsum == 1)
  {
    /* serial:1366: N_VLinearSumVectorArray(nvec, a[0], X, 1, Y[0], Z[0]); the row
       arrays Y[0], Z[0] are identical objects iff the caller aliased them */
    return orc_linear_sum_vector_array(nvec, a[0], X, 1.0, Y, Z, n);
  }
  for (int i = 0; i < nvec; i++)
  {
    const double* x = X[i];
    for (int j = 0; j < nsum; j++)
    {
      const double* y = Y[j * nvec + i];
      double* z       = Z[j * nvec + i];
      if (Y == Z)
      {
        for (orc_index k = 0; k < n; k++) z[k] += a[j] * x[k]; /* serial:1388 */
      }
      else
      {
        for (orc_index k = 0; k < n; k++) z[k] = a[j] * x[k] + y[k]; /* serial:1404 */
      }
    }
  }
  return ORC_SUCCESS;
}

/* serial:1410-1544.  X flattened as X[i*nvec + j] (i = term, j = vector).
 * "X[0] == Z" (row-array identity) is expressed by X (row 0) == Z pointer. */
int orc_linear_combination_vector_array(int nvec, int nsum, const double* c, double* const* X,
                                        double* const* Z, orc_index n)
{
  if (nvec < 1 || nsum < 1) return ORC_ERR_ARG;
  if (nvec == 1)
  {
    if (nsum == 1)
    {
      orc_scale(c[0], X[0], Z[0], n);
      return ORC_SUCCESS;
    }
    if (nsum == 2)
    {
      orc_linear_sum(c[0], X[0], c[1], X[1], Z[0], n);
      return ORC_SUCCESS;
    }
    /* nvec == 1: X[i*1 + 0] is already the gathered array (serial:1450-1455) */
    return orc_linear_combination(nsum, c, X, Z[0], n);
  }
  if (nsum == 1)
  {
    /* serial:1467-1477: ScaleVectorArray with c[0] replicated */
    double* ctmp = (double*)malloc((size_t)nvec * sizeof(double));
    for (int j = 0; j < nvec; j++) ctmp[j] = c[0];
    int rc = orc_scale_vector_array(nvec, ctmp, X, Z, n);
    free(ctmp);
    return rc;
  }
  if (nsum == 2)
  {
    /* serial:1483 */
    return orc_linear_sum_vector_array(nvec, c[0], X, c[1], X + nvec, Z, n);
  }
  for (int j = 0; j < nvec; j++)
  {
    double* z = Z[j];
    orc_index k;
    if (X == Z) /* row 0 of X is the Z array itself */
    {
      if (c[0] != 1.0)
      {
        for (k = 0; k < n; k++) z[k] *= c[0]; /* serial:1519 */
      }
    }
    else
    {
      const double* x0 = X[j];
      for (k = 0; k < n; k++) z[k] = c[0] * x0[k]; /* serial:1536 */
    }
    for (int i = 1; i < nsum; i++)
    {
      const double* xi = X[i * nvec + j];
      for (k = 0; k < n; k++) z[k] += c[i] * xi[k]; /* serial:1505,1523,1540 */
    }
  }
  return ORC_SUCCESS;
}

/* ---------------------------------------------------------------------------
 * multi-rank semantics (src/nvector/manyvector/nvector_manyvector.c:940-965):
 * WrmsNorm = sqrt( allreduce_sum(local WSqrSum) / global_length ), summed in
 * rank order here.
 * ------------------------------------------------------------------------- */
double orc_mpi_wrms_from_local(const double* local_sqrsums, int nranks, orc_index global_n)
{
  double g = 0.0;
  for (int r = 0; r < nranks; r++) g += local_sqrsums[r];
  return guarded_sqrt(g / (double)global_n);
}

/* ---------------------------------------------------------------------------
 * input generator: the C99 LCG the reference benchmark uses
 * (benchmarks/nvector/test_nvector_performance.c:2751-2768) but with a FIXED
 * seed (the reference seeds from time(NULL)).
 * ------------------------------------------------------------------------- */
void orc_fill_uniform(double* x, orc_index n, uint32_t seed, double lo, double hi)
{
  const uint32_t mask = 0x7fffffffu;
  uint32_t s          = seed & mask;
  double range        = hi - lo;
  for (orc_index i = 0; i < n; i++)
  {
    s    = (1103515245u * s + 12345u) & mask;
    x[i] = range * ((double)s / (double)mask) + lo;
  }
}
