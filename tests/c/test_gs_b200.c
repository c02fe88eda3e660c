/* test_gs_b200.c -- Gram-Schmidt on the new vector.
 *
 * Runs the reference's UNMODIFIED SUNModifiedGS / SUNClassicalGS
 * (src/sundials/sundials_iterative.c:45-170, from libsundials_ref.so) on a
 * Krylov basis held (a) in nvector_serial and (b) in NVECTOR_B200 with fused ops
 * enabled, building the basis column by column exactly like SPGMR does
 * (orthogonalise v[k], normalise by the returned norm).  Compares the
 * Hessenberg entries, the returned norms and the final basis.
 *
 *   test_gs_b200 <n> <maxl> <tol>      tol = 0 demands bit-identical results
 *
 * A third and a fourth pass compare the reference's SUNClassicalGS / SUNModifiedGS on nvector_serial
 * with the fused SUNClassicalGS_B200 / SUNModifiedGS_B200 (include/sundials_iterative_b200.h) on
 * NVECTOR_B200.
 * Prints one line per (gstype, k) and exits with the number of mismatches.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <nvector/nvector_serial_ref.h>
#include <sundials/sundials_iterative.h>
#include <sundials/sundials_math.h>

#include "nvector_b200.h"
#include "sundials_iterative_b200.h"

#define FUSED_CLASSICAL_GS 3 /* serial: reference SUNClassicalGS; B200: SUNClassicalGS_B200 (2 kernels per column) */
#define FUSED_MODIFIED_GS  4 /* serial: reference SUNModifiedGS;  B200: SUNModifiedGS_B200 (k + 1 kernels per column) */
static const char* gsname(int t)
{
  return t == SUN_MODIFIED_GS ? "modified " : t == SUN_CLASSICAL_GS ? "classical" : t == FUSED_CLASSICAL_GS ? "fused-cgs" : "fused-mgs";
}

static void fill(sunrealtype* d, sunindextype n, unsigned seed)
{
  unsigned s = seed & 0x7fffffffu;
  for (sunindextype i = 0; i < n; i++)
  {
    s    = (1103515245u * s + 12345u) & 0x7fffffffu;
    d[i] = 2.0 * ((double)s / (double)0x7fffffff) - 1.0;
  }
}

static int differ(double a, double b, double tol, double scale)
{
  if (tol == 0.0) return memcmp(&a, &b, sizeof a) != 0;
  return fabs(a - b) > tol * scale;
}

int main(int argc, char** argv)
{
  sunindextype n = (argc > 1) ? atol(argv[1]) : 1000;
  int maxl       = (argc > 2) ? atoi(argv[2]) : 5;
  double tol     = (argc > 3) ? atof(argv[3]) : 0.0;
  SUNContext ctx;
  if (SUNContext_Create(SUN_COMM_NULL, &ctx)) return 99;
  int bad = 0;

  for (int gstype = SUN_MODIFIED_GS; gstype <= FUSED_MODIFIED_GS; gstype++)
  {
    N_Vector ts = N_VNew_Serial(n, ctx);
    N_Vector tb = N_VNew_B200(n, ctx);
    N_VEnableFusedOps_Serial(ts, SUNTRUE);
    N_VEnableFusedOps_B200(tb, SUNTRUE);
    N_Vector* Vs = N_VCloneVectorArray(maxl + 1, ts);
    N_Vector* Vb = N_VCloneVectorArray(maxl + 1, tb);
    N_Vector* ws = (N_Vector*)malloc((maxl + 1) * sizeof(N_Vector));
    N_Vector* wb = (N_Vector*)malloc((maxl + 1) * sizeof(N_Vector));
    sunrealtype* ss = (sunrealtype*)malloc((maxl + 1) * sizeof(sunrealtype));
    sunrealtype* sb = (sunrealtype*)malloc((maxl + 1) * sizeof(sunrealtype));
    sunrealtype** Hs = (sunrealtype**)malloc((maxl + 1) * sizeof(sunrealtype*));
    sunrealtype** Hb = (sunrealtype**)malloc((maxl + 1) * sizeof(sunrealtype*));
    for (int i = 0; i <= maxl; i++)
    {
      Hs[i] = (sunrealtype*)calloc(maxl, sizeof(sunrealtype));
      Hb[i] = (sunrealtype*)calloc(maxl, sizeof(sunrealtype));
      fill(N_VGetArrayPointer(Vs[i]), n, 1000u + (unsigned)i);
      memcpy(N_VGetHostArrayPointer_B200(Vb[i]), N_VGetArrayPointer(Vs[i]), (size_t)n * sizeof(sunrealtype));
      N_VCopyToDevice_B200(Vb[i]);
    }
    /* normalise v[0] as SPGMR does */
    sunrealtype n0s = SUNRsqrt(N_VDotProd(Vs[0], Vs[0])), n0b = SUNRsqrt(N_VDotProd(Vb[0], Vb[0]));
    N_VScale(1.0 / n0s, Vs[0], Vs[0]);
    N_VScale(1.0 / n0b, Vb[0], Vb[0]);
    bad += differ(n0s, n0b, tol, n0s);

    for (int k = 1; k <= maxl; k++)
    {
      sunrealtype nrm_s = 0, nrm_b = 0;
      if (gstype == SUN_MODIFIED_GS)
      {
        SUNModifiedGS(Vs, Hs, k, maxl, &nrm_s);
        SUNModifiedGS(Vb, Hb, k, maxl, &nrm_b);
      }
      else if (gstype == SUN_CLASSICAL_GS)
      {
        SUNClassicalGS(Vs, Hs, k, maxl, &nrm_s, ss, ws);
        SUNClassicalGS(Vb, Hb, k, maxl, &nrm_b, sb, wb);
      }
      else if (gstype == FUSED_CLASSICAL_GS)
      {
        SUNClassicalGS(Vs, Hs, k, maxl, &nrm_s, ss, ws);
        SUNClassicalGS_B200(Vb, Hb, k, maxl, &nrm_b, sb, wb);
      }
      else
      {
        SUNModifiedGS(Vs, Hs, k, maxl, &nrm_s);
        SUNModifiedGS_B200(Vb, Hb, k, maxl, &nrm_b);
      }
      int kb = differ(nrm_s, nrm_b, tol, nrm_s);
      double hmax = 0;
      for (int i = 0; i < k; i++)
      {
        kb += differ(Hs[i][k - 1], Hb[i][k - 1], tol, 1.0);
        double d = fabs(Hs[i][k - 1] - Hb[i][k - 1]);
        if (d > hmax) hmax = d;
      }
      N_VScale(1.0 / nrm_s, Vs[k], Vs[k]);
      N_VScale(1.0 / nrm_b, Vb[k], Vb[k]);
      printf("%s k=%d  norm serial %.17g b200 %.17g  max|dh| %.3g  %s\n",
             gsname(gstype), k, nrm_s, nrm_b, hmax, kb ? "MISMATCH" : "ok");
      bad += kb;
    }
    /* final basis */
    double vmax = 0;
    for (int i = 0; i <= maxl; i++)
    {
      N_VCopyFromDevice_B200(Vb[i]);
      sunrealtype *a = N_VGetArrayPointer(Vs[i]), *b = N_VGetHostArrayPointer_B200(Vb[i]);
      for (sunindextype j = 0; j < n; j++)
      {
        double d = fabs(a[j] - b[j]);
        if (d > vmax) vmax = d;
        if (tol == 0.0 && memcmp(&a[j], &b[j], sizeof(double))) bad++;
      }
    }
    if (tol > 0.0 && vmax > tol * 10) bad++;
    printf("%s basis max|dv| = %.3g\n", gsname(gstype), vmax);
    for (int i = 0; i <= maxl; i++)
    {
      free(Hs[i]);
      free(Hb[i]);
    }
    free(Hs); free(Hb); free(ss); free(sb); free(ws); free(wb);
    N_VDestroyVectorArray(Vs, maxl + 1);
    N_VDestroyVectorArray(Vb, maxl + 1);
    N_VDestroy(ts);
    N_VDestroy(tb);
  }
  printf(bad ? "FAIL: %d mismatches\n" : "SUCCESS: Gram-Schmidt on NVECTOR_B200 matches nvector_serial (%d mismatches)\n", bad);
  SUNContext_Free(&ctx);
  return bad;
}
