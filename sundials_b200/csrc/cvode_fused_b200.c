/* cvode_fused_b200.c -- libsundials_cvode_fused_b200.so: the seven functions of the reference's
 * fused-kernel plugin boundary (include/cvode_fused_b200.h) on NVECTOR_B200.  Host side only: argument
 * unpacking like the reference's CUDA library (src/cvode/cvode_fused_gpu.cpp:83-105 reads the vector
 * content directly), one b200vec_cv_* launch on the vector's context, no host wait unless the memory
 * is host-visible (managed / pinned vectors are coherent on return, like every op of the vector).
 * Returns 0, or -1 on a CUDA error (recorded by the library, b200vec_last_error) and for the atolmin0
 * failure of the error-weight functions. */
#include "cvode_fused_b200.h"

#include <stdio.h>
#include <stdlib.h>

/* B200CVF_REPORT=1: at exit, how often CVODE came through each entry point (stderr) */
enum { F_EWT_SS, F_EWT_SV, F_CONSTR, F_NLSRES, F_FORMY, F_BUILDM, F_UPDATEM, F_COUNT };
static long g_calls[F_COUNT];
static void report(void)
{
  static const char* names[F_COUNT] = {"cvEwtSetSS_fused", "cvEwtSetSV_fused", "cvCheckConstraints_fused", "cvNlsResid_fused",
                                       "cvDiagSetup_formY", "cvDiagSetup_buildM", "cvDiagSolve_updateM"};
  for (int i = 0; i < F_COUNT; i++) fprintf(stderr, "[libsundials_cvode_fused_b200] %s calls: %ld\n", names[i], g_calls[i]);
}
__attribute__((constructor)) static void init(void)
{
  const char* e = getenv("B200CVF_REPORT");
  if (e && e[0] && e[0] != '0') atexit(report);
}

#define NVC(v)  ((N_VectorContent_B200)((v)->content))
#define NDEV(v) (NVC(v)->device_data)
#define NCTX(v) (NVC(v)->ctx)
#define NLEN(v) ((int64_t)NVC(v)->length)

static int finish(N_Vector out, int rc)
{
  if (rc) return -1;
  if (NVC(out)->mem_kind != B200_MEM_DEVICE && b200vec_ctx_sync(NCTX(out))) return -1;
  return 0;
}

static int ewt_set(sunbooleantype atolmin0, sunrealtype reltol, sunrealtype Sabstol, N_Vector Vabstol, N_Vector ycur,
                   N_Vector tempv, N_Vector weight)
{
  const sunrealtype* av = Vabstol ? NDEV(Vabstol) : NULL;
  if (!atolmin0)
    return finish(weight, b200vec_cv_ewt(NCTX(weight), reltol, Sabstol, av, NDEV(ycur), NDEV(tempv), NDEV(weight),
                                         NLEN(weight)));
  /* some atol_i is zero: the stubs test min(tempv) before they invert (a global minimum on a
     distributed vector, through the ops table) */
  if (finish(tempv, b200vec_cv_ewt(NCTX(weight), reltol, Sabstol, av, NDEV(ycur), NDEV(tempv), NULL, NLEN(weight))))
    return -1;
  if (tempv->ops->nvmin(tempv) <= 0.0) return -1; /* N_VMin / N_VInv, without depending on sundials_core */
  weight->ops->nvinv(tempv, weight);
  return 0;
}

int cvEwtSetSS_fused(const sunbooleantype atolmin0, const sunrealtype reltol, const sunrealtype Sabstol,
                     const N_Vector ycur, N_Vector tempv, N_Vector weight)
{
  g_calls[F_EWT_SS]++;
  return ewt_set(atolmin0, reltol, Sabstol, NULL, ycur, tempv, weight);
}

int cvEwtSetSV_fused(const sunbooleantype atolmin0, const sunrealtype reltol, const N_Vector Vabstol,
                     const N_Vector ycur, N_Vector tempv, N_Vector weight)
{
  g_calls[F_EWT_SV]++;
  return ewt_set(atolmin0, reltol, 0.0, Vabstol, ycur, tempv, weight);
}

int cvCheckConstraints_fused(const N_Vector c, const N_Vector ewt, const N_Vector y, const N_Vector mm, N_Vector tempv)
{
  g_calls[F_CONSTR]++;
  return finish(tempv, b200vec_cv_constraints(NCTX(tempv), NDEV(c), NDEV(ewt), NDEV(y), NDEV(mm), NDEV(tempv),
                                              NLEN(tempv)));
}

int cvNlsResid_fused(const sunrealtype rl1, const sunrealtype ngamma, const N_Vector zn1, const N_Vector ycor,
                     const N_Vector ftemp, N_Vector res)
{
  g_calls[F_NLSRES]++;
  return finish(res, b200vec_cv_nls_resid(NCTX(res), rl1, ngamma, NDEV(zn1), NDEV(ycor), NDEV(ftemp), NDEV(res),
                                          NLEN(res)));
}

int cvDiagSetup_formY(const sunrealtype h, const sunrealtype r, const N_Vector fpred, const N_Vector zn1,
                      const N_Vector ypred, N_Vector ftemp, N_Vector y)
{
  g_calls[F_FORMY]++;
  return finish(y, b200vec_cv_diag_form_y(NCTX(y), h, r, NDEV(fpred), NDEV(zn1), NDEV(ypred), NDEV(ftemp), NDEV(y),
                                          NLEN(y)));
}

int cvDiagSetup_buildM(const sunrealtype fract, const sunrealtype uround, const sunrealtype h, const N_Vector ftemp,
                       const N_Vector fpred, const N_Vector ewt, N_Vector bit, N_Vector bitcomp, N_Vector y, N_Vector M)
{
  (void)fract; /* the stubs use the constant FRACT = 0.1 (cvode_fused_stubs.c:28), and so does the kernel */
  g_calls[F_BUILDM]++;
  return finish(M, b200vec_cv_diag_build_m(NCTX(M), uround, h, NDEV(ftemp), NDEV(fpred), NDEV(ewt), NDEV(bit),
                                           NDEV(bitcomp), NDEV(y), NDEV(M), NLEN(M)));
}

int cvDiagSolve_updateM(const sunrealtype r, N_Vector M)
{
  g_calls[F_UPDATEM]++;
  return finish(M, b200vec_cv_diag_update_m(NCTX(M), r, NDEV(M), NLEN(M)));
}
