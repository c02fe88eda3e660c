/* cvode_fused_b200.h -- CVODE's integrator-level fused kernels for NVECTOR_B200.
 *
 * libsundials_cvode_fused_b200.so takes the place of libsundials_cvode_fused_cuda / _hip / _stubs
 * (src/cvode/CMakeLists.txt:43-75) under a libsundials_cvode built with the reference's own switch
 * SUNDIALS_BUILD_PACKAGE_FUSED_KERNELS: it exports the seven functions CVODE calls when
 * CVodeSetUseIntegratorFusedKernels(mem, SUNTRUE) is in effect, with the prototypes of
 * src/cvode/cvode_impl.h:639-672 (repeated below; that header is private to the reference).
 *
 *   function                    call site in the reference           N_V* ops it replaces (stubs)
 *   cvEwtSetSS_fused            cvode.c:4798  (cvEwtSetSS)            4-5   cvode_fused_stubs.c:38-50
 *   cvEwtSetSV_fused            cvode.c:4838  (cvEwtSetSV)            3-4   :59-71
 *   cvCheckConstraints_fused    cvode.c:3182, cvode_constraints.c:61  5     :80-89
 *   cvNlsResid_fused            cvode_nls.c:388 (cvNlsResidual)       2     :97-104
 *   cvDiagSetup_formY           cvode_diag.c:344                      2     :112-119
 *   cvDiagSetup_buildM          cvode_diag.c:373                      10    :128-147
 *   cvDiagSolve_updateM         cvode_diag.c:434                      4     :154-161
 *
 * Each is ONE kernel (b200vec_cv_*, include/b200vec.h) whose element-wise arithmetic is the stubs' op
 * sequence, so a run with the fused kernels on prints what the same run prints with them off, bit for
 * bit -- the reference's CUDA kernels do not promise that (FMA contraction, `>` for N_VCompare's `>=`,
 * the atolmin0 test dropped).  The atolmin0 test is kept: tempv is formed, its minimum taken, and the
 * weights are only written when it is positive, exactly as the stubs do.
 *
 * The reference admits these kernels only for vectors whose N_VGetVectorID is SUNDIALS_NVEC_CUDA or
 * _HIP (src/cvode/cvode_io.c:1022-1029): call N_VSetVectorID_B200(v, SUNDIALS_NVEC_CUDA) on the
 * template vector (clones inherit it) before CVodeInit -- NVECTOR_B200 then answers as the vector it
 * replaces.  No other code in the reference's src/ or include/ trees tests that ID.
 */
#ifndef CVODE_FUSED_B200_H
#define CVODE_FUSED_B200_H

#include "nvector_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

int cvEwtSetSS_fused(const sunbooleantype atolmin0, const sunrealtype reltol, const sunrealtype Sabstol,
                     const N_Vector ycur, N_Vector tempv, N_Vector weight);
int cvEwtSetSV_fused(const sunbooleantype atolmin0, const sunrealtype reltol, const N_Vector Vabstol,
                     const N_Vector ycur, N_Vector tempv, N_Vector weight);
int cvCheckConstraints_fused(const N_Vector c, const N_Vector ewt, const N_Vector y, const N_Vector mm,
                             N_Vector tempv);
int cvNlsResid_fused(const sunrealtype rl1, const sunrealtype ngamma, const N_Vector zn1, const N_Vector ycor,
                     const N_Vector ftemp, N_Vector res);
int cvDiagSetup_formY(const sunrealtype h, const sunrealtype r, const N_Vector fpred, const N_Vector zn1,
                      const N_Vector ypred, N_Vector ftemp, N_Vector y);
int cvDiagSetup_buildM(const sunrealtype fract, const sunrealtype uround, const sunrealtype h, const N_Vector ftemp,
                       const N_Vector fpred, const N_Vector ewt, N_Vector bit, N_Vector bitcomp, N_Vector y,
                       N_Vector M);
int cvDiagSolve_updateM(const sunrealtype r, N_Vector M);

#ifdef __cplusplus
}
#endif
#endif
