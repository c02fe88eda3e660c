/* test_nvector_b200.c -- runs the REFERENCE's own N_Vector unit-test harness
 * (test/unit_tests/nvector/test_nvector.c, compiled by path, unmodified) against
 * NVECTOR_B200.  This file only supplies the backend hooks the harness asks for
 * (test_nvector.h:43-50) and the list of tests to run.
 *
 *   test_nvector_b200 <length> <print_timing> [kinds]
 *
 * For every memory kind (device / managed / pinned) all standard, fused,
 * vector-array, local-reduction and XBraid-buffer tests are run, the fused ones
 * twice: with the fused slots disabled (SUNDIALS' generic fallback loops over
 * our streaming kernels) and enabled (our fused kernels).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <sundials/sundials_math.h>
#include <sundials/sundials_nvector.h>
#include <sundials/sundials_types.h>

#include "nvector_b200.h"
#include "test_nvector.h"

/* ---- hooks ---- */
int check_ans(sunrealtype ans, N_Vector X, sunindextype local_length)
{
  int failure = 0;
  N_VCopyFromDevice_B200(X);
  sunrealtype* h = N_VGetHostArrayPointer_B200(X);
  if (local_length != N_VGetLocalLength_B200(X)) return 1;
  for (sunindextype i = 0; i < local_length; i++) failure += SUNRCompare(h[i], ans);
  return (failure > 0) ? 1 : 0;
}

sunbooleantype has_data(N_Vector X)
{
  if (X == NULL || X->content == NULL) return SUNFALSE;
  return (N_VGetDeviceArrayPointer_B200(X) != NULL || N_VGetLocalLength_B200(X) == 0) ? SUNTRUE : SUNFALSE;
}

void set_element_range(N_Vector X, sunindextype is, sunindextype ie, sunrealtype val)
{
  sunindextype cnt = ie - is + 1;
  if (cnt <= 0) return;
  sunrealtype* tmp = (sunrealtype*)malloc((size_t)cnt * sizeof(sunrealtype));
  for (sunindextype i = 0; i < cnt; i++) tmp[i] = val;
  b200vec_copy_h2d(N_VGetCtx_B200(X), N_VGetDeviceArrayPointer_B200(X) + is, tmp, (size_t)cnt * sizeof(sunrealtype), 1);
  free(tmp);
}

void set_element(N_Vector X, sunindextype i, sunrealtype val) { set_element_range(X, i, i, val); }

sunrealtype get_element(N_Vector X, sunindextype i)
{
  sunrealtype v = 0;
  b200vec_copy_d2h(N_VGetCtx_B200(X), &v, N_VGetDeviceArrayPointer_B200(X) + i, sizeof(sunrealtype), 1);
  return v;
}

double max_time(N_Vector X, double time)
{
  (void)X;
  return time;
}

void sync_device(N_Vector X) { b200vec_ctx_sync(N_VGetCtx_B200(X)); }

/* ---- driver ---- */
static int run_kind(int kind, const char* kname, sunindextype length)
{
  int fails = 0;
  printf("\n=== NVECTOR_B200 memory kind: %s, length %ld ===\n", kname, (long)length);
  N_Vector X = N_VNewWithCtx_B200(length, kind, NULL, sunctx);
  if (!X)
  {
    printf(">>> FAILED: constructor returned NULL\n");
    return 1;
  }
  N_Vector Y = N_VClone(X), Z = N_VClone(X);
  if (!Y || !Z) return 1;

  fails += Test_N_VMake(X, length, 0);
  fails += Test_N_VGetVectorID(X, SUNDIALS_NVEC_CUSTOM, 0);
  fails += Test_N_VGetLength(X, 0);
  fails += Test_N_VGetCommunicator(X, SUN_COMM_NULL, 0);
  fails += Test_N_VCloneEmpty(X, 0);
  fails += Test_N_VClone(X, length, 0);
  fails += Test_N_VCloneEmptyVectorArray(5, X, 0);
  fails += Test_N_VCloneVectorArray(5, X, length, 0);
  if (kind != B200_MEM_DEVICE) fails += Test_N_VGetArrayPointer(X, length, 0); /* host-coherent kinds */

  fails += Test_N_VConst(X, length, 0);
  fails += Test_N_VLinearSum(X, Y, Z, length, 0);
  fails += Test_N_VProd(X, Y, Z, length, 0);
  fails += Test_N_VDiv(X, Y, Z, length, 0);
  fails += Test_N_VScale(X, Z, length, 0);
  fails += Test_N_VAbs(X, Z, length, 0);
  fails += Test_N_VInv(X, Z, length, 0);
  fails += Test_N_VAddConst(X, Z, length, 0);
  fails += Test_N_VDotProd(X, Y, length, 0);
  fails += Test_N_VMaxNorm(X, length, 0);
  fails += Test_N_VWrmsNorm(X, Y, length, 0);
  fails += Test_N_VWrmsNormMask(X, Y, Z, length, 0);
  fails += Test_N_VMin(X, length, 0);
  fails += Test_N_VWL2Norm(X, Y, length, 0);
  fails += Test_N_VL1Norm(X, length, 0);
  if (length >= 3) fails += Test_N_VCompare(X, Z, length, 0);
  fails += Test_N_VInvTest(X, Z, length, 0);
  if (length >= 7) fails += Test_N_VConstrMask(X, Y, Z, length, 0);
  fails += Test_N_VMinQuotient(X, Y, length, 0);

  for (int enabled = 0; enabled <= 1; enabled++)
  {
    printf("\n--- fused and vector-array operations %s ---\n", enabled ? "ENABLED" : "DISABLED (generic fallback)");
    N_Vector U = N_VClone(X);
    if (N_VEnableFusedOps_B200(U, enabled ? SUNTRUE : SUNFALSE)) return fails + 1;
    fails += Test_N_VLinearCombination(U, length, 0);
    fails += Test_N_VScaleAddMulti(U, length, 0);
    fails += Test_N_VDotProdMulti(U, length, 0);
    fails += Test_N_VLinearSumVectorArray(U, length, 0);
    fails += Test_N_VScaleVectorArray(U, length, 0);
    fails += Test_N_VConstVectorArray(U, length, 0);
    fails += Test_N_VWrmsNormVectorArray(U, length, 0);
    fails += Test_N_VWrmsNormMaskVectorArray(U, length, 0);
    fails += Test_N_VScaleAddMultiVectorArray(U, length, 0);
    fails += Test_N_VLinearCombinationVectorArray(U, length, 0);
    if (enabled) fails += Test_N_VDotProdMultiLocal(U, length, 0);
    N_VDestroy(U);
  }

  printf("\n--- local reduction operations ---\n");
  fails += Test_N_VDotProdLocal(X, Y, length, 0);
  fails += Test_N_VMaxNormLocal(X, length, 0);
  fails += Test_N_VMinLocal(X, length, 0);
  fails += Test_N_VL1NormLocal(X, length, 0);
  fails += Test_N_VWSqrSumLocal(X, Y, length, 0);
  fails += Test_N_VWSqrSumMaskLocal(X, Y, Z, length, 0);
  fails += Test_N_VInvTestLocal(X, Z, length, 0);
  if (length >= 7) fails += Test_N_VConstrMaskLocal(X, Y, Z, length, 0);
  fails += Test_N_VMinQuotientLocal(X, Y, length, 0);

  printf("\n--- XBraid buffer operations ---\n");
  fails += Test_N_VBufSize(X, length, 0);
  fails += Test_N_VBufPack(X, length, 0);
  fails += Test_N_VBufUnpack(X, length, 0);

  N_VDestroy(X);
  N_VDestroy(Y);
  N_VDestroy(Z);
  return fails;
}

int main(int argc, char* argv[])
{
  if (argc < 3)
  {
    printf("usage: %s <length> <print_timing> [device|managed|pinned ...]\n", argv[0]);
    return 2;
  }
  sunindextype length = (sunindextype)atol(argv[1]);
  if (length <= 0) return 2;
  Test_Init(SUN_COMM_NULL);
  SetTiming(atoi(argv[2]), 0);

  const char* names[3] = {"device", "managed", "pinned"};
  int fails            = 0;
  for (int k = 0; k < 3; k++)
  {
    int want = (argc == 3);
    for (int a = 3; a < argc; a++) want |= !strcmp(argv[a], names[k]);
    if (want) fails += run_kind(k, names[k], length);
  }

  if (fails) printf("\nFAIL: NVECTOR_B200 module failed %d tests\n\n", fails);
  else printf("\nSUCCESS: NVECTOR_B200 module passed all tests\n\n");
  Test_Finalize();
  return fails;
}
