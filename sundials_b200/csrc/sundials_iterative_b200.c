/* sundials_iterative_b200.c -- classical Gram-Schmidt for NVECTOR_B200 in two kernels per column.
 * Same signature, outputs and re-orthogonalisation rule as the reference's SUNClassicalGS
 * (src/sundials/sundials_iterative.c:116-170, "ref:<line>" below); see
 * include/sundials_iterative_b200.h for what is fused. */
#include "sundials_iterative_b200.h"

#include <math.h>

#define FACTOR 1000.0 /* ref:30 */

static long g_calls = 0;
long SUNClassicalGS_B200_Calls(void) { return g_calls; }

static sunrealtype rsqrt_guard(sunrealtype x) { return (x <= 0.0) ? 0.0 : sqrt(x); } /* SUNRsqrt */

SUNErrCode SUNClassicalGS_B200(N_Vector* v, sunrealtype** h, int k, int p, sunrealtype* new_vk_norm, sunrealtype* stemp,
                               N_Vector* vtemp)
{
  const int k_minus_1 = k - 1;
  const int i0        = (k - p > 0) ? k - p : 0; /* ref:126 */
  SUNErrCode err;
  sunrealtype sq = 0.0;
  g_calls++;

  /* all projections and v[k].v[k] in ONE multi-dot (ref:130); x = v[k] is itself the last Y, the
     kernel reads it once */
  err = N_VDotProdMulti_B200(k - i0 + 1, v[k], v + i0, stemp);
  if (err) return err;

  const sunrealtype vk_norm = rsqrt_guard(stemp[k - i0]);
  for (int i = k - i0 - 1; i >= 0; i--) /* ref:133-138, including its indexing of h and v from 0 */
  {
    h[i][k_minus_1] = stemp[i];
    stemp[i + 1]    = -stemp[i];
    vtemp[i + 1]    = v[i];
  }
  stemp[0] = 1.0;
  vtemp[0] = v[k];

  /* v[k] <- v[k] - sum h_i v_i  and  ||v[k]||^2 of the result, one pass (ref:142 + ref:146) */
  err = N_VLinearCombinationSqNorm_B200(k - i0 + 1, stemp, vtemp, v[k], &sq);
  if (err) return err;
  *new_vk_norm = rsqrt_guard(sq);

  /* re-orthogonalise if the new vector is tiny against the old one (ref:151-168) */
  if ((FACTOR * (*new_vk_norm)) < vk_norm)
  {
    err = N_VDotProdMulti_B200(k - i0, v[k], v + i0, stemp + 1);
    if (err) return err;
    stemp[0] = 1.0;
    vtemp[0] = v[k];
    for (int i = i0; i < k; i++)
    {
      h[i][k_minus_1] += stemp[i - i0 + 1];
      stemp[i - i0 + 1] = -stemp[i - i0 + 1];
      vtemp[i - i0 + 1] = v[i - i0];
    }
    err = N_VLinearCombinationSqNorm_B200(k + 1, stemp, vtemp, v[k], &sq);
    if (err) return err;
    *new_vk_norm = rsqrt_guard(sq);
  }
  return SUN_SUCCESS;
}
