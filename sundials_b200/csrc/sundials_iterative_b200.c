/* sundials_iterative_b200.c -- classical Gram-Schmidt for NVECTOR_B200 in two kernels per column.
 * Same signature, outputs and re-orthogonalisation rule as the reference's SUNClassicalGS
 * (src/sundials/sundials_iterative.c:116-170, "ref:<line>" below); see
 * include/sundials_iterative_b200.h for what is fused. */
#include "sundials_iterative_b200.h"

#include <math.h>

#define FACTOR 1000.0 /* ref:30 */
#define B200_MGS_CHAIN_MAX 60 /* longest sweep one chain handles (result slots per context - 2) */
#define B200_GS_CHAIN_MAX 24 /* widest column one chain handles (multi-dot outputs per launch) */

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

  const int nvec = k - i0 + 1;
  /* operands of the combination in the reference's order (ref:133-140, including its indexing of h
     and v from 0): vtemp = {v[k], v[0], v[1], ...} */
  for (int i = nvec - 2; i >= 0; i--) vtemp[i + 1] = v[i];
  vtemp[0] = v[k];
  if (nvec <= B200_GS_CHAIN_MAX)
  {
    /* ONE chain, one host wait: the (k+1)-wide multi-dot (ref:130; x = v[k] is itself the last Y and is read
       once), then v[k] <- v[k] - sum h_i v_i together with ||v[k]||^2 (ref:142 + ref:146) with the
       coefficients -h_i taken from the device result slots */
    err = N_VClassicalGSStep_B200(nvec, v[k], v + i0, vtemp, v[k], stemp, &sq);
    if (err) return err;
  }
  else
  {
    err = N_VDotProdMulti_B200(nvec, v[k], v + i0, stemp);
    if (err) return err;
  }
  const sunrealtype vk_norm = rsqrt_guard(stemp[nvec - 1]);
  for (int i = nvec - 2; i >= 0; i--)
  {
    h[i][k_minus_1] = stemp[i];
    stemp[i + 1]    = -stemp[i];
  }
  stemp[0] = 1.0;
  if (nvec > B200_GS_CHAIN_MAX)
  {
    err = N_VLinearCombinationSqNorm_B200(nvec, stemp, vtemp, v[k], &sq);
    if (err) return err;
  }
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

/* ------------------------------------------------------------------ modified Gram-Schmidt
 * ref:45-105.  Every N_VLinearSum of the sweep is fused with the N_VDotProd that follows it. */
static long g_mgs_calls = 0;
long SUNModifiedGS_B200_Calls(void) { return g_mgs_calls; }

SUNErrCode SUNModifiedGS_B200(N_Vector* v, sunrealtype** h, int k, int p, sunrealtype* new_vk_norm)
{
  const int k_minus_1 = k - 1;
  const int i0        = (k - p > 0) ? k - p : 0; /* ref:56 */
  SUNErrCode err;
  sunrealtype d2[2], hi, sq = 0.0;
  g_mgs_calls++;
  if (i0 >= k)
  { /* nothing to orthogonalise against (k == 0): the reference computes the norm twice */
    *new_vk_norm = rsqrt_guard(N_VDotProd_B200(v[k], v[k]));
    return SUN_SUCCESS;
  }

  sunrealtype vk_norm;
  if (k - i0 <= B200_MGS_CHAIN_MAX)
  {
    /* the whole column as ONE chain of k - i0 + 1 kernels and one host wait: every kernel reads the
       projection it needs from the device result slot its predecessor wrote */
    sunrealtype hcol[B200_MGS_CHAIN_MAX], norms[2];
    err = N_VModifiedGSSweep_B200(k - i0, v + i0, v[k], hcol, norms);
    if (err) return err;
    for (int i = i0; i < k; i++) h[i][k_minus_1] = hcol[i - i0];
    vk_norm = rsqrt_guard(norms[0]);
    sq      = norms[1];
  }
  else
  {
    /* ||v_k||^2 and the first projection in one 2-wide multi-dot (ref:52 + first pass of ref:62) */
    N_Vector y2[2] = {v[k], v[i0]};
    err = N_VDotProdMulti_B200(2, v[k], y2, d2);
    if (err) return err;
    vk_norm = rsqrt_guard(d2[0]);
    hi      = d2[1];
    for (int i = i0; i < k; i++) /* ref:60-66 */
    {
      h[i][k_minus_1] = hi;
      if (i + 1 < k)
      { /* v_k <- v_k - h_i v_i, and h_{i+1} = v_{i+1} . v_k on the updated vector */
        err = N_VAxpyDot_B200(-hi, v[i], v[k], v[i + 1], &hi);
        if (err) return err;
      }
      else
      { /* last update together with the new norm (ref:70) */
        sunrealtype c2[2] = {1.0, -hi};
        N_Vector x2[2]    = {v[k], v[i]};
        err               = N_VLinearCombinationSqNorm_B200(2, c2, x2, v[k], &sq);
        if (err) return err;
      }
    }
  }
  *new_vk_norm = rsqrt_guard(sq);

  /* ref:79-80: reorthogonalise only if the new vector is tiny against the old one */
  sunrealtype temp = FACTOR * vk_norm;
  if ((temp + (*new_vk_norm)) != temp) return SUN_SUCCESS;

  sunrealtype new_norm_2 = 0.0; /* ref:82-102, rare: the unfused ops */
  for (int i = i0; i < k; i++)
  {
    sunrealtype new_product = N_VDotProd_B200(v[i], v[k]);
    temp                    = FACTOR * h[i][k_minus_1];
    if ((temp + new_product) == temp) continue;
    h[i][k_minus_1] += new_product;
    N_VLinearSum_B200(1.0, v[k], -new_product, v[i], v[k]);
    new_norm_2 += new_product * new_product;
  }
  if (new_norm_2 != 0.0)
  {
    sunrealtype new_product = (*new_vk_norm) * (*new_vk_norm) - new_norm_2;
    *new_vk_norm            = (new_product > 0.0) ? rsqrt_guard(new_product) : 0.0;
  }
  return SUN_SUCCESS;
}
