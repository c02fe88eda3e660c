/* sundials_iterative_b200.h -- Gram-Schmidt orthogonalisation fused for NVECTOR_B200.
 *
 * SUNClassicalGS_B200 has the exact signature and semantics of the reference's
 * SUNClassicalGS (include/sundials/sundials_iterative.h; src/sundials/sundials_iterative.c:116-170)
 * -- the routine SPGMR / SPFGMR call once per Krylov column
 * (src/sunlinsol/spgmr/sunlinsol_spgmr.c:724, spfgmr/sunlinsol_spfgmr.c:691) -- but issues TWO
 * kernels and TWO host round trips per column instead of three:
 *
 *   reference                                 here
 *   N_VDotProdMulti(k+1, v[k], v, stemp)      N_VDotProdMulti_B200        (same kernel)
 *   N_VLinearCombination(k+1, ...) -> v[k]    N_VLinearCombinationSqNorm_B200: the in-place
 *   N_VDotProd(v[k], v[k])                    combination AND sum v[k]^2 of the values it writes
 *
 * HBM traffic per column 8N(2k+3) instead of 8N(2k+4).  The Hessenberg column h[.][k-1] and
 * *new_vk_norm agree with the reference routine on nvector_serial within 1e-13 (relative) and
 * bit-for-bit for n <= 1024 (exact-order reductions).
 *
 * libsundials_b200gs.so (sundials_b200/csrc/gs_interpose.c) exports SUNClassicalGS and SUNModifiedGS
 * that forward to the _B200 routines when v[0] is an NVECTOR_B200 vector and to the next definition in
 * link order otherwise: LD_PRELOAD it (or link it ahead of sundials_core) and the UNMODIFIED SPGMR /
 * SPFGMR use the fused routine -- "interposed, reference unmodified".
 */
#ifndef SUNDIALS_ITERATIVE_B200_H
#define SUNDIALS_ITERATIVE_B200_H

#include "nvector_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

SUNErrCode SUNClassicalGS_B200(N_Vector* v, sunrealtype** h, int k, int p, sunrealtype* new_vk_norm, sunrealtype* stemp,
                               N_Vector* vtemp);

/* SUNModifiedGS (src/sundials/sundials_iterative.c:45-105; SPGMR's DEFAULT orthogonalisation,
 * sunlinsol_spgmr.c:728) with every update fused with the dot product that follows it:
 *
 *   reference (2k + 2 ops, k + 2 host round trips)      here (k + 1 kernels, k + 1 round trips)
 *   N_VDotProd(v_k, v_k)                                N_VDotProdMulti_B200(2, v_k, {v_k, v_0})
 *   for i: h_i = N_VDotProd(v_i, v_k)                   for i < k-1: N_VAxpyDot_B200(-h_i, v_i, v_k, v_{i+1}) -> h_{i+1}
 *          N_VLinearSum(1, v_k, -h_i, v_i, v_k)
 *   N_VDotProd(v_k, v_k)                                N_VLinearCombinationSqNorm_B200({1, -h_{k-1}}, {v_k, v_{k-1}})
 *
 * HBM traffic per column 8N(4k+1) instead of 8N(5k+2); same values (each dot is taken on the
 * vector updated so far, exactly as the reference does), bit-identical for n <= 1024. */
SUNErrCode SUNModifiedGS_B200(N_Vector* v, sunrealtype** h, int k, int p, sunrealtype* new_vk_norm);
long SUNModifiedGS_B200_Calls(void);
/* calls made so far in this process (tests / the interposition proof) */
long SUNClassicalGS_B200_Calls(void);

#ifdef __cplusplus
}
#endif
#endif
