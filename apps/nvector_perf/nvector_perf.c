/* nvector_perf.c -- the N_Vector performance suite as a C driver over the ops table.
 *
 * Re-host of the reference's benchmarks/nvector harness (benchmarks/nvector/test_nvector_performance.c:
 * every standard, reduction, fused and vector-array operation -- 16 N_VLinearSum cases :70-472, the
 * four N_VScale cases, the fused cases :1312-1700 and the vector-array cases :1700-2690 -- called in
 * C loops on vectors of one length, nvecs = 8, nsums = 4).  The operations are reached ONLY through
 * `v->ops->nv...`, the drop-in boundary (include/sundials/sundials_nvector.h:101-195), exactly as the
 * integrators reach them -- so the very same compiled driver times NVECTOR_B200 and the reference's
 * nvector_openmp / nvector_serial (bench.py runs it for both arms), and no interpreter sits between
 * two operations (a scalar-returning op costs ~7 us of launch + return; a Python call per op added
 * another ~5 us to round 1's per-op figures).
 *
 * Differences from the reference harness, all deliberate: fused ops are
 * ENABLED (the reference drivers leave them off), seeds are fixed by the caller, timing is done by
 * the caller (CUDA events for the device arm, the host clock for the CPU arm), in-place forms act on
 * scratch vectors so that a step leaves its inputs unchanged.
 *
 * Vector indices wrap modulo nvecs, so the same op list runs with fewer vectors (the length
 * sweep uses nvecs = 2 at 2^30, where a vector is 8 GiB).
 *
 * bytes per element = the algorithmic count of SURVEY.md section 8d (each distinct operand read
 * once, each output written once).
 */
#include <stdlib.h>
#include <string.h>

#include <sundials/sundials_nvector.h>

#define MAXV 16
#define MAXS 8

typedef struct nvperf_suite_s
{
  int nv, ns;
  N_Vector X[MAXV], Y[MAXV], Z[MAXV], YS[MAXV];
  N_Vector S, T, W, ID, CN;
  N_Vector YYrow[MAXS][MAXV], ZZrow[MAXS][MAXV];
  N_Vector* YY[MAXS];
  N_Vector* ZZ[MAXS];
  sunrealtype c8[MAXV], c8_one[MAXV], c4[MAXS], cs8[MAXV];
  sunrealtype dots[MAXV], nrm[MAXV];
  double result[64];
  int err;
}* nvperf_suite;

enum
{
  OP_SCALE_COPY, OP_LS_1A, OP_LS_1B, OP_LS_1C, OP_LS_2A, OP_LS_2B, OP_LS_2C, OP_LS_3, OP_LS_4A, OP_LS_4B, OP_LS_5A,
  OP_LS_5B, OP_LS_6A, OP_LS_6B, OP_LS_7, OP_LS_8, OP_LS_9, OP_CONST, OP_PROD, OP_DIV, OP_SCALE_INPLACE, OP_SCALE_NEG,
  OP_SCALE, OP_ABS, OP_INV, OP_ADDCONST, OP_DOT, OP_MAXNORM, OP_WRMS, OP_WRMSMASK, OP_MIN, OP_WL2, OP_L1, OP_COMPARE,
  OP_INVTEST, OP_CONSTRMASK, OP_MINQUOT, OP_LC_1, OP_LC_2, OP_LC_3, OP_SAM_1, OP_SAM_2, OP_DOTMULTI, OP_LSVA, OP_SVA,
  OP_CVA, OP_WRMSVA, OP_WRMSMASKVA, OP_SAMVA, OP_LCVA, OP_DOTLOCAL, OP_MAXNORMLOCAL, OP_WSQRSUMLOCAL, OP_DOTMULTILOCAL,
  OP_RESULT, OP_COUNT
};

static const char* const kNames[OP_COUNT] = {
  "N_VScale-2(copy)", "N_VLinearSum-1a", "N_VLinearSum-1b", "N_VLinearSum-1c", "N_VLinearSum-2a", "N_VLinearSum-2b",
  "N_VLinearSum-2c", "N_VLinearSum-3", "N_VLinearSum-4a", "N_VLinearSum-4b", "N_VLinearSum-5a", "N_VLinearSum-5b",
  "N_VLinearSum-6a", "N_VLinearSum-6b", "N_VLinearSum-7", "N_VLinearSum-8", "N_VLinearSum-9", "N_VConst", "N_VProd",
  "N_VDiv", "N_VScale-1(inplace)", "N_VScale-3(neg)", "N_VScale-4", "N_VAbs", "N_VInv", "N_VAddConst", "N_VDotProd",
  "N_VMaxNorm", "N_VWrmsNorm", "N_VWrmsNormMask", "N_VMin", "N_VWL2Norm", "N_VL1Norm", "N_VCompare", "N_VInvTest",
  "N_VConstrMask", "N_VMinQuotient", "N_VLinearCombination-1", "N_VLinearCombination-2", "N_VLinearCombination-3",
  "N_VScaleAddMulti-1", "N_VScaleAddMulti-2", "N_VDotProdMulti", "N_VLinearSumVectorArray", "N_VScaleVectorArray",
  "N_VConstVectorArray", "N_VWrmsNormVectorArray", "N_VWrmsNormMaskVectorArray", "N_VScaleAddMultiVectorArray",
  "N_VLinearCombinationVectorArray", "N_VDotProdLocal", "N_VMaxNormLocal", "N_VWSqrSumLocal", "N_VDotProdMultiLocal",
  "N_VWrmsNorm(result)"};

int nvperf_num_ops(void) { return OP_COUNT; }
const char* nvperf_op_name(int op) { return (op >= 0 && op < OP_COUNT) ? kNames[op] : ""; }

double nvperf_op_bytes_per_elt(nvperf_suite s, int op)
{
  const int nv = s->nv, ns = s->ns;
  switch (op)
  {
  case OP_CONST: case OP_MAXNORM: case OP_MIN: case OP_L1: case OP_MAXNORMLOCAL: return 8;
  case OP_SCALE_COPY: case OP_SCALE_INPLACE: case OP_SCALE_NEG: case OP_SCALE: case OP_ABS: case OP_INV:
  case OP_ADDCONST: case OP_DOT: case OP_WRMS: case OP_WL2: case OP_COMPARE: case OP_INVTEST: case OP_MINQUOT:
  case OP_DOTLOCAL: case OP_WSQRSUMLOCAL: case OP_RESULT: return 16;
  case OP_WRMSMASK: case OP_CONSTRMASK: case OP_PROD: case OP_DIV: return 24;
  case OP_LC_1: case OP_LC_2: case OP_LC_3: case OP_DOTMULTI: case OP_DOTMULTILOCAL: return 8.0 * (nv + 1);
  case OP_SAM_1: case OP_SAM_2: return 8.0 * (2 * nv + 1);
  case OP_LSVA: return 24.0 * nv;
  case OP_SVA: return 16.0 * nv;
  case OP_CVA: return 8.0 * nv;
  case OP_WRMSVA: return 16.0 * nv;
  case OP_WRMSMASKVA: return 8.0 * (2 * nv + 1);
  case OP_SAMVA: return 8.0 * (nv + 2 * nv * ns);
  case OP_LCVA: return 8.0 * (nv * ns + nv);
  default: return 24; /* the N_VLinearSum cases */
  }
}

/* 1 if the op hands a scalar (or an array of scalars) back to the host, i.e. blocks */
int nvperf_op_returns_scalar(int op)
{
  switch (op)
  {
  case OP_DOT: case OP_MAXNORM: case OP_WRMS: case OP_WRMSMASK: case OP_MIN: case OP_WL2: case OP_L1: case OP_INVTEST:
  case OP_CONSTRMASK: case OP_MINQUOT: case OP_DOTMULTI: case OP_WRMSVA: case OP_WRMSMASKVA: case OP_DOTLOCAL:
  case OP_MAXNORMLOCAL: case OP_WSQRSUMLOCAL: case OP_DOTMULTILOCAL: case OP_RESULT: return 1;
  default: return 0;
  }
}

/* X, Y, Z: nv handles each; YY, ZZ: ns * nv handles, [j * nv + i].  The handles stay owned by the caller. */
nvperf_suite nvperf_create(N_Vector* X, N_Vector* Y, N_Vector* Z, N_Vector S, N_Vector T, N_Vector W, N_Vector ID,
                           N_Vector CN, N_Vector* YY, N_Vector* ZZ, int nv, int ns)
{
  if (nv < 2 || nv > MAXV || ns < 1 || ns > MAXS) return NULL;
  nvperf_suite s = (nvperf_suite)calloc(1, sizeof *s);
  if (!s) return NULL;
  s->nv = nv;
  s->ns = ns;
  for (int i = 0; i < nv; i++)
  {
    s->X[i]  = X[i];
    s->Y[i]  = Y[i];
    s->Z[i]  = Z[i];
    s->YS[i] = (i == 0) ? S : Y[i]; /* fused in-place forms act on the scratch vector */
    s->c8[i]     = 0.11 * (i + 1) * ((i & 1) ? -1.0 : 1.0);
    s->c8_one[i] = (i == 0) ? 1.0 : 0.11 * (i + 1);
    s->cs8[i]    = 1.0 + 0.01 * i;
  }
  for (int j = 0; j < ns; j++)
  {
    s->c4[j] = 0.21 * (j + 1) * ((j & 1) ? -1.0 : 1.0);
    for (int i = 0; i < nv; i++)
    {
      s->YYrow[j][i] = YY[j * nv + i];
      s->ZZrow[j][i] = ZZ[j * nv + i];
    }
    s->YY[j] = s->YYrow[j];
    s->ZZ[j] = s->ZZrow[j];
  }
  s->S = S; s->T = T; s->W = W; s->ID = ID; s->CN = CN;
  /* every fused slot must be filled: this suite times the fused kernels, not the generic fallback */
  N_Vector_Ops o = X[0]->ops;
  if (!o->nvlinearcombination || !o->nvscaleaddmulti || !o->nvdotprodmulti || !o->nvlinearsumvectorarray ||
      !o->nvscalevectorarray || !o->nvconstvectorarray || !o->nvwrmsnormvectorarray || !o->nvwrmsnormmaskvectorarray ||
      !o->nvscaleaddmultivectorarray || !o->nvlinearcombinationvectorarray || !o->nvdotprodmultilocal)
  {
    free(s);
    return NULL;
  }
  return s;
}

void nvperf_destroy(nvperf_suite s) { free(s); }
double nvperf_result(nvperf_suite s, int op) { return (op >= 0 && op < OP_COUNT) ? s->result[op] : 0.0; }
int nvperf_error(nvperf_suite s) { return s->err; }

static void run_one(nvperf_suite s, int op)
{
  const double a = 0.37, b = -1.63;
  N_Vector *X = s->X, *Y = s->Y, *Z = s->Z;
  N_Vector S = s->S, T = s->T, W = s->W;
  const int nv = s->nv, ns = s->ns;
  N_Vector_Ops o = X[0]->ops;
  int e = 0;
  switch (op)
  {
  /* N_VLinearSum cases 1a..9 (test_nvector_performance.c:70-472) */
  case OP_SCALE_COPY: o->nvscale(1.0, Y[0 % nv], S); break;
  case OP_LS_1A: o->nvlinearsum(1.0, X[0 % nv], 1.0, S, S); break;
  case OP_LS_1B: o->nvlinearsum(-1.0, X[0 % nv], 1.0, S, S); break;
  case OP_LS_1C: o->nvlinearsum(a, X[0 % nv], 1.0, S, S); break;
  case OP_LS_2A: o->nvlinearsum(1.0, S, 1.0, Y[0 % nv], S); break;
  case OP_LS_2B: o->nvlinearsum(1.0, S, -1.0, Y[0 % nv], S); break;
  case OP_LS_2C: o->nvlinearsum(1.0, S, b, Y[0 % nv], S); break;
  case OP_LS_3: o->nvlinearsum(1.0, X[0 % nv], 1.0, Y[0 % nv], Z[0 % nv]); break;
  case OP_LS_4A: o->nvlinearsum(1.0, X[1 % nv], -1.0, Y[1 % nv], Z[1 % nv]); break;
  case OP_LS_4B: o->nvlinearsum(-1.0, X[2 % nv], 1.0, Y[2 % nv], Z[2 % nv]); break;
  case OP_LS_5A: o->nvlinearsum(1.0, X[3 % nv], b, Y[3 % nv], Z[3 % nv]); break;
  case OP_LS_5B: o->nvlinearsum(a, X[4 % nv], 1.0, Y[4 % nv], Z[4 % nv]); break;
  case OP_LS_6A: o->nvlinearsum(-1.0, X[5 % nv], b, Y[5 % nv], Z[5 % nv]); break;
  case OP_LS_6B: o->nvlinearsum(a, X[6 % nv], -1.0, Y[6 % nv], Z[6 % nv]); break;
  case OP_LS_7: o->nvlinearsum(a, X[7 % nv], a, Y[7 % nv], Z[7 % nv]); break;
  case OP_LS_8: o->nvlinearsum(a, X[0 % nv], -a, Y[1 % nv], Z[0 % nv]); break;
  case OP_LS_9: o->nvlinearsum(a, X[1 % nv], b, Y[2 % nv], Z[1 % nv]); break;
  case OP_CONST: o->nvconst(1.5, T); break;
  case OP_PROD: o->nvprod(X[2 % nv], Y[3 % nv], Z[2 % nv]); break;
  case OP_DIV: o->nvdiv(X[3 % nv], Y[4 % nv], Z[3 % nv]); break;
  case OP_SCALE_INPLACE: o->nvscale(1.0009765625, S, S); break;
  case OP_SCALE_NEG: o->nvscale(-1.0, X[4 % nv], Z[4 % nv]); break;
  case OP_SCALE: o->nvscale(a, X[5 % nv], Z[5 % nv]); break;
  case OP_ABS: o->nvabs(X[6 % nv], Z[6 % nv]); break;
  case OP_INV: o->nvinv(X[7 % nv], Z[7 % nv]); break;
  case OP_ADDCONST: o->nvaddconst(X[0 % nv], b, Z[0 % nv]); break;
  case OP_DOT: s->result[op] = o->nvdotprod(X[1 % nv], Y[1 % nv]); break;
  case OP_MAXNORM: s->result[op] = o->nvmaxnorm(X[2 % nv]); break;
  case OP_WRMS: s->result[op] = o->nvwrmsnorm(X[3 % nv], W); break;
  case OP_WRMSMASK: s->result[op] = o->nvwrmsnormmask(X[4 % nv], W, s->ID); break;
  case OP_MIN: s->result[op] = o->nvmin(X[5 % nv]); break;
  case OP_WL2: s->result[op] = o->nvwl2norm(X[6 % nv], W); break;
  case OP_L1: s->result[op] = o->nvl1norm(X[7 % nv]); break;
  case OP_COMPARE: o->nvcompare(0.75, X[0 % nv], Z[0 % nv]); break;
  case OP_INVTEST: s->result[op] = o->nvinvtest(X[1 % nv], Z[1 % nv]); break;
  case OP_CONSTRMASK: s->result[op] = o->nvconstrmask(s->CN, X[2 % nv], Z[2 % nv]); break;
  case OP_MINQUOT: s->result[op] = o->nvminquotient(X[3 % nv], Y[3 % nv]); break;
  /* fused (:1312-1700); -1 / -2 are the in-place forms */
  case OP_LC_1: e = o->nvlinearcombination(nv, s->c8_one, s->YS, S); break;
  case OP_LC_2: e = o->nvlinearcombination(nv, s->c8, s->YS, S); break;
  case OP_LC_3: e = o->nvlinearcombination(nv, s->c8, X, T); break;
  case OP_SAM_1: e = o->nvscaleaddmulti(nv, s->c8, X[0 % nv], Z, Z); break;
  case OP_SAM_2: e = o->nvscaleaddmulti(nv, s->c8, X[1 % nv], Y, Z); break;
  case OP_DOTMULTI: e = o->nvdotprodmulti(nv, X[2 % nv], Y, s->dots); s->result[op] = s->dots[0]; break;
  /* vector arrays (:1700-2690) */
  case OP_LSVA: e = o->nvlinearsumvectorarray(nv, a, X, b, Y, Z); break;
  case OP_SVA: e = o->nvscalevectorarray(nv, s->cs8, X, Z); break;
  case OP_CVA: e = o->nvconstvectorarray(nv, 0.5, Z); break;
  case OP_WRMSVA: e = o->nvwrmsnormvectorarray(nv, X, Y, s->nrm); s->result[op] = s->nrm[0]; break;
  case OP_WRMSMASKVA: e = o->nvwrmsnormmaskvectorarray(nv, X, Y, s->ID, s->nrm); s->result[op] = s->nrm[0]; break;
  case OP_SAMVA: e = o->nvscaleaddmultivectorarray(nv, ns, s->c4, X, s->YY, s->ZZ); break;
  case OP_LCVA: e = o->nvlinearcombinationvectorarray(nv, ns, s->c4, s->YY, Z); break;
  /* local reductions (no communication even on a distributed vector) */
  case OP_DOTLOCAL: s->result[op] = o->nvdotprodlocal(X[4 % nv], Y[4 % nv]); break;
  case OP_MAXNORMLOCAL: s->result[op] = o->nvmaxnormlocal(X[5 % nv]); break;
  case OP_WSQRSUMLOCAL: s->result[op] = o->nvwsqrsumlocal(X[6 % nv], W); break;
  case OP_DOTMULTILOCAL: e = o->nvdotprodmultilocal(nv, X[7 % nv], Y, s->dots); s->result[op] = s->dots[0]; break;
  /* the step's result: a checksum of an output vector */
  case OP_RESULT: s->result[op] = o->nvwrmsnorm(Z[1 % nv], W); break;
  default: e = -1;
  }
  if (e) s->err = e;
}

void nvperf_run_op(nvperf_suite s, int op, int reps)
{
  for (int r = 0; r < reps; r++) run_one(s, op);
}

void nvperf_run_step(nvperf_suite s, int steps)
{
  for (int k = 0; k < steps; k++)
    for (int op = 0; op < OP_COUNT; op++) run_one(s, op);
}
