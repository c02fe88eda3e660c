/* nvector_b200.h -- public header of NVECTOR_B200, a B200 (sm_100a) native
 * N_Vector for SUNDIALS 7.x.
 *
 * Drop-in at the boundary `struct _generic_N_Vector_Ops`
 * (include/sundials/sundials_nvector.h:101-195 of the reference): C host code
 * (sundials_b200/csrc/nvector_b200.c) fills the standard ops table, so CVODE,
 * ARKODE, IDA, KINSOL and the SPGMR/SPFGMR/PCG solvers use it unmodified, and
 * calls hand-written CUDA kernels through the C ABI in b200vec.h.  This header
 * plays the role of include/nvector/nvector_cuda.h:30-215 (constructors,
 * accessors, copies, enable functions); semantics follow nvector_serial.
 *
 * Memory kinds (cf. N_VNew_Cuda / N_VNewManaged_Cuda, nvector_cuda.cu:238-345):
 *   N_VNew_B200         device array; host mirror allocated lazily (pinned) on the
 *                       first N_VGetHostArrayPointer_B200 / N_VGetArrayPointer;
 *                       use N_VCopyToDevice_B200 / N_VCopyFromDevice_B200.
 *   N_VNewManaged_B200  one cudaMallocManaged array, HOST-COHERENT: every op
 *                       synchronises the stream before returning, so host code
 *                       that reads/writes through N_VGetArrayPointer (the serial
 *                       examples and unit tests) works unchanged.
 *   N_VNewPinned_B200   one pinned, device-mapped host array (zero-copy), also
 *                       host-coherent; no page migration -- the better compat
 *                       mode for small problems.
 *   N_VMake_B200        wraps user host and/or device arrays (not owned).
 * Multi-GPU (one rank per GPU, MPIPlusX pattern nvector_mpiplusx.c:30):
 *   N_VMakeDistributed_B200 marks a vector as the local block of a global
 *   vector partitioned over the ranks of the context's NCCL communicator;
 *   reductions then allreduce, N_VGetLength returns the global length.
 */
#ifndef NVECTOR_B200_H
#define NVECTOR_B200_H

#include <stdio.h>
#include <sundials/sundials_memory.h>
#include <sundials/sundials_nvector.h>

#include "b200vec.h"

#ifdef __cplusplus
extern "C" {
#endif

#define B200_MEM_DEVICE  0
#define B200_MEM_MANAGED 1
#define B200_MEM_PINNED  2

struct _N_VectorContent_B200
{
  sunindextype length;        /* local length                                        */
  sunindextype global_length; /* == length unless distributed                        */
  sunbooleantype own_device;  /* device_data was allocated by the vector             */
  sunbooleantype own_host;    /* host_data was allocated by the vector               */
  int mem_kind;               /* B200_MEM_*                                          */
  sunbooleantype distributed; /* reductions allreduce over the ctx communicator      */
  sunrealtype* host_data;     /* host mirror (DEVICE kind) or the single array       */
  sunrealtype* device_data;   /* device address of the data                          */
  b200vec_ctx ctx;            /* shared execution context (retained)                 */
  /* vectors created by N_VNewWithMemHelp_B200: the arrays come from the user's allocator */
  SUNMemoryHelper mem_helper; /* NULL: the context's allocation cache                 */
  sunbooleantype own_helper;  /* clones own a clone of the helper (cuda:662-663)       */
  SUNMemory helper_host;      /* what mem_helper handed out for host_data / device_data */
  SUNMemory helper_device;
};
typedef struct _N_VectorContent_B200* N_VectorContent_B200;

/* ---- constructors (replace N_VNew_Serial serial:172, N_VNew_Cuda cuda:238, ...) ---- */
N_Vector N_VNewEmpty_B200(SUNContext sunctx);
N_Vector N_VNew_B200(sunindextype length, SUNContext sunctx);
N_Vector N_VNewManaged_B200(sunindextype length, SUNContext sunctx);
N_Vector N_VNewPinned_B200(sunindextype length, SUNContext sunctx);
/* explicit execution context (stream / workspace / communicator) and memory kind */
N_Vector N_VNewWithCtx_B200(sunindextype length, int mem_kind, b200vec_ctx ctx, SUNContext sunctx);
/* host and device arrays (or ONE SUNMEMTYPE_UVM array) allocated and freed through the caller's SUNMemoryHelper,
   for this vector and -- through a clone of the helper -- for its clones: N_VNewWithMemHelp_Cuda,
   nvector_cuda.cu:271-310, AllocateData :2226-2270.  Only helper->ops->alloc / dealloc / clone / destroy are used. */
N_Vector N_VNewWithMemHelp_B200(sunindextype length, sunbooleantype use_managed_mem, SUNMemoryHelper helper,
                                SUNContext sunctx);
N_Vector N_VMake_B200(sunindextype length, sunrealtype* h_vdata, sunrealtype* d_vdata, SUNContext sunctx);
/* wraps ONE user array that both sides can address (cudaMallocManaged / pinned-mapped memory), not
   owned: N_VMakeManaged_Cuda, nvector_cuda.cu:387 */
N_Vector N_VMakeManaged_B200(sunindextype length, sunrealtype* vdata, SUNContext sunctx);
N_Vector N_VMakeWithCtx_B200(sunindextype length, sunrealtype* h_vdata, sunrealtype* d_vdata, b200vec_ctx ctx,
                             SUNContext sunctx);
/* turn v (and its future clones) into the local block of a distributed vector;
 * global_length < 0: computed by an allreduce of the local lengths
 * (nvector_manyvector.c:231) */
SUNErrCode N_VMakeDistributed_B200(N_Vector v, sunindextype global_length);

/* ---- accessors / copies (nvector_cuda.h:86-104, cuda:553-599) ---- */
sunindextype N_VGetLength_B200(N_Vector v);
sunindextype N_VGetLocalLength_B200(N_Vector v);
sunrealtype* N_VGetHostArrayPointer_B200(N_Vector v);
sunrealtype* N_VGetDeviceArrayPointer_B200(N_Vector v);
void N_VSetHostArrayPointer_B200(sunrealtype* h_vdata, N_Vector v);
void N_VSetDeviceArrayPointer_B200(sunrealtype* d_vdata, N_Vector v);
sunbooleantype N_VIsManagedMemory_B200(N_Vector v);
void N_VCopyToDevice_B200(N_Vector v);
void N_VCopyFromDevice_B200(N_Vector v);
/* upload on the context's copy stream, overlapping the vector ops issued next; the context's
 * stream waits for all pending uploads at N_VCopyJoin_B200 (stream-side wait, the host goes on) */
void N_VCopyToDeviceAsync_B200(N_Vector v);
void N_VCopyJoin_B200(N_Vector v);
b200vec_ctx N_VGetCtx_B200(N_Vector v);
/* all vectors sharing v's context move to `stream` (a cudaStream_t); replaces
 * N_VSetKernelExecPolicy_Cuda cuda:514 */
SUNErrCode N_VSetStream_B200(N_Vector v, void* stream);

/* ---- ops-table entries (also callable directly, like N_V*_Serial) ---- */
N_Vector_ID N_VGetVectorID_B200(N_Vector v);
/* the ID v and its future clones report: SUNDIALS_NVEC_CUSTOM (default) or SUNDIALS_NVEC_CUDA, which the
   reference demands before it enables CVODE's fused kernels (cvode_io.c:1022-1029; include/cvode_fused_b200.h) */
SUNErrCode N_VSetVectorID_B200(N_Vector v, N_Vector_ID id);
N_Vector N_VCloneEmpty_B200(N_Vector w);
N_Vector N_VClone_B200(N_Vector w);
void N_VDestroy_B200(N_Vector v);
void N_VSpace_B200(N_Vector v, sunindextype* lrw, sunindextype* liw);
sunrealtype* N_VGetArrayPointer_B200(N_Vector v);
void N_VSetArrayPointer_B200(sunrealtype* v_data, N_Vector v);
void N_VPrint_B200(N_Vector v);
void N_VPrintFile_B200(N_Vector v, FILE* outfile);

void N_VLinearSum_B200(sunrealtype a, N_Vector x, sunrealtype b, N_Vector y, N_Vector z);
void N_VConst_B200(sunrealtype c, N_Vector z);
void N_VProd_B200(N_Vector x, N_Vector y, N_Vector z);
void N_VDiv_B200(N_Vector x, N_Vector y, N_Vector z);
void N_VScale_B200(sunrealtype c, N_Vector x, N_Vector z);
void N_VAbs_B200(N_Vector x, N_Vector z);
void N_VInv_B200(N_Vector x, N_Vector z);
void N_VAddConst_B200(N_Vector x, sunrealtype b, N_Vector z);
sunrealtype N_VDotProd_B200(N_Vector x, N_Vector y);
sunrealtype N_VMaxNorm_B200(N_Vector x);
sunrealtype N_VWrmsNorm_B200(N_Vector x, N_Vector w);
sunrealtype N_VWrmsNormMask_B200(N_Vector x, N_Vector w, N_Vector id);
sunrealtype N_VMin_B200(N_Vector x);
sunrealtype N_VWL2Norm_B200(N_Vector x, N_Vector w);
sunrealtype N_VL1Norm_B200(N_Vector x);
void N_VCompare_B200(sunrealtype c, N_Vector x, N_Vector z);
sunbooleantype N_VInvTest_B200(N_Vector x, N_Vector z);
sunbooleantype N_VConstrMask_B200(N_Vector c, N_Vector x, N_Vector m);
sunrealtype N_VMinQuotient_B200(N_Vector num, N_Vector denom);

SUNErrCode N_VLinearCombination_B200(int nvec, sunrealtype* c, N_Vector* X, N_Vector z);
SUNErrCode N_VScaleAddMulti_B200(int nvec, sunrealtype* a, N_Vector x, N_Vector* Y, N_Vector* Z);
SUNErrCode N_VDotProdMulti_B200(int nvec, N_Vector x, N_Vector* Y, sunrealtype* dotprods);
/* not an ops-table slot: z = sum c_i X_i and *sqnorm = z . z in ONE pass (used by
 * SUNClassicalGS_B200, include/sundials_iterative_b200.h) */
SUNErrCode N_VLinearCombinationSqNorm_B200(int nvec, sunrealtype* c, N_Vector* X, N_Vector z, sunrealtype* sqnorm);
/* not an ops-table slot: the integrators' error-weight vector ewt_i = 1 / (rtol |y_i| + atol_i) in ONE
 * kernel (16 B/elt instead of the 5 vector ops, 72 B/elt, of cvEwtSetSS / arkEwtSetSS).  Call it from a
 * CVEwtFn / ARKEwtFn / IDAEwtFn registered with CV/ARK/IDA...WFtolerances: same bits, same return
 * convention (0, or -1 when atolmin0 and a denominator is not positive).  vatol == NULL: scalar atol. */
int N_VEwtSet_B200(sunrealtype rtol, sunrealtype atol, N_Vector vatol, sunbooleantype atolmin0, N_Vector y, N_Vector ewt);
/* not an ops-table slot: z <- a x + z and *dot = w . z of the updated z in ONE pass (used by
 * SUNModifiedGS_B200) */
SUNErrCode N_VAxpyDot_B200(sunrealtype a, N_Vector x, N_Vector z, N_Vector w, sunrealtype* dot);
/* whole Gram-Schmidt columns as ONE chain of kernels with one host wait (coefficients stay on the device
 * between the kernels): h[nproj], norms[2] = {vk.vk before, after}; dots[nvec], *sqnorm */
SUNErrCode N_VModifiedGSSweep_B200(int nproj, N_Vector* V, N_Vector vk, sunrealtype* h, sunrealtype* norms);
SUNErrCode N_VClassicalGSStep_B200(int nvec, N_Vector x, N_Vector* Ydots, N_Vector* Xcomb, N_Vector z,
                                   sunrealtype* dots, sunrealtype* sqnorm);

SUNErrCode N_VLinearSumVectorArray_B200(int nvec, sunrealtype a, N_Vector* X, sunrealtype b, N_Vector* Y,
                                        N_Vector* Z);
SUNErrCode N_VScaleVectorArray_B200(int nvec, sunrealtype* c, N_Vector* X, N_Vector* Z);
SUNErrCode N_VConstVectorArray_B200(int nvec, sunrealtype c, N_Vector* Z);
SUNErrCode N_VWrmsNormVectorArray_B200(int nvec, N_Vector* X, N_Vector* W, sunrealtype* nrm);
SUNErrCode N_VWrmsNormMaskVectorArray_B200(int nvec, N_Vector* X, N_Vector* W, N_Vector id, sunrealtype* nrm);
SUNErrCode N_VScaleAddMultiVectorArray_B200(int nvec, int nsum, sunrealtype* a, N_Vector* X, N_Vector** Y,
                                            N_Vector** Z);
SUNErrCode N_VLinearCombinationVectorArray_B200(int nvec, int nsum, sunrealtype* c, N_Vector** X, N_Vector* Z);

/* local reductions (no communication) */
sunrealtype N_VDotProdLocal_B200(N_Vector x, N_Vector y);
sunrealtype N_VMaxNormLocal_B200(N_Vector x);
sunrealtype N_VMinLocal_B200(N_Vector x);
sunrealtype N_VL1NormLocal_B200(N_Vector x);
sunrealtype N_VWSqrSumLocal_B200(N_Vector x, N_Vector w);
sunrealtype N_VWSqrSumMaskLocal_B200(N_Vector x, N_Vector w, N_Vector id);
sunbooleantype N_VInvTestLocal_B200(N_Vector x, N_Vector z);
sunbooleantype N_VConstrMaskLocal_B200(N_Vector c, N_Vector x, N_Vector m);
sunrealtype N_VMinQuotientLocal_B200(N_Vector num, N_Vector denom);
SUNErrCode N_VDotProdMultiLocal_B200(int nvec, N_Vector x, N_Vector* Y, sunrealtype* dotprods);
SUNErrCode N_VDotProdMultiAllReduce_B200(int nvec_total, N_Vector x, sunrealtype* sum);

/* XBraid buffer ops (serial:1552-1592, cuda:2024-2081) */
SUNErrCode N_VBufSize_B200(N_Vector x, sunindextype* size);
SUNErrCode N_VBufPack_B200(N_Vector x, void* buf);
SUNErrCode N_VBufUnpack_B200(N_Vector x, void* buf);

/* ---- enable / disable fused and vector-array ops (serial:1948-2053).
 * As in every reference backend they are DISABLED (NULL) by default. ---- */
SUNErrCode N_VEnableFusedOps_B200(N_Vector v, sunbooleantype tf);
SUNErrCode N_VEnableLinearCombination_B200(N_Vector v, sunbooleantype tf);
SUNErrCode N_VEnableScaleAddMulti_B200(N_Vector v, sunbooleantype tf);
SUNErrCode N_VEnableDotProdMulti_B200(N_Vector v, sunbooleantype tf);
SUNErrCode N_VEnableLinearSumVectorArray_B200(N_Vector v, sunbooleantype tf);
SUNErrCode N_VEnableScaleVectorArray_B200(N_Vector v, sunbooleantype tf);
SUNErrCode N_VEnableConstVectorArray_B200(N_Vector v, sunbooleantype tf);
SUNErrCode N_VEnableWrmsNormVectorArray_B200(N_Vector v, sunbooleantype tf);
SUNErrCode N_VEnableWrmsNormMaskVectorArray_B200(N_Vector v, sunbooleantype tf);
SUNErrCode N_VEnableScaleAddMultiVectorArray_B200(N_Vector v, sunbooleantype tf);
SUNErrCode N_VEnableLinearCombinationVectorArray_B200(N_Vector v, sunbooleantype tf);

#ifdef __cplusplus
}
#endif
#endif /* NVECTOR_B200_H */
