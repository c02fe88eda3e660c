/* tests/c/shim_serial_managed/nvector/nvector_cuda.h -- TEST SHIM (our code, not a reference file).
 *
 * The ORACLE build of the reference's CUDA example programs: <nvector/nvector_cuda.h> resolves to this
 * header, which maps the N_V*_Cuda names onto the reference's own nvector_serial over cudaMallocManaged
 * arrays.  Every vector operation is then the reference's CPU code (nvector_serial.c, unmodified); only
 * the example's own RHS kernels run on the GPU, on the same arrays (the examples synchronise the device
 * after each kernel, examples/cvode/cuda/cvAdvDiff_diag_cuda.cu:334).  The program's output is what the
 * NVECTOR_B200 build of the same source must print bit for bit: streaming ops are bit-identical to
 * nvector_serial and the reductions of these small problems take the exact-order path.
 *
 * Clones must also live in managed memory (CVODE clones its work vectors from the template), so the
 * template's clone / destroy ops are replaced; N_VCloneEmpty_Serial copies the ops table, which carries
 * the replacement to every descendant.  N_VGetVectorID answers SUNDIALS_NVEC_CUDA like the vector this
 * stands in for, so that CVodeSetUseIntegratorFusedKernels (cvode_io.c:1022-1029) admits it -- the fused
 * functions are then the reference's CPU stubs (libsundials_cvode_fused_stubs.so).
 */
#ifndef B200_SHIM_SERIAL_MANAGED_NVECTOR_CUDA_H
#define B200_SHIM_SERIAL_MANAGED_NVECTOR_CUDA_H

#include <cuda_runtime.h>
#include <nvector/nvector_serial.h>

static inline N_Vector_ID shimsm_id(N_Vector) { return SUNDIALS_NVEC_CUDA; }
static inline N_Vector shimsm_wrap(N_Vector v);

static inline N_Vector shimsm_clone(N_Vector w)
{
  N_Vector v = N_VCloneEmpty_Serial(w);
  if (!v) return NULL;
  sunrealtype* p = NULL;
  const sunindextype n = NV_LENGTH_S(w);
  if (n > 0 && cudaMallocManaged((void**)&p, (size_t)n * sizeof(sunrealtype)) != cudaSuccess)
  {
    N_VDestroy_Serial(v);
    return NULL;
  }
  NV_DATA_S(v)     = p;
  NV_OWN_DATA_S(v) = SUNFALSE; /* freed by shimsm_destroy, not by free() */
  return v;
}

static inline void shimsm_destroy(N_Vector v)
{
  if (!v) return;
  if (v->content && NV_DATA_S(v)) cudaFree(NV_DATA_S(v));
  N_VDestroy_Serial(v);
}

static inline N_Vector shimsm_wrap(N_Vector v)
{
  if (!v) return NULL;
  v->ops->nvclone       = shimsm_clone;
  v->ops->nvdestroy     = shimsm_destroy;
  v->ops->nvgetvectorid = shimsm_id;
  return v;
}

static inline N_Vector N_VNew_Cuda(sunindextype n, SUNContext ctx)
{
  sunrealtype* p = NULL;
  if (n > 0 && cudaMallocManaged((void**)&p, (size_t)n * sizeof(sunrealtype)) != cudaSuccess) return NULL;
  return shimsm_wrap(N_VMake_Serial(n, p, ctx));
}
#define N_VNewManaged_Cuda N_VNew_Cuda
static inline N_Vector N_VClone_Cuda(N_Vector w) { return N_VClone(w); }
static inline void N_VDestroy_Cuda(N_Vector v) { N_VDestroy(v); }
static inline sunrealtype* N_VGetHostArrayPointer_Cuda(N_Vector v) { return NV_DATA_S(v); }
static inline sunrealtype* N_VGetDeviceArrayPointer_Cuda(N_Vector v) { return NV_DATA_S(v); }
static inline void N_VCopyToDevice_Cuda(N_Vector) {}
static inline void N_VCopyFromDevice_Cuda(N_Vector) {}
static inline sunbooleantype N_VIsManagedMemory_Cuda(N_Vector) { return SUNTRUE; }
#define N_VEnableFusedOps_Cuda N_VEnableFusedOps_Serial

/* execution policies mean nothing to a CPU vector: accepted and ignored (the oracle runs are made with
   CUDA_LAUNCH_BLOCKING=1, so that the example's kernels have finished when the host code reads the arrays) */
class SUNCudaExecPolicy
{
public:
  virtual ~SUNCudaExecPolicy() {}
};
class SUNCudaThreadDirectExecPolicy : public SUNCudaExecPolicy
{
public:
  SUNCudaThreadDirectExecPolicy(int, cudaStream_t = 0) {}
};
class SUNCudaGridStrideExecPolicy : public SUNCudaExecPolicy
{
public:
  SUNCudaGridStrideExecPolicy(int, int, cudaStream_t = 0) {}
};
class SUNCudaBlockReduceExecPolicy : public SUNCudaExecPolicy
{
public:
  SUNCudaBlockReduceExecPolicy(int, int = 0, cudaStream_t = 0) {}
};
class SUNCudaBlockReduceAtomicExecPolicy : public SUNCudaExecPolicy
{
public:
  SUNCudaBlockReduceAtomicExecPolicy(int, int = 0, cudaStream_t = 0) {}
};
static inline SUNErrCode N_VSetKernelExecPolicy_Cuda(N_Vector, SUNCudaExecPolicy*, SUNCudaExecPolicy*) { return SUN_SUCCESS; }

#endif
