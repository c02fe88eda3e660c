/* tests/c/shim_cuda/nvector/nvector_cuda.h -- TEST SHIM (our code, not a reference file).
 *
 * Lets the reference's CUDA example programs (examples/cvode/cuda/cvAdvDiff_kry_cuda.cu, ...)
 * compile UNMODIFIED on NVECTOR_B200: they include <nvector/nvector_cuda.h> and call the
 * N_V*_Cuda constructors / accessors (include/nvector/nvector_cuda.h:60-215 of the reference);
 * this header, found first on the include path, maps those names onto the B200 vector's
 * counterparts.  The execution-policy classes (include/sundials/sundials_cuda_policies.hpp:71-235)
 * shrink to "which stream": NVECTOR_B200 picks its own launch geometry.
 */
#ifndef B200_SHIM_NVECTOR_CUDA_H
#define B200_SHIM_NVECTOR_CUDA_H

#include <cuda_runtime.h>

#include "nvector_b200.h"

/* the constructors also make the vector answer N_VGetVectorID with SUNDIALS_NVEC_CUDA, as the vector it
   stands in for does (nvector_cuda.h:114): CVodeSetUseIntegratorFusedKernels tests it (cvode_io.c:1022-1029) */
static inline N_Vector b200_shim_as_cuda(N_Vector v)
{
  if (v) N_VSetVectorID_B200(v, SUNDIALS_NVEC_CUDA);
  return v;
}
#define N_VNew_Cuda(n, ctx)              b200_shim_as_cuda(N_VNew_B200(n, ctx))
#define N_VNewManaged_Cuda(n, ctx)       b200_shim_as_cuda(N_VNewManaged_B200(n, ctx))
#define N_VNewEmpty_Cuda(ctx)            b200_shim_as_cuda(N_VNewEmpty_B200(ctx))
#define N_VMake_Cuda(n, h, d, ctx)       b200_shim_as_cuda(N_VMake_B200(n, h, d, ctx))
#define N_VMakeManaged_Cuda(n, p, ctx)   b200_shim_as_cuda(N_VMakeManaged_B200(n, p, ctx))
#define N_VNewWithMemHelp_Cuda(n, managed, helper, ctx) b200_shim_as_cuda(N_VNewWithMemHelp_B200(n, managed, helper, ctx))
#define N_VClone_Cuda                 N_VClone_B200
#define N_VCloneEmpty_Cuda            N_VCloneEmpty_B200
#define N_VDestroy_Cuda               N_VDestroy_B200
#define N_VGetHostArrayPointer_Cuda   N_VGetHostArrayPointer_B200
#define N_VGetDeviceArrayPointer_Cuda N_VGetDeviceArrayPointer_B200
#define N_VCopyToDevice_Cuda          N_VCopyToDevice_B200
#define N_VCopyFromDevice_Cuda        N_VCopyFromDevice_B200
#define N_VIsManagedMemory_Cuda       N_VIsManagedMemory_B200
#define N_VEnableFusedOps_Cuda        N_VEnableFusedOps_B200

#ifdef __cplusplus
class SUNCudaExecPolicy
{
public:
  explicit SUNCudaExecPolicy(cudaStream_t s) : stream_(s) {}
  const cudaStream_t* stream() const { return &stream_; }

private:
  cudaStream_t stream_;
};
class SUNCudaThreadDirectExecPolicy : public SUNCudaExecPolicy
{
public:
  SUNCudaThreadDirectExecPolicy(int /*blockDim*/, cudaStream_t s = 0) : SUNCudaExecPolicy(s) {}
};
class SUNCudaGridStrideExecPolicy : public SUNCudaExecPolicy
{
public:
  SUNCudaGridStrideExecPolicy(int /*blockDim*/, int /*gridDim*/, cudaStream_t s = 0) : SUNCudaExecPolicy(s) {}
};
class SUNCudaBlockReduceExecPolicy : public SUNCudaExecPolicy
{
public:
  SUNCudaBlockReduceExecPolicy(int /*blockDim*/, int /*gridDim*/ = 0, cudaStream_t s = 0) : SUNCudaExecPolicy(s) {}
};
class SUNCudaBlockReduceAtomicExecPolicy : public SUNCudaExecPolicy
{
public:
  SUNCudaBlockReduceAtomicExecPolicy(int /*blockDim*/, int /*gridDim*/ = 0, cudaStream_t s = 0) : SUNCudaExecPolicy(s) {}
};

/* N_VSetKernelExecPolicy_Cuda (nvector_cuda.cu:514): all that carries over is the stream */
static inline SUNErrCode N_VSetKernelExecPolicy_Cuda(N_Vector x, SUNCudaExecPolicy* stream_exec_policy,
                                                     SUNCudaExecPolicy* /*reduce_exec_policy*/)
{
  return N_VSetStream_B200(x, stream_exec_policy ? (void*)*stream_exec_policy->stream() : (void*)0);
}
#endif

#endif
