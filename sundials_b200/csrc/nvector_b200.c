/* nvector_b200.c -- NVECTOR_B200: C host code that fills the standard SUNDIALS
 * N_Vector_Ops table (include/sundials/sundials_nvector.h:101-195 of the
 * reference) and forwards every operation to the hand-written sm_100a kernels
 * through the C ABI of include/b200vec.h.  No arithmetic on vector data happens
 * in this file and there is no CPU fallback: if a kernel launch fails the
 * process is stopped with a diagnostic (void/scalar ops have no error channel,
 * sundials_nvector.h; the reference CUDA vector silently ignores such errors,
 * src/sundials/sundials_cuda.h:61-76).
 *
 * Semantics follow nvector_serial.c ("serial:<line>"); structure follows the
 * role of nvector_cuda.cu ("cuda:<line>") but none of its code.
 *
 * The vector does not need libsundials_core at link time: the generic struct
 * and ops table are allocated here exactly as N_VNewEmpty / N_VCopyOps /
 * N_VFreeEmpty do (src/sundials/sundials_nvector.c:46-278 -- malloc'd struct +
 * malloc'd ops table, freed with free()), so the shared library loads into any
 * SUNDIALS 7.x application (or a bare test harness) as a plug-in.
 */
#include "nvector_b200.h"

#include <sundials/priv/sundials_context_impl.h>

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NVC(v)   ((N_VectorContent_B200)((v)->content))
#define NLEN(v)  (NVC(v)->length)
#define NDEV(v)  (NVC(v)->device_data)
#define NCTX(v)  (NVC(v)->ctx)
#define NDIST(v) (NVC(v)->distributed && b200vec_comm_size(NVC(v)->ctx) > 1)

#define MAX_STACK_VECS 64

/* ----------------------------------------------------------------------
 * failure handling
 * -------------------------------------------------------------------- */
/* A failed kernel launch / CUDA call inside a void or scalar op has no return channel
   (sundials_nvector.h).  It is (1) reported on stderr, (2) recorded in the vector's SUNContext
   (sunctx->last_err = SUN_ERR_EXT_FAIL, what SUNCheckLastErr() inspects), and scalar ops return
   NaN so that the calling integrator fails through its own error path (a NaN norm fails every
   error test).  Set B200VEC_ABORT_ON_ERROR=1 to stop the process at the first failure instead. */
static void b200_fail(N_Vector v, const char* where, int rc)
{
  static int reported = 0;
  if (reported < 8)
  {
    reported++;
    fprintf(stderr, "[nvector_b200] ERROR in %s: error %d: %s\n", where, rc, b200vec_last_error());
    fflush(stderr);
  }
  if (v && v->sunctx) v->sunctx->last_err = SUN_ERR_EXT_FAIL;
  const char* e = getenv("B200VEC_ABORT_ON_ERROR");
  if (e && e[0] && e[0] != '0') abort();
}

#define CHECK_ON(v, call)                                \
  do {                                                   \
    int rc_ = (call);                                    \
    if (rc_ != B200VEC_OK) b200_fail((v), __func__, rc_); \
  } while (0)

static SUNErrCode map_err(int rc)
{
  switch (rc)
  {
  case B200VEC_OK: return SUN_SUCCESS;
  case B200VEC_ERR_ARG: return SUN_ERR_ARG_OUTOFRANGE;
  case B200VEC_ERR_NOMEM: return SUN_ERR_MALLOC_FAIL;
  default: return SUN_ERR_EXT_FAIL;
  }
}

/* host-coherent kinds (managed / pinned): the host may look at the data right
   after any op returns, so streaming ops end with a stream sync */
static void coherent_sync(N_Vector v)
{
  if (NVC(v)->mem_kind != B200_MEM_DEVICE) CHECK_ON(v, b200vec_ctx_sync(NCTX(v)));
}

/* SUNRsqrt (include/sundials/sundials_math.h:83) */
static sunrealtype rsqrt_guard(sunrealtype x) { return (x <= 0.0) ? 0.0 : sqrt(x); }

/* ----------------------------------------------------------------------
 * generic object plumbing (what N_VNewEmpty/N_VCopyOps/N_VFreeEmpty do)
 * -------------------------------------------------------------------- */
static N_Vector generic_new(SUNContext sunctx)
{
  N_Vector v = (N_Vector)malloc(sizeof *v);
  if (!v) return NULL;
  v->ops = (N_Vector_Ops)calloc(1, sizeof *(v->ops)); /* all slots NULL */
  if (!v->ops)
  {
    free(v);
    return NULL;
  }
  v->content = NULL;
  v->sunctx  = sunctx;
  return v;
}

static void attach_ops(N_Vector v)
{
  N_Vector_Ops o = v->ops;
  /* constructors, destructors, utilities */
  o->nvgetvectorid           = N_VGetVectorID_B200;
  o->nvclone                 = N_VClone_B200;
  o->nvcloneempty            = N_VCloneEmpty_B200;
  o->nvdestroy               = N_VDestroy_B200;
  o->nvspace                 = N_VSpace_B200;
  o->nvgetarraypointer       = N_VGetArrayPointer_B200;
  o->nvgetdevicearraypointer = N_VGetDeviceArrayPointer_B200;
  o->nvsetarraypointer       = N_VSetArrayPointer_B200;
  o->nvgetlength             = N_VGetLength_B200;
  o->nvgetlocallength        = N_VGetLocalLength_B200;
  /* standard operations */
  o->nvlinearsum    = N_VLinearSum_B200;
  o->nvconst        = N_VConst_B200;
  o->nvprod         = N_VProd_B200;
  o->nvdiv          = N_VDiv_B200;
  o->nvscale        = N_VScale_B200;
  o->nvabs          = N_VAbs_B200;
  o->nvinv          = N_VInv_B200;
  o->nvaddconst     = N_VAddConst_B200;
  o->nvdotprod      = N_VDotProd_B200;
  o->nvmaxnorm      = N_VMaxNorm_B200;
  o->nvwrmsnorm     = N_VWrmsNorm_B200;
  o->nvwrmsnormmask = N_VWrmsNormMask_B200;
  o->nvmin          = N_VMin_B200;
  o->nvwl2norm      = N_VWL2Norm_B200;
  o->nvl1norm       = N_VL1Norm_B200;
  o->nvcompare      = N_VCompare_B200;
  o->nvinvtest      = N_VInvTest_B200;
  o->nvconstrmask   = N_VConstrMask_B200;
  o->nvminquotient  = N_VMinQuotient_B200;
  /* fused and vector-array operations: disabled (NULL) by default, as in every
     reference backend (serial:127, cuda:174) -- see N_VEnableFusedOps_B200 */
  /* local reductions */
  o->nvdotprodlocal     = N_VDotProdLocal_B200;
  o->nvmaxnormlocal     = N_VMaxNormLocal_B200;
  o->nvminlocal         = N_VMinLocal_B200;
  o->nvl1normlocal      = N_VL1NormLocal_B200;
  o->nvinvtestlocal     = N_VInvTestLocal_B200;
  o->nvconstrmasklocal  = N_VConstrMaskLocal_B200;
  o->nvminquotientlocal = N_VMinQuotientLocal_B200;
  o->nvwsqrsumlocal     = N_VWSqrSumLocal_B200;
  o->nvwsqrsummasklocal = N_VWSqrSumMaskLocal_B200;
  /* single-buffer reductions */
  o->nvdotprodmultilocal     = N_VDotProdMultiLocal_B200;
  o->nvdotprodmultiallreduce = N_VDotProdMultiAllReduce_B200;
  /* XBraid buffers */
  o->nvbufsize   = N_VBufSize_B200;
  o->nvbufpack   = N_VBufPack_B200;
  o->nvbufunpack = N_VBufUnpack_B200;
  /* debugging */
  o->nvprint     = N_VPrint_B200;
  o->nvprintfile = N_VPrintFile_B200;
}

static N_Vector new_shell(SUNContext sunctx, b200vec_ctx ctx)
{
  N_Vector v = generic_new(sunctx);
  if (!v) return NULL;
  attach_ops(v);
  N_VectorContent_B200 c = (N_VectorContent_B200)calloc(1, sizeof *c);
  if (!c)
  {
    free(v->ops);
    free(v);
    return NULL;
  }
  if (!ctx && b200vec_ctx_default(&ctx) != B200VEC_OK)
  {
    fprintf(stderr, "[nvector_b200] cannot create the execution context: %s\n", b200vec_last_error());
    free(c);
    free(v->ops);
    free(v);
    return NULL;
  }
  b200vec_ctx_retain(ctx);
  c->ctx        = ctx;
  c->mem_kind   = B200_MEM_DEVICE;
  v->content    = c;
  return v;
}

/* the user's allocator (N_VNewWithMemHelp_B200): through the helper's ops table, so that no symbol of
   sundials_core is needed (SUNMemoryHelper_Alloc is ops->alloc plus argument checks, sundials_memory.c:137) */
static int helper_alloc_data(N_VectorContent_B200 c, size_t bytes)
{
  SUNMemoryHelper h = c->mem_helper;
  if (c->mem_kind == B200_MEM_MANAGED)
  {
    if (h->ops->alloc(h, &c->helper_device, bytes, SUNMEMTYPE_UVM, h->queue) || !c->helper_device) return B200VEC_ERR_NOMEM;
    c->device_data = c->host_data = (sunrealtype*)c->helper_device->ptr;
    return B200VEC_OK;
  }
  if (h->ops->alloc(h, &c->helper_host, bytes, SUNMEMTYPE_HOST, h->queue) || !c->helper_host) return B200VEC_ERR_NOMEM;
  if (h->ops->alloc(h, &c->helper_device, bytes, SUNMEMTYPE_DEVICE, h->queue) || !c->helper_device) return B200VEC_ERR_NOMEM;
  c->host_data   = (sunrealtype*)c->helper_host->ptr;
  c->device_data = (sunrealtype*)c->helper_device->ptr;
  return B200VEC_OK;
}

/* SUNMemoryHelper_Clone (sundials_memory.c:235-253) without sundials_core: the helper's own clone op, or --
   for a helper without content -- a copy of the object and of its ops table */
static SUNMemoryHelper helper_clone(SUNMemoryHelper h)
{
  if (h->ops->clone) return h->ops->clone(h);
  if (h->content) return NULL;
  SUNMemoryHelper k = (SUNMemoryHelper)malloc(sizeof *k);
  if (!k) return NULL;
  *k     = *h;
  k->ops = (SUNMemoryHelper_Ops)malloc(sizeof *(k->ops));
  if (!k->ops)
  {
    free(k);
    return NULL;
  }
  *(k->ops) = *(h->ops);
  return k;
}

static void helper_release(N_VectorContent_B200 c)
{
  SUNMemoryHelper h = c->mem_helper;
  if (!h) return;
  if (c->helper_device) h->ops->dealloc(h, c->helper_device, h->queue);
  if (c->helper_host) h->ops->dealloc(h, c->helper_host, h->queue);
  c->helper_device = c->helper_host = NULL;
  if (c->own_helper)
  {
    if (h->ops->destroy) h->ops->destroy(h);
    else
    {
      free(h->ops);
      free(h);
    }
  }
  c->mem_helper = NULL;
  c->own_helper = SUNFALSE;
}

static int alloc_data(N_Vector v)
{
  N_VectorContent_B200 c = NVC(v);
  size_t bytes           = (size_t)c->length * sizeof(sunrealtype);
  if (c->length == 0) return B200VEC_OK;
  if (c->mem_helper) return helper_alloc_data(c, bytes);
  int rc;
  void* p = NULL;
  switch (c->mem_kind)
  {
  case B200_MEM_DEVICE:
    rc = b200vec_malloc_device(c->ctx, bytes, &p);
    if (rc) return rc;
    c->device_data = (sunrealtype*)p;
    c->own_device  = SUNTRUE;
    break;
  case B200_MEM_MANAGED:
    rc = b200vec_malloc_managed(c->ctx, bytes, &p);
    if (rc) return rc;
    c->device_data = c->host_data = (sunrealtype*)p;
    c->own_device                 = SUNTRUE;
    break;
  case B200_MEM_PINNED:
    rc = b200vec_malloc_host(c->ctx, bytes, &p);
    if (rc) return rc;
    c->device_data = c->host_data = (sunrealtype*)p; /* UVA: same address on both sides */
    c->own_device                 = SUNTRUE;
    break;
  default: return B200VEC_ERR_ARG;
  }
  return B200VEC_OK;
}

static void free_data(N_Vector v)
{
  N_VectorContent_B200 c = NVC(v);
  if (c->helper_device || c->helper_host)
  { /* arrays from the user's allocator go back to it (the helper itself is released by N_VDestroy_B200) */
    SUNMemoryHelper h = c->mem_helper;
    if (c->helper_device) h->ops->dealloc(h, c->helper_device, h->queue);
    if (c->helper_host) h->ops->dealloc(h, c->helper_host, h->queue);
    c->helper_device = c->helper_host = NULL;
    c->device_data = c->host_data = NULL;
  }
  if (c->own_device && c->device_data)
  {
    size_t bytes = (size_t)c->length * sizeof(sunrealtype);
    if (c->mem_kind == B200_MEM_DEVICE) b200vec_free_device(c->ctx, c->device_data, bytes);
    else if (c->mem_kind == B200_MEM_MANAGED) b200vec_free_managed(c->ctx, c->device_data);
    else b200vec_free_host(c->ctx, c->device_data);
  }
  if (c->mem_kind == B200_MEM_DEVICE && c->own_host && c->host_data) b200vec_free_host(c->ctx, c->host_data);
  c->device_data = c->host_data = NULL;
  c->own_device = c->own_host = SUNFALSE;
}

/* ----------------------------------------------------------------------
 * constructors
 * -------------------------------------------------------------------- */
N_Vector N_VNewEmpty_B200(SUNContext sunctx)
{
  if (sunctx == NULL) return NULL; /* as N_VNewEmpty, sundials_nvector.c:48 */
  return new_shell(sunctx, NULL);
}

N_Vector N_VNewWithCtx_B200(sunindextype length, int mem_kind, b200vec_ctx ctx, SUNContext sunctx)
{
  if (length < 0) return NULL;
  if (mem_kind != B200_MEM_DEVICE && mem_kind != B200_MEM_MANAGED && mem_kind != B200_MEM_PINNED) return NULL;
  N_Vector v = new_shell(sunctx, ctx);
  if (!v) return NULL;
  NVC(v)->length = NVC(v)->global_length = length;
  NVC(v)->mem_kind                       = mem_kind;
  if (alloc_data(v) != B200VEC_OK)
  {
    fprintf(stderr, "[nvector_b200] allocation of %lld elements failed: %s\n", (long long)length,
            b200vec_last_error());
    N_VDestroy_B200(v);
    return NULL;
  }
  return v;
}

N_Vector N_VNewWithMemHelp_B200(sunindextype length, sunbooleantype use_managed_mem, SUNMemoryHelper helper,
                                SUNContext sunctx)
{
  if (sunctx == NULL || length < 0) return NULL;
  /* cuda:277-288: a helper with the required ops (SUNMemoryHelper_ImplementsRequiredOps, sundials_memory.c:266) */
  if (!helper || !helper->ops || !helper->ops->alloc || !helper->ops->dealloc || !helper->ops->copy) return NULL;
  N_Vector v = new_shell(sunctx, NULL);
  if (!v) return NULL;
  NVC(v)->length = NVC(v)->global_length = length;
  NVC(v)->mem_kind                       = use_managed_mem ? B200_MEM_MANAGED : B200_MEM_DEVICE;
  NVC(v)->mem_helper                     = helper; /* not owned (cuda:298) */
  if (alloc_data(v) != B200VEC_OK)
  {
    fprintf(stderr, "[nvector_b200] the memory helper could not allocate %lld elements\n", (long long)length);
    N_VDestroy_B200(v);
    return NULL;
  }
  return v;
}

N_Vector N_VNew_B200(sunindextype length, SUNContext sunctx)
{
  if (sunctx == NULL) return NULL;
  return N_VNewWithCtx_B200(length, B200_MEM_DEVICE, NULL, sunctx);
}

N_Vector N_VNewManaged_B200(sunindextype length, SUNContext sunctx)
{
  if (sunctx == NULL) return NULL;
  return N_VNewWithCtx_B200(length, B200_MEM_MANAGED, NULL, sunctx);
}

N_Vector N_VNewPinned_B200(sunindextype length, SUNContext sunctx)
{
  if (sunctx == NULL) return NULL;
  return N_VNewWithCtx_B200(length, B200_MEM_PINNED, NULL, sunctx);
}

N_Vector N_VMakeWithCtx_B200(sunindextype length, sunrealtype* h_vdata, sunrealtype* d_vdata, b200vec_ctx ctx,
                             SUNContext sunctx)
{
  if (length < 0) return NULL;
  if (length > 0 && d_vdata == NULL) return NULL; /* a device-accessible array is required */
  N_Vector v = new_shell(sunctx, ctx);
  if (!v) return NULL;
  NVC(v)->length = NVC(v)->global_length = length;
  NVC(v)->mem_kind    = (h_vdata != NULL && h_vdata == d_vdata) ? B200_MEM_MANAGED : B200_MEM_DEVICE;
  NVC(v)->host_data   = h_vdata;
  NVC(v)->device_data = d_vdata;
  return v;
}

N_Vector N_VMake_B200(sunindextype length, sunrealtype* h_vdata, sunrealtype* d_vdata, SUNContext sunctx)
{
  if (sunctx == NULL) return NULL;
  return N_VMakeWithCtx_B200(length, h_vdata, d_vdata, NULL, sunctx);
}

N_Vector N_VMakeManaged_B200(sunindextype length, sunrealtype* vdata, SUNContext sunctx)
{
  if (sunctx == NULL) return NULL;
  if (length > 0 && vdata == NULL) return NULL;
  /* h_vdata == d_vdata selects the host-coherent kind (cuda:387-437; here every op on it synchronises) */
  N_Vector v = N_VMakeWithCtx_B200(length, vdata, vdata, NULL, sunctx);
  if (v) NVC(v)->mem_kind = B200_MEM_MANAGED; /* also for length 0, where both pointers may be NULL */
  return v;
}

SUNErrCode N_VMakeDistributed_B200(N_Vector v, sunindextype global_length)
{
  if (!v || !v->content) return SUN_ERR_ARG_CORRUPT;
  if (global_length < 0)
  {
    int64_t g = (int64_t)NLEN(v);
    int rc    = b200vec_allreduce_i64_host(NCTX(v), &g, B200VEC_SUM);
    if (rc) return map_err(rc);
    global_length = (sunindextype)g;
  }
  NVC(v)->distributed   = SUNTRUE;
  NVC(v)->global_length = global_length;
  return SUN_SUCCESS;
}

/* ----------------------------------------------------------------------
 * accessors
 * -------------------------------------------------------------------- */
N_Vector_ID N_VGetVectorID_B200(N_Vector v)
{
  (void)v;
  return SUNDIALS_NVEC_CUSTOM;
}

/* NVECTOR_B200 answering as the vector it replaces: the one place the reference keys on the CUDA ID is the
   gate of CVodeSetUseIntegratorFusedKernels (src/cvode/cvode_io.c:1022-1029).  The ID lives in the ops
   table, which clones copy (N_VCloneEmpty_B200), so the template vector's choice propagates. */
static N_Vector_ID getvectorid_as_cuda(N_Vector v)
{
  (void)v;
  return SUNDIALS_NVEC_CUDA;
}

SUNErrCode N_VSetVectorID_B200(N_Vector v, N_Vector_ID id)
{
  if (!v || !v->ops) return SUN_ERR_ARG_CORRUPT;
  if (id == SUNDIALS_NVEC_CUSTOM) v->ops->nvgetvectorid = N_VGetVectorID_B200;
  else if (id == SUNDIALS_NVEC_CUDA) v->ops->nvgetvectorid = getvectorid_as_cuda;
  else return SUN_ERR_ARG_OUTOFRANGE;
  return SUN_SUCCESS;
}

sunindextype N_VGetLength_B200(N_Vector v) { return NVC(v)->global_length; }
sunindextype N_VGetLocalLength_B200(N_Vector v) { return NVC(v)->length; }
b200vec_ctx N_VGetCtx_B200(N_Vector v) { return NCTX(v); }
sunrealtype* N_VGetDeviceArrayPointer_B200(N_Vector v) { return NDEV(v); }
sunbooleantype N_VIsManagedMemory_B200(N_Vector v) { return NVC(v)->mem_kind == B200_MEM_MANAGED; }

sunrealtype* N_VGetHostArrayPointer_B200(N_Vector v)
{
  N_VectorContent_B200 c = NVC(v);
  if (c->mem_kind == B200_MEM_DEVICE)
  {
    if (!c->host_data && c->device_data && c->length > 0)
    { /* lazily created pinned mirror (the reference allocates it eagerly, cuda:2226); an empty clone
         (no device array yet) has no host side either, as Test_N_VCloneEmpty's has_data expects */
      void* p = NULL;
      CHECK_ON(v, b200vec_malloc_host(c->ctx, (size_t)c->length * sizeof(sunrealtype), &p));
      c->host_data = (sunrealtype*)p;
      c->own_host  = SUNTRUE;
    }
    return c->host_data;
  }
  /* host-coherent kinds: make pending kernels visible before the host looks */
  CHECK_ON(v, b200vec_ctx_sync(c->ctx));
  return c->host_data;
}

sunrealtype* N_VGetArrayPointer_B200(N_Vector v) { return N_VGetHostArrayPointer_B200(v); }

void N_VSetHostArrayPointer_B200(sunrealtype* h_vdata, N_Vector v)
{
  N_VectorContent_B200 c = NVC(v);
  if (c->mem_kind == B200_MEM_DEVICE)
  {
    if (c->own_host && c->host_data) b200vec_free_host(c->ctx, c->host_data);
    c->host_data = h_vdata;
    c->own_host  = SUNFALSE;
  }
  else
  { /* single array visible from both sides */
    if (c->own_device && c->device_data) free_data(v);
    c->host_data = c->device_data = h_vdata;
    c->own_device                 = SUNFALSE;
  }
}

void N_VSetDeviceArrayPointer_B200(sunrealtype* d_vdata, N_Vector v)
{
  N_VectorContent_B200 c = NVC(v);
  if (c->mem_kind == B200_MEM_DEVICE)
  {
    if (c->own_device && c->device_data)
      b200vec_free_device(c->ctx, c->device_data, (size_t)c->length * sizeof(sunrealtype));
    c->device_data = d_vdata;
    c->own_device  = SUNFALSE;
  }
  else N_VSetHostArrayPointer_B200(d_vdata, v);
}

void N_VSetArrayPointer_B200(sunrealtype* v_data, N_Vector v)
{
  if (NLEN(v) > 0) N_VSetHostArrayPointer_B200(v_data, v); /* serial:380-385, cuda:151 */
}

void N_VCopyToDevice_B200(N_Vector v)
{
  N_VectorContent_B200 c = NVC(v);
  if (c->mem_kind != B200_MEM_DEVICE || !c->host_data) return;
  CHECK_ON(v, b200vec_copy_h2d(c->ctx, c->device_data, c->host_data, (size_t)c->length * sizeof(sunrealtype), 1));
}

void N_VCopyFromDevice_B200(N_Vector v)
{
  N_VectorContent_B200 c = NVC(v);
  if (c->mem_kind != B200_MEM_DEVICE)
  {
    CHECK_ON(v, b200vec_ctx_sync(c->ctx));
    return;
  }
  (void)N_VGetHostArrayPointer_B200(v);
  CHECK_ON(v, b200vec_copy_d2h(c->ctx, c->host_data, c->device_data, (size_t)c->length * sizeof(sunrealtype), 1));
}

/* asynchronous upload on the context's copy stream: ordered after the vector work issued so far,
   overlapping whatever is issued next; N_VCopyJoin_B200 makes the context's stream wait for it
   (no host blocking).  The reference has only the synchronous form (cuda:553). */
void N_VCopyToDeviceAsync_B200(N_Vector v)
{
  N_VectorContent_B200 c = NVC(v);
  if (c->mem_kind != B200_MEM_DEVICE || !c->host_data) return;
  CHECK_ON(v, b200vec_copy_h2d_async(c->ctx, c->device_data, c->host_data, (size_t)c->length * sizeof(sunrealtype)));
}

void N_VCopyJoin_B200(N_Vector v) { CHECK_ON(v, b200vec_copy_join(NCTX(v))); }

SUNErrCode N_VSetStream_B200(N_Vector v, void* stream) { return map_err(b200vec_ctx_set_stream(NCTX(v), stream)); }

void N_VSpace_B200(N_Vector v, sunindextype* lrw, sunindextype* liw)
{
  *lrw = NVC(v)->global_length; /* serial:362-373 */
  *liw = 1;
}

void N_VPrintFile_B200(N_Vector v, FILE* outfile)
{
  N_VCopyFromDevice_B200(v);
  sunrealtype* h = N_VGetHostArrayPointer_B200(v);
  for (sunindextype i = 0; i < NLEN(v); i++) fprintf(outfile, "%35.32e\n", h[i]);
}

void N_VPrint_B200(N_Vector v) { N_VPrintFile_B200(v, stdout); }

/* ----------------------------------------------------------------------
 * clone / destroy  (serial:275-360, cuda:603-748)
 * -------------------------------------------------------------------- */
N_Vector N_VCloneEmpty_B200(N_Vector w)
{
  if (!w || !w->content) return NULL;
  N_Vector v = generic_new(w->sunctx);
  if (!v) return NULL;
  *(v->ops) = *(w->ops); /* N_VCopyOps: enabled/disabled fused ops propagate */
  N_VectorContent_B200 c = (N_VectorContent_B200)calloc(1, sizeof *c);
  if (!c)
  {
    free(v->ops);
    free(v);
    return NULL;
  }
  *c             = *NVC(w);
  c->host_data   = NULL;
  c->device_data = NULL;
  c->own_device = c->own_host = SUNFALSE;
  c->helper_host = c->helper_device = NULL;
  c->own_helper                     = SUNFALSE;
  if (c->mem_helper)
  { /* cuda:662-663: every clone owns a clone of the helper */
    c->mem_helper = helper_clone(c->mem_helper);
    if (!c->mem_helper)
    {
      free(c);
      free(v->ops);
      free(v);
      return NULL;
    }
    c->own_helper = SUNTRUE;
  }
  b200vec_ctx_retain(c->ctx);
  v->content = c;
  return v;
}

N_Vector N_VClone_B200(N_Vector w)
{
  N_Vector v = N_VCloneEmpty_B200(w);
  if (!v) return NULL;
  if (alloc_data(v) != B200VEC_OK)
  {
    fprintf(stderr, "[nvector_b200] clone of %lld elements failed: %s\n", (long long)NLEN(w), b200vec_last_error());
    N_VDestroy_B200(v);
    return NULL;
  }
  return v;
}

void N_VDestroy_B200(N_Vector v)
{
  if (v == NULL) return;
  if (v->content != NULL)
  {
    free_data(v);
    helper_release(NVC(v));
    b200vec_ctx_release(NCTX(v));
    free(v->content);
    v->content = NULL;
  }
  if (v->ops != NULL)
  {
    free(v->ops);
    v->ops = NULL;
  }
  free(v);
}

/* ----------------------------------------------------------------------
 * streaming ops
 * -------------------------------------------------------------------- */
void N_VLinearSum_B200(sunrealtype a, N_Vector x, sunrealtype b, N_Vector y, N_Vector z)
{
  CHECK_ON(z, b200vec_linear_sum(NCTX(z), a, NDEV(x), b, NDEV(y), NDEV(z), NLEN(z)));
  coherent_sync(z);
}

void N_VConst_B200(sunrealtype c, N_Vector z)
{
  CHECK_ON(z, b200vec_const(NCTX(z), c, NDEV(z), NLEN(z)));
  coherent_sync(z);
}

void N_VProd_B200(N_Vector x, N_Vector y, N_Vector z)
{
  CHECK_ON(z, b200vec_prod(NCTX(z), NDEV(x), NDEV(y), NDEV(z), NLEN(z)));
  coherent_sync(z);
}

void N_VDiv_B200(N_Vector x, N_Vector y, N_Vector z)
{
  CHECK_ON(z, b200vec_div(NCTX(z), NDEV(x), NDEV(y), NDEV(z), NLEN(z)));
  coherent_sync(z);
}

void N_VScale_B200(sunrealtype c, N_Vector x, N_Vector z)
{
  CHECK_ON(z, b200vec_scale(NCTX(z), c, NDEV(x), NDEV(z), NLEN(z)));
  coherent_sync(z);
}

void N_VAbs_B200(N_Vector x, N_Vector z)
{
  CHECK_ON(z, b200vec_abs(NCTX(z), NDEV(x), NDEV(z), NLEN(z)));
  coherent_sync(z);
}

void N_VInv_B200(N_Vector x, N_Vector z)
{
  CHECK_ON(z, b200vec_inv(NCTX(z), NDEV(x), NDEV(z), NLEN(z)));
  coherent_sync(z);
}

void N_VAddConst_B200(N_Vector x, sunrealtype b, N_Vector z)
{
  CHECK_ON(z, b200vec_add_const(NCTX(z), NDEV(x), b, NDEV(z), NLEN(z)));
  coherent_sync(z);
}

void N_VCompare_B200(sunrealtype c, N_Vector x, N_Vector z)
{
  CHECK_ON(z, b200vec_compare(NCTX(z), c, NDEV(x), NDEV(z), NLEN(z)));
  coherent_sync(z);
}

/* ----------------------------------------------------------------------
 * reductions.  Local forms return the rank-local value (the kernel stores the
 * scalar in pinned host memory; no memcpy, no stream sync).  Global forms on a
 * distributed vector set the one-shot GLOBAL scope: the same kernel then folds
 * the ranks' partials over NVLink peer memory in its last CTA (or the library
 * falls back to ncclAllReduce) -- nvector_manyvector.c:754-1400 semantics,
 * SUM / MAX / MIN implied by the op.
 * -------------------------------------------------------------------- */
#define GLOBAL_SCOPE(v) \
  do { if (NDIST(v)) CHECK_ON(v, b200vec_ctx_set_scope(NCTX(v), B200VEC_SCOPE_GLOBAL)); } while (0)

sunrealtype N_VDotProdLocal_B200(N_Vector x, N_Vector y)
{
  double r = NAN;
  CHECK_ON(x, b200vec_dot_prod(NCTX(x), NDEV(x), NDEV(y), NLEN(x), &r));
  return r;
}

sunrealtype N_VDotProd_B200(N_Vector x, N_Vector y)
{
  double r = NAN;
  GLOBAL_SCOPE(x); /* nvector_manyvector.c:815 */
  CHECK_ON(x, b200vec_dot_prod(NCTX(x), NDEV(x), NDEV(y), NLEN(x), &r));
  return r;
}

sunrealtype N_VMaxNormLocal_B200(N_Vector x)
{
  double r = NAN;
  CHECK_ON(x, b200vec_max_norm(NCTX(x), NDEV(x), NLEN(x), &r));
  return r;
}

sunrealtype N_VMaxNorm_B200(N_Vector x)
{
  double r = NAN;
  GLOBAL_SCOPE(x); /* :869 */
  CHECK_ON(x, b200vec_max_norm(NCTX(x), NDEV(x), NLEN(x), &r));
  return r;
}

sunrealtype N_VMinLocal_B200(N_Vector x)
{
  double r = NAN;
  CHECK_ON(x, b200vec_min(NCTX(x), NDEV(x), NLEN(x), &r));
  return r;
}

sunrealtype N_VMin_B200(N_Vector x)
{
  double r = NAN;
  GLOBAL_SCOPE(x); /* :1107 */
  CHECK_ON(x, b200vec_min(NCTX(x), NDEV(x), NLEN(x), &r));
  return r;
}

sunrealtype N_VL1NormLocal_B200(N_Vector x)
{
  double r = NAN;
  CHECK_ON(x, b200vec_l1_norm(NCTX(x), NDEV(x), NLEN(x), &r));
  return r;
}

sunrealtype N_VL1Norm_B200(N_Vector x)
{
  double r = NAN;
  GLOBAL_SCOPE(x); /* :1203 */
  CHECK_ON(x, b200vec_l1_norm(NCTX(x), NDEV(x), NLEN(x), &r));
  return r;
}

sunrealtype N_VWSqrSumLocal_B200(N_Vector x, N_Vector w)
{
  double r = NAN;
  CHECK_ON(x, b200vec_wsqr_sum(NCTX(x), NDEV(x), NDEV(w), NLEN(x), &r));
  return r;
}

sunrealtype N_VWSqrSumMaskLocal_B200(N_Vector x, N_Vector w, N_Vector id)
{
  double r = NAN;
  CHECK_ON(x, b200vec_wsqr_sum_mask(NCTX(x), NDEV(x), NDEV(w), NDEV(id), NLEN(x), &r));
  return r;
}

static sunrealtype wsqr_global(N_Vector x, N_Vector w, N_Vector id)
{
  double r = NAN;
  GLOBAL_SCOPE(x); /* :956, :1050, :1128 */
  if (id) CHECK_ON(x, b200vec_wsqr_sum_mask(NCTX(x), NDEV(x), NDEV(w), NDEV(id), NLEN(x), &r));
  else CHECK_ON(x, b200vec_wsqr_sum(NCTX(x), NDEV(x), NDEV(w), NLEN(x), &r));
  return r;
}

/* serial:641-648 -- divides by the GLOBAL length (nvector_manyvector.c:963) */
sunrealtype N_VWrmsNorm_B200(N_Vector x, N_Vector w)
{
  return rsqrt_guard(wsqr_global(x, w, NULL) / (sunrealtype)NVC(x)->global_length);
}

sunrealtype N_VWrmsNormMask_B200(N_Vector x, N_Vector w, N_Vector id)
{
  return rsqrt_guard(wsqr_global(x, w, id) / (sunrealtype)NVC(x)->global_length);
}

sunrealtype N_VWL2Norm_B200(N_Vector x, N_Vector w) { return rsqrt_guard(wsqr_global(x, w, NULL)); }

sunbooleantype N_VInvTestLocal_B200(N_Vector x, N_Vector z)
{
  double r = NAN;
  CHECK_ON(z, b200vec_inv_test(NCTX(z), NDEV(x), NDEV(z), NLEN(z), &r));
  coherent_sync(z); /* the scalar arrives before the kernel has drained: host-coherent z needs the sync */
  return (r > 0.5) ? SUNTRUE : SUNFALSE;
}

sunbooleantype N_VInvTest_B200(N_Vector x, N_Vector z)
{
  double r = NAN;
  GLOBAL_SCOPE(z); /* :1277 -- on the context that launches */
  CHECK_ON(z, b200vec_inv_test(NCTX(z), NDEV(x), NDEV(z), NLEN(z), &r));
  coherent_sync(z);
  return (r > 0.5) ? SUNTRUE : SUNFALSE;
}

sunbooleantype N_VConstrMaskLocal_B200(N_Vector c, N_Vector x, N_Vector m)
{
  double r = NAN;
  CHECK_ON(m, b200vec_constr_mask(NCTX(m), NDEV(c), NDEV(x), NDEV(m), NLEN(m), &r));
  coherent_sync(m);
  return (r > 0.5) ? SUNTRUE : SUNFALSE;
}

sunbooleantype N_VConstrMask_B200(N_Vector c, N_Vector x, N_Vector m)
{
  double r = NAN;
  GLOBAL_SCOPE(m); /* :1339 -- on the context that launches */
  CHECK_ON(m, b200vec_constr_mask(NCTX(m), NDEV(c), NDEV(x), NDEV(m), NLEN(m), &r));
  coherent_sync(m);
  return (r > 0.5) ? SUNTRUE : SUNFALSE;
}

sunrealtype N_VMinQuotientLocal_B200(N_Vector num, N_Vector denom)
{
  double r = NAN;
  CHECK_ON(num, b200vec_min_quotient(NCTX(num), NDEV(num), NDEV(denom), NLEN(num), &r));
  return r;
}

sunrealtype N_VMinQuotient_B200(N_Vector num, N_Vector denom)
{
  double r = NAN;
  GLOBAL_SCOPE(num); /* :1399 */
  CHECK_ON(num, b200vec_min_quotient(NCTX(num), NDEV(num), NDEV(denom), NLEN(num), &r));
  return r;
}

/* ----------------------------------------------------------------------
 * fused ops: gather device pointers of the handles into a small host table
 * (stack for the usual nvec, heap beyond) and make ONE C-ABI call.
 * -------------------------------------------------------------------- */
typedef struct
{
  const double** p;
  const double* stack[MAX_STACK_VECS];
} ptr_table;

static int table_init(ptr_table* t, int count)
{
  t->p = (count <= MAX_STACK_VECS) ? t->stack : (const double**)malloc((size_t)count * sizeof(double*));
  return t->p ? 0 : -1;
}

static void table_free(ptr_table* t)
{
  if (t->p != t->stack) free((void*)t->p);
}

static int gather(ptr_table* t, N_Vector* V, int count)
{
  if (table_init(t, count)) return -1;
  for (int i = 0; i < count; i++) t->p[i] = NDEV(V[i]);
  return 0;
}

/* V2[j][i] -> t[j*nvec + i] */
static int gather2d(ptr_table* t, N_Vector** V2, int nsum, int nvec)
{
  if (table_init(t, nsum * nvec)) return -1;
  for (int j = 0; j < nsum; j++)
    for (int i = 0; i < nvec; i++) t->p[j * nvec + i] = NDEV(V2[j][i]);
  return 0;
}

SUNErrCode N_VLinearCombination_B200(int nvec, sunrealtype* c, N_Vector* X, N_Vector z)
{
  if (nvec < 1) return SUN_ERR_ARG_OUTOFRANGE;
  ptr_table tx;
  if (gather(&tx, X, nvec)) return SUN_ERR_MALLOC_FAIL;
  /* nvec <= 2 delegate to Scale / LinearSum whose in-place forms are chosen on
     HANDLE identity (serial:885-898); data-pointer identity is equivalent here */
  int rc = b200vec_linear_combination(NCTX(z), nvec, c, tx.p, NDEV(z), NLEN(z));
  table_free(&tx);
  if (!rc) coherent_sync(z);
  return map_err(rc);
}

/* z = sum c_i X_i and *sqnorm = z . z (GLOBAL on a distributed vector) in one kernel and one host
   round trip: N_VLinearCombination + N_VDotProd(z, z) of sundials_iterative.c:137-152 */
SUNErrCode N_VLinearCombinationSqNorm_B200(int nvec, sunrealtype* c, N_Vector* X, N_Vector z, sunrealtype* sqnorm)
{
  if (nvec < 1 || !sqnorm) return SUN_ERR_ARG_OUTOFRANGE;
  ptr_table tx;
  if (gather(&tx, X, nvec)) return SUN_ERR_MALLOC_FAIL;
  int rc = NDIST(z) ? b200vec_ctx_set_scope(NCTX(z), B200VEC_SCOPE_GLOBAL) : B200VEC_OK;
  if (!rc) rc = b200vec_linear_combination_sqnorm(NCTX(z), nvec, c, tx.p, NDEV(z), NLEN(z), sqnorm);
  table_free(&tx);
  if (!rc) coherent_sync(z);
  return map_err(rc);
}

/* ewt = 1 / (rtol |y| + atol) in ONE kernel -- the body of a user error-weight function for
   CVodeWFtolerances / ARKodeWFtolerances / IDAWFtolerances (the public hook of the UNMODIFIED integrators),
   bit-identical to their built-in cvEwtSetSS/SV, arkEwtSetSS/SV sequences (5 vector ops).  vatol == NULL:
   scalar atol.  Returns 0, or -1 when atolmin0 (some absolute tolerance is zero) and a denominator is <= 0,
   exactly as the built-in routines do (src/cvode/cvode.c:4794-4860). */
int N_VEwtSet_B200(sunrealtype rtol, sunrealtype atol, N_Vector vatol, sunbooleantype atolmin0, N_Vector y, N_Vector ewt)
{
  double mind = NAN;
  GLOBAL_SCOPE(ewt);
  CHECK_ON(ewt, b200vec_ewt_set(NCTX(ewt), rtol, atol, vatol ? NDEV(vatol) : NULL, NDEV(y), NDEV(ewt), NLEN(ewt), &mind));
  coherent_sync(ewt);
  if (mind != mind) return -1; /* the kernel did not run */
  return (atolmin0 && mind <= 0.0) ? -1 : 0;
}

/* z <- a x + z and *dot = w . z(updated) (GLOBAL on a distributed vector): N_VLinearSum(1, z, a, x, z)
   + N_VDotProd(w, z) of a modified Gram-Schmidt sweep (sundials_iterative.c:62-67) in one kernel */
SUNErrCode N_VAxpyDot_B200(sunrealtype a, N_Vector x, N_Vector z, N_Vector w, sunrealtype* dot)
{
  if (!dot) return SUN_ERR_ARG_OUTOFRANGE;
  int rc = NDIST(z) ? b200vec_ctx_set_scope(NCTX(z), B200VEC_SCOPE_GLOBAL) : B200VEC_OK;
  if (!rc) rc = b200vec_axpy_dot(NCTX(z), a, NDEV(x), NDEV(z), NDEV(w), NLEN(z), dot);
  if (!rc) coherent_sync(z);
  return map_err(rc);
}

/* one modified Gram-Schmidt column as a chain of nproj + 1 kernels with ONE host wait (b200vec_mgs_sweep) */
SUNErrCode N_VModifiedGSSweep_B200(int nproj, N_Vector* V, N_Vector vk, sunrealtype* h, sunrealtype* norms)
{
  if (nproj < 1 || !h || !norms) return SUN_ERR_ARG_OUTOFRANGE;
  ptr_table tv;
  if (gather(&tv, V, nproj)) return SUN_ERR_MALLOC_FAIL;
  int rc = NDIST(vk) ? b200vec_ctx_set_scope(NCTX(vk), B200VEC_SCOPE_GLOBAL) : B200VEC_OK;
  if (!rc) rc = b200vec_mgs_sweep(NCTX(vk), nproj, NDEV(vk), tv.p, NLEN(vk), h, norms);
  table_free(&tv);
  if (!rc) coherent_sync(vk);
  return map_err(rc);
}

/* one classical Gram-Schmidt column as a chain of 2 kernels with ONE host wait (b200vec_cgs_step) */
SUNErrCode N_VClassicalGSStep_B200(int nvec, N_Vector x, N_Vector* Ydots, N_Vector* Xcomb, N_Vector z,
                                   sunrealtype* dots, sunrealtype* sqnorm)
{
  if (nvec < 2 || !dots || !sqnorm) return SUN_ERR_ARG_OUTOFRANGE;
  ptr_table ty, tx;
  if (gather(&ty, Ydots, nvec)) return SUN_ERR_MALLOC_FAIL;
  if (gather(&tx, Xcomb, nvec))
  {
    table_free(&ty);
    return SUN_ERR_MALLOC_FAIL;
  }
  int rc = NDIST(z) ? b200vec_ctx_set_scope(NCTX(z), B200VEC_SCOPE_GLOBAL) : B200VEC_OK;
  if (!rc) rc = b200vec_cgs_step(NCTX(z), nvec, NDEV(x), ty.p, tx.p, NDEV(z), NLEN(z), dots, sqnorm);
  table_free(&ty);
  table_free(&tx);
  if (!rc) coherent_sync(z);
  return map_err(rc);
}

SUNErrCode N_VScaleAddMulti_B200(int nvec, sunrealtype* a, N_Vector x, N_Vector* Y, N_Vector* Z)
{
  if (nvec < 1) return SUN_ERR_ARG_OUTOFRANGE;
  ptr_table ty, tz;
  if (gather(&ty, Y, nvec)) return SUN_ERR_MALLOC_FAIL;
  if (gather(&tz, Z, nvec))
  {
    table_free(&ty);
    return SUN_ERR_MALLOC_FAIL;
  }
  int rc = b200vec_scale_add_multi(NCTX(x), nvec, a, NDEV(x), ty.p, (double* const*)tz.p, NLEN(x));
  table_free(&ty);
  table_free(&tz);
  if (!rc) coherent_sync(x);
  return map_err(rc);
}

SUNErrCode N_VDotProdMultiLocal_B200(int nvec, N_Vector x, N_Vector* Y, sunrealtype* dotprods)
{
  if (nvec < 1) return SUN_ERR_ARG_OUTOFRANGE;
  ptr_table ty;
  if (gather(&ty, Y, nvec)) return SUN_ERR_MALLOC_FAIL;
  int rc = B200VEC_OK;
  /* the context has a fixed number of result slots: long lists go in slices */
  for (int j0 = 0; j0 < nvec && !rc; j0 += MAX_STACK_VECS)
  {
    int nj = (nvec - j0 < MAX_STACK_VECS) ? nvec - j0 : MAX_STACK_VECS;
    rc     = b200vec_dot_prod_multi(NCTX(x), nj, NDEV(x), ty.p + j0, NLEN(x), dotprods + j0);
  }
  table_free(&ty);
  return map_err(rc);
}

SUNErrCode N_VDotProdMulti_B200(int nvec, N_Vector x, N_Vector* Y, sunrealtype* dotprods)
{
  if (nvec < 1) return SUN_ERR_ARG_OUTOFRANGE;
  if (!NDIST(x)) return N_VDotProdMultiLocal_B200(nvec, x, Y, dotprods);
  ptr_table ty;
  if (gather(&ty, Y, nvec)) return SUN_ERR_MALLOC_FAIL;
  int rc = B200VEC_OK;
  for (int j0 = 0; j0 < nvec && !rc; j0 += MAX_STACK_VECS)
  {
    int nj = (nvec - j0 < MAX_STACK_VECS) ? nvec - j0 : MAX_STACK_VECS;
    /* fused local dots and ONE nj-wide exchange inside the same kernel
       (nvector_manyvector.c:1576; the reference's local part is nj separate
       kernels + syncs, :1566-1570) */
    rc = b200vec_ctx_set_scope(NCTX(x), B200VEC_SCOPE_GLOBAL);
    if (!rc) rc = b200vec_dot_prod_multi(NCTX(x), nj, NDEV(x), ty.p + j0, NLEN(x), dotprods + j0);
  }
  table_free(&ty);
  return map_err(rc);
}

/* nvector_manyvector.c:1452-1462: in-place SUM allreduce of a host array */
SUNErrCode N_VDotProdMultiAllReduce_B200(int nvec_total, N_Vector x, sunrealtype* sum)
{
  if (nvec_total < 1) return SUN_ERR_ARG_OUTOFRANGE;
  if (!NDIST(x)) return SUN_SUCCESS;
  b200vec_ctx ctx = NCTX(x);
  int rc          = B200VEC_OK;
  for (int j0 = 0; j0 < nvec_total && !rc; j0 += MAX_STACK_VECS)
  {
    int nj = (nvec_total - j0 < MAX_STACK_VECS) ? nvec_total - j0 : MAX_STACK_VECS;
    rc     = b200vec_copy_h2d(ctx, b200vec_result_device(ctx), sum + j0, (size_t)nj * sizeof(double), 0);
    if (!rc) rc = b200vec_allreduce(ctx, nj, B200VEC_SUM);
    if (!rc) rc = b200vec_result_fetch(ctx, nj, sum + j0);
  }
  return map_err(rc);
}

/* ----------------------------------------------------------------------
 * vector-array ops
 * -------------------------------------------------------------------- */
SUNErrCode N_VLinearSumVectorArray_B200(int nvec, sunrealtype a, N_Vector* X, sunrealtype b, N_Vector* Y,
                                        N_Vector* Z)
{
  if (nvec < 1) return SUN_ERR_ARG_OUTOFRANGE;
  ptr_table tx, ty, tz;
  if (gather(&tx, X, nvec)) return SUN_ERR_MALLOC_FAIL;
  if (gather(&ty, Y, nvec))
  {
    table_free(&tx);
    return SUN_ERR_MALLOC_FAIL;
  }
  if (gather(&tz, Z, nvec))
  {
    table_free(&tx);
    table_free(&ty);
    return SUN_ERR_MALLOC_FAIL;
  }
  int rc = b200vec_linear_sum_vector_array(NCTX(Z[0]), nvec, a, tx.p, b, ty.p, (double* const*)tz.p, Z == X, Z == Y,
                                           NLEN(Z[0]));
  table_free(&tx);
  table_free(&ty);
  table_free(&tz);
  if (!rc) coherent_sync(Z[0]);
  return map_err(rc);
}

SUNErrCode N_VScaleVectorArray_B200(int nvec, sunrealtype* c, N_Vector* X, N_Vector* Z)
{
  if (nvec < 1) return SUN_ERR_ARG_OUTOFRANGE;
  ptr_table tx, tz;
  if (gather(&tx, X, nvec)) return SUN_ERR_MALLOC_FAIL;
  if (gather(&tz, Z, nvec))
  {
    table_free(&tx);
    return SUN_ERR_MALLOC_FAIL;
  }
  int rc = b200vec_scale_vector_array(NCTX(Z[0]), nvec, c, tx.p, (double* const*)tz.p, NLEN(Z[0]));
  table_free(&tx);
  table_free(&tz);
  if (!rc) coherent_sync(Z[0]);
  return map_err(rc);
}

SUNErrCode N_VConstVectorArray_B200(int nvec, sunrealtype c, N_Vector* Z)
{
  if (nvec < 1) return SUN_ERR_ARG_OUTOFRANGE;
  ptr_table tz;
  if (gather(&tz, Z, nvec)) return SUN_ERR_MALLOC_FAIL;
  int rc = b200vec_const_vector_array(NCTX(Z[0]), nvec, c, (double* const*)tz.p, NLEN(Z[0]));
  table_free(&tz);
  if (!rc) coherent_sync(Z[0]);
  return map_err(rc);
}

static SUNErrCode wrms_va(int nvec, N_Vector* X, N_Vector* W, N_Vector id, sunrealtype* nrm)
{
  if (nvec < 1) return SUN_ERR_ARG_OUTOFRANGE;
  ptr_table tx, tw;
  if (gather(&tx, X, nvec)) return SUN_ERR_MALLOC_FAIL;
  if (gather(&tw, W, nvec))
  {
    table_free(&tx);
    return SUN_ERR_MALLOC_FAIL;
  }
  N_Vector x0         = X[0];
  b200vec_ctx ctx     = NCTX(x0);
  const int dist      = NDIST(x0);
  const double* idd   = id ? NDEV(id) : NULL;
  const sunrealtype N = (sunrealtype)NVC(x0)->global_length;
  int rc              = B200VEC_OK;
  for (int j0 = 0; j0 < nvec && !rc; j0 += MAX_STACK_VECS)
  {
    int nj = (nvec - j0 < MAX_STACK_VECS) ? nvec - j0 : MAX_STACK_VECS;
    if (dist) rc = b200vec_ctx_set_scope(ctx, B200VEC_SCOPE_GLOBAL); /* nvector_manyvector.c:1749,1793 */
    if (!rc) rc = b200vec_wsqr_sum_vector_array(ctx, nj, tx.p + j0, tw.p + j0, idd, NLEN(x0), nrm + j0);
  }
  table_free(&tx);
  table_free(&tw);
  if (rc) return map_err(rc);
  for (int i = 0; i < nvec; i++) nrm[i] = rsqrt_guard(nrm[i] / N); /* serial:1262,1303 */
  return SUN_SUCCESS;
}

SUNErrCode N_VWrmsNormVectorArray_B200(int nvec, N_Vector* X, N_Vector* W, sunrealtype* nrm)
{
  return wrms_va(nvec, X, W, NULL, nrm);
}

SUNErrCode N_VWrmsNormMaskVectorArray_B200(int nvec, N_Vector* X, N_Vector* W, N_Vector id, sunrealtype* nrm)
{
  return wrms_va(nvec, X, W, id, nrm);
}

SUNErrCode N_VScaleAddMultiVectorArray_B200(int nvec, int nsum, sunrealtype* a, N_Vector* X, N_Vector** Y,
                                            N_Vector** Z)
{
  if (nvec < 1 || nsum < 1) return SUN_ERR_ARG_OUTOFRANGE;
  ptr_table tx, ty, tz;
  if (gather(&tx, X, nvec)) return SUN_ERR_MALLOC_FAIL;
  if (gather2d(&ty, Y, nsum, nvec))
  {
    table_free(&tx);
    return SUN_ERR_MALLOC_FAIL;
  }
  if (gather2d(&tz, Z, nsum, nvec))
  {
    table_free(&tx);
    table_free(&ty);
    return SUN_ERR_MALLOC_FAIL;
  }
  /* nsum == 1 delegates to LinearSumVectorArray(…, Y[0], Z[0]) whose axpy form
     keys on Y[0] == Z[0] as arrays (serial:1366, 1062) */
  int y_is_z = (Y == Z) || (Y[0] == Z[0]);
  int rc = b200vec_scale_add_multi_vector_array(NCTX(X[0]), nvec, nsum, a, tx.p, ty.p, (double* const*)tz.p, y_is_z,
                                                NLEN(X[0]));
  table_free(&tx);
  table_free(&ty);
  table_free(&tz);
  if (!rc) coherent_sync(X[0]);
  return map_err(rc);
}

SUNErrCode N_VLinearCombinationVectorArray_B200(int nvec, int nsum, sunrealtype* c, N_Vector** X, N_Vector* Z)
{
  if (nvec < 1 || nsum < 1) return SUN_ERR_ARG_OUTOFRANGE;
  ptr_table tx, tz;
  if (gather2d(&tx, X, nsum, nvec)) return SUN_ERR_MALLOC_FAIL;
  if (gather(&tz, Z, nvec))
  {
    table_free(&tx);
    return SUN_ERR_MALLOC_FAIL;
  }
  int rc = b200vec_linear_combination_vector_array(NCTX(Z[0]), nvec, nsum, c, tx.p, (double* const*)tz.p, X[0] == Z,
                                                   NLEN(Z[0]));
  table_free(&tx);
  table_free(&tz);
  if (!rc) coherent_sync(Z[0]);
  return map_err(rc);
}

/* ----------------------------------------------------------------------
 * XBraid buffer ops
 * -------------------------------------------------------------------- */
SUNErrCode N_VBufSize_B200(N_Vector x, sunindextype* size)
{
  if (!size) return SUN_ERR_ARG_CORRUPT;
  *size = NLEN(x) * (sunindextype)sizeof(sunrealtype);
  return SUN_SUCCESS;
}

SUNErrCode N_VBufPack_B200(N_Vector x, void* buf)
{
  if (!buf) return SUN_ERR_ARG_CORRUPT;
  return map_err(b200vec_copy_d2h(NCTX(x), buf, NDEV(x), (size_t)NLEN(x) * sizeof(sunrealtype), 1));
}

SUNErrCode N_VBufUnpack_B200(N_Vector x, void* buf)
{
  if (!buf) return SUN_ERR_ARG_CORRUPT;
  return map_err(b200vec_copy_h2d(NCTX(x), NDEV(x), buf, (size_t)NLEN(x) * sizeof(sunrealtype), 1));
}

/* ----------------------------------------------------------------------
 * enable / disable fused and vector-array ops  (serial:1948-2053)
 * -------------------------------------------------------------------- */
SUNErrCode N_VEnableFusedOps_B200(N_Vector v, sunbooleantype tf)
{
  if (!v || !v->ops) return SUN_ERR_ARG_CORRUPT;
  N_VEnableLinearCombination_B200(v, tf);
  N_VEnableScaleAddMulti_B200(v, tf);
  N_VEnableDotProdMulti_B200(v, tf);
  N_VEnableLinearSumVectorArray_B200(v, tf);
  N_VEnableScaleVectorArray_B200(v, tf);
  N_VEnableConstVectorArray_B200(v, tf);
  N_VEnableWrmsNormVectorArray_B200(v, tf);
  N_VEnableWrmsNormMaskVectorArray_B200(v, tf);
  N_VEnableScaleAddMultiVectorArray_B200(v, tf);
  N_VEnableLinearCombinationVectorArray_B200(v, tf);
  return SUN_SUCCESS;
}

SUNErrCode N_VEnableLinearCombination_B200(N_Vector v, sunbooleantype tf)
{
  v->ops->nvlinearcombination = tf ? N_VLinearCombination_B200 : NULL;
  return SUN_SUCCESS;
}

SUNErrCode N_VEnableScaleAddMulti_B200(N_Vector v, sunbooleantype tf)
{
  v->ops->nvscaleaddmulti = tf ? N_VScaleAddMulti_B200 : NULL;
  return SUN_SUCCESS;
}

SUNErrCode N_VEnableDotProdMulti_B200(N_Vector v, sunbooleantype tf)
{
  v->ops->nvdotprodmulti      = tf ? N_VDotProdMulti_B200 : NULL;
  v->ops->nvdotprodmultilocal = tf ? N_VDotProdMultiLocal_B200 : NULL;
  return SUN_SUCCESS;
}

SUNErrCode N_VEnableLinearSumVectorArray_B200(N_Vector v, sunbooleantype tf)
{
  v->ops->nvlinearsumvectorarray = tf ? N_VLinearSumVectorArray_B200 : NULL;
  return SUN_SUCCESS;
}

SUNErrCode N_VEnableScaleVectorArray_B200(N_Vector v, sunbooleantype tf)
{
  v->ops->nvscalevectorarray = tf ? N_VScaleVectorArray_B200 : NULL;
  return SUN_SUCCESS;
}

SUNErrCode N_VEnableConstVectorArray_B200(N_Vector v, sunbooleantype tf)
{
  v->ops->nvconstvectorarray = tf ? N_VConstVectorArray_B200 : NULL;
  return SUN_SUCCESS;
}

SUNErrCode N_VEnableWrmsNormVectorArray_B200(N_Vector v, sunbooleantype tf)
{
  v->ops->nvwrmsnormvectorarray = tf ? N_VWrmsNormVectorArray_B200 : NULL;
  return SUN_SUCCESS;
}

SUNErrCode N_VEnableWrmsNormMaskVectorArray_B200(N_Vector v, sunbooleantype tf)
{
  v->ops->nvwrmsnormmaskvectorarray = tf ? N_VWrmsNormMaskVectorArray_B200 : NULL;
  return SUN_SUCCESS;
}

SUNErrCode N_VEnableScaleAddMultiVectorArray_B200(N_Vector v, sunbooleantype tf)
{
  v->ops->nvscaleaddmultivectorarray = tf ? N_VScaleAddMultiVectorArray_B200 : NULL;
  return SUN_SUCCESS;
}

SUNErrCode N_VEnableLinearCombinationVectorArray_B200(N_Vector v, sunbooleantype tf)
{
  v->ops->nvlinearcombinationvectorarray = tf ? N_VLinearCombinationVectorArray_B200 : NULL;
  return SUN_SUCCESS;
}
