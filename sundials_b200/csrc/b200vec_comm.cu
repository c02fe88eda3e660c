/* b200vec_comm.cu -- multi-GPU plumbing: one rank (process) per GPU, contiguous
 * 1-D partition of the vector (the MPIPlusX pattern of
 * src/nvector/mpiplusx/nvector_mpiplusx.c:30).  The ONLY communication a vector
 * op needs is an allreduce of 1..nv doubles that the local reduction kernel left
 * in the context's device result slots -- the MPI_Allreduce call sites of
 * src/nvector/manyvector/nvector_manyvector.c:815,869,956,1050,1107,1128,1203,
 * 1277,1339,1399,1461,1576,1749,1793 -- done here with ncclAllReduce on the
 * context stream (NVLink 5 / NVSwitch), followed by one D2H of the slots.
 * One communicator per context, shared by every clone (the reference
 * MPI_Comm_dup's per clone, nvector_manyvector.c:195,2088).
 *
 * NCCL is loaded at run time (dlopen "libnccl.so.2"): single-GPU users have no
 * NCCL dependency, and inside a PyTorch process the already-loaded
 * torch-bundled NCCL is picked up by SONAME.
 */
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "b200vec_internal.h"

namespace b200 {

/* minimal NCCL ABI (stable since NCCL 2.x) */
typedef struct ncclComm* ncclComm_t;
typedef struct
{
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
enum
{
  nccl_Int64   = 4,
  nccl_Float64 = 8
};
enum
{
  nccl_Sum = 0,
  nccl_Max = 2,
  nccl_Min = 3
};

struct Nccl
{
  void* handle                                                                                  = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*)                                                    = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int)                             = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t)                                                       = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t)     = nullptr;
  const char* (*GetErrorString)(ncclResult_t)                                                   = nullptr;
};

static Nccl g_nccl;
static std::once_flag g_nccl_once;

static void load_nccl()
{
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names)
  {
    g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) return;
  g_nccl.GetUniqueId    = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.handle, "ncclGetUniqueId");
  g_nccl.CommInitRank   = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.handle, "ncclCommInitRank");
  g_nccl.CommDestroy    = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.handle, "ncclCommDestroy");
  g_nccl.AllReduce      = (decltype(g_nccl.AllReduce))dlsym(g_nccl.handle, "ncclAllReduce");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.handle, "ncclGetErrorString");
}

static int need_nccl()
{
  std::call_once(g_nccl_once, load_nccl);
  if (!g_nccl.handle || !g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
    return set_error(B200VEC_ERR_COMM, "NCCL (libnccl.so.2) could not be loaded: %s", dlerror());
  return B200VEC_OK;
}

static int check_nccl(ncclResult_t r, const char* what)
{
  if (r == 0) return B200VEC_OK;
  return set_error(B200VEC_ERR_COMM, "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
}

static int nccl_op(int op)
{
  return op == B200VEC_MAX ? nccl_Max : op == B200VEC_MIN ? nccl_Min : nccl_Sum;
}

} // namespace b200

using namespace b200;

extern "C" {

int b200vec_comm_get_unique_id(unsigned char id[B200VEC_UNIQUE_ID_BYTES])
{
  int rc = need_nccl();
  if (rc) return rc;
  ncclUniqueId u;
  rc = check_nccl(g_nccl.GetUniqueId(&u), "ncclGetUniqueId");
  if (rc) return rc;
  static_assert(sizeof(u) == B200VEC_UNIQUE_ID_BYTES, "ncclUniqueId size");
  memcpy(id, &u, sizeof(u));
  return B200VEC_OK;
}

int b200vec_comm_init(b200vec_ctx ctx, const unsigned char id[B200VEC_UNIQUE_ID_BYTES], int rank, int nranks)
{
  B200_CHECK_CTX(ctx);
  if (nranks < 1 || rank < 0 || rank >= nranks) return set_error(B200VEC_ERR_ARG, "comm_init: bad rank/nranks");
  if (ctx->nccl_comm) return set_error(B200VEC_ERR_ARG, "comm_init: context already has a communicator");
  ctx->rank   = rank;
  ctx->nranks = nranks;
  if (nranks == 1) return B200VEC_OK;
  int rc = need_nccl();
  if (rc) return rc;
  DeviceGuard g(ctx->device);
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclComm_t comm = nullptr;
  rc              = check_nccl(g_nccl.CommInitRank(&comm, nranks, u, rank), "ncclCommInitRank");
  if (rc) return rc;
  ctx->nccl_comm = comm;
  return check_cuda(cudaMalloc((void**)&ctx->d_commbuf, sizeof(double) * kMaxRows), "cudaMalloc(commbuf)");
}

int b200vec_comm_finalize(b200vec_ctx ctx)
{
  B200_CHECK_CTX(ctx);
  if (!ctx->nccl_comm) return B200VEC_OK;
  DeviceGuard g(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  int rc         = check_nccl(g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm), "ncclCommDestroy");
  ctx->nccl_comm = nullptr;
  ctx->nranks    = 1;
  ctx->rank      = 0;
  return rc;
}

int b200vec_comm_rank(b200vec_ctx ctx) { return ctx ? ctx->rank : 0; }
int b200vec_comm_size(b200vec_ctx ctx) { return ctx ? ctx->nranks : 1; }

int b200vec_allreduce_buffer(b200vec_ctx ctx, double* buf, int count, int op)
{
  B200_CHECK_CTX(ctx);
  if (count < 0 || (count > 0 && !buf)) return set_error(B200VEC_ERR_ARG, "allreduce: bad buffer");
  if (ctx->nranks == 1 || count == 0) return B200VEC_OK;
  if (!ctx->nccl_comm) return set_error(B200VEC_ERR_COMM, "allreduce: no communicator attached");
  DeviceGuard g(ctx->device);
  return check_nccl(g_nccl.AllReduce(buf, buf, (size_t)count, nccl_Float64, nccl_op(op), (ncclComm_t)ctx->nccl_comm,
                                     ctx->stream),
                    "ncclAllReduce");
}

int b200vec_allreduce(b200vec_ctx ctx, int count, int op)
{
  B200_CHECK_CTX(ctx);
  if (count < 0 || count > kMaxRows) return set_error(B200VEC_ERR_ARG, "allreduce: bad slot count");
  return b200vec_allreduce_buffer(ctx, ctx->d_result, count, op);
}

int b200vec_allreduce_i64_host(b200vec_ctx ctx, int64_t* value, int op)
{
  B200_CHECK_CTX(ctx);
  if (!value) return set_error(B200VEC_ERR_ARG, "allreduce_i64: NULL");
  if (ctx->nranks == 1) return B200VEC_OK;
  if (!ctx->nccl_comm) return set_error(B200VEC_ERR_COMM, "allreduce: no communicator attached");
  DeviceGuard g(ctx->device);
  int64_t* d = (int64_t*)ctx->d_commbuf;
  int rc     = check_cuda(cudaMemcpyAsync(d, value, sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream), "H2D(i64)");
  if (!rc)
    rc = check_nccl(g_nccl.AllReduce(d, d, 1, nccl_Int64, nccl_op(op), (ncclComm_t)ctx->nccl_comm, ctx->stream),
                    "ncclAllReduce(i64)");
  if (!rc) rc = check_cuda(cudaMemcpyAsync(value, d, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream), "D2H(i64)");
  if (!rc) rc = check_cuda(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
  return rc;
}

} /* extern "C" */
