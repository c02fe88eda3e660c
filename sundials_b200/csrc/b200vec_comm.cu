/* b200vec_comm.cu -- multi-GPU plumbing: one rank (process) per GPU, contiguous
 * 1-D partition of the vector (the MPIPlusX pattern of
 * src/nvector/mpiplusx/nvector_mpiplusx.c:30).  The ONLY communication a vector
 * op needs is an allreduce of 1..nv doubles -- the MPI_Allreduce call sites of
 * src/nvector/manyvector/nvector_manyvector.c:815,869,956,1050,1107,1128,1203,
 * 1277,1339,1399,1461,1576,1749,1793.  Two transports:
 *
 *  peer memory (default): every rank owns a small mailbox in HBM that all peers
 *     map through CUDA IPC; the LAST CTA of the reduction kernel itself posts
 *     the rank's partial(s) into every peer's mailbox over NVLink and folds the
 *     peers' partials in rank order (xrank_combine_warp, b200vec_device.cuh):
 *     local reduction + allreduce + host publication are ONE kernel, no
 *     collective launch, no extra sync.
 *  NCCL: ncclAllReduce on the context stream after the local kernel, then one
 *     D2H of the slots.  Used when IPC peer mapping is unavailable or the
 *     "p2p" tuning knob is 0, and for bootstrap (handle exchange, int64 sums).
 *
 * One communicator per context, shared by every clone (the reference
 * MPI_Comm_dup's per clone, nvector_manyvector.c:195,2088).
 *
 * NCCL is loaded at run time (dlopen "libnccl.so.2"): single-GPU users have no
 * NCCL dependency, and inside a PyTorch process the already-loaded
 * torch-bundled NCCL is picked up by SONAME.
 */
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "b200vec_device.cuh"

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
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t)          = nullptr;
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
  g_nccl.AllGather      = (decltype(g_nccl.AllGather))dlsym(g_nccl.handle, "ncclAllGather");
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

/* ---- peer-memory transport --------------------------------------------- */

/* Symmetric peer allocation (collective): every rank allocates `bytes` of HBM,
   zeroes them, and maps the allocations of all peers through CUDA IPC (handles
   allgathered over the already initialised NCCL communicator).  All ranks must
   agree, so the success flags are min-reduced.  On success ptrs[r] is rank r's
   buffer as addressable from THIS process (ptrs[me] = the own buffer) and
   *ok_out = 1; otherwise nothing stays allocated and *ok_out = 0. */
static int peer_alloc(b200vec_ctx ctx, size_t bytes, void** ptrs, int* ok_out)
{
  const int nr = ctx->nranks, me = ctx->rank;
  *ok_out = 0;
  for (int r = 0; r < nr && r < kMaxPeers; r++) ptrs[r] = nullptr;
  const char* env = getenv("B200VEC_P2P");
  const int want  = (nr <= kMaxPeers) && !(env && env[0] == '0') && g_nccl.AllGather != nullptr;
  void* own       = nullptr;
  cudaIpcMemHandle_t* d_handles = nullptr;
  std::vector<cudaIpcMemHandle_t> handles(nr);
  int ok = want;
  /* at least a whole 2 MiB block of its own: the IPC handle then maps exactly this allocation */
  const size_t alloc = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
  if (ok && cudaMalloc(&own, alloc) != cudaSuccess) ok = 0;
  if (ok && cudaMemsetAsync(own, 0, alloc, ctx->stream) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&handles[me], own) != cudaSuccess) ok = 0;
  (void)cudaGetLastError();
  /* every rank takes part in the exchange even if its own setup failed */
  int rc = check_cuda(cudaMalloc((void**)&d_handles, sizeof(cudaIpcMemHandle_t) * nr), "cudaMalloc(ipc handles)");
  if (rc) return rc;
  if (g_nccl.AllGather)
  {
    rc = check_cuda(cudaMemcpyAsync(d_handles + me, &handles[me], sizeof(cudaIpcMemHandle_t), cudaMemcpyHostToDevice,
                                    ctx->stream),
                    "H2D(ipc handle)");
    if (!rc)
      rc = check_nccl(g_nccl.AllGather(d_handles + me, d_handles, sizeof(cudaIpcMemHandle_t), /*ncclInt8*/ 0,
                                       (ncclComm_t)ctx->nccl_comm, ctx->stream),
                      "ncclAllGather(ipc handles)");
    if (!rc)
      rc = check_cuda(cudaMemcpyAsync(handles.data(), d_handles, sizeof(cudaIpcMemHandle_t) * nr,
                                      cudaMemcpyDeviceToHost, ctx->stream),
                      "D2H(ipc handles)");
    if (!rc) rc = check_cuda(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize(ipc handles)");
  }
  cudaFree(d_handles);
  if (rc) return rc;
  int64_t all_ok = ok;
  rc             = b200vec_allreduce_i64_host(ctx, &all_ok, B200VEC_MIN);
  if (rc) return rc;
  if (all_ok)
  {
    for (int r = 0; r < nr && ok; r++)
    {
      if (r == me)
      {
        ptrs[r] = own;
        continue;
      }
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, handles[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
      {
        (void)cudaGetLastError();
        ok = 0;
      }
      ptrs[r] = p;
    }
    all_ok = ok;
    rc     = b200vec_allreduce_i64_host(ctx, &all_ok, B200VEC_MIN); /* also the barrier before first use */
    if (rc) return rc;
  }
  if (all_ok)
  {
    *ok_out = 1;
    return B200VEC_OK;
  }
  for (int r = 0; r < nr && r < kMaxPeers; r++)
  {
    if (r != me && ptrs[r]) cudaIpcCloseMemHandle(ptrs[r]);
    ptrs[r] = nullptr;
  }
  if (own) cudaFree(own);
  (void)cudaGetLastError();
  return B200VEC_OK;
}

static void peer_free(b200vec_ctx ctx, void** ptrs)
{
  for (int r = 0; r < ctx->nranks && r < kMaxPeers; r++)
  {
    if (!ptrs[r]) continue;
    if (r == ctx->rank) cudaFree(ptrs[r]);
    else cudaIpcCloseMemHandle(ptrs[r]);
    ptrs[r] = nullptr;
  }
  (void)cudaGetLastError();
}

static int setup_peer_mailboxes(b200vec_ctx ctx)
{
  ctx->p2p_ready = false;
  int ok         = 0;
  int rc         = peer_alloc(ctx, kMboxWords * sizeof(unsigned long long), (void**)ctx->mbox, &ok);
  if (rc) return rc;
  ctx->p2p_ready = (ok != 0);
  ctx->xseq      = 0;
  if (!ok && getenv("B200VEC_VERBOSE"))
    fprintf(stderr, "[b200vec] rank %d: peer-memory transport unavailable, using NCCL\n", ctx->rank);
  return B200VEC_OK;
}

static void teardown_peer_mailboxes(b200vec_ctx ctx)
{
  peer_free(ctx, (void**)ctx->mbox);
  ctx->p2p_ready = false;
}

static void fill_xargs(b200vec_ctx ctx, XArgs* x)
{
  for (int r = 0; r < kMaxPeers; r++) x->mbox[r] = ctx->mbox[r];
  x->nranks = ctx->nranks;
  x->rank   = ctx->rank;
  x->seq    = ++ctx->xseq;
  /* tag 0 is the "never written" value.  At the 32-bit wrap skip TWO values (0xffffffff -> 2) so that
     consecutive collectives keep alternating mailbox halves (parity = seq & 1): skipping only 0 would
     put two odd tags back to back and let a fast rank overwrite a slot a slower peer still polls. */
  if (x->seq == 0) x->seq = ctx->xseq = 2;
  x->prof   = ctx->tune.profile ? ctx->d_prof : nullptr;
}

int take_scope(b200vec_ctx ctx, XArgs* x)
{
  const bool global = ctx->scope_global && ctx->nranks > 1;
  ctx->scope_global = false;
  for (int r = 0; r < kMaxPeers; r++) x->mbox[r] = nullptr;
  x->nranks = 1;
  x->rank   = 0;
  x->seq    = 0;
  x->prof   = ctx->tune.profile ? ctx->d_prof : nullptr; /* also local reductions stamp start / publication */
  if (!global) return 0;
  if (ctx->p2p_ready && ctx->tune.p2p)
  {
    fill_xargs(ctx, x);
    return 1;
  }
  return 2;
}

void next_xargs(b200vec_ctx ctx, XArgs* x) { fill_xargs(ctx, x); }

/* standalone exchange of result slots [0,count) (count <= kMaxOut): warp j folds
   slot j across ranks in place */
template <int OP>
__global__ void __launch_bounds__(kBlock) k_xrank(double* d_res, int count, const __grid_constant__ XArgs x)
{
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (int j = threadIdx.x >> 5; j < count; j += kBlock / 32)
  {
    const double v = d_res[j];
    double r;
    if (OP == B200VEC_MAX) r = xrank_combine_warp<CombMax>(v, j, x);
    else if (OP == B200VEC_MIN) r = xrank_combine_warp<CombMin>(v, j, x);
    else r = xrank_combine_warp<CombSum>(v, j, x);
    if ((threadIdx.x & 31) == 0) d_res[j] = r;
  }
}

} // namespace b200

using namespace b200;

extern "C" {

int b200vec_ctx_set_scope(b200vec_ctx ctx, int scope)
{
  B200_CHECK_CTX(ctx);
  ctx->scope_global = (scope == B200VEC_SCOPE_GLOBAL);
  return B200VEC_OK;
}

const char* b200vec_comm_transport(b200vec_ctx ctx)
{
  if (!ctx || ctx->nranks <= 1) return "none";
  return (ctx->p2p_ready && ctx->tune.p2p) ? "peer-memory" : "nccl";
}

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
  rc = check_cuda(cudaMalloc((void**)&ctx->d_commbuf, sizeof(double) * kMaxRows), "cudaMalloc(commbuf)");
  if (rc) return rc;
  return setup_peer_mailboxes(ctx);
}

int b200vec_comm_finalize(b200vec_ctx ctx)
{
  B200_CHECK_CTX(ctx);
  if (!ctx->nccl_comm) return B200VEC_OK;
  DeviceGuard g(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  teardown_peer_mailboxes(ctx);
  int rc         = check_nccl(g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm), "ncclCommDestroy");
  ctx->nccl_comm = nullptr;
  ctx->nranks    = 1;
  ctx->rank      = 0;
  return rc;
}

int b200vec_comm_peer_alloc(b200vec_ctx ctx, size_t bytes, void** ptrs)
{
  B200_CHECK_CTX(ctx);
  if (!ptrs || bytes == 0) return set_error(B200VEC_ERR_ARG, "comm_peer_alloc: bad argument");
  DeviceGuard g(ctx->device);
  if (ctx->nranks == 1)
  {
    int rc = check_cuda(cudaMalloc(&ptrs[0], bytes), "cudaMalloc(peer_alloc)");
    if (!rc) rc = check_cuda(cudaMemsetAsync(ptrs[0], 0, bytes, ctx->stream), "cudaMemsetAsync(peer_alloc)");
    return rc;
  }
  if (!ctx->nccl_comm) return set_error(B200VEC_ERR_COMM, "comm_peer_alloc: no communicator attached");
  int ok = 0;
  int rc = peer_alloc(ctx, bytes, ptrs, &ok);
  if (rc) return rc;
  if (!ok) return set_error(B200VEC_ERR_COMM, "comm_peer_alloc: CUDA IPC peer mapping unavailable on this system");
  return B200VEC_OK;
}

int b200vec_comm_peer_free(b200vec_ctx ctx, void** ptrs)
{
  B200_CHECK_CTX(ctx);
  if (!ptrs) return B200VEC_OK;
  DeviceGuard g(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->nranks == 1)
  {
    if (ptrs[0]) cudaFree(ptrs[0]);
    ptrs[0] = nullptr;
    return B200VEC_OK;
  }
  peer_free(ctx, ptrs);
  return B200VEC_OK;
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
  if (ctx->nranks == 1 || count == 0) return B200VEC_OK;
  if (!(ctx->p2p_ready && ctx->tune.p2p)) return b200vec_allreduce_buffer(ctx, ctx->d_result, count, op);
  DeviceGuard g(ctx->device);
  for (int j0 = 0; j0 < count; j0 += kMaxOut)
  {
    const int nj = (count - j0 < kMaxOut) ? count - j0 : kMaxOut;
    XArgs xa;
    fill_xargs(ctx, &xa);
    if (op == B200VEC_MAX) launch_k(ctx, k_xrank<B200VEC_MAX>, dim3(1), dim3(kBlock), ctx->d_result + j0, nj, xa);
    else if (op == B200VEC_MIN) launch_k(ctx, k_xrank<B200VEC_MIN>, dim3(1), dim3(kBlock), ctx->d_result + j0, nj, xa);
    else launch_k(ctx, k_xrank<B200VEC_SUM>, dim3(1), dim3(kBlock), ctx->d_result + j0, nj, xa);
    int rc = check_launch(ctx, "xrank");
    if (rc) return rc;
  }
  return B200VEC_OK;
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
