/* b200vec_ctx.cu -- execution context, workspace, memory and error plumbing
 * of the B200-native N_Vector kernel library (C ABI in include/b200vec.h).
 *
 * Replaces, for this vector, what the reference spreads over
 *   - SUNCudaExecPolicy objects      (include/sundials/sundials_cuda_policies.hpp:71-235)
 *   - per-vector reduction / fused scratch (src/nvector/cuda/nvector_cuda.cu:2277-2630)
 *   - SUNMemoryHelper_Cuda           (src/sunmemory/cuda/sundials_cuda_memory.cu:131-351)
 * with ONE workspace per context that every clone shares.
 */
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <vector>

#include "b200vec_internal.h"

namespace b200 {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  if (getenv("B200VEC_VERBOSE")) fprintf(stderr, "[b200vec] error %d: %s\n", code, g_err);
  return code;
}

int check_cuda(cudaError_t e, const char* what)
{
  if (e == cudaSuccess) return B200VEC_OK;
  int code = (e == cudaErrorMemoryAllocation) ? B200VEC_ERR_NOMEM
             : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? B200VEC_ERR_NODEVICE
                                                                            : B200VEC_ERR_CUDA;
  return set_error(code, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

int check_launch(b200vec_ctx ctx, const char* kernel)
{
  ctx->launches++;
  ctx->launches_total++;
  return check_cuda(cudaGetLastError(), kernel);
}

/* B200VEC_REPORT=1: when the process ends, one line per context on stderr with its lifetime (context
   creation -> exit, i.e. WITHOUT CUDA initialisation and process start-up) and the kernels it launched:
   what separates "the program took 0.9 s" from "the vector spent 0.4 s" on launch-bound problems */
static std::mutex g_report_mu;
static std::vector<b200vec_ctx> g_report_ctx;
static double now_s()
{
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static void report_at_exit()
{
  for (b200vec_ctx c : g_report_ctx)
    fprintf(stderr, "[b200vec] report: {\"ctx_lifetime_s\": %.4f, \"kernel_launches\": %lld, \"device\": %d}\n",
            now_s() - c->t_created, (long long)c->launches_total, c->device);
}
static void report_register(b200vec_ctx c)
{
  const char* e = getenv("B200VEC_REPORT");
  c->t_created  = now_s();
  if (!e || !e[0] || e[0] == '0') return;
  std::lock_guard<std::mutex> lk(g_report_mu);
  if (g_report_ctx.empty()) atexit(report_at_exit);
  c->refcount++; /* stays alive until the report has been printed */
  g_report_ctx.push_back(c);
}

} // namespace b200

using namespace b200;

extern "C" {

const char* b200vec_last_error(void) { return g_err; }
const char* b200vec_version(void) { return "b200vec 0.1 (sm_100a, fp64, -fmad=false)"; }

int b200vec_ctx_create(b200vec_ctx* out, int device, void* stream)
{
  if (!out) return set_error(B200VEC_ERR_ARG, "ctx_create: NULL out");
  *out    = nullptr;
  int ndev = 0;
  int rc   = check_cuda(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount");
  if (rc) return rc;
  if (ndev < 1) return set_error(B200VEC_ERR_NODEVICE, "no CUDA device visible");
  if (device < 0)
  {
    rc = check_cuda(cudaGetDevice(&device), "cudaGetDevice");
    if (rc) return rc;
  }
  if (device >= ndev) return set_error(B200VEC_ERR_ARG, "device %d out of range (%d visible)", device, ndev);

  b200vec_ctx c = new (std::nothrow) b200vec_ctx_s();
  if (!c) return set_error(B200VEC_ERR_NOMEM, "ctx_create: host allocation failed");
  c->device = device;
  c->stream = (cudaStream_t)stream;
  DeviceGuard g(device);

  rc = check_cuda(cudaMalloc((void**)&c->d_partials, sizeof(double) * kMaxOut * kMaxPartialBlocks),
                  "cudaMalloc(partials)");
  if (!rc)
    rc = check_cuda(cudaMalloc((void**)&c->d_tagged, sizeof(unsigned long long) * 2 * kMaxPartialBlocks),
                    "cudaMalloc(tagged partials)");
  if (!rc)
    rc = check_cuda(cudaMemset(c->d_tagged, 0, sizeof(unsigned long long) * 2 * kMaxPartialBlocks),
                    "cudaMemset(tagged partials)");
  if (!rc) rc = check_cuda(cudaMalloc((void**)&c->d_count, sizeof(unsigned int) * kMaxRows), "cudaMalloc(count)");
  if (!rc) rc = check_cuda(cudaMemset(c->d_count, 0, sizeof(unsigned int) * kMaxRows), "cudaMemset(count)");
  if (!rc) rc = check_cuda(cudaMalloc((void**)&c->d_result, sizeof(double) * kMaxRows), "cudaMalloc(result)");
  if (!rc) rc = check_cuda(cudaMemset(c->d_result, 0, sizeof(double) * kMaxRows), "cudaMemset(result)");
  if (!rc) rc = check_cuda(cudaMalloc((void**)&c->d_prof, sizeof(unsigned long long) * 8), "cudaMalloc(prof)");
  if (!rc) rc = check_cuda(cudaMemset(c->d_prof, 0, sizeof(unsigned long long) * 8), "cudaMemset(prof)");
  if (!rc)
    rc = check_cuda(cudaHostAlloc((void**)&c->h_result, sizeof(double) * kHostSlots, cudaHostAllocMapped),
                    "cudaHostAlloc(result)");
  if (!rc)
    rc = check_cuda(cudaHostGetDevicePointer((void**)&c->h_result_dev, c->h_result, 0),
                    "cudaHostGetDevicePointer(result)");
  if (rc)
  {
    b200vec_ctx_release(c);
    return rc;
  }
  memset(c->h_result, 0, sizeof(double) * kHostSlots);
  report_register(c);
  *out = c;
  return B200VEC_OK;
}

int b200vec_ctx_retain(b200vec_ctx ctx)
{
  B200_CHECK_CTX(ctx);
  ctx->refcount++;
  return B200VEC_OK;
}

int b200vec_ctx_release(b200vec_ctx ctx)
{
  B200_CHECK_CTX(ctx);
  if (--ctx->refcount > 0) return B200VEC_OK;
  DeviceGuard g(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream)
  {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
    cudaEventDestroy(ctx->ev_compute);
    cudaEventDestroy(ctx->ev_copy);
  }
  if (ctx->nccl_comm) b200vec_comm_finalize(ctx);
  for (auto& kv : ctx->cache) cudaFree(kv.second);
  ctx->cache.clear();
  if (ctx->d_partials) cudaFree(ctx->d_partials);
  if (ctx->d_tagged) cudaFree(ctx->d_tagged);
  if (ctx->d_count) cudaFree(ctx->d_count);
  if (ctx->d_result) cudaFree(ctx->d_result);
  if (ctx->d_prof) cudaFree(ctx->d_prof);
  if (ctx->d_commbuf) cudaFree(ctx->d_commbuf);
  if (ctx->h_result) cudaFreeHost(ctx->h_result);
  delete ctx;
  return B200VEC_OK;
}

static std::mutex g_default_mu;
static std::map<int, b200vec_ctx> g_default;

int b200vec_ctx_default(b200vec_ctx* out)
{
  if (!out) return set_error(B200VEC_ERR_ARG, "ctx_default: NULL out");
  int dev = 0;
  int rc  = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(g_default_mu);
  auto it = g_default.find(dev);
  if (it == g_default.end())
  {
    b200vec_ctx c = nullptr;
    rc            = b200vec_ctx_create(&c, dev, nullptr);
    if (rc) return rc;
    it = g_default.emplace(dev, c).first; /* lives for the whole process */
  }
  *out = it->second;
  return B200VEC_OK;
}

int b200vec_ctx_set_stream(b200vec_ctx ctx, void* stream)
{
  B200_CHECK_CTX(ctx);
  DeviceGuard g(ctx->device);
  if (ctx->stream == (cudaStream_t)stream) return B200VEC_OK;
  /* pending work on the old stream must finish before results on the new one can depend on it.  The old
     handle itself is not touched: its owner may have destroyed it already (the reference's own unit-test
     driver destroys its stream and then re-policies fresh vectors, test_nvector_cuda.cu:94,158,394), and
     passing a destroyed stream to the runtime is undefined.  Waiting for the whole device covers it. */
  int rc      = check_cuda(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
  ctx->stream = (cudaStream_t)stream;
  return rc;
}

void* b200vec_ctx_get_stream(b200vec_ctx ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int b200vec_ctx_device(b200vec_ctx ctx) { return ctx ? ctx->device : -1; }

int b200vec_ctx_sync(b200vec_ctx ctx)
{
  B200_CHECK_CTX(ctx);
  DeviceGuard g(ctx->device);
  return check_cuda(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
}

int b200vec_ctx_set_tuning(b200vec_ctx ctx, const char* key, int64_t value)
{
  B200_CHECK_CTX(ctx);
  if (!key) return set_error(B200VEC_ERR_ARG, "set_tuning: NULL key");
  if (!strcmp(key, "max_blocks"))
  {
    if (value < 1 || value > kMaxPartialBlocks)
      return set_error(B200VEC_ERR_ARG, "max_blocks must be in [1,%d]", kMaxPartialBlocks);
    ctx->tune.max_blocks = value;
  }
  else if (!strcmp(key, "stream_max_blocks"))
  {
    if (value < 0 || value > 0x7fffffff) return set_error(B200VEC_ERR_ARG, "stream_max_blocks must be >= 0");
    ctx->tune.stream_max_blocks = value;
  }
  else if (!strcmp(key, "vec_width"))
  {
    if (value != 0 && value != 1 && value != 2 && value != 4)
      return set_error(B200VEC_ERR_ARG, "vec_width must be 0,1,2 or 4");
    ctx->tune.vec_width = value;
  }
  else if (!strcmp(key, "unroll"))
  {
    if (value != 0 && value != 1 && value != 2 && value != 4)
      return set_error(B200VEC_ERR_ARG, "unroll must be 0,1,2 or 4");
    ctx->tune.unroll = value;
  }
  else if (!strcmp(key, "exact_threshold"))
  {
    if (value < 0 || value > kExactMaxElems)
      return set_error(B200VEC_ERR_ARG, "exact_threshold must be in [0,%d]", kExactMaxElems);
    ctx->tune.exact_threshold = value;
  }
  else if (!strcmp(key, "spin_wait")) ctx->tune.spin_wait = value ? 1 : 0;
  else if (!strcmp(key, "pdl")) ctx->tune.pdl = value ? 1 : 0;
  else if (!strcmp(key, "p2p")) ctx->tune.p2p = value ? 1 : 0;
  else if (!strcmp(key, "profile"))
  {
    ctx->tune.profile = value ? 1 : 0;
    DeviceGuard g(ctx->device);
    int rc = check_cuda(cudaMemsetAsync(ctx->d_prof, 0, sizeof(unsigned long long) * 8, ctx->stream), "cudaMemset(prof)");
    if (rc) return rc;
  }
  else if (!strcmp(key, "prof_stamp_reset"))
  { /* arm the start/publish stamps of the next single-output reduction ([6] = min over CTAs) */
    DeviceGuard g(ctx->device);
    const unsigned long long init[2] = {~0ull, 0ull};
    int rc = check_cuda(cudaMemcpyAsync(ctx->d_prof + 6, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream),
                        "cudaMemcpy(prof stamps)");
    if (!rc) rc = check_cuda(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
    if (rc) return rc;
  }
  else if (!strcmp(key, "l2_prefetch"))
  {
    if (value < 0 || value > 4) return set_error(B200VEC_ERR_ARG, "l2_prefetch must be in [0,4]");
    ctx->tune.l2_prefetch = value;
  }
  else if (!strcmp(key, "count_launches"))
  { /* (re)starts the resettable launch counter read by b200vec_ctx_launch_count */
    ctx->tune.count_launches = value ? 1 : 0;
    ctx->launches            = 0;
  }
  else return set_error(B200VEC_ERR_ARG, "unknown tuning key '%s'", key);
  return B200VEC_OK;
}

int64_t b200vec_ctx_get_tuning(b200vec_ctx ctx, const char* key)
{
  if (!ctx || !key) return -1;
  if (!strcmp(key, "max_blocks")) return ctx->tune.max_blocks;
  if (!strcmp(key, "stream_max_blocks")) return ctx->tune.stream_max_blocks;
  if (!strcmp(key, "vec_width")) return ctx->tune.vec_width;
  if (!strcmp(key, "unroll")) return ctx->tune.unroll;
  if (!strcmp(key, "exact_threshold")) return ctx->tune.exact_threshold;
  if (!strcmp(key, "count_launches")) return ctx->tune.count_launches;
  if (!strcmp(key, "spin_wait")) return ctx->tune.spin_wait;
  if (!strcmp(key, "pdl")) return ctx->tune.pdl;
  if (!strcmp(key, "p2p")) return ctx->tune.p2p;
  if (!strcmp(key, "l2_prefetch")) return ctx->tune.l2_prefetch;
  if (!strcmp(key, "profile")) return ctx->tune.profile;
  if (!strcmp(key, "prof_counters_ptr")) return (int64_t)(uintptr_t)ctx->d_prof;
  if (!strncmp(key, "prof_counter_", 13))
  {
    /* prof_counter_<i>: synchronises the stream and reads counter i */
    const int i = atoi(key + 13);
    if (i < 0 || i >= 8) return -1;
    DeviceGuard g(ctx->device);
    unsigned long long v = 0;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
    if (cudaMemcpy(&v, ctx->d_prof + i, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t)v;
  }
  return -1;
}

int64_t b200vec_ctx_launch_count(b200vec_ctx ctx) { return ctx ? ctx->launches : -1; }

/* ---------------------------------------------------------------- memory */

int b200vec_malloc_device(b200vec_ctx ctx, size_t bytes, void** ptr)
{
  B200_CHECK_CTX(ctx);
  if (!ptr) return set_error(B200VEC_ERR_ARG, "malloc_device: NULL ptr");
  *ptr = nullptr;
  if (bytes == 0) return B200VEC_OK;
  auto it = ctx->cache.find(bytes);
  if (it != ctx->cache.end())
  {
    *ptr = it->second;
    ctx->cache.erase(it);
    ctx->cached_bytes -= bytes;
    return B200VEC_OK;
  }
  DeviceGuard g(ctx->device);
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e == cudaErrorMemoryAllocation && !ctx->cache.empty())
  {
    /* give cached blocks back to the driver and retry once */
    (void)cudaGetLastError();
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->cache) cudaFree(kv.second);
    ctx->cache.clear();
    ctx->cached_bytes = 0;
    e                 = cudaMalloc(ptr, bytes);
  }
  return check_cuda(e, "cudaMalloc");
}

int b200vec_free_device(b200vec_ctx ctx, void* ptr, size_t bytes)
{
  B200_CHECK_CTX(ctx);
  if (!ptr) return B200VEC_OK;
  /* stream-ordered reuse is safe: every consumer of a cached block is enqueued
     on the same ctx stream after all earlier users */
  if (bytes > 0 && ctx->cached_bytes + bytes <= ctx->cache_limit)
  {
    ctx->cache.emplace(bytes, ptr);
    ctx->cached_bytes += bytes;
    return B200VEC_OK;
  }
  DeviceGuard g(ctx->device);
  return check_cuda(cudaFree(ptr), "cudaFree");
}

int b200vec_malloc_host(b200vec_ctx ctx, size_t bytes, void** ptr)
{
  B200_CHECK_CTX(ctx);
  if (!ptr) return set_error(B200VEC_ERR_ARG, "malloc_host: NULL ptr");
  *ptr = nullptr;
  if (bytes == 0) return B200VEC_OK;
  DeviceGuard g(ctx->device);
  return check_cuda(cudaHostAlloc(ptr, bytes, cudaHostAllocMapped | cudaHostAllocPortable), "cudaHostAlloc");
}

int b200vec_free_host(b200vec_ctx ctx, void* ptr)
{
  B200_CHECK_CTX(ctx);
  if (!ptr) return B200VEC_OK;
  DeviceGuard g(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  return check_cuda(cudaFreeHost(ptr), "cudaFreeHost");
}

int b200vec_malloc_managed(b200vec_ctx ctx, size_t bytes, void** ptr)
{
  B200_CHECK_CTX(ctx);
  if (!ptr) return set_error(B200VEC_ERR_ARG, "malloc_managed: NULL ptr");
  *ptr = nullptr;
  if (bytes == 0) return B200VEC_OK;
  DeviceGuard g(ctx->device);
  return check_cuda(cudaMallocManaged(ptr, bytes, cudaMemAttachGlobal), "cudaMallocManaged");
}

int b200vec_free_managed(b200vec_ctx ctx, void* ptr)
{
  B200_CHECK_CTX(ctx);
  if (!ptr) return B200VEC_OK;
  DeviceGuard g(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  return check_cuda(cudaFree(ptr), "cudaFree(managed)");
}

int b200vec_copy_h2d(b200vec_ctx ctx, void* dst, const void* src, size_t bytes, int sync)
{
  B200_CHECK_CTX(ctx);
  if (bytes == 0) return B200VEC_OK;
  DeviceGuard g(ctx->device);
  int rc = check_cuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(H2D)");
  if (!rc && sync) rc = check_cuda(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
  return rc;
}

int b200vec_copy_h2d_async(b200vec_ctx ctx, void* dst, const void* src, size_t bytes)
{
  B200_CHECK_CTX(ctx);
  if (bytes == 0) return B200VEC_OK;
  DeviceGuard g(ctx->device);
  int rc = B200VEC_OK;
  if (!ctx->copy_stream)
  {
    rc = check_cuda(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate(copy)");
    if (!rc) rc = check_cuda(cudaEventCreateWithFlags(&ctx->ev_compute, cudaEventDisableTiming), "cudaEventCreate");
    if (!rc) rc = check_cuda(cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming), "cudaEventCreate");
    if (rc) return rc;
  }
  rc = check_cuda(cudaEventRecord(ctx->ev_compute, ctx->stream), "cudaEventRecord(compute)");
  if (!rc) rc = check_cuda(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_compute, 0), "cudaStreamWaitEvent(copy)");
  if (!rc) rc = check_cuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->copy_stream), "cudaMemcpyAsync(H2D, copy stream)");
  if (!rc) rc = check_cuda(cudaEventRecord(ctx->ev_copy, ctx->copy_stream), "cudaEventRecord(copy)");
  if (!rc) ctx->copies_pending = true;
  return rc;
}

int b200vec_copy_join(b200vec_ctx ctx)
{
  B200_CHECK_CTX(ctx);
  if (!ctx->copies_pending) return B200VEC_OK;
  DeviceGuard g(ctx->device);
  ctx->copies_pending = false;
  return check_cuda(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0), "cudaStreamWaitEvent(join)");
}

int b200vec_copy_d2h(b200vec_ctx ctx, void* dst, const void* src, size_t bytes, int sync)
{
  B200_CHECK_CTX(ctx);
  if (bytes == 0) return B200VEC_OK;
  DeviceGuard g(ctx->device);
  int rc = check_cuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(D2H)");
  if (!rc && sync) rc = check_cuda(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
  return rc;
}

int b200vec_copy_d2d(b200vec_ctx ctx, void* dst, const void* src, size_t bytes)
{
  B200_CHECK_CTX(ctx);
  if (bytes == 0) return B200VEC_OK;
  DeviceGuard g(ctx->device);
  return check_cuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream), "cudaMemcpyAsync(D2D)");
}

double* b200vec_result_device(b200vec_ctx ctx) { return ctx ? ctx->d_result : nullptr; }

int b200vec_result_fetch(b200vec_ctx ctx, int count, double* result_host)
{
  B200_CHECK_CTX(ctx);
  if (count < 0 || count > kMaxRows || !result_host) return set_error(B200VEC_ERR_ARG, "result_fetch: bad count/ptr");
  DeviceGuard g(ctx->device);
  /* after an allreduce the pinned slots are stale: copy from the device slots */
  int rc = check_cuda(cudaMemcpyAsync(ctx->h_result, ctx->d_result, sizeof(double) * count, cudaMemcpyDeviceToHost,
                                      ctx->stream),
                      "cudaMemcpyAsync(result)");
  if (!rc) rc = check_cuda(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
  if (!rc)
    for (int i = 0; i < count; i++) result_host[i] = ctx->h_result[i];
  return rc;
}

} /* extern "C" */
