/* b200vec_internal.h -- host-side internals shared by the .cu translation units.
 * Not installed; the public C ABI is include/b200vec.h. */
#ifndef B200VEC_INTERNAL_H
#define B200VEC_INTERNAL_H

#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <vector>

#include "b200vec.h"

namespace b200 {

constexpr int kBlock        = 256;  /* threads per CTA for every kernel               */
constexpr int kSMs          = 148;  /* B200: 2 dies x 74 SMs                          */
constexpr int kRBlock       = 512;  /* threads per CTA of the single-output reduction kernels */
constexpr int kMaxBlocksDef = kSMs * 2; /* reductions: 2 CTAs x 512 threads per SM measured best
                                           (profiles/r01_mb_reduce_24b.txt) */
constexpr int kMaxPartialBlocks = 4096; /* reduction partial rows (>= any max_blocks) */
constexpr int kMaxOut       = 24;   /* outputs per multi-reduction launch (GMRES maxl = 20 needs a 21-wide
                                       classical Gram-Schmidt multi-dot, sunlinsol_spgmr.c:790)  */
constexpr int kMaxPair      = 8;    /* outputs per launch of the two-operand-per-output modes (A_j, B_j pairs) */
constexpr int kMaxRows      = 64;   /* result slots per context                       */
/* pinned host area: [0, kMaxRows) doubles = staging of b200vec_result_fetch; then 2 tagged 8-byte
   words per result slot -- word = (sequence << 32) | 32 bits of the value -- which a reduction
   kernel's final pass stores and the host polls.  Every 8-byte store is single-copy atomic on
   PCIe and carries its own tag, so no fence and no separate flag is needed (NCCL-LL style). */
constexpr int kHostSlots    = kMaxRows + 2 * kMaxRows;
constexpr int kExactMaxElems = 4096; /* smem doubles available to the exact-order path */

struct Tuning
{
  int64_t max_blocks      = kMaxBlocksDef; /* reductions (partials rows)              */
  int64_t stream_max_blocks = 0;           /* streaming kernels; 0 = one tile per CTA  */
  int64_t vec_width       = 0; /* 0 = auto (widest the alignment allows) */
  int64_t unroll          = 0; /* 0 = auto */
  int64_t exact_threshold = 1024;
  int64_t count_launches  = 0;
  int64_t spin_wait       = 1; /* poll the pinned sequence word instead of cudaStreamSynchronize */
  int64_t p2p             = 1; /* global reductions exchange partials over NVLink peer memory inside
                                  the reduction kernel (0: ncclAllReduce after the kernel)         */
  int64_t l2_prefetch     = 0; /* single-output reductions: tiles of cp.async.bulk.prefetch.L2 look-ahead per
                                  CTA (0 = off); bytes in flight then do not depend on how many wide loads
                                  ptxas keeps outstanding per thread                                       */
  int64_t profile         = 0; /* accumulate device-side wait times of the cross-rank exchanges
                                  (read back with get_tuning "prof_xwait_ns" / "prof_xwait_calls")        */
  int64_t pdl             = 1; /* programmatic dependent launch: a kernel's launch ramp overlaps the
                                  tail of its predecessor on the stream (griddepcontrol)          */
};

} // namespace b200

namespace b200 {
/* cross-rank exchange over NVLink peer memory (b200vec_comm.cu sets it up, the
   reduction kernels' epilogue uses it).  Every rank owns a 2 KiB mailbox in its
   HBM, mapped into every peer through CUDA IPC:
     word[((parity*kMaxPeers + src)*kMaxOut + slot)*2 + half] = (seq << 32) | 32 bits of the value
   so each 8-byte store carries its own sequence tag (NCCL-LL style): no fence,
   no separate flag, one NVLink hop.  Two parities make reuse safe (a rank can be
   at most one collective ahead of any peer). */
constexpr int kMaxPeers     = 8;
constexpr size_t kMboxWords = (size_t)2 * kMaxPeers * kMaxOut * 2;
struct XArgs
{
  unsigned long long* mbox[kMaxPeers]; /* mbox[r]: rank r's mailbox as mapped in this process */
  int nranks;                          /* 1 = local reduction, no exchange                    */
  int rank;
  unsigned int seq; /* collective sequence number, identical on all ranks (SPMD) */
  unsigned long long* prof; /* NULL, or the context's profile counters ("profile" tuning key):
                               [0] ns spent between posting the own partial and having all peers', [1] calls;
                               [6] %globaltimer of the earliest CTA of the last single-output reduction
                               (atomicMin; the host resets it), [7] %globaltimer at its publication */
};
} // namespace b200

struct b200vec_ctx_s
{
  int device            = 0;
  cudaStream_t stream   = nullptr;
  int refcount          = 1;
  b200::Tuning tune;
  int64_t launches      = 0;      /* resettable ("count_launches" tuning key)        */
  int64_t launches_total = 0;     /* since the context was created (B200VEC_REPORT)  */
  double t_created       = 0.0;

  /* reduction workspace (device) */
  double* d_partials    = nullptr; /* [kMaxOut][kMaxPartialBlocks]  (multi-output kernels, ticket scheme) */
  unsigned long long* d_tagged = nullptr; /* [kMaxPartialBlocks][2] tagged CTA partials of the single-output
                                      kernels: CTA 0 polls them, no ticket, no fence       */
  unsigned int* d_count = nullptr; /* [kMaxRows] last-block-done tickets, self-resetting */
  double* d_result      = nullptr; /* [kMaxRows] result slots                       */
  unsigned long long* d_prof = nullptr; /* [8] profile counters (XArgs::prof; apps add theirs from [2]) */
  /* pinned + mapped host mirror of the result slots: the final pass of every
     reduction kernel stores here directly, so a scalar-returning op costs one
     stream sync and no memcpy */
  double* h_result      = nullptr; /* host address   */
  double* h_result_dev  = nullptr; /* device alias   */
  unsigned long long seq = 0;      /* reductions launched so far; its low 32 bits tag the published words */

  /* copy stream of b200vec_copy_h2d_async (lazily created) */
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_compute   = nullptr; /* "ctx stream reached this point": the copy stream waits for it */
  cudaEvent_t ev_copy      = nullptr; /* "copies issued so far are done": b200vec_copy_join waits for it */
  bool copies_pending      = false;

  /* exact-size free-list cache of device allocations (clone/destroy churn) */
  std::multimap<size_t, void*> cache;
  size_t cached_bytes = 0;
  size_t cache_limit  = (size_t)8 << 30;

  /* communicator (NULL = single rank) */
  void* nccl_comm = nullptr;
  int rank        = 0;
  int nranks      = 1;
  double* d_commbuf = nullptr;
  /* peer-memory transport (NVLink): own mailbox + IPC mappings of the peers' */
  bool p2p_ready               = false;
  unsigned long long* mbox[b200::kMaxPeers] = {nullptr};
  unsigned int xseq            = 0;     /* collectives issued so far (same on all ranks)   */
  bool scope_global            = false; /* one-shot: the next reduction call is global     */
};

namespace b200 {

/* launch geometry shared by the streaming / reduction / fused launchers */
struct MapCfg
{
  int W;    /* doubles per load/store: 4 = 256-bit, 2 = 128-bit, 1 = 64-bit */
  int U;    /* independent wide loads in flight per operand and thread        */
  int grid; /* CTAs                                                           */
};
MapCfg pick_map_cfg(b200vec_ctx ctx, int64_t n, int wmax, bool reduction, int block = kBlock);
static inline int align_width(const void* p)
{
  if (!p) return 4; /* absent operand does not constrain */
  uintptr_t a = (uintptr_t)p;
  return (a % 32 == 0) ? 4 : (a % 16 == 0) ? 2 : 1;
}
int finish_reduction(b200vec_ctx ctx, int count, double* result_host);
/* consume the one-shot global scope; fills the exchange arguments when the peer
   transport carries this reduction.  Returns: 0 local, 1 global via peer memory
   (x filled), 2 global via ncclAllReduce after the kernel */
int take_scope(b200vec_ctx ctx, XArgs* x);
void next_xargs(b200vec_ctx ctx, XArgs* x); /* a further collective of the same call */
int linear_sum_dispatch(b200vec_ctx ctx, double a, const double* x, double b, const double* y, double* z,
                        bool z_is_x, bool z_is_y, int64_t n);
int scale_dispatch(b200vec_ctx ctx, double c, const double* x, double* z, int64_t n);

int set_error(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int check_launch(b200vec_ctx ctx, const char* kernel);

/* Every kernel of the library is launched through this: programmatic dependent
   launch lets the next kernel's CTAs become resident (and run their prologue up
   to griddepcontrol.wait) while the previous kernel on the stream drains --
   measured -2 us per launch at n = 2^24 and -25% on chains of small kernels
   (profiles/r01_mb_reduce_24b.txt).  Stream order is preserved: every kernel
   executes griddepcontrol.wait before its first global-memory access. */
template <class... KArgs, class... Args>
static inline void launch_k(b200vec_ctx ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, Args&&... args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim            = grid;
  cfg.blockDim           = block;
  cfg.dynamicSmemBytes   = 0;
  cfg.stream             = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs                                        = at;
  cfg.numAttrs                                     = ctx->tune.pdl ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<Args&&>(args)...); /* error picked up by check_launch */
}

/* RAII device guard so a context bound to device k works from any thread */
struct DeviceGuard
{
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev)
  {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev)
    {
      cudaSetDevice(dev);
      switched = true;
    }
  }
  ~DeviceGuard()
  {
    if (switched) cudaSetDevice(prev);
  }
};

} // namespace b200

#define B200_CHECK_CTX(ctx)                                                   \
  do {                                                                        \
    if ((ctx) == nullptr) return b200::set_error(B200VEC_ERR_ARG, "%s: NULL context", __func__); \
  } while (0)

#endif
