/* b200vec.h -- kernel-level C ABI of the B200-native N_Vector (libsundials_nvecb200.so)
 *
 * This is the thin C-ABI layer between the C host code that fills the SUNDIALS
 * N_Vector_Ops table (sundials_b200/csrc/nvector_b200.c, public header
 * include/nvector_b200.h) and the hand-written sm_100a CUDA kernels.  Plain
 * pointers and sizes only: `const double*` are DEVICE pointers unless a
 * parameter is named *_host; `int64_t n` is the (local) vector length;
 * coefficient arrays, pointer tables and result arrays are HOST memory owned by
 * the caller and only read/written during the call (they travel to the GPU as
 * by-value kernel parameters -- there is no per-op H2D copy and no sync for
 * streaming ops).
 *
 * Each entry point names the reference function it replaces
 * (SUNDIALS 7.5.0; "serial" = src/nvector/serial/nvector_serial.c, the parity
 * oracle; "cuda" = src/nvector/cuda/nvector_cuda.cu, the implementation this
 * replaces).  Arithmetic follows serial exactly (the library is built with
 * -fmad=false): streaming and fused-streaming ops are bit-identical to serial,
 * reductions are bit-identical for n <= the exact-order threshold (default 1024)
 * and deterministic pairwise (fixed order, independent of scheduling) above it.
 *
 * All functions return 0 (B200VEC_OK) or a negative B200VEC_ERR_* code; the text
 * of the last error on the calling thread is b200vec_last_error().
 * There is NO CPU fallback anywhere in this library.
 */
#ifndef B200VEC_H
#define B200VEC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200VEC_OK            0
#define B200VEC_ERR_ARG      (-1)  /* bad argument (nvec < 1, NULL pointer, n < 0)        */
#define B200VEC_ERR_CUDA     (-2)  /* a CUDA runtime call or kernel launch failed         */
#define B200VEC_ERR_NOMEM    (-3)  /* device / pinned allocation failed                   */
#define B200VEC_ERR_COMM     (-4)  /* NCCL not loadable or a collective failed            */
#define B200VEC_ERR_NODEVICE (-5)  /* no CUDA device visible                              */

/* opaque execution context: device, stream, reduction workspace, pinned result
 * slots, allocation cache, optional communicator.  One per SUNContext / stream;
 * shared by all vectors cloned from one another (replaces the per-vector
 * reduction + fused scratch buffers of cuda:2277-2630 and the exec-policy
 * objects of include/sundials/sundials_cuda_policies.hpp:71-235). */
typedef struct b200vec_ctx_s* b200vec_ctx;

/* reduction combiners for b200vec_allreduce (MPI_SUM/MAX/MIN of
 * src/nvector/manyvector/nvector_manyvector.c:815,869,1107) */
#define B200VEC_SUM 0
#define B200VEC_MAX 1
#define B200VEC_MIN 2

/* ------------------------------------------------------------------------
 * context / stream / workspace
 * ---------------------------------------------------------------------- */
/* device < 0: current device.  stream: a cudaStream_t (as void*), NULL = the
 * legacy default stream (what the reference uses, cuda:126-127), so user RHS
 * kernels launched on stream 0 stay ordered with vector ops. */
int b200vec_ctx_create(b200vec_ctx* out, int device, void* stream);
int b200vec_ctx_retain(b200vec_ctx ctx);
int b200vec_ctx_release(b200vec_ctx ctx); /* destroys when the last reference goes */
/* process-wide default context on the current device / legacy stream */
int b200vec_ctx_default(b200vec_ctx* out);
/* all later work of the context goes to `stream`; waits for the DEVICE first (not for the previous stream,
 * whose handle the caller may already have destroyed).  The caller keeps the stream alive while the context
 * uses it -- the lifetime rule of the reference's execution-policy objects (sundials_cuda_policies.hpp:71-90). */
int b200vec_ctx_set_stream(b200vec_ctx ctx, void* stream);
void* b200vec_ctx_get_stream(b200vec_ctx ctx);
int b200vec_ctx_device(b200vec_ctx ctx);
int b200vec_ctx_sync(b200vec_ctx ctx); /* cudaStreamSynchronize on the ctx stream */
/* tuning knobs (sweepable from the bench without recompiling):
 *  "max_blocks"      grid cap of the reduction kernels (default 148*2 CTAs of 512 threads; 148 below
 *                    2^20 elements; one tile per CTA for the reductions that also write a vector)
 *  "pdl"             1 (default): programmatic dependent launch on every kernel
 *  "p2p"             1 (default): global reductions exchange over NVLink peer memory
 *                    inside the reduction kernel; 0: ncclAllReduce (same on all ranks!)
 *  "spin_wait"       1 (default): scalar-returning ops poll the tagged pinned result words
 *                    written by the kernel's final pass; 0: cudaStreamSynchronize
 *  "profile"         1: %globaltimer counters (cross-rank waits; start / publication of the last
 *                    reduction, armed by "prof_stamp_reset"); read "prof_counter_<i>"
 *  "stream_max_blocks" grid cap of streaming/fused kernels (default 0 = one tile per CTA)
 *  "vec_width"       doubles per load: 4 (256-bit LDG/STG), 2 (128-bit), 1; 0 = auto
 *  "unroll"          independent loads in flight per operand: 4, 2, 1; 0 = auto
 *  "exact_threshold" n at or below which reductions sum strictly left-to-right
 *  "count_launches"  1 = count kernel launches (b200vec_ctx_launch_count)      */
int b200vec_ctx_set_tuning(b200vec_ctx ctx, const char* key, int64_t value);
int64_t b200vec_ctx_get_tuning(b200vec_ctx ctx, const char* key);
int64_t b200vec_ctx_launch_count(b200vec_ctx ctx);
const char* b200vec_last_error(void);
const char* b200vec_version(void);

/* ------------------------------------------------------------------------
 * memory (replaces SUNMemoryHelper_Cuda alloc/copy, src/sunmemory/cuda/
 * sundials_cuda_memory.cu:131-351, for this vector)
 * ---------------------------------------------------------------------- */
int b200vec_malloc_device(b200vec_ctx ctx, size_t bytes, void** ptr);  /* cached cudaMalloc */
int b200vec_free_device(b200vec_ctx ctx, void* ptr, size_t bytes);
int b200vec_malloc_host(b200vec_ctx ctx, size_t bytes, void** ptr);    /* pinned, device-mapped */
int b200vec_free_host(b200vec_ctx ctx, void* ptr);
int b200vec_malloc_managed(b200vec_ctx ctx, size_t bytes, void** ptr); /* cudaMallocManaged */
int b200vec_free_managed(b200vec_ctx ctx, void* ptr);
/* stream-ordered copies; sync != 0 waits for completion (cuda:553-599) */
int b200vec_copy_h2d(b200vec_ctx ctx, void* dst_dev, const void* src_host, size_t bytes, int sync);
int b200vec_copy_d2h(b200vec_ctx ctx, void* dst_host, const void* src_dev, size_t bytes, int sync);
int b200vec_copy_d2d(b200vec_ctx ctx, void* dst_dev, const void* src_dev, size_t bytes);
/* H2D on the context's own COPY stream (created on first use, non-blocking): starts once the work
 * already issued on the ctx stream has finished (so a buffer still being read is not overwritten)
 * and overlaps everything issued afterwards.  b200vec_copy_join makes the ctx stream wait for all
 * copies issued so far -- an event wait on the device, the host does not block. */
int b200vec_copy_h2d_async(b200vec_ctx ctx, void* dst_dev, const void* src_host, size_t bytes);
int b200vec_copy_join(b200vec_ctx ctx);

/* ------------------------------------------------------------------------
 * streaming ops (asynchronous on the ctx stream)
 * ---------------------------------------------------------------------- */
/* z = a x + b y with the reference's 12 algebraic forms selected from (a, b, z==x,
 * z==y): replaces N_VLinearSum  serial:387-480 / cuda:771 (linearSumKernel) */
int b200vec_linear_sum(b200vec_ctx ctx, double a, const double* x, double b, const double* y,
                       double* z, int64_t n);
int b200vec_const(b200vec_ctx ctx, double c, double* z, int64_t n);                 /* N_VConst serial:482 cuda:755 */
int b200vec_prod(b200vec_ctx ctx, const double* x, const double* y, double* z, int64_t n); /* N_VProd serial:497 cuda:790 */
int b200vec_div(b200vec_ctx ctx, const double* x, const double* y, double* z, int64_t n);  /* N_VDiv serial:514 cuda:808 */
int b200vec_scale(b200vec_ctx ctx, double c, const double* x, double* z, int64_t n);       /* N_VScale serial:531 cuda:826 */
int b200vec_abs(b200vec_ctx ctx, const double* x, double* z, int64_t n);                   /* N_VAbs serial:557 cuda:843 */
int b200vec_inv(b200vec_ctx ctx, const double* x, double* z, int64_t n);                   /* N_VInv serial:573 cuda:860 */
int b200vec_add_const(b200vec_ctx ctx, const double* x, double b, double* z, int64_t n);   /* N_VAddConst serial:589 cuda:877 */
int b200vec_compare(b200vec_ctx ctx, double c, const double* x, double* z, int64_t n);     /* N_VCompare serial:762 cuda:1197 */

/* ------------------------------------------------------------------------
 * local reductions.  Each launches ONE two-stage kernel (per-thread sequential
 * partial -> warp shuffle -> block -> fixed-order final pass by CTA 0 over tagged
 * CTA partials) that leaves the LOCAL result in the context's device result buffer
 * (b200vec_result_device) AND, as two tagged 8-byte words, in pinned host memory.
 *   result_host != NULL : the call waits for the tagged words and stores the value
 *                         (what N_VDotProd etc. need: no stream sync, zero memcpys);
 *   result_host == NULL : asynchronous; combine across ranks with
 *                         b200vec_allreduce and read with b200vec_result_fetch.
 * Replaces the H2D-init + kernel + D2H + sync pattern of cuda:894-940,
 * 2277-2411 and the atomic reducers of src/sundials/sundials_cuda_kernels.cuh:297-424.
 * ---------------------------------------------------------------------- */
int b200vec_dot_prod(b200vec_ctx ctx, const double* x, const double* y, int64_t n, double* result_host);      /* N_VDotProd[Local] serial:605 cuda:894 */
int b200vec_max_norm(b200vec_ctx ctx, const double* x, int64_t n, double* result_host);                       /* N_VMaxNorm[Local] serial:622 cuda:942 */
int b200vec_min(b200vec_ctx ctx, const double* x, int64_t n, double* result_host);                            /* N_VMin[Local] serial:705 cuda:1099 */
int b200vec_l1_norm(b200vec_ctx ctx, const double* x, int64_t n, double* result_host);                        /* N_VL1Norm[Local] serial:746 cuda:1151 */
/* sum_i (x_i w_i)^2 : N_VWSqrSumLocal serial:650 cuda:989; callers form
 * WrmsNorm = sqrt(sum/N_global) (serial:646) and WL2Norm = sqrt(sum) (serial:743) */
int b200vec_wsqr_sum(b200vec_ctx ctx, const double* x, const double* w, int64_t n, double* result_host);
/* same restricted to id_i > 0 : N_VWSqrSumMaskLocal serial:680 cuda:1043 */
int b200vec_wsqr_sum_mask(b200vec_ctx ctx, const double* x, const double* w, const double* id, int64_t n,
                          double* result_host);
/* z_i = 1/x_i where x_i != 0 (z untouched elsewhere); result 1.0 iff no zero:
 * N_VInvTest[Local] serial:778 cuda:1214 */
int b200vec_inv_test(b200vec_ctx ctx, const double* x, double* z, int64_t n, double* result_host);
/* m_i = constraint-violation mask; result 1.0 iff no violation:
 * N_VConstrMask[Local] serial:800 cuda:1262 */
int b200vec_constr_mask(b200vec_ctx ctx, const double* c, const double* x, double* m, int64_t n,
                        double* result_host);
/* min over denom_i != 0 of num_i/denom_i, DBL_MAX if none:
 * N_VMinQuotient[Local] serial:833 cuda:1312 */
int b200vec_min_quotient(b200vec_ctx ctx, const double* num, const double* denom, int64_t n,
                         double* result_host);

/* w_i = 1 / (rtol |y_i| + atol_i) and *min_denominator_host = min_i (rtol |y_i| + atol_i): the error-weight
 * vector of CVODE / ARKODE / IDA in one pass, bit-identical to the op sequence of cvEwtSetSS/SV
 * (src/cvode/cvode.c:4794-4860: N_VAbs, N_VScale, N_VAddConst | N_VLinearSum, [N_VMin], N_VInv) -- what
 * src/cvode/cvode_fused_gpu.cpp:62 fuses for nvector_cuda.  atol_vec == NULL: scalar atol. */
int b200vec_ewt_set(b200vec_ctx ctx, double rtol, double atol, const double* atol_vec, const double* y,
                    double* w, int64_t n, double* min_denominator_host);

/* ---- integrator-level fused streaming kernels: the bodies of libsundials_cvode_fused_b200.so
 * (include/cvode_fused_b200.h).  One launch each, no host wait; results bit-identical to the N_V* op
 * sequences of src/cvode/cvode_fused_stubs.c on nvector_serial (the line ranges below), which the
 * reference's CUDA kernels (src/cvode/cvode_fused_gpu.cpp) fuse for nvector_cuda.  Outputs may alias
 * inputs as the integrators do (M, y, tempv are updated in place). */
/* tempv = rtol |y| + atol_i, weight = 1 / tempv (stubs:38-72).  weight == NULL: tempv only.  atol_vec == NULL:
 * scalar atol.  24 B/elt (32 with a vector atol) instead of 64 (56). */
int b200vec_cv_ewt(b200vec_ctx ctx, double rtol, double atol, const double* atol_vec, const double* y, double* tempv,
                   double* weight, int64_t n);
/* tmp = mm (y - 0.1 a c / ewt), a_i = 1 where |c_i| >= 1.5 (stubs:80-89): 40 B/elt instead of 112 */
int b200vec_cv_constraints(b200vec_ctx ctx, const double* c, const double* ewt, const double* y, const double* mm,
                           double* tmp, int64_t n);
/* res = rl1 zn1 + ycor + ngamma ftemp (stubs:97-104): 32 B/elt instead of 48 */
int b200vec_cv_nls_resid(b200vec_ctx ctx, double rl1, double ngamma, const double* zn1, const double* ycor,
                         const double* ftemp, double* res, int64_t n);
/* ftemp = h fpred - zn1, y = r ftemp + ypred (stubs:112-119): 40 B/elt instead of 48 */
int b200vec_cv_diag_form_y(b200vec_ctx ctx, double h, double r, const double* fpred, const double* zn1,
                           const double* ypred, double* ftemp, double* y, int64_t n);
/* the 10-op construction of M = I - gamma J with the round-off guard (stubs:128-147): 64 B/elt instead of 224 */
int b200vec_cv_diag_build_m(b200vec_ctx ctx, double uround, double h, const double* ftemp, const double* fpred,
                            const double* ewt, double* bit, double* bitcomp, double* y, double* M, int64_t n);
/* M = 1 + r (1/M - 1) (stubs:154-161): 16 B/elt instead of 64 */
int b200vec_cv_diag_update_m(b200vec_ctx ctx, double r, double* M, int64_t n);

/* z <- a x + z (serial's Vaxpy form, serial:1734) and result = sum_i w_i z_i of the UPDATED z, one
 * pass: a modified Gram-Schmidt step -- N_VLinearSum(1, v_k, -h_i, v_i, v_k) + N_VDotProd(v_{i+1}, v_k),
 * src/sundials/sundials_iterative.c:62-67 -- at 32 B/elt instead of 24 + 16.  w may alias x. */
int b200vec_axpy_dot(b200vec_ctx ctx, double a, const double* x, double* z, const double* w, int64_t n,
                     double* result_host);

/* ---- chained Gram-Schmidt sweeps: all kernels of one orthogonalisation column launched back to back,
 * each reading the coefficient(s) it needs from the DEVICE result slots its predecessor wrote, ONE host
 * wait per column (src/sundials/sundials_iterative.c:45-80 and :130-146; see sundials_iterative_b200.h).
 * V / Ydots / Xcomb: HOST arrays of DEVICE pointers. */
/* modified: res = {v_k.v_k, h_0..h_{nproj-1}, ||v_k||^2 after}; v_k <- v_k - sum h_i V_i, every h_i taken on
 * the vector updated so far.  nproj + 1 kernels. */
int b200vec_mgs_sweep(b200vec_ctx ctx, int nproj, double* vk, const double* const* V, int64_t n,
                      double* h_host, double* norms_host /* [2] */);
/* classical: dots_j = x . Ydots_j (j < nvec <= 24);  z <- Xcomb_0 - sum_{j>=1} dots_{j-1} Xcomb_j and
 * *sqnorm_host = z . z.  2 kernels. */
int b200vec_cgs_step(b200vec_ctx ctx, int nvec, const double* x, const double* const* Ydots,
                     const double* const* Xcomb, double* z, int64_t n, double* dots_host, double* sqnorm_host);

/* device-resident result slots of the last reduction(s): slot k of the context */
double* b200vec_result_device(b200vec_ctx ctx);
/* sync the stream and copy `count` slots to host (after b200vec_allreduce) */
int b200vec_result_fetch(b200vec_ctx ctx, int count, double* result_host);

/* ------------------------------------------------------------------------
 * fused ops.  X / Y / Z are HOST arrays of DEVICE pointers.
 * ---------------------------------------------------------------------- */
/* z = sum_j c_j X_j, each X_j read once, register accumulation in j order
 * (bit-identical to the j-ordered passes of serial:871-942); nvec==1 -> scale,
 * nvec==2 -> linear_sum forms.  z may alias X[0] only.  cuda:1368 */
int b200vec_linear_combination(b200vec_ctx ctx, int nvec, const double* c_host, const double* const* X,
                               double* z, int64_t n);
/* Z_j = a_j x + Y_j, x read once; Y_j may alias Z_j.  serial:944-992 cuda:1424 */
int b200vec_scale_add_multi(b200vec_ctx ctx, int nvec, const double* a_host, const double* x,
                            const double* const* Y, double* const* Z, int64_t n);
/* d_j = sum_i x_i Y_j,i ; x read once per group of 24 outputs, full grid.
 * N_VDotProdMulti / N_VDotProdMultiLocal serial:994-1027 cuda:1489.
 * result_host as for the scalar reductions (nvec slots). */
int b200vec_dot_prod_multi(b200vec_ctx ctx, int nvec, const double* x, const double* const* Y, int64_t n,
                           double* result_host);

/* z = sum_j c_j X_j (as b200vec_linear_combination) AND result = sum_i z_i^2 of the values just
 * written, in one kernel and one host round trip: the N_VLinearCombination + N_VDotProd(v[k], v[k])
 * pair of a classical Gram-Schmidt step (src/sundials/sundials_iterative.c:137-152).  8 (nvec + 1)
 * bytes per element instead of 8 (nvec + 1) + 8.  result_host as for the scalar reductions. */
int b200vec_linear_combination_sqnorm(b200vec_ctx ctx, int nvec, const double* c_host, const double* const* X,
                                      double* z, int64_t n, double* result_host);

/* ------------------------------------------------------------------------
 * vector-array ops.  arrays_alias_* tell the launcher whether the reference
 * call had Z == Y / Z == X as N_Vector* ARRAYS (serial:1062,1069 decide the
 * axpy forms on array identity, not on data pointers).
 * ---------------------------------------------------------------------- */
int b200vec_linear_sum_vector_array(b200vec_ctx ctx, int nvec, double a, const double* const* X, double b,
                                    const double* const* Y, double* const* Z, int z_is_x, int z_is_y,
                                    int64_t n);                                  /* serial:1035-1151 cuda:1555 */
int b200vec_scale_vector_array(b200vec_ctx ctx, int nvec, const double* c_host, const double* const* X,
                               double* const* Z, int64_t n);                     /* serial:1153-1199 cuda:1620 */
int b200vec_const_vector_array(b200vec_ctx ctx, int nvec, double c, double* const* Z, int64_t n); /* serial:1201 cuda:1685 */
/* nvec local sums of (x w)^2 (optionally masked by id > 0, id read once per
 * group of 8 vectors): N_VWrmsNorm[Mask]VectorArray serial:1232-1307
 * cuda:1732,1803.  Results are the raw LOCAL sums; the caller forms
 * sqrt(sum/N_global).  id == NULL -> unmasked. */
int b200vec_wsqr_sum_vector_array(b200vec_ctx ctx, int nvec, const double* const* X, const double* const* W,
                                  const double* id, int64_t n, double* result_host);
/* Z[j][i] = a_j X_i + Y[j][i]; Y, Z flattened as P[j*nvec + i].
 * serial:1309-1408 cuda:1874.  y_is_z: the reference call had Y == Z. */
int b200vec_scale_add_multi_vector_array(b200vec_ctx ctx, int nvec, int nsum, const double* a_host,
                                         const double* const* X, const double* const* Y, double* const* Z,
                                         int y_is_z, int64_t n);
/* Z_j = sum_i c_i X[i][j]; X flattened as P[i*nvec + j].  serial:1410-1544
 * cuda:1949.  x0_is_z: the reference call had X[0] == Z (array identity). */
int b200vec_linear_combination_vector_array(b200vec_ctx ctx, int nvec, int nsum, const double* c_host,
                                            const double* const* X, double* const* Z, int x0_is_z,
                                            int64_t n);

/* ------------------------------------------------------------------------
 * multi-GPU: one rank per GPU, contiguous 1-D partition (MPIPlusX pattern,
 * src/nvector/mpiplusx/nvector_mpiplusx.c:30); the only communication is an
 * allreduce of the context's result slots (table in SURVEY.md section 2b,
 * src/nvector/manyvector/nvector_manyvector.c:815-1793).  NCCL is loaded with
 * dlopen("libnccl.so.2") -- no link-time dependency.
 * ---------------------------------------------------------------------- */
/* reduction scope (one-shot): after b200vec_ctx_set_scope(ctx, B200VEC_SCOPE_GLOBAL)
 * the NEXT reduction call on the context returns the result over ALL ranks of the
 * communicator (SUM / MAX / MIN as the op implies, manyvector.c:815,869,1107):
 * over peer memory the exchange happens inside the reduction kernel's last CTA,
 * otherwise as ncclAllReduce after it.  Every rank must make the same sequence of
 * global reductions (SPMD, as with MPI).  The scope falls back to LOCAL after
 * each reduction call. */
#define B200VEC_SCOPE_LOCAL  0
#define B200VEC_SCOPE_GLOBAL 1
int b200vec_ctx_set_scope(b200vec_ctx ctx, int scope);
/* "none" (single rank), "peer-memory" or "nccl" */
const char* b200vec_comm_transport(b200vec_ctx ctx);

#define B200VEC_UNIQUE_ID_BYTES 128
int b200vec_comm_get_unique_id(unsigned char id[B200VEC_UNIQUE_ID_BYTES]);   /* rank 0, then broadcast */
int b200vec_comm_init(b200vec_ctx ctx, const unsigned char id[B200VEC_UNIQUE_ID_BYTES], int rank, int nranks);
int b200vec_comm_finalize(b200vec_ctx ctx);
/* symmetric peer allocation (collective over the communicator, <= 8 ranks):
 * every rank allocates `bytes` of zeroed HBM and maps all peers' allocations
 * through CUDA IPC; ptrs[r] (array of comm_size entries) is rank r's buffer as
 * addressable from this process -- loads/stores to it travel over NVLink.
 * Used by the reductions' mailboxes and by applications for halo exchange. */
int b200vec_comm_peer_alloc(b200vec_ctx ctx, size_t bytes, void** ptrs);
int b200vec_comm_peer_free(b200vec_ctx ctx, void** ptrs);
int b200vec_comm_rank(b200vec_ctx ctx);
int b200vec_comm_size(b200vec_ctx ctx);   /* 1 when no communicator is attached */
/* in-place allreduce of result slots [0,count) on the ctx stream (no-op when size==1) */
int b200vec_allreduce(b200vec_ctx ctx, int count, int op);
/* generic in-place allreduce of a device buffer of doubles / int64 (global length) */
int b200vec_allreduce_buffer(b200vec_ctx ctx, double* buf_dev, int count, int op);
int b200vec_allreduce_i64_host(b200vec_ctx ctx, int64_t* value_host, int op);

#ifdef __cplusplus
}
#endif
#endif /* B200VEC_H */
