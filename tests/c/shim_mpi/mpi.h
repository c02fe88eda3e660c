/* mpi.h -- SINGLE-RANK MPI stand-in (test infrastructure, our own code).
 *
 * This container has no MPI.  The reference's benchmarks/diffusion_2D and its
 * nvector_parallel need <mpi.h>; with this header (every communicator has
 * exactly one rank, collectives copy, point-to-point never happens because a
 * single rank has no neighbours) the UNMODIFIED reference sources compile by
 * path into a np = 1 CPU program: the golden/CPU-baseline for our re-hosted
 * diffusion_2D.  MPI_Comm is `int`, so objects stay ABI-compatible with the
 * non-MPI libsundials_ref.so (SUNComm is int there too).
 */
#ifndef B200_SHIM_MPI_H
#define B200_SHIM_MPI_H

#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef ptrdiff_t MPI_Aint;
typedef struct
{
  int MPI_SOURCE, MPI_TAG, MPI_ERROR;
} MPI_Status;
typedef void(MPI_User_function)(void*, void*, int*, MPI_Datatype*);

#define MPI_SUCCESS    0
#define MPI_COMM_NULL  0
#define MPI_COMM_WORLD 1
#define MPI_ANY_TAG    (-1)
#define MPI_IN_PLACE   ((void*)1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)

/* datatype handle = size in bytes (user types get handles >= 1024) */
#define MPI_CHAR        1
#define MPI_BYTE        1
#define MPI_INT         4
#define MPI_FLOAT       4
#define MPI_DOUBLE      8
#define MPI_LONG        8
#define MPI_INT32_T     4
#define MPI_INT64_T     8
#define MPI_LONG_DOUBLE 16
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3

static int b200_mpi_user_sizes[64];
static int b200_mpi_user_count = 0;
static inline size_t b200_mpi_size(MPI_Datatype t) { return t >= 1024 ? (size_t)b200_mpi_user_sizes[t - 1024] : (size_t)t; }

static inline int MPI_Init(int* argc, char*** argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code); return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int* n) { (void)c; *n = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_dup(MPI_Comm c, MPI_Comm* out) { *out = c; return MPI_SUCCESS; }
static inline int MPI_Comm_free(MPI_Comm* c) { *c = MPI_COMM_NULL; return MPI_SUCCESS; }
#define MPI_IDENT     0
#define MPI_CONGRUENT 1
#define MPI_UNEQUAL   3
static inline int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int* r) { *r = (a == b) ? MPI_IDENT : MPI_CONGRUENT; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }
static inline double MPI_Wtime(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* 1 x 1 Cartesian grid */
static inline int MPI_Dims_create(int n, int nd, int* dims) { (void)n; for (int i = 0; i < nd; i++) if (dims[i] == 0) dims[i] = 1; return MPI_SUCCESS; }
static inline int MPI_Cart_create(MPI_Comm c, int nd, const int* dims, const int* per, int reorder, MPI_Comm* out)
{ (void)nd; (void)dims; (void)per; (void)reorder; *out = c; return MPI_SUCCESS; }
static inline int MPI_Cart_get(MPI_Comm c, int nd, int* dims, int* per, int* coords)
{ (void)c; for (int i = 0; i < nd; i++) { dims[i] = 1; per[i] = 0; coords[i] = 0; } return MPI_SUCCESS; }
static inline int MPI_Cart_rank(MPI_Comm c, const int* coords, int* r) { (void)c; (void)coords; *r = 0; return MPI_SUCCESS; }

/* collectives over one rank: copy (or nothing when in place) */
static inline int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{ (void)op; (void)c; if (s != MPI_IN_PLACE && s != r) memcpy(r, s, (size_t)n * b200_mpi_size(t)); return MPI_SUCCESS; }
static inline int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{ (void)root; return MPI_Allreduce(s, r, n, t, op, c); }
static inline int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void)b; (void)n; (void)t; (void)root; (void)c; return MPI_SUCCESS; }

/* point-to-point: the only possible peer of the single rank is itself (periodic
   Cartesian grids, benchmarks/advection_reaction_3D).  Messages are matched by tag: an
   Irecv posts its buffer, an Isend copies into the matching posted buffer -- or is kept
   until the matching Irecv arrives.  Anything else (a peer other than rank 0) is a bug. */
#define MPI_PROC_NULL    (-2)
#define MPI_REQUEST_NULL 0
typedef struct { void* buf; const void* sbuf; size_t bytes; int tag; int used; } b200_mpi_msg;
static b200_mpi_msg b200_mpi_recvq[64], b200_mpi_sendq[64];
static inline int MPI_Irecv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* q)
{
  (void)c;
  if (src != 0) abort();
  size_t bytes = (size_t)n * b200_mpi_size(t);
  *q = 1;
  for (int i = 0; i < 64; i++)
    if (b200_mpi_sendq[i].used && b200_mpi_sendq[i].tag == tag)
    { if (b200_mpi_sendq[i].bytes != bytes) abort(); memcpy(b, b200_mpi_sendq[i].sbuf, bytes); b200_mpi_sendq[i].used = 0; return MPI_SUCCESS; }
  for (int i = 0; i < 64; i++)
    if (!b200_mpi_recvq[i].used)
    { b200_mpi_recvq[i].buf = b; b200_mpi_recvq[i].bytes = bytes; b200_mpi_recvq[i].tag = tag; b200_mpi_recvq[i].used = 1; return MPI_SUCCESS; }
  abort();
  return 1;
}
static inline int MPI_Isend(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request* q)
{
  (void)c;
  if (dst != 0) abort();
  size_t bytes = (size_t)n * b200_mpi_size(t);
  *q = 1;
  for (int i = 0; i < 64; i++)
    if (b200_mpi_recvq[i].used && b200_mpi_recvq[i].tag == tag)
    { if (b200_mpi_recvq[i].bytes != bytes) abort(); memcpy(b200_mpi_recvq[i].buf, b, bytes); b200_mpi_recvq[i].used = 0; return MPI_SUCCESS; }
  for (int i = 0; i < 64; i++)
    if (!b200_mpi_sendq[i].used)
    { b200_mpi_sendq[i].sbuf = b; b200_mpi_sendq[i].bytes = bytes; b200_mpi_sendq[i].tag = tag; b200_mpi_sendq[i].used = 1; return MPI_SUCCESS; }
  abort();
  return 1;
}
static inline int MPI_Wait(MPI_Request* q, MPI_Status* s) { (void)q; (void)s; return MPI_SUCCESS; }
static inline int MPI_Waitall(int n, MPI_Request* q, MPI_Status* s) { (void)n; (void)q; (void)s; return MPI_SUCCESS; }
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)

/* derived types / user ops (used by the profiler only) */
static inline int MPI_Type_create_struct(int n, const int* bl, const MPI_Aint* d, const MPI_Datatype* ty, MPI_Datatype* out)
{
  size_t sz = 0;
  for (int i = 0; i < n; i++) { size_t e = (size_t)d[i] + (size_t)bl[i] * b200_mpi_size(ty[i]); if (e > sz) sz = e; }
  b200_mpi_user_sizes[b200_mpi_user_count] = (int)sz;
  *out = 1024 + b200_mpi_user_count++;
  return MPI_SUCCESS;
}
static inline int MPI_Type_get_extent(MPI_Datatype t, MPI_Aint* lb, MPI_Aint* ext) { *lb = 0; *ext = (MPI_Aint)b200_mpi_size(t); return MPI_SUCCESS; }
static inline int MPI_Type_create_resized(MPI_Datatype t, MPI_Aint lb, MPI_Aint ext, MPI_Datatype* out)
{ (void)lb; (void)t; b200_mpi_user_sizes[b200_mpi_user_count] = (int)ext; *out = 1024 + b200_mpi_user_count++; return MPI_SUCCESS; }
static inline int MPI_Type_commit(MPI_Datatype* t) { (void)t; return MPI_SUCCESS; }
static inline int MPI_Type_free(MPI_Datatype* t) { (void)t; return MPI_SUCCESS; }
static inline int MPI_Op_create(MPI_User_function* f, int commute, MPI_Op* op) { (void)f; (void)commute; *op = 100; return MPI_SUCCESS; }
static inline int MPI_Op_free(MPI_Op* op) { (void)op; return MPI_SUCCESS; }

#ifdef __cplusplus
}
#endif
#endif
