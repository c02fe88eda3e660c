/* ar3d_b200.cu -- re-host of the reference's benchmarks/advection_reaction_3D/raja on
 * NVECTOR_B200, without MPI and without RAJA.
 *
 * What the reference does per advection evaluation (rhs3D.hpp:30-319):
 *   3 pack kernels (FillSendBuffers, advection_reaction_3D.cpp:178-262) -> 6 x MPI_Irecv/Isend
 *   -> N_VConst(0) -> interior kernel -> MPI_Waitall -> 3 face kernels, then (DIRK / ERK / BDF)
 *   a separate reaction kernel that reads y and read-modify-writes ydot (rhs3D.hpp:322-383):
 *   ~56 bytes of HBM traffic per unknown and 9 launches.
 * Here it is ONE kernel, 16 bytes per unknown (y read once, ydot written once):
 *   - 1-D slabs in x: the local block is nxl contiguous planes of ny*nz*3 doubles; the y and z
 *     faces wrap periodically inside the rank (the reference's self-exchange when npy = npz = 1);
 *   - k_ar3d_march (c > 0, nz % 4 == 0): a thread owns 4 consecutive nodes of a z-line (12
 *     doubles, three 256-bit loads) and marches through the x-planes of its chunk keeping the
 *     upstream plane in registers; the j-1 line comes through L1/L2 (it is the centre line of
 *     a neighbouring thread), the k-1 node by warp shuffle; the CTA's elected thread prefetches
 *     the tile two planes ahead into L2 with cp.async.bulk.prefetch;
 *   - upwind halo (one face): the CTAs that own the slab's LAST plane are scheduled first and
 *     PUSH that plane into the east neighbour's buffer over NVLink peer memory, then bump its
 *     arrival counter; the CTAs that own plane 0 are scheduled last, march their other planes
 *     first and only then wait for the west neighbour's plane.  An acknowledge counter keeps a
 *     producer from overwriting a buffer its consumer has not read yet (two buffers, by call
 *     parity), so no reduction between two evaluations is required for safety;
 *   - the reaction term is added in registers in the same kernel (AdvectionReaction), or runs
 *     alone for the implicit part of IMEX-ARK.
 *   - k_ar3d_generic: one node per thread, any mesh size and either upwind direction; same
 *     exchange protocol on a grid that is resident in one wave.
 * Arithmetic follows the reference's expressions and operation order (-fmad=false), including
 * its different summation order on the faces (rhs3D.hpp:82-87 vs :182-187), so results are
 * bit-identical to the CPU build of the reference for the same decomposition.
 * The initial condition is separable: per-axis Gaussian factor tables are evaluated with the
 * HOST libm (advection_reaction_3D.cpp:539-557) and summed on the device in the reference's order.
 */
#include <arkode/arkode_arkstep.h>
#include <arkode/arkode_erkstep.h>
#include <cuda_runtime.h>
#include <cvode/cvode.h>
#include <sunlinsol/sunlinsol_spgmr.h>
#include <sunnonlinsol/sunnonlinsol_fixedpoint.h>
#include <sunnonlinsol/sunnonlinsol_newton.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ar3d_b200.h"
#include "nvector_b200.h"

namespace {

constexpr int kT    = 128; /* threads per CTA of the marching kernel: 128 x 12 doubles = 12 KB per plane tile */
constexpr int kTG   = 256; /* generic / node-wise kernels */
constexpr int kAhead = 2;  /* planes of L2 prefetch distance */

struct Rates
{
  double A, B, k1, k2, k3, k4, k5, k6;
};

struct ArArgs
{
  const double* y;
  double* f;
  int nxl, ny, nz;
  int dir;           /* +1: c > 0 (backward differences), -1: c < 0 */
  double cx, cy, cz; /* -c/dx, -c/dy, -c/dz (rhs3D.hpp:43-45) */
  Rates r;
  /* plane supplying the out-of-slab x-neighbour of the halo-dependent plane: the receive
     buffer (np > 1) or the slab's own opposite plane (np == 1, periodic) */
  const double* halo;
  int np;
  int chunk; /* planes per CTA (marching kernel) */
  int npush; /* CTAs that push / consume the halo (generic kernel) */
  /* exchange over peer memory */
  double* send;                          /* downstream neighbour's receive buffer, this call's parity */
  unsigned long long* ctr_remote;        /* its arrival counter                                        */
  const unsigned long long* ctr_local;   /* my arrival counter                                         */
  unsigned long long expected;           /* cumulative arrivals that complete this call                */
  unsigned long long* ack_remote;        /* upstream neighbour's acknowledge counter                   */
  const unsigned long long* ack_local;   /* my acknowledge counter (bumped by my downstream neighbour) */
  unsigned long long ack_expected;       /* acknowledges that free this call's send buffer             */
  unsigned long long* prof;              /* NULL, or the context's profile counters: [2],[3] ns / CTAs waiting for
                                            the halo plane, [4],[5] ns / CTAs waiting for acknowledges           */
};

__device__ __forceinline__ void pdl_prologue()
{
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ void ld4(const double* p, double* v)
{
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
}
/* streaming: no reuse inside the kernel */
__device__ __forceinline__ void ld4s(const double* p, double* v)
{
  asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
               : "l"(p)
               : "memory");
}
/* data written by another GPU: read at L2 */
__device__ __forceinline__ void ld4cg(const double* p, double* v)
{
  asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ double ld1cg(const double* p)
{
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st4(double* p, const double* v)
{
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
__device__ __forceinline__ void ld12(const double* p, double (&v)[12])
{
  ld4(p, v);
  ld4(p + 4, v + 4);
  ld4(p + 8, v + 8);
}
__device__ __forceinline__ void ld12s(const double* p, double (&v)[12])
{
  ld4s(p, v);
  ld4s(p + 4, v + 4);
  ld4s(p + 8, v + 8);
}
__device__ __forceinline__ void ld12cg(const double* p, double (&v)[12])
{
  ld4cg(p, v);
  ld4cg(p + 4, v + 4);
  ld4cg(p + 8, v + 8);
}
__device__ __forceinline__ void st12(double* p, const double (&v)[12])
{
  st4(p, v);
  st4(p + 4, v + 4);
  st4(p + 8, v + 8);
}

/* prof (optional, the context's profile counters): [0] += ns this CTA spun, [1] += 1 */
__device__ __forceinline__ void wait_counter(const unsigned long long* ctr, unsigned long long expected,
                                             unsigned long long* prof = nullptr)
{
  if (threadIdx.x == 0)
  {
    unsigned long long v, t0 = 0, p0 = 0;
    unsigned int spins = 0;
    if (prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(p0));
    for (;;)
    {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
      if (v >= expected) break;
      if ((++spins & 0xfffu) == 0)
      {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 30000000000ull) asm volatile("trap;"); /* dead neighbour: fail loudly, do not hang */
      }
    }
    if (prof)
    {
      unsigned long long p1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(p1));
      atomicAdd(prof, p1 - p0);
      atomicAdd(prof + 1, 1ull);
    }
  }
  __syncthreads();
}

/* all threads of the CTA have finished their peer stores / halo loads -> one system-scope bump */
__device__ __forceinline__ void signal_counter(unsigned long long* ctr)
{
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd_system(ctr, 1ull);
}

/* upwind difference of one unknown, c > 0.  Interior points sum z, y, x (rhs3D.hpp:82-87); points
   of the slab's west / south / back faces sum x, y, z (rhs3D.hpp:182-187, 204-209, 226-231). */
__device__ __forceinline__ double upwind(const ArArgs& a, bool face, double y, double yi, double yj, double yk)
{
  double d;
  if (face)
  {
    d = a.cx * (y - yi);
    d += a.cy * (y - yj);
    d += a.cz * (y - yk);
  }
  else
  {
    d = a.cz * (y - yk);
    d += a.cy * (y - yj);
    d += a.cx * (y - yi);
  }
  return d;
}
/* c < 0: forward differences (rhs3D.hpp:120-125 interior, :264-269 faces) */
__device__ __forceinline__ double downwind(const ArArgs& a, bool face, double y, double yi, double yj, double yk)
{
  double d;
  if (face)
  {
    d = a.cx * (yi - y);
    d += a.cy * (yj - y);
    d += a.cz * (yk - y);
  }
  else
  {
    d = a.cz * (yk - y);
    d += a.cy * (yj - y);
    d += a.cx * (yi - y);
  }
  return d;
}

/* reaction terms g(u,v,w) added to d (rhs3D.hpp:373-378) */
__device__ __forceinline__ void react_add(const Rates& r, double u, double v, double w, double& du, double& dv,
                                          double& dw)
{
  du += r.k1 * r.A - r.k2 * w * u + r.k3 * u * u * v - r.k4 * u;
  dv += r.k2 * w * u - r.k3 * u * u * v;
  dw += -r.k2 * w * u + r.k5 * r.B - r.k6 * w;
}

/* ------------------------------------------------------------------ marching kernel
 * grid = (tiles of the plane, chunks of planes); thread = 4 nodes of one z-line. */
template <bool REACT>
__device__ __forceinline__ void march_plane(const ArArgs& a, int i, int64_t P, int64_t m0, int64_t mj, int64_t me,
                                            bool need_edge, bool jface, bool kface0, bool act, double (&prev)[12],
                                            const double* plane_prev_override)
{
  const double* yp = a.y + (int64_t)i * P;
  double c[12], jm[12], e[3];
  ld12s(yp + m0, c);
  ld12(yp + mj, jm);
  e[0] = e[1] = e[2] = 0.0;
  if (need_edge)
  {
    e[0] = yp[me];
    e[1] = yp[me + 1];
    e[2] = yp[me + 2];
  }
  if (plane_prev_override) ld12cg(plane_prev_override + m0, prev);
  /* last node of the previous thread's span = k-1 neighbour of my first node */
  double km0[3];
#pragma unroll
  for (int l = 0; l < 3; l++)
  {
    const double s = __shfl_up_sync(0xffffffffu, c[9 + l], 1);
    km0[l]         = need_edge ? e[l] : s;
  }
  const bool iface = (i == 0);
  double d[12];
#pragma unroll
  for (int n = 0; n < 4; n++)
  {
    const bool face = iface || jface || (kface0 && n == 0);
#pragma unroll
    for (int l = 0; l < 3; l++)
    {
      const double yk = (n == 0) ? km0[l] : c[3 * (n - 1) + l];
      d[3 * n + l]    = upwind(a, face, c[3 * n + l], prev[3 * n + l], jm[3 * n + l], yk);
    }
    if (REACT) react_add(a.r, c[3 * n], c[3 * n + 1], c[3 * n + 2], d[3 * n], d[3 * n + 1], d[3 * n + 2]);
  }
  if (act) st12(a.f + (int64_t)i * P + m0, d);
#pragma unroll
  for (int q = 0; q < 12; q++) prev[q] = c[q];
}

__device__ __forceinline__ void prefetch_tile(const double* p, int64_t doubles_left)
{
  /* one elected thread: the CTA's 12 KB tile of a later plane -> L2 */
  if (threadIdx.x == 0 && doubles_left > 0)
  {
    const int64_t n   = doubles_left < (int64_t)kT * 12 ? doubles_left : (int64_t)kT * 12;
    const uint32_t by = (uint32_t)(n * 8);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(by) : "memory");
  }
}

template <bool REACT>
__global__ void __launch_bounds__(kT, 3) k_ar3d_march(const __grid_constant__ ArArgs a)
{
  pdl_prologue();
  const int nq_row = a.nz >> 2;
  const int64_t nq = (int64_t)a.ny * nq_row; /* 4-node spans per plane */
  const int64_t P  = nq * 12;
  const int64_t q0 = (int64_t)blockIdx.x * kT + threadIdx.x;
  const bool act   = q0 < nq;
  const int64_t q  = act ? q0 : nq - 1; /* idle threads shadow a valid span (no stores) */
  const int lane   = threadIdx.x & 31;
  const int j      = (int)(q / nq_row);
  const int kq     = (int)(q - (int64_t)j * nq_row);
  const int64_t R  = (int64_t)a.nz * 3;
  const int64_t m0 = q * 12;
  const int64_t mj = (j > 0) ? m0 - R : m0 + (int64_t)(a.ny - 1) * R; /* periodic south (Srecv) */
  const bool kface0    = (kq == 0);
  const bool need_edge = kface0 || lane == 0; /* the others take it from lane-1 by shuffle */
  const int64_t me     = kface0 ? m0 + R - 3 : m0 - 3; /* periodic back (Brecv) */
  const bool jface     = (j == 0);
  const int64_t tile0  = (int64_t)blockIdx.x * kT * 12;

  /* chunk order: the chunk with the slab's last plane first (its push leaves at kernel
     start), the chunk with plane 0 last (it waits for the neighbour) */
  const int nch = gridDim.y;
  int ci        = blockIdx.y;
  if (nch > 1) ci = (blockIdx.y == 0) ? nch - 1 : (blockIdx.y == (unsigned)nch - 1) ? 0 : blockIdx.y;
  const int ib = ci * a.chunk;
  const int ie = (ib + a.chunk < a.nxl) ? ib + a.chunk : a.nxl;
  const bool xch = a.np > 1;

  if (xch && ie == a.nxl)
  {
    if (a.ack_expected) wait_counter(a.ack_local, a.ack_expected, a.prof ? a.prof + 4 : nullptr);
    if (act)
    {
      double v[12];
      ld12(a.y + (int64_t)(a.nxl - 1) * P + m0, v);
      st12(a.send + m0, v);
    }
    signal_counter(a.ctr_remote);
  }

  double prev[12];
  if (xch && ib == 0)
  {
    /* planes 1 .. ie-1 first (upstream plane 0 is local), plane 0 after the halo has arrived */
    if (ie > 1)
    {
      ld12(a.y + m0, prev);
      for (int i = 1; i < ie; i++)
      {
        if (i + kAhead < ie) prefetch_tile(a.y + (int64_t)(i + kAhead) * P + tile0, P - tile0);
        march_plane<REACT>(a, i, P, m0, mj, me, need_edge, jface, kface0, act, prev, nullptr);
      }
    }
    wait_counter(a.ctr_local, a.expected, a.prof ? a.prof + 2 : nullptr);
    march_plane<REACT>(a, 0, P, m0, mj, me, need_edge, jface, kface0, act, prev, a.halo);
    signal_counter(a.ack_remote);
  }
  else
  {
    if (ib == 0) ld12(a.halo + m0, prev);
    else ld12(a.y + (int64_t)(ib - 1) * P + m0, prev);
    for (int i = ib; i < ie; i++)
    {
      if (i + kAhead < ie) prefetch_tile(a.y + (int64_t)(i + kAhead) * P + tile0, P - tile0);
      march_plane<REACT>(a, i, P, m0, mj, me, need_edge, jface, kface0, act, prev, nullptr);
    }
  }
}

/* ------------------------------------------------------------------ generic kernel
 * one node per thread, persistent grid (<= one resident wave), either direction */
template <bool REACT>
__device__ __forceinline__ void generic_node(const ArArgs& a, int64_t node, int64_t npl, bool halo_remote)
{
  const int i   = (int)(node / npl);
  int64_t rem   = node - (int64_t)i * npl;
  const int j   = (int)(rem / a.nz);
  const int k   = (int)(rem - (int64_t)j * a.nz);
  const int64_t P = npl * 3;
  const double* yc = a.y + node * 3;
  const double u = yc[0], v = yc[1], w = yc[2];
  double d[3];
  if (a.dir > 0)
  {
    const bool face    = (i == 0) || (j == 0) || (k == 0);
    const int jn       = (j > 0) ? j - 1 : a.ny - 1;
    const int kn       = (k > 0) ? k - 1 : a.nz - 1;
    const double* yj   = a.y + (((int64_t)i * a.ny + jn) * a.nz + k) * 3;
    const double* yk   = a.y + (((int64_t)i * a.ny + j) * a.nz + kn) * 3;
    const int64_t off  = ((int64_t)j * a.nz + k) * 3;
    double yi[3];
    if (i > 0)
    {
      const double* p = a.y + (int64_t)(i - 1) * P + off;
      yi[0] = p[0]; yi[1] = p[1]; yi[2] = p[2];
    }
    else if (halo_remote) { yi[0] = ld1cg(a.halo + off); yi[1] = ld1cg(a.halo + off + 1); yi[2] = ld1cg(a.halo + off + 2); }
    else { yi[0] = a.halo[off]; yi[1] = a.halo[off + 1]; yi[2] = a.halo[off + 2]; }
    d[0] = upwind(a, face, u, yi[0], yj[0], yk[0]);
    d[1] = upwind(a, face, v, yi[1], yj[1], yk[1]);
    d[2] = upwind(a, face, w, yi[2], yj[2], yk[2]);
  }
  else
  {
    const bool face    = (i == a.nxl - 1) || (j == a.ny - 1) || (k == a.nz - 1);
    const int jn       = (j < a.ny - 1) ? j + 1 : 0;
    const int kn       = (k < a.nz - 1) ? k + 1 : 0;
    const double* yj   = a.y + (((int64_t)i * a.ny + jn) * a.nz + k) * 3;
    const double* yk   = a.y + (((int64_t)i * a.ny + j) * a.nz + kn) * 3;
    const int64_t off  = ((int64_t)j * a.nz + k) * 3;
    double yi[3];
    if (i < a.nxl - 1)
    {
      const double* p = a.y + (int64_t)(i + 1) * P + off;
      yi[0] = p[0]; yi[1] = p[1]; yi[2] = p[2];
    }
    else if (halo_remote) { yi[0] = ld1cg(a.halo + off); yi[1] = ld1cg(a.halo + off + 1); yi[2] = ld1cg(a.halo + off + 2); }
    else { yi[0] = a.halo[off]; yi[1] = a.halo[off + 1]; yi[2] = a.halo[off + 2]; }
    d[0] = downwind(a, face, u, yi[0], yj[0], yk[0]);
    d[1] = downwind(a, face, v, yi[1], yj[1], yk[1]);
    d[2] = downwind(a, face, w, yi[2], yj[2], yk[2]);
  }
  if (REACT) react_add(a.r, u, v, w, d[0], d[1], d[2]);
  double* fo = a.f + node * 3;
  fo[0] = d[0];
  fo[1] = d[1];
  fo[2] = d[2];
}

template <bool REACT>
__global__ void __launch_bounds__(kTG) k_ar3d_generic(const __grid_constant__ ArArgs a)
{
  pdl_prologue();
  const int64_t npl   = (int64_t)a.ny * a.nz; /* nodes per plane */
  const int64_t total = npl * a.nxl;
  const bool xch      = a.np > 1;
  const int send_pl   = (a.dir > 0) ? a.nxl - 1 : 0; /* plane pushed downstream        */
  const int halo_pl   = (a.dir > 0) ? 0 : a.nxl - 1; /* plane that needs the neighbour */
  const bool pusher   = xch && (int)blockIdx.x < a.npush;

  if (pusher)
  {
    if (a.ack_expected) wait_counter(a.ack_local, a.ack_expected);
    const double* src = a.y + (int64_t)send_pl * npl * 3;
    for (int64_t m = (int64_t)blockIdx.x * kTG + threadIdx.x; m < npl * 3; m += (int64_t)a.npush * kTG)
      a.send[m] = src[m];
    signal_counter(a.ctr_remote);
  }
  for (int64_t node = (int64_t)blockIdx.x * kTG + threadIdx.x; node < total; node += (int64_t)gridDim.x * kTG)
  {
    if (xch && (int)(node / npl) == halo_pl) continue;
    generic_node<REACT>(a, node, npl, false);
  }
  if (pusher)
  {
    wait_counter(a.ctr_local, a.expected);
    for (int64_t n = (int64_t)blockIdx.x * kTG + threadIdx.x; n < npl; n += (int64_t)a.npush * kTG)
      generic_node<REACT>(a, (int64_t)halo_pl * npl + n, npl, true);
    signal_counter(a.ack_remote);
  }
}

/* ------------------------------------------------------------------ node-wise kernels
 * VEC: a thread handles 4 nodes (12 doubles, 256-bit accesses); else one node. */
template <bool VEC>
__global__ void __launch_bounds__(kTG) k_ar3d_reaction(const double* __restrict__ y, double* __restrict__ f, int64_t nodes, Rates r)
{
  pdl_prologue();
  constexpr int NPT = VEC ? 4 : 1;
  for (int64_t t = (int64_t)blockIdx.x * kTG + threadIdx.x; t * NPT < nodes; t += (int64_t)gridDim.x * kTG)
  {
    double c[3 * NPT], d[3 * NPT];
    if constexpr (VEC)
    {
      ld4s(y + t * 12, c);
      ld4s(y + t * 12 + 4, c + 4);
      ld4s(y + t * 12 + 8, c + 8);
    }
    else { c[0] = y[t * 3]; c[1] = y[t * 3 + 1]; c[2] = y[t * 3 + 2]; }
#pragma unroll
    for (int n = 0; n < NPT; n++)
    {
      /* N_VConst(0, ydot) then "+=" (rhs3D.hpp:344, 373) */
      d[3 * n] = d[3 * n + 1] = d[3 * n + 2] = 0.0;
      react_add(r, c[3 * n], c[3 * n + 1], c[3 * n + 2], d[3 * n], d[3 * n + 1], d[3 * n + 2]);
    }
    if constexpr (VEC)
    {
      st4(f + t * 12, d);
      st4(f + t * 12 + 4, d + 4);
      st4(f + t * 12 + 8, d + 8);
    }
    else { f[t * 3] = d[0]; f[t * 3 + 1] = d[1]; f[t * 3 + 2] = d[2]; }
  }
}

/* x = (I - gamma J)^-1 b per node, J = dg/dy (rhs3D.hpp:441-550).  The closed-form 3x3 solve is
   evaluated with the reference's intermediate products and operation order. */
__device__ __forceinline__ void block_solve(const Rates& r, double gamma, double u, double v, double w, double b0,
                                            double b1, double b2, double& x0, double& x1, double& x2)
{
  /* J rows (u, v, w columns) */
  double J0 = -r.k2 * w + 2.0 * r.k3 * u * v - r.k4;
  double J1 = r.k3 * u * u;
  double J2 = -r.k2 * u;
  double J3 = r.k2 * w - 2.0 * r.k3 * u * v;
  double J4 = -r.k3 * u * u;
  double J5 = r.k2 * u;
  double J6 = -r.k2 * w;
  double J7 = 0.0;
  double J8 = -r.k2 * u - r.k6;
  /* M = I - gamma J */
  J0 = 1. - (gamma * J0);
  J1 = -gamma * J1;
  J2 = -gamma * J2;
  J3 = -gamma * J3;
  J4 = 1. - (gamma * J4);
  J5 = -gamma * J5;
  J6 = -gamma * J6;
  J7 = -gamma * J7;
  J8 = 1. - (gamma * J8);
  /* adjugate / determinant for rows 0 and 1, elimination for row 2 */
  const double p48   = J4 * J8;
  const double p15   = J1 * J5;
  const double p27   = J2 * J7;
  const double p57   = J5 * J7;
  const double p18   = J1 * J8;
  const double p24   = J2 * J4;
  const double idet  = 1.0 / (J0 * p48 - J0 * p57 + J3 * p27 - J3 * p18 + J6 * p15 - J6 * p24);
  const double p23   = J2 * J3;
  const double p6b   = J6 * b0;
  const double p26   = J2 * J6;
  const double p3b   = J3 * b0;
  const double i0    = 1.0 / J0;
  const double q10   = J1 * i0;
  const double ratio = (-J6 * q10 + J7) / (-J3 * q10 + J4);
  x0 = idet * (b0 * (p48 - p57) + b1 * (p27 - p18) + b2 * (p15 - p24));
  x1 = idet * (b2 * (p23 - J0 * J5) + b1 * (J0 * J8 - p26) + J5 * p6b - J8 * p3b);
  x2 = (-b2 + i0 * p6b + ratio * (b1 - p3b * i0)) / (-J8 + i0 * p26 + ratio * (J5 - i0 * p23));
}

template <bool VEC>
__global__ void __launch_bounds__(kTG) k_ar3d_psolve(const double* __restrict__ y, const double* b, double* x, int64_t nodes, Rates r,
                                                     double gamma)
{
  pdl_prologue();
  constexpr int NPT = VEC ? 4 : 1;
  for (int64_t t = (int64_t)blockIdx.x * kTG + threadIdx.x; t * NPT < nodes; t += (int64_t)gridDim.x * kTG)
  {
    double c[3 * NPT], bb[3 * NPT], xx[3 * NPT];
    if constexpr (VEC)
    {
      ld4s(y + t * 12, c);
      ld4s(y + t * 12 + 4, c + 4);
      ld4s(y + t * 12 + 8, c + 8);
      ld4s(b + t * 12, bb);
      ld4s(b + t * 12 + 4, bb + 4);
      ld4s(b + t * 12 + 8, bb + 8);
    }
    else
    {
      c[0] = y[t * 3]; c[1] = y[t * 3 + 1]; c[2] = y[t * 3 + 2];
      bb[0] = b[t * 3]; bb[1] = b[t * 3 + 1]; bb[2] = b[t * 3 + 2];
    }
#pragma unroll
    for (int n = 0; n < NPT; n++)
      block_solve(r, gamma, c[3 * n], c[3 * n + 1], c[3 * n + 2], bb[3 * n], bb[3 * n + 1], bb[3 * n + 2], xx[3 * n],
                  xx[3 * n + 1], xx[3 * n + 2]);
    if constexpr (VEC)
    {
      st4(x + t * 12, xx);
      st4(x + t * 12 + 4, xx + 4);
      st4(x + t * 12 + 8, xx + 8);
    }
    else { x[t * 3] = xx[0]; x[t * 3 + 1] = xx[1]; x[t * 3 + 2] = xx[2]; }
  }
}

/* y = steady state + Gaussian bump; gx/gy/gz = per-axis factors from the host libm
   (advection_reaction_3D.cpp:600-611: p = x + y + z) */
__global__ void __launch_bounds__(kTG) k_ar3d_ic(double* y, int nxl, int ny, int nz, const double* gx, const double* gy,
                                                 const double* gz, double us, double vs, double ws)
{
  const int64_t npl = (int64_t)ny * nz, total = npl * nxl;
  for (int64_t node = (int64_t)blockIdx.x * kTG + threadIdx.x; node < total; node += (int64_t)gridDim.x * kTG)
  {
    const int i      = (int)(node / npl);
    const int64_t rm = node - (int64_t)i * npl;
    const int j      = (int)(rm / nz);
    const int k      = (int)(rm - (int64_t)j * nz);
    const double p   = gx[i] + gy[j] + gz[k];
    y[node * 3]      = us + p;
    y[node * 3 + 1]  = vs + p;
    y[node * 3 + 2]  = ws + p;
  }
}

__global__ void __launch_bounds__(kTG) k_ar3d_mask(double* m, int64_t nodes, int component)
{
  for (int64_t node = (int64_t)blockIdx.x * kTG + threadIdx.x; node < nodes; node += (int64_t)gridDim.x * kTG)
  {
    m[node * 3]     = (component == 0) ? 1.0 : 0.0;
    m[node * 3 + 1] = (component == 1) ? 1.0 : 0.0;
    m[node * 3 + 2] = (component == 2) ? 1.0 : 0.0;
  }
}

} // namespace

struct b200_ar3d_plan_s
{
  b200vec_ctx ctx = nullptr;
  b200_ar3d_opts o;
  int rank = 0, np = 1;
  int nx = 0, ny = 0, nz = 0, nxl = 0, is = 0;
  double dx = 0, dy = 0, dz = 0;
  int64_t neq = 0, neq_loc = 0, plane = 0; /* plane = doubles per x-plane */
  Rates r;
  bool fast = false;
  int chunk = 8, tiles = 1, npush = 1, grid_generic = 1, sms = 148;
  double* d_tables = nullptr; /* gx (nxl) | gy (ny) | gz (nz) */
  /* symmetric halo region: [16 u64 counters][2 receive planes] */
  void* peers[8]         = {nullptr};
  bool have_peers        = false;
  unsigned long long seq = 0, arrivals = 0;
  unsigned long long acks_by_parity[2] = {0, 0}; /* cumulative acknowledges after the last use of each buffer */
  unsigned long long acks = 0;
  /* timing */
  bool time_rhs  = false;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  double rhs_ms  = 0;
  long rhs_calls = 0, psolve_calls = 0;
};

namespace {

inline unsigned long long* arrive_ctr(void* base) { return (unsigned long long*)base; }
inline unsigned long long* ack_ctr(void* base) { return (unsigned long long*)base + 8; }
inline double* recv_plane(void* base, int64_t plane, int par) { return (double*)((unsigned long long*)base + 16) + (size_t)par * plane; }

int fail(const char* what)
{
  fprintf(stderr, "[ar3d_b200] ERROR: %s (%s)\n", what, b200vec_last_error());
  return -1;
}

template <class K, class... Args>
cudaError_t launch_pdl(K kernel, dim3 grid, dim3 block, cudaStream_t s, Args... args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim            = grid;
  cfg.blockDim           = block;
  cfg.stream             = s;
  cudaLaunchAttribute at[1];
  at[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs                                        = at;
  cfg.numAttrs                                     = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

int64_t node_grid(b200_ar3d_plan p, int64_t work_items)
{
  int64_t blocks = (work_items + kTG - 1) / kTG;
  const int64_t cap = (int64_t)p->sms * 8;
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : blocks;
}

} // namespace

extern "C" {

void b200_ar3d_default_opts(b200_ar3d_opts* o)
{
  memset(o, 0, sizeof(*o));
  o->npts = 100;
  o->xmax = 1.0;
  o->A    = 1.0;
  o->B    = 3.5;
  o->k1 = o->k2 = o->k3 = o->k4 = 1.0;
  o->k5 = o->k6 = 1.0 / 5.0e-6;
  o->c          = 0.01;
  o->method     = AR3D_METHOD_ARK_DIRK;
  o->nls        = AR3D_NLS_NEWTON;
  o->order      = 3;
  o->fpaccel    = 3;
  o->precond    = 1;
  o->fused      = 0;
  o->t0         = 0.0;
  o->tf         = 10.0;
  o->rtol       = 1.0e-6;
  o->atol       = 1.0e-9;
  o->nout       = 10;
  o->save       = 0;
  strcpy(o->outputdir, ".");
  o->output    = 1;
  o->fused_ewt = 1;
}

int b200_ar3d_plan_create(b200vec_ctx ctx, const b200_ar3d_opts* opts, b200_ar3d_plan* out)
{
  if (!ctx || !opts || !out) return -1;
  if (opts->npts < 1 || opts->npts > 2000000) return fail("npts out of range");
  auto* p = new b200_ar3d_plan_s();
  p->ctx  = ctx;
  p->o    = *opts;
  p->rank = b200vec_comm_rank(ctx);
  p->np   = b200vec_comm_size(ctx);
  p->nx = p->ny = p->nz = (int)opts->npts;
  /* ParallelGrid.hpp:127-157 with dims = {np, 1, 1} */
  p->is  = (int)((int64_t)p->nx * p->rank / p->np);
  p->nxl = (int)((int64_t)p->nx * (p->rank + 1) / p->np) - p->is;
  p->dx  = (opts->xmax - 0.0) / (double)p->nx;
  p->dy  = (opts->xmax - 0.0) / (double)p->ny;
  p->dz  = (opts->xmax - 0.0) / (double)p->nz;
  if (p->nxl < 1)
  {
    delete p;
    return fail("every rank needs at least one x-plane");
  }
  p->plane   = (int64_t)p->ny * p->nz * 3;
  p->neq     = (int64_t)p->nx * p->plane;
  p->neq_loc = (int64_t)p->nxl * p->plane;
  p->r       = Rates{opts->A, opts->B, opts->k1, opts->k2, opts->k3, opts->k4, opts->k5, opts->k6};
  int dev    = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&p->sms, cudaDevAttrMultiProcessorCount, dev);

  p->fast = !opts->force_generic && opts->c > 0.0 && (p->nz % 4 == 0);
  if (const char* v = getenv("B200_AR3D_GENERIC"))
    if (atoi(v)) p->fast = false;
  const int64_t nq = (int64_t)p->ny * (p->nz / 4);
  p->tiles         = (int)((nq + kT - 1) / kT);
  p->chunk         = opts->planes_per_cta > 0 ? opts->planes_per_cta : 8;
  if (const char* v = getenv("B200_AR3D_CHUNK"))
    if (atoi(v) > 0) p->chunk = atoi(v);
  if (p->fast && p->np > 1)
  {
    /* The marching kernel gives the CTAs that push the slab's last plane the lowest block indices and the
       CTAs that wait for the neighbour's plane the highest, but CUDA does not promise dispatch in index
       order.  The exchange cannot dead-lock whatever the order as long as the waiting CTAs (one per plane
       tile) cannot fill every resident CTA slot of the GPU: a free slot always goes to a CTA that never
       waits on a peer.  Keep the fast kernel across ranks only while the waiters leave at least one slot
       per SM free (512^2 planes: 256 waiters, 444 slots); larger planes (ny * nz >~ 3e5 nodes) use the
       generic kernel, whose grid is one resident wave. */
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ar3d_march<true>, kT, 0);
    if (occ < 1) occ = 1;
    if ((int64_t)p->tiles + p->sms > (int64_t)p->sms * occ) p->fast = false;
  }
  {
    /* generic kernel: persistent grid of at most one resident wave (its CTAs wait on peers) */
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ar3d_generic<true>, kTG, 0);
    if (occ < 1) occ = 1;
    int64_t want = ((int64_t)p->nxl * p->ny * p->nz + kTG - 1) / kTG;
    const int64_t cap = (int64_t)p->sms * occ;
    /* pushers: depends on the plane size only, so every rank uses the same count */
    int64_t npush = ((int64_t)p->ny * p->nz + kTG - 1) / kTG;
    if (npush > 32) npush = 32;
    p->npush = (int)npush;
    if (want < npush) want = npush;
    if (want > cap) want = cap;
    p->grid_generic = (int)want;
  }

  /* per-axis Gaussian factors with the HOST libm (Gaussian3D, advection_reaction_3D.cpp:539-557) */
  {
    const double xmax  = opts->xmax;
    const double alpha = 0.1;
    const double mu[]    = {xmax / 2.0, xmax / 2.0, xmax / 2.0};
    const double sigma[] = {xmax / 4.0, xmax / 4.0, xmax / 4.0};
    const double denom   = 2.0 * sqrt((sigma[0] * sigma[1] * sigma[2]) * pow(2 * M_PI, 3));
    std::vector<double> t((size_t)p->nxl + p->ny + p->nz);
    for (int i = 0; i < p->nxl; i++)
    {
      /* x = (xcrd * nxl + i) * dx with the rank's own nxl, as SetIC writes it (:600) */
      const double x = (p->rank * p->nxl + i) * p->dx;
      t[i]           = alpha * exp(-((x - mu[0]) * (x - mu[0]) * (1.0 / sigma[0])) / denom);
    }
    for (int j = 0; j < p->ny; j++)
    {
      const double y = (0 * p->ny + j) * p->dy;
      t[p->nxl + j]  = alpha * exp(-((y - mu[1]) * (y - mu[1]) * (1.0 / sigma[1])) / denom);
    }
    for (int k = 0; k < p->nz; k++)
    {
      const double z        = (0 * p->nz + k) * p->dz;
      t[p->nxl + p->ny + k] = alpha * exp(-((z - mu[2]) * (z - mu[2]) * (1.0 / sigma[2])) / denom);
    }
    void* d = nullptr;
    if (b200vec_malloc_device(ctx, t.size() * sizeof(double), &d)) { delete p; return fail("table allocation"); }
    p->d_tables = (double*)d;
    if (b200vec_copy_h2d(ctx, d, t.data(), t.size() * sizeof(double), 1)) { delete p; return fail("table upload"); }
  }

  if (p->np > 1)
  {
    const size_t bytes = 16 * sizeof(unsigned long long) + (size_t)2 * p->plane * sizeof(double);
    if (b200vec_comm_peer_alloc(ctx, bytes, p->peers)) { delete p; return fail("peer allocation for the halo planes"); }
    p->have_peers = true;
  }
  p->time_rhs = getenv("B200_AR3D_TIME_RHS") != nullptr;
  if (p->time_rhs)
  {
    cudaEventCreate(&p->e0);
    cudaEventCreate(&p->e1);
  }
  *out = p;
  return 0;
}

int64_t b200_ar3d_plan_local_neq(b200_ar3d_plan p) { return p ? p->neq_loc : -1; }
int b200_ar3d_plan_is_fast(b200_ar3d_plan p) { return p && p->fast; }

void b200_ar3d_plan_destroy(b200_ar3d_plan p)
{
  if (!p) return;
  b200vec_ctx_sync(p->ctx);
  if (p->have_peers) b200vec_comm_peer_free(p->ctx, p->peers);
  if (p->d_tables) b200vec_free_device(p->ctx, p->d_tables, ((size_t)p->nxl + p->ny + p->nz) * sizeof(double));
  if (p->e0) cudaEventDestroy(p->e0);
  if (p->e1) cudaEventDestroy(p->e1);
  delete p;
}

int b200_ar3d_set_ic(b200_ar3d_plan p, double* y)
{
  const Rates& r  = p->r;
  const double us = r.k1 * r.A / r.k4;
  const double vs = r.k2 * r.k4 * r.B / (r.k1 * r.k3 * r.A);
  const double ws = 3.0;
  cudaStream_t s  = (cudaStream_t)b200vec_ctx_get_stream(p->ctx);
  k_ar3d_ic<<<(unsigned)node_grid(p, p->neq_loc / 3), kTG, 0, s>>>(y, p->nxl, p->ny, p->nz, p->d_tables,
                                                                   p->d_tables + p->nxl, p->d_tables + p->nxl + p->ny,
                                                                   us, vs, ws);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int b200_ar3d_component_mask(b200_ar3d_plan p, int component, double* m)
{
  cudaStream_t s = (cudaStream_t)b200vec_ctx_get_stream(p->ctx);
  k_ar3d_mask<<<(unsigned)node_grid(p, p->neq_loc / 3), kTG, 0, s>>>(m, p->neq_loc / 3, component);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int b200_ar3d_rhs(b200_ar3d_plan p, int which, const double* y, double* f)
{
  if (!p || !y || !f || y == f) return -1;
  cudaStream_t s = (cudaStream_t)b200vec_ctx_get_stream(p->ctx);
  cudaError_t e  = cudaSuccess;
  if (p->time_rhs) cudaEventRecord(p->e0, s);
  const int64_t nodes = p->neq_loc / 3;
  const bool vec      = (nodes % 4 == 0) && ((uintptr_t)y % 32 == 0) && ((uintptr_t)f % 32 == 0);
  if (which == AR3D_RHS_REACTION || p->o.c == 0.0)
  {
    if (which == AR3D_RHS_ADVECTION)
    {
      /* c == 0: the reference's Advection only zeroes ydot (rhs3D.hpp:59, no branch taken) */
      if (b200vec_const(p->ctx, 0.0, f, p->neq_loc)) return fail("const");
    }
    else if (vec)
      e = launch_pdl(k_ar3d_reaction<true>, dim3((unsigned)node_grid(p, nodes / 4)), dim3(kTG), s, y, f, nodes, p->r);
    else e = launch_pdl(k_ar3d_reaction<false>, dim3((unsigned)node_grid(p, nodes)), dim3(kTG), s, y, f, nodes, p->r);
  }
  else
  {
    const bool react = (which == AR3D_RHS_ADVECTION_REACTION);
    ArArgs a;
    memset(&a, 0, sizeof(a));
    a.y = y;
    a.f = f;
    a.nxl = p->nxl; a.ny = p->ny; a.nz = p->nz;
    a.dir = (p->o.c > 0.0) ? 1 : -1;
    a.cx  = -p->o.c / p->dx;
    a.cy  = -p->o.c / p->dy;
    a.cz  = -p->o.c / p->dz;
    a.r   = p->r;
    a.np  = p->np;
    a.chunk = p->chunk;
    a.npush = p->npush;
    bool fast = p->fast;
    if (fast && (((uintptr_t)y % 32) || ((uintptr_t)f % 32)))
    {
      /* wrapped user pointers without 32-byte alignment: the generic kernel serves them on one
         rank; across ranks the two kernels count pushes differently, so refuse loudly */
      if (p->np > 1) return fail("RHS vectors must be 32-byte aligned on a multi-rank run");
      fast = false;
    }
    const unsigned long long per_call = fast ? (unsigned long long)p->tiles : (unsigned long long)p->npush;
    if (p->np > 1)
    {
      const unsigned long long seq = ++p->seq;
      const int par                = (int)(seq & 1ull);
      /* c > 0: my last plane goes east, the west neighbour's last plane comes in; c < 0 mirrored */
      const int down = (a.dir > 0) ? (p->rank + 1) % p->np : (p->rank + p->np - 1) % p->np;
      const int up   = (a.dir > 0) ? (p->rank + p->np - 1) % p->np : (p->rank + 1) % p->np;
      void* me       = p->peers[p->rank];
      a.send         = recv_plane(p->peers[down], p->plane, par);
      a.ctr_remote   = arrive_ctr(p->peers[down]);
      a.ctr_local    = arrive_ctr(me);
      a.halo         = recv_plane(me, p->plane, par);
      a.ack_remote   = ack_ctr(p->peers[up]);
      a.ack_local    = ack_ctr(me);
      p->arrivals += per_call;
      a.expected = p->arrivals;
      /* the buffer of this parity was last read in call seq-2: wait for that call's acknowledges */
      a.ack_expected = p->acks_by_parity[par];
      p->acks += per_call;
      p->acks_by_parity[par] = p->acks;
      if (b200vec_ctx_get_tuning(p->ctx, "profile") > 0)
        a.prof = (unsigned long long*)(uintptr_t)b200vec_ctx_get_tuning(p->ctx, "prof_counters_ptr");
    }
    else a.halo = y + (int64_t)((a.dir > 0) ? p->nxl - 1 : 0) * p->plane;
    if (fast)
    {
      const int nch = (p->nxl + p->chunk - 1) / p->chunk;
      dim3 grid((unsigned)p->tiles, (unsigned)nch);
      e = react ? launch_pdl(k_ar3d_march<true>, grid, dim3(kT), s, a) : launch_pdl(k_ar3d_march<false>, grid, dim3(kT), s, a);
    }
    else
    {
      dim3 grid((unsigned)p->grid_generic);
      e = react ? launch_pdl(k_ar3d_generic<true>, grid, dim3(kTG), s, a)
                : launch_pdl(k_ar3d_generic<false>, grid, dim3(kTG), s, a);
    }
  }
  if (e != cudaSuccess)
  {
    fprintf(stderr, "[ar3d_b200] RHS launch failed: %s\n", cudaGetErrorString(e));
    return -1;
  }
  if (p->time_rhs)
  {
    cudaEventRecord(p->e1, s);
    cudaEventSynchronize(p->e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, p->e0, p->e1);
    p->rhs_ms += ms;
  }
  p->rhs_calls++;
  return 0;
}

int b200_ar3d_psolve(b200_ar3d_plan p, const double* y, const double* b, double* x, double gamma)
{
  if (!p || !y || !b || !x) return -1;
  cudaStream_t s      = (cudaStream_t)b200vec_ctx_get_stream(p->ctx);
  const int64_t nodes = p->neq_loc / 3;
  const bool vec = (nodes % 4 == 0) && ((uintptr_t)y % 32 == 0) && ((uintptr_t)b % 32 == 0) && ((uintptr_t)x % 32 == 0);
  cudaError_t e;
  if (vec) e = launch_pdl(k_ar3d_psolve<true>, dim3((unsigned)node_grid(p, nodes / 4)), dim3(kTG), s, y, b, x, nodes, p->r, gamma);
  else e = launch_pdl(k_ar3d_psolve<false>, dim3((unsigned)node_grid(p, nodes)), dim3(kTG), s, y, b, x, nodes, p->r, gamma);
  p->psolve_calls++;
  if (e != cudaSuccess)
  {
    fprintf(stderr, "[ar3d_b200] PSolve launch failed: %s\n", cudaGetErrorString(e));
    return -1;
  }
  return 0;
}

} /* extern "C" */

/* ---- the driver: callbacks handed to the (unmodified) integrators ------------------ */
namespace {

struct Driver
{
  b200_ar3d_plan p = nullptr;
  SUNContext sunctx = nullptr;
  N_Vector umask = nullptr, vmask = nullptr, wmask = nullptr;
  long nnlfi = 0;
  FILE *TFID = nullptr, *UFID = nullptr, *VFID = nullptr, *WFID = nullptr;
  /* non-distributed aliases of distributed vectors for the task-local Newton solver
     (N_VGetLocalVector_MPIPlusX in arkode_driver.cpp:559-602) */
  std::vector<std::pair<N_Vector, N_Vector>> local_views;
  void* arkode_mem = nullptr;
  double rtol = 0.0, atol = 0.0; /* for the fused error-weight function */
};

double* dptr(N_Vector v) { return N_VGetDeviceArrayPointer_B200(v); }

/* rhs3D.hpp:30 / :322 / :386 */
int f_advection(sunrealtype, N_Vector y, N_Vector ydot, void* ud)
{
  return b200_ar3d_rhs(((Driver*)ud)->p, AR3D_RHS_ADVECTION, dptr(y), dptr(ydot));
}
int f_reaction(sunrealtype, N_Vector y, N_Vector ydot, void* ud)
{
  return b200_ar3d_rhs(((Driver*)ud)->p, AR3D_RHS_REACTION, dptr(y), dptr(ydot));
}
int f_advection_reaction(sunrealtype, N_Vector y, N_Vector ydot, void* ud)
{
  return b200_ar3d_rhs(((Driver*)ud)->p, AR3D_RHS_ADVECTION_REACTION, dptr(y), dptr(ydot));
}
/* ARKEwtFn / CVEwtFn: ewt = 1 / (rtol |y| + atol) in one kernel, the bits of arkEwtSetSS (arkode.c:2935-2947)
   and cvEwtSetSS (cvode.c:4794-4822) */
int ewt_cb(N_Vector y, N_Vector ewt, void* ud)
{
  auto* d = (Driver*)ud;
  return N_VEwtSet_B200(d->rtol, d->atol, nullptr, d->atol == 0.0, y, ewt);
}

/* rhs3D.hpp:672-686 */
int psolve_cb(sunrealtype, N_Vector y, N_Vector, N_Vector r, N_Vector z, sunrealtype gamma, sunrealtype, int, void* ud)
{
  return b200_ar3d_psolve(((Driver*)ud)->p, dptr(y), dptr(r), dptr(z), gamma);
}

/* ---- task-local Newton (arkode_driver.cpp:28-44, 545-791): every rank solves its own
 * nonlinear system (the reaction term couples nothing across nodes) with a Newton solver on
 * NON-distributed views -- no global reductions in the iteration -- and the ranks agree on
 * the outcome with two integer allreduces. */
struct TLContent
{
  Driver* drv;
  SUNNonlinearSolver local_nls;
  long ncnf;
};
#define TL(NLS) ((TLContent*)((NLS)->content))

N_Vector local_view(Driver* d, N_Vector v)
{
  for (auto& pr : d->local_views)
    if (pr.first == v)
    {
      if (dptr(pr.second) != dptr(v)) N_VSetDeviceArrayPointer_B200(dptr(v), pr.second);
      return pr.second;
    }
  N_Vector lv = N_VMakeWithCtx_B200(N_VGetLocalLength_B200(v), nullptr, dptr(v), N_VGetCtx_B200(v), d->sunctx);
  d->local_views.emplace_back(v, lv);
  return lv;
}

int tl_residual(N_Vector ycor, N_Vector F, void* arkode_mem)
{
  N_Vector z, zpred, Fi, sdata;
  sunrealtype tcur, gamma;
  void* user_data;
  ARKodeGetNonlinearSystemData(arkode_mem, &tcur, &zpred, &z, &Fi, &gamma, &sdata, &user_data);
  Driver* d = (Driver*)user_data;
  /* z = zpred + ycor on the local data (arkode_driver.cpp:568-569) */
  N_VLinearSum(1.0, local_view(d, zpred), 1.0, ycor, local_view(d, z));
  int rc = b200_ar3d_rhs(d->p, AR3D_RHS_REACTION, dptr(z), dptr(Fi));
  d->nnlfi++;
  if (rc < 0) return -1;
  if (rc > 0) return +1;
  sunrealtype c[3] = {1.0, -1.0, -gamma};
  N_Vector X[3]    = {ycor, local_view(d, sdata), local_view(d, Fi)};
  return N_VLinearCombination(3, c, X, F) != 0 ? -1 : 0;
}

int tl_lsolve(N_Vector delta, void* arkode_mem)
{
  N_Vector z, zpred, Fi, sdata;
  sunrealtype tcur, gamma;
  void* user_data;
  ARKodeGetNonlinearSystemData(arkode_mem, &tcur, &zpred, &z, &Fi, &gamma, &sdata, &user_data);
  Driver* d = (Driver*)user_data;
  return b200_ar3d_psolve(d->p, dptr(z), dptr(delta), dptr(delta), gamma);
}

SUNNonlinearSolver_Type tl_gettype(SUNNonlinearSolver) { return SUNNONLINEARSOLVER_ROOTFIND; }
int tl_initialize(SUNNonlinearSolver NLS)
{
  if (!NLS) return SUN_ERR_ARG_CORRUPT;
  SUNNonlinSolSetSysFn(TL(NLS)->local_nls, tl_residual);
  SUNNonlinSolSetLSolveFn(TL(NLS)->local_nls, tl_lsolve);
  return SUNNonlinSolInitialize(TL(NLS)->local_nls);
}
int tl_solve(SUNNonlinearSolver NLS, N_Vector y0, N_Vector ycor, N_Vector w, sunrealtype tol, sunbooleantype callLSetup,
             void* mem)
{
  if (!NLS || !y0 || !ycor || !w || !mem) return SUN_ERR_ARG_CORRUPT;
  Driver* d        = TL(NLS)->drv;
  int solve_status = SUNNonlinSolSolve(TL(NLS)->local_nls, local_view(d, y0), local_view(d, ycor), local_view(d, w), tol,
                                       callLSetup, mem);
  /* MPI_Allreduce(MIN) / (MAX) of the status (arkode_driver.cpp:660-668) */
  int64_t lo = solve_status, hi = solve_status;
  if (d->p->np > 1)
  {
    if (b200vec_allreduce_i64_host(d->p->ctx, &lo, B200VEC_MIN)) return SUN_ERR_GENERIC;
    if (lo < 0) return (int)lo;
    if (b200vec_allreduce_i64_host(d->p->ctx, &hi, B200VEC_MAX)) return SUN_ERR_GENERIC;
  }
  else if (lo < 0) return (int)lo;
  if (hi == SUN_NLS_CONV_RECVR) TL(NLS)->ncnf++;
  return (int)hi;
}
int tl_free(SUNNonlinearSolver NLS)
{
  if (!NLS) return SUN_SUCCESS;
  if (NLS->content)
  {
    SUNNonlinSolFree(TL(NLS)->local_nls);
    delete TL(NLS);
    NLS->content = nullptr;
  }
  SUNNonlinSolFreeEmpty(NLS);
  return SUN_SUCCESS;
}
int tl_setsysfn(SUNNonlinearSolver NLS, SUNNonlinSolSysFn fn) { return SUNNonlinSolSetSysFn(TL(NLS)->local_nls, fn); }
int tl_setctestfn(SUNNonlinearSolver NLS, SUNNonlinSolConvTestFn fn, void* data)
{
  return SUNNonlinSolSetConvTestFn(TL(NLS)->local_nls, fn, data);
}
int tl_getnumconvfails(SUNNonlinearSolver NLS, long int* n)
{
  *n = TL(NLS)->ncnf;
  return 0;
}

SUNNonlinearSolver TaskLocalNewton(Driver* d, N_Vector y)
{
  SUNNonlinearSolver NLS = SUNNonlinSolNewEmpty(d->sunctx);
  if (!NLS) return nullptr;
  NLS->ops->gettype         = tl_gettype;
  NLS->ops->initialize      = tl_initialize;
  NLS->ops->solve           = tl_solve;
  NLS->ops->free            = tl_free;
  NLS->ops->setsysfn        = tl_setsysfn;
  NLS->ops->setctestfn      = tl_setctestfn;
  NLS->ops->getnumconvfails = tl_getnumconvfails;
  auto* c                   = new TLContent();
  c->drv                    = d;
  c->ncnf                   = 0;
  c->local_nls              = SUNNonlinSol_Newton(local_view(d, y), d->sunctx);
  NLS->content              = c;
  if (!c->local_nls)
  {
    tl_free(NLS);
    return nullptr;
  }
  return NLS;
}

/* WriteOutput, advection_reaction_3D.cpp:619-700 */
int write_output(Driver* d, double t, N_Vector y, const b200_ar3d_opts* o, double* rms)
{
  b200_ar3d_plan p = d->p;
  const double N   = (double)((int64_t)p->nx * p->ny * p->nz);
  double u         = N_VWL2Norm(y, d->umask);
  u                = sqrt(u * u / N);
  double v         = N_VWL2Norm(y, d->vmask);
  v                = sqrt(v * v / N);
  double w         = N_VWL2Norm(y, d->wmask);
  w                = sqrt(w * w / N);
  if (p->rank == 0 && o->output)
  {
    printf("     %10.6f   %10.6f   %10.6f   %10.6f\n", t, u, v, w);
    fflush(stdout);
  }
  rms[0] = u; rms[1] = v; rms[2] = w;
  if (o->save)
  {
    if (p->rank == 0 && d->TFID)
    {
      fprintf(d->TFID, " %.16e\n", t);
      fflush(d->TFID);
    }
    N_VCopyFromDevice_B200(y);
    const double* h = N_VGetHostArrayPointer_B200(y);
    if (!h) return -1;
    const int64_t nodes = p->neq_loc / 3;
    for (int64_t n = 0; n < nodes; n++)
    {
      fprintf(d->UFID, " %.16e", h[3 * n]);
      fprintf(d->VFID, " %.16e", h[3 * n + 1]);
      fprintf(d->WFID, " %.16e", h[3 * n + 2]);
    }
    fprintf(d->UFID, "\n");
    fprintf(d->VFID, "\n");
    fprintf(d->WFID, "\n");
    fflush(d->UFID);
    fflush(d->VFID);
    fflush(d->WFID);
  }
  return 0;
}

#define CHK(call, what)                                                    \
  do {                                                                     \
    int flag_ = (call);                                                    \
    if (flag_ < 0)                                                         \
    {                                                                      \
      fprintf(stderr, "[ar3d_b200] %s failed with flag %d\n", what, flag_); \
      return -1;                                                           \
    }                                                                      \
  } while (0)

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

const char* method_name(int m)
{
  switch (m)
  {
  case AR3D_METHOD_ERK: return "ERK";
  case AR3D_METHOD_ARK_DIRK: return "ARK-DIRK";
  case AR3D_METHOD_ARK_IMEX: return "ARK-IMEX";
  case AR3D_METHOD_CV_BDF: return "CV-BDF";
  case AR3D_METHOD_CV_ADAMS: return "CV-ADAMS";
  }
  return "?";
}
const char* nls_name(int n, int method)
{
  if (method == AR3D_METHOD_ERK) return "none";
  if (method == AR3D_METHOD_CV_ADAMS) return "fixedpoint";
  return n == AR3D_NLS_NEWTON ? "newton" : n == AR3D_NLS_TL_NEWTON ? "tl-newton" : "fixedpoint";
}

} // namespace

extern "C" int b200_ar3d_run(b200vec_ctx ctx, const b200_ar3d_opts* opts, b200_ar3d_stats* st)
{
  if (!ctx || !opts || !st) return -1;
  memset(st, 0, sizeof(*st));
  const double t_setup0 = now_s();
  b200_ar3d_opts o      = *opts;
  /* ParseArgs, advection_reaction_3D.cpp:432-437 */
  if (o.method == AR3D_METHOD_CV_ADAMS) o.nls = AR3D_NLS_FIXEDPOINT;
  Driver d;
  d.rtol = o.rtol;
  d.atol = o.atol;
  if (b200_ar3d_plan_create(ctx, &o, &d.p)) return -1;
  b200_ar3d_plan p = d.p;
  CHK(SUNContext_Create(SUN_COMM_NULL, &d.sunctx), "SUNContext_Create");
  SUNContext sunctx = d.sunctx;
  const bool say    = o.output && p->rank == 0;

  N_Vector y = N_VNewWithCtx_B200(p->neq_loc, B200_MEM_DEVICE, ctx, sunctx);
  if (!y) return fail("N_VNewWithCtx_B200");
  if (p->np > 1 && N_VMakeDistributed_B200(y, p->neq)) return fail("N_VMakeDistributed_B200");
  if (o.fused) N_VEnableFusedOps_B200(y, SUNTRUE);
  d.umask = N_VClone(y);
  d.vmask = N_VClone(y);
  d.wmask = N_VClone(y);
  b200_ar3d_component_mask(p, 0, dptr(d.umask));
  b200_ar3d_component_mask(p, 1, dptr(d.vmask));
  b200_ar3d_component_mask(p, 2, dptr(d.wmask));
  if (o.save)
  {
    std::string dir = o.outputdir;
    char fname[2048];
    if (p->rank == 0)
    {
      snprintf(fname, sizeof(fname), "%s/t.%06d.txt", dir.c_str(), p->rank);
      d.TFID = fopen(fname, "w");
    }
    snprintf(fname, sizeof(fname), "%s/u.%06d.txt", dir.c_str(), p->rank);
    d.UFID = fopen(fname, "w");
    snprintf(fname, sizeof(fname), "%s/v.%06d.txt", dir.c_str(), p->rank);
    d.VFID = fopen(fname, "w");
    snprintf(fname, sizeof(fname), "%s/w.%06d.txt", dir.c_str(), p->rank);
    d.WFID = fopen(fname, "w");
    if (!d.UFID || !d.VFID || !d.WFID) return fail("cannot open the output files");
  }
  if (say)
  {
    /* SetupProblem + ParallelGrid::PrintInfo, advection_reaction_3D.cpp:530-553, ParallelGrid.hpp:418-431 */
    printf("\n\t\tAdvection-Reaction Test Problem\n\n");
    printf("Using the MPI+%s NVECTOR\n", "B200");
    printf("Number of Processors = %li\n", (long int)p->np);
    printf("ParallelGrid Info:\n");
    printf("    dimensions = %d\n", 3);
    printf("    processors = {%d, %d, %d}\n", p->np, 1, 1);
    printf("        domain = {[%g,%g], [%g,%g], [%g,%g]}\n", 0.0, o.xmax, 0.0, o.xmax, 0.0, o.xmax);
    printf("   global npts = {%li, %li, %li}\n", (long int)p->nx, (long int)p->ny, (long int)p->nz);
    printf("    local npts = {%d, %d, %d}\n", p->nxl, p->ny, p->nz);
    printf("  mesh spacing = {%g, %g, %g}\n", p->dx, p->dy, p->dz);
    printf(o.c > 0.0 ? "    upwind dir = right\n" : "    upwind dir = left\n");
    printf("Problem Parameters:\n");
    printf("  A = %g\n", o.A);
    printf("  B = %g\n", o.B);
    printf("  k = %g\n", o.k1);
    printf("  c = %g\n", o.c);
    printf("Integrator Options:\n");
    printf("  order            = %d\n", o.order);
    printf("  method           = %s\n", method_name(o.method));
    printf("  nonlinear solver = %s\n", nls_name(o.nls, o.method));
    printf("  fpaccel          = %d\n", o.fpaccel);
    printf("  preconditioner   = %d\n", o.precond);
    printf("  fused vector ops = %d\n", o.fused);
    printf("  t0               = %g\n", o.t0);
    printf("  tf               = %g\n", o.tf);
    printf("  reltol           = %.1e\n", o.rtol);
    printf("  abstol           = %.1e\n", o.atol);
    printf("  nout             = %d\n", o.nout);
    printf("Output directory: %s\n", o.outputdir);
  }
  if (b200_ar3d_set_ic(p, dptr(y))) return fail("SetIC");
  if (say && o.save && o.nout > 0)
  {
    /* ParallelGrid::MeshToFile, ParallelGrid.hpp:438-450 */
    std::string fn = std::string(o.outputdir) + "/mesh.txt";
    if (FILE* mf = fopen(fn.c_str(), "w"))
    {
      for (int a = 0; a < 3; a++)
      {
        const int n    = a == 0 ? p->nx : a == 1 ? p->ny : p->nz;
        const double h = a == 0 ? p->dx : a == 1 ? p->dy : p->dz;
        for (int i = 0; i < n; i++) fprintf(mf, " %.16g", h * i);
        fprintf(mf, "\n");
      }
      fclose(mf);
    }
  }

  /* ---- integrator setup: EvolveProblemExplicit / DIRK / IMEX (arkode_driver.cpp:51-543),
     EvolveProblemBDF / Adams (cvode_driver.cpp:26-329) */
  const bool is_cv  = (o.method == AR3D_METHOD_CV_BDF || o.method == AR3D_METHOD_CV_ADAMS);
  const bool is_erk = (o.method == AR3D_METHOD_ERK);
  void* mem              = nullptr;
  SUNNonlinearSolver NLS = nullptr;
  SUNLinearSolver LS     = nullptr;
  bool newton            = false;
  if (is_erk)
  {
    mem = ERKStepCreate(f_advection_reaction, o.t0, y, sunctx);
    if (!mem) return fail("ERKStepCreate");
    CHK(ARKodeSetOrder(mem, o.order), "ARKodeSetOrder");
    CHK(ARKodeSetUserData(mem, &d), "ARKodeSetUserData");
    CHK(ARKodeSStolerances(mem, o.rtol, o.atol), "ARKodeSStolerances");
    CHK(ARKodeSetMaxNumSteps(mem, 1000000), "ARKodeSetMaxNumSteps");
    CHK(ARKodeSetFixedStep(mem, 1e-5), "ARKodeSetFixedStep");
  }
  else if (!is_cv)
  {
    const bool imex = (o.method == AR3D_METHOD_ARK_IMEX);
    mem = imex ? ARKStepCreate(f_advection, f_reaction, o.t0, y, sunctx)
               : ARKStepCreate(nullptr, f_advection_reaction, o.t0, y, sunctx);
    if (!mem) return fail("ARKStepCreate");
    d.arkode_mem = mem;
    CHK(ARKodeSetOrder(mem, o.order), "ARKodeSetOrder");
    CHK(ARKodeSetUserData(mem, &d), "ARKodeSetUserData");
    CHK(ARKodeSStolerances(mem, o.rtol, o.atol), "ARKodeSStolerances");
    /* implicit / IMEX: the built-in routine would be arkEwtSetSS (the fixed-step ERK path above keeps the
       built-in choice: ARKODE swaps in arkEwtSetSmallReal there, arkode_erkstep.c:423) */
    if (o.fused_ewt) CHK(ARKodeWFtolerances(mem, ewt_cb), "ARKodeWFtolerances");
    CHK(ARKodeSetMaxNumSteps(mem, 100000), "ARKodeSetMaxNumSteps");
    if (o.nls == AR3D_NLS_NEWTON)
    {
      newton = true;
      NLS    = SUNNonlinSol_Newton(y, sunctx);
      if (!NLS) return fail("SUNNonlinSol_Newton");
      CHK(ARKodeSetNonlinearSolver(mem, NLS), "ARKodeSetNonlinearSolver");
      /* IMEX always preconditions (arkode_driver.cpp:309), DIRK honours --nopre (:114-115) */
      const int pt = (imex || o.precond) ? SUN_PREC_LEFT : SUN_PREC_NONE;
      LS           = SUNLinSol_SPGMR(y, pt, 0, sunctx);
      if (!LS) return fail("SUNLinSol_SPGMR");
      CHK(ARKodeSetLinearSolver(mem, LS, nullptr), "ARKodeSetLinearSolver");
      CHK(ARKodeSetPreconditioner(mem, nullptr, psolve_cb), "ARKodeSetPreconditioner");
    }
    else if (o.nls == AR3D_NLS_TL_NEWTON && imex)
    {
      NLS = TaskLocalNewton(&d, y);
      if (!NLS) return fail("TaskLocalNewton");
      CHK(ARKodeSetNonlinearSolver(mem, NLS), "ARKodeSetNonlinearSolver");
    }
    else if (o.nls == AR3D_NLS_FIXEDPOINT)
    {
      NLS = SUNNonlinSol_FixedPoint(y, o.fpaccel, sunctx);
      if (!NLS) return fail("SUNNonlinSol_FixedPoint");
      CHK(ARKodeSetNonlinearSolver(mem, NLS), "ARKodeSetNonlinearSolver");
    }
    else
    {
      fprintf(stderr, "\nERROR: %s is not compatible with the nls option provided\n", method_name(o.method));
      return 1;
    }
  }
  else
  {
    mem = CVodeCreate(o.method == AR3D_METHOD_CV_BDF ? CV_BDF : CV_ADAMS, sunctx);
    if (!mem) return fail("CVodeCreate");
    CHK(CVodeInit(mem, f_advection_reaction, o.t0, y), "CVodeInit");
    CHK(CVodeSetUserData(mem, &d), "CVodeSetUserData");
    CHK(CVodeSStolerances(mem, o.rtol, o.atol), "CVodeSStolerances");
    if (o.fused_ewt) CHK(CVodeWFtolerances(mem, ewt_cb), "CVodeWFtolerances");
    CHK(CVodeSetMaxNumSteps(mem, 100000), "CVodeSetMaxNumSteps");
    if (o.method == AR3D_METHOD_CV_BDF && o.nls == AR3D_NLS_NEWTON)
    {
      newton = true;
      NLS    = SUNNonlinSol_Newton(y, sunctx);
      if (!NLS) return fail("SUNNonlinSol_Newton");
      CHK(CVodeSetNonlinearSolver(mem, NLS), "CVodeSetNonlinearSolver");
      LS = SUNLinSol_SPGMR(y, o.precond ? SUN_PREC_LEFT : SUN_PREC_NONE, 0, sunctx);
      if (!LS) return fail("SUNLinSol_SPGMR");
      CHK(CVodeSetLinearSolver(mem, LS, nullptr), "CVodeSetLinearSolver");
      CHK(CVodeSetPreconditioner(mem, nullptr, psolve_cb), "CVodeSetPreconditioner");
    }
    else if (o.nls == AR3D_NLS_FIXEDPOINT)
    {
      NLS = SUNNonlinSol_FixedPoint(y, o.fpaccel, sunctx);
      if (!NLS) return fail("SUNNonlinSol_FixedPoint");
      CHK(CVodeSetNonlinearSolver(mem, NLS), "CVodeSetNonlinearSolver");
    }
    else
    {
      fprintf(stderr, "\nERROR: %s method is not compatible with the nls option provided\n", method_name(o.method));
      return 1;
    }
  }

  double rms[3] = {0, 0, 0};
  if (o.nout > 0)
  {
    if (say)
    {
      printf("\n          t         ||u||_rms   ||v||_rms   ||w||_rms\n");
      printf("   ----------------------------------------------------\n");
    }
    write_output(&d, o.t0, y, &o, rms);
  }
  b200vec_ctx_sync(ctx);
  st->setup_seconds = now_s() - t_setup0;

  double t = o.t0, dtout = (o.tf - o.t0);
  if (o.nout != 0) dtout /= o.nout;
  double tout   = t + dtout;
  int iout      = 0;
  int rc        = 0;
  const double t_ev0 = now_s();
  do {
    const int flag = is_cv ? CVode(mem, tout, y, &t, CV_NORMAL) : ARKodeEvolve(mem, tout, y, &t, ARK_NORMAL);
    if (flag < 0)
    {
      fprintf(stderr, "[ar3d_b200] %s failed with flag %d\n", is_cv ? "CVode" : "ARKodeEvolve", flag);
      rc = -1;
      break;
    }
    if (o.nout > 0) write_output(&d, t, y, &o, rms);
    tout += dtout;
    tout = (tout > o.tf) ? o.tf : tout;
    iout++;
  }
  while (iout < o.nout);
  b200vec_ctx_sync(ctx);
  st->evolve_seconds = now_s() - t_ev0;

  if (is_cv)
  {
    CVodeGetNumSteps(mem, &st->nst);
    CVodeGetNumRhsEvals(mem, &st->nfi);
    CVodeGetNumErrTestFails(mem, &st->netf);
    CVodeGetNumNonlinSolvIters(mem, &st->nni);
    CVodeGetNumNonlinSolvConvFails(mem, &st->ncnf);
    if (newton)
    {
      CVodeGetNumLinIters(mem, &st->nli);
      CVodeGetNumPrecSolves(mem, &st->npsol);
    }
    if (say)
    {
      printf("\nFinal Solver Statistics (for processor 0):\n");
      printf("   Internal solver steps = %li\n", st->nst);
      printf("   Total RHS evals: %li\n", st->nfi + d.nnlfi);
      printf("   Total number of error test failures = %li\n", st->netf);
      printf("   Total number of nonlinear solver convergence failures = %li\n", st->ncnf);
      /* EvolveProblemAdams stops here (cvode_driver.cpp:315-320); BDF goes on (:163-170) */
      if (o.method == AR3D_METHOD_CV_BDF) printf("   Total number of nonlinear iterations = %li\n", st->nni);
      if (newton)
      {
        printf("   Total number of linear iterations = %li\n", st->nli);
        printf("   Total number of preconditioner solves = %li\n", st->npsol);
      }
    }
  }
  else
  {
    ARKodeGetNumSteps(mem, &st->nst);
    ARKodeGetNumStepAttempts(mem, &st->nst_a);
    ARKodeGetNumRhsEvals(mem, 0, &st->nfe);
    if (!is_erk) ARKodeGetNumRhsEvals(mem, 1, &st->nfi);
    ARKodeGetNumErrTestFails(mem, &st->netf);
    if (!is_erk)
    {
      ARKodeGetNumNonlinSolvIters(mem, &st->nni);
      ARKodeGetNumNonlinSolvConvFails(mem, &st->ncnf);
    }
    if (newton)
    {
      ARKodeGetNumLinIters(mem, &st->nli);
      ARKodeGetNumPrecSolves(mem, &st->npsol);
    }
    if (say)
    {
      printf("\nFinal Solver Statistics (for processor 0):\n");
      printf("   Internal solver steps = %li (attempted = %li)\n", st->nst, st->nst_a);
      if (is_erk) printf("   Total RHS evals:  Fe = %li\n", st->nfe);
      else printf("   Total RHS evals:  Fe = %li,  Fi = %li\n", st->nfe, st->nfi + d.nnlfi);
      printf("   Total number of error test failures = %li\n", st->netf);
      if (!is_erk)
      {
        printf("   Total number of nonlinear solver convergence failures = %li\n", st->ncnf);
        printf("   Total number of nonlinear iterations = %li\n", st->nni);
        if (newton)
        {
          printf("   Total number of linear iterations = %li\n", st->nli);
          printf("   Total number of preconditioner solves = %li\n", st->npsol);
        }
      }
    }
  }
  if (say) fflush(stdout);
  st->nnlfi        = d.nnlfi;
  st->t_final      = t;
  st->urms         = rms[0];
  st->vrms         = rms[1];
  st->wrms         = rms[2];
  st->rhs_seconds  = p->rhs_ms * 1e-3;
  st->rhs_calls    = p->rhs_calls;
  st->psolve_calls = p->psolve_calls;
  st->neq          = p->neq;
  st->neq_loc      = p->neq_loc;
  st->nranks       = p->np;

  if (is_cv) CVodeFree(&mem);
  else ARKodeFree(&mem);
  if (NLS) SUNNonlinSolFree(NLS);
  if (LS) SUNLinSolFree(LS);
  for (auto& pr : d.local_views) N_VDestroy(pr.second);
  if (d.TFID) fclose(d.TFID);
  if (d.UFID) fclose(d.UFID);
  if (d.VFID) fclose(d.VFID);
  if (d.WFID) fclose(d.WFID);
  N_VDestroy(d.umask);
  N_VDestroy(d.vmask);
  N_VDestroy(d.wmask);
  N_VDestroy(y);
  b200_ar3d_plan_destroy(p);
  SUNContext_Free(&d.sunctx);
  return rc;
}
