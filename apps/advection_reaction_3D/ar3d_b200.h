/* ar3d_b200.h -- C ABI of the re-hosted 3-D advection-reaction benchmark.
 *
 * Re-host of the reference's benchmarks/advection_reaction_3D/raja (ARKODE ERK /
 * DIRK / IMEX-ARK with Newton+SPGMR, task-local Newton or Anderson-accelerated
 * fixed point: advection_reaction_3D.cpp, arkode_driver.cpp; RHS, Jacobian
 * solves and preconditioner: rhs3D.hpp; mesh + halo exchange: ParallelGrid.hpp)
 * WITHOUT MPI and without RAJA: one process per GPU, NVECTOR_B200 state vectors,
 * the right-hand side as hand-written sm_100a kernels that perform the upwind
 * halo exchange over NVLink peer memory inside the kernel, global reductions
 * inside the vector's reduction kernels.  Integrator, nonlinear and Krylov
 * solvers are the unmodified reference reached through their public C API.
 *
 * Same problem, options and screen output as the reference benchmark
 * (benchmarks/advection_reaction_3D/raja/README.md).  State layout is the
 * reference's: y[((i*ny + j)*nz + k)*3 + l], l = (u, v, w) fastest
 * (RAJA::Layout<4>(nxl, nyl, nzl, dof), rhs3D.hpp:62).  The decomposition is
 * 1-D slabs in x (npxyz = {ranks, 1, 1}), which makes every rank's unknowns a
 * contiguous block of the global vector -- the MPIPlusX partition -- and leaves
 * one exchanged face (west for c > 0, east for c < 0); the y and z faces wrap
 * periodically inside the rank.
 */
#ifndef AR3D_B200_H
#define AR3D_B200_H

#include <stdint.h>

#include "b200vec.h"

#ifdef __cplusplus
extern "C" {
#endif

/* method (UserOptions::method, advection_reaction_3D.cpp:368-379) */
#define AR3D_METHOD_ERK      0
#define AR3D_METHOD_ARK_DIRK 1
#define AR3D_METHOD_ARK_IMEX 2
#define AR3D_METHOD_CV_BDF   3
#define AR3D_METHOD_CV_ADAMS 4
/* nonlinear solver (UserOptions::nls, :384-395) */
#define AR3D_NLS_NEWTON     0
#define AR3D_NLS_TL_NEWTON  1
#define AR3D_NLS_FIXEDPOINT 2

/* which terms b200_ar3d_rhs evaluates (rhs3D.hpp:30 Advection, :322 Reaction, :386 AdvectionReaction) */
#define AR3D_RHS_ADVECTION          0
#define AR3D_RHS_REACTION           1
#define AR3D_RHS_ADVECTION_REACTION 2

typedef struct
{
  /* problem (UserData defaults, advection_reaction_3D.cpp:453-471) */
  int64_t npts;  /* GLOBAL mesh points per direction              (100)      */
  double xmax;   /* domain [0, xmax]^3                            (1)        */
  double A, B;   /* concentrations                                (1, 3.5)   */
  double k1, k2, k3, k4, k5, k6; /* rates                         (1,1,1,1, 2e5, 2e5) */
  double c;      /* advection speed                               (0.01)     */
  /* integrator (UserOptions defaults, :474-490) */
  int method;    /* AR3D_METHOD_*                                 (ARK_DIRK) */
  int nls;       /* AR3D_NLS_*                                    (NEWTON)   */
  int order;     /*                                               (3)        */
  int fpaccel;   /* Anderson vectors of the fixed-point solver    (3)        */
  int precond;   /* reaction-block preconditioner on/off          (1)        */
  int fused;     /* N_VEnableFusedOps_B200                        (0, as the reference) */
  double t0, tf; /*                                               (0, 10)    */
  double rtol, atol; /*                                           (1e-6, 1e-9) */
  int nout;      /* output times                                  (10)       */
  int save;      /* write u/v/w.<rank>.txt + t.000000.txt + mesh.txt into outputdir (0) */
  char outputdir[1024];
  /* execution */
  int output;       /* 0: silent, 1: the reference's screen output on rank 0 (1)                */
  int force_generic;/* 1: always use the one-node-per-thread kernel (testing)            (0)   */
  int planes_per_cta; /* fast RHS kernel: x-planes marched per CTA, 0 = default          (0)   */
  int fused_ewt;    /* implicit / IMEX ARKODE and CVODE runs: error weights by ONE kernel (N_VEwtSet_B200 as the
                       EwtFn registered with ARKodeWFtolerances / CVodeWFtolerances) instead of the five
                       vector ops of arkEwtSetSS / cvEwtSetSS; same bits                  (1)   */
} b200_ar3d_opts;

typedef struct
{
  long nst, nst_a, netf, nfe, nfi, nni, ncnf, nli, npsol, nnlfi;
  double t_final, urms, vrms, wrms;
  double evolve_seconds; /* wall time of the Evolve loop incl. the per-output norms (device synchronised) */
  double setup_seconds;
  double rhs_seconds;    /* device time in the RHS kernels (CUDA events; only with B200_AR3D_TIME_RHS=1) */
  long rhs_calls, psolve_calls;
  int64_t neq, neq_loc;
  int nranks;
} b200_ar3d_stats;

void b200_ar3d_default_opts(b200_ar3d_opts* o);
/* ctx: execution context of this rank; for > 1 rank it must carry a communicator
 * (b200vec_comm_init) with the peer-memory transport.  Returns 0 on success. */
int b200_ar3d_run(b200vec_ctx ctx, const b200_ar3d_opts* opts, b200_ar3d_stats* stats);

/* the building blocks on raw device arrays of neq_loc doubles (tests, profiling) */
typedef struct b200_ar3d_plan_s* b200_ar3d_plan;
int b200_ar3d_plan_create(b200vec_ctx ctx, const b200_ar3d_opts* opts, b200_ar3d_plan* out);
void b200_ar3d_plan_destroy(b200_ar3d_plan p);
int64_t b200_ar3d_plan_local_neq(b200_ar3d_plan p);
/* 1 when the plane-marching kernel serves this plan, 0 for the generic kernel */
int b200_ar3d_plan_is_fast(b200_ar3d_plan p);
/* SetIC, advection_reaction_3D.cpp:560-616 */
int b200_ar3d_set_ic(b200_ar3d_plan p, double* y_dev);
/* ComponentMask, :405-428 */
int b200_ar3d_component_mask(b200_ar3d_plan p, int component, double* mask_dev);
/* which = AR3D_RHS_*; collective over the ranks when it contains advection */
int b200_ar3d_rhs(b200_ar3d_plan p, int which, const double* y_dev, double* f_dev);
/* x = (I - gamma dg/dy)^-1 b, SolveReactionLinSys rhs3D.hpp:441-550; x may alias b */
int b200_ar3d_psolve(b200_ar3d_plan p, const double* y_dev, const double* b_dev, double* x_dev, double gamma);

#ifdef __cplusplus
}
#endif
#endif
