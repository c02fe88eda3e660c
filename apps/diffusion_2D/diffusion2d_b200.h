/* diffusion2d_b200.h -- C ABI of the re-hosted 2-D diffusion benchmark.
 *
 * Re-host of the reference's benchmarks/diffusion_2D (ARKODE DIRK + PCG/GMRES +
 * Jacobi, main_arkode.cpp; RHS/halo code mpi_gpu/diffusion.cpp, buffers.cpp,
 * solution.cpp) WITHOUT MPI: one process per GPU, NVECTOR_B200 state vectors,
 * the right-hand side as one fused sm_100a kernel that also performs the halo
 * exchange over NVLink peer memory, global reductions inside the vector's
 * reduction kernels.  The integrator and Krylov solver are the unmodified
 * reference (ARKODE, SUNLinSol_PCG) reached through their public C API.
 *
 * Same problem, options and output as the reference benchmark
 * (benchmarks/diffusion_2D/README.md); the spatial decomposition is 1-D strips
 * in y (npx = 1, npy = ranks), which makes each rank's unknowns a contiguous
 * block of the global row-major vector -- the MPIPlusX partition.
 */
#ifndef DIFFUSION2D_B200_H
#define DIFFUSION2D_B200_H

#include <stdint.h>

#include "b200vec.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct
{
  /* problem (UserData, diffusion_2D.hpp:68-100) */
  int64_t nx, ny;  /* GLOBAL mesh points                       (32, 32) */
  double xu, yu;   /* domain upper bounds                      (1, 1)   */
  double kx, ky;   /* diffusion coefficients                   (1, 1)   */
  double tf;       /* final time                               (1)      */
  int forcing;     /* forcing term on/off                      (1)      */
  /* integrator and solver (UserOptions, main_arkode.cpp:26-52) */
  double rtol, atol; /* (1e-5, 1e-10) */
  int order;         /* DIRK order (3)                                   */
  int linear;        /* linearly implicit (1)                            */
  int ls_gmres;      /* 0 = PCG (default), 1 = SPGMR                     */
  int preconditioning; /* Jacobi on/off (1)                              */
  int liniters;      /* (20) */
  int msbp;          /* LSetup frequency, 0 = default (0)                */
  double epslin;     /* 0 = default (0)                                  */
  int maxsteps;      /* 0 = default (0)                                  */
  char controller[16]; /* ARKodeSetAdaptControllerByName ("I")           */
  /* output (UserOutput, diffusion_2D.hpp:202-207) */
  int output; /* 0 none, 1 table + statistics on rank 0 (1)              */
  int nout;   /* number of output times (20)                             */
  /* vector options */
  int fused_ops; /* N_VEnableFusedOps_B200 (1)                           */
  int rows_per_cta; /* RHS kernel: rows marched per CTA; 0 = one wave of equal row blocks (0) */
  int fused_ewt;    /* error weights by ONE kernel (N_VEwtSet_B200 as the ARKEwtFn registered with
                       ARKodeWFtolerances) instead of arkEwtSetSS's five vector ops; same bits (1) */
} b200_diffusion2d_opts;

typedef struct
{
  long nst, nst_a, netf, nfe, nfi, nni, ncfn, nsetups, nli, nlcf, npe, nps, njv, nfeLS;
  double t_final, urms, max_err;
  double evolve_seconds; /* wall time of the ARKodeEvolve loop (host clock, device synchronised) */
  double setup_seconds;
  double rhs_seconds;    /* device time spent in the RHS kernel (CUDA events), rhs_calls launches */
  long rhs_calls;
  int64_t nodes, nodes_loc;
  int nranks;
} b200_diffusion2d_stats;

void b200_diffusion2d_default_opts(b200_diffusion2d_opts* o);
/* ctx: execution context of this rank; for > 1 rank it must carry a
 * communicator (b200vec_comm_init) with the peer-memory transport available.
 * Returns 0 on success. */
int b200_diffusion2d_run(b200vec_ctx ctx, const b200_diffusion2d_opts* opts, b200_diffusion2d_stats* stats);

/* one RHS evaluation f = L u + b(t) on the rank's strip (testing / profiling):
 * u_dev, f_dev device arrays of nodes_loc doubles.  `plan` is created once. */
typedef struct b200_diffusion2d_plan_s* b200_diffusion2d_plan;
int b200_diffusion2d_plan_create(b200vec_ctx ctx, const b200_diffusion2d_opts* opts, b200_diffusion2d_plan* out);
int64_t b200_diffusion2d_plan_local_nodes(b200_diffusion2d_plan p);
int b200_diffusion2d_rhs(b200_diffusion2d_plan p, double t, const double* u_dev, double* f_dev);
int b200_diffusion2d_solution(b200_diffusion2d_plan p, double t, double* u_dev);
void b200_diffusion2d_plan_destroy(b200_diffusion2d_plan p);

#ifdef __cplusplus
}
#endif
#endif
