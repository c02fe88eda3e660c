/* ar3d_rhs_dump.cpp -- TEST INFRASTRUCTURE: golden-vector generator for apps/advection_reaction_3D.
 *
 * Calls the REFERENCE's own right-hand-side functions -- Advection, Reaction,
 * AdvectionReaction, SolveReactionLinSys (benchmarks/advection_reaction_3D/raja/rhs3D.hpp,
 * included by path) and SetIC / SetupProblem (advection_reaction_3D.cpp, compiled by path with
 * its main() renamed) -- on inputs read from a file, and writes their outputs:
 *
 *   ar3d_rhs_dump <npts> <c> <gamma> <in.bin> <out.bin>
 *     in.bin : y[neq], b[neq]                               (doubles)
 *     out.bin: ic[neq], fe[neq], fi[neq], f[neq], x[neq]    (doubles)
 *       ic = SetIC; fe = Advection(y); fi = Reaction(y) alone; f = AdvectionReaction(y);
 *       x  = SolveReactionLinSys(y, b, gamma)
 * Built by tests/c/Makefile against the sequential RAJA and single-rank MPI stand-ins;
 * tests/golden/make_ar3d_golden.py turns the outputs into tests/golden/advection_reaction_3D/.
 */
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "advection_reaction_3D.hpp"
#include "rhs3D.hpp"

static int read_doubles(FILE* f, sunrealtype* p, size_t n) { return fread(p, sizeof(sunrealtype), n, f) == n ? 0 : -1; }

int main(int argc, char* argv[])
{
  if (argc != 6)
  {
    fprintf(stderr, "usage: %s npts c gamma in.bin out.bin\n", argv[0]);
    return 2;
  }
  const double gamma = strtod(argv[3], NULL);
  MPI_Comm comm      = MPI_COMM_WORLD;
  MPI_Init(&argc, &argv);
  SUNContext ctx;
  SUNContext_Create(comm, &ctx);
  SUNMemoryHelper mem_helper = SUNMemoryHelper_Sys(ctx);
  int rc = 0;
  {
    UserData udata(ctx);
    UserOptions uopt;
    std::string s_npts = argv[1], s_c = argv[2];
    char* av[] = {argv[0], (char*)"--npts", (char*)s_npts.c_str(), (char*)"--c", (char*)s_c.c_str(),
                  (char*)"--dont-save", (char*)"--nout", (char*)"0"};
    if (SetupProblem(8, av, &udata, &uopt, mem_helper, ctx)) return 3;
    const sunindextype neq = udata.grid->neq;
    N_Vector y  = N_VMake_MPIPlusX(udata.comm, LocalNvector(neq, ctx), ctx);
    N_Vector b  = N_VClone(y);
    N_Vector o  = N_VClone(y);
    FILE* fin   = fopen(argv[4], "rb");
    FILE* fout  = fopen(argv[5], "wb");
    if (!fin || !fout) return 4;
    if (read_doubles(fin, GetVecData(y), neq) || read_doubles(fin, GetVecData(b), neq)) return 5;
    fclose(fin);

    SetIC(o, &udata);
    fwrite(GetVecData(o), sizeof(sunrealtype), neq, fout);

    udata.add_reactions = false;
    N_VConst(-7.0, o);
    rc |= Advection(0.0, y, o, &udata);
    fwrite(GetVecData(o), sizeof(sunrealtype), neq, fout);
    N_VConst(-7.0, o);
    rc |= Reaction(0.0, y, o, &udata);
    fwrite(GetVecData(o), sizeof(sunrealtype), neq, fout);

    udata.add_reactions = true;
    N_VConst(-7.0, o);
    rc |= AdvectionReaction(0.0, y, o, &udata);
    fwrite(GetVecData(o), sizeof(sunrealtype), neq, fout);

    N_VConst(-7.0, o);
    rc |= SolveReactionLinSys(y, o, b, gamma, &udata);
    fwrite(GetVecData(o), sizeof(sunrealtype), neq, fout);
    fclose(fout);

    N_VDestroy(b); /* clones own their local vector */
    N_VDestroy(o);
    N_VDestroy(N_VGetLocalVector_MPIPlusX(y));
    N_VDestroy(y);
  }
  SUNMemoryHelper_Destroy(mem_helper);
  SUNContext_Free(&ctx);
  MPI_Finalize();
  return rc ? 6 : 0;
}
