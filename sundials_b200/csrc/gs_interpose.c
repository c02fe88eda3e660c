/* gs_interpose.c -> libsundials_b200gs.so: symbol interposition of SUNClassicalGS / SUNModifiedGS.
 *
 * The reference's SPGMR / SPFGMR call SUNClassicalGS by name
 * (src/sunlinsol/spgmr/sunlinsol_spgmr.c:724,728, spfgmr/sunlinsol_spfgmr.c:691,695).  With this library
 * ahead of sundials_core in symbol resolution order (LD_PRELOAD, or linked first) the call lands
 * here: NVECTOR_B200 vectors go to the fused SUNClassicalGS_B200, anything else to the next
 * definition in link order -- the reference's own routine.  The reference stays unmodified. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>

#include "sundials_iterative_b200.h"

/* B200GS_REPORT=1: print how many calls were routed to the fused routine when the process ends */
static void report(void)
{
  fprintf(stderr, "[libsundials_b200gs] SUNClassicalGS_B200 calls: %ld\n", SUNClassicalGS_B200_Calls());
  fprintf(stderr, "[libsundials_b200gs] SUNModifiedGS_B200 calls: %ld\n", SUNModifiedGS_B200_Calls());
}
__attribute__((constructor)) static void init(void)
{
  const char* e = getenv("B200GS_REPORT");
  if (e && e[0] && e[0] != '0') atexit(report);
}

typedef SUNErrCode (*cgs_fn)(N_Vector*, sunrealtype**, int, int, sunrealtype*, sunrealtype*, N_Vector*);

SUNErrCode SUNClassicalGS(N_Vector* v, sunrealtype** h, int k, int p, sunrealtype* new_vk_norm, sunrealtype* stemp,
                          N_Vector* vtemp)
{
  if (v && v[0] && v[0]->ops && v[0]->ops->nvcloneempty == N_VCloneEmpty_B200)
    return SUNClassicalGS_B200(v, h, k, p, new_vk_norm, stemp, vtemp);
  static cgs_fn next = NULL;
  if (!next)
  {
    next = (cgs_fn)dlsym(RTLD_NEXT, "SUNClassicalGS");
    if (!next)
    {
      fprintf(stderr, "[libsundials_b200gs] no SUNClassicalGS after this library in link order\n");
      abort();
    }
  }
  return next(v, h, k, p, new_vk_norm, stemp, vtemp);
}

typedef SUNErrCode (*mgs_fn)(N_Vector*, sunrealtype**, int, int, sunrealtype*);

SUNErrCode SUNModifiedGS(N_Vector* v, sunrealtype** h, int k, int p, sunrealtype* new_vk_norm)
{
  if (v && v[0] && v[0]->ops && v[0]->ops->nvcloneempty == N_VCloneEmpty_B200)
    return SUNModifiedGS_B200(v, h, k, p, new_vk_norm);
  static mgs_fn next = NULL;
  if (!next)
  {
    next = (mgs_fn)dlsym(RTLD_NEXT, "SUNModifiedGS");
    if (!next)
    {
      fprintf(stderr, "[libsundials_b200gs] no SUNModifiedGS after this library in link order\n");
      abort();
    }
  }
  return next(v, h, k, p, new_vk_norm);
}
