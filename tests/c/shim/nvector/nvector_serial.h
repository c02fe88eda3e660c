/* tests/c/shim/nvector/nvector_serial.h -- TEST SHIM (our own file, not a copy).
 *
 * Placed ahead of the reference's include directory, this header makes an
 * UNMODIFIED reference example or unit test that says
 *     #include <nvector/nvector_serial.h>   ...   N_VNew_Serial(n, ctx)
 * allocate the B200 vector instead -- no source edit, the reference .c file is
 * compiled by path.  The serial examples read and write vector data on the host
 * through N_VGetArrayPointer / NV_Ith_S without any copy calls, so the shim
 * picks the host-coherent zero-copy kind (N_VNewPinned_B200): every op still
 * runs as a CUDA kernel on the GPU, and synchronises before returning.
 *
 * Build with -DB200_SHIM_MANAGED to use cudaMallocManaged storage instead.
 */
#ifndef B200_SHIM_NVECTOR_SERIAL_H
#define B200_SHIM_NVECTOR_SERIAL_H

#include "nvector_b200.h"

#ifdef B200_SHIM_MANAGED
#define N_VNew_Serial(n, ctx) N_VNewManaged_B200((n), (ctx))
#else
#define N_VNew_Serial(n, ctx) N_VNewPinned_B200((n), (ctx))
#endif
#define N_VEnableFusedOps_Serial(v, tf) N_VEnableFusedOps_B200((v), (tf))
#define NV_LENGTH_S(v)                  (N_VGetLength_B200(v))
#define NV_DATA_S(v)                    (N_VGetArrayPointer(v))
#define NV_Ith_S(v, i)                  (N_VGetArrayPointer(v)[i])

#endif
