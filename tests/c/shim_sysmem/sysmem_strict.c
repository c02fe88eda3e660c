/* sysmem_strict.c -- host-only SUNMemoryHelper stand-in (test infrastructure, our own code).
 *
 * The reference's benchmarks/advection_reaction_3D probes its memory helper for the
 * "best" memory type by asking for PINNED, DEVICE, UVM in turn and treating a non-zero
 * return as "not available" (raja/ParallelGrid.hpp:552-574).  The reference's system
 * helper only *asserts* on an unsupported type (src/sunmemory/system/
 * sundials_system_memory.c:83), so in a release build (assertions off, the configuration
 * oracle/Makefile uses) it hands back a NULL "pinned" buffer and the benchmark's serial
 * configuration dereferences it.  This stand-in defines SUNMemoryHelper_Sys with the
 * contract the probe expects -- allocation fails with an error code for anything but
 * host memory -- and is linked ahead of libsundials_ref.so for that one program only.
 */
#include <stdlib.h>
#include <string.h>
#include <sundials/sundials_errors.h>
#include <sundials/sundials_memory.h>

static SUNErrCode strict_alloc(SUNMemoryHelper h, SUNMemory* out, size_t bytes, SUNMemoryType type, void* queue)
{
  (void)queue;
  if (type != SUNMEMTYPE_HOST) return SUN_ERR_ARG_INCOMPATIBLE;
  SUNMemory m = SUNMemoryNewEmpty(h->sunctx);
  if (!m) return SUN_ERR_MALLOC_FAIL;
  m->ptr   = malloc(bytes ? bytes : 1);
  m->own   = SUNTRUE;
  m->type  = SUNMEMTYPE_HOST;
  m->bytes = bytes;
  if (!m->ptr)
  {
    free(m);
    return SUN_ERR_MALLOC_FAIL;
  }
  *out = m;
  return SUN_SUCCESS;
}

static SUNErrCode strict_dealloc(SUNMemoryHelper h, SUNMemory m, void* queue)
{
  (void)h;
  (void)queue;
  if (!m) return SUN_SUCCESS;
  if (m->own && m->ptr) free(m->ptr);
  free(m);
  return SUN_SUCCESS;
}

static SUNErrCode strict_copy(SUNMemoryHelper h, SUNMemory dst, SUNMemory src, size_t bytes, void* queue)
{
  (void)h;
  (void)queue;
  memcpy(dst->ptr, src->ptr, bytes);
  return SUN_SUCCESS;
}

static SUNErrCode strict_destroy(SUNMemoryHelper h)
{
  if (h)
  {
    free(h->ops);
    free(h);
  }
  return SUN_SUCCESS;
}

SUNMemoryHelper SUNMemoryHelper_Sys(SUNContext sunctx);
static SUNMemoryHelper strict_clone(SUNMemoryHelper h) { return SUNMemoryHelper_Sys(h->sunctx); }

SUNMemoryHelper SUNMemoryHelper_Sys(SUNContext sunctx)
{
  SUNMemoryHelper h = SUNMemoryHelper_NewEmpty(sunctx);
  if (!h) return NULL;
  h->ops->alloc   = strict_alloc;
  h->ops->dealloc = strict_dealloc;
  h->ops->copy    = strict_copy;
  h->ops->clone   = strict_clone;
  h->ops->destroy = strict_destroy;
  h->content      = NULL;
  return h;
}
