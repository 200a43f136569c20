/* libmisc.h — stand-in for the Cell SDK header of that name (absent).  TEST INFRASTRUCTURE ONLY.
 * cell/spu/trace_spu.cpp includes it and uses nothing from it; with -DTARGET_PPU cell/alignedarray.h:27-41 takes
 * malloc_align(size, log2_alignment) / free_align(ptr) from it. */
#ifndef YV_REF_SHIM_LIBMISC_H
#define YV_REF_SHIM_LIBMISC_H
#include <stdlib.h>
static inline void *malloc_align(size_t size, unsigned int log2_align) {
  void *p = NULL;
  size_t a = (size_t)1 << log2_align;
  if (a < sizeof(void *)) a = sizeof(void *);
  return posix_memalign(&p, a, size ? size : 1) == 0 ? p : NULL;
}
static inline void free_align(void *p) { free(p); }
#endif
