/* libmisc.h — stand-in for the Cell SDK header of that name (absent).  TEST INFRASTRUCTURE ONLY.
 * cell/spu/trace_spu.cpp includes it and uses nothing from it; with -DTARGET_PPU cell/alignedarray.h:27-41 takes
 * malloc_align(size, log2_alignment) / free_align(ptr) from it. With YV_SHIM_LOW_MEMORY the memory comes from below
 * 2 GiB (MAP_32BIT): the SPU program reaches the node pool and the colour buffer through 32-bit effective addresses
 * (cell/spu/trace_spu.cpp:28,174), so when it runs on the host those arrays have to live there. */
#ifndef YV_REF_SHIM_LIBMISC_H
#define YV_REF_SHIM_LIBMISC_H
#include <stdlib.h>
#ifdef YV_SHIM_LOW_MEMORY
#include <sys/mman.h>
static inline void *malloc_align(size_t size, unsigned int log2_align) {
  (void)log2_align;                                       /* pages are aligned to 4096 */
  const size_t bytes = size + 4096;
  char *p = (char *)mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_32BIT, -1, 0);
  if (p == (char *)MAP_FAILED) return NULL;
  *(size_t *)p = bytes;
  return p + 4096;
}
static inline void free_align(void *p) { if (p) { char *b = (char *)p - 4096; munmap(b, *(size_t *)b); } }
#else
static inline void *malloc_align(size_t size, unsigned int log2_align) {
  void *p = NULL;
  size_t a = (size_t)1 << log2_align;
  if (a < sizeof(void *)) a = sizeof(void *);
  return posix_memalign(&p, a, size ? size : 1) == 0 ? p : NULL;
}
static inline void free_align(void *p) { free(p); }
#endif
#endif
