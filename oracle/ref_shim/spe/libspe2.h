/* libspe2.h — stand-in for the Cell SDK's SPE runtime management library (absent).  TEST INFRASTRUCTURE ONLY.
 * cell/spu_renderer.cpp:34-59 creates one SPE context per thread, loads the embedded program `trace_spu`, runs it with a
 * pointer to a parameter block, destroys the context; :73 asks how many SPEs there are. Here an "SPE" is the host CPU
 * running the program's entry point (cell/spu/trace_spu.cpp:149, renamed trace_spu_main by the build):
 *   - the program keeps its state in globals (params, node cache, result block), as one program image per SPE would;
 *     on the host there is one image, so spe_context_run holds a lock and the "SPEs" run one after another — the
 *     block-interleaved split of the frame over them (blockStart / blockStride, :80-83) is exercised all the same;
 *   - effective addresses are 32-bit on the SPU side (`(unsigned int)parm`, trace_spu.cpp:153), so the parameter block,
 *     which the caller keeps on its stack, is copied below 2 GiB before the program starts (the pool and the colour
 *     buffer already live there: oracle/ref_shim/spu/libmisc.h, YV_SHIM_LOW_MEMORY).
 * YV_SHIM_SPES (environment) = number of usable SPEs reported, default 6 (a PlayStation 3's). */
#ifndef YV_REF_SHIM_LIBSPE2_H
#define YV_REF_SHIM_LIBSPE2_H
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <sys/mman.h>

typedef struct yv_shim_spe_context { int loaded; } *spe_context_ptr_t;
typedef struct spe_program_handle { int unused; } spe_program_handle_t;
typedef struct spe_stop_info spe_stop_info_t;
#define SPE_DEFAULT_ENTRY 0xffffffffu
#define SPE_COUNT_USABLE_SPES 3

int trace_spu_main(unsigned long long spu_id, unsigned long long parm);      /* cell/spu/trace_spu.cpp:149 */

static inline spe_context_ptr_t spe_context_create(unsigned int, void *) { return new yv_shim_spe_context(); }
static inline int spe_program_load(spe_context_ptr_t ctx, spe_program_handle_t *) { ctx->loaded = 1; return 0; }
static inline int spe_context_destroy(spe_context_ptr_t ctx) { delete ctx; return 0; }
static inline int spe_cpu_info_get(int, int) { const char *e = getenv("YV_SHIM_SPES"); int n = e ? atoi(e) : 6; return n > 0 ? n : 1; }
static inline int spe_context_run(spe_context_ptr_t ctx, unsigned int *, unsigned int, void *argp, void *, spe_stop_info_t *) {
  static std::mutex one_program_image;
  if (!ctx || !ctx->loaded) return -1;
  std::lock_guard<std::mutex> lock(one_program_image);
  static void *low = mmap(NULL, 4096, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_32BIT, -1, 0);
  if (low == MAP_FAILED) return -1;
  memcpy(low, argp, 256);                       /* trace_spu_params is ~100 bytes; the caller's stack is readable beyond it */
  trace_spu_main(0, (unsigned long long)(size_t)low);
  return 0;
}
#endif
