// harness_spu.cpp — runs the reference's SPU tracing program on the host.  TEST INFRASTRUCTURE ONLY.
//
// oracle/Makefile (target ref) compiles cell/spu/trace_spu.cpp — unmodified, from where it lies under $(REFERENCE),
// with -DTARGET_SPU as its own Makefile does and its entry point renamed (-Dmain=trace_spu_main) — together with this
// file into oracle/_ref/libtrace_spu_f32_ref.so. FetchNode (the 2048-entry software node cache, trace_spu.cpp:15-35),
// FindFirstChildSPU (:48-68), GoNextSPU (:70-93), RecTrace (:97-116), RenderBlock (:121-146) and the block loop of
// main (:149-181) run as written; the SIMD intrinsics and the MFC DMA calls come from oracle/ref_shim/spu, the absent
// cpp/*.h from oracle/ref_shim/cpp.
// This file plays the PPU side (cell/spu_renderer.cpp:44-56): it fills trace_spu_params and starts the program.
// Effective addresses on the SPU are 32-bit (`(unsigned int)node_ptr`, trace_spu.cpp:28,151,174), so everything the
// program reaches by DMA — the parameter block, the node pool, the colour buffer — is placed below 2 GiB (MAP_32BIT).
#include "stdafx.h"
#include "trace_spu.h"
#include <sys/mman.h>

int trace_spu_main(unsigned long long spu_id, unsigned long long parm);     // cell/spu/trace_spu.cpp:149 (renamed)
extern int missCount, fetchCount;                                            // trace_spu.cpp:18-19

extern "C" {

int yv_ref_shader_probe = 0;

static void *low_alloc(size_t bytes) {
  void *p = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_32BIT, -1, 0);
  return p == MAP_FAILED ? NULL : p;
}

// Renders the blocks the program covers (viewSize / BlockSize in each direction, trace_spu.cpp:162) into out[W*H];
// pixels outside them keep the value `fill`. stats[0..1] = node fetches / cache misses of this run.
int yv_ref_spu_render(const void *nodes, unsigned int count, unsigned int root, const float *pos, const float *dir0,
                      const float *du, const float *dv, int width, int height, int probe, unsigned int fill,
                      unsigned int *out, int *stats) {
  const size_t node_bytes = (size_t)count * sizeof(VoxNode), pix_bytes = (size_t)width * height * sizeof(Color32);
  char *pool = (char *)low_alloc(node_bytes + 64);
  char *frame = (char *)low_alloc(pix_bytes + 64);
  trace_spu_params *p = (trace_spu_params *)low_alloc(4096);
  if (!pool || !frame || !p) return 0;
  memcpy(pool, nodes, node_bytes);
  for (size_t i = 0; i < (size_t)width * height; ++i) memcpy(frame + 4 * i, &fill, 4);
  p->pos = point_3f(pos[0], pos[1], pos[2]);
  p->rdd.dir0 = point_3f(dir0[0], dir0[1], dir0[2]);
  p->rdd.du = point_3f(du[0], du[1], du[2]);
  p->rdd.dv = point_3f(dv[0], dv[1], dv[2]);
  p->viewSize = point_2i(width, height);
  p->blockStart = 0;
  p->blockStride = 1;
  p->colorBuf = (Color32 *)frame;
  p->root = root;
  p->nodes = (const VoxNode *)pool;
  const int f0 = fetchCount, m0 = missCount;
  yv_ref_shader_probe = probe;
  trace_spu_main(0, (unsigned long long)(uintptr_t)p);
  yv_ref_shader_probe = 0;
  if (stats) { stats[0] = fetchCount - f0; stats[1] = missCount - m0; }
  memcpy(out, frame, pix_bytes);
  munmap(pool, node_bytes + 64); munmap(frame, pix_bytes + 64); munmap(p, 4096);
  return 1;
}

}  // extern "C"
