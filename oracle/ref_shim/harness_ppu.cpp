// harness_ppu.cpp — C entry points around the reference's CPU renderer.  TEST INFRASTRUCTURE ONLY.
//
// oracle/Makefile (target ref) compiles cell/ppu_renderer.cpp — unmodified, from where it lies under
// $(REFERENCE) — together with this file into oracle/_ref/libppu_renderer_ref.so. Everything on the path that the
// snapshot contains therefore runs as the reference wrote it: SVOData::Load (cell/svodata.h:31-50), the RendererBase
// setters and InitRayDir (cell/renderer_base.h:25-61) over the reference's own cg::point_t operators
// (nest/include/geometry/primitives/point.h), the per-pixel loop RenderRect and the recursion RecTrace
// (cell/ppu_renderer.cpp:18-70), SimpleRenderer / TreadedRenderer::RenderFrame (:76-85,:121-144). What the snapshot
// lacks (cpp/*.h) is supplied by the stand-ins in oracle/ref_shim/cpp, each of which says what it is.
// This file only drives the ISVORenderer interface (cell/svorenderer.h:5-30) the way cell/main.cpp:21-40 does.
#include "stdafx.h"
#include "svodata.h"
#include "svorenderer.h"
#include "renderer_base.h"

namespace {
// RendererBase::InitRayDir is protected (renderer_base.h:49-61); a subclass hands its result out
struct RayDirProbe : public RendererBase {
  virtual const Color32 *RenderFrame() { return NULL; }
  void Get(RayDirData &rdd) { InitRayDir(rdd); }
};
}

extern "C" {

int yv_ref_shader_probe = 0;

void *yv_ref_scene_load(const char *path) {
  SVOData *s = new SVOData;
  std::streambuf *out = std::cout.rdbuf(NULL);      // Load announces itself on std::cout (svodata.h:33,49); callers
  s->Load(path);                                    // such as bench.py own stdout, so the text is dropped
  std::cout.rdbuf(out);
  std::cout.clear();
  return s;
}
unsigned int yv_ref_scene_root(void *scene) { return static_cast<SVOData *>(scene)->GetRoot(); }
void yv_ref_scene_free(void *scene) { delete static_cast<SVOData *>(scene); }

// kind: 0 = CreateSimpleRenderer, 1 = CreateThreadedRenderer. up == NULL / fov <= 0 / width <= 0 keep the
// renderer's defaults (renderer_base.h:25). Returns 1 and fills out[W*H] (and *w,*h) if RenderFrame gave a frame,
// 0 if it returned NULL.
int yv_ref_ppu_render(void *scene, int kind, const float *pos, const float *dir, const float *up, float fov,
                      int width, int height, int probe, unsigned int *out, int *w, int *h) {
  shared_ptr<ISVORenderer> r = kind ? CreateThreadedRenderer() : CreateSimpleRenderer();
  if (scene) r->SetScene(static_cast<SVOData *>(scene));
  if (width > 0) r->SetResolution(width, height);
  if (pos) r->SetViewPos(point_3f(pos[0], pos[1], pos[2]));
  if (dir) r->SetViewDir(point_3f(dir[0], dir[1], dir[2]));
  if (up) r->SetViewUp(point_3f(up[0], up[1], up[2]));
  if (fov > 0) r->SetFOV(fov);
  const point_2i size = r->GetResolution();
  if (w) *w = size.x;
  if (h) *h = size.y;
  yv_ref_shader_probe = probe;
  const Color32 *frame = r->RenderFrame();
  yv_ref_shader_probe = 0;
  if (!frame) return 0;
  if (out) memcpy(out, frame, sizeof(Color32) * (size_t)size.x * (size_t)size.y);
  return 1;
}

// The reference's own InitRayDir: out[0..2] = dir0, [3..5] = du, [6..8] = dv.
void yv_ref_init_ray_dir(const float *dir, const float *up, float fov, int width, int height, float *out) {
  RayDirProbe r;
  r.SetResolution(width, height);
  r.SetViewDir(point_3f(dir[0], dir[1], dir[2]));
  if (up) r.SetViewUp(point_3f(up[0], up[1], up[2]));
  if (fov > 0) r.SetFOV(fov);
  RayDirData rdd;
  r.Get(rdd);
  const point_3f v[3] = { rdd.dir0, rdd.du, rdd.dv };
  for (int i = 0; i < 3; ++i) { out[3 * i] = v[i].x; out[3 * i + 1] = v[i].y; out[3 * i + 2] = v[i].z; }
}

}  // extern "C"
