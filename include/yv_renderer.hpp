// yv_renderer.hpp — header-only C++ adapter: the reference's renderer interface over the C ABI.
//
// Mirrors, name for name, what a caller of the reference sees:
//   SVOData ........ cell/svodata.h:22-55        (Load, GetRoot)
//   ISVORenderer ... cell/svorenderer.h:5-24     (SetScene, SetViewPos/Dir/Up, SetResolution,
//                                                 GetResolution, SetFOV, RenderFrame)
//   factory ........ cell/svorenderer.h:26-30    (CreateSimpleRenderer / CreateThreadedRenderer /
//                                                 CreateSPURenderer -> CreateB200Renderer)
//   Render(d_ptr) .. demo/SVORenderer.h:36       (device-pointer variant)
// so a driver written like cell/main.cpp:21-40 compiles unchanged against it. Inside the reference
// tree, include the tree's svorenderer.h and define YV_USE_REFERENCE_TYPES before including this header: the
// adapter then derives from the tree's own ISVORenderer, takes the tree's SVOData / point_3f / point_2i / Color32
// and imports the scene's node pool on SetScene (integration/b200_renderer.cpp is that translation unit; the
// tests build the reference's unmodified cell/main.cpp against it — see INTEGRATION.md).
// Error behaviour follows the reference: no exceptions; RenderFrame() returns NULL when it cannot
// render (no scene, no GPU — there is no CPU fallback); yv_last_error() has the reason.
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "yv_b200.h"

#ifndef YV_USE_REFERENCE_TYPES
namespace yv {
struct point_3f { float x, y, z; point_3f() : x(0), y(0), z(0) {} point_3f(float a, float b, float c) : x(a), y(b), z(c) {} };
struct point_2i { int x, y; point_2i() : x(0), y(0) {} point_2i(int a, int b) : x(a), y(b) {} };
struct Color32 { uint8_t r, g, b, a; };

// SVOData (cell/svodata.h:22-55)
class SVOData {
 public:
  SVOData() : h_(nullptr) {}
  ~SVOData() { yv_svo_free(h_); }
  SVOData(const SVOData &) = delete;
  SVOData &operator=(const SVOData &) = delete;
  // :31 — reloads in place, as the reference does: renderers that were given this object (SetScene keeps the pointer,
  // renderer_base.h:28) stay valid. Failures: a fresh object keeps GetRoot()==EmptyNode, a loaded one keeps its scene.
  void Load(const char *fn) { if (h_) yv_svo_load_into(h_, fn); else yv_svo_load(fn, &h_); }
  bool BuildSphereFractal(int depth, int threads) { reset(); return yv_svo_build_sphere_fractal(depth, threads, &h_) == 0; }
  bool BuildIsoVolume(int depth, uint32_t seed, int iso, int threads) { reset(); return yv_svo_build_iso_volume(depth, seed, iso, threads, &h_) == 0; }
  bool Save(const char *fn) const { return h_ && yv_svo_save(h_, fn) == 0; }
  yv_node_id GetRoot() const { return h_ ? yv_svo_root(h_) : YV_EMPTY_NODE; }        // :52
  const yv_vox_node &operator[](yv_node_id id) const { return yv_svo_nodes(h_)[id]; }  // :54
  uint32_t GetNodeCount() const { return h_ ? yv_svo_node_count(h_) : 0; }
  yv_svo *handle() const { return h_; }
 private:
  void reset() { yv_svo_free(h_); h_ = nullptr; }
  yv_svo *h_;
};

// ISVORenderer (cell/svorenderer.h:5-24)
class ISVORenderer {
 protected:
  ISVORenderer() {}
 public:
  virtual ~ISVORenderer() {}
  virtual void SetScene(SVOData *svo) = 0;
  virtual void SetViewPos(const point_3f &pos) = 0;
  virtual void SetViewDir(const point_3f &dir) = 0;
  virtual void SetViewUp(const point_3f &up) = 0;
  virtual void SetResolution(int width, int height) = 0;
  virtual point_2i GetResolution() const = 0;
  virtual void SetFOV(float fov) = 0;
  virtual const Color32 *RenderFrame() = 0;
};
}  // namespace yv
#define YV_NS yv::
#else
#define YV_NS
#endif

namespace yv {

class B200Renderer : public YV_NS ISVORenderer {
 public:
  explicit B200Renderer(int device = 0) : r_(nullptr), own_(nullptr) { yv_renderer_create(device, &r_); }
  // SPURenderer's "every usable SPE" (cell/spu_renderer.cpp:73): one renderer over the GPUs of `device_mask`
  // (YV_ALL_DEVICES = all of them); each RenderFrame() splits the frame's blocks over them (:80-83)
  struct Devices { uint64_t mask; };
  explicit B200Renderer(Devices d) : r_(nullptr), own_(nullptr) { yv_renderer_create_multi(d.mask, &r_); }
  explicit B200Renderer(const std::vector<int> &ordinals) : r_(nullptr), own_(nullptr) {
    if (!ordinals.empty()) yv_renderer_create_group(ordinals.data(), (int)ordinals.size(), &r_);
  }
  int DeviceCount() const { return r_ ? yv_renderer_device_count(r_) : 0; }
  ~B200Renderer() override { yv_renderer_destroy(r_); yv_svo_free(own_); }
  bool ok() const { return r_ != nullptr; }

#ifndef YV_USE_REFERENCE_TYPES
  void SetScene(YV_NS SVOData *svo) override { if (r_) yv_set_scene(r_, svo ? svo->handle() : nullptr); }
#else
  // The tree's SVOData (cell/svodata.h:22-55) owns a host pool of 40-byte VoxNodes — the layout of yv_vox_node — and
  // exposes GetRoot() and operator[] but not its size, so the pool handed to the library ends at the highest node id
  // reachable from the root. The scene pointer stays borrowed (renderer_base.h:28); the library keeps its own copy.
  void SetScene(SVOData *svo) override {
    if (r_) yv_set_scene(r_, nullptr);
    yv_svo_free(own_); own_ = nullptr;
    if (svo) {
      const VoxNodeId root = svo->GetRoot();
      uint32_t count = 0;
      if (!IsNull(root)) {
        std::vector<VoxNodeId> todo(1, root);
        std::vector<bool> seen;
        while (!todo.empty()) {
          const VoxNodeId id = todo.back(); todo.pop_back();
          if (id >= seen.size()) seen.resize(id + 1 > 2 * seen.size() ? id + 1 : 2 * seen.size(), false);
          if (seen[id]) continue;
          seen[id] = true;
          if (id + 1 > count) count = id + 1;
          const VoxNode &n = (*svo)[id];
          for (int c = 0; c < 8; ++c)
            if (!GetLeafFlag(n.flags, c) && !IsNull(n.child[c])) todo.push_back(n.child[c]);
        }
      }
      yv_svo_from_memory(root, count ? reinterpret_cast<const yv_vox_node *>(&(*svo)[0]) : nullptr, count, &own_);
    }
    if (r_) yv_set_scene(r_, own_);
  }
#endif
  void SetViewPos(const YV_NS point_3f &p) override { const float v[3] = { p.x, p.y, p.z }; if (r_) yv_set_view_pos(r_, v); }
  void SetViewDir(const YV_NS point_3f &p) override { const float v[3] = { p.x, p.y, p.z }; if (r_) yv_set_view_dir(r_, v); }
  void SetViewUp(const YV_NS point_3f &p) override { const float v[3] = { p.x, p.y, p.z }; if (r_) yv_set_view_up(r_, v); }
  void SetResolution(int w, int h) override { if (r_) yv_set_resolution(r_, w, h); }
  YV_NS point_2i GetResolution() const override { int w = 0, h = 0; if (r_) yv_get_resolution(r_, &w, &h); return YV_NS point_2i(w, h); }
  void SetFOV(float fov) override { if (r_) yv_set_fov(r_, fov); }
  // NULL when no scene is set (cell/ppu_renderer.cpp:78-79) or when rendering is impossible
  const YV_NS Color32 *RenderFrame() override {
    const uint8_t *px = nullptr;
    if (!r_ || yv_render_frame(r_, &px) != 0) return nullptr;
    return reinterpret_cast<const YV_NS Color32 *>(px);
  }
  // SVORenderer::Render(void* d_dstBuf) (demo/SVORenderer.h:36)
  bool Render(void *d_dst) { return r_ && yv_render_frame_device(r_, d_dst) == 0; }
  // the remaining setters of the CUDA demo's SVORenderer (demo/SVORenderer.h:20-38)
  void SetViewSize(int w, int h) { SetResolution(w, h); }                                        // :20
  float GetFOV() const { float f = 0; if (r_) yv_get_fov(r_, &f); return f; }                    // :23
  void SetDetailCoef(float coef) { if (r_) yv_set_detail_coef(r_, coef); }                       // :25
  float GetDetailCoef() const { float c = 0; if (r_) yv_get_detail_coef(r_, &c); return c; }     // :26
  void SetSSNA(bool enable) { if (r_) yv_set_ssna(r_, enable ? 1 : 0); }                         // :28
  bool GetSSNA() const { int v = 0; if (r_) yv_get_ssna(r_, &v); return v != 0; }                // :29
  void SetShowNormals(bool enable) { if (r_) yv_set_show_normals(r_, enable ? 1 : 0); }          // :31
  bool GetShowNormals() const { int v = 0; if (r_) yv_get_show_normals(r_, &v); return v != 0; } // :32
  void SetLigth(int i, const yv_light &lp) { if (r_) yv_set_light(r_, i, &lp); }                 // :34 (reference spelling)
  void DumpTraceData(const std::string &fnbase) { if (r_) yv_dump_trace_data(r_, fnbase.c_str()); }   // :38
  float LastFrameMs() const { return r_ ? yv_last_frame_ms(r_) : -1.0f; }
  yv_renderer *handle() const { return r_; }

 private:
  yv_renderer *r_;
  yv_svo *own_;           // reference-types mode: the library-side copy of the tree's pool
};

#ifndef YV_USE_REFERENCE_TYPES
// CreateSimpleRenderer / CreateThreadedRenderer / CreateSPURenderer (cell/svorenderer.h:26-30)
inline std::shared_ptr<YV_NS ISVORenderer> CreateB200Renderer(int device = 0) {
  return std::shared_ptr<YV_NS ISVORenderer>(new B200Renderer(device));
}
// all GPUs of `device_mask` behind one ISVORenderer (CreateSPURenderer's role on a multi-accelerator machine)
inline std::shared_ptr<YV_NS ISVORenderer> CreateB200MultiRenderer(uint64_t device_mask = YV_ALL_DEVICES) {
  return std::shared_ptr<YV_NS ISVORenderer>(new B200Renderer(B200Renderer::Devices{ device_mask }));
}
#endif     // in the tree the factory is written with the tree's own shared_ptr (integration/b200_renderer.cpp)

}  // namespace yv
