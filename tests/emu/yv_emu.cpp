// yv_emu.cpp — HOST build of the kernel's per-ray code (yoxel-voxel_b200/csrc/trace_core.cuh).
// TEST INFRASTRUCTURE ONLY: lets the explicit-stack state machine, the packed-record addressing
// and the shading helpers be checked against the oracle in a container without a GPU. It is never
// linked into libyv_b200.so and no product path can reach it.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../yoxel-voxel_b200/csrc/trace_core.cuh"

using namespace yv;

namespace {
float g_detail = 0.0f;   // rp.detailCoef; > 0 switches the LOD cut-off on (lean mode only)
bool g_lod_hit = false;
const uint32_t *g_node_data = nullptr;
uint64_t g_trips = 0;   // lean_step / trace_step calls of the last yve_render
int g_mode = 2;   // 0 = trace_step (classic form), 2 = lean_step, 3 = lean_step with octant culling, 4 = lean_step with
                  // lean_descend_once in front of every secondary ray (what render_frame<SEC> runs)
uint64_t g_fast_levels = 0;   // levels lean_descend_once took over in the last yve_render
struct HostStack {
  StackEntry e[kMaxStack];
  int max_sp = 0;
  void push(int sp, const StackEntry &en) { e[sp] = en; if (sp + 1 > max_sp) max_sp = sp + 1; }
  StackEntry pop(int sp) const { return e[sp]; }
};
struct LeanHostStack {
  U4 a[kMaxStack], b[kMaxStack];
  int max_sp = 0;
  void push(int sp, const U4 &x, const U4 &y) { a[sp] = x; b[sp] = y; if (sp + 1 > max_sp) max_sp = sp + 1; }
  void pop(int sp, U4 &x, U4 &y) const { x = a[sp]; y = b[sp]; }
};
LeanHostStack g_lean_stack;
struct HostFetch {
  const Rec *recs;
  mutable uint64_t fetches = 0, visits = 0;
  Rec operator()(uint32_t idx) const { ++fetches; ++visits; return recs[idx]; }
  Rec get(uint32_t idx, bool visit) const { ++fetches; if (visit) ++visits; return recs[idx]; }
  // layout policy used by the lean form (packed pool)
  void node(uint32_t idx, bool visit, uint32_t &masks, uint32_t &child_base) const {
    const Rec r = get(idx, visit); masks = r.masks; child_base = r.child_base;
  }
  uint32_t child_index(uint32_t, uint32_t child_base, uint32_t masks, uint32_t c) const {
    return child_base + (uint32_t)YV_POPC((masks >> 8) & ((1u << c) - 1u));
  }
  uint32_t root_index() const { return 0u; }
};

// the culling policy: the node's grandchild mask is put together from its children's records here (the product's
// repack stores it in the record; tests/test_gpu_pack.py checks that array against the same definition)
struct HostCullFetch : HostFetch {
  void node(uint32_t idx, bool visit, uint32_t &masks, uint32_t &child_base, uint32_t &gm_lo, uint32_t &gm_hi) const {
    const Rec r = get(idx, visit); masks = r.masks; child_base = r.child_base;
    uint64_t g = 0; uint32_t k = 0;
    for (uint32_t c = 0; c < 8; ++c)
      if ((r.masks >> (8 + c)) & 1u) { const uint32_t m = recs[r.child_base + k++].masks; g |= (uint64_t)((m | (m >> 8)) & 0xffu) << (8 * c); }
    gm_lo = (uint32_t)g; gm_hi = (uint32_t)(g >> 32);
  }
  uint32_t box(uint32_t flags, uint32_t ch0, uint32_t chx) const { return box_mask_stored(flags, ch0, chx); }
};
}  // namespace
namespace yv { template <> struct FetchTraits<HostCullFetch> { static constexpr bool kGuardDepth = false; static constexpr bool kCull = true; }; }
namespace {

template <class Fetch>
bool trace(const Fetch &fetch, bool root_valid, HostStack &stk, float ox, float oy, float oz,
           float dx, float dy, float dz, bool front_only, RayState &s, Rec &rec, uint64_t &steps,
           float tlimit = __builtin_huge_valf()) {
  dx = adjust_dir1(dx); dy = adjust_dir1(dy); dz = adjust_dir1(dz);
  if (g_mode >= 2) {
    LeanState ls;
    if (!lean_begin(ls, fetch, root_valid, ox, oy, oz, dx, dy, dz)) return false;
    ls.tlimit = tlimit;
    if ((g_mode == 4 && front_only) || g_mode == 5) {        // 5: in front of primary rays too
      for (;;) {
        const bool more = g_detail > 0.0f ? lean_descend_once<true>(ls, fetch, g_lean_stack, front_only)
                                          : lean_descend_once<false>(ls, fetch, g_lean_stack, front_only);
        if (!more) break;
        ++g_fast_levels;
      }
    }
    for (;;) {
      ++steps;
      const int r = g_detail > 0.0f ? lean_step<true>(ls, fetch, g_lean_stack, front_only, g_detail)
                                    : lean_step<false>(ls, fetch, g_lean_stack, front_only);
      if (r == kStepContinue) continue;
      if (g_lean_stack.max_sp > stk.max_sp) stk.max_sp = g_lean_stack.max_sp;
      if (r == kStepMiss) return false;
      g_lod_hit = r == kStepLodHit;
      // present the result through the classic state the caller reads
      s.t1x = lean_t1x(ls); s.t1y = lean_t1y(ls); s.t1z = lean_t1z(ls); s.ch = lean_ch(ls); s.flags = ls.flags; s.idx = ls.idx;
      rec = fetch.recs[ls.idx];
      return true;
    }
  }
  if (!setup_trace(ox, oy, oz, dx, dy, dz, s)) return false;
  if (!trace_enter_root(s, rec, fetch, root_valid)) return false;
  for (;;) {
    ++steps;
    int r = trace_step(s, rec, fetch, stk, front_only);
    if (r == kStepHit) return true;
    if (r == kStepMiss) return false;
  }
}
}  // namespace

extern "C" void yve_set_mode(int mode) { g_mode = mode; }
extern "C" uint64_t yve_trips() { return g_trips; }      // lean_step / trace_step calls of the last yve_render
extern "C" uint64_t yve_fast_levels() { return g_fast_levels; }
extern "C" void yve_set_lod(float detail, const uint32_t *node_data) { g_detail = detail; g_node_data = node_data; }

template <class Fetch>
int render_with(const Fetch &fetch, const uint32_t *leaves, int root_valid,
                const float pos[3], const float dir0[3], const float du[3], const float dv[3],
                const float light[3], int width, int height,
                int shadow, int ao_samples, uint32_t seed, float voxel_size, float ao_max_t,
                uint32_t *hit_node, int32_t *hit_child, float *hit_t, uint32_t *rgba,
                uint64_t *out_fetches, int *out_max_sp, uint64_t *out_visits) {
  HostStack stk;
  uint64_t steps = 0;
  g_fast_levels = 0;
  const bool sec = shadow || ao_samples > 0;
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) {
      const uint32_t pixel = (uint32_t)y * (uint32_t)width + (uint32_t)x;
      float dx, dy, dz;
      primary_dir(dir0, du, dv, x, y, dx, dy, dz);
      dx = adjust_dir1(dx); dy = adjust_dir1(dy); dz = adjust_dir1(dz);
      RayState s; Rec rec;
      uint32_t out = 0, hn = YV_MISS_NODE; int32_t hc = YV_MISS_CHILD; float ht = 0.0f;
      g_lod_hit = false;
      if (trace(fetch, root_valid != 0, stk, pos[0], pos[1], pos[2], dx, dy, dz, false, s, rec, steps)) {
        const uint32_t c = s.ch ^ s.flags;
        const bool lod = g_lod_hit;
        hn = rec.orig_id; hc = lod ? -1 : (int32_t)c; ht = max3f(s.t1x, s.t1y, s.t1z);
        const uint32_t data = lod ? g_node_data[s.idx]
                                  : leaves[rec.leaf_base + (uint32_t)YV_POPC(rec.masks & 0xffu & ((1u << c) - 1u))];
        float nx, ny, nz;
        unpack_normal(data, nx, ny, nz);
        const float Px = YV_FADD(pos[0], YV_FMUL(dx, ht)), Py = YV_FADD(pos[1], YV_FMUL(dy, ht)), Pz = YV_FADD(pos[2], YV_FMUL(dz, ht));
        const float dl = lambert(nx, ny, nz, Px, Py, Pz, light[0], light[1], light[2]);
        float vis = 1.0f, ao = 1.0f;
        if (sec) {
          const float Ox = YV_FADD(Px, YV_FMUL(nx, voxel_size)), Oy = YV_FADD(Py, YV_FMUL(ny, voxel_size)), Oz = YV_FADD(Pz, YV_FMUL(nz, voxel_size));
          RayState s2; Rec r2;
          if (shadow) {
            const float vx = YV_FSUB(light[0], Ox), vy = YV_FSUB(light[1], Oy), vz = YV_FSUB(light[2], Oz);
            const float len = YV_FSQRT(YV_FADD(YV_FADD(YV_FMUL(vx, vx), YV_FMUL(vy, vy)), YV_FMUL(vz, vz)));
            if (len > 0 && trace(fetch, root_valid != 0, stk, Ox, Oy, Oz, YV_FDIV(vx, len), YV_FDIV(vy, len), YV_FDIV(vz, len), true, s2, r2, steps, len)) {
              const float ts = max3f(s2.t1x, s2.t1y, s2.t1z);
              if (ts > 0 && ts < len) vis = 0.0f;
            }
          }
          if (ao_samples > 0) {
            int occ = 0;
            for (int k = 0; k < ao_samples; ++k) {
              float ax, ay, az;
              ao_direction(nx, ny, nz, pixel, (uint32_t)k, seed, ax, ay, az);
              if (trace(fetch, root_valid != 0, stk, Ox, Oy, Oz, ax, ay, az, true, s2, r2, steps, ao_max_t)) {
                const float ts = max3f(s2.t1x, s2.t1y, s2.t1z);
                if (ts > 0 && ts < ao_max_t) ++occ;
              }
            }
            ao = YV_FSUB(1.0f, YV_FDIV((float)occ, (float)ao_samples));
          }
        }
        const float k = sec ? YV_FMUL(YV_FADD(YV_SHADE_AMBIENT, YV_FMUL(YV_SHADE_DIFFUSE, YV_FMUL(dl, vis))), ao)
                            : YV_FADD(YV_SHADE_AMBIENT, YV_FMUL(YV_SHADE_DIFFUSE, YV_FMUL(dl, 1.0f)));
        out = shade_rgba(data, k);
      }
      if (hit_node) hit_node[pixel] = hn;
      if (hit_child) hit_child[pixel] = hc;
      if (hit_t) hit_t[pixel] = ht;
      if (rgba) rgba[pixel] = out;
    }
  g_trips = steps;
  if (out_fetches) *out_fetches = fetch.fetches;
  if (out_visits) *out_visits = fetch.visits;
  if (out_max_sp) *out_max_sp = stk.max_sp;
  return 0;
}

extern "C" int yve_render(const uint32_t *records, const uint32_t *leaves, int root_valid,
                          const float pos[3], const float dir0[3], const float du[3], const float dv[3],
                          const float light[3], int width, int height,
                          int shadow, int ao_samples, uint32_t seed, float voxel_size, float ao_max_t,
                          uint32_t *hit_node, int32_t *hit_child, float *hit_t, uint32_t *rgba,
                          uint64_t *out_fetches, int *out_max_sp, uint64_t *out_visits) {
  const Rec *recs = reinterpret_cast<const Rec *>(records);
  if (g_mode == 3) {
    HostCullFetch fetch; fetch.recs = recs;
    return render_with(fetch, leaves, root_valid, pos, dir0, du, dv, light, width, height, shadow, ao_samples, seed, voxel_size,
                       ao_max_t, hit_node, hit_child, hit_t, rgba, out_fetches, out_max_sp, out_visits);
  }
  HostFetch fetch{ recs };
  return render_with(fetch, leaves, root_valid, pos, dir0, du, dv, light, width, height, shadow, ao_samples, seed, voxel_size,
                     ao_max_t, hit_node, hit_child, hit_t, rgba, out_fetches, out_max_sp, out_visits);
}
