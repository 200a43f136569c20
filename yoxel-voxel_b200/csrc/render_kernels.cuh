// render_kernels.cuh — sm_100a kernels of the SVO ray caster.
//
// One fused kernel per frame: per-pixel ray generation (cell/ppu_renderer.cpp:49-57), slab set-up,
// octree descent over the packed 16-byte records (RecTrace, :18-41), leaf shading (:67) and the
// RGBA8 store (:54,:67) — the reference CUDA path's InitEyeRays / Trace / ShadeSimple sequence
// (demo/SVORenderer.cpp:95-149, trace_cuda.py:20-23) collapsed into one launch.
//
// Two schedules share the per-ray code in trace_core.cuh:
//   render_tiles       one CTA per 16x8 pixel tile, one thread per pixel (8x4 pixels per warp)
//   render_persistent  resident CTAs; every warp pulls 8x8 pixel tiles from an atomic counter and
//                      re-fills idle lanes with new rays (ballot + popc prefix) once fewer than
//                      kRefillThreshold lanes are still traversing; secondary rays (shadow, AO)
//                      re-enter the same traversal loop as further stages of the pixel's state
//                      machine instead of running as divergent tails.
// The top `smem_nodes` records of the breadth-first pool are staged in shared memory per CTA
// (precedent: the SPU's software node cache, cell/spu/trace_spu.cpp:15-35).
#pragma once

#include <cuda_runtime.h>

#include "trace_core.cuh"

namespace yv {

struct RenderParams {
  const uint4 *recs;           // packed records (svo_pack.h)
  const uint32_t *leaves;      // inline VoxData words
  uint32_t root_valid;
  uint32_t smem_nodes;         // records staged in shared memory (<= record count)
  float pos[3];                // eye (m_pos) — also shader viewerPos (renderer_base.h:30-35)
  float dir0[3], du[3], dv[3]; // RayDirData (renderer_base.h:50-61), computed on the host
  float light[3];              // shader lightPos
  int width, height;           // m_viewSize
  int y0, y1;                  // row band rendered by this launch
  uint32_t *out_rgba;          // full-frame addressed: out_rgba[y*width + x]
  uint32_t *hit_node;          // optional TraceResult planes (ppu_renderer.cpp:7-12)
  int32_t *hit_child;
  float *hit_t;
  uint32_t *counters;          // optional: low 16 bits node visits, high 16 bits pop re-fetches
  unsigned int *tile_counter;  // persistent schedule: next tile to hand out
  int tiles_x, num_tiles;      // 8x8 tiles covering [0,width) x [y0,y1)
  int shadow, ao_samples;      // secondary rays
  uint32_t seed;
  float voxel_size, ao_max_t;
};

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kRefillThreshold = 20;   // refill once <= this many lanes are still traversing

// per-thread traversal stack in local memory: two 16-byte words per entry (LDL.128 / STL.128)
struct LocalStack {
  uint4 w[2 * kMaxStack];
  __device__ __forceinline__ void push(int sp, const StackEntry &e) {
    w[2 * sp]     = make_uint4(__float_as_uint(e.t1x), __float_as_uint(e.t1y), __float_as_uint(e.t1z), e.idx);
    w[2 * sp + 1] = make_uint4(__float_as_uint(e.t2x), __float_as_uint(e.t2y), __float_as_uint(e.t2z), e.ch);
  }
  __device__ __forceinline__ StackEntry pop(int sp) const {
    const uint4 a = w[2 * sp], b = w[2 * sp + 1];
    StackEntry e = { __uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), a.w,
                     __uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), b.w };
    return e;
  }
};

// node fetch: shared memory for the staged top of the tree, read-only global path otherwise
template <bool COUNT>
struct NodeFetch {
  const uint4 *recs;
  const uint4 *staged;
  uint32_t staged_n;
  mutable uint32_t visits, revisits;
  __device__ __forceinline__ Rec load(uint32_t idx) const {
    uint4 v;
    if (idx < staged_n) v = staged[idx];
    else v = __ldg(recs + idx);
    Rec r = { v.x, v.y, v.z, v.w };
    return r;
  }
  __device__ __forceinline__ Rec operator()(uint32_t idx) const { if (COUNT) ++visits; return load(idx); }
};

// trace_step's pop path calls fetch(idx) as well; to separate algorithmic visits from re-fetches
// the counter variant tracks stack depth changes outside (see trace_to_end).

__device__ __forceinline__ void stage_top_records(const RenderParams &p, uint4 *staged) {
  for (uint32_t i = threadIdx.x; i < p.smem_nodes; i += blockDim.x) staged[i] = __ldg(p.recs + i);
  __syncthreads();
}

// Run one ray to completion (used by the tile schedule and by yv_trace_rays).
template <bool FRONT_ONLY, bool COUNT>
__device__ __forceinline__ bool trace_to_end(const NodeFetch<COUNT> &fetch, bool root_valid, LocalStack &stk,
                                             float ox, float oy, float oz, float dx, float dy, float dz,
                                             RayState &s, Rec &rec, uint32_t &pops) {
  dx = adjust_dir1(dx); dy = adjust_dir1(dy); dz = adjust_dir1(dz);
  if (!setup_trace(ox, oy, oz, dx, dy, dz, s)) return false;
  if (!trace_enter_root(s, rec, fetch, root_valid)) return false;
  for (;;) {
    const int sp0 = s.sp;
    const int r = trace_step(s, rec, fetch, stk, FRONT_ONLY);
    if (COUNT && s.sp < sp0) ++pops;
    if (r == kStepHit) return true;
    if (r == kStepMiss) return false;
  }
}

__device__ __forceinline__ uint32_t leaf_data(const RenderParams &p, const Rec &rec, uint32_t c) {
  return __ldg(p.leaves + rec.leaf_base + (uint32_t)__popc(rec.masks & 0xffu & ((1u << c) - 1u)));
}

// Shade one pixel from its primary hit, tracing the secondary rays in-thread (tile schedule).
template <bool SEC, bool COUNT>
__device__ __forceinline__ uint32_t shade_hit(const RenderParams &p, const NodeFetch<COUNT> &fetch, LocalStack &stk,
                                              uint32_t pixel, uint32_t data, float dx, float dy, float dz, float t,
                                              uint32_t &pops) {
  float nx, ny, nz;
  unpack_normal(data, nx, ny, nz);
  const float Px = YV_FADD(p.pos[0], YV_FMUL(dx, t));
  const float Py = YV_FADD(p.pos[1], YV_FMUL(dy, t));
  const float Pz = YV_FADD(p.pos[2], YV_FMUL(dz, t));
  const float dl = lambert(nx, ny, nz, Px, Py, Pz, p.light[0], p.light[1], p.light[2]);
  if (!SEC) {
    const float k = YV_FADD(YV_SHADE_AMBIENT, YV_FMUL(YV_SHADE_DIFFUSE, YV_FMUL(dl, 1.0f)));
    return shade_rgba(data, k);
  }
  const float Ox = YV_FADD(Px, YV_FMUL(nx, p.voxel_size));
  const float Oy = YV_FADD(Py, YV_FMUL(ny, p.voxel_size));
  const float Oz = YV_FADD(Pz, YV_FMUL(nz, p.voxel_size));
  float vis = 1.0f;
  RayState s; Rec rec;
  if (p.shadow) {
    const float vx = YV_FSUB(p.light[0], Ox), vy = YV_FSUB(p.light[1], Oy), vz = YV_FSUB(p.light[2], Oz);
    const float len = YV_FSQRT(YV_FADD(YV_FADD(YV_FMUL(vx, vx), YV_FMUL(vy, vy)), YV_FMUL(vz, vz)));
    if (len > 0) {
      const bool h = trace_to_end<true, COUNT>(fetch, p.root_valid != 0u, stk, Ox, Oy, Oz,
                                               YV_FDIV(vx, len), YV_FDIV(vy, len), YV_FDIV(vz, len), s, rec, pops);
      if (h) { const float ts = max3f(s.t1x, s.t1y, s.t1z); if (ts > 0 && ts < len) vis = 0.0f; }
    }
  }
  float ao = 1.0f;
  if (p.ao_samples > 0) {
    int occ = 0;
    for (int smp = 0; smp < p.ao_samples; ++smp) {
      float ax, ay, az;
      ao_direction(nx, ny, nz, pixel, (uint32_t)smp, p.seed, ax, ay, az);
      const bool h = trace_to_end<true, COUNT>(fetch, p.root_valid != 0u, stk, Ox, Oy, Oz, ax, ay, az, s, rec, pops);
      if (h) { const float ts = max3f(s.t1x, s.t1y, s.t1z); if (ts > 0 && ts < p.ao_max_t) ++occ; }
    }
    ao = YV_FSUB(1.0f, YV_FDIV((float)occ, (float)p.ao_samples));
  }
  const float k = YV_FMUL(YV_FADD(YV_SHADE_AMBIENT, YV_FMUL(YV_SHADE_DIFFUSE, YV_FMUL(dl, vis))), ao);
  return shade_rgba(data, k);
}

// ---------------------------------------------------------------------------------------------
// schedule 1: one CTA per 16x8 tile, one thread per pixel
// ---------------------------------------------------------------------------------------------
template <bool HITS, bool SEC, bool COUNT>
__global__ void __launch_bounds__(128) render_tiles(const __grid_constant__ RenderParams p) {
  extern __shared__ uint4 staged[];
  stage_top_records(p, staged);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles_x16 = (p.width + 15) >> 4;
  const int tx = blockIdx.x % tiles_x16, ty = blockIdx.x / tiles_x16;
  const int x = tx * 16 + (warp & 1) * 8 + (lane & 7);
  const int y = p.y0 + ty * 8 + (warp >> 1) * 4 + (lane >> 3);
  if (x >= p.width || y >= p.y1) return;
  const uint32_t pixel = (uint32_t)y * (uint32_t)p.width + (uint32_t)x;

  NodeFetch<COUNT> fetch = { p.recs, staged, p.smem_nodes, 0u, 0u };
  LocalStack stk;
  RayState s; Rec rec;
  uint32_t pops = 0;

  float dx, dy, dz;
  primary_dir(p.dir0, p.du, p.dv, x, y, dx, dy, dz);
  dx = adjust_dir1(dx); dy = adjust_dir1(dy); dz = adjust_dir1(dz);
  const bool hit = trace_to_end<false, COUNT>(fetch, p.root_valid != 0u, stk, p.pos[0], p.pos[1], p.pos[2],
                                              dx, dy, dz, s, rec, pops);
  uint32_t rgba = 0u;
  uint32_t hn = YV_MISS_NODE; int32_t hc = YV_MISS_CHILD; float ht = 0.0f;
  if (hit) {
    const uint32_t c = s.ch ^ s.flags;
    hn = rec.orig_id; hc = (int32_t)c; ht = max3f(s.t1x, s.t1y, s.t1z);
    const uint32_t data = leaf_data(p, rec, c);
    rgba = shade_hit<SEC, COUNT>(p, fetch, stk, pixel, data, dx, dy, dz, ht, pops);
  }
  p.out_rgba[pixel] = rgba;
  if (HITS) { p.hit_node[pixel] = hn; p.hit_child[pixel] = hc; p.hit_t[pixel] = ht; }
  if (COUNT) p.counters[pixel] = ((fetch.visits - pops) & 0xffffu) | (pops << 16);
}

// ---------------------------------------------------------------------------------------------
// schedule 2: persistent warps, atomic tile queue, lane refill, secondary rays as stages
// ---------------------------------------------------------------------------------------------
enum : int { kLaneIdle = 0, kLaneActive = 1, kLaneHit = 2, kLaneMiss = 3, kLaneNew = 4 };

template <bool HITS, bool SEC, bool COUNT, int THREADS>
__global__ void __launch_bounds__(THREADS) render_persistent(const __grid_constant__ RenderParams p) {
  extern __shared__ uint4 staged[];
  stage_top_records(p, staged);

  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  NodeFetch<COUNT> fetch = { p.recs, staged, p.smem_nodes, 0u, 0u };
  LocalStack stk;
  RayState s; Rec rec;
  uint32_t pops = 0;

  // warp-uniform tile pool
  constexpr int kTilePix = 64;            // 8x8 pixels, handed out in Morton order
  int pool_next = kTilePix, tile_x0 = 0, tile_y0 = 0;
  bool pool_empty = false;

  // per-lane pixel state
  int state = kLaneIdle;
  int x = 0, y = 0;
  uint32_t pixel = 0;
  float dx = 0.f, dy = 0.f, dz = 0.f;     // direction of the ray being traversed
  // secondary-ray stage machine (SEC only): stage 0 = primary, 1 = shadow, 2.. = AO samples
  int stage = 0;
  uint32_t sdata = 0; float nx = 0.f, ny = 0.f, nz = 0.f, Ox = 0.f, Oy = 0.f, Oz = 0.f;
  float dl = 0.f, vis = 1.f, slen = 0.f; int occ = 0;

  for (;;) {
    // ---- 1. hand new pixels to idle lanes --------------------------------------------------
    unsigned idle = __ballot_sync(kFullMask, state == kLaneIdle);
    while (idle != 0u && !pool_empty) {
      if (pool_next >= kTilePix) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(p.tile_counter, 1u);
        t = __shfl_sync(kFullMask, t, 0);
        if (t >= (unsigned)p.num_tiles) { pool_empty = true; break; }
        tile_x0 = (int)(t % (unsigned)p.tiles_x) * 8;
        tile_y0 = p.y0 + (int)(t / (unsigned)p.tiles_x) * 8;
        pool_next = 0;
      }
      const int take = min(__popc(idle), kTilePix - pool_next);
      const int rank = __popc(idle & lt_mask);
      if (state == kLaneIdle && rank < take) {
        const int m = pool_next + rank;   // Morton index inside the tile
        x = tile_x0 + ((m & 1) | ((m >> 1) & 2) | ((m >> 2) & 4));
        y = tile_y0 + (((m >> 1) & 1) | ((m >> 2) & 2) | ((m >> 3) & 4));
        if (x < p.width && y < p.y1) state = kLaneNew;
      }
      pool_next += take;
      idle = __ballot_sync(kFullMask, state == kLaneIdle);
    }

    // ---- 2. primary ray set-up for the new lanes --------------------------------------------
    if (state == kLaneNew) {
      pixel = (uint32_t)y * (uint32_t)p.width + (uint32_t)x;
      if (COUNT) { fetch.visits = 0; pops = 0; }
      primary_dir(p.dir0, p.du, p.dv, x, y, dx, dy, dz);
      dx = adjust_dir1(dx); dy = adjust_dir1(dy); dz = adjust_dir1(dz);
      stage = 0;
      state = kLaneMiss;
      if (setup_trace(p.pos[0], p.pos[1], p.pos[2], dx, dy, dz, s) &&
          trace_enter_root(s, rec, fetch, p.root_valid != 0u))
        state = kLaneActive;
    }

    // ---- 3. traversal: all lanes step in lock-step until too few are left --------------------
    for (;;) {
      const unsigned am = __ballot_sync(kFullMask, state == kLaneActive);
      if (am == 0u) break;
      if (!pool_empty && __popc(am) <= kRefillThreshold) break;
      if (state == kLaneActive) {
        const int sp0 = s.sp;
        const int r = trace_step(s, rec, fetch, stk, SEC && stage > 0);
        if (COUNT && s.sp < sp0) ++pops;
        if (r == kStepHit) state = kLaneHit;
        else if (r == kStepMiss) state = kLaneMiss;
      }
    }

    // ---- 4. finished rays: shade / spawn the next secondary ray / store ----------------------
    if (state == kLaneHit || state == kLaneMiss) {
      const bool hit = state == kLaneHit;
      bool done = true;
      uint32_t rgba = 0u;
      if (!SEC || stage == 0) {
        uint32_t hn = YV_MISS_NODE; int32_t hc = YV_MISS_CHILD; float ht = 0.0f;
        if (hit) {
          const uint32_t c = s.ch ^ s.flags;
          hn = rec.orig_id; hc = (int32_t)c; ht = max3f(s.t1x, s.t1y, s.t1z);
          sdata = leaf_data(p, rec, c);
          unpack_normal(sdata, nx, ny, nz);
          const float Px = YV_FADD(p.pos[0], YV_FMUL(dx, ht));
          const float Py = YV_FADD(p.pos[1], YV_FMUL(dy, ht));
          const float Pz = YV_FADD(p.pos[2], YV_FMUL(dz, ht));
          dl = lambert(nx, ny, nz, Px, Py, Pz, p.light[0], p.light[1], p.light[2]);
          if (!SEC) {
            rgba = shade_rgba(sdata, YV_FADD(YV_SHADE_AMBIENT, YV_FMUL(YV_SHADE_DIFFUSE, YV_FMUL(dl, 1.0f))));
          } else {
            Ox = YV_FADD(Px, YV_FMUL(nx, p.voxel_size));
            Oy = YV_FADD(Py, YV_FMUL(ny, p.voxel_size));
            Oz = YV_FADD(Pz, YV_FMUL(nz, p.voxel_size));
            vis = 1.0f; occ = 0;
            done = false;           // secondary stages follow
          }
        }
        if (HITS) { p.hit_node[pixel] = hn; p.hit_child[pixel] = hc; p.hit_t[pixel] = ht; }
      } else {
        // a secondary ray came back
        const float ts = max3f(s.t1x, s.t1y, s.t1z);
        if (stage == 1) { if (hit && ts > 0 && ts < slen) vis = 0.0f; }
        else if (hit && ts > 0 && ts < p.ao_max_t) ++occ;
        done = false;
      }
      if (SEC && !done) {
        // pick the next secondary ray of this pixel, or finish
        const int last_stage = 1 + p.ao_samples;
        bool launched = false;
        while (!launched && stage < last_stage) {
          ++stage;
          float rx, ry, rz;
          if (stage == 1) {
            if (!p.shadow) continue;
            const float vx = YV_FSUB(p.light[0], Ox), vy = YV_FSUB(p.light[1], Oy), vz = YV_FSUB(p.light[2], Oz);
            slen = YV_FSQRT(YV_FADD(YV_FADD(YV_FMUL(vx, vx), YV_FMUL(vy, vy)), YV_FMUL(vz, vz)));
            if (!(slen > 0)) continue;
            rx = YV_FDIV(vx, slen); ry = YV_FDIV(vy, slen); rz = YV_FDIV(vz, slen);
          } else {
            ao_direction(nx, ny, nz, pixel, (uint32_t)(stage - 2), p.seed, rx, ry, rz);
          }
          rx = adjust_dir1(rx); ry = adjust_dir1(ry); rz = adjust_dir1(rz);
          if (setup_trace(Ox, Oy, Oz, rx, ry, rz, s) && trace_enter_root(s, rec, fetch, p.root_valid != 0u))
            launched = true;          // otherwise this secondary ray misses outright: unoccluded
        }
        if (launched) state = kLaneActive;
        else {
          float ao = 1.0f;
          if (p.ao_samples > 0) ao = YV_FSUB(1.0f, YV_FDIV((float)occ, (float)p.ao_samples));
          const float k = YV_FMUL(YV_FADD(YV_SHADE_AMBIENT, YV_FMUL(YV_SHADE_DIFFUSE, YV_FMUL(dl, vis))), ao);
          rgba = shade_rgba(sdata, k);
          done = true;
        }
      }
      if (done) {
        p.out_rgba[pixel] = rgba;
        if (COUNT) p.counters[pixel] = ((fetch.visits - pops) & 0xffffu) | (pops << 16);
        state = kLaneIdle;
      }
    }

    // ---- 5. the warp retires when the queue is dry and every lane is idle --------------------
    if (pool_empty && __ballot_sync(kFullMask, state != kLaneIdle) == 0u) break;
  }
}

// ---------------------------------------------------------------------------------------------
// arbitrary rays (DynamicSVO::TraceRay, ore/src/main.cpp:125)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) trace_rays_kernel(const uint4 *recs, uint32_t root_valid,
                                                         const float *pos, const float *dir, uint32_t count,
                                                         uint32_t *node, int32_t *child, float *t) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  NodeFetch<false> fetch = { recs, nullptr, 0u, 0u, 0u };
  LocalStack stk;
  RayState s; Rec rec;
  uint32_t pops = 0;
  const bool hit = trace_to_end<false, false>(fetch, root_valid != 0u, stk, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2],
                                              dir[3 * i], dir[3 * i + 1], dir[3 * i + 2], s, rec, pops);
  node[i] = hit ? rec.orig_id : YV_MISS_NODE;
  child[i] = hit ? (int32_t)(s.ch ^ s.flags) : YV_MISS_CHILD;
  t[i] = hit ? max3f(s.t1x, s.t1y, s.t1z) : 0.0f;
}

}  // namespace yv
