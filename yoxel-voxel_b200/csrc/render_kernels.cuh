// render_kernels.cuh — sm_100a kernels of the SVO ray caster.
//
// One fused kernel per frame: per-pixel ray generation (cell/ppu_renderer.cpp:49-57), slab set-up,
// octree descent over the packed 16-byte records (RecTrace, :18-41), leaf shading (:67) and the
// RGBA8 store (:54,:67) — the reference CUDA path's InitEyeRays / Trace / ShadeSimple sequence
// (demo/SVORenderer.cpp:95-149, trace_cuda.py:20-23) collapsed into one launch.
//
// Kernels in this file:
//   render_frame<SEC, COUNT, STACK, PERSISTENT, STAGED, LOD, RAW>   the renderer (default: all false / local stack)
//   render_queue, render_sec_queue                                   measured alternatives (see below), off by default
//   shade_pass                                                       ShadeSimple pass for Phong lights / show-normals
//   trace_rays_kernel<RAW>                                           DynamicSVO::TraceRay, batched
//
// render_frame is one warp-synchronous state machine:
//   * every lane owns one pixel at a time; a pixel's rays (primary, then shadow and AO samples when
//     SEC) are stages of the lane's state, so secondary rays re-enter the same traversal loop
//     instead of running as divergent tails, and ray set-up / shading run batched over many lanes;
//   * the traversal loop is flat: each trip every live lane performs one lean_step (trace_core.cuh) —
//     up to two sibling steps, then one (descend | pop); the warp votes every kStepsPerVote trips;
//   * PERSISTENT = false: one CTA per 16x8 pixel tile, one pixel per lane (8x4 pixels per warp);
//     PERSISTENT = true : resident CTAs; each warp pulls 8x8-pixel tiles from an atomic counter and
//     re-fills idle lanes (ballot + popc prefix) once <= refill_threshold lanes are still traversing;
//   * STAGED: the top `smem_nodes` records of the breadth-first pool are read from shared memory
//     (precedent: the SPU's software node cache, cell/spu/trace_spu.cpp:15-35). Measured on B200 the
//     126 MB L2 + L1 already serve these records (L1 hit rate 94 % without staging) and the extra
//     branch costs issue slots, so the default is smem_nodes = 0; the knob stays for ablation;
//   * STACK selects where the explicit traversal stack lives: local memory, or a 4-entry
//     shared-memory ring for the hot top of the stack that spills to local memory;
//   * LOD: SetDetailCoef cut-off; RAW: traverse the reference's 40-byte pool (scenes under edit).
// What the measurements say (profiles/README.md): the kernel is instruction-issue bound, and lanes of a warp
// that descend in lock-step share their node fetches. Every schedule that fills idle lanes by de-synchronising
// them (PERSISTENT refill, render_queue, render_sec_queue) executes fewer instructions and still loses.
#pragma once

#include <cuda.h>                      // CUtensorMap (types only: the encode function is fetched through the runtime)
#include <cuda_runtime.h>
#include <cooperative_groups.h>

#include "trace_core.cuh"

namespace yv {

struct RenderParams {
  const uint4 *recs;           // packed records, device form (svo_pack.h): { child_base, masks, leaf_base, orig_id }
  const uint2 *octs;           // { octants lo, octants hi } per record, read by the culling traversal only
  const uint32_t *leaves;      // inline VoxData words
  const uint32_t *node_data;   // VoxNode::data per record (LOD hits only; NULL when detail == 0)
  float detail;                // rp.detailCoef (demo/SVORenderer.cpp:104); 0 = no LOD cut-off
  uint32_t root_valid;
  uint32_t root_index;         // packed pool: 0; raw pool: the reference root id
  uint32_t smem_nodes;         // records staged in shared memory (<= record count)
  float pos[3];                // eye (m_pos) — also shader viewerPos (renderer_base.h:30-35)
  float dir0[3], du[3], dv[3]; // RayDirData (renderer_base.h:50-61), computed on the host
  float light[3];              // shader lightPos
  int width, height;           // m_viewSize
  int y0, y1;                  // row band rendered by this launch
  int band_rows8, band_stride, band_phase;   // interleaved partition: blocks of band_rows8*8 rows, this launch
                                             // renders blocks b with b % band_stride == band_phase (stride 1 = all)
  uint32_t *out_rgba;          // full-frame addressed: out_rgba[y*width + x]
  uint32_t *final_rgba;        // frames with a ShadeSimple pass: where that pass writes (NULL = in place). Lets the trace
                               // kernel draw in HBM while the pass that finishes the frame stores into a pinned host frame.
  uint32_t *hit_node;          // optional TraceResult planes (ppu_renderer.cpp:7-12); NULL = off
  int32_t *hit_child;
  float *hit_t;
  uint32_t *counters;          // COUNT: low 16 bits node visits, high 16 bits pop re-fetches
  unsigned int *tile_counter;  // persistent schedule: next tile to hand out
  int tiles_x, num_tiles;      // 8x8 tiles covering [0,width) x [y0,y1)
  int refill_threshold;        // persistent schedule: refill once <= this many lanes are still traversing
  int sec_threshold;           // secondary rays: hand waiting lanes their next ray once <= this many lanes traverse
  int shade_mode;              // 0 = head-light Lambert (SimpleShader), 1 = Phong point lights (ShadeSimple), 2 = normals
  yv_light lights[YV_MAX_LIGHTS];   // SetLigth(i, LightParams) (demo/SVORenderer.h:34)
  uint2 *shade_rec;            // shade_mode != 0: (VoxData, t) of every hit pixel for the ShadeSimple pass; else NULL
  int shadow, ao_samples;      // secondary rays
  uint32_t seed;
  float voxel_size, ao_max_t;
  // SSNA (SetSSNA, demo/SVORenderer.h:28): view-space z-buffer + the camera basis InitRayDir builds
  int ssna;
  float *zbuf;                 // ssna_z_pass: z0 out; shade_pass: the blurred z in; full-frame
  float fwd[3], right[3], down[3], d2;   // down = -(right x fwd); d2 = 2*da
  // displaced ray origins (reaction/report/main.tex:107-114): 0 = off
  float jitter_amp;
  uint32_t jitter_seed;
};

// BlurZ launch parameters (demo/SVORenderer.cpp:55-79,137): ping-pong buffers, zlimit, K*K Gaussian taps
struct BlurParams {
  const float *src;
  float *dst;
  int width, height;
  float zlimit;
  float wsum;                  // ((0 + w[0]) + w[1]) + ... over all K*K taps in row-major order: wacc of a pixel whose taps all pass
  float taps[YV_BLURZ_KERN * YV_BLURZ_KERN];
};

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kCtaThreads = 128;
#ifndef YV_FRAME_CTA
#define YV_FRAME_CTA 128               // threads per render_frame CTA (a 16x8-pixel tile is always four warps: smaller CTAs take
#endif                                 // a share of one tile's warps and free their SM slot without waiting for the others)
constexpr int kFrameCta = YV_FRAME_CTA;
#ifndef YV_MINBLOCKS
#define YV_MINBLOCKS 8                 // __launch_bounds__ residency target (register cap = 65536 / (128 * this))
#endif
constexpr int kFrameMinBlocks = YV_MINBLOCKS * kCtaThreads / kFrameCta;      // same register cap whatever the CTA size
#ifndef YV_WARP_W
#define YV_WARP_W 8                    // pixels per warp: 8x4 (4 = 4x8, 16 = 16x2); 8x4 measured best (profiles/README.md)
#endif
#ifndef YV_SEC_FAST_DESCENT
#define YV_SEC_FAST_DESCENT 1          // secondary rays: lean_descend_once in front of the general loop (0 = ablation)
#endif
#ifndef YV_PRIMARY_FAST_DESCENT
#define YV_PRIMARY_FAST_DESCENT 1      // primary rays with the eye inside the cube: the same, on levels without leaf children
                                       // (-1.4 % on config 2, -1.8 % at 8K, -0.5 % on the iso volume: profiles/README.md round 2)
#endif
#ifndef YV_STEPS_PER_VOTE
#define YV_STEPS_PER_VOTE 6            // lean_steps between two warp votes on "anyone still traversing?" (4 -> 6: -1.0..1.3 % frame
                                       // time on configs 2, 3, 4 and at 8K, 8 no better: profiles/README.md round 2)
#endif
constexpr int kStepsPerVote = YV_STEPS_PER_VOTE;

enum : int { kStackLocal = 0, kStackRing4 = 4 };

// ---- traversal stacks: two 16-byte words per entry ------------------------------------------------
__device__ __forceinline__ uint4 to_uint4(const U4 &v) { return make_uint4(v.x, v.y, v.z, v.w); }
__device__ __forceinline__ U4 to_u4(const uint4 &v) { U4 r = { v.x, v.y, v.z, v.w }; return r; }

// per-thread stack in local memory (LDL.128 / STL.128)
struct LocalStack {
  uint4 w[2 * kMaxStack];
  __device__ __forceinline__ LocalStack(uint4 *) {}
  __device__ __forceinline__ void reset() {}
  __device__ __forceinline__ void push(int sp, const U4 &a, const U4 &b) { w[2 * sp] = to_uint4(a); w[2 * sp + 1] = to_uint4(b); }
  __device__ __forceinline__ void pop(int sp, U4 &a, U4 &b) { a = to_u4(w[2 * sp]); b = to_u4(w[2 * sp + 1]); }
};

// YV_STACK_TOP (build-time variant, measured in profiles/README.md round 2): the newest entry stays in registers; it
// is written to local memory only when another push follows, so a push that is popped again before the next push
// (a child visit that finds nothing, 31 % of the visits on config 2) touches no memory at all.
#ifndef YV_STACK_TOP
#define YV_STACK_TOP 0
#endif
struct LocalStackTop {
  uint4 w[2 * kMaxStack];
  uint4 ta, tb;
  bool has;
  __device__ __forceinline__ LocalStackTop(uint4 *) : has(false) {}
  __device__ __forceinline__ void reset() { has = false; }
  __device__ __forceinline__ void push(int sp, const U4 &a, const U4 &b) {
    if (has) { w[2 * (sp - 1)] = ta; w[2 * (sp - 1) + 1] = tb; }
    ta = to_uint4(a); tb = to_uint4(b); has = true;
  }
  __device__ __forceinline__ void pop(int sp, U4 &a, U4 &b) {
    if (has) { a = to_u4(ta); b = to_u4(tb); has = false; }
    else { a = to_u4(w[2 * sp]); b = to_u4(w[2 * sp + 1]); }
  }
};

// the K most recent entries in shared memory ([slot][half][thread]: a warp's 128-bit accesses never
// bank-conflict), older ones spilled to local memory
template <int K>
struct RingStack {
  uint4 *base;
  uint4 spill[2 * kMaxStack];
  int lo;        // entries [lo, sp) live in the ring
  const int pitch;   // threads per CTA: the ring is laid out [slot][half][thread]
  __device__ __forceinline__ RingStack(uint4 *area) : base(area + threadIdx.x), lo(0), pitch((int)blockDim.x) {}
  __device__ __forceinline__ void reset() { lo = 0; }
  __device__ __forceinline__ void push(int sp, const U4 &a, const U4 &b) {
    if (sp - lo == K) {
      const int slot = lo & (K - 1);
      spill[2 * lo] = base[(2 * slot) * pitch];
      spill[2 * lo + 1] = base[(2 * slot + 1) * pitch];
      ++lo;
    }
    const int slot = sp & (K - 1);
    base[(2 * slot) * pitch] = to_uint4(a); base[(2 * slot + 1) * pitch] = to_uint4(b);
  }
  __device__ __forceinline__ void pop(int sp, U4 &a, U4 &b) {
    if (sp < lo) { lo = sp; a = to_u4(spill[2 * sp]); b = to_u4(spill[2 * sp + 1]); return; }
    const int slot = sp & (K - 1);
    a = to_u4(base[(2 * slot) * pitch]); b = to_u4(base[(2 * slot + 1) * pitch]);
  }
};

#if YV_STACK_TOP
template <int STACK> struct StackOf { using type = LocalStackTop; };
#else
template <int STACK> struct StackOf { using type = LocalStack; };
#endif
template <> struct StackOf<kStackRing4> { using type = RingStack<4>; };

// shared-memory bytes the stack variant needs per CTA of `threads` threads
__host__ __device__ inline size_t stack_smem_bytes(int stack, int threads = kCtaThreads) {
  return stack == kStackRing4 ? 4 * 2 * sizeof(uint4) * (size_t)threads : 0;
}

// node fetch policy (see trace_core.cuh). RAW = false: packed 16-byte records, optionally with the top of
// the tree staged in shared memory; RAW = true: the reference's 40-byte VoxNode pool as uploaded page by
// page for scenes under edit (CudaSVO::Update, demo/SVORenderer.cpp:33-53): `recs` then points at the
// pool viewed as uint32 words (flags, data, child[8]).
// the 512 box masks of trace_core.cuh's octant culling, index (dirFlags << 6) | (ch0 << 3) | chx, in stored-child space
struct BoxLut { uint8_t v[512]; };
constexpr BoxLut make_box_lut() {
  BoxLut t{};
  for (uint32_t f = 0; f < 8; ++f)
    for (uint32_t a = 0; a < 8; ++a)
      for (uint32_t b = 0; b < 8; ++b) {
        uint32_t box = 0;
        for (uint32_t o = 0; o < 8; ++o)
          if ((o & a) == a && (o | b) == b) box |= 1u << (o ^ f);
        t.v[(f << 6) | (a << 3) | b] = (uint8_t)box;
      }
  return t;
}
__device__ const BoxLut kBoxLut = make_box_lut();

template <bool COUNT, bool STAGED, bool RAW = false, bool CULL = false>
struct NodeFetch {
  const uint4 *recs;
  const uint4 *staged;
  uint32_t staged_n;
  uint32_t root;                 // index of the root node (0 for the packed pool)
  mutable uint32_t visits, revisits;
  const uint2 *octs;             // CULL: octant occupancy of every record's children (packed pool)
  const uint8_t *lut;            // CULL: the box masks, staged in shared memory
  __device__ __forceinline__ const uint32_t *pool() const { return reinterpret_cast<const uint32_t *>(recs); }
  __device__ __forceinline__ uint4 load(uint32_t idx) const {
    if (STAGED && idx < staged_n) return staged[idx];
    return __ldg(recs + idx);
  }
  // the two words a descent needs: the first half of the record
  __device__ __forceinline__ uint2 load2(uint32_t idx) const {
    if (STAGED && idx < staged_n) { const uint4 v = staged[idx]; return make_uint2(v.x, v.y); }
    return __ldg(reinterpret_cast<const uint2 *>(recs + idx));
  }
  __device__ __forceinline__ uint32_t root_index() const { return root; }
  // one node dereference: the two child masks (+ child base for the packed layout)
  __device__ __forceinline__ void node(uint32_t idx, bool visit, uint32_t &masks, uint32_t &child_base) const {
    if (COUNT) { if (visit) ++visits; else ++revisits; }
    if (RAW) {
      const uint32_t flags = __ldg(pool() + (size_t)idx * 10u);
      const uint32_t leaf = flags & 0xffu;
      masks = leaf | ((~(flags >> 8) & ~leaf & 0xffu) << 8);              // child = not leaf, not null
      child_base = 0u;
    } else {
      const uint2 r = load2(idx);
      child_base = r.x; masks = r.y;
    }
  }
  // ... and, for the culling traversal, the occupancy of the children's octants that rides in the same 16 bytes
  __device__ __forceinline__ void node(uint32_t idx, bool visit, uint32_t &masks, uint32_t &child_base,
                                       uint32_t &gm_lo, uint32_t &gm_hi) const {
    if (COUNT) { if (visit) ++visits; else ++revisits; }
    const uint2 r = load2(idx);
    const uint2 g = __ldg(octs + idx);
    child_base = r.x; masks = r.y; gm_lo = g.x; gm_hi = g.y;
  }
  __device__ __forceinline__ uint32_t box(uint32_t flags, uint32_t ch0, uint32_t chx) const {
    return lut[(flags << 6) | (ch0 << 3) | chx];
  }
  __device__ __forceinline__ uint32_t child_index(uint32_t idx, uint32_t child_base, uint32_t masks, uint32_t c) const {
    if (RAW) return __ldg(pool() + (size_t)idx * 10u + 2u + c);
    return child_base + (uint32_t)__popc((masks >> 8) & ((1u << c) - 1u));
  }
  // a finished ray: reference node id and the VoxData to shade (leaf slot c, or the node's own data for LOD).
  // `masks` = the masks of record idx (the traversal holds them when it reports a leaf hit; unused for a LOD hit)
  __device__ __forceinline__ void hit_info(const uint32_t *leaves, const uint32_t *node_data, uint32_t idx, uint32_t c,
                                           uint32_t masks, bool lod_hit, uint32_t &orig_id, uint32_t &data) const {
    if (RAW) {
      orig_id = idx;
      data = __ldg(pool() + (size_t)idx * 10u + (lod_hit ? 1u : 2u + c));
    } else {
      const uint2 in = __ldg(reinterpret_cast<const uint2 *>(recs + idx) + 1);     // { leaf_base, orig_id }: same sector as the descent's half
      orig_id = in.y;
      data = lod_hit ? __ldg(node_data + idx)
                     : __ldg(leaves + in.x + (uint32_t)__popc(masks & 0xffu & ((1u << c) - 1u)));
    }
  }
};

// the raw pool is traversed as it was handed in: bound the descent in the kernel (trace_core.cuh, FetchTraits)
template <bool COUNT, bool STAGED, bool RAW, bool CULL> struct FetchTraits<NodeFetch<COUNT, STAGED, RAW, CULL>> {
  static constexpr bool kGuardDepth = RAW;
  static constexpr bool kCull = CULL && !RAW;
};

// tile row (8 pixel rows) of this launch -> first pixel row, for contiguous and interleaved partitions
__device__ __forceinline__ int tile_row_y(const RenderParams &p, int ty) {
  const int blk = ty / p.band_rows8, within = ty - blk * p.band_rows8;
  return p.y0 + ((blk * p.band_stride + p.band_phase) * p.band_rows8 + within) * 8;
}

enum : int { kLaneIdle = 0, kLaneActive = 1, kLaneHit = 2, kLaneMiss = 3, kLaneNew = 4, kLaneLodHit = 5 };

// ZB (SSNA, the default schedule only): the hit epilogue also writes the view-space z-buffer BlurZ reads, so the frame
// needs no ssna_z_pass. A template parameter, not a run-time test of p.zbuf: the test alone cost the plain primary
// kernel 1.4 % (profiles/README.md, round 2).
template <bool SEC, bool COUNT, int STACK, bool PERSISTENT, bool STAGED, bool LOD, bool RAW, bool JIT = false, bool CULL = false, bool ZB = false>
__global__ void __launch_bounds__(kFrameCta, kFrameMinBlocks) render_frame(const __grid_constant__ RenderParams p) {
  extern __shared__ uint4 smem[];
  __shared__ uint32_t box_lut[CULL ? 128 : 1];
  uint4 *staged = smem;
  uint4 *stack_area = smem + (STAGED ? p.smem_nodes : 0u);
  if (CULL)
    for (uint32_t i = threadIdx.x; i < 128u; i += kFrameCta) box_lut[i] = reinterpret_cast<const uint32_t *>(kBoxLut.v)[i];
  if (STAGED)
    for (uint32_t i = threadIdx.x; i < p.smem_nodes; i += kFrameCta) staged[i] = __ldg(p.recs + i);
  if (STAGED || CULL) __syncthreads();

  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  NodeFetch<COUNT, STAGED, RAW, CULL> fetch = { p.recs, staged, p.smem_nodes, p.root_index, 0u, 0u, p.octs,
                                                reinterpret_cast<const uint8_t *>(box_lut) };
  typename StackOf<STACK>::type stk(stack_area);
  LeanState s;

  // warp-uniform tile pool (PERSISTENT)
  constexpr int kTilePix = 64;            // 8x8 pixels, handed out in Morton order
  int pool_next = kTilePix, tile_x0 = 0, tile_y0 = 0;
  bool pool_empty = !PERSISTENT;

  // per-lane pixel state
  int state = kLaneIdle;
  int x = 0, y = 0;
  // secondary-ray stage machine (SEC only): stage 0 = primary, 1 = shadow, 2.. = AO samples
  int stage = 0;
  uint32_t sdata = 0; float Ox = 0.f, Oy = 0.f, Oz = 0.f;
  float dl = 0.f, vis = 1.f, slen = 0.f; int occ = 0;
  bool fresh = false;                     // SEC: this lane's secondary ray has just been set up (lean_descend_once comes first)

  if (!PERSISTENT) {
    // one CTA per 16x8 tile: warp w covers an 8x4 block
    const int vwarp = (int)blockIdx.x * (kFrameCta / 32) + (threadIdx.x >> 5);     // four consecutive warps share a 16x8 tile
    const int warp = vwarp & 3, tile = vwarp >> 2;
    const int tiles_x16 = (p.width + 15) >> 4;
    const int tx = tile % tiles_x16, ty = tile / tiles_x16;
    // warp footprint YV_WARP_W x (32 / YV_WARP_W) pixels inside the CTA's 16x8 tile
    constexpr int kWW = YV_WARP_W, kWH = 32 / YV_WARP_W, kWarpsX = 16 / YV_WARP_W;
    x = tx * 16 + (warp % kWarpsX) * kWW + (lane % kWW);
    y = tile_row_y(p, ty) + (warp / kWarpsX) * kWH + (lane / kWW);
    if (x < p.width && y < p.y1) state = kLaneNew;
  }

  for (;;) {
    // ---- 1. hand new pixels to idle lanes (persistent schedule) -----------------------------
    if (PERSISTENT) {
      unsigned idle = __ballot_sync(kFullMask, state == kLaneIdle);
      while (idle != 0u && !pool_empty) {
        if (pool_next >= kTilePix) {
          unsigned t = 0;
          if (lane == 0) t = atomicAdd(p.tile_counter, 1u);
          t = __shfl_sync(kFullMask, t, 0);
          if (t >= (unsigned)p.num_tiles) { pool_empty = true; break; }
          tile_x0 = (int)(t % (unsigned)p.tiles_x) * 8;
          tile_y0 = tile_row_y(p, (int)(t / (unsigned)p.tiles_x));
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
    }

    // ---- 2. primary ray set-up for the new lanes ----------------------------------------------
    if (state == kLaneNew) {
      if (COUNT) { fetch.visits = 0; fetch.revisits = 0; }
      float dx, dy, dz;
      primary_dir(p.dir0, p.du, p.dv, x, y, dx, dy, dz);
      dx = adjust_dir1(dx); dy = adjust_dir1(dy); dz = adjust_dir1(dz);
      stage = 0;
      stk.reset();
      float ex = p.pos[0], ey = p.pos[1], ez = p.pos[2];
      if (JIT) jitter_origin(p.pos, p.jitter_amp, p.jitter_seed, (uint32_t)y * (uint32_t)p.width + (uint32_t)x, ex, ey, ez);
      state = lean_begin(s, fetch, p.root_valid != 0u, ex, ey, ez, dx, dy, dz) ? kLaneActive : kLaneMiss;
      if (YV_PRIMARY_FAST_DESCENT) fresh = true;
    }

    // ---- 3a. secondary rays: the descent to the origin's cell, level by level in closed form, the warp together ----
    // (trace_core.cuh, lean_descend_once: a third of a secondary ray's trips on config 4)
    if ((SEC && YV_SEC_FAST_DESCENT) || YV_PRIMARY_FAST_DESCENT) {
      bool fast = fresh && state == kLaneActive;
      fresh = false;
      while (__any_sync(kFullMask, fast)) {
        if (fast) fast = lean_descend_once<LOD>(s, fetch, stk, SEC && stage > 0);
      }
    }

    // ---- 3. traversal: one lean_step per live lane per iteration --------------------------------
    for (;;) {
      const unsigned am = __ballot_sync(kFullMask, state == kLaneActive);
      if (am == 0u) break;
      if (PERSISTENT && !pool_empty && __popc(am) <= p.refill_threshold) break;
      // secondary stages: lanes whose ray has ended wait for their next ray; serve them once few lanes
      // are still traversing instead of only when the whole warp has drained
      if (SEC && __popc(am) <= p.sec_threshold &&
          __ballot_sync(kFullMask, state == kLaneHit || state == kLaneMiss || state == kLaneLodHit) != 0u) break;
#pragma unroll
      for (int u = 0; u < kStepsPerVote; ++u) {
        if (state == kLaneActive) {
          const int r = lean_step<LOD>(s, fetch, stk, SEC && stage > 0, p.detail);
          if (r == kStepHit) state = kLaneHit;
          else if (r == kStepMiss) state = kLaneMiss;
          else if (LOD && r == kStepLodHit) state = kLaneLodHit;
        }
      }
    }

    // ---- 4. finished rays: shade / spawn the next secondary ray / store -------------------------
    if (state == kLaneHit || state == kLaneMiss || (LOD && state == kLaneLodHit)) {
      const bool lod_hit = LOD && state == kLaneLodHit;
      const bool hit = state == kLaneHit || lod_hit;
      const uint32_t pixel = (uint32_t)y * (uint32_t)p.width + (uint32_t)x;
      bool done = true;
      uint32_t rgba = 0u;
      float nx = 0.f, ny = 0.f, nz = 0.f;
      if (!SEC || stage == 0) {
        uint32_t hn = YV_MISS_NODE; int32_t hc = YV_MISS_CHILD; float ht = 0.0f, zview = 0.0f;
        if (hit) {
          const uint32_t c = lean_ch(s) ^ s.flags;
          fetch.hit_info(p.leaves, p.node_data, s.idx, c, s.masks, lod_hit, hn, sdata);
          hc = lod_hit ? -1 : (int32_t)c; ht = lean_hit_t(s);
          unpack_normal(sdata, nx, ny, nz);
          float dx, dy, dz;                                       // the primary direction, recomputed
          primary_dir(p.dir0, p.du, p.dv, x, y, dx, dy, dz);
          dx = adjust_dir1(dx); dy = adjust_dir1(dy); dz = adjust_dir1(dz);
          float ex = p.pos[0], ey = p.pos[1], ez = p.pos[2];
          if (JIT) jitter_origin(p.pos, p.jitter_amp, p.jitter_seed, pixel, ex, ey, ez);
          const float Px = YV_FADD(ex, YV_FMUL(dx, ht));
          const float Py = YV_FADD(ey, YV_FMUL(dy, ht));
          const float Pz = YV_FADD(ez, YV_FMUL(dz, ht));
          dl = lambert(nx, ny, nz, Px, Py, Pz, p.light[0], p.light[1], p.light[2]);
          // SSNA: the view-space depth of the hit, z = t * (d . forward), goes straight into the z-buffer BlurZ reads
          if (ZB) zview = YV_FMUL(ht, YV_FADD(YV_FADD(YV_FMUL(dx, p.fwd[0]), YV_FMUL(dy, p.fwd[1])), YV_FMUL(dz, p.fwd[2])));
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
        if (p.hit_node) { p.hit_node[pixel] = hn; p.hit_child[pixel] = hc; p.hit_t[pixel] = ht; }
        if (!SEC && p.shade_rec && hit) p.shade_rec[pixel] = make_uint2(sdata, __float_as_uint(ht));
        if (ZB) p.zbuf[pixel] = zview;                // 0 = no hit (what ssna_z_pass wrote as a pass of its own)
      } else {
        // a secondary ray came back
        const float ts = lean_hit_t(s);
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
            if (stage > 1 && nx == 0.f && ny == 0.f && nz == 0.f) unpack_normal(sdata, nx, ny, nz);
            ao_direction(nx, ny, nz, pixel, (uint32_t)(stage - 2), p.seed, rx, ry, rz);
          }
          rx = adjust_dir1(rx); ry = adjust_dir1(ry); rz = adjust_dir1(rz);
          stk.reset();
          if (lean_begin(s, fetch, p.root_valid != 0u, Ox, Oy, Oz, rx, ry, rz)) {
            s.tlimit = stage == 1 ? slen : p.ao_max_t;
            launched = true;          // otherwise this secondary ray misses outright: unoccluded
            fresh = true;
          }
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
        if (COUNT) p.counters[pixel] = (fetch.visits & 0xffffu) | (fetch.revisits << 16);
        state = kLaneIdle;
      }
    }

    // ---- 5. the warp retires when no pixel is pending and the queue is dry -----------------------
    if (pool_empty && __ballot_sync(kFullMask, state != kLaneIdle) == 0u) break;
  }
}

// ---------------------------------------------------------------------------------------------
// schedule 2 ("queue"): a warp owns a 16x8-pixel tile and schedules its 128 rays over its 32 lanes
// ---------------------------------------------------------------------------------------------
// Ray lengths inside one warp differ (mean/max ~ 0.77 on config 2), and an in-loop refill is only
// worth it if handing a lane its next ray is cheap. So the warp does the expensive, perfectly
// parallel parts up front and at the end, converged over all 32 lanes, and keeps only the descent in
// the divergent middle:
//   phase 0  every lane sets up 4 rays (ray generation, SetupTrace, root entry test, FindFirstChild in
//            the root) and parks the 7-word states in shared memory; rays that miss the cube outright
//            are resolved here; the others are appended to the warp's queue (ballot + popc prefix);
//   phase 1  lanes pull queue entries (7 LDS + 6 FADD), run lean_steps, and park the 3-word hit record
//            in the ray's slot when done; a warp vote every kStepsPerVote steps refills idle lanes;
//   phase 2  every lane shades and stores its 4 pixels.
// Slot order is the Morton order of the 16x8 tile, so the rays in flight stay spatially close.
constexpr int kQueueRays = 128;                      // rays per warp
constexpr int kQueueSlotWords = 8;                   // 7 state words + 1 (visit counters)
constexpr size_t kQueueSmemPerWarp = kQueueRays * kQueueSlotWords * 4 + kQueueRays;   // slots + queue bytes

__device__ __forceinline__ void queue_slot_xy(int slot, int &dx, int &dy) {
  // 7-bit Morton index -> (x in 0..15, y in 0..7): bits x0 y0 x1 y1 x2 y2 x3
  dx = (slot & 1) | ((slot >> 1) & 2) | ((slot >> 2) & 4) | ((slot >> 3) & 8);
  dy = ((slot >> 1) & 1) | ((slot >> 2) & 2) | ((slot >> 3) & 4);
}

template <bool COUNT, int STACK>
__global__ void __launch_bounds__(kCtaThreads, YV_MINBLOCKS) render_queue(const __grid_constant__ RenderParams p) {
  extern __shared__ uint4 smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint4 *stack_area = smem;                                // ring stack first (16-byte aligned), then the slots
  uint32_t *slots = reinterpret_cast<uint32_t *>(smem) + stack_smem_bytes(STACK) / 4 + warp * (kQueueSmemPerWarp / 4);   // [word][slot]
  uint8_t *queue = reinterpret_cast<uint8_t *>(slots + kQueueRays * kQueueSlotWords);

  // CTA tile: 32x16 pixels = 2x2 warp tiles of 16x8
  const int tiles_x32 = (p.width + 31) >> 5;
  const int tx = blockIdx.x % tiles_x32, ty = blockIdx.x / tiles_x32;
  const int wx0 = tx * 32 + (warp & 1) * 16, wy0 = tile_row_y(p, ty * 2 + (warp >> 1));

  NodeFetch<COUNT, false> fetch = { p.recs, nullptr, 0u, 0u, 0u, 0u, p.octs, nullptr };
  typename StackOf<STACK>::type stk(stack_area);
  LeanState s;
  const bool root_valid = p.root_valid != 0u;
  uint32_t root_masks = 0u, root_child_base = 0u;
  if (root_valid) { const uint4 r = fetch.load(0u); root_masks = r.y; root_child_base = r.x; }

  // ---- phase 0: set up 4 rays per lane, build the queue --------------------------------------------
  int qcount = 0;
#pragma unroll 1
  for (int k = 0; k < kQueueRays / 32; ++k) {
    const int slot = k * 32 + lane;
    int ox, oy; queue_slot_xy(slot, ox, oy);
    const int x = wx0 + ox, y = wy0 + oy;
    bool live = false;
    if (x < p.width && y < p.y1) {
      float dx, dy, dz;
      primary_dir(p.dir0, p.du, p.dv, x, y, dx, dy, dz);
      dx = adjust_dir1(dx); dy = adjust_dir1(dy); dz = adjust_dir1(dz);
      live = lean_setup_root(s, root_valid, p.pos[0], p.pos[1], p.pos[2], dx, dy, dz);
    }
    if (live) {
      slots[0 * kQueueRays + slot] = __float_as_uint(s.t1x); slots[1 * kQueueRays + slot] = __float_as_uint(s.t1y);
      slots[2 * kQueueRays + slot] = __float_as_uint(s.t1z); slots[3 * kQueueRays + slot] = __float_as_uint(s.Tx);
      slots[4 * kQueueRays + slot] = __float_as_uint(s.Ty);  slots[5 * kQueueRays + slot] = __float_as_uint(s.Tz);
      slots[6 * kQueueRays + slot] = s.ch | (s.flags << 3);
    } else {
      slots[0 * kQueueRays + slot] = 0xffffffffu;          // resolved: miss (or outside the frame)
    }
    if (COUNT) slots[7 * kQueueRays + slot] = live ? 1u : 0u;   // the root visit
    const unsigned lv = __ballot_sync(kFullMask, live);
    if (live) queue[qcount + __popc(lv & lt_mask)] = (uint8_t)slot;
    qcount += __popc(lv);
  }
  __syncwarp();

  // ---- phase 1: lanes pull rays from the queue -------------------------------------------------------
  int qnext = 0;
  int cur = -1;                                            // slot of the ray this lane is tracing
  for (;;) {
    const unsigned idle = __ballot_sync(kFullMask, cur < 0);
    if (idle != 0u && qnext < qcount) {
      const int take = min(__popc(idle), qcount - qnext);
      const int rank = __popc(idle & lt_mask);
      if (cur < 0 && rank < take) {
        cur = queue[qnext + rank];
        s.t1x = __uint_as_float(slots[0 * kQueueRays + cur]); s.t1y = __uint_as_float(slots[1 * kQueueRays + cur]);
        s.t1z = __uint_as_float(slots[2 * kQueueRays + cur]); s.Tx = __uint_as_float(slots[3 * kQueueRays + cur]);
        s.Ty = __uint_as_float(slots[4 * kQueueRays + cur]);  s.Tz = __uint_as_float(slots[5 * kQueueRays + cur]);
        const uint32_t w = slots[6 * kQueueRays + cur];
        s.ch = w & 7u; s.flags = w >> 3; s.idx = 0u; s.sp = 0; s.pend = 0u; s.st = 0u; s.tlimit = __builtin_huge_valf();
        s.masks = root_masks; s.child_base = root_child_base;
        stk.reset();
        lean_eval_next(s);
        if (COUNT) { fetch.visits = 1u; fetch.revisits = 0u; }
      }
      qnext += take;
    }
    if (__ballot_sync(kFullMask, cur >= 0) == 0u) break;
#pragma unroll
    for (int u = 0; u < kStepsPerVote; ++u) {
      if (cur >= 0) {
        const int r = lean_step<false>(s, fetch, stk, false);
        if (r != kStepContinue) {
          const bool hit = r == kStepHit;
          slots[0 * kQueueRays + cur] = hit ? s.idx : 0xffffffffu;
          slots[1 * kQueueRays + cur] = lean_ch(s) ^ s.flags;
          slots[2 * kQueueRays + cur] = __float_as_uint(lean_hit_t(s));
          if (COUNT) slots[7 * kQueueRays + cur] = (fetch.visits & 0xffffu) | (fetch.revisits << 16);
          cur = -1;
        }
      }
    }
  }
  __syncwarp();

  // ---- phase 2: shade and store 4 pixels per lane ------------------------------------------------------
#pragma unroll 1
  for (int k = 0; k < kQueueRays / 32; ++k) {
    const int slot = k * 32 + lane;
    int ox, oy; queue_slot_xy(slot, ox, oy);
    const int x = wx0 + ox, y = wy0 + oy;
    if (x >= p.width || y >= p.y1) continue;
    const uint32_t pixel = (uint32_t)y * (uint32_t)p.width + (uint32_t)x;
    const uint32_t idx = slots[0 * kQueueRays + slot];
    uint32_t rgba = 0u, hn = YV_MISS_NODE; int32_t hc = YV_MISS_CHILD; float ht = 0.0f;
    if (idx != 0xffffffffu) {
      const uint32_t c = slots[1 * kQueueRays + slot];
      ht = __uint_as_float(slots[2 * kQueueRays + slot]);
      hc = (int32_t)c;
      uint32_t data;
      fetch.hit_info(p.leaves, p.node_data, idx, c, fetch.load(idx).y, false, hn, data);
      float nx, ny, nz, dx, dy, dz;
      unpack_normal(data, nx, ny, nz);
      primary_dir(p.dir0, p.du, p.dv, x, y, dx, dy, dz);
      dx = adjust_dir1(dx); dy = adjust_dir1(dy); dz = adjust_dir1(dz);
      const float Px = YV_FADD(p.pos[0], YV_FMUL(dx, ht));
      const float Py = YV_FADD(p.pos[1], YV_FMUL(dy, ht));
      const float Pz = YV_FADD(p.pos[2], YV_FMUL(dz, ht));
      const float dl = lambert(nx, ny, nz, Px, Py, Pz, p.light[0], p.light[1], p.light[2]);
      rgba = shade_rgba(data, YV_FADD(YV_SHADE_AMBIENT, YV_FMUL(YV_SHADE_DIFFUSE, YV_FMUL(dl, 1.0f))));
      if (p.shade_rec) p.shade_rec[pixel] = make_uint2(data, __float_as_uint(ht));
    }
    p.out_rgba[pixel] = rgba;
    if (p.hit_node) { p.hit_node[pixel] = hn; p.hit_child[pixel] = hc; p.hit_t[pixel] = ht; }
    if (COUNT) p.counters[pixel] = slots[7 * kQueueRays + slot];
  }
}

// ---------------------------------------------------------------------------------------------
// secondary rays with a per-warp AO-ray queue (BASELINE config 4)
// ---------------------------------------------------------------------------------------------
// In the stage machine of render_frame a lane traces its pixel's shadow and AO rays one after the other,
// so a warp runs as long as its slowest pixel: ncu shows 11 of 32 lanes busy in the child test on config 4
// (42 % of the pixels miss and have no secondary rays; AO rays differ wildly in length). AO rays are
// incoherent by construction, so unlike primary rays they lose nothing by being handed to whichever lane
// is free. This kernel therefore keeps the coherent rays in lock-step and pools the incoherent ones:
//   A  primary rays, one per lane, lock-step (as render_frame's tile schedule);
//   B  shadow rays of the hit pixels, one per lane, lock-step (they all aim at the light);
//   C  AO rays: every hit lane sets up its pixel's next (up to 4) AO rays and parks them in shared memory;
//      the warp's lanes then pull rays from that queue until it is dry, adding an occlusion to the pixel's
//      shared-memory counter when a ray ends occluded; repeated while samples remain;
//   D  every hit lane shades and stores its pixel.
// Results are order-independent sums, so the frame is bit-identical to the stage machine's and the oracle's.
constexpr int kAoBatch = 4;                               // AO rays queued per pixel and pass
constexpr int kAoQueueRays = 32 * kAoBatch;               // per warp
constexpr size_t kAoSmemPerWarp = kAoQueueRays * 8 * 4 + 32 * 4 + 32 * 4;

template <bool COUNT, bool LOD>
__global__ void __launch_bounds__(kCtaThreads, YV_MINBLOCKS) render_sec_queue(const __grid_constant__ RenderParams p) {
  extern __shared__ uint4 smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint32_t *slots = reinterpret_cast<uint32_t *>(smem) + warp * (kAoSmemPerWarp / 4);     // [word][slot], 8 words
  uint32_t *occ_cnt = slots + kAoQueueRays * 8;                                             // per pixel (lane)
  uint32_t *vis_cnt = occ_cnt + 32;                                                          // per pixel: extra node visits (COUNT)

  const int tiles_x16 = (p.width + 15) >> 4;
  const int tx = blockIdx.x % tiles_x16, ty = blockIdx.x / tiles_x16;
  const int x = tx * 16 + (warp & 1) * 8 + (lane & 7);
  const int y = tile_row_y(p, ty) + (warp >> 1) * 4 + (lane >> 3);
  const bool in_frame = x < p.width && y < p.y1;
  const uint32_t pixel = (uint32_t)y * (uint32_t)p.width + (uint32_t)x;

  NodeFetch<COUNT, false> fetch = { p.recs, nullptr, 0u, 0u, 0u, 0u, p.octs, nullptr };
  LocalStack stk(nullptr);
  LeanState s;
  const bool root_valid = p.root_valid != 0u;
  uint32_t root_masks = 0u, root_child_base = 0u;
  if (root_valid) { const uint4 r = fetch.load(0u); root_masks = r.y; root_child_base = r.x; }

  // ---- A: primary ray, lock-step ------------------------------------------------------------------
  int state = kLaneIdle;
  if (in_frame) {
    float dx, dy, dz;
    primary_dir(p.dir0, p.du, p.dv, x, y, dx, dy, dz);
    dx = adjust_dir1(dx); dy = adjust_dir1(dy); dz = adjust_dir1(dz);
    state = lean_begin(s, fetch, root_valid, p.pos[0], p.pos[1], p.pos[2], dx, dy, dz) ? kLaneActive : kLaneMiss;
  }
  auto run_lockstep = [&](bool front_only) {
    for (;;) {
      if (__ballot_sync(kFullMask, state == kLaneActive) == 0u) break;
#pragma unroll
      for (int u = 0; u < kStepsPerVote; ++u) {
        if (state == kLaneActive) {
          const int r = lean_step<LOD>(s, fetch, stk, front_only, p.detail);
          if (r == kStepHit) state = kLaneHit;
          else if (r == kStepMiss) state = kLaneMiss;
          else if (LOD && r == kStepLodHit) state = kLaneLodHit;
        }
      }
    }
  };
  run_lockstep(false);

  // ---- primary result; hit point, normal, secondary origin ---------------------------------------------
  const bool lod_hit = LOD && state == kLaneLodHit;
  const bool hit = state == kLaneHit || lod_hit;
  uint32_t sdata = 0u, hn = YV_MISS_NODE; int32_t hc = YV_MISS_CHILD; float ht = 0.0f;
  float nx = 0.f, ny = 0.f, nz = 0.f, Ox = 0.f, Oy = 0.f, Oz = 0.f, dl = 0.f, vis = 1.0f;
  if (hit) {
    const uint32_t c = lean_ch(s) ^ s.flags;
    fetch.hit_info(p.leaves, p.node_data, s.idx, c, s.masks, lod_hit, hn, sdata);
    hc = lod_hit ? -1 : (int32_t)c; ht = lean_hit_t(s);
    unpack_normal(sdata, nx, ny, nz);
    float dx, dy, dz;
    primary_dir(p.dir0, p.du, p.dv, x, y, dx, dy, dz);
    dx = adjust_dir1(dx); dy = adjust_dir1(dy); dz = adjust_dir1(dz);
    const float Px = YV_FADD(p.pos[0], YV_FMUL(dx, ht));
    const float Py = YV_FADD(p.pos[1], YV_FMUL(dy, ht));
    const float Pz = YV_FADD(p.pos[2], YV_FMUL(dz, ht));
    dl = lambert(nx, ny, nz, Px, Py, Pz, p.light[0], p.light[1], p.light[2]);
    Ox = YV_FADD(Px, YV_FMUL(nx, p.voxel_size));
    Oy = YV_FADD(Py, YV_FMUL(ny, p.voxel_size));
    Oz = YV_FADD(Pz, YV_FMUL(nz, p.voxel_size));
  }
  if (in_frame && p.hit_node) { p.hit_node[pixel] = hn; p.hit_child[pixel] = hc; p.hit_t[pixel] = ht; }

  // ---- B: shadow rays, lock-step -------------------------------------------------------------------------
  state = kLaneIdle;
  float slen = 0.0f;
  if (hit && p.shadow) {
    const float vx = YV_FSUB(p.light[0], Ox), vy = YV_FSUB(p.light[1], Oy), vz = YV_FSUB(p.light[2], Oz);
    slen = YV_FSQRT(YV_FADD(YV_FADD(YV_FMUL(vx, vx), YV_FMUL(vy, vy)), YV_FMUL(vz, vz)));
    if (slen > 0) {
      const float rx = adjust_dir1(YV_FDIV(vx, slen)), ry = adjust_dir1(YV_FDIV(vy, slen)), rz = adjust_dir1(YV_FDIV(vz, slen));
      if (lean_begin(s, fetch, root_valid, Ox, Oy, Oz, rx, ry, rz)) { s.tlimit = slen; state = kLaneActive; }
    }
  }
  run_lockstep(true);
  if (state == kLaneHit || (LOD && state == kLaneLodHit)) {
    const float ts = lean_hit_t(s);
    if (ts > 0 && ts < slen) vis = 0.0f;
  }

  // ---- C: AO rays through the warp's queue ------------------------------------------------------------------
  const uint32_t own_visits = fetch.visits, own_revisits = fetch.revisits;   // this pixel's primary + shadow rays
  occ_cnt[lane] = 0u;
  if (COUNT) vis_cnt[lane] = 0u;
  __syncwarp();
  for (int k0 = 0; k0 < p.ao_samples; k0 += kAoBatch) {
    const int nb = min(kAoBatch, p.ao_samples - k0);
    // set-up: hit lanes park their pixel's next nb rays
    const unsigned hm = __ballot_sync(kFullMask, hit);
    const int my_rank = __popc(hm & lt_mask);
    int qcount = __popc(hm) * nb;
    if (hit) {
      for (int k = 0; k < nb; ++k) {
        float rx, ry, rz;
        ao_direction(nx, ny, nz, pixel, (uint32_t)(k0 + k), p.seed, rx, ry, rz);
        rx = adjust_dir1(rx); ry = adjust_dir1(ry); rz = adjust_dir1(rz);
        const int slot = my_rank * nb + k;
        LeanState q;
        if (lean_setup_root(q, root_valid, Ox, Oy, Oz, rx, ry, rz)) {
          slots[0 * kAoQueueRays + slot] = __float_as_uint(q.t1x); slots[1 * kAoQueueRays + slot] = __float_as_uint(q.t1y);
          slots[2 * kAoQueueRays + slot] = __float_as_uint(q.t1z); slots[3 * kAoQueueRays + slot] = __float_as_uint(q.Tx);
          slots[4 * kAoQueueRays + slot] = __float_as_uint(q.Ty);  slots[5 * kAoQueueRays + slot] = __float_as_uint(q.Tz);
          slots[6 * kAoQueueRays + slot] = q.ch | (q.flags << 3) | ((uint32_t)lane << 6);
        } else {
          slots[6 * kAoQueueRays + slot] = 0xffffffffu;        // misses the cube outright: unoccluded, nothing to trace
        }
      }
    }
    __syncwarp();
    // lanes pull rays
    int qnext = 0, cur = -1, owner = 0;
    for (;;) {
      const unsigned idle = __ballot_sync(kFullMask, cur < 0);
      if (idle != 0u && qnext < qcount) {
        const int take = min(__popc(idle), qcount - qnext);
        const int rank = __popc(idle & lt_mask);
        if (cur < 0 && rank < take) {
          const int slot = qnext + rank;
          const uint32_t w = slots[6 * kAoQueueRays + slot];
          if (w != 0xffffffffu) {
            cur = slot; owner = (int)((w >> 6) & 31u);
            s.t1x = __uint_as_float(slots[0 * kAoQueueRays + slot]); s.t1y = __uint_as_float(slots[1 * kAoQueueRays + slot]);
            s.t1z = __uint_as_float(slots[2 * kAoQueueRays + slot]); s.Tx = __uint_as_float(slots[3 * kAoQueueRays + slot]);
            s.Ty = __uint_as_float(slots[4 * kAoQueueRays + slot]);  s.Tz = __uint_as_float(slots[5 * kAoQueueRays + slot]);
            s.ch = w & 7u; s.flags = (w >> 3) & 7u; s.idx = 0u; s.sp = 0; s.pend = 0u; s.st = 0u; s.level = 0u; s.tlimit = p.ao_max_t;
            s.masks = root_masks; s.child_base = root_child_base;
            lean_eval_next(s);
            if (COUNT) { atomicAdd(&vis_cnt[owner], 1u); }
          }
        }
        qnext += take;
      }
      if (__ballot_sync(kFullMask, cur >= 0) == 0u) { if (qnext >= qcount) break; else continue; }
#pragma unroll
      for (int u = 0; u < kStepsPerVote; ++u) {
        if (cur >= 0) {
          const uint32_t v0 = COUNT ? fetch.visits : 0u, r0 = COUNT ? fetch.revisits : 0u;
          const int r = lean_step<LOD>(s, fetch, stk, true, p.detail);
          if (COUNT) { const uint32_t dv = fetch.visits - v0, dr = fetch.revisits - r0; if (dv | dr) atomicAdd(&vis_cnt[owner], dv | (dr << 16)); }
          if (r != kStepContinue) {
            if (r != kStepMiss) {
              const float ts = lean_hit_t(s);
              if (ts > 0 && ts < p.ao_max_t) atomicAdd(&occ_cnt[owner], 1u);
            }
            cur = -1;
          }
        }
      }
    }
    __syncwarp();
  }

  // ---- D: shade + store ------------------------------------------------------------------------------------
  if (!in_frame) return;
  uint32_t rgba = 0u;
  if (hit) {
    float ao = 1.0f;
    if (p.ao_samples > 0) ao = YV_FSUB(1.0f, YV_FDIV((float)occ_cnt[lane], (float)p.ao_samples));
    const float k = YV_FMUL(YV_FADD(YV_SHADE_AMBIENT, YV_FMUL(YV_SHADE_DIFFUSE, YV_FMUL(dl, vis))), ao);
    rgba = shade_rgba(sdata, k);
  }
  p.out_rgba[pixel] = rgba;
  if (COUNT) {
    const uint32_t extra = vis_cnt[lane];
    p.counters[pixel] = ((own_visits + (extra & 0xffffu)) & 0xffffu) | ((own_revisits + (extra >> 16)) << 16);
  }
}

// ---------------------------------------------------------------------------------------------
// ShadeSimple pass (demo/SVORenderer.cpp:147): the CUDA renderer's second kernel. Used only for the shading
// models the CPU tracer does not have (point-light Phong, show-normals): the trace kernel leaves (VoxData, t)
// per hit pixel, this pass re-derives the ray and overwrites the pixel. Keeping it out of the trace kernel
// costs that kernel no registers.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void shade_pixel(const RenderParams &p, const float *__restrict__ zbuf, const int x, const int y) {
  if (x >= p.width || y >= p.y1) return;
  const uint32_t pixel = (uint32_t)y * (uint32_t)p.width + (uint32_t)x;
  uint32_t *dst = p.final_rgba ? p.final_rgba : p.out_rgba;
  if (p.out_rgba[pixel] == 0u) {                          // miss (hit pixels carry alpha 255)
    if (p.final_rgba) dst[pixel] = 0u;
    return;
  }
  const uint2 rec = p.shade_rec[pixel];
  const float t = __uint_as_float(rec.y);
  float nx = 0.f, ny = 0.f, nz = 0.f, dx, dy, dz;
  bool have_n = false;                                    // the stored normal is unpacked only where SSNA has none to offer
  if (p.ssna) {                                           // normal from the blurred z-buffer (yv_format.h "SSNA")
    const float z = zbuf[pixel];
    if (z != 0.0f) {
      const float zr = x + 1 < p.width ? zbuf[pixel + 1] : 0.0f, zl = x > 0 ? zbuf[pixel - 1] : 0.0f;
      const float zd = y + 1 < p.height ? zbuf[pixel + p.width] : 0.0f, zu = y > 0 ? zbuf[pixel - p.width] : 0.0f;
      const float ddx = abs_min_diff(zr != 0.0f, YV_FSUB(zr, z), zl != 0.0f, YV_FSUB(z, zl));
      const float ddy = abs_min_diff(zd != 0.0f, YV_FSUB(zd, z), zu != 0.0f, YV_FSUB(z, zu));
      const float nvx = YV_FMUL(YV_FMUL(p.d2, ddx), z);
      const float nvy = YV_FMUL(YV_FMUL(p.d2, ddy), z);
      const float nvz = -YV_FMUL(YV_FMUL(p.d2, p.d2), YV_FMUL(z, z));
      const float len = YV_FSQRT(YV_FADD(YV_FADD(YV_FMUL(nvx, nvx), YV_FMUL(nvy, nvy)), YV_FMUL(nvz, nvz)));
      if (len > 0.0f) {
        nx = YV_FDIV(YV_FADD(YV_FADD(YV_FMUL(p.right[0], nvx), YV_FMUL(p.down[0], nvy)), YV_FMUL(p.fwd[0], nvz)), len);
        ny = YV_FDIV(YV_FADD(YV_FADD(YV_FMUL(p.right[1], nvx), YV_FMUL(p.down[1], nvy)), YV_FMUL(p.fwd[1], nvz)), len);
        nz = YV_FDIV(YV_FADD(YV_FADD(YV_FMUL(p.right[2], nvx), YV_FMUL(p.down[2], nvy)), YV_FMUL(p.fwd[2], nvz)), len);
        have_n = true;
      }
    }
  }
  if (!have_n) unpack_normal(rec.x, nx, ny, nz);
  uint32_t rgba;
  if (p.shade_mode == 2) rgba = shade_normal(nx, ny, nz);
  else {
    primary_dir(p.dir0, p.du, p.dv, x, y, dx, dy, dz);
    dx = adjust_dir1(dx); dy = adjust_dir1(dy); dz = adjust_dir1(dz);
    float ex = p.pos[0], ey = p.pos[1], ez = p.pos[2];
    if (p.jitter_amp > 0.0f) jitter_origin(p.pos, p.jitter_amp, p.jitter_seed, pixel, ex, ey, ez);
    const float Px = YV_FADD(ex, YV_FMUL(dx, t));
    const float Py = YV_FADD(ey, YV_FMUL(dy, t));
    const float Pz = YV_FADD(ez, YV_FMUL(dz, t));
    if (p.shade_mode == 1) rgba = shade_phong(rec.x, nx, ny, nz, Px, Py, Pz, p.pos, p.lights);
    else {                                                // head-light Lambert with the SSNA normal
      const float dl = lambert(nx, ny, nz, Px, Py, Pz, p.pos[0], p.pos[1], p.pos[2]);
      rgba = shade_rgba(rec.x, YV_FADD(YV_SHADE_AMBIENT, YV_FMUL(YV_SHADE_DIFFUSE, YV_FMUL(dl, 1.0f))));
    }
  }
  dst[pixel] = rgba;
}

__global__ void __launch_bounds__(256) shade_pass(const __grid_constant__ RenderParams p) {
  // a CTA = 32x8 pixels of the rows of this launch's band / blocks
  shade_pixel(p, p.zbuf, blockIdx.x * 32 + (threadIdx.x & 31), tile_row_y(p, blockIdx.y) + (threadIdx.x >> 5));
}

// ---------------------------------------------------------------------------------------------
// SSNA passes (SVORenderer::Render, demo/SVORenderer.cpp:126-147): the trace kernel leaves (VoxData, t) per hit
// pixel; ssna_z_pass turns t into the view-space z-buffer, blur_z_pass runs five times on ping-pong buffers,
// shade_pass (above) rebuilds the normal from the result. All three are HBM/L2 streaming passes over 4 B/pixel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ssna_z_pass(const __grid_constant__ RenderParams p) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= p.width || y >= p.height) return;
  const uint32_t pixel = (uint32_t)y * (uint32_t)p.width + (uint32_t)x;
  float z = 0.0f;
  if (p.out_rgba[pixel] != 0u) {
    const float t = __uint_as_float(p.shade_rec[pixel].y);
    float dx, dy, dz;
    primary_dir(p.dir0, p.du, p.dv, x, y, dx, dy, dz);
    dx = adjust_dir1(dx); dy = adjust_dir1(dy); dz = adjust_dir1(dz);
    z = YV_FMUL(t, YV_FADD(YV_FADD(YV_FMUL(dx, p.fwd[0]), YV_FMUL(dy, p.fwd[1])), YV_FMUL(dz, p.fwd[2])));
  }
  p.zbuf[pixel] = z;
}

// One BlurZ pass: a 32x32 output tile per CTA with its 3-pixel apron staged in shared memory (row pitch 39 words: a
// warp's 32 consecutive columns never share a bank). A thread owns four vertically adjacent outputs and walks the ten
// tile rows they touch once, so each staged value is read once per column offset instead of once per tap (70 LDS
// for 196 taps). Invalid pixels are staged as +inf: |inf - zc| < zlimit is false, which folds the validity test into
// the depth test. Per output the taps still arrive in row-major (ky, kx) order with separate multiply and add, so the
// result matches the CPU statement bit for bit. 5 ALU instructions per tap: the pass is issue-bound, not HBM-bound
// (245 instructions against 8 bytes per pixel) — so a warp first finds out how much of that it needs (round 2):
//   * background: none of the warp's 32x4 outputs is valid -> zeros, no taps (27 % of the warps on config 2);
//   * smooth:     the 38x10 staged values the warp's taps can reach are all valid and max - min < zlimit. Then every
//                 tap of every output passes its test — rounding is monotone, so fl(|zq - zc|) <= fl(max - min) — and
//                 the pass is 2 instructions per tap; wacc is the constant the full row-major sum of the taps gives
//                 (BlurParams::wsum, summed on the host in that order);
//   * otherwise the tested form.
constexpr int kBlurTile = 32, kBlurApron = YV_BLURZ_KERN / 2, kBlurSpan = kBlurTile + 2 * kBlurApron, kBlurRows = 4;
#ifndef YV_POST_MINB
#define YV_POST_MINB 6
#endif
#ifndef YV_BLUR_INLINE
#define YV_BLUR_INLINE __forceinline__
#endif
typedef float BlurTile[kBlurSpan][kBlurSpan + 1];
template <int PITCH, int XOFF>
__device__ YV_BLUR_INLINE void blur_compute(const BlurParams &b, float *__restrict__ dst, const float zlimit,
                                            const float (&tile)[kBlurSpan][PITCH], const int bx, const int by);

// one 32x32 output tile at (bx, by) by one 256-thread CTA; contains one __syncthreads, threads return at different points after it
__device__ YV_BLUR_INLINE void blur_tile(const BlurParams &b, const float *__restrict__ src, float *__restrict__ dst, const float zlimit,
                                          BlurTile &tile, const int bx, const int by) {
  const float kInvalid = __int_as_float(0x7f800000);
  // every thread runs the same number of trips (the last one predicated): with `i < span*span` as the loop condition the
  // lanes of one warp leave the loop at different trips, and inside ssna_post's tile loop ptxas then emitted the barrier
  // below without re-converging them first — the early lanes ran on into the taps while lanes 0-3 of warps 0-5 were
  // still staging (compute-sanitizer racecheck; an out-of-range address through a clobbered uniform register)
#pragma unroll
  for (int k = 0; k < (kBlurSpan * kBlurSpan + 255) / 256; ++k) {
    const int i = (int)threadIdx.x + 256 * k;
    if (i < kBlurSpan * kBlurSpan) {
      const int ty = i / kBlurSpan, tx = i - ty * kBlurSpan;
      const int gx = bx + tx - kBlurApron, gy = by + ty - kBlurApron;
      float v = 0.0f;                                     // outside the frame = invalid
      if (gx >= 0 && gx < b.width && gy >= 0 && gy < b.height) v = src[(size_t)gy * b.width + gx];
      tile[ty][tx] = v != 0.0f ? v : kInvalid;
    }
  }
  __syncwarp();
  __syncthreads();
  blur_compute<kBlurSpan + 1, 0>(b, dst, zlimit, tile, bx, by);
}

// the taps of one staged tile (invalid pixels hold +inf; tile column XOFF is frame column bx - 3); no barrier inside,
// threads return at different points
template <int PITCH, int XOFF>
__device__ YV_BLUR_INLINE void blur_compute(const BlurParams &b, float *__restrict__ dst, const float zlimit,
                                            const float (&tile)[kBlurSpan][PITCH], const int bx, const int by) {
  const float kInvalid = __int_as_float(0x7f800000);
  const int lx = threadIdx.x & 31, gx = bx + lx;
  const int ly = (threadIdx.x >> 5) * kBlurRows;           // first of this thread's four output rows
  float zc[kBlurRows], acc[kBlurRows], wacc[kBlurRows];
  bool any_valid = false;
#pragma unroll
  for (int j = 0; j < kBlurRows; ++j) {
    zc[j] = tile[ly + j + kBlurApron][lx + kBlurApron + XOFF]; acc[j] = 0.0f; wacc[j] = 0.0f;
    any_valid = any_valid || zc[j] != kInvalid;            // outputs outside the frame were staged invalid
  }
  const bool in_x = gx < b.width;
  if (__ballot_sync(kFullMask, any_valid) == 0u) {         // background warp
#pragma unroll
    for (int j = 0; j < kBlurRows; ++j)
      if (in_x && by + ly + j < b.height) dst[(size_t)(by + ly + j) * b.width + gx] = 0.0f;
    return;
  }
  // range of everything this warp's taps can reach: tile rows ly .. ly+9, columns 0 .. 37
  float mn = kInvalid, mx = -kInvalid;
#pragma unroll
  for (int r = 0; r < kBlurRows + YV_BLURZ_KERN - 1; ++r) {
    const float v = tile[ly + r][lx + XOFF];
    mn = fminf(mn, v); mx = fmaxf(mx, v);
    if (lx < 2 * kBlurApron) { const float u = tile[ly + r][kBlurTile + lx + XOFF]; mn = fminf(mn, u); mx = fmaxf(mx, u); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(kFullMask, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, o));
  }
  const bool smooth = YV_FSUB(mx, mn) < zlimit;           // false when anything is invalid (inf - x, inf - inf)
  if (smooth) {
#pragma unroll
    for (int r = 0; r < kBlurRows + YV_BLURZ_KERN - 1; ++r) {
      float zq[YV_BLURZ_KERN];
#pragma unroll
      for (int kx = 0; kx < YV_BLURZ_KERN; ++kx) zq[kx] = tile[ly + r][lx + kx + XOFF];
#pragma unroll
      for (int j = 0; j < kBlurRows; ++j) {
        const int ky = r - j;
        if (ky < 0 || ky >= YV_BLURZ_KERN) continue;
#pragma unroll
        for (int kx = 0; kx < YV_BLURZ_KERN; ++kx) acc[j] = YV_FADD(acc[j], YV_FMUL(b.taps[ky * YV_BLURZ_KERN + kx], zq[kx]));
      }
    }
#pragma unroll
    for (int j = 0; j < kBlurRows; ++j)
      if (in_x && by + ly + j < b.height) dst[(size_t)(by + ly + j) * b.width + gx] = YV_FDIV(acc[j], b.wsum);
    return;
  }
#pragma unroll
  for (int r = 0; r < kBlurRows + YV_BLURZ_KERN - 1; ++r) {           // tile row ly + r feeds output j as tap row ky = r - j
    float zq[YV_BLURZ_KERN];
#pragma unroll
    for (int kx = 0; kx < YV_BLURZ_KERN; ++kx) zq[kx] = tile[ly + r][lx + kx + XOFF];
#pragma unroll
    for (int j = 0; j < kBlurRows; ++j) {
      const int ky = r - j;
      if (ky < 0 || ky >= YV_BLURZ_KERN) continue;
#pragma unroll
      for (int kx = 0; kx < YV_BLURZ_KERN; ++kx) {
        const float w = b.taps[ky * YV_BLURZ_KERN + kx];
        const bool ok = fabsf(YV_FSUB(zq[kx], zc[j])) < zlimit;
        acc[j] = ok ? YV_FADD(acc[j], YV_FMUL(w, zq[kx])) : acc[j];
        wacc[j] = ok ? YV_FADD(wacc[j], w) : wacc[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kBlurRows; ++j) {
    const int gy = by + ly + j;
    if (!in_x || gy >= b.height) continue;
    float out = 0.0f;
    if (zc[j] != kInvalid) out = wacc[j] > 0.0f ? YV_FDIV(acc[j], wacc[j]) : zc[j];
    dst[(size_t)gy * b.width + gx] = out;
  }
}

__global__ void __launch_bounds__(256) blur_z_pass(const __grid_constant__ BlurParams b) {
  __shared__ BlurTile tile;
  blur_tile(b, b.src, b.dst, b.zlimit, tile, blockIdx.x * kBlurTile, blockIdx.y * kBlurTile);
}

constexpr int kTmaPitch = kBlurSpan + 2;            // the TMA box is 40 x 38 floats: its inner extent must be a multiple of 16 bytes,
constexpr int kTmaXoff = 1;                         // and so must its start: the box begins at column bx - 4 (an unaligned inner
                                                    // coordinate is an illegal-instruction fault), tile column 1 = frame column bx - 3
constexpr unsigned kTmaTileBytes = kBlurSpan * kTmaPitch * sizeof(float);
struct SsnaPostParams {
  RenderParams p;                         // as for shade_pass; p.zbuf is set by the kernel
  BlurParams b;                           // width, height, taps, wsum (src / dst / zlimit come from the fields below)
  float *zbuf[2];                         // ping-pong; pass i reads zbuf[i & 1]
  float zlimit[YV_BLURZ_PASSES];
  unsigned int *counters;                 // [YV_BLURZ_PASSES], zero on entry, zero again on exit
  int tiles_x, tiles_y;                   // 32x32 blur tiles
  alignas(64) CUtensorMap tmap[2];        // 2-D maps of zbuf[0] / zbuf[1], box 40 x 38, out-of-frame elements read as 0
};

// ---- TMA / mbarrier plumbing (PTX: cp.async.bulk.tensor, mbarrier) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}"
      :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one 2-D tile: box corner (x, y) in elements, may lie outside the tensor (those elements arrive as zeros)
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int x, int y, unsigned long long *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               :: "r"(smem_u32(smem_dst)), "l"((unsigned long long)map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ssna_post: the five BlurZ passes and the ShadeSimple pass as ONE persistent cooperative launch. A grid that exactly fills
// the GPU pulls 32x32 tiles from one atomic counter per pass, passes are separated by grid barriers, the pixels are shaded
// by the same CTAs after the last barrier. With use_tma the staging of a tile — 38 % of a BlurZ pass's stall samples sit on
// the global -> shared copy in front of the barrier (profiles/r02_prof_ssna.lines.txt) — is taken off the critical path:
// one thread issues a 2-D TMA load of the NEXT tile into the other shared-memory buffer (the hardware fills what lies
// outside the frame with zeros, completion is counted on an mbarrier) while the CTA computes the current one.
// Arithmetic per pixel is blur_compute / shade_pixel unchanged.
template <bool TMA>
__global__ void __launch_bounds__(256, YV_POST_MINB) ssna_post(const __grid_constant__ SsnaPostParams q) {
  // TMA destinations / the plain tile; every buffer starts on a 128-byte boundary (38 * 40 * 4 = 6080 bytes is no multiple)
  constexpr int kTileFloats = (kBlurSpan * kTmaPitch + 31) / 32 * 32;
  __shared__ alignas(128) float tile_mem[TMA ? 2 : 1][kTileFloats];
  typedef float TmaTile[kBlurSpan][kTmaPitch];
  __shared__ alignas(8) unsigned long long mbar[2];
  __shared__ unsigned int s_next[2];
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const unsigned int ntiles = (unsigned int)(q.tiles_x * q.tiles_y);
  const unsigned int tx_n = (unsigned int)q.tiles_x;
  const float kInvalid = __int_as_float(0x7f800000);
  if (TMA) {
    if (threadIdx.x == 0) {
      mbar_init(&mbar[0], 1u); mbar_init(&mbar[1], 1u);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  unsigned int phase0 = 0u, phase1 = 0u;                   // parity of the next completion of mbar[0] / mbar[1]
#pragma unroll 1
  for (int pass = 0; pass < YV_BLURZ_PASSES; ++pass) {
    const float *src = q.zbuf[pass & 1];
    float *dst = q.zbuf[(pass & 1) ^ 1];
    const float zlimit = q.zlimit[pass];
    if (TMA) {
      const CUtensorMap *map = &q.tmap[pass & 1];
      // prologue: this CTA's first tile of the pass
      if (threadIdx.x == 0) {
        const unsigned int t0 = atomicAdd(&q.counters[pass], 1u);
        s_next[0] = t0;
        if (t0 < ntiles) {
          fence_proxy_async();
          mbar_expect_tx(&mbar[0], kTmaTileBytes);
          tma_load_2d(&tile_mem[0][0], map, (int)(t0 % tx_n) * kBlurTile - kBlurApron - kTmaXoff, (int)(t0 / tx_n) * kBlurTile - kBlurApron, &mbar[0]);
        }
      }
      __syncthreads();
      unsigned int t = s_next[0];
      int buf = 0;
      while (t < ntiles) {
        // the next tile's load goes out before this one is touched; its buffer was last read before barrier (B) below
        if (threadIdx.x == 0) {
          const unsigned int t2 = atomicAdd(&q.counters[pass], 1u);
          s_next[buf ^ 1] = t2;
          if (t2 < ntiles) {
            fence_proxy_async();
            mbar_expect_tx(&mbar[buf ^ 1], kTmaTileBytes);
            tma_load_2d(&tile_mem[buf ^ 1][0], map, (int)(t2 % tx_n) * kBlurTile - kBlurApron - kTmaXoff, (int)(t2 / tx_n) * kBlurTile - kBlurApron, &mbar[buf ^ 1]);
          }
        }
        mbar_wait(&mbar[buf], buf ? phase1 : phase0);
        if (buf) phase1 ^= 1u; else phase0 ^= 1u;
        // invalid pixels (0 in the z-buffer, and everything the TMA filled in outside the frame) become +inf
        float *flat = &tile_mem[buf][0];
#pragma unroll
        for (int k = 0; k < (kBlurSpan * kTmaPitch + 255) / 256; ++k) {
          const int i = (int)threadIdx.x + 256 * k;
          if (i < kBlurSpan * kTmaPitch && flat[i] == 0.0f) flat[i] = kInvalid;
        }
        __syncwarp();
        __syncthreads();                                   // (A) the tile is ready; s_next[buf ^ 1] is visible
        blur_compute<kTmaPitch, kTmaXoff>(q.b, dst, zlimit, *reinterpret_cast<const TmaTile *>(&tile_mem[buf][0]), (int)(t % tx_n) * kBlurTile, (int)(t / tx_n) * kBlurTile);
        t = s_next[buf ^ 1];
        __syncwarp();
        __syncthreads();                                   // (B) nobody reads tiles[buf] any more
        buf ^= 1;
      }
      fence_proxy_async();                                 // this pass's stores (generic proxy) before the next pass's TMA reads
    } else {
      BlurTile &tile = *reinterpret_cast<BlurTile *>(&tile_mem[0][0]);
      for (;;) {
        __syncthreads();                                   // the previous tile's readers are done with `tile` and s_next
        if (threadIdx.x == 0) s_next[0] = atomicAdd(&q.counters[pass], 1u);
        __syncthreads();
        const unsigned int t = s_next[0];
        if (t >= ntiles) break;
        blur_tile(q.b, src, dst, zlimit, tile, (int)(t % tx_n) * kBlurTile, (int)(t / tx_n) * kBlurTile);
      }
    }
    grid.sync();
    if (TMA) fence_proxy_async();
  }
  if (blockIdx.x == 0 && threadIdx.x < YV_BLURZ_PASSES) q.counters[threadIdx.x] = 0u;      // for the next frame
  // ShadeSimple with the normals of the blurred z-buffer (five passes: the result sits in zbuf[1])
  const RenderParams &p = q.p;
  const float *zfinal = q.zbuf[YV_BLURZ_PASSES & 1];
  const int bw = (p.width + 31) / 32, bh = p.num_tiles / p.tiles_x;       // 32x8-pixel blocks over the rows of this launch
  for (int blk = (int)blockIdx.x; blk < bw * bh; blk += (int)gridDim.x)
    shade_pixel(p, zfinal, (blk % bw) * 32 + (threadIdx.x & 31), tile_row_y(p, blk / bw) + (threadIdx.x >> 5));
}

// ---------------------------------------------------------------------------------------------
// frame averaging (reaction/report/main.tex:111): per-channel integer sums of n frames, then (sum + n/2) / n
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) accumulate_frame(const uint32_t *rgba, uint4 *sum, uint32_t pixels, int first) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= pixels) return;
  const uint32_t c = rgba[i];
  uint4 a = first ? make_uint4(0u, 0u, 0u, 0u) : sum[i];
  a.x += c & 255u; a.y += (c >> 8) & 255u; a.z += (c >> 16) & 255u; a.w += c >> 24;
  sum[i] = a;
}
__global__ void __launch_bounds__(256) resolve_frames(const uint4 *sum, uint32_t *rgba, uint32_t pixels, uint32_t n) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= pixels) return;
  const uint4 a = sum[i];
  const uint32_t h = n / 2u;
  rgba[i] = ((a.x + h) / n) | (((a.y + h) / n) << 8) | (((a.z + h) / n) << 16) | (((a.w + h) / n) << 24);
}

// ---------------------------------------------------------------------------------------------
// arbitrary rays (DynamicSVO::TraceRay, ore/src/main.cpp:125)
// ---------------------------------------------------------------------------------------------
template <bool RAW>
__global__ void __launch_bounds__(128) trace_rays_kernel(const uint4 *recs, const uint2 *octs, const uint32_t *leaves, uint32_t root_valid, uint32_t root_index,
                                                         const float *pos, const float *dir, uint32_t count,
                                                         uint32_t *node, int32_t *child, float *t) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  NodeFetch<false, false, RAW> fetch = { recs, nullptr, 0u, root_index, 0u, 0u, octs, nullptr };
  LocalStack stk(nullptr);
  LeanState s;
  const float dx = adjust_dir1(dir[3 * i]), dy = adjust_dir1(dir[3 * i + 1]), dz = adjust_dir1(dir[3 * i + 2]);
  bool hit = false;
  if (lean_begin(s, fetch, root_valid != 0u, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], dx, dy, dz)) {
    for (;;) {
      const int r = lean_step<false>(s, fetch, stk, false);
      if (r == kStepHit) { hit = true; break; }
      if (r == kStepMiss) break;
    }
  }
  uint32_t hn = YV_MISS_NODE, data;
  if (hit) fetch.hit_info(leaves, nullptr, s.idx, lean_ch(s) ^ s.flags, s.masks, false, hn, data);
  node[i] = hn;
  child[i] = hit ? (int32_t)(lean_ch(s) ^ s.flags) : YV_MISS_CHILD;
  t[i] = hit ? lean_hit_t(s) : 0.0f;
}

}  // namespace yv
