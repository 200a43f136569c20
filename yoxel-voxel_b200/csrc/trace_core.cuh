// trace_core.cuh — per-ray arithmetic of the SVO ray caster: ray generation, slab set-up,
// iterative (explicit-stack) octree descent over the packed records, VoxData decode and shading.
//
// Every float operation that can influence a traversal decision or a pixel value goes through
// YV_F* wrappers: on the device they are the round-to-nearest intrinsics (__fadd_rn, ... — never
// contracted into FMA), on the host (tests/emu build only, to check the state machine without a
// GPU) they are plain operators compiled with -ffp-contract=off. Divide and square root are the
// IEEE correctly-rounded forms on both sides, so hit ids are bit-exact against the CPU oracle.
//
// Reference behaviour restated here (paths relative to the znah/yoxel-voxel tree):
//   ray generation ...... cell/ppu_renderer.cpp:56-57
//   AdjustDir ........... reaction/report/voxel.tex:316-318
//   SetupTrace .......... reaction/report/voxel.tex:319-326, cell/spu/trace_spu.c_:84-93
//   FindFirstChild ...... cell/spu/trace_spu.cpp:48-68
//   GoNext .............. cell/spu/trace_spu.cpp:70-93
//   RecTrace ............ cell/ppu_renderer.cpp:18-41  (recursion -> explicit stack, see trace_step)
//   Shade ............... include/yv_format.h (body absent from the snapshot)
#pragma once

#include <cstdint>
#include <cstring>
#include <cmath>

#include "../../include/yv_format.h"

#if defined(__CUDACC__)
#define YV_HD __host__ __device__ __forceinline__
#elif defined(YV_TEST_HOST_BUILD)
#define YV_HD inline          // tests/emu only: checks the state machine against the oracle without a GPU
#else
#error "trace_core.cuh is device code; a host build exists only for tests/emu (define YV_TEST_HOST_BUILD there)"
#endif

#if defined(__CUDA_ARCH__)
#define YV_FADD(a, b) __fadd_rn((a), (b))
#define YV_FSUB(a, b) __fsub_rn((a), (b))
#define YV_FMUL(a, b) __fmul_rn((a), (b))
#define YV_FDIV(a, b) __fdiv_rn((a), (b))
#define YV_FSQRT(a)   __fsqrt_rn((a))
#define YV_POPC(a)    __popc((a))
#else
#define YV_FADD(a, b) ((a) + (b))
#define YV_FSUB(a, b) ((a) - (b))
#define YV_FMUL(a, b) ((a) * (b))
#define YV_FDIV(a, b) ((a) / (b))
#define YV_FSQRT(a)   sqrtf((a))
#define YV_POPC(a)    __builtin_popcount((a))
#endif

namespace yv {

constexpr int kMaxStack = 23;          // supports trees up to 24 levels below the root

struct Rec { uint32_t child_base, leaf_base, masks, orig_id; };   // == PackedRecord == uint4

struct RayState {
  float t1x, t1y, t1z;   // slab entry parameters of the current child cube
  float t2x, t2y, t2z;   // slab exit parameters
  uint32_t idx;          // packed index of the current node
  uint32_t ch;           // logical (mirrored-space) child index 0..7
  uint32_t flags;        // dirFlags: axes along which the ray was mirrored
  int sp;                // stack entries in use
};

enum : int { kStepContinue = 0, kStepHit = 1, kStepMiss = 2 };

YV_HD float max3f(float a, float b, float c) { float m = a > b ? a : b; return m > c ? m : c; }
YV_HD float min3f(float a, float b, float c) { float m = a < b ? a : b; return m < c ? m : c; }

// dir = normalized(dir0 + du*x + dv*y); AdjustDir(dir)      (ppu_renderer.cpp:56-57)
YV_HD void primary_dir(const float dir0[3], const float du[3], const float dv[3], int x, int y,
                       float &dx, float &dy, float &dz) {
  const float fx = (float)x, fy = (float)y;
  float ax = YV_FADD(YV_FADD(dir0[0], YV_FMUL(du[0], fx)), YV_FMUL(dv[0], fy));
  float ay = YV_FADD(YV_FADD(dir0[1], YV_FMUL(du[1], fx)), YV_FMUL(dv[1], fy));
  float az = YV_FADD(YV_FADD(dir0[2], YV_FMUL(du[2], fx)), YV_FMUL(dv[2], fy));
  float n = YV_FSQRT(YV_FADD(YV_FADD(YV_FMUL(ax, ax), YV_FMUL(ay, ay)), YV_FMUL(az, az)));
  dx = YV_FDIV(ax, n); dy = YV_FDIV(ay, n); dz = YV_FDIV(az, n);
}

YV_HD float adjust_dir1(float d) { return fabsf(d) < YV_DIR_EPS ? copysignf(YV_DIR_EPS, d) : d; }

// SetupTrace: mirror negative axes, slab parameters of the unit cube. Returns max(t1) < min(t2).
YV_HD bool setup_trace(float px, float py, float pz, float dx, float dy, float dz, RayState &s) {
  uint32_t f = 0;
  if (dx < 0) { px = YV_FSUB(1.0f, px); dx = -dx; f |= 1u; }
  if (dy < 0) { py = YV_FSUB(1.0f, py); dy = -dy; f |= 2u; }
  if (dz < 0) { pz = YV_FSUB(1.0f, pz); dz = -dz; f |= 4u; }
  s.t1x = YV_FDIV(YV_FSUB(0.0f, px), dx); s.t2x = YV_FDIV(YV_FSUB(1.0f, px), dx);
  s.t1y = YV_FDIV(YV_FSUB(0.0f, py), dy); s.t2y = YV_FDIV(YV_FSUB(1.0f, py), dy);
  s.t1z = YV_FDIV(YV_FSUB(0.0f, pz), dz); s.t2z = YV_FDIV(YV_FSUB(1.0f, pz), dz);
  s.flags = f;
  s.sp = 0;
  return max3f(s.t1x, s.t1y, s.t1z) < min3f(s.t2x, s.t2y, s.t2z);
}

// FindFirstChild: narrow (t1,t2) to the first child's interval and return its logical index.
YV_HD void find_first_child(RayState &s) {
  const float tmx = YV_FMUL(0.5f, YV_FADD(s.t1x, s.t2x));
  const float tmy = YV_FMUL(0.5f, YV_FADD(s.t1y, s.t2y));
  const float tmz = YV_FMUL(0.5f, YV_FADD(s.t1z, s.t2z));
  const float te = max3f(s.t1x, s.t1y, s.t1z);
  uint32_t ch = 0;
  if (te > tmx) { ch |= 1u; s.t1x = tmx; } else s.t2x = tmx;
  if (te > tmy) { ch |= 2u; s.t1y = tmy; } else s.t2y = tmy;
  if (te > tmz) { ch |= 4u; s.t1z = tmz; } else s.t2z = tmz;
  s.ch = ch;
}

// Explicit stack entry: the parent's state *after* its GoNext, i.e. what the recursion would
// resume with when RecTrace(child) returns false (ppu_renderer.cpp:35-38). An entry is pushed
// only if the parent still has a sibling to visit, so a pop never lands on an exhausted node.
struct StackEntry { float t1x, t1y, t1z; uint32_t idx; float t2x, t2y, t2z; uint32_t ch; };

// One iteration of RecTrace's child loop, with the recursion unrolled onto `stk`.
//   front_only = false : reference behaviour (leaf test precedes the child's t2 > 0 test,
//                        so a leaf behind the origin can be reported with t < 0 — SURVEY §8a10)
//   front_only = true  : secondary rays: a leaf counts only if its own min(t2) > 0
// (a run-time flag, so primary and secondary rays of one warp share a single instruction stream)
// fetch.get(idx, visit) returns the packed record (visit = false for the re-fetch after a pop).
template <class Fetch, class Stack>
YV_HD int trace_step(RayState &s, Rec &rec, const Fetch &fetch, Stack &stk, const bool front_only) {
  const uint32_t c = s.ch ^ s.flags;
  const uint32_t bit = 1u << c;
  const float t2min = min3f(s.t2x, s.t2y, s.t2z);
  if ((rec.masks & bit) && (!front_only || t2min > 0.0f)) return kStepHit;        // :27-33

  // would RecTrace(child) pass its entry test (:20) and fetch a node?
  const bool descend = (((rec.masks >> 8) & bit) != 0u) && (t2min > 0.0f);

  // GoNext on the current state (:38), needed by both the push and the plain advance
  const uint32_t e = (s.t2x > s.t2y) ? ((s.t2y < s.t2z) ? 1u : 2u) : ((s.t2x < s.t2z) ? 0u : 2u);
  const bool can_adv = (s.ch & (1u << e)) == 0u;
  const float a = e == 0u ? s.t1x : (e == 1u ? s.t1y : s.t1z);
  const float b = e == 0u ? s.t2x : (e == 1u ? s.t2y : s.t2z);
  const float nb = YV_FADD(b, YV_FSUB(b, a));          // t2[e] += (t2[e] - t1[e])
  const float n1x = e == 0u ? b : s.t1x, n1y = e == 1u ? b : s.t1y, n1z = e == 2u ? b : s.t1z;
  const float n2x = e == 0u ? nb : s.t2x, n2y = e == 1u ? nb : s.t2y, n2z = e == 2u ? nb : s.t2z;
  const uint32_t nch = s.ch ^ (1u << e);

  if (descend) {
    if (can_adv) {
      StackEntry en = { n1x, n1y, n1z, s.idx, n2x, n2y, n2z, nch };
      stk.push(s.sp, en);
      ++s.sp;
    }
    s.idx = rec.child_base + (uint32_t)YV_POPC((rec.masks >> 8) & (bit - 1u));
    rec = fetch.get(s.idx, true);                                                   // :23
    find_first_child(s);                                                            // :24
    return kStepContinue;
  }
  if (can_adv) {
    s.t1x = n1x; s.t1y = n1y; s.t1z = n1z; s.t2x = n2x; s.t2y = n2y; s.t2z = n2z; s.ch = nch;
    return kStepContinue;
  }
  if (s.sp == 0) return kStepMiss;
  --s.sp;
  const StackEntry en = stk.pop(s.sp);
  s.t1x = en.t1x; s.t1y = en.t1y; s.t1z = en.t1z; s.t2x = en.t2x; s.t2y = en.t2y; s.t2z = en.t2z;
  s.idx = en.idx; s.ch = en.ch;
  rec = fetch.get(s.idx, false);   // re-fetch of the parent (not an algorithmic node visit)
  return kStepContinue;
}

// ---- lean form (what render_frame runs) ----------------------------------------------------------
// Same decisions and the same float operations as trace_step, arranged for instruction count:
//   * the exit parameter an axis takes when it is stepped, N = t2 + (t2 - t1) (GoNext,
//     trace_spu.cpp:84-90), is evaluated for all three axes once per node (after FindFirstChild or a
//     pop) instead of being selected per sibling step, so a step is six predicated moves;
//   * argmin(t2) yields the exit axis and min(t2) together;
//   * push stores the parent's registers as they are plus the exit axis; the GoNext the recursion
//     would run after the child returns (ppu_renderer.cpp:38) is applied when the entry is popped;
//   * descend and pop share one node fetch and one evaluation of N.
// Only child_base and the two masks of a record stay in registers; a hit re-reads its record.
struct LeanState {
  float t1x, t1y, t1z;   // entry parameters of the current child cube
  float Tx, Ty, Tz;      // exit parameters (t2)
  float Nx, Ny, Nz;      // t2 + (t2 - t1) per axis, valid while that axis' ch bit is clear
  uint32_t ch, flags, idx;
  uint32_t masks, child_base;
  uint32_t gm_lo, gm_hi; // octant occupancy of the current node's eight children (byte c = child c), culling fetch policies only
  uint32_t pend;         // exit axis (one-hot) of a sibling step chosen but not yet applied; 0 = none
  uint32_t st;           // YV_LEAN_ABC: axes (one-hot, or-ed) stepped inside the current node; 0 otherwise
  uint32_t level;        // depth of the current node (root = 0); only maintained when LOD is on
  float tlimit;          // secondary rays: nothing at or beyond this ray parameter matters (shadow: distance to
                         // the light, AO: ao_max_t); cells are met in non-decreasing entry parameter, so the ray
                         // ends as a miss at the first child entered at t >= tlimit. +inf for primary rays.
  int sp;
};

struct U4 { uint32_t x, y, z, w; };

#if defined(__CUDA_ARCH__)
#define YV_F2U(f) __float_as_uint(f)
#define YV_U2F(u) __uint_as_float(u)
#else
YV_HD uint32_t yv_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
YV_HD float yv_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#define YV_F2U(f) yv_f2u(f)
#define YV_U2F(u) yv_u2f(u)
#endif

// YV_LEAN_ABC (build-time variant, measured in profiles/README.md round 2): a sibling step moves values, it computes
// nothing — t1 <- T, T <- N on the stepped axis, and an axis is stepped at most once per node. So instead of moving
// them (six selects per step), the three values an axis has at node entry stay where they are, A = t1, B = T,
// C = N = B + (B - A), and a 3-bit mask `st` says which axes have been stepped: the current interval of an axis is
// (A, B) before its step and (B, C) after. The exit parameters for the argmin are three selects, the entry parameters
// are selected only where they are used (descent, hit, range limit), and a stack entry is (A, B, ch, st).
#ifndef YV_LEAN_ABC
#define YV_LEAN_ABC 0
#endif

// entry parameters of the current child cube / its logical child index, whatever the representation
YV_HD float lean_t1x(const LeanState &s) { return (YV_LEAN_ABC && (s.st & 1u)) ? s.Tx : s.t1x; }
YV_HD float lean_t1y(const LeanState &s) { return (YV_LEAN_ABC && (s.st & 2u)) ? s.Ty : s.t1y; }
YV_HD float lean_t1z(const LeanState &s) { return (YV_LEAN_ABC && (s.st & 4u)) ? s.Tz : s.t1z; }
YV_HD uint32_t lean_ch(const LeanState &s) { return YV_LEAN_ABC ? (s.ch | s.st) : s.ch; }
// hit distance t = maxCoord(t1) in the reference's a > b ? a : b form (stored, so no FMNMX)
YV_HD float lean_hit_t(const LeanState &s) {
  const float a = lean_t1x(s), b = lean_t1y(s), c = lean_t1z(s);
  const float m = a > b ? a : b;
  return m > c ? m : c;
}

YV_HD void lean_eval_next(LeanState &s) {
  s.Nx = YV_FADD(s.Tx, YV_FSUB(s.Tx, s.t1x));
  s.Ny = YV_FADD(s.Ty, YV_FSUB(s.Ty, s.t1y));
  s.Nz = YV_FADD(s.Tz, YV_FSUB(s.Tz, s.t1z));
}

// The exit axis is carried one-hot (bit 0 = x, 1 = y, 2 = z; 0 = no step) so that "already in the
// upper half" is one AND with ch and the step is six selects — branch-free on purpose: lanes of one
// warp step different axes, selects keep them converged.
// YV_STEP_IMAD (build-time variant, profiles/README.md round 2): the six selects as integer multiply-adds on the float
// bits, x + m * (y - x) with m = 0 / 1 — exact, and issued to the FMA pipe instead of the ALU pipe the kernel saturates.
#ifndef YV_STEP_IMAD
#define YV_STEP_IMAD 0
#endif
YV_HD float lean_isel(uint32_t m, float y, float x) { return YV_U2F(YV_F2U(x) + m * (YV_F2U(y) - YV_F2U(x))); }
YV_HD void lean_apply_step(LeanState &s, const uint32_t ebits) {
#if YV_STEP_IMAD
  const uint32_t mx = ebits & 1u, my = (ebits >> 1) & 1u, mz = ebits >> 2;
  s.t1x = lean_isel(mx, s.Tx, s.t1x); s.Tx = lean_isel(mx, s.Nx, s.Tx);
  s.t1y = lean_isel(my, s.Ty, s.t1y); s.Ty = lean_isel(my, s.Ny, s.Ty);
  s.t1z = lean_isel(mz, s.Tz, s.t1z); s.Tz = lean_isel(mz, s.Nz, s.Tz);
  s.ch |= ebits;
  return;
#endif
  const bool ex = (ebits & 1u) != 0u, ey = (ebits & 2u) != 0u, ez = (ebits & 4u) != 0u;
  s.t1x = ex ? s.Tx : s.t1x; s.Tx = ex ? s.Nx : s.Tx;
  s.t1y = ey ? s.Ty : s.t1y; s.Ty = ey ? s.Ny : s.Ty;
  s.t1z = ez ? s.Tz : s.t1z; s.Tz = ez ? s.Nz : s.Tz;
  s.ch |= ebits;
}

// FindFirstChild on (t1, T); fmaxf is used only where the result is compared, never stored
YV_HD void lean_first_child(LeanState &s) {
  const float tmx = YV_FMUL(0.5f, YV_FADD(s.t1x, s.Tx));
  const float tmy = YV_FMUL(0.5f, YV_FADD(s.t1y, s.Ty));
  const float tmz = YV_FMUL(0.5f, YV_FADD(s.t1z, s.Tz));
  const float te = fmaxf(fmaxf(s.t1x, s.t1y), s.t1z);
  const bool fx = te > tmx, fy = te > tmy, fz = te > tmz;
  s.t1x = fx ? tmx : s.t1x; s.Tx = fx ? s.Tx : tmx;
  s.t1y = fy ? tmy : s.t1y; s.Ty = fy ? s.Ty : tmy;
  s.t1z = fz ? tmz : s.t1z; s.Tz = fz ? s.Tz : tmz;
  s.ch = (fx ? 1u : 0u) | (fy ? 2u : 0u) | (fz ? 4u : 0u);
}

// The Fetch policy hides the node-pool layout:
//   fetch.node(idx, visit, masks, child_base)  one node dereference -> leaf|child masks (+ child base)
//   fetch.child_index(idx, child_base, masks, c)  index of existing child c
//   fetch.root_index()
// packed pool (svo_pack.h): one 16-byte record, children contiguous; raw pool (the reference's 40-byte
// VoxNode array, used for scenes that are being edited): flags word, then the child slot itself.
//
// Octant culling (FetchTraits<Fetch>::kCull; the policy then also returns the node's 64-bit "grandchild mask": byte c =
// which of the eight octants of child node c hold anything, leaf or node). RecTrace enters every existing child node
// the ray's interval reaches (cell/ppu_renderer.cpp:35); on BASELINE config 2, 31 % of those visits find nothing: the
// ray crosses the child through octants that are all empty, and the visit costs a push, a node fetch, one or two sibling
// steps and a pop. Which octants the ray can touch inside the child is known BEFORE entering it: FindFirstChild of the
// child's interval gives the first octant ch0 (arithmetic the visit performs anyway), GoNext only ever sets bits of ch,
// and an axis whose mid-plane parameter tm lies beyond the interval's exit cannot be stepped — so every octant the
// visit could test lies in the box { o : ch0 ⊆ o ⊆ chx }, chx = ch0 | { axes with tm <= exit (+ rounding margin) }.
// If no occupied octant of the child is in that box the visit is skipped; the child is treated as the reference
// treats it after RecTrace(child) returned false. The test is conservative (it never skips a visit that could hit or
// descend: proof and margin below), every float operation of the visits that do happen is unchanged, so hit ids, t and
// pixels stay bit-identical — only the number of node fetches drops (34.9 -> 24.9 per ray on config 2,
// tools/model/cull_model.cpp).
//
// Why chx is a superset of the axes the visit steps: inside the child, an axis i whose ch0 bit is clear has T_i = tm_i
// and is stepped only when it is argmin(T) (trace_spu.cpp:75-90) while some axis j* = argmin of the child's own exit
// parameters is (a) still clear: then tm_i <= tm_j* <= exit, or (b) set: then tm_i <= T_j*, and T_j* is either the
// child's exit parameter itself (set by FindFirstChild) or N_j* = tm + (tm - t1), which differs from it by at most
// 2 ulp of the interval's largest magnitude M. The test uses tm_i <= exit + 1e-6 * M  (8 ulp of M).
YV_HD uint32_t box_mask_stored(uint32_t flags, uint32_t ch0, uint32_t chx) {
  uint32_t box = 0u;
  for (uint32_t o = 0u; o < 8u; ++o)
    if ((o & ch0) == ch0 && (o | chx) == chx) box |= 1u << (o ^ flags);        // logical (mirrored) octant -> stored index
  return box;
}

template <class Fetch> struct FetchTraits { static constexpr bool kGuardDepth = false; static constexpr bool kCull = false; };

template <class Fetch>
YV_HD void lean_load_node(LeanState &s, const Fetch &fetch, const bool visit) {
  if constexpr (FetchTraits<Fetch>::kCull) fetch.node(s.idx, visit, s.masks, s.child_base, s.gm_lo, s.gm_hi);
  else fetch.node(s.idx, visit, s.masks, s.child_base);
}

// SetupTrace + RecTrace's entry test on the root + FindFirstChild in the root (needs no node data).
// Returns false on an immediate miss. The caller loads the root record and evaluates N.
YV_HD bool lean_setup_root(LeanState &s, const bool root_valid,
                           float px, float py, float pz, float dx, float dy, float dz) {
  RayState r;
  if (!setup_trace(px, py, pz, dx, dy, dz, r)) return false;
  if (!root_valid || fminf(fminf(r.t2x, r.t2y), r.t2z) <= 0.0f) return false;
  s.t1x = r.t1x; s.t1y = r.t1y; s.t1z = r.t1z; s.Tx = r.t2x; s.Ty = r.t2y; s.Tz = r.t2z;
  s.flags = r.flags; s.sp = 0; s.idx = 0u; s.pend = 0u; s.st = 0u; s.level = 0u; s.tlimit = __builtin_huge_valf();
  lean_first_child(s);
  return true;
}

template <class Fetch>
YV_HD bool lean_begin(LeanState &s, const Fetch &fetch, const bool root_valid,
                      float px, float py, float pz, float dx, float dy, float dz) {
  if (!lean_setup_root(s, root_valid, px, py, pz, dx, dy, dz)) return false;
  s.idx = fetch.root_index();
  lean_load_node(s, fetch, true);
  lean_eval_next(s);
  return true;
}

// One call = up to YV_STEPS_PER_CALL sibling steps followed by at most one (descend | pop): test the
// current child; while it is empty and a sibling follows, step (register-only); then descend or pop.
// Folding the cheap steps into the call that performs the expensive node entry means a warp issues the
// (descend | pop) block once per node visit instead of once per loop trip.
// A sibling step is *deferred*: the chosen exit axis is parked in s.pend and applied at the top of the
// next trip, which is also where a popped parent's GoNext is applied — one copy of the step code
// serves both.
// Stack: push(sp, U4, U4) / pop(sp, U4&, U4&).
#ifndef YV_STEPS_PER_CALL
#define YV_STEPS_PER_CALL 2
#endif
//
// LOD (SVORenderer::SetDetailCoef, demo/SVORenderer.h:25; rp.detailCoef, demo/SVORenderer.cpp:104): with
// LOD on, a child NODE whose cube is smaller than detail * (distance at which the ray enters it) is not
// entered; it is reported as the hit with child = -1 and shaded with its sub-tree average VoxNode::data
// (endNodeChild < 0 -> node.data, demo/SVORenderer.cpp:176-179). Returns kStepLodHit with s.idx = that node.
enum : int { kStepLodHit = 3 };
// A Fetch policy over a pool that nothing has checked for depth (the raw reference pool of a scene under edit, or a
// .vox / caller-supplied pool traversed in place) sets FetchTraits<Fetch>::kGuardDepth: the traversal then counts
// levels and treats a child node below level kMaxStack as empty, so neither a pool deeper than the explicit stack
// nor a cyclic one can run the stack over or keep a ray descending for ever. (The packed layout is depth-checked
// when it is made.)
#if YV_LEAN_ABC
template <bool LOD, class Fetch, class Stack>
YV_HD int lean_step(LeanState &s, const Fetch &fetch, Stack &stk, const bool front_only, const float detail = 0.0f) {
  constexpr bool GUARD = FetchTraits<Fetch>::kGuardDepth;
  // (no culling form: a kCull policy is traversed as the reference traversal in this variant)
  constexpr bool LEVELS = LOD || GUARD;
  uint32_t bit, e, chc;
  bool descend, can_adv;
  float Tx, Ty, Tz;
#pragma unroll
  for (int k = 0;; ++k) {
    s.st |= s.pend;                                                // deferred GoNext: mark the axis, move nothing
    s.pend = 0u;
    const bool px = (s.st & 1u) != 0u, py = (s.st & 2u) != 0u, pz = (s.st & 4u) != 0u;
    Tx = px ? s.Nx : s.Tx; Ty = py ? s.Ny : s.Ty; Tz = pz ? s.Nz : s.Tz;
    if (front_only && fmaxf(fmaxf(px ? s.Tx : s.t1x, py ? s.Ty : s.t1y), pz ? s.Tz : s.t1z) >= s.tlimit) return kStepMiss;
    chc = s.ch | s.st;
    bit = 1u << (chc ^ s.flags);
    const bool xy = Tx > Ty;
    const bool nz = xy ? (Ty < Tz) : (Tx < Tz);
    e = nz ? (xy ? 2u : 1u) : 4u;                                  // argmin(t2), one-hot, the reference's tie order
    const float tmin = fminf(fminf(Tx, Ty), Tz);                   // compared only
    if (((s.masks & bit) != 0u) && (!front_only || tmin > 0.0f)) return kStepHit;           // :27
    descend = (((s.masks >> 8) & bit) != 0u) && (tmin > 0.0f);                               // :20,:35
    if (GUARD) descend = descend && s.level < (uint32_t)kMaxStack;
    can_adv = (chc & e) == 0u;                                                               // :38
    if (LOD) {
      const float tent = fmaxf(fmaxf(px ? s.Tx : s.t1x, py ? s.Ty : s.t1y), pz ? s.Tz : s.t1z);
      const float lodk = YV_U2F(YV_F2U(detail) + ((s.level + 1u) << 23));
      if (descend && tent > 0.0f && YV_FMUL(tent, lodk) > 1.0f) {
        s.idx = fetch.child_index(s.idx, s.child_base, s.masks, chc ^ s.flags);
        return kStepLodHit;
      }
    }
    if (descend || !can_adv) break;
    s.pend = e;
    if (k + 1 == YV_STEPS_PER_CALL) return kStepContinue;
  }

  if (descend) {
    if (can_adv) {
      const U4 a = { YV_F2U(s.t1x), YV_F2U(s.t1y), YV_F2U(s.t1z), s.idx };
      const U4 b = { YV_F2U(s.Tx), YV_F2U(s.Ty), YV_F2U(s.Tz), s.ch | (e << 3) | (s.st << 6) | (LEVELS ? (s.level << 9) : 0u) };
      stk.push(s.sp, a, b);
      ++s.sp;
    }
    s.idx = fetch.child_index(s.idx, s.child_base, s.masks, chc ^ s.flags);
    if (LEVELS) ++s.level;
    // the child's own interval becomes (t1, T); FindFirstChild narrows it below
    const bool px = (s.st & 1u) != 0u, py = (s.st & 2u) != 0u, pz = (s.st & 4u) != 0u;
    s.t1x = px ? s.Tx : s.t1x; s.t1y = py ? s.Ty : s.t1y; s.t1z = pz ? s.Tz : s.t1z;
    s.Tx = Tx; s.Ty = Ty; s.Tz = Tz;
    s.st = 0u;
  } else {
    if (s.sp == 0) return kStepMiss;
    --s.sp;
    U4 a, b;
    stk.pop(s.sp, a, b);
    s.t1x = YV_U2F(a.x); s.t1y = YV_U2F(a.y); s.t1z = YV_U2F(a.z); s.idx = a.w;
    s.Tx = YV_U2F(b.x); s.Ty = YV_U2F(b.y); s.Tz = YV_U2F(b.z);
    s.ch = b.w & 7u; s.pend = (b.w >> 3) & 7u; s.st = (b.w >> 6) & 7u;       // the parent's GoNext, applied next trip
    if (LEVELS) s.level = b.w >> 9;
  }
  lean_load_node(s, fetch, descend);                                                         // :23
  if (descend) lean_first_child(s);                                                          // :24
  lean_eval_next(s);
  return kStepContinue;
}
#else
template <bool LOD, class Fetch, class Stack>
YV_HD int lean_step(LeanState &s, const Fetch &fetch, Stack &stk, const bool front_only, const float detail = 0.0f) {
  constexpr bool GUARD = FetchTraits<Fetch>::kGuardDepth;
  constexpr bool CULL = FetchTraits<Fetch>::kCull;
  constexpr bool LEVELS = LOD || GUARD;
  uint32_t bit, e;
  bool descend, can_adv;
  float tmx = 0.0f, tmy = 0.0f, tmz = 0.0f;                        // CULL: the child's FindFirstChild, evaluated before entering it
  bool fx = false, fy = false, fz = false;
#pragma unroll
  for (int k = 0;; ++k) {
    lean_apply_step(s, s.pend);                                    // deferred GoNext (no-op when pend == 0)
    s.pend = 0u;
    if (front_only && fmaxf(fmaxf(s.t1x, s.t1y), s.t1z) >= s.tlimit) return kStepMiss;      // secondary rays: range limit
    bit = 1u << (s.ch ^ s.flags);
    const bool xy = s.Tx > s.Ty;
    const bool nz = xy ? (s.Ty < s.Tz) : (s.Tx < s.Tz);
    e = nz ? (xy ? 2u : 1u) : 4u;                                  // argmin(t2), one-hot, the reference's tie order
    const float tmin = fminf(fminf(s.Tx, s.Ty), s.Tz);             // compared only
    if (((s.masks & bit) != 0u) && (!front_only || tmin > 0.0f)) return kStepHit;           // :27
    descend = (((s.masks >> 8) & bit) != 0u) && (tmin > 0.0f);                               // :20,:35
    if (GUARD) descend = descend && s.level < (uint32_t)kMaxStack;
    can_adv = (s.ch & e) == 0u;                                                              // :38
    if (LOD) {
      // child cube edge 2^-(level+1) < detail * t_enter   <=>   t_enter * (detail * 2^(level+1)) > 1
      // (scaling by a power of two commutes with rounding); t_enter is only compared here, the stored
      // hit distance is recomputed in the exact form by the caller.
      const float tent = fmaxf(fmaxf(s.t1x, s.t1y), s.t1z);
      const float lodk = YV_U2F(YV_F2U(detail) + ((s.level + 1u) << 23));
      if (descend && tent > 0.0f && YV_FMUL(tent, lodk) > 1.0f) {
        s.idx = fetch.child_index(s.idx, s.child_base, s.masks, s.ch ^ s.flags);
        return kStepLodHit;
      }
    }
    if constexpr (CULL) if (descend) {
      tmx = YV_FMUL(0.5f, YV_FADD(s.t1x, s.Tx));                   // FindFirstChild of the child (trace_spu.cpp:50-64):
      tmy = YV_FMUL(0.5f, YV_FADD(s.t1y, s.Ty));                   // kept for the entry below, so nothing is computed twice
      tmz = YV_FMUL(0.5f, YV_FADD(s.t1z, s.Tz));
      const float te = fmaxf(fmaxf(s.t1x, s.t1y), s.t1z);
      fx = te > tmx; fy = te > tmy; fz = te > tmz;
      const float mag = fmaxf(fmaxf(fmaxf(fabsf(s.t1x), fabsf(s.t1y)), fabsf(s.t1z)), fmaxf(fmaxf(fabsf(s.Tx), fabsf(s.Ty)), fabsf(s.Tz)));
      const float lim = tmin + 1e-6f * mag;                        // compared only (may contract to an FMA)
      const uint32_t ch0 = (fx ? 1u : 0u) | (fy ? 2u : 0u) | (fz ? 4u : 0u);
      const uint32_t chx = ch0 | (tmx <= lim ? 1u : 0u) | (tmy <= lim ? 2u : 0u) | (tmz <= lim ? 4u : 0u);
      const uint32_t c = s.ch ^ s.flags;
      const uint32_t occ = ((c & 4u) ? s.gm_hi : s.gm_lo) >> ((c & 3u) << 3);
      descend = (fetch.box(s.flags, ch0, chx) & occ & 0xffu) != 0u;
    }
    if (descend || !can_adv) break;
    s.pend = e;
    if (k + 1 == YV_STEPS_PER_CALL) return kStepContinue;
  }

  if (descend) {
    if (can_adv) {
      const U4 a = { YV_F2U(s.t1x), YV_F2U(s.t1y), YV_F2U(s.t1z), s.idx };
      const U4 b = { YV_F2U(s.Tx), YV_F2U(s.Ty), YV_F2U(s.Tz), s.ch | (e << 3) | (LEVELS ? (s.level << 6) : 0u) };
      stk.push(s.sp, a, b);
      ++s.sp;
    }
    s.idx = fetch.child_index(s.idx, s.child_base, s.masks, s.ch ^ s.flags);
    if (LEVELS) ++s.level;
  } else {
    if (s.sp == 0) return kStepMiss;
    --s.sp;
    U4 a, b;
    stk.pop(s.sp, a, b);
    s.t1x = YV_U2F(a.x); s.t1y = YV_U2F(a.y); s.t1z = YV_U2F(a.z); s.idx = a.w;
    s.Tx = YV_U2F(b.x); s.Ty = YV_U2F(b.y); s.Tz = YV_U2F(b.z);
    s.ch = b.w & 7u; s.pend = (b.w >> 3) & 7u;                     // the parent's GoNext, applied next trip
    if (LEVELS) s.level = b.w >> 6;
  }
  lean_load_node(s, fetch, descend);                                                         // :23
  if (descend) {                                                                             // :24
    if (CULL) {
      s.t1x = fx ? tmx : s.t1x; s.Tx = fx ? s.Tx : tmx;
      s.t1y = fy ? tmy : s.t1y; s.Ty = fy ? s.Ty : tmy;
      s.t1z = fz ? tmz : s.t1z; s.Tz = fz ? s.Tz : tmz;
      s.ch = (fx ? 1u : 0u) | (fy ? 2u : 0u) | (fz ? 4u : 0u);
    } else lean_first_child(s);
  }
  lean_eval_next(s);
  return kStepContinue;
}
#endif  // YV_LEAN_ABC

// ---- secondary rays: the descent to the ray's origin, specialised -----------------------------------------------------
// A shadow or AO ray starts one voxel off a surface, i.e. INSIDE the cube, and a leaf counts for it only in front of the
// origin (front_only). Until the traversal reaches the cell that holds the origin, RecTrace's loop does the same thing at
// every level: the children the line crosses behind the origin have min(t2) <= 0, so they are neither hit nor entered
// (cell/ppu_renderer.cpp:20,27 with the front_only rule) — GoNext steps over them — and the first child with
// min(t2) > 0 is the one that contains the origin; if it is a node it is entered. That is 12-13 of the 39 lean_step trips
// an average secondary ray of BASELINE config 4 needs, each a full general-purpose trip (hit test, step, push | pop)
// whose lanes are at different phases. lean_descend_once performs one such level in closed form, for all lanes of a warp
// at once and without the blocks that cannot fire:
//   * GoNext steps exactly the axes whose exit parameter is <= 0, each once, in argmin order; the order does not change
//     the outcome (an axis is stepped t1 <- T, T <- N independently of the others), so they are applied together;
//   * anything irregular — the origin not strictly inside on some axis, an axis that would have to be stepped twice
//     (N <= 0 or already in the upper half: RecTrace leaves the node), a range limit that could fire, a leaf or an empty
//     slot at the origin's child, a depth guard — returns false WITHOUT touching the state, and lean_step carries on from
//     there: the state between two levels is a state lean_step itself passes through (pend = 0).
// Same float operations on the same operands, same node fetches (the counters agree with the oracle's), same stack
// entries as the general loop; tests/emu runs it against the oracle on the CPU.
// front_only = false (a primary ray whose eye is inside the cube — configs 2, 3 and 5): the reference tests a child for a
// leaf BEFORE its t2 > 0 test (:27 before :20, the "leaf behind the eye" quirk), so children behind the eye can be hits;
// the closed form then only takes levels whose node has no leaf child at all.
template <bool LOD, class Fetch, class Stack>
YV_HD bool lean_descend_once(LeanState &s, const Fetch &fetch, Stack &stk, const bool front_only = true) {
#if YV_LEAN_ABC
  return false;
#else
  constexpr bool GUARD = FetchTraits<Fetch>::kGuardDepth;
  constexpr bool LEVELS = LOD || GUARD;
  if (FetchTraits<Fetch>::kCull) return false;                     // the culling traversal decides descents differently
  // pre-condition: node s.idx loaded, FindFirstChild done, N evaluated, pend == 0 (the state after lean_begin or a descent)
  if (!(fmaxf(fmaxf(s.t1x, s.t1y), s.t1z) < 0.0f) || !(s.tlimit > 0.0f) || s.pend != 0u) return false;
  if (!front_only && (s.masks & 0xffu) != 0u) return false;      // (a test of just the children the steps pass through took no
                                                                 // further level on the bench scenes: the eye's cell is empty one level down)
  const bool sx = !(s.Tx > 0.0f), sy = !(s.Ty > 0.0f), sz = !(s.Tz > 0.0f);       // the axes GoNext steps (:38, trace_spu.cpp:75-90)
  const uint32_t S = (sx ? 1u : 0u) | (sy ? 2u : 0u) | (sz ? 4u : 0u);
  if ((S & s.ch) != 0u) return false;                              // already in the upper half there: the ray leaves the node
  if ((sx && !(s.Nx > 0.0f)) || (sy && !(s.Ny > 0.0f)) || (sz && !(s.Nz > 0.0f))) return false;   // ... or would after the step
  const uint32_t ch = s.ch | S;
  const uint32_t bit = 1u << (ch ^ s.flags);
  if ((s.masks & bit) != 0u) return false;                         // a leaf holds the origin: lean_step reports it (:27)
  if (((s.masks >> 8) & bit) == 0u) return false;                  // empty there: lean_step steps on
  if (GUARD && !(s.level < (uint32_t)kMaxStack)) return false;
  // commit the steps, then RecTrace(child) (:35): push the parent if it still has a sibling to offer, enter the child
  s.t1x = sx ? s.Tx : s.t1x; s.Tx = sx ? s.Nx : s.Tx;
  s.t1y = sy ? s.Ty : s.t1y; s.Ty = sy ? s.Ny : s.Ty;
  s.t1z = sz ? s.Tz : s.t1z; s.Tz = sz ? s.Nz : s.Tz;
  s.ch = ch;
  const bool xy = s.Tx > s.Ty;
  const bool nz = xy ? (s.Ty < s.Tz) : (s.Tx < s.Tz);
  const uint32_t e = nz ? (xy ? 2u : 1u) : 4u;                     // argmin(t2), one-hot, the reference's tie order
  if ((ch & e) == 0u) {
    const U4 a = { YV_F2U(s.t1x), YV_F2U(s.t1y), YV_F2U(s.t1z), s.idx };
    const U4 b = { YV_F2U(s.Tx), YV_F2U(s.Ty), YV_F2U(s.Tz), ch | (e << 3) | (LEVELS ? (s.level << 6) : 0u) };
    stk.push(s.sp, a, b);
    ++s.sp;
  }
  s.idx = fetch.child_index(s.idx, s.child_base, s.masks, ch ^ s.flags);
  if (LEVELS) ++s.level;
  lean_load_node(s, fetch, true);                                                            // :23
  lean_first_child(s);                                                                       // :24
  lean_eval_next(s);
  return true;
#endif
}

// Entry test of RecTrace(root): the caller has run setup_trace. Returns false on an immediate miss.
template <class Fetch>
YV_HD bool trace_enter_root(RayState &s, Rec &rec, const Fetch &fetch, bool root_valid) {
  if (!root_valid || min3f(s.t2x, s.t2y, s.t2z) <= 0.0f) return false;               // :20
  s.idx = 0u;
  rec = fetch(0u);
  find_first_child(s);
  return true;
}

// ---- VoxData decode + shading (spec: include/yv_format.h) -----------------------------------

YV_HD void unpack_normal(uint32_t data, float &nx, float &ny, float &nz) {
  float fx = YV_FSUB(YV_FDIV((float)((data >> 16) & 255u), 127.5f), 1.0f);
  float fy = YV_FSUB(YV_FDIV((float)((data >> 24) & 255u), 127.5f), 1.0f);
  float fz = YV_FSUB(YV_FSUB(1.0f, fabsf(fx)), fabsf(fy));
  if (fz < 0) {
    float ox = YV_FMUL(YV_FSUB(1.0f, fabsf(fy)), fx >= 0 ? 1.0f : -1.0f);
    float oy = YV_FMUL(YV_FSUB(1.0f, fabsf(fx)), fy >= 0 ? 1.0f : -1.0f);
    fx = ox; fy = oy;
  }
  float len = YV_FSQRT(YV_FADD(YV_FADD(YV_FMUL(fx, fx), YV_FMUL(fy, fy)), YV_FMUL(fz, fz)));
  nx = YV_FDIV(fx, len); ny = YV_FDIV(fy, len); nz = YV_FDIV(fz, len);
}

// Lambert term max(0, n . normalize(light - P))
YV_HD float lambert(float nx, float ny, float nz, float Px, float Py, float Pz,
                    float lx, float ly, float lz) {
  float vx = YV_FSUB(lx, Px), vy = YV_FSUB(ly, Py), vz = YV_FSUB(lz, Pz);
  float len = YV_FSQRT(YV_FADD(YV_FADD(YV_FMUL(vx, vx), YV_FMUL(vy, vy)), YV_FMUL(vz, vz)));
  float Lx = 0.0f, Ly = 0.0f, Lz = 0.0f;
  if (len > 0) { Lx = YV_FDIV(vx, len); Ly = YV_FDIV(vy, len); Lz = YV_FDIV(vz, len); }
  float ndl = YV_FADD(YV_FADD(YV_FMUL(nx, Lx), YV_FMUL(ny, Ly)), YV_FMUL(nz, Lz));
  return ndl > 0 ? ndl : 0.0f;
}

// colour * k -> RGBA8 word in memory order R,G,B,A (little-endian: R in bits 0..7)
// dx / dy of the SSNA normal: the smaller-magnitude one-sided difference among the defined ones
// (demo/dumps/ztools.py:21-22,33-34; spec in include/yv_format.h)
YV_HD float abs_min_diff(bool has_f, float f, bool has_b, float b) {
  if (has_f && has_b) return fabsf(f) < fabsf(b) ? f : b;
  return has_f ? f : (has_b ? b : 0.0f);
}

YV_HD uint32_t shade_rgba(uint32_t data, float k) {
  const uint32_t r5 = (data >> 11) & 31u, g6 = (data >> 5) & 63u, b5 = data & 31u;
  const uint32_t c0 = (r5 << 3) | (r5 >> 2), c1 = (g6 << 2) | (g6 >> 4), c2 = (b5 << 3) | (b5 >> 2);
  float v0 = floorf(YV_FADD(YV_FMUL((float)c0, k), 0.5f));
  float v1 = floorf(YV_FADD(YV_FMUL((float)c1, k), 0.5f));
  float v2 = floorf(YV_FADD(YV_FMUL((float)c2, k), 0.5f));
  const uint32_t o0 = (uint32_t)(v0 < 255.0f ? v0 : 255.0f);
  const uint32_t o1 = (uint32_t)(v1 < 255.0f ? v1 : 255.0f);
  const uint32_t o2 = (uint32_t)(v2 < 255.0f ? v2 : 255.0f);
  return o0 | (o1 << 8) | (o2 << 16) | 0xff000000u;
}

// ShadeSimple with point lights / SetShowNormals (spec: include/yv_format.h)
YV_HD uint32_t shade_phong(uint32_t data, float nx, float ny, float nz, float Px, float Py, float Pz,
                           const float viewer[3], const yv_light *lights) {
  const uint32_t r5 = (data >> 11) & 31u, g6 = (data >> 5) & 63u, b5 = data & 31u;
  const float c[3] = { (float)((r5 << 3) | (r5 >> 2)), (float)((g6 << 2) | (g6 >> 4)), (float)((b5 << 3) | (b5 >> 2)) };
  float acc[3] = { YV_FMUL(YV_SHADE_AMBIENT, c[0]), YV_FMUL(YV_SHADE_AMBIENT, c[1]), YV_FMUL(YV_SHADE_AMBIENT, c[2]) };
  float Vx = YV_FSUB(viewer[0], Px), Vy = YV_FSUB(viewer[1], Py), Vz = YV_FSUB(viewer[2], Pz);
  const float vl = YV_FSQRT(YV_FADD(YV_FADD(YV_FMUL(Vx, Vx), YV_FMUL(Vy, Vy)), YV_FMUL(Vz, Vz)));
  if (vl > 0) { Vx = YV_FDIV(Vx, vl); Vy = YV_FDIV(Vy, vl); Vz = YV_FDIV(Vz, vl); } else { Vx = Vy = Vz = 0.0f; }
  for (int i = 0; i < YV_MAX_LIGHTS; ++i) {
    if (!lights[i].enabled) continue;
    const float lx = YV_FSUB(lights[i].pos[0], Px), ly = YV_FSUB(lights[i].pos[1], Py), lz = YV_FSUB(lights[i].pos[2], Pz);
    const float d = YV_FSQRT(YV_FADD(YV_FADD(YV_FMUL(lx, lx), YV_FMUL(ly, ly)), YV_FMUL(lz, lz)));
    if (!(d > 0)) continue;
    const float Lx = YV_FDIV(lx, d), Ly = YV_FDIV(ly, d), Lz = YV_FDIV(lz, d);
    const float att = YV_FDIV(1.0f, YV_FADD(YV_FADD(lights[i].attenuation[0], YV_FMUL(lights[i].attenuation[1], d)),
                                             YV_FMUL(YV_FMUL(lights[i].attenuation[2], d), d)));
    const float nl = YV_FADD(YV_FADD(YV_FMUL(nx, Lx), YV_FMUL(ny, Ly)), YV_FMUL(nz, Lz));
    const float ndl = nl > 0 ? nl : 0.0f;
    const float k2 = YV_FMUL(2.0f, nl);
    const float Rx = YV_FSUB(YV_FMUL(k2, nx), Lx), Ry = YV_FSUB(YV_FMUL(k2, ny), Ly), Rz = YV_FSUB(YV_FMUL(k2, nz), Lz);
    float rv = YV_FADD(YV_FADD(YV_FMUL(Rx, Vx), YV_FMUL(Ry, Vy)), YV_FMUL(Rz, Vz));
    rv = (nl > 0 && rv > 0) ? rv : 0.0f;
    const float s2 = YV_FMUL(rv, rv), s4 = YV_FMUL(s2, s2), s8 = YV_FMUL(s4, s4), spec = YV_FMUL(s8, s2);
    for (int ch = 0; ch < 3; ++ch)
      acc[ch] = YV_FADD(acc[ch], YV_FMUL(att, YV_FADD(YV_FMUL(YV_FMUL(lights[i].diffuse[ch], ndl), c[ch]),
                                                       YV_FMUL(YV_FMUL(lights[i].specular[ch], spec), 255.0f))));
  }
  uint32_t out = 0xff000000u;
  for (int ch = 0; ch < 3; ++ch) {
    const float v = floorf(YV_FADD(acc[ch], 0.5f));
    out |= (uint32_t)(v < 255.0f ? (v > 0.0f ? v : 0.0f) : 255.0f) << (8 * ch);
  }
  return out;
}

YV_HD uint32_t shade_normal(float nx, float ny, float nz) {
  const float n[3] = { nx, ny, nz };
  uint32_t out = 0xff000000u;
  for (int ch = 0; ch < 3; ++ch) {
    const float v = floorf(YV_FADD(YV_FMUL(YV_FADD(YV_FMUL(n[ch], 0.5f), 0.5f), 255.0f), 0.5f));
    out |= (uint32_t)(v < 255.0f ? (v > 0.0f ? v : 0.0f) : 255.0f) << (8 * ch);
  }
  return out;
}

// ---- secondary-ray helpers (BASELINE config 4; spec: include/yv_b200.h YV secondary rays) -----

YV_HD uint32_t hash_u32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// unit vector by rejection sampling in a cube of integer lattice points (no transcendentals)
YV_HD void hash_unit_vector(uint32_t key, float &ux, float &uy, float &uz) {
  for (int k = 0; k < 8; ++k) {
    const uint32_t h = hash_u32(key + 0x9e3779b9U * (uint32_t)k);
    const float x = YV_FSUB((float)(int)(h & 1023u), 511.5f);
    const float y = YV_FSUB((float)(int)((h >> 10) & 1023u), 511.5f);
    const float z = YV_FSUB((float)(int)((h >> 20) & 1023u), 511.5f);
    const float l2 = YV_FADD(YV_FADD(YV_FMUL(x, x), YV_FMUL(y, y)), YV_FMUL(z, z));
    if (l2 <= 261632.25f && l2 >= 1.0f) {
      const float l = YV_FSQRT(l2);
      ux = YV_FDIV(x, l); uy = YV_FDIV(y, l); uz = YV_FDIV(z, l);
      return;
    }
  }
  ux = 0.0f; uy = 0.0f; uz = 1.0f;
}

// displaced ray origin (reaction/report/main.tex:107-114; spec: include/yv_format.h "Hiding voxelisation artefacts")
YV_HD void jitter_origin(const float pos[3], float amplitude, uint32_t seed, uint32_t pixel, float &ox, float &oy, float &oz) {
  float ux, uy, uz;
  hash_unit_vector(hash_u32(pixel) ^ hash_u32(seed ^ YV_JITTER_SALT), ux, uy, uz);
  ox = YV_FADD(pos[0], YV_FMUL(amplitude, ux));
  oy = YV_FADD(pos[1], YV_FMUL(amplitude, uy));
  oz = YV_FADD(pos[2], YV_FMUL(amplitude, uz));
}

// cosine-weighted AO direction: normalize(n + U), falling back to n when the sum degenerates
YV_HD void ao_direction(float nx, float ny, float nz, uint32_t pixel, uint32_t sample, uint32_t seed,
                        float &dx, float &dy, float &dz) {
  const uint32_t key = hash_u32(pixel * 16u + sample) ^ hash_u32(seed);
  float ux, uy, uz;
  hash_unit_vector(key, ux, uy, uz);
  float x = YV_FADD(nx, ux), y = YV_FADD(ny, uy), z = YV_FADD(nz, uz);
  const float l2 = YV_FADD(YV_FADD(YV_FMUL(x, x), YV_FMUL(y, y)), YV_FMUL(z, z));
  if (l2 < 1e-6f) { dx = nx; dy = ny; dz = nz; return; }
  const float l = YV_FSQRT(l2);
  dx = YV_FDIV(x, l); dy = YV_FDIV(y, l); dz = YV_FDIV(z, l);
}

}  // namespace yv
