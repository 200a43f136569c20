/* yv_oracle.c — CPU oracle for the SVO ray-caster path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference CPU tracer. The reference holds no golden vector for this path; the
 * restatement is pinned by running the reference's own sources instead (see yv_oracle.h): cell/ppu_renderer.cpp and
 * cell/spu/trace_spu.cpp compile unmodified into oracle/_ref behind stand-ins for the absent cpp/ headers, and this file
 * reproduces their frames, hit distances, hit ids and node-fetch counts bit for bit. Only what the snapshot does not
 * contain at all (AdjustDir eps, SetupTrace body, VoxData packing, Shade, LOD, SSNA, secondary rays) stays unpinned.
 *
 * Every function cites the reference lines whose behaviour it restates (paths relative to
 * /root/reference). All ray arithmetic is IEEE-754 binary32, round-to-nearest, and this file
 * must be compiled with -ffp-contract=off so that no FMA is formed (oracle/Makefile).
 */
#include "yv_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z; } v3;

/* cg::point_t operators (nest/include/geometry/primitives/point.h) ----------------------- */

/* scalar product: res = 0; res += a[n]*b[n] for n = 0..2 (point.h:416-425) */
static inline float v3_dot(v3 a, v3 b) {
  float r = 0.0f;
  r += a.x * b.x;
  r += a.y * b.y;
  r += a.z * b.z;
  return r;
}
/* vector product (point.h:435-440) */
static inline v3 v3_cross(v3 a, v3 b) {
  v3 r = { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
  return r;
}
/* normalized(): norm = sqrt(a*a); point /= norm (point.h:462-466,485-499,726-733) */
static inline v3 v3_normalized(v3 a) {
  float n = sqrtf(v3_dot(a, a));
  v3 r = { a.x / n, a.y / n, a.z / n };
  return r;
}
static inline v3 v3_scale(v3 a, float s) { v3 r = { a.x * s, a.y * s, a.z * s }; return r; }
static inline v3 v3_add(v3 a, v3 b) { v3 r = { a.x + b.x, a.y + b.y, a.z + b.z }; return r; }
static inline v3 v3_sub(v3 a, v3 b) { v3 r = { a.x - b.x, a.y - b.y, a.z - b.z }; return r; }
static inline v3 v3_from(const float p[3]) { v3 r = { p[0], p[1], p[2] }; return r; }

static inline float max3(v3 a) { float m = a.x > a.y ? a.x : a.y; return m > a.z ? m : a.z; }
static inline float min3(v3 a) { float m = a.x < a.y ? a.x : a.y; return m < a.z ? m : a.z; }

/* RendererBase::InitRayDir (cell/renderer_base.h:50-61).
 * grad2rad(float) = grad * float(pi/180) (nest/include/geometry/xmath.h:47,57).
 * tan(cg::grad2rad(m_fov / 2)) / m_viewSize.x (:56): the argument is a float, so the C++ overload set gives the
 * float tangent and a float division — confirmed against the reference's header compiled in oracle/_ref
 * (tests/test_reference_renderer.py); the product's host side does the same. */
void yvo_init_ray_dir(const yvo_camera *cam, yvo_raydir *out) {
  const double pi = 3.14159265358979323846;
  v3 vfwd = v3_normalized(v3_from(cam->dir));
  v3 vright = v3_normalized(v3_cross(vfwd, v3_from(cam->up)));
  v3 vup = v3_cross(vright, vfwd);

  float half_deg = cam->fov_deg / 2;
  float half_rad = half_deg * (float)(pi / 180.0);
  float da = tanf(half_rad) / (float)cam->width;

  v3 du = v3_scale(v3_scale(vright, 2.0f), da);     /* 2 * vright * da   (:58) */
  v3 dv = v3_scale(v3_scale(vup, -2.0f), da);       /* -2 * vup * da     (:59) */
  float w = (float)cam->width, h = (float)cam->height;
  /* vfwd - du*W/2 - dv*H/2 (:60) */
  v3 a = v3_scale(du, w); a.x /= 2.0f; a.y /= 2.0f; a.z /= 2.0f;
  v3 b = v3_scale(dv, h); b.x /= 2.0f; b.y /= 2.0f; b.z /= 2.0f;
  v3 d0 = v3_sub(v3_sub(vfwd, a), b);
  out->dir0[0] = d0.x; out->dir0[1] = d0.y; out->dir0[2] = d0.z;
  out->du[0] = du.x; out->du[1] = du.y; out->du[2] = du.z;
  out->dv[0] = dv.x; out->dv[1] = dv.y; out->dv[2] = dv.z;
}

/* AdjustDir (call: ppu_renderer.cpp:57; rule: reaction/report/voxel.tex:316-318) */
static inline float adjust1(float d) {
  return fabsf(d) < YV_DIR_EPS ? copysignf(YV_DIR_EPS, d) : d;
}
static inline v3 adjust_dir(v3 d) {
  v3 r = { adjust1(d.x), adjust1(d.y), adjust1(d.z) };
  return r;
}

/* SetupTrace (call: ppu_renderer.cpp:60; rule: voxel.tex:319-326, trace_spu.c_:84-93) */
static inline int setup_trace(v3 p, v3 d, v3 *t1, v3 *t2, uint32_t *dir_flags) {
  uint32_t f = 0;
  if (d.x < 0) { p.x = 1.0f - p.x; d.x = -d.x; f |= 1u; }
  if (d.y < 0) { p.y = 1.0f - p.y; d.y = -d.y; f |= 2u; }
  if (d.z < 0) { p.z = 1.0f - p.z; d.z = -d.z; f |= 4u; }
  t1->x = (0.0f - p.x) / d.x; t2->x = (1.0f - p.x) / d.x;
  t1->y = (0.0f - p.y) / d.y; t2->y = (1.0f - p.y) / d.y;
  t1->z = (0.0f - p.z) / d.z; t2->z = (1.0f - p.z) / d.z;
  *dir_flags = f;
  return max3(*t1) < min3(*t2);
}

/* FindFirstChild (scalar form of cell/spu/trace_spu.cpp:48-68) */
static inline int find_first_child(v3 *t1, v3 *t2) {
  float tmx = 0.5f * (t1->x + t2->x);
  float tmy = 0.5f * (t1->y + t2->y);
  float tmz = 0.5f * (t1->z + t2->z);
  float t_enter = max3(*t1);
  int ch = 0;
  if (t_enter > tmx) { ch |= 1; t1->x = tmx; } else t2->x = tmx;
  if (t_enter > tmy) { ch |= 2; t1->y = tmy; } else t2->y = tmy;
  if (t_enter > tmz) { ch |= 4; t1->z = tmz; } else t2->z = tmz;
  return ch;
}

/* Test-only switch: 0 = the path's argmin tie order (cell/spu/trace_spu.cpp:75-78, the default and the only order
 * the parity tests use); 1 = the tie order of the reference's scalar prototype (argMin, cell/spu/vector.h:45-59), so
 * that tests/test_reference_prototype.py can compare the traversal with that compiled prototype on rays that cross
 * cell edges exactly. The two differ only when two components of t2 are equal. */
static int g_tie_order = 0;
void yvo_set_tie_order(int order) { g_tie_order = order; }

/* GoNext (scalar form of cell/spu/trace_spu.cpp:70-93) */
static inline int go_next(int *ch, v3 *t1, v3 *t2) {
  int e;
  if (g_tie_order == 1) {
    if (t2->x < t2->y) e = (t2->x < t2->z) ? 0 : 2;
    else               e = (t2->y < t2->z) ? 1 : 2;
  } else
  if (t2->x > t2->y) e = (t2->y < t2->z) ? 1 : 2;
  else               e = (t2->x < t2->z) ? 0 : 2;
  int mask = 1 << e;
  if (*ch & mask) return 0;
  *ch ^= mask;
  float *a = e == 0 ? &t1->x : (e == 1 ? &t1->y : &t1->z);
  float *b = e == 0 ? &t2->x : (e == 1 ? &t2->y : &t2->z);
  float dt = *b - *a;
  *a = *b;
  *b += dt;
  return 1;
}

typedef struct {
  const yv_vox_node *nodes;
  uint32_t count;
  uint32_t dir_flags;
  int front_only;        /* 0 = reference behaviour; 1 = secondary rays: a leaf counts only
                            if its own min(t2) > 0 (no hits behind the origin)              */
  float tlimit;          /* secondary rays: range limit (shadow: distance to the light, AO: ao_max_t). Cells are
                            met in non-decreasing entry parameter, so the ray ends as a miss at the first child
                            entered at t >= tlimit; the outcome (occluded within the limit or not) is unchanged */
  int abort;
  float detail;          /* rp.detailCoef (demo/SVORenderer.cpp:104); 0 = off. A child node whose
                            cube is smaller than detail * entry distance is not entered: it is the
                            hit, child = -1, shaded with VoxNode::data (SVORenderer.cpp:176-179)  */
  /* result (TraceResult, ppu_renderer.cpp:7-12) */
  yv_node_id node;
  int child;
  float t;
  /* counters */
  uint64_t visits, iters;
  /* optional model of the SPU program's software node cache (cell/spu/trace_spu.cpp:15-35): 2048 direct-mapped
     slots indexed by id % 2048; a fetch of an id that is not in its slot is a miss (a DMA) and takes the slot */
  uint32_t *cache_ids;
  uint64_t cache_misses;
} trace_ctx;

/* PPURendererBase::RecTrace (cell/ppu_renderer.cpp:18-41); `level` = depth of node `id` (root 0) */
static int rec_trace(trace_ctx *c, yv_node_id id, v3 t1, v3 t2, int level) {
  if (YV_IS_NULL(id) || min3(t2) <= 0) return 0;            /* :20 */
  if (id >= c->count) return 0;                             /* malformed pool: assert in ref (:54) */
  const yv_vox_node *node = &c->nodes[id];                  /* :23  <- counted node fetch */
  c->visits++;
  if (c->cache_ids) {                                       /* FetchNode, trace_spu.cpp:21-35 */
    const uint32_t ofs = id % YVO_SPU_CACHE_SIZE;
    if (c->cache_ids[ofs] != id) { c->cache_ids[ofs] = id; c->cache_misses++; }
  }
  int ch = find_first_child(&t1, &t2);                      /* :24 */
  for (;;) {
    int cc = ch ^ (int)c->dir_flags;
    if (c->front_only && max3(t1) >= c->tlimit) { c->abort = 1; return 0; }
    c->iters++;
    if (YV_LEAF_FLAG(node->flags, cc) && (!c->front_only || min3(t2) > 0)) {   /* :27 */
      c->node = id; c->child = cc; c->t = max3(t1);         /* :29-31 */
      return 1;
    }
    if (!YV_LEAF_FLAG(node->flags, cc)) {
      const yv_node_id kid = node->child[cc];
      if (c->detail > 0 && !YV_IS_NULL(kid) && min3(t2) > 0) {           /* LOD cut-off (CUDA tracer) */
        float tent = max3(t1);
        float csize = ldexpf(1.0f, -(level + 1));
        if (tent > 0 && csize < c->detail * tent) {
          c->node = kid; c->child = -1; c->t = tent;
          return 1;
        }
      }
      if (rec_trace(c, kid, t1, t2, level + 1)) return 1;                /* :35 */
      if (c->abort) return 0;
    }
    if (!go_next(&ch, &t1, &t2)) return 0;                  /* :38 */
  }
}

static int trace_ray(trace_ctx *c, yv_node_id root, v3 pos, v3 dir) {
  v3 t1, t2;
  c->abort = 0;
  dir = adjust_dir(dir);
  if (!setup_trace(pos, dir, &t1, &t2, &c->dir_flags)) return 0;
  return rec_trace(c, root, t1, t2, 0);
}

/* ---- VoxData unpack + SimpleShader::Shade restatement (spec: include/yv_format.h) ------- */

void yvo_unpack_normal(yv_vox_data data, float n[3]) {
  float fx = (float)((data >> 16) & 255u) / 127.5f - 1.0f;
  float fy = (float)((data >> 24) & 255u) / 127.5f - 1.0f;
  float fz = (1.0f - fabsf(fx)) - fabsf(fy);
  if (fz < 0) {
    float ox = (1.0f - fabsf(fy)) * (fx >= 0 ? 1.0f : -1.0f);
    float oy = (1.0f - fabsf(fx)) * (fy >= 0 ? 1.0f : -1.0f);
    fx = ox; fy = oy;
  }
  float len = sqrtf((fx * fx + fy * fy) + fz * fz);
  n[0] = fx / len; n[1] = fy / len; n[2] = fz / len;
}

static inline void unpack_color(yv_vox_data data, uint32_t c[3]) {
  uint32_t r5 = (data >> 11) & 31u, g6 = (data >> 5) & 63u, b5 = data & 31u;
  c[0] = (r5 << 3) | (r5 >> 2);
  c[1] = (g6 << 2) | (g6 >> 4);
  c[2] = (b5 << 3) | (b5 >> 2);
}

static inline float lambert(const float n[3], v3 P, v3 light) {
  v3 Lv = v3_sub(light, P);
  float len = sqrtf((Lv.x * Lv.x + Lv.y * Lv.y) + Lv.z * Lv.z);
  v3 L = { 0, 0, 0 };
  if (len > 0) { L.x = Lv.x / len; L.y = Lv.y / len; L.z = Lv.z / len; }
  float ndl = (n[0] * L.x + n[1] * L.y) + n[2] * L.z;
  return ndl > 0 ? ndl : 0.0f;
}

static inline void write_color(yv_vox_data data, float k, uint8_t out[4]) {
  uint32_t c[3];
  unpack_color(data, c);
  for (int i = 0; i < 3; ++i) {
    float v = floorf((float)c[i] * k + 0.5f);
    out[i] = (uint8_t)(v < 255.0f ? v : 255.0f);
  }
  out[3] = 255;
}

void yvo_shade(yv_vox_data data, const float dir[3], float t,
               const float viewer[3], const float light[3], float visibility, uint8_t out[4]) {
  float n[3];
  yvo_unpack_normal(data, n);
  v3 P = { viewer[0] + dir[0] * t, viewer[1] + dir[1] * t, viewer[2] + dir[2] * t };
  float d = lambert(n, P, v3_from(light));
  float k = YV_SHADE_AMBIENT + YV_SHADE_DIFFUSE * (d * visibility);
  write_color(data, k, out);
}

/* ShadeSimple with point lights / SetShowNormals, restated from the spec in include/yv_format.h
 * (parameters: demo/SVORenderer.cpp:112-113, demo/Demo.cpp:141-147; kernel body absent) */
static void shade_phong(yv_vox_data data, const float n[3], v3 P, v3 viewer, const yv_light *lights, uint8_t out[4]) {
  uint32_t ci[3];
  unpack_color(data, ci);
  float c[3] = { (float)ci[0], (float)ci[1], (float)ci[2] };
  float acc[3] = { YV_SHADE_AMBIENT * c[0], YV_SHADE_AMBIENT * c[1], YV_SHADE_AMBIENT * c[2] };
  v3 V = v3_sub(viewer, P);
  float vl = sqrtf((V.x * V.x + V.y * V.y) + V.z * V.z);
  if (vl > 0) { V.x /= vl; V.y /= vl; V.z /= vl; } else { V.x = V.y = V.z = 0; }
  for (int i = 0; i < YV_MAX_LIGHTS; ++i) {
    const yv_light *lt = &lights[i];
    if (!lt->enabled) continue;
    v3 Lv = { lt->pos[0] - P.x, lt->pos[1] - P.y, lt->pos[2] - P.z };
    float d = sqrtf((Lv.x * Lv.x + Lv.y * Lv.y) + Lv.z * Lv.z);
    if (!(d > 0)) continue;
    v3 L = { Lv.x / d, Lv.y / d, Lv.z / d };
    float att = 1.0f / ((lt->attenuation[0] + lt->attenuation[1] * d) + (lt->attenuation[2] * d) * d);
    float nl = (n[0] * L.x + n[1] * L.y) + n[2] * L.z;
    float ndl = nl > 0 ? nl : 0.0f;
    float k2 = 2.0f * nl;
    v3 R = { k2 * n[0] - L.x, k2 * n[1] - L.y, k2 * n[2] - L.z };
    float rv = (R.x * V.x + R.y * V.y) + R.z * V.z;
    rv = (nl > 0 && rv > 0) ? rv : 0.0f;
    float s2 = rv * rv, s4 = s2 * s2, s8 = s4 * s4, spec = s8 * s2;
    for (int ch = 0; ch < 3; ++ch)
      acc[ch] = acc[ch] + att * ((lt->diffuse[ch] * ndl) * c[ch] + (lt->specular[ch] * spec) * 255.0f);
  }
  for (int ch = 0; ch < 3; ++ch) {
    float v = floorf(acc[ch] + 0.5f);
    out[ch] = (uint8_t)(v < 255.0f ? (v > 0.0f ? v : 0.0f) : 255.0f);
  }
  out[3] = 255;
}

static void shade_normal(const float n[3], uint8_t out[4]) {
  for (int ch = 0; ch < 3; ++ch) {
    float v = floorf((n[ch] * 0.5f + 0.5f) * 255.0f + 0.5f);
    out[ch] = (uint8_t)(v < 255.0f ? (v > 0.0f ? v : 0.0f) : 255.0f);
  }
  out[3] = 255;
}

/* ---- secondary rays (BASELINE config 4; our definition, mirrored by the kernel) ---------- */

static inline uint32_t hash_u32(uint32_t x) {           /* lowbias32 */
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

/* Unit vector from integer hashes by rejection in the cube: no transcendental functions, so
 * the CPU and GPU agree bit for bit. Up to 8 tries, then +z. */
static v3 hash_unit_vector(uint32_t key) {
  for (int k = 0; k < 8; ++k) {
    uint32_t h = hash_u32(key + 0x9e3779b9U * (uint32_t)k);
    float x = (float)(int)(h & 1023u) - 511.5f;
    float y = (float)(int)((h >> 10) & 1023u) - 511.5f;
    float z = (float)(int)((h >> 20) & 1023u) - 511.5f;
    float l2 = (x * x + y * y) + z * z;
    if (l2 <= 261632.25f && l2 >= 1.0f) {             /* inside the ball of radius 511.5 */
      float l = sqrtf(l2);
      v3 r = { x / l, y / l, z / l };
      return r;
    }
  }
  v3 up = { 0, 0, 1 };
  return up;
}

/* displaced ray origin (reaction/report/main.tex:107-114; spec: include/yv_format.h) */
static v3 jitter_origin(const yvo_camera *cam, size_t offs) {
  v3 pos = v3_from(cam->pos);
  if (!(cam->jitter_amp > 0)) return pos;
  v3 U = hash_unit_vector(hash_u32((uint32_t)offs) ^ hash_u32(cam->jitter_seed ^ YV_JITTER_SALT));
  v3 o = { pos.x + cam->jitter_amp * U.x, pos.y + cam->jitter_amp * U.y, pos.z + cam->jitter_amp * U.z };
  return o;
}

/* ---- SSNA (spec: include/yv_format.h "SSNA") ------------------------------------------------ */

/* SVORenderer::InitBlur (demo/SVORenderer.cpp:55-79) */
void yvo_blur_taps(float *taps) {
  const int K = YV_BLURZ_KERN;
  float h = (float)(K / 2);
  float scale = 2.0f;
  float sum = 0;
  for (int y = 0; y < K; ++y)
    for (int x = 0; x < K; ++x) {
      float tx = scale * ((float)x - h) / h;
      float ty = scale * ((float)y - h) / h;
      tx = tx * tx;
      ty = ty * ty;
      float v = (float)exp(-(double)(tx + ty));
      taps[y * K + x] = v;
      sum += v;
    }
  for (int i = 0; i < K * K; ++i) taps[i] /= sum;
}

static void blur_pass(const float *taps, const float *src, float *dst, int W, int H, float zlimit) {
  const int K = YV_BLURZ_KERN, h = K / 2;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float zc = src[(size_t)y * W + x];
      float out = 0.0f;
      if (zc != 0.0f) {
        float acc = 0.0f, wacc = 0.0f;
        for (int ky = 0; ky < K; ++ky) {
          int qy = y + ky - h;
          if (qy < 0 || qy >= H) continue;
          for (int kx = 0; kx < K; ++kx) {
            int qx = x + kx - h;
            if (qx < 0 || qx >= W) continue;
            float zq = src[(size_t)qy * W + qx];
            if (zq == 0.0f || !(fabsf(zq - zc) < zlimit)) continue;
            float w = taps[ky * K + kx];
            acc = acc + w * zq;
            wacc = wacc + w;
          }
        }
        out = wacc > 0 ? acc / wacc : zc;
      }
      dst[(size_t)y * W + x] = out;
    }
}

/* the BlurZ loop of SVORenderer::Render (demo/SVORenderer.cpp:126-141) */
int yvo_blur_z(const yvo_camera *cam, const float *z0, float *out) {
  const int W = cam->width, H = cam->height;
  float taps[YV_BLURZ_KERN * YV_BLURZ_KERN];
  yvo_blur_taps(taps);
  float *buf[2];
  buf[0] = (float *)malloc((size_t)W * H * sizeof(float));
  buf[1] = (float *)malloc((size_t)W * H * sizeof(float));
  if (!buf[0] || !buf[1]) { free(buf[0]); free(buf[1]); return -1; }
  memcpy(buf[0], z0, (size_t)W * H * sizeof(float));
  const float vox = cam->ssna_voxel_size > 0 ? cam->ssna_voxel_size : YV_SSNA_VOXEL_SIZE;
  const float pixel_ang = (cam->fov_deg * (float)(3.14159265358979323846 / 180.0)) / (float)W;   /* :105 */
  int src = 0;
  float blur_size = 3;
  for (int i = 0; i < YV_BLURZ_PASSES; ++i) {
    float zlimit = (5.0f * vox) / (pixel_ang * blur_size);
    blur_pass(taps, buf[src], buf[1 - src], W, H, zlimit);
    src = 1 - src;
    blur_size += 3;
  }
  memcpy(out, buf[src], (size_t)W * H * sizeof(float));
  free(buf[0]); free(buf[1]);
  return 0;
}

/* camera basis as InitRayDir builds it (cell/renderer_base.h:52-54) */
static void view_basis(const yvo_camera *cam, v3 *fwd, v3 *right, v3 *down, float *d2) {
  *fwd = v3_normalized(v3_from(cam->dir));
  *right = v3_normalized(v3_cross(*fwd, v3_from(cam->up)));
  v3 up = v3_cross(*right, *fwd);
  down->x = -up.x; down->y = -up.y; down->z = -up.z;
  float half_rad = (cam->fov_deg / 2) * (float)(3.14159265358979323846 / 180.0);
  float da = tanf(half_rad) / (float)cam->width;
  *d2 = 2.0f * da;
}

static inline float abs_min_diff(int has_f, float f, int has_b, float b) {      /* ztools.py:21-22,33-34 */
  if (has_f && has_b) return fabsf(f) < fabsf(b) ? f : b;
  if (has_f) return f;
  if (has_b) return b;
  return 0.0f;
}

/* normal from the z-buffer (demo/dumps/ztools.py:23-44), turned to face the camera, in world space */
int yvo_ssna_normal(const yvo_camera *cam, const float *zb, int32_t x, int32_t y, float n[3]) {
  const int W = cam->width, H = cam->height;
  const float z = zb[(size_t)y * W + x];
  if (z == 0.0f) return 0;
  v3 fwd, right, down; float d2;
  view_basis(cam, &fwd, &right, &down, &d2);
  float zr = x + 1 < W ? zb[(size_t)y * W + x + 1] : 0.0f, zl = x > 0 ? zb[(size_t)y * W + x - 1] : 0.0f;
  float zd = y + 1 < H ? zb[(size_t)(y + 1) * W + x] : 0.0f, zu = y > 0 ? zb[(size_t)(y - 1) * W + x] : 0.0f;
  float dx = abs_min_diff(zr != 0.0f, zr - z, zl != 0.0f, z - zl);
  float dy = abs_min_diff(zd != 0.0f, zd - z, zu != 0.0f, z - zu);
  float nvx = (d2 * dx) * z;
  float nvy = (d2 * dy) * z;
  float nvz = -((d2 * d2) * (z * z));
  float len = sqrtf((nvx * nvx + nvy * nvy) + nvz * nvz);
  if (!(len > 0)) return 0;
  n[0] = ((right.x * nvx + down.x * nvy) + fwd.x * nvz) / len;
  n[1] = ((right.y * nvx + down.y * nvy) + fwd.y * nvz) / len;
  n[2] = ((right.z * nvx + down.z * nvy) + fwd.z * nvz) / len;
  return 1;
}

typedef struct {
  const yv_vox_node *nodes; uint32_t count; yv_node_id root;
  const yvo_camera *cam; const yvo_secondary *sec; yvo_raydir rdd;
  uint32_t *ssna_data; float *ssna_t, *ssna_z;     /* SSNA: per-pixel VoxData, t and view-space z of the hit */
  int32_t y0, y1;
  uint32_t *hit_node; int32_t *hit_child; float *hit_t; uint8_t *rgba; uint32_t *visits;
  yvo_stats stats;
} strip_job;

/* PPURendererBase::RenderRect (cell/ppu_renderer.cpp:43-70) over rows [y0,y1) */
static void *render_strip(void *arg) {
  strip_job *j = (strip_job *)arg;
  const int W = j->cam->width;
  const v3 pos = v3_from(j->cam->pos);
  const v3 dir0 = v3_from(j->rdd.dir0), du = v3_from(j->rdd.du), dv = v3_from(j->rdd.dv);
  const yvo_secondary *sec = j->sec;
  const int want_sec = sec && (sec->shadow || sec->ao_samples > 0);
  trace_ctx c;
  memset(&c, 0, sizeof c);
  c.nodes = j->nodes; c.count = j->count;
  if (j->cam->detail_coef > 0) {
    /* rp.detailCoef = m_detailCoef * grad2rad(m_fov / 2) / m_viewSize.x  (demo/SVORenderer.cpp:104) */
    float half_rad = (j->cam->fov_deg / 2) * (float)(3.14159265358979323846 / 180.0);
    c.detail = (j->cam->detail_coef * half_rad) / (float)W;
  }

  for (int y = j->y0; y < j->y1; ++y) {
    for (int x = 0; x < W; ++x) {
      size_t offs = (size_t)y * (size_t)W + (size_t)x;                       /* :53 */
      uint8_t px[4] = { 0, 0, 0, 0 };                                       /* :54 */
      uint32_t hn = YV_MISS_NODE; int32_t hc = YV_MISS_CHILD; float ht = 0.0f;
      uint64_t v0 = c.visits;

      v3 d = v3_add(v3_add(dir0, v3_scale(du, (float)x)), v3_scale(dv, (float)y));
      d = v3_normalized(d);                                                  /* :56 */
      d = adjust_dir(d);                                                     /* :57 */
      j->stats.rays++;
      c.front_only = 0;
      const v3 org = want_sec ? pos : jitter_origin(j->cam, offs);           /* main.tex:109 */
      if (trace_ray(&c, j->root, org, d)) {                                  /* :60-65 */
        hn = c.node; hc = c.child; ht = c.t;
        j->stats.hits++;
        yv_vox_data data = hc < 0 ? j->nodes[hn].data : j->nodes[hn].child[hc];   /* :67; LOD: node.data */
        float dd[3] = { d.x, d.y, d.z };
        if (j->ssna_z) {                                   /* Trace leaves RayData + z (demo/SVORenderer.cpp:118-126) */
          v3 f = v3_normalized(v3_from(j->cam->dir));
          j->ssna_data[offs] = data; j->ssna_t[offs] = ht;
          j->ssna_z[offs] = ht * ((d.x * f.x + d.y * f.y) + d.z * f.z);
        }
        int any_light = 0;
        for (int li = 0; li < YV_MAX_LIGHTS; ++li) any_light |= j->cam->lights[li].enabled;
        if (!want_sec && (j->cam->show_normals || any_light)) {
          float n[3];
          yvo_unpack_normal(data, n);
          v3 P = { org.x + d.x * ht, org.y + d.y * ht, org.z + d.z * ht };
          if (j->cam->show_normals) shade_normal(n, px);
          else shade_phong(data, n, P, pos, j->cam->lights, px);
        } else if (!want_sec) {
          float oo[3] = { org.x, org.y, org.z };
          yvo_shade(data, dd, ht, oo, j->cam->pos, 1.0f, px);                /* light = eye (:33-34) */
        } else {
          float n[3];
          yvo_unpack_normal(data, n);
          v3 P = { pos.x + d.x * ht, pos.y + d.y * ht, pos.z + d.z * ht };
          v3 O = { P.x + n[0] * sec->voxel_size, P.y + n[1] * sec->voxel_size,
                   P.z + n[2] * sec->voxel_size };
          v3 light = sec->shadow ? v3_from(sec->light_pos) : pos;   /* no shadow: head light */
          float dl = lambert(n, P, light);
          float vis = 1.0f;
          c.front_only = 1;
          if (sec->shadow) {
            v3 Lv = v3_sub(light, O);
            float len = sqrtf(v3_dot(Lv, Lv));
            if (len > 0) {
              v3 sd = { Lv.x / len, Lv.y / len, Lv.z / len };
              j->stats.rays++;
              c.tlimit = len;
              if (trace_ray(&c, j->root, O, sd) && c.t > 0 && c.t < len) vis = 0.0f;
            }
          }
          float ao = 1.0f;
          if (sec->ao_samples > 0) {
            int occ = 0;
            for (int s = 0; s < sec->ao_samples; ++s) {
              uint32_t key = hash_u32((uint32_t)offs * 16u + (uint32_t)s) ^ hash_u32(sec->seed);
              v3 U = hash_unit_vector(key);
              v3 D = { n[0] + U.x, n[1] + U.y, n[2] + U.z };
              float l2 = v3_dot(D, D);
              if (l2 < 1e-6f) { D.x = n[0]; D.y = n[1]; D.z = n[2]; }
              else { float l = sqrtf(l2); D.x /= l; D.y /= l; D.z /= l; }
              j->stats.rays++;
              c.tlimit = sec->ao_max_t;
              if (trace_ray(&c, j->root, O, D) && c.t > 0 && c.t < sec->ao_max_t) occ++;
            }
            ao = 1.0f - (float)occ / (float)sec->ao_samples;
          }
          float k = (YV_SHADE_AMBIENT + YV_SHADE_DIFFUSE * (dl * vis)) * ao;
          write_color(data, k, px);
        }
      }
      if (j->hit_node)  j->hit_node[offs] = hn;
      if (j->hit_child) j->hit_child[offs] = hc;
      if (j->hit_t)     j->hit_t[offs] = ht;
      if (j->rgba)      memcpy(j->rgba + 4 * offs, px, 4);
      if (j->visits)    j->visits[offs] = (uint32_t)(c.visits - v0);
    }
  }
  j->stats.node_visits = c.visits;
  j->stats.iterations = c.iters;
  return NULL;
}

int yvo_render(const yv_vox_node *nodes, uint32_t node_count, yv_node_id root,
               const yvo_camera *cam, const yvo_secondary *sec,
               int32_t y0, int32_t y1, int32_t threads,
               uint32_t *hit_node, int32_t *hit_child, float *hit_t,
               uint8_t *rgba, uint32_t *visits_per_ray, yvo_stats *stats) {
  if (!cam || cam->width <= 0 || cam->height <= 0) return -1;
  if (!nodes && !YV_IS_NULL(root)) return -1;
  if (y0 < 0) y0 = 0;
  if (y1 > cam->height) y1 = cam->height;
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  int rows = y1 - y0;
  if (rows <= 0) { if (stats) memset(stats, 0, sizeof *stats); return 0; }
  if (threads > rows) threads = rows;

  const int want_sec = sec && (sec->shadow || sec->ao_samples > 0);
  const int ssna = cam->ssna && !want_sec && rgba;
  uint32_t *ssna_data = NULL; float *ssna_t = NULL, *ssna_z = NULL, *ssna_zb = NULL;
  if (ssna) {
    if (y0 != 0 || y1 != cam->height) return -2;                 /* SSNA needs the whole frame */
    size_t n = (size_t)cam->width * (size_t)cam->height;
    ssna_data = (uint32_t *)calloc(n, 4); ssna_t = (float *)calloc(n, 4);
    ssna_z = (float *)calloc(n, 4); ssna_zb = (float *)calloc(n, 4);
    if (!ssna_data || !ssna_t || !ssna_z || !ssna_zb) { free(ssna_data); free(ssna_t); free(ssna_z); free(ssna_zb); return -1; }
  }

  strip_job *jobs = (strip_job *)calloc((size_t)threads, sizeof(strip_job));
  pthread_t *tids = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
  yvo_raydir rdd;
  yvo_init_ray_dir(cam, &rdd);
  int ystep = rows / threads;                                   /* ppu_renderer.cpp:130 */
  for (int i = 0; i < threads; ++i) {
    strip_job *j = &jobs[i];
    j->nodes = nodes; j->count = node_count; j->root = root; j->cam = cam; j->sec = sec; j->rdd = rdd;
    j->y0 = y0 + ystep * i;
    j->y1 = (i == threads - 1) ? y1 : y0 + ystep * (i + 1);
    j->hit_node = hit_node; j->hit_child = hit_child; j->hit_t = hit_t; j->rgba = rgba;
    j->visits = visits_per_ray;
    j->ssna_data = ssna_data; j->ssna_t = ssna_t; j->ssna_z = ssna_z;
  }
  if (threads == 1) render_strip(&jobs[0]);
  else {
    for (int i = 0; i < threads; ++i) pthread_create(&tids[i], NULL, render_strip, &jobs[i]);
    for (int i = 0; i < threads; ++i) pthread_join(tids[i], NULL);
  }
  if (stats) {
    memset(stats, 0, sizeof *stats);
    for (int i = 0; i < threads; ++i) {
      stats->rays += jobs[i].stats.rays; stats->node_visits += jobs[i].stats.node_visits;
      stats->iterations += jobs[i].stats.iterations; stats->hits += jobs[i].stats.hits;
    }
  }
  free(jobs); free(tids);
  if (ssna) {                       /* BlurZ x5, then ShadeSimple with the z-buffer (demo/SVORenderer.cpp:126-147) */
    const int W = cam->width, H = cam->height;
    int rc = yvo_blur_z(cam, ssna_z, ssna_zb);
    const v3 pos = v3_from(cam->pos);
    const v3 dir0 = v3_from(rdd.dir0), du = v3_from(rdd.du), dv = v3_from(rdd.dv);
    int any_light = 0;
    for (int li = 0; li < YV_MAX_LIGHTS; ++li) any_light |= cam->lights[li].enabled;
    for (int y = 0; y < H && rc == 0; ++y)
      for (int x = 0; x < W; ++x) {
        size_t offs = (size_t)y * (size_t)W + (size_t)x;
        if (rgba[4 * offs + 3] == 0) continue;                   /* miss */
        yv_vox_data data = ssna_data[offs];
        float ht = ssna_t[offs], n[3];
        if (!yvo_ssna_normal(cam, ssna_zb, x, y, n)) yvo_unpack_normal(data, n);
        v3 d = v3_add(v3_add(dir0, v3_scale(du, (float)x)), v3_scale(dv, (float)y));
        d = adjust_dir(v3_normalized(d));
        v3 org = jitter_origin(cam, offs);
        v3 P = { org.x + d.x * ht, org.y + d.y * ht, org.z + d.z * ht };
        uint8_t *px = rgba + 4 * offs;
        if (cam->show_normals) shade_normal(n, px);
        else if (any_light) shade_phong(data, n, P, pos, cam->lights, px);
        else write_color(data, YV_SHADE_AMBIENT + YV_SHADE_DIFFUSE * (lambert(n, P, pos) * 1.0f), px);
      }
    free(ssna_data); free(ssna_t); free(ssna_z); free(ssna_zb);
    if (rc) return rc;
  }
  return 0;
}

/* TreadedRenderer::RenderFrame exactly as written (ppu_renderer.cpp:121-144) */
int yvo_render_threaded_ref(const yv_vox_node *nodes, uint32_t node_count, yv_node_id root,
                            const yvo_camera *cam, uint8_t *rgba) {
  const int ThreadNum = 4;
  int ystep = cam->height / ThreadNum;
  return yvo_render(nodes, node_count, root, cam, NULL, 0, ystep * ThreadNum, ThreadNum,
                    NULL, NULL, NULL, rgba, NULL, NULL);
}

int yvo_trace_ray(const yv_vox_node *nodes, uint32_t node_count, yv_node_id root,
                  const float pos[3], const float dir[3],
                  uint32_t *hit_node, int32_t *hit_child, float *hit_t) {
  trace_ctx c;
  memset(&c, 0, sizeof c);
  c.nodes = nodes; c.count = node_count;
  int hit = trace_ray(&c, root, v3_from(pos), v3_from(dir));
  if (hit_node)  *hit_node = hit ? c.node : YV_MISS_NODE;
  if (hit_child) *hit_child = hit ? c.child : YV_MISS_CHILD;
  if (hit_t)     *hit_t = hit ? c.t : 0.0f;
  return hit;
}

/* The SPU program's node traffic (cell/spu/trace_spu.cpp): one run with blockStart 0 / blockStride 1 over the
 * viewSize / 16 blocks (:162-168), pixels of a block row by row (:123-124), the cache cleared to EmptyNode at the
 * start of the run (:155-156). Returns the fetchCount / missCount the program prints (:179). */
int yvo_spu_cache_model(const yv_vox_node *nodes, uint32_t node_count, yv_node_id root, const yvo_camera *cam,
                        uint64_t *fetches, uint64_t *misses) {
  if (!nodes || !cam || cam->width <= 0 || cam->height <= 0) return -1;
  yvo_raydir rdd;
  yvo_init_ray_dir(cam, &rdd);
  const v3 pos = v3_from(cam->pos), dir0 = v3_from(rdd.dir0), du = v3_from(rdd.du), dv = v3_from(rdd.dv);
  uint32_t *ids = (uint32_t *)malloc(YVO_SPU_CACHE_SIZE * sizeof(uint32_t));
  if (!ids) return -2;
  for (uint32_t i = 0; i < YVO_SPU_CACHE_SIZE; ++i) ids[i] = YV_EMPTY_NODE;
  trace_ctx c;
  memset(&c, 0, sizeof c);
  c.nodes = nodes; c.count = node_count; c.cache_ids = ids;
  const int B = 16, gx = cam->width / B, gy = cam->height / B;            /* BlockSize, trace_spu.h:5 */
  for (int block = 0; block < gx * gy; ++block) {
    const int bx = block % gx, by = block / gx;
    for (int y = 0; y < B; ++y)
      for (int x = 0; x < B; ++x) {
        v3 d = v3_add(v3_add(dir0, v3_scale(du, (float)(bx * B + x))), v3_scale(dv, (float)(by * B + y)));
        d = adjust_dir(v3_normalized(d));
        trace_ray(&c, root, pos, d);
      }
  }
  if (fetches) *fetches = c.visits;
  if (misses) *misses = c.cache_misses;
  free(ids);
  return 0;
}

/* SVOData::Load (cell/svodata.h:31-50): root, two discarded words, count, raw nodes. */
int yvo_load_vox(const char *path, yv_node_id *root, uint32_t *count, yv_vox_node **nodes) {
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  uint32_t hdr[4];
  if (fread(hdr, 4, 4, f) != 4) { fclose(f); return -2; }
  *root = hdr[0];
  *count = hdr[3];
  *nodes = (yv_vox_node *)malloc((size_t)hdr[3] * sizeof(yv_vox_node) + 1);
  if (!*nodes) { fclose(f); return -3; }
  size_t got = fread(*nodes, sizeof(yv_vox_node), hdr[3], f);
  fclose(f);
  if (got != hdr[3]) { free(*nodes); *nodes = NULL; return -4; }
  return 0;
}

void yvo_free(void *p) { free(p); }
