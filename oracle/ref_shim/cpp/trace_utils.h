/* trace_utils.h — stand-in for the reference's cpp/trace_utils.h (absent).  TEST INFRASTRUCTURE ONLY.
 * The helpers cell/ppu_renderer.cpp and cell/spu/trace_spu.cpp call but do not define. What the snapshot says about
 * each, and what is therefore written here:
 *   minCoord / maxCoord / min_vec_float3 / max_vec_float3 ... smallest / largest of x,y,z (names only)
 *   AdjustDir ......... |d_i| < eps -> copysign(eps, d_i), reaction/report/voxel.tex:316-318 (eps: ours, 1e-6)
 *   SetupTrace ........ mirror the negative axes, slab parameters of the unit cube, accept iff max(t1) < min(t2):
 *                       reaction/report/voxel.tex:305-326, cell/spu/trace_spu.c_:79-93
 *   FindFirstChild .... scalar form of FindFirstChildSPU, cell/spu/trace_spu.cpp:48-68
 *   GoNext ............ scalar form of GoNextSPU, cell/spu/trace_spu.cpp:70-93
 * The SPU build never uses the last two: it runs the reference's own SIMD bodies. */
#ifndef YV_REF_SHIM_TRACE_UTILS_H
#define YV_REF_SHIM_TRACE_UTILS_H

inline float minCoord(const point_3f &p) { float m = p.x < p.y ? p.x : p.y; return m < p.z ? m : p.z; }
inline float maxCoord(const point_3f &p) { float m = p.x > p.y ? p.x : p.y; return m > p.z ? m : p.z; }

inline void AdjustDir(point_3f &dir) {
  const float eps = 1e-6f;
  if (fabsf(dir.x) < eps) dir.x = copysignf(eps, dir.x);
  if (fabsf(dir.y) < eps) dir.y = copysignf(eps, dir.y);
  if (fabsf(dir.z) < eps) dir.z = copysignf(eps, dir.z);
}

inline bool SetupTrace(const point_3f &pos, const point_3f &dir, point_3f &t1, point_3f &t2, uint &dirFlags) {
  float p[3] = { pos.x, pos.y, pos.z }, d[3] = { dir.x, dir.y, dir.z }, a[3], b[3];
  dirFlags = 0;
  for (int i = 0; i < 3; ++i) {
    if (d[i] < 0) { p[i] = 1.0f - p[i]; d[i] = -d[i]; dirFlags |= 1u << i; }
    a[i] = (0.0f - p[i]) / d[i];
    b[i] = (1.0f - p[i]) / d[i];
  }
  t1 = point_3f(a[0], a[1], a[2]);
  t2 = point_3f(b[0], b[1], b[2]);
  return maxCoord(t1) < minCoord(t2);
}

inline int FindFirstChild(point_3f &t1, point_3f &t2) {
  const float tmx = 0.5f * (t1.x + t2.x), tmy = 0.5f * (t1.y + t2.y), tmz = 0.5f * (t1.z + t2.z);
  const float tEnter = maxCoord(t1);
  int childId = 0;
  if (tEnter > tmx) { childId |= 1; t1.x = tmx; } else t2.x = tmx;
  if (tEnter > tmy) { childId |= 2; t1.y = tmy; } else t2.y = tmy;
  if (tEnter > tmz) { childId |= 4; t1.z = tmz; } else t2.z = tmz;
  return childId;
}

inline bool GoNext(int &childId, point_3f &t1, point_3f &t2) {
  int exitPlane;
  if (t2.x > t2.y) exitPlane = (t2.y < t2.z) ? 1 : 2;
  else             exitPlane = (t2.x < t2.z) ? 0 : 2;
  const int mask = 1 << exitPlane;
  if ((childId & mask) != 0) return false;
  childId ^= mask;
  float &a = exitPlane == 0 ? t1.x : (exitPlane == 1 ? t1.y : t1.z);
  float &b = exitPlane == 0 ? t2.x : (exitPlane == 1 ? t2.y : t2.z);
  const float dt = b - a;
  a = b;
  b += dt;
  return true;
}

#ifdef TARGET_SPU
inline float max_vec_float3(vector float v) { float m = v[0] > v[1] ? v[0] : v[1]; return m > v[2] ? m : v[2]; }
inline float min_vec_float3(vector float v) { float m = v[0] < v[1] ? v[0] : v[1]; return m < v[2] ? m : v[2]; }
#endif

#endif
