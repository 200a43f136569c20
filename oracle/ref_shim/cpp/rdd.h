/* rdd.h — stand-in for the reference's cpp/rdd.h (absent).  TEST INFRASTRUCTURE ONLY.
 * RayDirData: the three fields InitRayDir fills (cell/renderer_base.h:58-60). */
#ifndef YV_REF_SHIM_RDD_H
#define YV_REF_SHIM_RDD_H
struct RayDirData {
  point_3f dir0, du, dv;
};
#endif
