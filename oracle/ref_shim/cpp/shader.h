/* shader.h — stand-in for the reference's cpp/shader.h (listed in cell/cell.vcproj:134, not shipped).
 * TEST INFRASTRUCTURE ONLY. SimpleShader with the two fields the renderers set
 * (cell/renderer_base.h:33-34, cell/spu/trace_spu.cpp:159-160). The body of Shade is not in the snapshot; mode 0
 * below is the head-light Lambert term written down in include/yv_format.h, stated here a third time,
 * independently of the oracle and of the CUDA kernel. Modes 1 and 2 are probes: they return the 32 bits of the
 * hit distance / of the VoxData as the "colour", so a frame rendered by the reference's own loop carries its
 * TraceResult out bit for bit. */
#ifndef YV_REF_SHIM_SHADER_H
#define YV_REF_SHIM_SHADER_H

#include "vox_node.h"                   /* Color32 */

extern "C" int yv_ref_shader_probe;     /* 0 = shade, 1 = bits of t, 2 = bits of VoxData (defined in the harness) */

struct SimpleShader {
  point_3f viewerPos, lightPos;

  static Color32 from_bits(unsigned int u) { Color32 c; memcpy(&c, &u, 4); return c; }

  Color32 Shade(VoxData data, const point_3f &dir, float t) const {
    if (yv_ref_shader_probe == 1) { unsigned int u; memcpy(&u, &t, 4); return from_bits(u); }
    if (yv_ref_shader_probe == 2) return from_bits(data);
    /* normal: octahedral, bits 16..23 / 24..31 */
    float fx = (float)((data >> 16) & 255u) / 127.5f - 1.0f;
    float fy = (float)((data >> 24) & 255u) / 127.5f - 1.0f;
    float fz = (1.0f - fabsf(fx)) - fabsf(fy);
    if (fz < 0) {
      const float ox = (1.0f - fabsf(fy)) * (fx >= 0 ? 1.0f : -1.0f);
      const float oy = (1.0f - fabsf(fx)) * (fy >= 0 ? 1.0f : -1.0f);
      fx = ox; fy = oy;
    }
    const float nl = sqrtf((fx * fx + fy * fy) + fz * fz);
    const float nx = fx / nl, ny = fy / nl, nz = fz / nl;
    /* P = viewer + dir * t ; L = normalize(light - P) */
    const float Px = viewerPos.x + dir.x * t, Py = viewerPos.y + dir.y * t, Pz = viewerPos.z + dir.z * t;
    const float vx = lightPos.x - Px, vy = lightPos.y - Py, vz = lightPos.z - Pz;
    const float len = sqrtf((vx * vx + vy * vy) + vz * vz);
    float Lx = 0.0f, Ly = 0.0f, Lz = 0.0f;
    if (len > 0) { Lx = vx / len; Ly = vy / len; Lz = vz / len; }
    const float ndl = (nx * Lx + ny * Ly) + nz * Lz;
    const float d = ndl > 0 ? ndl : 0.0f;
    const float k = 0.1f + 0.9f * (d * 1.0f);
    /* colour: RGB565 -> 8 bits, scaled, rounded */
    const unsigned int r5 = (data >> 11) & 31u, g6 = (data >> 5) & 63u, b5 = data & 31u;
    const unsigned int c8[3] = { (r5 << 3) | (r5 >> 2), (g6 << 2) | (g6 >> 4), (b5 << 3) | (b5 >> 2) };
    int out[3];
    for (int i = 0; i < 3; ++i) {
      const float v = floorf((float)c8[i] * k + 0.5f);
      out[i] = (int)(v < 255.0f ? v : 255.0f);
    }
    return Color32(out[0], out[1], out[2], 255);
  }
};

#endif
