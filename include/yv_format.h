/* yv_format.h — data formats and numeric conventions of the SVO ray-caster path.
 *
 * Definitions only (struct layouts, bit positions, constants, and the prose spec of the
 * arithmetic the reference snapshot does not pin). Both the CUDA product
 * (yoxel-voxel_b200/csrc) and the CPU oracle (oracle/) include this file; each implements
 * the arithmetic described here on its own, so that a parity test compares two
 * independent implementations of one written spec.
 *
 * Reference pointers (relative to /root/reference):
 *   node record ............ reaction/report/main.tex:38-55  (cpp/vox_node.h is absent)
 *   .vox file .............. cell/svodata.h:31-50
 *   hit record ............. cell/ppu_renderer.cpp:7-12
 *   frame buffer ........... cell/renderer_base.h:22,42 ; cell/main.cpp:36 ("RGBA", CharPixel)
 *   near-zero clamp ........ reaction/report/voxel.tex:316-318
 */
#ifndef YV_FORMAT_H
#define YV_FORMAT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- node pool (main.tex:38-55) ------------------------------------------------------ */

typedef uint32_t yv_node_id;              /* VoxNodeId                                      */
typedef uint32_t yv_vox_data;             /* VoxData: 16-bit colour + 16-bit normal         */

#define YV_NULL_BIT    0x80000000u        /* IsNull(id): top bit set (main.tex:62)          */
#define YV_EMPTY_NODE  0x80000000u        /* EmptyNode  (main.tex:54)                       */
#define YV_FULL_NODE   0x80000001u        /* FullNode   (main.tex:55)                       */

/* VoxNode, #pragma pack(4): 40 bytes, array-of-structs, little-endian on disk.            */
typedef struct yv_vox_node {
  uint32_t flags;                         /* bits 0..7 leaf flags, 8..15 null flags,
                                             bit 19 empty flag (main.tex:40-42)             */
  yv_vox_data data;                       /* sub-tree average (coarser LOD)                 */
  uint32_t child[8];                      /* node id | inline VoxData (leaf bit set) |
                                             id with top bit set = empty/full               */
} yv_vox_node;

#define YV_NODE_BYTES 40u
#define YV_LEAF_FLAG(flags, i)  (((flags) >> (i)) & 1u)          /* GetLeafFlag              */
#define YV_NULL_FLAG(flags, i)  (((flags) >> (8 + (i))) & 1u)
#define YV_IS_NULL(id)          (((id) & YV_NULL_BIT) != 0u)     /* IsNull                   */

/* Child index bit i (0=x,1=y,2=z) set = upper half of the node along that axis
 * (trace_spu.cpp:62-64 builds childId that way and XORs it with dirFlags).                 */

/* ---- .vox container (svodata.h:31-50) ------------------------------------------------- */
/* header = 4 little-endian u32: root id, two words the loader discards, node count;
 * then count * 40 bytes of raw nodes. We write the two ignored words as
 * YV_VOX_MAGIC and the tree depth (levels below the root cube); readers ignore them.       */
#define YV_VOX_HEADER_BYTES 16u
#define YV_VOX_MAGIC        0x31584f56u   /* "VOX1" */

/* ---- hit record (ppu_renderer.cpp:7-12) ----------------------------------------------- */
/* Stored as three planar per-pixel arrays: node (u32), child (i32), t (f32).
 * A ray that hits nothing has node = YV_EMPTY_NODE, child = -1, t = 0.                     */
#define YV_MISS_NODE   YV_EMPTY_NODE
#define YV_MISS_CHILD  (-1)

/* ---- frame buffer (Color32) ------------------------------------------------------------ */
/* 4 bytes per pixel in memory order R,G,B,A; row-major; row 0 is the top row.
 * Miss = (0,0,0,0) (ppu_renderer.cpp:54). Hit alpha = 255.                                  */

/* ---- builder decisions: things the snapshot does not pin ------------------ */

/* AdjustDir: |d_i| < eps  =>  d_i = copysign(eps, d_i)   (voxel.tex:316-318)               */
#define YV_DIR_EPS 1e-6f

/* VoxData bit packing:
 *   bits  0..15  colour, RGB565: r5 = bits 11..15, g6 = bits 5..10, b5 = bits 0..4
 *   bits 16..23  normal octahedral u (0..255)
 *   bits 24..31  normal octahedral v (0..255)
 * Colour decode to 8 bits (integer): r8=(r5<<3)|(r5>>2), g8=(g6<<2)|(g6>>4), b8=(b5<<3)|(b5>>2).
 * Normal decode (float32, round-to-nearest, no FMA, in this order):
 *   fx = (float)u / 127.5f - 1.0f;  fy = (float)v / 127.5f - 1.0f;
 *   fz = (1.0f - |fx|) - |fy|;
 *   if (fz < 0) { ox = (1.0f - |fy|) * sgn(fx); oy = (1.0f - |fx|) * sgn(fy); fx = ox; fy = oy; }
 *       with sgn(a) = (a >= 0) ? 1.0f : -1.0f
 *   len = sqrt((fx*fx + fy*fy) + fz*fz);  n = (fx/len, fy/len, fz/len)
 * Normal encode (builder side, any precision): p = n.xy / (|nx|+|ny|+|nz|); fold if nz < 0;
 *   u = round((p.x*0.5+0.5)*255), v likewise, clamped to 0..255.                             */
#define YV_PACK_RGB565(r8, g8, b8) \
  ((uint32_t)((((r8) >> 3) << 11) | (((g8) >> 2) << 5) | ((b8) >> 3)))

/* SimpleShader::Shade(VoxData, dir, t) (ppu_renderer.cpp:67; body absent) is restated as a
 * head-light Lambert term (north_star: "Lambert"), float32, no FMA, in this order:
 *   P   = viewer + dir * t                       (component-wise: mul, then add)
 *   Lv  = light - P
 *   len = sqrt((Lv.x*Lv.x + Lv.y*Lv.y) + Lv.z*Lv.z)
 *   L   = len > 0 ? Lv / len : (0,0,0)
 *   ndl = (n.x*L.x + n.y*L.y) + n.z*L.z ;  d = ndl > 0 ? ndl : 0
 *   k   = YV_SHADE_AMBIENT + YV_SHADE_DIFFUSE * d      (mul, then add)
 *   c8' = (uint8) min(255.0f, floorf(c8 * k + 0.5f))    per channel (mul, add, floor)
 *   k is multiplied by the visibility terms of the secondary rays when they are enabled
 *   (see yv_b200.h, YV_SECONDARY_*).
 * ambient 0.1 follows the CUDA variant's rp.ambient (demo/SVORenderer.cpp:113).             */
#define YV_SHADE_AMBIENT 0.1f
#define YV_SHADE_DIFFUSE 0.9f

/* ShadeSimple with point lights (the CUDA renderer's shader: rp.ambient = 0.1, rp.specularExp = 10,
 * demo/SVORenderer.cpp:112-113; LightParams{enabled,pos,diffuse,specular,attenuationCoefs}, demo/Demo.cpp:141-147;
 * kernel body absent). Restated as Phong, float32, no FMA, in this order, per enabled light i (index order), with
 * c = the voxel colour channel as a float in 0..255, n = unpacked normal, P = viewer + dir*t:
 *   V   = (viewer - P) / |viewer - P|                     (zero vector if the length is 0)
 *   Lv  = pos_i - P ; d = sqrt((Lv.x*Lv.x + Lv.y*Lv.y) + Lv.z*Lv.z) ; L = Lv / d   (skip the light if d == 0)
 *   att = 1 / ((a0 + a1*d) + (a2*d)*d)
 *   nl  = (n.x*L.x + n.y*L.y) + n.z*L.z ; ndl = nl > 0 ? nl : 0
 *   R   = (2*nl)*n - L ; rv = (R.x*V.x + R.y*V.y) + R.z*V.z ; rv = (nl > 0 && rv > 0) ? rv : 0
 *   s2 = rv*rv ; s4 = s2*s2 ; s8 = s4*s4 ; spec = s8*s2            (rv^10)
 *   acc_ch += att * ((diffuse_i.ch * ndl) * c_ch + (specular_i.ch * spec) * 255)
 * starting from acc_ch = YV_SHADE_AMBIENT * c_ch; out_ch = (uint8) min(255, floorf(acc_ch + 0.5f)); alpha 255.
 * SetShowNormals (demo/SVORenderer.h:31): out_ch = (uint8) floorf((n_ch * 0.5f + 0.5f) * 255.0f + 0.5f).      */
/* SSNA — screen-space normal approximation (SetSSNA / GetSSNA, demo/SVORenderer.h:28-29). The host sequence is
 * demo/SVORenderer.cpp:55-79 (Gaussian taps) and :126-147 (Trace, BlurZ x5 on ping-pong z-buffers, ShadeSimple);
 * the BlurZ / ShadeSimple kernel bodies and BlurZKernSize are absent. Restated (float32, no FMA, in this order)
 * from that sequence and from the prototype's normal reconstruction (demo/dumps/ztools.py:23-44):
 *   z-buffer  z0[p] = hit ? t * ((d.x*f.x + d.y*f.y) + d.z*f.z) : 0     d = the pixel's adjusted unit ray,
 *             f = normalized(viewDir). A pixel is "valid" iff its z != 0.
 *   taps      K = YV_BLURZ_KERN, h = (float)(K/2), scale = 2; rows y then columns x:
 *             tx = scale*((float)x - h)/h ; ty likewise ; tx = tx*tx ; ty = ty*ty ;
 *             v[y][x] = (float)exp(-(double)(tx + ty)) ; sum += v[y][x] ; afterwards w[y][x] = v[y][x] / sum.
 *   pass i    (i = 0..4) blurSize = 3 + 3*i ; pixelAng = (fov * (float)(pi/180)) / W ;
 *             zlimit = (5*voxSize) / (pixelAng * blurSize) ; voxSize = 1/2048 in the reference (:129), settable here.
 *             dst[p] = 0 if src[p] is invalid, else acc / wacc, where, over the taps q = p + (kx - K/2, ky - K/2)
 *             in row-major (ky, kx) order that lie inside the frame, are valid and have |src[q] - src[p]| < zlimit:
 *             acc = acc + w*src[q] ; wacc = wacc + w.
 *   normal    z = blurred value at p ; d2 = 2*da (da from InitRayDir) ;
 *             fx = z[x+1,y] - z ; bx = z - z[x-1,y], each defined only if that neighbour is inside the frame and valid;
 *             dx = both defined ? (|fx| < |bx| ? fx : bx) : the defined one, else 0 ; dy likewise with rows y+1 / y-1 ;
 *             nvx = (d2*dx)*z ; nvy = (d2*dy)*z ; nvz = -((d2*d2)*(z*z)) ; len = sqrt((nvx*nvx + nvy*nvy) + nvz*nvz) ;
 *             len > 0: n_c = ((right_c*nvx + down_c*nvy) + fwd_c*nvz) / len with right = normalized(fwd x up),
 *             down = -(right x fwd) ; otherwise (or z invalid) n = the voxel's stored normal.
 *   shading   the Lambert / Phong / show-normals formulas above with this n ; P still uses the unblurred t.
 * SSNA needs the whole frame on one device (the taps reach 15 rows beyond any band).                            */
#define YV_BLURZ_KERN 7
#define YV_BLURZ_PASSES 5
#define YV_SSNA_VOXEL_SIZE (1.0f / 2048.0f)

/* Hiding voxelisation artefacts (reaction/report/main.tex:107-114; described in the report only, no code in the
 * snapshot): (1) every traced ray starts from a randomly displaced origin, the displacement comparable to a voxel;
 * (2) while the view is unchanged, consecutive frames drawn with different displacements are averaged. Restated:
 *   origin    key = hash(pixel) ^ hash(seed ^ YV_JITTER_SALT) with pixel = y*W + x and hash = the lowbias32 mix of
 *             the AO rays; U = the lattice-rejection unit vector of `key` (same generator as the AO rays);
 *             O_c = pos_c + amplitude * U_c (mul, then add). The ray direction is the pixel's usual direction; the
 *             shaded point is P = O + d*t; viewer and head light stay at pos.
 *   average   frame k of n is drawn with seed + k; per channel out = (sum_k c_k + n/2) / n in integers (alpha too, so
 *             a pixel hit in some frames only gets fractional coverage).                                           */
#define YV_JITTER_SALT 0x6a09e667u
#define YV_MAX_LIGHTS 4
#define YV_SPECULAR_EXP 10
typedef struct yv_light {            /* LightParams (demo/Demo.cpp:141-147) */
  int32_t enabled;
  float pos[3];
  float diffuse[3];
  float specular[3];
  float attenuation[3];              /* constant, linear, quadratic */
} yv_light;

#ifdef __cplusplus
}
#endif
#endif /* YV_FORMAT_H */
