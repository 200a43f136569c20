/* yv_oracle.h — CPU oracle for the SVO ray-caster path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's CPU tracer. Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it;
 * the product (yoxel-voxel_b200/) never links, imports or calls anything in oracle/.
 *
 * How the oracle is pinned. The reference snapshot ships no golden vectors, no scene files and no test for this path,
 * and as shipped its tracer does not compile (the cpp/ directory with stdafx.h, trace_utils.h, vox_node.h, shader.h,
 * rdd.h, utils.h is absent, as are Boost.Thread and the Cell SDK — SURVEY.md §0.2, §8c). oracle/Makefile (target `ref`)
 * compiles the reference's sources anyway — unmodified, from where they lie — behind stand-ins for exactly those
 * headers (oracle/ref_shim/, each stating what it replaces and on what evidence), into oracle/_ref:
 *   - cell/ppu_renderer.cpp (+ renderer_base.h, svorenderer.h, svodata.h, alignedarray.h, nest/include/geometry):
 *     SVOData::Load, the setters and defaults, InitRayDir, the per-pixel loop, RecTrace, Simple/TreadedRenderer;
 *   - cell/spu/trace_spu.cpp: the SPU program with the reference's own FindFirstChildSPU / GoNextSPU / RecTrace /
 *     FetchNode cache / RenderBlock;
 *   - cell/main.cpp + cell/spu_renderer.cpp + cell/spu/trace_spu.cpp: the complete Cell application as one executable
 *     (oracle/_ref/cell_main_spu), its SPEs played by the host CPU; the frame it writes equals this oracle's;
 *   - cell/spu/trace_spu.c_: the first scalar (double) tracer.
 *   tests/test_reference_renderer.py: on test scenes, seeded random pools and random cameras the oracle's RGBA frame,
 *   the bits of every hit distance, the VoxData, the hit ids (through leaf words that name node and child), the
 *   number of node fetches and InitRayDir's nine floats are identical to what those builds produce;
 *   tests/golden/reference_golden.npz holds their output for the boxes where /root/reference does not exist
 *   (tests/test_reference_golden.py, CPU and GPU); tests/test_reference_prototype.py covers the scalar prototype.
 *   - hand-computed rays, a brute-force voxel-grid marcher (a different algorithm), structural properties —
 *     tests/test_oracle.py.
 * STILL UNPINNED — no reference code exists for it, so the stand-ins carry OUR statement of it and agreement there
 * proves nothing: the AdjustDir epsilon, the body of SetupTrace (documented in voxel.tex:305-326 and mirrored by
 * trace_spu.c_:79-93), the VoxData bit layout, the Shade formula, and everything the CUDA renderer added (LOD, SSNA,
 * Phong lights) plus the secondary rays; those are builder decisions written down in include/yv_format.h.
 */
#ifndef YV_ORACLE_H
#define YV_ORACLE_H

#include "../include/yv_format.h"

#ifdef __cplusplus
extern "C" {
#endif

/* camera state of RendererBase (cell/renderer_base.h:7-47) */
typedef struct yvo_camera {
  float pos[3];
  float dir[3];
  float up[3];
  float fov_deg;      /* horizontal, degrees; default 70 (renderer_base.h:25) */
  int32_t width;
  int32_t height;
  float detail_coef;  /* SVORenderer::SetDetailCoef (demo/SVORenderer.h:25); 0 = no LOD cut-off   */
  int32_t show_normals;             /* SetShowNormals (demo/SVORenderer.h:31)                     */
  yv_light lights[YV_MAX_LIGHTS];   /* SetLigth (demo/SVORenderer.h:34): any enabled light -> Phong */
  int32_t ssna;                     /* SetSSNA (demo/SVORenderer.h:28): normals from the blurred z-buffer */
  float ssna_voxel_size;            /* voxSize of demo/SVORenderer.cpp:129; 0 = the reference's 1/2048    */
  float jitter_amp;                 /* displaced ray origins (reaction/report/main.tex:107-114); 0 = off  */
  uint32_t jitter_seed;
} yvo_camera;

/* RayDirData{dir0,du,dv} (cell/renderer_base.h:50-61) */
typedef struct yvo_raydir {
  float dir0[3];
  float du[3];
  float dv[3];
} yvo_raydir;

/* secondary-ray options (BASELINE config 4); all zero = primary + Lambert only */
typedef struct yvo_secondary {
  int32_t shadow;     /* 1: one shadow ray from the hit point towards light_pos            */
  int32_t ao_samples; /* 0..16 cosine-weighted hemisphere rays                             */
  uint32_t seed;      /* per-pixel hash seed                                               */
  float light_pos[3]; /* light used for shadow + Lambert when `shadow` is set              */
  float voxel_size;   /* origin offset along the normal (one voxel = 2^-depth)             */
  float ao_max_t;     /* AO rays count as occluded only if they hit within this distance   */
} yvo_secondary;

typedef struct yvo_stats {
  uint64_t rays;         /* rays traced (primary + secondary)                              */
  uint64_t node_visits;  /* executions of the node fetch at ppu_renderer.cpp:23            */
  uint64_t iterations;   /* child-loop iterations (leaf test at ppu_renderer.cpp:27)       */
  uint64_t hits;         /* primary rays that hit a leaf                                   */
} yvo_stats;

void yvo_init_ray_dir(const yvo_camera *cam, yvo_raydir *out);

/* Render rows [y0,y1) of the frame with `threads` horizontal strips
 * (SimpleRenderer when threads==1, TreadedRenderer's strip split otherwise;
 * the last strip takes the remainder rows so that every row in [y0,y1) is rendered).
 * Output arrays are full-frame sized (width*height); any may be NULL.                      */
int yvo_render(const yv_vox_node *nodes, uint32_t node_count, yv_node_id root,
               const yvo_camera *cam, const yvo_secondary *sec,
               int32_t y0, int32_t y1, int32_t threads,
               uint32_t *hit_node, int32_t *hit_child, float *hit_t,
               uint8_t *rgba, uint32_t *visits_per_ray, yvo_stats *stats);

/* Reference quirk mode: TreadedRenderer exactly as written (4 strips of H/4 rows, rows
 * 4*(H/4)..H-1 left untouched; ppu_renderer.cpp:129-142). rgba must be pre-filled by caller. */
int yvo_render_threaded_ref(const yv_vox_node *nodes, uint32_t node_count, yv_node_id root,
                            const yvo_camera *cam, uint8_t *rgba);

/* Trace one arbitrary ray (DynamicSVO::TraceRay-like; ore/src/main.cpp:125).
 * Returns 1 on hit. */
/* Model of the SPU program's software node cache (cell/spu/trace_spu.cpp:15-35,155-156,162-168): fetches and misses
 * of one run over the whole frame in the program's own block and pixel order. */
#define YVO_SPU_CACHE_SIZE 2048u
int yvo_spu_cache_model(const yv_vox_node *nodes, uint32_t node_count, yv_node_id root, const yvo_camera *cam,
                        uint64_t *fetches, uint64_t *misses);

int yvo_trace_ray(const yv_vox_node *nodes, uint32_t node_count, yv_node_id root,
                  const float pos[3], const float dir[3],
                  uint32_t *hit_node, int32_t *hit_child, float *hit_t);

/* Test-only: argmin tie order of GoNext. 0 (default) = cell/spu/trace_spu.cpp:75-78, the path's; 1 = argMin of the
 * reference's scalar prototype (cell/spu/vector.h:45-59), used only to compare against that prototype compiled into
 * oracle/_ref (tests/test_reference_prototype.py). Process-global; restore 0 after use. */
void yvo_set_tie_order(int order);

/* SimpleShader::Shade restatement, exposed for unit tests */
void yvo_shade(yv_vox_data data, const float dir[3], float t,
               const float viewer[3], const float light[3], float visibility, uint8_t out_rgba[4]);
void yvo_unpack_normal(yv_vox_data data, float n[3]);

/* SSNA building blocks (spec: include/yv_format.h "SSNA"), exposed for unit tests.
 * yvo_blur_taps: the K*K normalised Gaussian taps (demo/SVORenderer.cpp:55-79), row-major.
 * yvo_blur_z: the five BlurZ passes (:126-141) from z0 into out (both width*height floats).
 * yvo_ssna_normal: world-space normal at pixel (x,y) of a blurred z-buffer; returns 0 if undefined. */
void yvo_blur_taps(float *taps);
int yvo_blur_z(const yvo_camera *cam, const float *z0, float *out);
int yvo_ssna_normal(const yvo_camera *cam, const float *z, int32_t x, int32_t y, float n[3]);

/* SVOData::Load (cell/svodata.h:31-50). Caller frees *nodes with yvo_free. */
int yvo_load_vox(const char *path, yv_node_id *root, uint32_t *count, yv_vox_node **nodes);
void yvo_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
