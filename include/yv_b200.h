/* yv_b200.h — C ABI of the B200-native SVO ray caster (libyv_b200.so).
 *
 * Drop-in boundary for the reference's renderer path. Plain C types only; every entry point
 * returns 0 on success or a negative status, with a message available from yv_last_error().
 * Each declaration cites the reference interface it replaces (paths relative to the
 * znah/yoxel-voxel tree). INTEGRATION.md shows the reference-side binding.
 *
 * There is no CPU fallback: every render entry point fails with YV_ERR_CUDA when no
 * sm_100 device is usable.
 */
#ifndef YV_B200_H
#define YV_B200_H

#include <stddef.h>
#include <stdint.h>

#include "yv_format.h"

#ifdef __cplusplus
extern "C" {
#endif

#define YV_OK            0
#define YV_ERR_ARG      (-1)   /* bad argument                                              */
#define YV_ERR_IO       (-2)   /* file could not be read / written                          */
#define YV_ERR_FORMAT   (-3)   /* malformed node pool                                       */
#define YV_ERR_CUDA     (-4)   /* CUDA runtime failure or no usable device                  */
#define YV_ERR_NOSCENE  (-5)   /* RenderFrame without SetScene (reference returns NULL,
                                  cell/ppu_renderer.cpp:78-79)                              */
#define YV_ERR_NOMEM    (-6)

typedef struct yv_svo yv_svo;             /* SVOData        (cell/svodata.h:22-55)          */
typedef struct yv_renderer yv_renderer;   /* ISVORenderer   (cell/svorenderer.h:5-24)       */

const char *yv_last_error(void);          /* thread-local message of the last failure       */
int yv_abi_version(void);

/* ---- scene: SVOData (cell/svodata.h) --------------------------------------------------- */

/* SVOData::Load(const char*)  (cell/svodata.h:31-50) */
int yv_svo_load(const char *path, yv_svo **out);
/* SVOData::Load on an object renderers already hold: the reference reloads in place and SetScene keeps the pointer
 * (cell/svodata.h:31-50, cell/renderer_base.h:28). The pool inside the handle is replaced, device copies are re-made
 * at the next frame, bound renderers stay valid. On failure the old scene is kept. */
int yv_svo_load_into(yv_svo *svo, const char *path);
/* scene built elsewhere (demo/SVORenderer.h:14 SetScene(DynamicSVO*)): copies the pool */
int yv_svo_from_memory(yv_node_id root, const yv_vox_node *nodes, uint32_t count, yv_svo **out);
/* DynamicSVO::Save (ore/src/main.cpp:123) */
int yv_svo_save(const yv_svo *svo, const char *path);
/* renderers that still hold the scene fall back to "no scene" (RenderFrame -> NULL) */
void yv_svo_free(yv_svo *svo);
/* SVOData::GetRoot / operator[] (cell/svodata.h:52-54) */
yv_node_id yv_svo_root(const yv_svo *svo);
uint32_t yv_svo_node_count(const yv_svo *svo);          /* DynamicSVO::GetNodeCount            */
uint32_t yv_svo_depth(const yv_svo *svo);
const yv_vox_node *yv_svo_nodes(const yv_svo *svo);     /* host pool, reference layout         */

/* procedural scenes: gen_spheres.py:5-35, gen_largevol.py:8-40 (dataset replaced by seeded
 * noise), MakeSphereSource + BuildRange (ore/src/main.cpp:69,121), MakeRawSource (:37-52) */
int yv_svo_build_sphere_fractal(int depth, int threads, yv_svo **out);
int yv_svo_build_iso_volume(int depth, uint32_t seed, int iso_level, int threads, yv_svo **out);
int yv_svo_build_single_sphere(int depth, int cx, int cy, int cz, int radius,
                               uint8_t r, uint8_t g, uint8_t b, yv_svo **out);
int yv_svo_build_from_dense(int depth, const uint32_t *voxdata, yv_svo **out);
uint32_t yv_pack_voxdata(uint8_t r, uint8_t g, uint8_t b, float nx, float ny, float nz);

/* ---- editing: DynamicSVO + VoxelSource (ore/src/main.cpp:101-129; demo/Demo.cpp:82-114) ------------- */
typedef struct yv_source yv_source;        /* VoxelSource (ore/src/main.cpp:106-117)                         */
int yv_svo_create(yv_svo **out);                                         /* DynamicSVO() — empty scene        */
int yv_source_sphere(int radius, uint8_t r, uint8_t g, uint8_t b, int inverted, yv_source **out);   /* MakeSphereSource (:69) */
int yv_source_raw(const int size[3], const uint32_t *voxdata, yv_source **out);                      /* MakeRawSource (:37-52): VoxData words, x fastest, 0 = empty */
/* MakeRawSource in the reference's own argument form (:37-52, call site scene_gen.py:20-79): per voxel a Color32 (R,G,B,A)
 * and a Normal32 (int8 x,y,z,pad); alpha 0 = empty, 255 = surface voxel, anything else = buried (part of a FullNode) */
int yv_source_raw_colors_normals(const int size[3], const uint8_t *colors_rgba, const int8_t *normals_xyzw, yv_source **out);
int yv_source_iso(const int size[3], const uint8_t *data, int iso_level, int inside,
                  uint8_t r, uint8_t g, uint8_t b, yv_source **out);     /* MakeIsoSource + SetIsoLevel/SetInside/SetColor (:54-67,112-116) */
void yv_source_free(yv_source *src);
int yv_source_size(const yv_source *src, int size[3], int pivot[3]);     /* GetSize / GetPivot (:107-108)     */
/* DynamicSVO::BuildRange(level, pos, mode, src) (:121): merge the source, placed with its pivot at voxel `pos`
 * of the 2^level grid, into the scene. mode 0 = BUILD_MODE_GROW (union), 1 = BUILD_MODE_CLEAR (subtraction;
 * the source's surface voxels become the cavity wall — use an inverted sphere as demo/Demo.cpp:109 does). */
int yv_svo_build_range(yv_svo *svo, int level, const int pos[3], int mode, const yv_source *src);
uint32_t yv_svo_live_node_count(const yv_svo *svo);                      /* nodecount (:126), free-list excluded */
int yv_svo_node_count_by_level(const yv_svo *svo, int *counts, int capacity);   /* GetNodeCountByLevel1 (:129) */
/* page versions (256-node pages, reaction/report/main.tex:71) */
uint32_t yv_svo_version(const yv_svo *svo);
int yv_svo_count_changed_pages(const yv_svo *svo, uint32_t since_version);      /* CountChangedPages (:127)  */
/* CudaSVO::Update for a scene under edit: copy only the pages written since the last call into the device's
 * raw (reference-layout) mirror; *bytes_transferred = CountTransfrerSize (:128). Renderers with option
 * "layout" = 1 read that mirror (and call this implicitly before every frame). */
int yv_svo_update(yv_svo *svo, int device, uint64_t *bytes_transferred);

/* CudaSVO::Update (demo/SVORenderer.cpp:33-53): bring the packed 16-byte record form of the pool onto `device`:
 * the raw pool is copied page-wise and re-laid-out breadth-first by GPU kernels (environment YV_HOST_PACK=1, or a
 * pool with shared sub-trees, uses the host repack instead). Implicit at the first render if not called. */
int yv_svo_upload(yv_svo *svo, int device);
/* "Broadcast at load" for a replicated scene: make the packed pool resident on dst_device by copying it from
 * src_device over NVLink / PCIe P2P (uploading and re-packing it on src_device first if it is not there yet) instead
 * of a second upload through the host. A multi-device renderer does this by itself. */
int yv_svo_replicate(yv_svo *svo, int src_device, int dst_device);
/* bytes resident on `device` for this scene (0 if not uploaded) */
uint64_t yv_svo_device_bytes(const yv_svo *svo, int device);
/* repacked record / leaf counts, for roofline arithmetic and tests */
int yv_svo_packed_counts(yv_svo *svo, uint32_t *records, uint32_t *leaves);
/* the device's packed arrays copied back (tests: the GPU repack equals the host repack); any pointer may be NULL */
int yv_svo_device_packed_copy(yv_svo *svo, int device, uint32_t *n_records, uint32_t *n_leaves,
                              uint32_t *records_out, uint32_t *leaves_out, uint32_t *node_data_out);
/* copy of the repacked host arrays (tests): records = 4 u32 each, leaves = 1 u32 each */
int yv_svo_packed_copy(yv_svo *svo, uint32_t *records_out, uint32_t *leaves_out);
/* the octant-occupancy ("grandchild") masks the culling traversal reads, one uint64 per record — byte c = which octants
 * of child node c hold anything: from the host repack, and as they sit in the device's records (tests compare them) */
int yv_svo_octant_masks(yv_svo *svo, uint64_t *out);
int yv_svo_device_octant_masks(yv_svo *svo, int device, uint64_t *out);

/* ---- renderer: ISVORenderer (cell/svorenderer.h:5-24) + SVORenderer (demo/SVORenderer.h) -- */

/* CreateSimpleRenderer / CreateThreadedRenderer / CreateSPURenderer (cell/svorenderer.h:26-30) */
int yv_renderer_create(int device, yv_renderer **out);
/* CreateSPURenderer (cell/svorenderer.h:30; cell/spu_renderer.cpp:30-90): ONE renderer that drives every GPU whose bit is
 * set in device_mask inside each frame call, the way SPURenderer drives every SPE: GPU k of n renders the blocks b of
 * the frame with b % n == k (blockStart / blockStride, cell/spu_renderer.cpp:80-83, cell/spu/trace_spu.cpp:164) and
 * stores its pixels straight into the one frame the call returns (the SPEs' DMA into the PPU's colour buffer,
 * trace_spu.cpp:171-176). The scene is uploaded and re-packed once, on the first GPU of the mask, and copied to the
 * others over NVLink. Every other entry point takes the handle unchanged; yv_set_rows / yv_set_interleave are refused
 * (the group owns the partition: yv_set_partition), SSNA needs a single-device handle. */
#define YV_ALL_DEVICES (~(uint64_t)0)      /* every GPU of the machine: spe_cpu_info_get(SPE_COUNT_USABLE_SPES), spu_renderer.cpp:73 */
int yv_renderer_create_multi(uint64_t device_mask, yv_renderer **out);
/* the same with an explicit list of CUDA ordinals; a GPU may be listed more than once (its members then share that
 * GPU: useful to exercise the group machinery on a single-GPU machine) */
int yv_renderer_create_group(const int *devices, int count, yv_renderer **out);
int yv_renderer_device_count(const yv_renderer *r);              /* GPUs behind the handle (1 for yv_renderer_create) */
int yv_renderer_device(const yv_renderer *r, int k);             /* CUDA ordinal of member k, -1 if out of range       */
/* mode 0 (default): blocks of band_rows rows (multiple of 16, default 32) dealt round-robin over the GPUs;
 * mode 1: contiguous bands (band_rows ignored). */
int yv_set_partition(yv_renderer *r, int mode, int band_rows);
/* device time member k spent on its own share of the last frame (imbalance across the group), ms */
float yv_member_frame_ms(const yv_renderer *r, int k);
/* wall time and bytes of the last replication of the scene over peer copies (0 when none happened): the copies to all
 * peers run concurrently and are timed from the first one issued to the last one complete, allocations excluded */
int yv_replicate_stats(const yv_renderer *r, double *ms, uint64_t *bytes);
void yv_renderer_destroy(yv_renderer *r);

int yv_set_scene(yv_renderer *r, yv_svo *svo);                 /* SetScene      (:12) borrowed   */
int yv_set_view_pos(yv_renderer *r, const float pos[3]);       /* SetViewPos    (:14) also moves
                                                                  the head light (renderer_base.h:30-35) */
int yv_set_view_dir(yv_renderer *r, const float dir[3]);       /* SetViewDir    (:15)            */
int yv_set_view_up(yv_renderer *r, const float up[3]);         /* SetViewUp     (:16)            */
int yv_set_resolution(yv_renderer *r, int width, int height);  /* SetResolution (:18) / SetViewSize */
int yv_get_resolution(const yv_renderer *r, int *width, int *height);   /* GetResolution (:19)   */
int yv_set_fov(yv_renderer *r, float fov_deg);                 /* SetFOV        (:21)            */
int yv_get_fov(const yv_renderer *r, float *fov_deg);          /* GetFOV (demo/SVORenderer.h:23) */
/* SetDetailCoef / GetDetailCoef (demo/SVORenderer.h:25-26): level-of-detail cut-off of the CUDA tracer.
 * With coef > 0 a child node whose cube is smaller than coef * rad(fov/2) / width * (entry distance)
 * (rp.detailCoef, demo/SVORenderer.cpp:104) is not descended into: it is the hit, reported with child = -1
 * and shaded with its sub-tree average VoxNode::data (demo/SVORenderer.cpp:176-179). 0 (default) = off. */
int yv_set_detail_coef(yv_renderer *r, float coef);
/* SetLigth(i, LightParams) (demo/SVORenderer.h:34; demo/Demo.cpp:141-167): point lights of the CUDA renderer's
 * ShadeSimple. While any light is enabled, primary-ray frames are shaded with the Phong model written down in
 * yv_format.h (ambient 0.1, specular exponent 10, attenuation 1/(a0 + a1 d + a2 d^2)) instead of the head-light
 * Lambert of the CPU tracer. index 0..YV_MAX_LIGHTS-1. Secondary-ray frames keep the Lambert model. */
int yv_set_light(yv_renderer *r, int index, const yv_light *light);
/* SetShowNormals / GetShowNormals (demo/SVORenderer.h:31-32): write the unpacked normal as the colour */
int yv_set_show_normals(yv_renderer *r, int enable);
int yv_get_show_normals(const yv_renderer *r, int *enable);
/* SetSSNA / GetSSNA (demo/SVORenderer.h:28-29; Demo.cpp:204): screen-space normal approximation. The frame's
 * view-space z-buffer is blurred five times (SVORenderer::Render, demo/SVORenderer.cpp:126-141) and every hit pixel is
 * shaded (Lambert, Phong or show-normals as selected above) with the normal rebuilt from it instead of the voxel's
 * stored normal; arithmetic in yv_format.h "SSNA". Primary-ray frames only, whole frame on one device: a frame call
 * with a row band or an interleaved partition set returns YV_ERR_ARG. The reference constructs with SSNA on
 * (demo/SVORenderer.cpp:15); this handle starts with it off so that the default frame is ISVORenderer's.
 * yv_set_ssna_voxel_size replaces the hard-coded voxSize = 1/2048 of demo/SVORenderer.cpp:129 (0 restores it). */
int yv_set_ssna(yv_renderer *r, int enable);
int yv_get_ssna(const yv_renderer *r, int *enable);
int yv_set_ssna_voxel_size(yv_renderer *r, float voxel_size);
/* Hiding voxelisation artefacts (reaction/report/main.tex:107-114 — report text only, no code in the snapshot):
 * yv_set_jitter displaces every primary ray's origin by `amplitude` (scene units; a voxel is 2^-depth) along a
 * per-pixel hashed unit vector (arithmetic in yv_format.h); 0 switches it off. Primary-ray frames with the default
 * schedule, stack and layout only (anything else: YV_ERR_ARG at the frame call).
 * yv_render_accumulated draws `frames` frames with seeds seed, seed+1, ... and returns their per-channel integer mean
 * ((sum + frames/2) / frames), the report's "average of several consecutive frames"; *rgba as for yv_render_frame.
 * Row bands / interleave apply to the traced frames; the mean always covers the whole frame buffer. */
int yv_set_jitter(yv_renderer *r, float amplitude, uint32_t seed);
int yv_render_accumulated(yv_renderer *r, int frames, const uint8_t **rgba);
int yv_get_detail_coef(const yv_renderer *r, float *coef);

/* const Color32* RenderFrame()  (cell/svorenderer.h:23): synchronous; *rgba aliases
 * renderer-owned pinned host memory (width*height*4 bytes, R,G,B,A), valid until the next
 * yv_set_resolution / yv_render_frame* / destroy on this handle. */
int yv_render_frame(yv_renderer *r, const uint8_t **rgba);
/* void Render(void* d_dstBuf)  (demo/SVORenderer.h:36): writes uchar4 pixels of the current
 * row band into a DEVICE pointer addressed as a full frame (base + (y*width+x)*4). The pointer
 * may be a peer / IPC mapping of another GPU's frame buffer. Synchronous. */
int yv_render_frame_device(yv_renderer *r, void *d_rgba);
/* same, asynchronous on the renderer's stream (pair with yv_sync) */
int yv_render_frame_device_async(yv_renderer *r, void *d_rgba);
int yv_sync(yv_renderer *r);
/* Frames in flight, for flythrough batches: yv_render_frame_async starts a frame with the camera as set and returns
 * at once; yv_wait_frame blocks until that frame is complete where its consumer reads it. Up to "slots" (option,
 * default 2) frames may be outstanding, so the delivery of frame k overlaps the traversal of frame k+1 — what the
 * CUDA demo gets from rendering into a mapped PBO while the previous one is displayed (demo/Demo.cpp:172-181).
 * dst == NULL: the frame lands in a renderer-owned pinned host slot (returned by yv_wait_frame, valid until that slot
 * is reused); else dst is a caller-owned full-frame buffer — page-locked / registered host memory or device memory.
 * Delivery follows option "zero_copy": 1 = the kernels store into the target; 0 (and every frame with a second pass)
 * = the frame is drawn in HBM and moved by the copy engine(s), each GPU of a group moving its own rows. */
int yv_render_frame_async(yv_renderer *r, void *dst, int *ticket);
int yv_wait_frame(yv_renderer *r, int ticket, const uint8_t **rgba);
/* renderer-owned device frame buffer (full frame), for callers that have none */
int yv_device_framebuffer(yv_renderer *r, void **d_rgba);

/* Screen-space partition for multi-GPU rendering (SPURenderer splits blocks across SPEs,
 * cell/spu_renderer.cpp:73-87): render only rows [y0,y1). Default: the whole frame. */
int yv_set_rows(yv_renderer *r, int y0, int y1);
/* Interleaved partition for load balance (the SPU program's block stride, cell/spu/trace_spu.cpp:164):
 * the frame is cut into blocks of band_rows rows (multiple of 16); this renderer draws the blocks b with
 * b % stride == phase. stride 1 restores the contiguous mode. Cancels yv_set_rows and vice versa. */
int yv_set_interleave(yv_renderer *r, int band_rows, int stride, int phase);

/* Secondary rays (BASELINE config 4). shadow: 0/1; ao_samples: 0..16. light_pos is used for
 * the Lambert term and the shadow ray when shadow != 0 (otherwise the light rides on the eye).
 * A shadow ray is occluded by a hit closer than the light, an AO ray by a hit closer than ao_max_t; both are
 * traced with that range limit (the traversal meets cells front to back, so it stops at the first cell entered
 * beyond the limit — same outcome as tracing to the end, far fewer node visits). */
int yv_set_secondary(yv_renderer *r, int shadow, int ao_samples, uint32_t seed,
                     const float light_pos[3], float voxel_size, float ao_max_t);

/* TraceResult per pixel (cell/ppu_renderer.cpp:7-12). Hit buffers cost 12 B/ray of extra
 * stores, so they are off unless enabled. yv_get_hits copies the last frame's records to host
 * arrays of width*height entries (any pointer may be NULL). */
int yv_enable_hits(yv_renderer *r, int enable);
int yv_get_hits(yv_renderer *r, uint32_t *node, int32_t *child, float *t);
/* SVORenderer::DumpTraceData(fnbase) (demo/SVORenderer.cpp:158-192): writes <fnbase>_<W>x<H>.dist (f32 hit
 * distance), .color (RGBA8 unpacked voxel colour) and .normal (3 x f32 unpacked normal) for the last frame.
 * Needs yv_enable_hits. (The reference's .dist is zero-filled because its assignment is commented out.) */
int yv_dump_trace_data(yv_renderer *r, const char *fnbase);
/* optional per-ray counters (node fetches incl. re-fetches after a pop) for profiling */
int yv_enable_counters(yv_renderer *r, int enable);
int yv_get_counters(yv_renderer *r, uint32_t *fetches_per_ray);

/* device time of the last frame's kernels (CUDA events on the renderer's stream), ms */
float yv_last_frame_ms(const yv_renderer *r);
/* kernels launched by the last frame */
int yv_last_frame_launches(const yv_renderer *r);

/* Run the renderer on a caller-owned CUDA stream (cudaStream_t as void*; NULL = own stream) */
int yv_set_stream(yv_renderer *r, void *cuda_stream);

/* Kernel variant knobs (for ablation runs; defaults are the tuned ones):
 *   "smem_nodes"  number of top-of-tree records staged in shared memory per CTA
 *   "schedule"    0 = one CTA per 16x8 tile, one pixel per lane; 1 = persistent CTAs pulling tiles
 *                 from an atomic counter; 2 = per-warp ray queue (a warp schedules the 128 rays of a
 *                 16x8 tile over its lanes; primary rays). "persistent" is an alias.
 *                 (with warp-level lane refill: ballot + popc compaction of finished rays)
 *   "refill"      persistent schedule: refill a warp once <= this many lanes are still traversing
 *   "sec_queue"   secondary rays: 0 (default) = every lane traces its own pixel's rays as stages
 *                 (render_frame<SEC>); 1 = primary and shadow rays in lock-step, AO rays pooled per warp and pulled
 *                 by whichever lane is free (render_sec_queue; fills more lanes but loses lock-step fetches: slower)
 *   "sec_threshold" secondary rays (stage machine): lanes whose ray has ended are handed their pixel's next ray once <= this many
 *                 lanes of the warp are still traversing (-1 = only when the whole warp has drained)
 *   "zero_copy"   1 (default) = yv_render_frame's kernel stores its pixels straight into the pinned host frame
 *                 (device-addressable under UVA): the posted PCIe writes overlap the traversal, there is no copy and
 *                 no second launch. With a second pass over the image (Phong / show-normals / SSNA) the trace kernel
 *                 draws in HBM and the pass that finishes the pixels stores them into the host frame. 0 = the copy
 *                 paths below
 *   "pipeline"    with zero_copy 0: number of row chunks (2..8, default 4) yv_render_frame cuts the frame into:
 *                 chunks render on two alternating streams and each chunk's device->host copy overlaps the next
 *                 chunk's kernel; 0 or 1 = one launch, then one copy
 *   "pipeline_taper" each chunk is this many percent (10..100, default 100) of the rows of the one before it
 *   "layout"      0 = packed 16-byte records (default; re-packed and re-uploaded in full after an edit),
 *                 1 = the raw reference pool mirrored page by page (yv_svo_update) — for scenes under edit
 *   "stack"       where the traversal stack lives: 0 local memory, 4 = four-entry
 *                 shared-memory ring spilling to local memory
 *   "slots"       frames in flight for yv_render_frame_async (2..4, default 2)
 *   "group_threads" multi-device handles: 1 (default) = every peer GPU's share of a frame is issued by its own persistent
 *                 host thread (SPURenderer's one thread per SPE, cell/spu_renderer.cpp:76-87), so the N launches start
 *                 together; 0 = one loop on the calling thread (the launches of 8 GPUs then start ~18 us apart). The
 *                 threads poll for the next frame for YV_WORKER_SPIN_US microseconds (environment, default 2000, 0 = off)
 *                 before they sleep: a condition-variable wake-up is 10-30 us of launch skew
 *   "ssna_fused"  SSNA's BlurZ x5 + ShadeSimple (demo/SVORenderer.cpp:126-147) as ONE persistent cooperative launch that
 *                 pulls 32x32 tiles from a counter per pass, grid barriers between the passes: 1 = with the next tile
 *                 prefetched into a second shared-memory buffer by a 2-D TMA load (needs width % 4 == 0, else as 2),
 *                 2 = with plain staging loads; 0 (default) = six launches. Same pixels in all three; measured per SSNA
 *                 frame on config 2: 0.841 (six launches) / 0.846 (TMA) / 0.861 ms (plain) — the grid barriers and the
 *                 per-tile hand-shake cost what the launch boundaries did, so the fused forms stay options.
 *                 Packed layout, local stack, no staging. */
int yv_set_option(yv_renderer *r, const char *name, int value);
int yv_get_option(const yv_renderer *r, const char *name, int *value);

/* DynamicSVO::TraceRay (ore/src/main.cpp:125): trace `count` arbitrary rays on the device.
 * pos/dir are count*3 floats; outputs are count entries (any may be NULL). */
int yv_trace_rays(yv_renderer *r, const float *pos, const float *dir, uint32_t count,
                  uint32_t *node, int32_t *child, float *t);

/* ---- multi-process frame gather over NVLink (one process per GPU) ------------------------ */
/* Export a device allocation made by this library (yv_device_framebuffer) as a 64-byte CUDA
 * IPC handle; open it in another process; the mapping is a valid target for
 * yv_render_frame_device there (direct peer stores into GPU 0's frame). */
int yv_device_alloc(int device, size_t bytes, void **d_ptr);      /* plain cudaMalloc (IPC-exportable) */
int yv_device_free(int device, void *d_ptr);
int yv_copy_to_host(int device, void *dst_host, const void *src_device, size_t bytes);
int yv_ipc_export(void *d_ptr, uint8_t handle[64]);
int yv_ipc_open(int device, const uint8_t handle[64], void **d_ptr);
int yv_ipc_close(void *d_ptr);

/* ---- a caller-owned host frame as render target -------------------------------------------- */
/* RenderFrame's consumer reads host memory (const Color32*, cell/svorenderer.h:23; cell/main.cpp:36). With several
 * processes (one per GPU) rendering parts of one frame, or frames of one batch, the cheapest way to the host is for
 * every GPU to store its pixels straight into ONE host frame over its own PCIe link: map a shared-memory segment in
 * every process, register it here, and pass the returned address to yv_render_frame_device. yv_host_register
 * page-locks [host_ptr, host_ptr + bytes) for `device` and returns the address the kernel stores through. */
int yv_host_register(int device, void *host_ptr, size_t bytes, void **device_ptr);
int yv_host_unregister(void *host_ptr);

/* RendererBase::InitRayDir (cell/renderer_base.h:50-61): the per-frame ray basis the kernel
 * consumes, computed on the host (pure function; exposed for tests and external ray generators) */
int yv_init_ray_dir(const float dir[3], const float up[3], float fov_deg, int width, int height,
                    float dir0[3], float du[3], float dv[3]);

int yv_device_count(void);
int yv_device_name(int device, char *buf, size_t len);

#ifdef __cplusplus
}
#endif
#endif /* YV_B200_H */
