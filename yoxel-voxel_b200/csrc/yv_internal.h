// yv_internal.h — host-side state behind the opaque handles of include/yv_b200.h, shared by the translation
// units of libyv_b200.so (yv_api.cu: scenes, single-device frames; yv_multi.cu: device groups, frame slots).
//
//   yv_svo ........ SVOData (cell/svodata.h:22-55) + the editing state of DynamicSVO + the per-device copies
//                   (CudaSVO, demo/SVORenderer.cpp:33-53)
//   yv_renderer ... RendererBase (cell/renderer_base.h:7-61) + the CUDA renderer's extras (demo/SVORenderer.h:8-64);
//                   a handle made by yv_renderer_create_multi additionally leads one peer renderer per further GPU,
//                   the way SPURenderer leads one thread per SPE (cell/spu_renderer.cpp:30-90)
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/yv_b200.h"
#include "dynamic_svo.h"
#include "svo_host.h"
#include "svo_pack.h"

namespace yvi {

int fail(int code, const std::string &msg);      // sets the thread-local message of yv_last_error()

#define YV_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return yvi::fail(YV_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
  } while (0)

// The C ABI never lets a C++ exception reach the caller (a ctypes / cgo / JNI frame cannot unwind it):
// allocation failures become YV_ERR_NOMEM, anything else YV_ERR_ARG with the exception text.
template <class F>
int guarded(F &&body) {
  try { return body(); }
  catch (const std::bad_alloc &) { return fail(YV_ERR_NOMEM, "out of host memory"); }
  catch (const std::exception &e) { return fail(YV_ERR_ARG, std::string("internal error: ") + e.what()); }
  catch (...) { return fail(YV_ERR_ARG, "internal error"); }
}

struct DeviceSVO {
  uint4 *recs = nullptr;              // { child_base, masks, leaf_base, orig_id } per record (svo_pack.h, device form)
  uint2 *octs = nullptr;              // { octants lo, octants hi } per record: read by the culling traversal only
  uint32_t *leaves = nullptr;
  uint32_t *node_data = nullptr;      // uploaded on first use of the LOD cut-off
  size_t n_recs = 0, n_leaves = 0;
  bool root_null = true;
  int levels = 0;
  uint32_t packed_version = 0;        // scene edit version the packed copy was made from
  // raw mirror of the reference-layout pool, kept in step page by page (CudaSVO::Update)
  yv_vox_node *raw = nullptr;
  size_t raw_capacity = 0;            // nodes
  uint32_t raw_version = 0;           // every page with a version <= this is on the device
  uint32_t raw_checked_version = 0;   // scene version whose raw pool passed the depth / cycle check
};

}  // namespace yvi

struct yv_svo {
  yv::HostSVO host;
  yv::DynamicSVO dyn{ host };         // editing state (free list, page versions) over `host`
  yv::PackedSVO packed;
  bool packed_ok = false;
  uint32_t packed_version = 0;
  std::map<int, yvi::DeviceSVO> dev;
  std::vector<yv_renderer *> bound;   // renderers whose scene this is: un-set when the handle is freed
  std::mutex mu;
  uint32_t version() const { return dyn.version(); }
};

struct yv_frame_slot {                // one frame in flight (yv_render_frame_async / yv_wait_frame)
  uint8_t *h_fb = nullptr;            // pinned, portable: the frame as the host consumer sees it
  uint32_t *d_fb = nullptr;           // staged delivery: the frame in this device's HBM before the copy engine moves it
  void *target = nullptr;             // where this slot's frame is delivered (h_fb or a caller-owned buffer)
  cudaEvent_t ev_begin = nullptr, ev_done = nullptr;
  long ticket = -1;                   // -1 = free
  int launches = 0;
};

struct yv_renderer {
  int device = 0;
  int sm_count = 0;
  yv_svo *svo = nullptr;
  // RendererBase state (renderer_base.h:10-18,25)
  float pos[3] = { 0, 0, 0 }, dir[3] = { 1, 0, 0 }, up[3] = { 0, 0, 1 };
  float fov = 70.0f;
  yv_light lights[YV_MAX_LIGHTS] = {};   // SetLigth (demo/SVORenderer.h:34); any enabled light switches to Phong
  bool show_normals = false;          // SetShowNormals (demo/SVORenderer.h:31)
  bool ssna = false;                  // SetSSNA (demo/SVORenderer.h:28); the reference defaults to true, off here so that
                                      // the default frame is the CPU tracer's (ISVORenderer) image
  float ssna_voxel_size = YV_SSNA_VOXEL_SIZE;   // voxSize of demo/SVORenderer.cpp:129
  float blur_taps[YV_BLURZ_KERN * YV_BLURZ_KERN] = {};
  float jitter_amp = 0.0f;            // displaced ray origins (reaction/report/main.tex:107-114); 0 = off
  uint32_t jitter_seed = 1;
  uint4 *d_accum = nullptr;           // per-channel sums of yv_render_accumulated
  float detail_coef = 0.0f;           // SVORenderer::m_detailCoef (demo/SVORenderer.h:56); 0 = off
  int width = 0, height = 0;
  int y0 = 0, y1 = 0;
  bool rows_set = false;
  int il_rows = 0, il_stride = 1, il_phase = 0;   // interleaved partition (il_stride > 1)
  // secondary rays
  int shadow = 0, ao_samples = 0;
  uint32_t seed = 1;
  float light[3] = { 0, 0, 0 }, voxel_size = 0.0f, ao_max_t = 0.0f;
  // buffers
  uint32_t *d_fb = nullptr;
  uint8_t *h_fb = nullptr;            // pinned
  size_t fb_pixels = 0;
  uint32_t *d_hit_node = nullptr; int32_t *d_hit_child = nullptr; float *d_hit_t = nullptr;
  uint32_t *d_counters = nullptr;
  uint2 *d_shade_rec = nullptr;       // (VoxData, t) per pixel for the ShadeSimple pass
  float *d_zbuf[2] = { nullptr, nullptr };   // m_zbuf[2] (demo/SVORenderer.cpp:85-86): BlurZ ping-pong
  unsigned int *d_ssna_counters = nullptr;   // ssna_post: one tile counter per BlurZ pass (zero between frames)
  int ssna_post_grid = 0, ssna_post_grid_tma = 0;   // resident CTAs of ssna_post<false> / <true> on this device
  int opt_ssna_fused = 0;                    // 1 = BlurZ x5 + ShadeSimple as one persistent cooperative launch, tiles prefetched by TMA;
                                             // 2 = the same with plain staging loads (measured 1.5 % slower
                                             // per frame than the six launches it replaces: profiles/README.md); 0 = six launches
  unsigned int *d_tile_counter = nullptr;
  bool hits = false, counters = false;
  // launch
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // RenderFrame pipelining: row chunks rendered on two alternating streams, each chunk's D2H copy overlapped
  static constexpr int kChunks = 8;
  cudaStream_t aux[2] = { nullptr, nullptr }, copy_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_chunk[kChunks] = {}, ev_copy = nullptr;
  bool suppress_events = false;
  int opt_pipeline = 4;               // row chunks per RenderFrame (0/1 = no pipelining); 4 measured best at 1080p
  int opt_zero_copy = 1;              // 1 = RenderFrame's kernel stores its pixels straight into the pinned host frame
                                      // (posted PCIe writes overlap the traversal; no copy, one launch)
  int opt_pipeline_taper = 100;       // each chunk is this many percent of the one before it (100 = equal chunks): the copy
                                      // of the last chunk is the only one that is not hidden behind a kernel
  bool timed = false;
  int launches = 0;
  int last_launches = 1;              // kernels launched by the most recent launch_frame call
  float last_ms = -1.0f;              // device time of the last frame when it was measured at yv_wait_frame (else events)
  // options
  int opt_smem_nodes = 0;             // records staged in shared memory (585 = four levels)
  int opt_persistent = 0;
  int opt_refill = 20;                // persistent schedule: refill when <= this many lanes are live
  int opt_sec_threshold = -1;         // secondary rays: serve waiting lanes when <= this many lanes are traversing (-1 = only when the warp has drained: best once the rays are range-limited)
  int opt_sec_queue = 0;              // 1 = AO rays pooled per warp (render_sec_queue); measured slower than the per-lane stage machine (6.31 vs 5.60 ms on config 4)
  int opt_layout = 0;                 // 0 = packed records (static scenes), 1 = raw reference pool (scenes under edit)
  int opt_stack = 0;                  // yv::kStackLocal / kStackRing4
  int opt_cull = 0;                   // 1 = skip child nodes the ray crosses through empty octants only (trace_core.cuh); measured slower, off

  // ---- device group (yv_renderer_create_multi) ---------------------------------------------------------------
  // The handle the caller holds is the leader (its own `device` is the first of the mask); peers[i] drives one further GPU.
  // Every frame call copies the leader's camera / options into the peers, gives member k the blocks b with b % n == k
  // (the SPU program's block stride, cell/spu/trace_spu.cpp:164) and joins the peers' streams into the leader's.
  std::vector<yv_renderer *> peers;
  yv_renderer *leader = nullptr;      // set on peers
  bool peer_access = true;            // every member can store into the leader's HBM (NVLink / PCIe P2P)
  void *worker = nullptr;             // peers: the host thread that issues this member's launches (yv_multi.cu, MemberWorker) —
                                      // SPURenderer's one thread per SPE (cell/spu_renderer.cpp:76-87)
  int opt_group_threads = 1;          // leader: 1 = every peer's share is launched by its own host thread, concurrently;
                                      // 0 = one loop on the calling thread (the launches of 8 GPUs then start ~18 us apart)
  int part_rows = 32;                 // rows per interleaved block (multiple of 16)
  int part_mode = 0;                  // 0 = interleaved blocks, 1 = contiguous bands
  cudaEvent_t ev_join = nullptr;      // on a peer: its share of the frame (and its copy) is done
  cudaEvent_t ev_own0 = nullptr, ev_own1 = nullptr;   // device time of this member's own share of the last group frame
  bool own_timed = false;
  double replicate_ms = 0.0;          // wall time of the last pool replication: the concurrent peer copies alone
  double replicate_alloc_ms = 0.0;    // ... and of the allocations on the peers that preceded them
  uint64_t replicate_bytes = 0;
  // ---- frames in flight -----------------------------------------------------------------------------------------
  static constexpr int kSlots = 4;
  yv_frame_slot slots[kSlots];
  size_t slot_pixels = 0;
  int opt_slots = 2;                  // frames in flight for yv_render_frame_async (2..4)
  long next_ticket = 0;
};

namespace yvi {

// yv_api.cu
int ensure_packed(yv_svo *svo);
int ensure_uploaded(yv_svo *svo, int device, DeviceSVO **out);
int ensure_frame_buffers(yv_renderer *r);
int launch_frame(yv_renderer *r, void *d_rgba);                 // asynchronous on r->stream
bool needs_second_pass(const yv_renderer *r);                   // Phong / show-normals / SSNA: the frame is re-read
bool single_pass_ssna(const yv_renderer *r);
void unbind_scene(yv_renderer *r);
void bind_scene(yv_renderer *r, yv_svo *svo);

// yv_multi.cu
inline int group_size(const yv_renderer *r) { return 1 + (int)r->peers.size(); }
inline yv_renderer *group_member(yv_renderer *r, int k) { return k == 0 ? r : r->peers[(size_t)k - 1]; }
// One frame over the whole group, asynchronous: every member renders its share into `target` (direct: a buffer all
// members can address — the leader's pinned host frame, or a buffer in the leader's HBM) or, when `staged`, into its
// own HBM (slot < 0: its frame buffer; else its staging buffer of that slot) and then moves its rows to `target` with
// its copy engine. done_stream: the leader stream on which the frame is complete when this returns (its work joined).
int group_launch(yv_renderer *r, void *target, bool staged, int slot, cudaStream_t *done_stream);
int group_render_frame(yv_renderer *r, const uint8_t **rgba);              // yv_render_frame on a group handle
int group_render_device(yv_renderer *r, void *d_rgba);                     // yv_render_frame_device_async on a group handle
void group_destroy_peers(yv_renderer *r);
void free_slots(yv_renderer *r);
// does member k of a group of n (the leader's partition settings) draw pixel row y?
bool member_owns_row(const yv_renderer *lead, int k, int n, int y);

}  // namespace yvi
