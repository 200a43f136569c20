// yv_api.cu — C ABI (include/yv_b200.h) over the host-side renderer state and kernel launches.
//
// Mirrors RendererBase (cell/renderer_base.h:7-61: camera state, setters, InitRayDir, the
// renderer-owned colour buffer) and the CUDA host sequence of demo/SVORenderer.cpp:95-149, with
// the Trace -> ShadeSimple launches fused into one kernel (render_kernels.cuh).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "yv_internal.h"
#include "render_kernels.cuh"
#include "svo_pack_gpu.h"

namespace yvi {

thread_local std::string g_err;

int fail(int code, const std::string &msg) { g_err = msg; return code; }

}  // namespace yvi

using namespace yvi;

namespace yvi {

int ensure_packed(yv_svo *svo) {
  if (svo->packed_ok && svo->packed_version == svo->dyn.version()) return YV_OK;
  std::string err;
  if (yv::pack_svo(svo->host, svo->packed, err) != 0) return fail(YV_ERR_FORMAT, err);
  if ((int)svo->packed.level_start.size() - 1 > yv::kMaxStack + 1)
    return fail(YV_ERR_FORMAT, "tree deeper than the traversal stack supports");
  svo->packed_ok = true;
  svo->packed_version = svo->dyn.version();
  return YV_OK;
}

// Host->device copy of a large pageable array through two pinned bounce buffers (a plain cudaMemcpy from
// pageable memory runs at ~1-2 GB/s here; staged, the copy of chunk k overlaps the host memcpy of chunk k+1).
int upload_staged(void *dst, const void *src, size_t bytes) {
  constexpr size_t kChunk = 32u << 20;
  if (bytes < 4 * kChunk) { YV_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice)); return YV_OK; }
  uint8_t *pin[2] = { nullptr, nullptr };
  cudaStream_t st = nullptr; cudaEvent_t ev[2] = { nullptr, nullptr };
  cudaError_t e = cudaMallocHost(&pin[0], kChunk);
  if (e == cudaSuccess) e = cudaMallocHost(&pin[1], kChunk);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
  size_t off = 0; int k = 0;
  while (e == cudaSuccess && off < bytes) {
    const size_t n = std::min(kChunk, bytes - off);
    if (k >= 2) e = cudaEventSynchronize(ev[k & 1]);             // bounce buffer free again?
    if (e != cudaSuccess) break;
    std::memcpy(pin[k & 1], (const uint8_t *)src + off, n);
    e = cudaMemcpyAsync((uint8_t *)dst + off, pin[k & 1], n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaEventRecord(ev[k & 1], st);
    off += n; ++k;
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (ev[0]) cudaEventDestroy(ev[0]);
  if (ev[1]) cudaEventDestroy(ev[1]);
  if (st) cudaStreamDestroy(st);
  cudaFreeHost(pin[0]); cudaFreeHost(pin[1]);
  if (e != cudaSuccess) return fail(YV_ERR_CUDA, std::string("staged upload: ") + cudaGetErrorString(e));
  return YV_OK;
}

void free_packed_device(DeviceSVO &d) {
  cudaFree(d.recs); cudaFree(d.octs); cudaFree(d.leaves); cudaFree(d.node_data);
  d.recs = nullptr; d.octs = nullptr; d.leaves = nullptr; d.node_data = nullptr; d.n_recs = d.n_leaves = 0;
}

// CudaSVO::Update (demo/SVORenderer.cpp:33-53; paging: reaction/report/main.tex:71): bring the device's raw
// mirror of the pool up to date by copying only the 256-node pages written since the last call.
int sync_raw_locked(yv_svo *svo, int device, DeviceSVO **out, uint64_t *bytes_out);

int sync_raw(yv_svo *svo, int device, DeviceSVO **out, uint64_t *bytes_out) {
  std::lock_guard<std::mutex> lock(svo->mu);
  return sync_raw_locked(svo, device, out, bytes_out);
}

int sync_raw_locked(yv_svo *svo, int device, DeviceSVO **out, uint64_t *bytes_out) {
  YV_CUDA(cudaSetDevice(device));
  DeviceSVO &d = svo->dev[device];
  if (svo->dyn.page_versions().empty() && !svo->host.nodes.empty()) svo->dyn.adopt_existing();
  const size_t n = svo->host.nodes.size();
  uint64_t bytes = 0;
  if (n > d.raw_capacity) {                       // (re)allocate with slack, then everything is dirty
    cudaFree(d.raw); d.raw = nullptr;
    d.raw_capacity = std::max<size_t>(n + n / 2, 1u << 16);
    YV_CUDA(cudaMalloc(&d.raw, d.raw_capacity * sizeof(yv_vox_node)));
    d.raw_version = 0;
  }
  const std::vector<uint32_t> &pv = svo->dyn.page_versions();
  size_t page = 0;
  bool quiesced = false;
  while (page < pv.size()) {
    if (pv[page] <= d.raw_version) { ++page; continue; }
    // Renderer streams are non-blocking, so nothing orders them against these copies: a frame launched with
    // yv_render_frame_device_async may still be traversing the pages about to be overwritten. Wait for the device
    // once per update that has anything to ship (an edit is host work of milliseconds; this is not the frame path).
    if (!quiesced) { YV_CUDA(cudaDeviceSynchronize()); quiesced = true; }
    size_t end = page;
    while (end < pv.size() && pv[end] > d.raw_version) ++end;          // one copy per run of dirty pages
    const size_t first = page * yv::kPageNodes, last = std::min(n, end * yv::kPageNodes);
    if (last > first) {
      int urc = upload_staged(d.raw + first, svo->host.nodes.data() + first, (last - first) * sizeof(yv_vox_node));
      if (urc) return urc;
      bytes += (last - first) * sizeof(yv_vox_node);
    }
    page = end;
  }
  // ... and the copies themselves (legacy-stream / private-stream, a pageable cudaMemcpy may return before its last DMA
  // has landed) are complete before any renderer stream can read the pages
  if (quiesced) YV_CUDA(cudaDeviceSynchronize());
  d.raw_version = svo->dyn.version();
  if (out) *out = &d;
  if (bytes_out) *bytes_out = bytes;
  return YV_OK;
}

// The device copy of the packed pool. Default: ship the raw pool (page-wise, sync_raw) and re-lay it out on the
// GPU (svo_pack_gpu.cu); YV_HOST_PACK=1, or a pool that is not a tree, uses the host BFS of svo_pack.cpp.
int ensure_uploaded(yv_svo *svo, int device, DeviceSVO **out) {
  std::lock_guard<std::mutex> lock(svo->mu);
  const uint32_t want = svo->dyn.version();
  auto it = svo->dev.find(device);
  if (it != svo->dev.end() && it->second.recs && it->second.packed_version != want)
    free_packed_device(it->second);               // the scene was edited since this copy was made
  if (it == svo->dev.end() || !it->second.recs) {
    YV_CUDA(cudaSetDevice(device));
    const char *host_env = std::getenv("YV_HOST_PACK");
    bool done = false;
    if (!(host_env && host_env[0] == '1') && !svo->host.nodes.empty()) {
      DeviceSVO *d = nullptr;
      const bool had_raw = it != svo->dev.end() && it->second.raw != nullptr;
      int rc = sync_raw_locked(svo, device, &d, nullptr);
      if (rc) return rc;
      yv::DevicePacked dp; std::string err;
      if (yv::pack_svo_on_device(d->raw, svo->host.nodes.size(), svo->host.root, dp, err) == 0) {
        if (dp.levels > yv::kMaxStack + 1) {
          cudaFree(dp.recs); cudaFree(dp.octs); cudaFree(dp.leaves); cudaFree(dp.node_data);
          return fail(YV_ERR_FORMAT, "tree deeper than the traversal stack supports");
        }
        d->recs = (uint4 *)dp.recs; d->octs = (uint2 *)dp.octs; d->leaves = dp.leaves; d->node_data = dp.node_data;
        d->n_recs = dp.n_recs; d->n_leaves = dp.n_leaves; d->levels = dp.levels;
        d->root_null = YV_IS_NULL(svo->host.root);
        if (!d->recs) {                            // null root: keep valid (dummy) pointers
          YV_CUDA(cudaMalloc(&d->recs, sizeof(uint4))); YV_CUDA(cudaMalloc(&d->octs, sizeof(uint2)));
          YV_CUDA(cudaMalloc(&d->leaves, sizeof(uint32_t)));
        }
        d->packed_version = want;
        done = true;
      }
      if (!had_raw) { cudaFree(d->raw); d->raw = nullptr; d->raw_capacity = 0; d->raw_version = 0; }   // only needed for the repack
      // "not a tree" (shared sub-trees) and CUDA failures (the repack transiently needs ~96 B/node against ~25 B/node
      // resident: a large scene can fit packed and still not fit the repack) go to the host pack; what is left is
      // a structural error
      const bool structural = err.find("not a tree") == std::string::npos && err.find("cuda") == std::string::npos &&
                              err.find("CUDA") == std::string::npos && err.find("memory") == std::string::npos;
      if (!done) cudaGetLastError();
      if (!done && structural) return fail(YV_ERR_FORMAT, err);
    }
    if (!done) {
      int rc = ensure_packed(svo);
      if (rc) return rc;
      DeviceSVO &d = svo->dev[device];
      d.n_recs = svo->packed.records.size();
      d.n_leaves = svo->packed.leaves.size();
      d.root_null = svo->packed.root_null;
      d.levels = (int)svo->packed.level_start.size() - 1;
      YV_CUDA(cudaMalloc(&d.recs, std::max<size_t>(1, d.n_recs) * sizeof(uint4)));
      YV_CUDA(cudaMalloc(&d.octs, std::max<size_t>(1, d.n_recs) * sizeof(uint2)));
      YV_CUDA(cudaMalloc(&d.leaves, std::max<size_t>(1, d.n_leaves) * sizeof(uint32_t)));
      if (d.n_recs) {
        std::vector<yv::DeviceRecord> trav; std::vector<yv::DeviceRecordOctants> octs;
        yv::device_layout(svo->packed, trav, octs);
        int urc = upload_staged(d.recs, trav.data(), d.n_recs * sizeof(uint4));
        if (!urc) urc = upload_staged(d.octs, octs.data(), d.n_recs * sizeof(uint2));
        if (urc) return urc;
      }
      if (d.n_leaves) { int urc = upload_staged(d.leaves, svo->packed.leaves.data(), d.n_leaves * sizeof(uint32_t)); if (urc) return urc; }
      d.packed_version = want;
    }
    YV_CUDA(cudaDeviceSynchronize());               // the copy is complete before any (non-blocking) renderer stream reads it
    it = svo->dev.find(device);
  }
  if (out) *out = &it->second;
  return YV_OK;
}

void free_frame_buffers(yv_renderer *r) {
  cudaSetDevice(r->device);
  cudaFree(r->d_fb); r->d_fb = nullptr;
  cudaFreeHost(r->h_fb); r->h_fb = nullptr;
  cudaFree(r->d_hit_node); cudaFree(r->d_hit_child); cudaFree(r->d_hit_t); cudaFree(r->d_counters);
  cudaFree(r->d_shade_rec); r->d_shade_rec = nullptr;
  cudaFree(r->d_zbuf[0]); cudaFree(r->d_zbuf[1]); r->d_zbuf[0] = r->d_zbuf[1] = nullptr;
  cudaFree(r->d_accum); r->d_accum = nullptr;
  cudaFree(r->d_ssna_counters); r->d_ssna_counters = nullptr;
  r->d_hit_node = nullptr; r->d_hit_child = nullptr; r->d_hit_t = nullptr; r->d_counters = nullptr;
  r->fb_pixels = 0;
}

int ensure_frame_buffers(yv_renderer *r) {
  const size_t n = (size_t)r->width * (size_t)r->height;
  YV_CUDA(cudaSetDevice(r->device));
  if (r->fb_pixels != n) {
    free_frame_buffers(r);
    YV_CUDA(cudaMalloc(&r->d_fb, std::max<size_t>(1, n) * 4));
    // portable + mapped: one address for the host and for every GPU (the members of a device group store into it);
    // a group's peers deliver into the leader's frame and have no host frame of their own
    if (!r->leader) YV_CUDA(cudaHostAlloc(&r->h_fb, std::max<size_t>(1, n) * 4, cudaHostAllocPortable | cudaHostAllocMapped));
    r->fb_pixels = n;
  }
  if (r->hits && !r->d_hit_node) {
    YV_CUDA(cudaMalloc(&r->d_hit_node, std::max<size_t>(1, n) * 4));
    YV_CUDA(cudaMalloc(&r->d_hit_child, std::max<size_t>(1, n) * 4));
    YV_CUDA(cudaMalloc(&r->d_hit_t, std::max<size_t>(1, n) * 4));
  }
  if (r->counters && !r->d_counters) YV_CUDA(cudaMalloc(&r->d_counters, std::max<size_t>(1, n) * 4));
  return YV_OK;
}

// RendererBase::InitRayDir (cell/renderer_base.h:50-61), float32 with the cg:: operator order
// (nest/include/geometry/primitives/point.h:416-440,462-499). `tan(cg::grad2rad(m_fov / 2)) / m_viewSize.x` (:56) has a
// float argument, so C++ picks the float overload and the division is a float division: da = tanf(rad) / (float)W —
// that is what the reference's own header computes when compiled (tests/test_reference_renderer.py).
struct ViewBasis { float fwd[3], right[3], down[3], d2; };   // SSNA: the frame InitRayDir builds; down = -up', d2 = 2*da

void init_ray_dir_raw(const float vdir[3], const float vup[3], float fov, int width, int height,
                      float dir0[3], float du[3], float dv[3], ViewBasis *basis = nullptr) {
  auto norm3 = [](const float v[3], float o[3]) {
    float d = 0.0f;
    d += v[0] * v[0]; d += v[1] * v[1]; d += v[2] * v[2];
    float n = sqrtf(d);
    o[0] = v[0] / n; o[1] = v[1] / n; o[2] = v[2] / n;
  };
  auto cross3 = [](const float a[3], const float b[3], float o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
  };
  float fwd[3], rightu[3], right[3], upv[3];
  norm3(vdir, fwd);
  cross3(fwd, vup, rightu);
  norm3(rightu, right);
  cross3(right, fwd, upv);
  const float half_deg = fov / 2;
  const float half_rad = half_deg * (float)(3.14159265358979323846 / 180.0);
  const float da = tanf(half_rad) / (float)width;
  const float w = (float)width, h = (float)height;
  for (int i = 0; i < 3; ++i) {
    du[i] = (2.0f * right[i]) * da;
    dv[i] = (-2.0f * upv[i]) * da;
    const float a = (du[i] * w) / 2.0f;
    const float b = (dv[i] * h) / 2.0f;
    dir0[i] = (fwd[i] - a) - b;
  }
  if (basis) {
    for (int i = 0; i < 3; ++i) { basis->fwd[i] = fwd[i]; basis->right[i] = right[i]; basis->down[i] = -upv[i]; }
    basis->d2 = 2.0f * da;
  }
}

// SVORenderer::InitBlur (demo/SVORenderer.cpp:55-79) with K = YV_BLURZ_KERN (spec: include/yv_format.h "SSNA")
void init_blur_taps(float *taps) {
  const int K = YV_BLURZ_KERN;
  const float h = (float)(K / 2), scale = 2.0f;
  float sum = 0.0f;
  for (int y = 0; y < K; ++y)
    for (int x = 0; x < K; ++x) {
      float tx = scale * ((float)x - h) / h, ty = scale * ((float)y - h) / h;
      tx = tx * tx; ty = ty * ty;
      const float v = (float)std::exp(-(double)(tx + ty));
      taps[y * K + x] = v;
      sum += v;
    }
  for (int i = 0; i < K * K; ++i) taps[i] /= sum;
}

// 2-D TMA descriptor of a z-buffer (width x height floats, dense rows): box = one staged BlurZ tile, 40 x 38 elements (the
// inner extent must be a multiple of 16 bytes), elements outside the tensor are delivered as zeros = "invalid". The encode
// function belongs to the driver API; it is fetched through the runtime, so nothing links against libcuda.
bool make_zbuf_tensor_map(CUtensorMap *map, float *zbuf, int width, int height) {
  typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiled encode = nullptr;
  static bool looked = false;
  if (!looked) {
    looked = true;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      encode = (EncodeTiled)fn;
    else
      cudaGetLastError();
  }
  if (!encode || !zbuf || width <= 0 || height <= 0 || width % 4 != 0) return false;
  const cuuint64_t gdim[2] = { (cuuint64_t)width, (cuuint64_t)height };
  const cuuint64_t gstride[1] = { (cuuint64_t)width * sizeof(float) };
  const cuuint32_t box[2] = { (cuuint32_t)yv::kTmaPitch, (cuuint32_t)yv::kBlurSpan };
  const cuuint32_t estride[2] = { 1u, 1u };
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2u, zbuf, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

void init_ray_dir(const yv_renderer *r, float dir0[3], float du[3], float dv[3], ViewBasis *basis = nullptr) {
  init_ray_dir_raw(r->dir, r->up, r->fov, r->width, r->height, dir0, du, dv, basis);
}

template <bool SEC, bool COUNT, int STACK, bool PERSISTENT, bool STAGED, bool LOD = false, bool RAW = false, bool JIT = false, bool CULL = false, bool ZB = false>
int launch_kernel(yv_renderer *r, const yv::RenderParams &p, size_t smem) {
  auto kern = yv::render_frame<SEC, COUNT, STACK, PERSISTENT, STAGED, LOD, RAW, JIT, CULL, ZB>;
  if (smem > 48 * 1024) YV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long grid;
  if (PERSISTENT) {
    int per_sm = 0;
    YV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, yv::kFrameCta, smem));
    if (per_sm < 1) per_sm = 1;
    grid = (long)r->sm_count * per_sm;
    const long warps_needed = ((long)p.num_tiles * 64 + 31) / 32;
    const long max_useful = (warps_needed + yv::kFrameCta / 32 - 1) / (yv::kFrameCta / 32);
    if (grid > max_useful) grid = std::max(1l, max_useful);
    YV_CUDA(cudaMemsetAsync(p.tile_counter, 0, sizeof(unsigned int), r->stream));
  } else {
    const long tiles16x8 = (long)((p.width + 15) / 16) * (p.num_tiles / p.tiles_x);      // four warps each
    grid = (tiles16x8 * 4 + yv::kFrameCta / 32 - 1) / (yv::kFrameCta / 32);
  }
  if (grid > 0) kern<<<(unsigned)grid, yv::kFrameCta, smem, r->stream>>>(p);
  YV_CUDA(cudaGetLastError());
  return YV_OK;
}

template <bool COUNT, int STACK>
int launch_queue(yv_renderer *r, const yv::RenderParams &p) {
  auto kern = yv::render_queue<COUNT, STACK>;
  const size_t smem = yv::stack_smem_bytes(STACK) + yv::kQueueSmemPerWarp * (yv::kCtaThreads / 32);
  YV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long grid = (long)((p.width + 31) / 32) * ((p.num_tiles / p.tiles_x + 1) / 2);
  if (grid > 0) kern<<<(unsigned)grid, yv::kCtaThreads, smem, r->stream>>>(p);
  YV_CUDA(cudaGetLastError());
  return YV_OK;
}

template <bool COUNT, bool LOD>
int launch_sec_queue(yv_renderer *r, const yv::RenderParams &p) {
  auto kern = yv::render_sec_queue<COUNT, LOD>;
  const size_t smem = yv::kAoSmemPerWarp * (yv::kCtaThreads / 32);
  const long grid = (long)((p.width + 15) / 16) * (p.num_tiles / p.tiles_x);
  if (grid > 0) kern<<<(unsigned)grid, yv::kCtaThreads, smem, r->stream>>>(p);
  YV_CUDA(cudaGetLastError());
  return YV_OK;
}

template <bool SEC, bool COUNT, int STACK, bool PERSISTENT>
int launch_staged(yv_renderer *r, const yv::RenderParams &p, size_t smem) {
  return p.smem_nodes > 0 ? launch_kernel<SEC, COUNT, STACK, PERSISTENT, true>(r, p, smem)
                          : launch_kernel<SEC, COUNT, STACK, PERSISTENT, false>(r, p, smem);
}

template <bool SEC, bool COUNT, int STACK>
int launch_schedule(yv_renderer *r, const yv::RenderParams &p, size_t smem) {
  return r->opt_persistent == 1 ? launch_staged<SEC, COUNT, STACK, true>(r, p, smem)
                           : launch_staged<SEC, COUNT, STACK, false>(r, p, smem);
}

// raw (reference-layout) pool: tiles schedule, local-memory stack
template <bool SEC, bool COUNT, bool LOD>
int launch_raw(yv_renderer *r, const yv::RenderParams &p) {
  return launch_kernel<SEC, COUNT, yv::kStackLocal, false, false, LOD, true>(r, p, 0);
}

// LOD variants exist for the local-memory stack without staging (the defaults)
template <bool SEC, bool COUNT>
int launch_lod(yv_renderer *r, const yv::RenderParams &p) {
  return r->opt_persistent == 1 ? launch_kernel<SEC, COUNT, yv::kStackLocal, true, false, true>(r, p, 0)
                                : launch_kernel<SEC, COUNT, yv::kStackLocal, false, false, true>(r, p, 0);
}

// octant culling (option "cull", off by default): packed pool, local-memory stack, no staging
template <bool SEC, bool COUNT, bool LOD>
int launch_cull(yv_renderer *r, const yv::RenderParams &p) {
  return r->opt_persistent == 1 ? launch_kernel<SEC, COUNT, yv::kStackLocal, true, false, LOD, false, false, true>(r, p, 0)
                                : launch_kernel<SEC, COUNT, yv::kStackLocal, false, false, LOD, false, false, true>(r, p, 0);
}

// displaced ray origins: primary rays, packed pool, tiles schedule, local-memory stack
template <bool COUNT, bool LOD>
int launch_jitter(yv_renderer *r, const yv::RenderParams &p) {
  return launch_kernel<false, COUNT, yv::kStackLocal, false, false, LOD, false, true>(r, p, 0);
}

template <bool SEC, bool COUNT>
int launch_stack(yv_renderer *r, const yv::RenderParams &p, size_t smem) {
  return r->opt_stack == yv::kStackRing4 ? launch_schedule<SEC, COUNT, yv::kStackRing4>(r, p, smem)
                                         : launch_schedule<SEC, COUNT, yv::kStackLocal>(r, p, smem);
}

int launch_frame(yv_renderer *r, void *d_rgba) {
  if (!r->svo) return fail(YV_ERR_NOSCENE, "no scene set");
  if (r->width <= 0 || r->height <= 0) return fail(YV_ERR_ARG, "resolution not set");
  DeviceSVO *ds = nullptr;
  const bool raw = r->opt_layout == 1;
  int rc = raw ? sync_raw(r->svo, r->device, &ds, nullptr) : ensure_uploaded(r->svo, r->device, &ds);
  if (rc) return rc;
  rc = ensure_frame_buffers(r);
  if (rc) return rc;
  YV_CUDA(cudaSetDevice(r->device));

  yv::RenderParams p;
  std::memset(&p, 0, sizeof p);
  if (raw) {
    p.recs = reinterpret_cast<const uint4 *>(ds->raw);
    p.root_valid = YV_IS_NULL(r->svo->host.root) ? 0u : 1u;
    p.root_index = p.root_valid ? r->svo->host.root : 0u;
  } else {
    p.recs = ds->recs; p.octs = ds->octs; p.leaves = ds->leaves;
    p.root_valid = ds->root_null ? 0u : 1u;
  }
  p.smem_nodes = raw ? 0u : (uint32_t)std::min<size_t>((size_t)std::max(0, r->opt_smem_nodes), ds->n_recs);
  for (int i = 0; i < 3; ++i) p.pos[i] = r->pos[i];
  ViewBasis basis;
  init_ray_dir(r, p.dir0, p.du, p.dv, &basis);
  const bool sec = r->shadow || r->ao_samples > 0;
  const bool ssna = r->ssna && !sec;
  for (int i = 0; i < 3; ++i) p.light[i] = (sec && r->shadow) ? r->light[i] : r->pos[i];
  p.width = r->width; p.height = r->height;
  p.y0 = r->rows_set ? std::max(0, r->y0) : 0;
  p.y1 = r->rows_set ? std::min(r->height, r->y1) : r->height;
  if (p.y1 < p.y0) p.y1 = p.y0;
  p.out_rgba = (uint32_t *)d_rgba;
  if (r->hits) { p.hit_node = r->d_hit_node; p.hit_child = r->d_hit_child; p.hit_t = r->d_hit_t; }
  p.counters = r->d_counters;
  p.tile_counter = r->d_tile_counter;
  // tile rows (8 pixel rows each) this launch covers
  int tile_rows;
  if (r->il_stride > 1) {
    p.y0 = 0; p.y1 = r->height;
    p.band_rows8 = r->il_rows / 8; p.band_stride = r->il_stride; p.band_phase = r->il_phase;
    const int blocks_total = (r->height + r->il_rows - 1) / r->il_rows;
    const int my_blocks = blocks_total > r->il_phase ? (blocks_total - r->il_phase + r->il_stride - 1) / r->il_stride : 0;
    tile_rows = my_blocks * p.band_rows8;
  } else {
    tile_rows = (p.y1 - p.y0 + 7) / 8;
    p.band_rows8 = std::max(1, tile_rows + 1); p.band_stride = 1; p.band_phase = 0;
  }
  p.tiles_x = (p.width + 7) / 8;
  p.num_tiles = p.tiles_x * tile_rows;
  p.refill_threshold = r->opt_refill;
  p.sec_threshold = r->opt_sec_threshold;
  bool any_light = false;
  for (int i = 0; i < YV_MAX_LIGHTS; ++i) { p.lights[i] = r->lights[i]; any_light = any_light || r->lights[i].enabled; }
  p.shade_mode = sec ? 0 : (r->show_normals ? 2 : (any_light ? 1 : 0));
  if (ssna) {
    if (p.y0 != 0 || p.y1 != r->height || r->il_stride > 1)
      return fail(YV_ERR_ARG, "SSNA needs the whole frame on one device (clear yv_set_rows / yv_set_interleave)");
    for (int i = 0; i < 2; ++i)
      if (!r->d_zbuf[i]) YV_CUDA(cudaMalloc(&r->d_zbuf[i], std::max<size_t>(1, r->fb_pixels) * sizeof(float)));
    p.ssna = 1;
    for (int i = 0; i < 3; ++i) { p.fwd[i] = basis.fwd[i]; p.right[i] = basis.right[i]; p.down[i] = basis.down[i]; }
    p.d2 = basis.d2;
  }
  if (p.shade_mode != 0 || ssna) {
    if (!r->d_shade_rec) YV_CUDA(cudaMalloc(&r->d_shade_rec, std::max<size_t>(1, r->fb_pixels) * sizeof(uint2)));
    p.shade_rec = r->d_shade_rec;
    // the trace kernel draws into the renderer's own HBM frame; the pass that finishes the pixels writes them where the
    // caller wants them (a pinned host frame for yv_render_frame: no D2H copy afterwards)
    if (d_rgba != (void *)r->d_fb) { p.out_rgba = r->d_fb; p.final_rgba = (uint32_t *)d_rgba; }
  }
  const bool jitter = r->jitter_amp > 0.0f;
  if (jitter) {
    if (sec || raw || r->opt_persistent != 0 || r->opt_stack != yv::kStackLocal || r->opt_smem_nodes > 0)
      return fail(YV_ERR_ARG, "displaced ray origins: primary rays with the default schedule, stack and layout only");
    p.jitter_amp = r->jitter_amp; p.jitter_seed = r->jitter_seed;
  }
  p.shadow = r->shadow; p.ao_samples = r->ao_samples; p.seed = r->seed;
  p.voxel_size = r->voxel_size; p.ao_max_t = r->ao_max_t;
  const bool lod = r->detail_coef > 0.0f;
  if (lod) {
    if (!raw && !ds->node_data) {
      { std::lock_guard<std::mutex> lock(r->svo->mu); int prc = ensure_packed(r->svo); if (prc) return prc; }
      const std::vector<uint32_t> &nd = r->svo->packed.node_data;
      YV_CUDA(cudaMalloc(&ds->node_data, std::max<size_t>(1, nd.size()) * sizeof(uint32_t)));
      if (!nd.empty()) YV_CUDA(cudaMemcpy(ds->node_data, nd.data(), nd.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    // rp.detailCoef = m_detailCoef * grad2rad(m_fov / 2) / m_viewSize.x   (demo/SVORenderer.cpp:104)
    const float half_rad = (r->fov / 2) * (float)(3.14159265358979323846 / 180.0);
    p.detail = (r->detail_coef * half_rad) / (float)r->width;
    p.node_data = ds->node_data;
    p.smem_nodes = 0;
  }
  const size_t smem = (lod || raw) ? 0 : (size_t)p.smem_nodes * sizeof(uint4) + yv::stack_smem_bytes(r->opt_stack, yv::kFrameCta);
  if (smem > 227 * 1024) return fail(YV_ERR_ARG, "shared-memory request exceeds 227 KB (lower smem_nodes or change stack)");

  if (!r->suppress_events) YV_CUDA(cudaEventRecord(r->ev0, r->stream));
  const int key = (sec ? 2 : 0) | (r->counters ? 1 : 0);
  const bool cull = r->opt_cull && !raw && !jitter && r->opt_stack == yv::kStackLocal && p.smem_nodes == 0 &&
                    r->opt_persistent != 2 && !(sec && r->opt_sec_queue && r->opt_persistent != 1);
  if (cull) {
    switch (key | (lod ? 4 : 0)) {
      case 0: rc = launch_cull<false, false, false>(r, p); break;
      case 1: rc = launch_cull<false, true, false>(r, p); break;
      case 2: rc = launch_cull<true, false, false>(r, p); break;
      case 3: rc = launch_cull<true, true, false>(r, p); break;
      case 4: rc = launch_cull<false, false, true>(r, p); break;
      case 5: rc = launch_cull<false, true, true>(r, p); break;
      case 6: rc = launch_cull<true, false, true>(r, p); break;
      default: rc = launch_cull<true, true, true>(r, p); break;
    }
  } else if (jitter) {
    if (lod) rc = r->counters ? launch_jitter<true, true>(r, p) : launch_jitter<false, true>(r, p);
    else rc = r->counters ? launch_jitter<true, false>(r, p) : launch_jitter<false, false>(r, p);
  } else if (sec && !raw && r->opt_sec_queue && r->opt_persistent != 1) {      // pooled AO rays (config 4)
    if (lod) rc = r->counters ? launch_sec_queue<true, true>(r, p) : launch_sec_queue<false, true>(r, p);
    else rc = r->counters ? launch_sec_queue<true, false>(r, p) : launch_sec_queue<false, false>(r, p);
  } else if (raw) {
    switch (key | (lod ? 4 : 0)) {
      case 0: rc = launch_raw<false, false, false>(r, p); break;
      case 1: rc = launch_raw<false, true, false>(r, p); break;
      case 2: rc = launch_raw<true, false, false>(r, p); break;
      case 3: rc = launch_raw<true, true, false>(r, p); break;
      case 4: rc = launch_raw<false, false, true>(r, p); break;
      case 5: rc = launch_raw<false, true, true>(r, p); break;
      case 6: rc = launch_raw<true, false, true>(r, p); break;
      default: rc = launch_raw<true, true, true>(r, p); break;
    }
  } else if (lod) {
    switch (key) {
      case 0: rc = launch_lod<false, false>(r, p); break;
      case 1: rc = launch_lod<false, true>(r, p); break;
      case 2: rc = launch_lod<true, false>(r, p); break;
      default: rc = launch_lod<true, true>(r, p); break;
    }
  } else if (r->opt_persistent == 2 && !sec) {        // per-warp ray queue (primary rays)
    if (r->opt_stack == yv::kStackRing4) rc = r->counters ? launch_queue<true, yv::kStackRing4>(r, p) : launch_queue<false, yv::kStackRing4>(r, p);
    else rc = r->counters ? launch_queue<true, yv::kStackLocal>(r, p) : launch_queue<false, yv::kStackLocal>(r, p);
  } else if (ssna && key == 0 && r->opt_persistent == 0 && r->opt_stack == yv::kStackLocal && p.smem_nodes == 0) {
    // SSNA on the default schedule: the trace kernel's hit epilogue writes the z-buffer (render_frame<..., ZB>); every
    // other variant leaves it to ssna_z_pass below
    p.zbuf = r->d_zbuf[0];
    rc = launch_kernel<false, false, yv::kStackLocal, false, false, false, false, false, false, true>(r, p, smem);
  } else
  switch (key) {
    case 0: rc = launch_stack<false, false>(r, p, smem); break;
    case 1: rc = launch_stack<false, true>(r, p, smem); break;
    case 2: rc = launch_stack<true, false>(r, p, smem); break;
    default: rc = launch_stack<true, true>(r, p, smem); break;
  }
  if (rc) return rc;
  int launches = 1;
  if (ssna && p.num_tiles > 0) {                    // z-buffer, then BlurZ x5 (demo/SVORenderer.cpp:126-141)
    if (!p.zbuf) {
      dim3 zgrid((p.width + 31) / 32, (p.height + 7) / 8);
      p.zbuf = r->d_zbuf[0];
      yv::ssna_z_pass<<<zgrid, 256, 0, r->stream>>>(p);
      ++launches;
    }
    yv::BlurParams b;
    std::memcpy(b.taps, r->blur_taps, sizeof b.taps);
    b.wsum = 0.0f;
    for (int i = 0; i < YV_BLURZ_KERN * YV_BLURZ_KERN; ++i) b.wsum += b.taps[i];      // float, in tap order: what the tested form accumulates
    b.width = p.width; b.height = p.height;
    const float pixel_ang = (r->fov * (float)(3.14159265358979323846 / 180.0)) / (float)r->width;    // rp.pixelAng (:105)
    float zlimit[YV_BLURZ_PASSES];
    float blur_size = 3;
    for (int i = 0; i < YV_BLURZ_PASSES; ++i, blur_size += 3) zlimit[i] = (5.0f * r->ssna_voxel_size) / (pixel_ang * blur_size);
    const int tiles_x = (p.width + yv::kBlurTile - 1) / yv::kBlurTile, tiles_y = (p.height + yv::kBlurTile - 1) / yv::kBlurTile;
    if (r->opt_ssna_fused) {
      // BlurZ x5 + ShadeSimple as one persistent cooperative launch (render_kernels.cuh, ssna_post); tiles staged by TMA
      // when the rows of the z-buffers are 16-byte multiples (the tensor map's pitch rule), otherwise by plain loads
      if (!r->d_ssna_counters) {
        YV_CUDA(cudaMalloc(&r->d_ssna_counters, YV_BLURZ_PASSES * sizeof(unsigned int)));
        YV_CUDA(cudaMemsetAsync(r->d_ssna_counters, 0, YV_BLURZ_PASSES * sizeof(unsigned int), r->stream));
      }
      yv::SsnaPostParams q;
      std::memset(&q, 0, sizeof q);
      bool tma = r->opt_ssna_fused == 1 && p.width % 4 == 0;
      if (tma) {
        for (int i = 0; i < 2 && tma; ++i) tma = make_zbuf_tensor_map(&q.tmap[i], r->d_zbuf[i], p.width, p.height);
        if (!tma) cudaGetLastError();
      }
      int &grid_cap = tma ? r->ssna_post_grid_tma : r->ssna_post_grid;
      if (grid_cap == 0) {
        int per_sm = 0;
        if (tma) YV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, yv::ssna_post<true>, 256, 0));
        else YV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, yv::ssna_post<false>, 256, 0));
        grid_cap = std::max(1, per_sm) * r->sm_count;
      }
      q.p = p; q.b = b; q.b.src = nullptr; q.b.dst = nullptr; q.b.zlimit = 0.0f;
      q.zbuf[0] = r->d_zbuf[0]; q.zbuf[1] = r->d_zbuf[1];
      for (int i = 0; i < YV_BLURZ_PASSES; ++i) q.zlimit[i] = zlimit[i];
      q.counters = r->d_ssna_counters;
      q.tiles_x = tiles_x; q.tiles_y = tiles_y;
      const int grid = std::max(1, std::min(grid_cap, tiles_x * tiles_y));
      void *args[] = { (void *)&q };
      const void *kern = tma ? (const void *)yv::ssna_post<true> : (const void *)yv::ssna_post<false>;
      YV_CUDA(cudaLaunchCooperativeKernel(kern, dim3((unsigned)grid), dim3(256), args, 0, r->stream));
      ++launches;
    } else {
      int src = 0;
      dim3 bgrid(tiles_x, tiles_y);
      for (int i = 0; i < YV_BLURZ_PASSES; ++i) {
        b.zlimit = zlimit[i];
        b.src = r->d_zbuf[src]; b.dst = r->d_zbuf[1 - src];
        yv::blur_z_pass<<<bgrid, 256, 0, r->stream>>>(b);
        ++launches;
        src = 1 - src;
      }
      YV_CUDA(cudaGetLastError());
      p.zbuf = r->d_zbuf[src];
    }
  }
  if ((p.shade_mode != 0 || ssna) && !(ssna && r->opt_ssna_fused) && p.num_tiles > 0) {       // ShadeSimple pass over the rows this launch rendered
    dim3 grid((p.width + 31) / 32, p.num_tiles / p.tiles_x);
    yv::shade_pass<<<grid, 256, 0, r->stream>>>(p);
    YV_CUDA(cudaGetLastError());
    ++launches;
  }
  r->last_launches = launches;
  if (!r->suppress_events) {
    YV_CUDA(cudaEventRecord(r->ev1, r->stream));
    r->timed = true;
    r->launches = launches;
  }
  return YV_OK;
}

// RenderFrame with the device->host copy overlapped: the frame is cut into row chunks, chunk k renders on
// aux[k & 1] (so the tail wave of one chunk overlaps the head of the next) and its rows start travelling to the
// pinned host buffer as soon as its kernel is done. One kernel launch per chunk.
int render_frame_pipelined(yv_renderer *r) {
  const int H = r->height, W = r->width;
  const int chunks = std::min(std::max(r->opt_pipeline, 2), (int)yv_renderer::kChunks);
  // chunk k starts at row bound[k]; sizes fall geometrically by opt_pipeline_taper percent, rounded to 8-row tiles
  int bound[yv_renderer::kChunks + 1];
  {
    const double q = std::min(100, std::max(10, r->opt_pipeline_taper)) / 100.0;
    double total = 0.0, w = 1.0;
    for (int k = 0; k < chunks; ++k, w *= q) total += w;
    double acc = 0.0; w = 1.0;
    bound[0] = 0;
    for (int k = 0; k < chunks; ++k, w *= q) {
      acc += w;
      int b = (int)(H * (acc / total) / 8.0 + 0.5) * 8;
      bound[k + 1] = k == chunks - 1 ? H : std::min(H, std::max(b, bound[k]));
    }
  }
  cudaStream_t main_stream = r->stream;
  YV_CUDA(cudaEventRecord(r->ev0, main_stream));
  YV_CUDA(cudaEventRecord(r->ev_fork, main_stream));
  for (int i = 0; i < 2; ++i) YV_CUDA(cudaStreamWaitEvent(r->aux[i], r->ev_fork, 0));
  int rc = YV_OK, launched = 0;
  r->suppress_events = true;
  for (int k = 0; k < chunks && rc == YV_OK; ++k) {
    const int y0 = bound[k], y1 = bound[k + 1];
    if (y1 <= y0) continue;
    r->rows_set = true; r->y0 = y0; r->y1 = y1;
    r->stream = r->aux[k & 1];
    rc = launch_frame(r, r->d_fb);
    if (rc == YV_OK) {
      cudaEventRecord(r->ev_chunk[k], r->stream);
      cudaStreamWaitEvent(r->copy_stream, r->ev_chunk[k], 0);
      const size_t off = (size_t)y0 * W * 4, bytes = (size_t)(y1 - y0) * W * 4;
      cudaMemcpyAsync(r->h_fb + off, (const uint8_t *)r->d_fb + off, bytes, cudaMemcpyDeviceToHost, r->copy_stream);
      launched += r->last_launches;
    }
  }
  r->suppress_events = false;
  r->rows_set = false;
  r->stream = main_stream;
  if (rc) { cudaDeviceSynchronize(); return rc; }
  YV_CUDA(cudaEventRecord(r->ev_copy, r->copy_stream));
  YV_CUDA(cudaStreamWaitEvent(main_stream, r->ev_copy, 0));
  YV_CUDA(cudaEventRecord(r->ev1, main_stream));
  r->timed = true;
  r->launches = launched;
  YV_CUDA(cudaStreamSynchronize(main_stream));
  return YV_OK;
}

void unbind_scene(yv_renderer *r) {
  if (!r->svo) return;
  std::lock_guard<std::mutex> lock(r->svo->mu);
  std::vector<yv_renderer *> &b = r->svo->bound;
  b.erase(std::remove(b.begin(), b.end(), r), b.end());
  r->svo = nullptr;
}

void bind_scene(yv_renderer *r, yv_svo *svo) {
  if (r->svo == svo) return;
  unbind_scene(r);
  r->svo = svo;
  if (svo) { std::lock_guard<std::mutex> lock(svo->mu); svo->bound.push_back(r); }
}

bool single_pass_ssna(const yv_renderer *r) { return r->ssna && !(r->shadow || r->ao_samples > 0); }

bool needs_second_pass(const yv_renderer *r) {
  bool any_light = false;
  for (int i = 0; i < YV_MAX_LIGHTS; ++i) any_light = any_light || r->lights[i].enabled;
  return single_pass_ssna(r) || ((r->show_normals || any_light) && !(r->shadow || r->ao_samples > 0));
}

static void free_device_copies(yv_svo *svo) {
  for (auto &kv : svo->dev) {
    cudaSetDevice(kv.first);
    cudaDeviceSynchronize();                      // no renderer stream is still reading the copies
    cudaFree(kv.second.recs);
    cudaFree(kv.second.octs);
    cudaFree(kv.second.leaves);
    cudaFree(kv.second.node_data);
    cudaFree(kv.second.raw);
  }
  svo->dev.clear();
}

// run a scene builder into a fresh handle; a failed or throwing build leaves nothing behind
template <class Build>
int build_scene(yv_svo **out, Build &&build) {
  return guarded([&]() -> int {
    yv_svo *s = new yv_svo; std::string err;
    int rc = 0;
    try { rc = build(s->host, err); } catch (...) { delete s; throw; }
    if (rc) { delete s; return fail(YV_ERR_ARG, err); }
    *out = s;
    return YV_OK;
  });
}

}  // namespace yvi

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

const char *yv_last_error(void) { return g_err.c_str(); }
int yv_abi_version(void) { return 1; }

int yv_svo_load(const char *path, yv_svo **out) {
  if (!path || !out) return fail(YV_ERR_ARG, "null argument");
  return guarded([&]() -> int {
    yv_svo *s = new yv_svo;
    std::string err;
    int rc = 0;
    try { rc = yv::load_vox(path, s->host, err); } catch (...) { delete s; throw; }
    if (rc) { delete s; return fail(rc <= -10 ? YV_ERR_FORMAT : YV_ERR_IO, err); }
    *out = s;
    return YV_OK;
  });
}

// SVOData::Load on an object that renderers already hold (cell/svodata.h:31-50 reloads in place; SetScene keeps the
// pointer, renderer_base.h:28): the pool is replaced inside the handle, device copies are dropped and re-made at the
// next frame, and every renderer bound to the handle keeps a valid scene. On failure the old scene stays.
int yv_svo_load_into(yv_svo *svo, const char *path) {
  if (!svo || !path) return fail(YV_ERR_ARG, "null argument");
  return guarded([&]() -> int {
    yv::HostSVO fresh;
    std::string err;
    const int rc = yv::load_vox(path, fresh, err);
    if (rc) return fail(rc <= -10 ? YV_ERR_FORMAT : YV_ERR_IO, err);
    std::lock_guard<std::mutex> lock(svo->mu);
    free_device_copies(svo);
    svo->host.root = fresh.root; svo->host.depth = fresh.depth; svo->host.nodes.swap(fresh.nodes);
    svo->dyn.reset_after_reload();
    svo->packed = yv::PackedSVO(); svo->packed_ok = false;
    return YV_OK;
  });
}

int yv_svo_from_memory(yv_node_id root, const yv_vox_node *nodes, uint32_t count, yv_svo **out) {
  if (!out || (!nodes && count)) return fail(YV_ERR_ARG, "null argument");
  return guarded([&]() -> int {
    yv_svo *s = new yv_svo;
    try {
      s->host.root = root;
      s->host.nodes.assign(nodes, nodes + count);
      yv::normalize_flags(s->host);
    } catch (...) { delete s; throw; }
    std::string err;
    if (yv::validate(s->host, err)) { delete s; return fail(YV_ERR_FORMAT, err); }
    *out = s;
    return YV_OK;
  });
}

int yv_svo_save(const yv_svo *svo, const char *path) {
  if (!svo || !path) return fail(YV_ERR_ARG, "null argument");
  std::string err;
  if (yv::save_vox(path, svo->host, err)) return fail(YV_ERR_IO, err);
  return YV_OK;
}

void yv_svo_free(yv_svo *svo) {
  if (!svo) return;
  // renderers that still hold this scene go back to "no scene" (RenderFrame -> NULL, cell/ppu_renderer.cpp:78-79)
  // instead of keeping a dangling pointer
  for (yv_renderer *r : std::vector<yv_renderer *>(svo->bound)) {
    if (r->own_stream) { cudaSetDevice(r->device); cudaStreamSynchronize(r->stream); }
    r->svo = nullptr;
  }
  svo->bound.clear();
  free_device_copies(svo);
  delete svo;
}

yv_node_id yv_svo_root(const yv_svo *svo) { return svo ? svo->host.root : YV_EMPTY_NODE; }
uint32_t yv_svo_node_count(const yv_svo *svo) { return svo ? (uint32_t)svo->host.nodes.size() : 0u; }
uint32_t yv_svo_depth(const yv_svo *svo) { return svo ? svo->host.depth : 0u; }
const yv_vox_node *yv_svo_nodes(const yv_svo *svo) { return svo && !svo->host.nodes.empty() ? svo->host.nodes.data() : nullptr; }

int yv_svo_build_sphere_fractal(int depth, int threads, yv_svo **out) {
  if (!out) return fail(YV_ERR_ARG, "null argument");
  return build_scene(out, [&](yv::HostSVO &h, std::string &err) { return yv::build_sphere_fractal(depth, threads, h, err); });
}
int yv_svo_build_iso_volume(int depth, uint32_t seed, int iso_level, int threads, yv_svo **out) {
  if (!out) return fail(YV_ERR_ARG, "null argument");
  return build_scene(out, [&](yv::HostSVO &h, std::string &err) { return yv::build_iso_volume(depth, seed, iso_level, threads, h, err); });
}
int yv_svo_build_single_sphere(int depth, int cx, int cy, int cz, int radius, uint8_t r, uint8_t g, uint8_t b, yv_svo **out) {
  if (!out) return fail(YV_ERR_ARG, "null argument");
  return build_scene(out, [&](yv::HostSVO &h, std::string &err) { return yv::build_single_sphere(depth, cx, cy, cz, radius, r, g, b, h, err); });
}
int yv_svo_build_from_dense(int depth, const uint32_t *voxdata, yv_svo **out) {
  if (!out || !voxdata) return fail(YV_ERR_ARG, "null argument");
  return build_scene(out, [&](yv::HostSVO &h, std::string &err) { return yv::build_from_dense(depth, voxdata, h, err); });
}
uint32_t yv_pack_voxdata(uint8_t r, uint8_t g, uint8_t b, float nx, float ny, float nz) {
  return yv::pack_voxdata(r, g, b, nx, ny, nz);
}

int yv_svo_upload(yv_svo *svo, int device) {
  if (!svo) return fail(YV_ERR_ARG, "null scene");
  return guarded([&]() -> int { return ensure_uploaded(svo, device, nullptr); });
}

uint64_t yv_svo_device_bytes(const yv_svo *svo, int device) {
  if (!svo) return 0;
  auto it = svo->dev.find(device);
  if (it == svo->dev.end()) return 0;
  return (uint64_t)it->second.n_recs * 24u + (uint64_t)it->second.n_leaves * 4u +
         (it->second.node_data ? (uint64_t)it->second.n_recs * 4u : 0u);
}

// device-side packed arrays copied back (tests: the GPU repack must equal the host repack bit for bit)
int yv_svo_device_packed_copy(yv_svo *svo, int device, uint32_t *n_records, uint32_t *n_leaves,
                              uint32_t *records_out, uint32_t *leaves_out, uint32_t *node_data_out) {
  if (!svo) return fail(YV_ERR_ARG, "null scene");
  DeviceSVO *d = nullptr;
  int rc = guarded([&]() -> int { return ensure_uploaded(svo, device, &d); });
  if (rc) return rc;
  if (n_records) *n_records = (uint32_t)d->n_recs;
  if (n_leaves) *n_leaves = (uint32_t)d->n_leaves;
  YV_CUDA(cudaSetDevice(device));
  if (records_out && d->n_recs) {                    // the canonical record { child_base, leaf_base, masks, orig_id } (svo_pack.h)
    return guarded([&]() -> int {
      std::vector<uint4> trav(d->n_recs);
      YV_CUDA(cudaMemcpy(trav.data(), d->recs, d->n_recs * 16, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < d->n_recs; ++i) {
        records_out[4 * i] = trav[i].x; records_out[4 * i + 1] = trav[i].z; records_out[4 * i + 2] = trav[i].y; records_out[4 * i + 3] = trav[i].w;
      }
      if (leaves_out && d->n_leaves) YV_CUDA(cudaMemcpy(leaves_out, d->leaves, d->n_leaves * 4, cudaMemcpyDeviceToHost));
      if (node_data_out && d->node_data) YV_CUDA(cudaMemcpy(node_data_out, d->node_data, d->n_recs * 4, cudaMemcpyDeviceToHost));
      return YV_OK;
    });
  }
  if (leaves_out && d->n_leaves) YV_CUDA(cudaMemcpy(leaves_out, d->leaves, d->n_leaves * 4, cudaMemcpyDeviceToHost));
  if (node_data_out && d->n_recs && d->node_data) YV_CUDA(cudaMemcpy(node_data_out, d->node_data, d->n_recs * 4, cudaMemcpyDeviceToHost));
  return YV_OK;
}

// the grandchild masks (one uint64 per record, svo_pack.h): host repack / as they sit on the device
int yv_svo_octant_masks(yv_svo *svo, uint64_t *out) {
  if (!svo || !out) return fail(YV_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> lock(svo->mu);
  int rc = guarded([&]() -> int { return ensure_packed(svo); });
  if (rc) return rc;
  if (!svo->packed.octants.empty()) std::memcpy(out, svo->packed.octants.data(), svo->packed.octants.size() * 8u);
  return YV_OK;
}

int yv_svo_device_octant_masks(yv_svo *svo, int device, uint64_t *out) {
  if (!svo || !out) return fail(YV_ERR_ARG, "null argument");
  DeviceSVO *d = nullptr;
  int rc = guarded([&]() -> int { return ensure_uploaded(svo, device, &d); });
  if (rc) return rc;
  return guarded([&]() -> int {
    YV_CUDA(cudaSetDevice(device));
    if (d->n_recs) YV_CUDA(cudaMemcpy(out, d->octs, d->n_recs * 8, cudaMemcpyDeviceToHost));   // { lo, hi } = little-endian uint64
    return YV_OK;
  });
}

int yv_svo_packed_counts(yv_svo *svo, uint32_t *records, uint32_t *leaves) {
  if (!svo) return fail(YV_ERR_ARG, "null scene");
  std::lock_guard<std::mutex> lock(svo->mu);
  int rc = guarded([&]() -> int { return ensure_packed(svo); });
  if (rc) return rc;
  if (records) *records = (uint32_t)svo->packed.records.size();
  if (leaves) *leaves = (uint32_t)svo->packed.leaves.size();
  return YV_OK;
}

int yv_svo_packed_copy(yv_svo *svo, uint32_t *records_out, uint32_t *leaves_out) {
  if (!svo) return fail(YV_ERR_ARG, "null scene");
  std::lock_guard<std::mutex> lock(svo->mu);
  int rc = guarded([&]() -> int { return ensure_packed(svo); });
  if (rc) return rc;
  if (records_out && !svo->packed.records.empty())
    std::memcpy(records_out, svo->packed.records.data(), svo->packed.records.size() * 16u);
  if (leaves_out && !svo->packed.leaves.empty())
    std::memcpy(leaves_out, svo->packed.leaves.data(), svo->packed.leaves.size() * 4u);
  return YV_OK;
}

int yv_init_ray_dir(const float dir[3], const float up[3], float fov_deg, int width, int height,
                    float dir0[3], float du[3], float dv[3]) {
  if (!dir || !up || !dir0 || !du || !dv || width <= 0 || height <= 0) return fail(YV_ERR_ARG, "bad argument");
  init_ray_dir_raw(dir, up, fov_deg, width, height, dir0, du, dv);
  return YV_OK;
}

int yv_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int yv_device_name(int device, char *buf, size_t len) {
  cudaDeviceProp prop;
  YV_CUDA(cudaGetDeviceProperties(&prop, device));
  if (buf && len) std::snprintf(buf, len, "%s", prop.name);
  return YV_OK;
}

int yv_renderer_create(int device, yv_renderer **out) {
  if (!out) return fail(YV_ERR_ARG, "null argument");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return fail(YV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  }
  if (device < 0 || device >= n) return fail(YV_ERR_ARG, "device index out of range");
  YV_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  YV_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(YV_ERR_CUDA, std::string("device is not sm_100-class: ") + prop.name);
  yv_renderer *r = new yv_renderer;
  r->device = device;
  r->sm_count = prop.multiProcessorCount;
  cudaError_t e = cudaStreamCreateWithFlags(&r->own_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&r->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&r->ev1);
  if (e == cudaSuccess) e = cudaEventCreate(&r->ev_own0);
  if (e == cudaSuccess) e = cudaEventCreate(&r->ev_own1);
  if (e == cudaSuccess) e = cudaMalloc(&r->d_tile_counter, sizeof(unsigned int));
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&r->aux[i], cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->ev_copy, cudaEventDisableTiming);
  for (int i = 0; i < yv_renderer::kChunks && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&r->ev_chunk[i], cudaEventDisableTiming);
  if (e != cudaSuccess) { std::string m = cudaGetErrorString(e); yv_renderer_destroy(r); return fail(YV_ERR_CUDA, m); }
  r->stream = r->own_stream;
  r->width = 640; r->height = 480;            // renderer_base.h:25
  init_blur_taps(r->blur_taps);               // InitBlur in the constructor (demo/SVORenderer.cpp:22)
  if (const char *e = std::getenv("YV_SSNA_FUSED")) r->opt_ssna_fused = std::max(0, std::min(2, std::atoi(e)));   // A/B runs of bench.py
  *out = r;
  return YV_OK;
}

void yv_renderer_destroy(yv_renderer *r) {
  if (!r) return;
  group_destroy_peers(r);
  cudaSetDevice(r->device);
  if (r->own_stream) cudaStreamSynchronize(r->own_stream);
  if (r->copy_stream) cudaStreamSynchronize(r->copy_stream);
  unbind_scene(r);
  free_slots(r);
  free_frame_buffers(r);
  if (r->ev_join) cudaEventDestroy(r->ev_join);
  if (r->ev_own0) cudaEventDestroy(r->ev_own0);
  if (r->ev_own1) cudaEventDestroy(r->ev_own1);
  cudaFree(r->d_tile_counter);
  if (r->ev0) cudaEventDestroy(r->ev0);
  if (r->ev1) cudaEventDestroy(r->ev1);
  if (r->ev_fork) cudaEventDestroy(r->ev_fork);
  if (r->ev_copy) cudaEventDestroy(r->ev_copy);
  for (int i = 0; i < yv_renderer::kChunks; ++i) if (r->ev_chunk[i]) cudaEventDestroy(r->ev_chunk[i]);
  for (int i = 0; i < 2; ++i) if (r->aux[i]) cudaStreamDestroy(r->aux[i]);
  if (r->copy_stream) cudaStreamDestroy(r->copy_stream);
  if (r->own_stream) cudaStreamDestroy(r->own_stream);
  delete r;
}

int yv_set_scene(yv_renderer *r, yv_svo *svo) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  if (r->stream) { cudaSetDevice(r->device); cudaStreamSynchronize(r->stream); }   // no frame of the old scene in flight
  bind_scene(r, svo);
  for (yv_renderer *p : r->peers) bind_scene(p, svo);
  return YV_OK;
}

int yv_set_view_pos(yv_renderer *r, const float pos[3]) {
  if (!r || !pos) return fail(YV_ERR_ARG, "null argument");
  for (int i = 0; i < 3; ++i) r->pos[i] = pos[i];
  return YV_OK;
}
int yv_set_view_dir(yv_renderer *r, const float dir[3]) {
  if (!r || !dir) return fail(YV_ERR_ARG, "null argument");
  for (int i = 0; i < 3; ++i) r->dir[i] = dir[i];
  return YV_OK;
}
int yv_set_view_up(yv_renderer *r, const float up[3]) {
  if (!r || !up) return fail(YV_ERR_ARG, "null argument");
  for (int i = 0; i < 3; ++i) r->up[i] = up[i];
  return YV_OK;
}

int yv_set_resolution(yv_renderer *r, int width, int height) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  if (width <= 0 || height <= 0 || (uint64_t)width * (uint64_t)height > 0x10000000ull)
    return fail(YV_ERR_ARG, "bad resolution");
  r->width = width; r->height = height;
  r->rows_set = false;
  r->il_stride = 1;
  return YV_OK;
}
int yv_get_resolution(const yv_renderer *r, int *width, int *height) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  if (width) *width = r->width;
  if (height) *height = r->height;
  return YV_OK;
}
int yv_set_fov(yv_renderer *r, float fov_deg) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  r->fov = fov_deg;
  return YV_OK;
}
int yv_set_light(yv_renderer *r, int index, const yv_light *light) {
  if (!r || !light) return fail(YV_ERR_ARG, "null argument");
  if (index < 0 || index >= YV_MAX_LIGHTS) return fail(YV_ERR_ARG, "light index out of range");
  r->lights[index] = *light;
  return YV_OK;
}
int yv_set_show_normals(yv_renderer *r, int enable) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  r->show_normals = enable != 0;
  return YV_OK;
}
int yv_get_show_normals(const yv_renderer *r, int *enable) {
  if (!r || !enable) return fail(YV_ERR_ARG, "null argument");
  *enable = r->show_normals ? 1 : 0;
  return YV_OK;
}
int yv_set_ssna(yv_renderer *r, int enable) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  r->ssna = enable != 0;
  return YV_OK;
}
int yv_get_ssna(const yv_renderer *r, int *enable) {
  if (!r || !enable) return fail(YV_ERR_ARG, "null argument");
  *enable = r->ssna ? 1 : 0;
  return YV_OK;
}
int yv_set_ssna_voxel_size(yv_renderer *r, float voxel_size) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  if (!(voxel_size >= 0.0f)) return fail(YV_ERR_ARG, "voxel size must be >= 0");
  r->ssna_voxel_size = voxel_size > 0.0f ? voxel_size : YV_SSNA_VOXEL_SIZE;
  return YV_OK;
}
int yv_set_jitter(yv_renderer *r, float amplitude, uint32_t seed) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  if (!(amplitude >= 0.0f)) return fail(YV_ERR_ARG, "jitter amplitude must be >= 0");
  r->jitter_amp = amplitude; r->jitter_seed = seed;
  return YV_OK;
}
int yv_set_detail_coef(yv_renderer *r, float coef) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  if (!(coef >= 0.0f)) return fail(YV_ERR_ARG, "detail coefficient must be >= 0");
  r->detail_coef = coef;
  return YV_OK;
}
int yv_get_detail_coef(const yv_renderer *r, float *coef) {
  if (!r || !coef) return fail(YV_ERR_ARG, "null argument");
  *coef = r->detail_coef;
  return YV_OK;
}
int yv_get_fov(const yv_renderer *r, float *fov_deg) {
  if (!r || !fov_deg) return fail(YV_ERR_ARG, "null argument");
  *fov_deg = r->fov;
  return YV_OK;
}

int yv_set_rows(yv_renderer *r, int y0, int y1) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  if (!r->peers.empty()) return fail(YV_ERR_ARG, "a multi-device renderer partitions the frame itself (yv_set_partition)");
  if (y0 < 0 || y1 < y0) return fail(YV_ERR_ARG, "bad row band");
  r->y0 = y0; r->y1 = y1; r->rows_set = true;
  r->il_stride = 1;
  return YV_OK;
}

int yv_set_interleave(yv_renderer *r, int band_rows, int stride, int phase) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  if (!r->peers.empty() && stride > 1) return fail(YV_ERR_ARG, "a multi-device renderer partitions the frame itself (yv_set_partition)");
  if (stride < 1 || phase < 0 || phase >= stride || band_rows < 16 || band_rows % 16 != 0)
    return fail(YV_ERR_ARG, "interleave: band_rows must be a multiple of 16, 0 <= phase < stride");
  r->il_rows = band_rows; r->il_stride = stride; r->il_phase = phase;
  if (stride > 1) r->rows_set = false;
  return YV_OK;
}

int yv_set_secondary(yv_renderer *r, int shadow, int ao_samples, uint32_t seed,
                     const float light_pos[3], float voxel_size, float ao_max_t) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  if (ao_samples < 0 || ao_samples > 16) return fail(YV_ERR_ARG, "ao_samples must be 0..16");
  r->shadow = shadow ? 1 : 0; r->ao_samples = ao_samples; r->seed = seed;
  if (light_pos) for (int i = 0; i < 3; ++i) r->light[i] = light_pos[i];
  r->voxel_size = voxel_size; r->ao_max_t = ao_max_t;
  return YV_OK;
}

int yv_enable_hits(yv_renderer *r, int enable) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  r->hits = enable != 0;
  return YV_OK;
}
int yv_enable_counters(yv_renderer *r, int enable) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  r->counters = enable != 0;
  return YV_OK;
}

int yv_render_frame_device_async(yv_renderer *r, void *d_rgba) {
  if (!r || !d_rgba) return fail(YV_ERR_ARG, "null argument");
  return guarded([&]() -> int {
    r->last_ms = -1.0f;
    return r->peers.empty() ? launch_frame(r, d_rgba) : group_render_device(r, d_rgba);
  });
}

int yv_sync(yv_renderer *r) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  YV_CUDA(cudaSetDevice(r->device));
  YV_CUDA(cudaStreamSynchronize(r->stream));      // a group's peers are joined into the leader's stream
  return YV_OK;
}

int yv_render_frame_device(yv_renderer *r, void *d_rgba) {
  int rc = yv_render_frame_device_async(r, d_rgba);
  if (rc) return rc;
  return yv_sync(r);
}

int yv_device_framebuffer(yv_renderer *r, void **d_rgba) {
  if (!r || !d_rgba) return fail(YV_ERR_ARG, "null argument");
  int rc = ensure_frame_buffers(r);
  if (rc) return rc;
  *d_rgba = r->d_fb;
  return YV_OK;
}

int yv_render_frame(yv_renderer *r, const uint8_t **rgba) {
  if (!r || !rgba) return fail(YV_ERR_ARG, "null argument");
  *rgba = nullptr;                                     // reference returns NULL on failure
  if (!r->svo) return fail(YV_ERR_NOSCENE, "no scene set");
  r->last_ms = -1.0f;
  if (!r->peers.empty()) return guarded([&]() -> int { return group_render_frame(r, rgba); });
  int rc = guarded([&]() -> int { return ensure_frame_buffers(r); });
  if (rc) return rc;
  const bool ssna = single_pass_ssna(r);               // BlurZ reaches across row chunks: one launch
  if (r->opt_zero_copy) {                       // (with a ShadeSimple / SSNA pass the trace kernel draws in HBM and that pass
    rc = launch_frame(r, r->h_fb);              // stores the finished pixels) pinned memory is device-addressable under UVA
    if (rc) return rc;
    YV_CUDA(cudaStreamSynchronize(r->stream));
    *rgba = r->h_fb;
    return YV_OK;
  }
  if (r->opt_pipeline > 1 && !r->rows_set && r->il_stride == 1 && r->opt_persistent != 1 && r->height >= 256 && !ssna) {
    rc = render_frame_pipelined(r);
    if (rc) return rc;
    *rgba = r->h_fb;
    return YV_OK;
  }
  rc = launch_frame(r, r->d_fb);
  if (rc) return rc;
  const int y0 = r->rows_set ? std::max(0, r->y0) : 0;
  const int y1 = r->rows_set ? std::min(r->height, r->y1) : r->height;
  if (y1 > y0) {   // (interleaved partitions copy the whole frame; rows of other ranks keep their old content)
    const size_t off = (size_t)y0 * r->width * 4, bytes = (size_t)(y1 - y0) * r->width * 4;
    YV_CUDA(cudaMemcpyAsync(r->h_fb + off, (const uint8_t *)r->d_fb + off, bytes, cudaMemcpyDeviceToHost, r->stream));
  }
  YV_CUDA(cudaStreamSynchronize(r->stream));
  *rgba = r->h_fb;
  return YV_OK;
}

int yv_render_accumulated(yv_renderer *r, int frames, const uint8_t **rgba) {
  if (!r || !rgba) return fail(YV_ERR_ARG, "null argument");
  *rgba = nullptr;
  if (frames < 1 || frames > 4096) return fail(YV_ERR_ARG, "frames must be 1..4096");
  if (!r->svo) return fail(YV_ERR_NOSCENE, "no scene set");
  int rc = ensure_frame_buffers(r);
  if (rc) return rc;
  YV_CUDA(cudaSetDevice(r->device));
  if (!r->d_accum) YV_CUDA(cudaMalloc(&r->d_accum, std::max<size_t>(1, r->fb_pixels) * sizeof(uint4)));
  const uint32_t pixels = (uint32_t)r->fb_pixels, grid = (pixels + 255u) / 256u;
  const uint32_t seed0 = r->jitter_seed;
  YV_CUDA(cudaEventRecord(r->ev0, r->stream));
  r->suppress_events = true;
  int launches = 0;
  for (int k = 0; k < frames && rc == YV_OK; ++k) {
    r->jitter_seed = seed0 + (uint32_t)k;
    rc = launch_frame(r, r->d_fb);
    if (rc == YV_OK && grid) {
      yv::accumulate_frame<<<grid, 256, 0, r->stream>>>(r->d_fb, r->d_accum, pixels, k == 0);
      launches += r->last_launches + 1;
    }
  }
  r->suppress_events = false;
  r->jitter_seed = seed0;
  if (rc) return rc;
  if (grid) { yv::resolve_frames<<<grid, 256, 0, r->stream>>>(r->d_accum, r->d_fb, pixels, (uint32_t)frames); ++launches; }
  YV_CUDA(cudaGetLastError());
  YV_CUDA(cudaEventRecord(r->ev1, r->stream));
  r->timed = true; r->launches = launches;
  YV_CUDA(cudaMemcpyAsync(r->h_fb, r->d_fb, r->fb_pixels * 4, cudaMemcpyDeviceToHost, r->stream));
  YV_CUDA(cudaStreamSynchronize(r->stream));
  *rgba = r->h_fb;
  return YV_OK;
}

// one per-pixel plane of a frame drawn by a group: every member holds the rows it rendered
static int gather_plane(yv_renderer *r, void *out, const void *(*plane)(const yv_renderer *)) {
  const int n = group_size(r);
  const size_t row = (size_t)r->width * 4, bytes = r->fb_pixels * 4;
  if (n == 1) { YV_CUDA(cudaSetDevice(r->device)); YV_CUDA(cudaMemcpy(out, plane(r), bytes, cudaMemcpyDeviceToHost)); return YV_OK; }
  std::vector<uint8_t> tmp(bytes);
  for (int k = 0; k < n; ++k) {
    yv_renderer *m = group_member(r, k);
    if (!plane(m) || m->fb_pixels != r->fb_pixels) return fail(YV_ERR_ARG, "per-pixel buffers not enabled before the last frame");
    YV_CUDA(cudaSetDevice(m->device));
    YV_CUDA(cudaMemcpy(tmp.data(), plane(m), bytes, cudaMemcpyDeviceToHost));
    for (int y = 0; y < r->height; ++y)
      if (member_owns_row(r, k, n, y)) std::memcpy((uint8_t *)out + y * row, tmp.data() + y * row, row);
  }
  return YV_OK;
}

int yv_get_hits(yv_renderer *r, uint32_t *node, int32_t *child, float *t) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  if (!r->hits || !r->d_hit_node) return fail(YV_ERR_ARG, "hit buffers not enabled before the last frame");
  YV_CUDA(cudaSetDevice(r->device));
  YV_CUDA(cudaStreamSynchronize(r->stream));
  return guarded([&]() -> int {
    int rc = YV_OK;
    if (node) rc = gather_plane(r, node, [](const yv_renderer *m) -> const void * { return m->d_hit_node; });
    if (!rc && child) rc = gather_plane(r, child, [](const yv_renderer *m) -> const void * { return m->d_hit_child; });
    if (!rc && t) rc = gather_plane(r, t, [](const yv_renderer *m) -> const void * { return m->d_hit_t; });
    return rc;
  });
}

int yv_get_counters(yv_renderer *r, uint32_t *fetches_per_ray) {
  if (!r || !fetches_per_ray) return fail(YV_ERR_ARG, "null argument");
  if (!r->counters || !r->d_counters) return fail(YV_ERR_ARG, "counters not enabled before the last frame");
  YV_CUDA(cudaSetDevice(r->device));
  YV_CUDA(cudaStreamSynchronize(r->stream));
  return guarded([&]() -> int {
    return gather_plane(r, fetches_per_ray, [](const yv_renderer *m) -> const void * { return m->d_counters; });
  });
}

// SVORenderer::DumpTraceData (demo/SVORenderer.cpp:158-192): <base>_<W>x<H>.dist / .color / .normal
int yv_dump_trace_data(yv_renderer *r, const char *fnbase) {
  if (!r || !fnbase) return fail(YV_ERR_ARG, "null argument");
  if (!r->svo) return fail(YV_ERR_NOSCENE, "no scene set");
  const size_t n = r->fb_pixels;
  std::vector<uint32_t> node(n); std::vector<int32_t> child(n); std::vector<float> dist(n);
  int rc = yv_get_hits(r, node.data(), child.data(), dist.data());
  if (rc) return rc;
  std::vector<uint8_t> color(n * 4, 0); std::vector<float> normal(n * 3, 0.0f);
  const std::vector<yv_vox_node> &pool = r->svo->host.nodes;
  for (size_t i = 0; i < n; ++i) {
    if (YV_IS_NULL(node[i]) || node[i] >= pool.size()) continue;                       // :171
    const yv_vox_node &nd = pool[node[i]];
    const uint32_t d = child[i] < 0 ? nd.data : nd.child[child[i] & 7];                // :176-179
    const uint32_t r5 = (d >> 11) & 31u, g6 = (d >> 5) & 63u, b5 = d & 31u;            // UnpackColor
    color[4 * i] = (uint8_t)((r5 << 3) | (r5 >> 2)); color[4 * i + 1] = (uint8_t)((g6 << 2) | (g6 >> 4));
    color[4 * i + 2] = (uint8_t)((b5 << 3) | (b5 >> 2)); color[4 * i + 3] = 255;
    float fx = (float)((d >> 16) & 255u) / 127.5f - 1.0f, fy = (float)((d >> 24) & 255u) / 127.5f - 1.0f;   // UnpackNormal
    float fz = (1.0f - fabsf(fx)) - fabsf(fy);
    if (fz < 0) { const float ox = (1.0f - fabsf(fy)) * (fx >= 0 ? 1.0f : -1.0f), oy = (1.0f - fabsf(fx)) * (fy >= 0 ? 1.0f : -1.0f); fx = ox; fy = oy; }
    const float len = sqrtf((fx * fx + fy * fy) + fz * fz);
    normal[3 * i] = fx / len; normal[3 * i + 1] = fy / len; normal[3 * i + 2] = fz / len;
  }
  const std::string base = std::string(fnbase) + "_" + std::to_string(r->width) + "x" + std::to_string(r->height);   // :188
  auto write = [&](const std::string &fn, const void *data, size_t bytes) {
    FILE *f = std::fopen(fn.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(data, 1, bytes, f) == bytes;
    return (std::fclose(f) == 0) && ok;
  };
  // the reference leaves .dist zero-filled (its `distBuf[i] = rd.t` is commented out, :168); the hit distance is written here
  if (!write(base + ".dist", dist.data(), n * 4) || !write(base + ".color", color.data(), n * 4) ||
      !write(base + ".normal", normal.data(), n * 12))
    return fail(YV_ERR_IO, "cannot write trace dump " + base);
  return YV_OK;
}

float yv_last_frame_ms(const yv_renderer *r) {
  if (!r || !r->timed) return -1.0f;
  if (r->last_ms >= 0.0f) return r->last_ms;          // measured at yv_wait_frame
  cudaSetDevice(r->device);
  if (cudaEventSynchronize(r->ev1) != cudaSuccess) return -1.0f;
  float ms = -1.0f;
  if (cudaEventElapsedTime(&ms, r->ev0, r->ev1) != cudaSuccess) return -1.0f;
  return ms;
}

int yv_last_frame_launches(const yv_renderer *r) { return r ? r->launches : 0; }

int yv_set_stream(yv_renderer *r, void *cuda_stream) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  r->stream = cuda_stream ? (cudaStream_t)cuda_stream : r->own_stream;
  return YV_OK;
}

int yv_set_option(yv_renderer *r, const char *name, int value) {
  if (!r || !name) return fail(YV_ERR_ARG, "null argument");
  std::string n(name);
  if (n == "smem_nodes") { if (value < 0 || value > 12288) return fail(YV_ERR_ARG, "smem_nodes must be 0..12288"); r->opt_smem_nodes = value; }
  else if (n == "persistent" || n == "schedule") {
    if (value < 0 || value > 2) return fail(YV_ERR_ARG, "schedule must be 0 (tiles), 1 (persistent) or 2 (queue)");
    r->opt_persistent = value;
  }
  else if (n == "sec_queue") r->opt_sec_queue = value ? 1 : 0;
  else if (n == "sec_threshold") { if (value < -1 || value > 31) return fail(YV_ERR_ARG, "sec_threshold must be -1..31"); r->opt_sec_threshold = value; }
  else if (n == "zero_copy") r->opt_zero_copy = value != 0;
  else if (n == "ssna_fused") r->opt_ssna_fused = value < 0 ? 0 : (value > 2 ? 2 : value);
  else if (n == "group_threads") r->opt_group_threads = value != 0;
  else if (n == "pipeline_taper") { if (value < 10 || value > 100) return fail(YV_ERR_ARG, "pipeline_taper must be 10..100 percent"); r->opt_pipeline_taper = value; }
  else if (n == "pipeline") { if (value < 0 || value > yv_renderer::kChunks) return fail(YV_ERR_ARG, "pipeline must be 0..8 chunks"); r->opt_pipeline = value; }
  else if (n == "layout") { if (value != 0 && value != 1) return fail(YV_ERR_ARG, "layout must be 0 (packed) or 1 (raw)"); r->opt_layout = value; }
  else if (n == "cull") r->opt_cull = value != 0;
  else if (n == "refill") { if (value < 0 || value > 31) return fail(YV_ERR_ARG, "refill must be 0..31"); r->opt_refill = value; }
  else if (n == "slots") {
    if (value < 2 || value > yv_renderer::kSlots) return fail(YV_ERR_ARG, "slots must be 2..4 frames in flight");
    for (int s = 0; s < yv_renderer::kSlots; ++s) if (r->slots[s].ticket >= 0) return fail(YV_ERR_ARG, "frames are in flight: yv_wait_frame first");
    r->opt_slots = value; r->next_ticket = 0;
  }
  else if (n == "stack") {
    if (value != yv::kStackLocal && value != yv::kStackRing4)
      return fail(YV_ERR_ARG, "stack must be 0 (local memory) or 4 (4-entry shared ring + local spill)");
    r->opt_stack = value;
  }
  else return fail(YV_ERR_ARG, "unknown option " + n);
  return YV_OK;
}

int yv_get_option(const yv_renderer *r, const char *name, int *value) {
  if (!r || !name || !value) return fail(YV_ERR_ARG, "null argument");
  std::string n(name);
  if (n == "smem_nodes") *value = r->opt_smem_nodes;
  else if (n == "persistent" || n == "schedule") *value = r->opt_persistent;
  else if (n == "sec_queue") *value = r->opt_sec_queue;
  else if (n == "sec_threshold") *value = r->opt_sec_threshold;
  else if (n == "pipeline") *value = r->opt_pipeline;
  else if (n == "pipeline_taper") *value = r->opt_pipeline_taper;
  else if (n == "zero_copy") *value = r->opt_zero_copy;
  else if (n == "ssna_fused") *value = r->opt_ssna_fused;
  else if (n == "group_threads") *value = r->opt_group_threads;
  else if (n == "layout") *value = r->opt_layout;
  else if (n == "cull") *value = r->opt_cull;
  else if (n == "refill") *value = r->opt_refill;
  else if (n == "stack") *value = r->opt_stack;
  else if (n == "slots") *value = r->opt_slots;
  else return fail(YV_ERR_ARG, "unknown option " + n);
  return YV_OK;
}

int yv_trace_rays(yv_renderer *r, const float *pos, const float *dir, uint32_t count,
                  uint32_t *node, int32_t *child, float *t) {
  if (!r || (!pos && count) || (!dir && count)) return fail(YV_ERR_ARG, "null argument");
  if (!r->svo) return fail(YV_ERR_NOSCENE, "no scene set");
  if (count == 0) return YV_OK;
  DeviceSVO *ds = nullptr;
  const bool raw = r->opt_layout == 1;
  int rc = raw ? sync_raw(r->svo, r->device, &ds, nullptr) : ensure_uploaded(r->svo, r->device, &ds);
  if (rc) return rc;
  YV_CUDA(cudaSetDevice(r->device));
  float *d_pos = nullptr, *d_dir = nullptr, *d_t = nullptr; uint32_t *d_node = nullptr; int32_t *d_child = nullptr;
  const size_t n = count;
  cudaError_t e = cudaMalloc(&d_pos, n * 12);
  if (e == cudaSuccess) e = cudaMalloc(&d_dir, n * 12);
  if (e == cudaSuccess) e = cudaMalloc(&d_node, n * 4);
  if (e == cudaSuccess) e = cudaMalloc(&d_child, n * 4);
  if (e == cudaSuccess) e = cudaMalloc(&d_t, n * 4);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_pos, pos, n * 12, cudaMemcpyHostToDevice, r->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_dir, dir, n * 12, cudaMemcpyHostToDevice, r->stream);
  if (e == cudaSuccess) {
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (raw) {
      const uint32_t valid = YV_IS_NULL(r->svo->host.root) ? 0u : 1u;
      yv::trace_rays_kernel<true><<<grid, 128, 0, r->stream>>>(reinterpret_cast<const uint4 *>(ds->raw), nullptr, nullptr, valid,
                                                               valid ? r->svo->host.root : 0u, d_pos, d_dir, count, d_node, d_child, d_t);
    } else {
      yv::trace_rays_kernel<false><<<grid, 128, 0, r->stream>>>(ds->recs, ds->octs, ds->leaves, ds->root_null ? 0u : 1u, 0u,
                                                                d_pos, d_dir, count, d_node, d_child, d_t);
    }
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && node) e = cudaMemcpyAsync(node, d_node, n * 4, cudaMemcpyDeviceToHost, r->stream);
  if (e == cudaSuccess && child) e = cudaMemcpyAsync(child, d_child, n * 4, cudaMemcpyDeviceToHost, r->stream);
  if (e == cudaSuccess && t) e = cudaMemcpyAsync(t, d_t, n * 4, cudaMemcpyDeviceToHost, r->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(r->stream);
  cudaFree(d_pos); cudaFree(d_dir); cudaFree(d_node); cudaFree(d_child); cudaFree(d_t);
  if (e != cudaSuccess) return fail(YV_ERR_CUDA, cudaGetErrorString(e));
  return YV_OK;
}

// ---- DynamicSVO editing (ore/src/main.cpp:101-129) ------------------------------------------------------
struct yv_source { yv::VoxelSource *src; };

int yv_svo_create(yv_svo **out) {
  if (!out) return fail(YV_ERR_ARG, "null argument");
  *out = new yv_svo;
  return YV_OK;
}

int yv_source_sphere(int radius, uint8_t r, uint8_t g, uint8_t b, int inverted, yv_source **out) {
  if (!out || radius < 0) return fail(YV_ERR_ARG, "bad argument");
  *out = new yv_source{ new yv::SphereSource(radius, r, g, b, inverted != 0) };
  return YV_OK;
}
int yv_source_raw(const int size[3], const uint32_t *voxdata, yv_source **out) {
  if (!out || !size || !voxdata || size[0] < 1 || size[1] < 1 || size[2] < 1) return fail(YV_ERR_ARG, "bad argument");
  *out = new yv_source{ new yv::RawSource(size, voxdata) };
  return YV_OK;
}
int yv_source_raw_colors_normals(const int size[3], const uint8_t *colors_rgba, const int8_t *normals_xyzw, yv_source **out) {
  if (!out || !size || !colors_rgba || !normals_xyzw || size[0] < 1 || size[1] < 1 || size[2] < 1) return fail(YV_ERR_ARG, "bad argument");
  *out = new yv_source{ new yv::RawSource(size, colors_rgba, normals_xyzw) };
  return YV_OK;
}
int yv_source_iso(const int size[3], const uint8_t *data, int iso_level, int inside, uint8_t r, uint8_t g, uint8_t b, yv_source **out) {
  if (!out || !size || !data || size[0] < 1 || size[1] < 1 || size[2] < 1) return fail(YV_ERR_ARG, "bad argument");
  yv::IsoBrickSource *s = new yv::IsoBrickSource(size, data);
  s->SetIsoLevel(iso_level); s->SetInside(inside != 0); s->SetColor(r, g, b);
  *out = new yv_source{ s };
  return YV_OK;
}
void yv_source_free(yv_source *src) { if (src) { delete src->src; delete src; } }
int yv_source_size(const yv_source *src, int size[3], int pivot[3]) {
  if (!src) return fail(YV_ERR_ARG, "null source");
  if (size) src->src->GetSize(size);
  if (pivot) src->src->GetPivot(pivot);
  return YV_OK;
}

int yv_svo_build_range(yv_svo *svo, int level, const int pos[3], int mode, const yv_source *src) {
  if (!svo || !pos || !src) return fail(YV_ERR_ARG, "null argument");
  if (mode != 0 && mode != 1) return fail(YV_ERR_ARG, "mode must be 0 (GROW) or 1 (CLEAR)");
  std::lock_guard<std::mutex> lock(svo->mu);
  std::string err;
  if (svo->dyn.BuildRange(level, pos, mode ? yv::BuildMode::Clear : yv::BuildMode::Grow, *src->src, err))
    return fail(YV_ERR_ARG, err);
  return YV_OK;
}

uint32_t yv_svo_live_node_count(const yv_svo *svo) { return svo ? svo->dyn.GetNodeCount() : 0u; }
int yv_svo_node_count_by_level(const yv_svo *svo, int *counts, int capacity) {
  if (!svo) return 0;
  const std::vector<int> c = svo->dyn.GetNodeCountByLevel();
  for (int i = 0; i < (int)c.size() && i < capacity && counts; ++i) counts[i] = c[i];
  return (int)c.size();
}
uint32_t yv_svo_version(const yv_svo *svo) { return svo ? svo->dyn.version() : 0u; }
int yv_svo_count_changed_pages(const yv_svo *svo, uint32_t since_version) {
  return svo ? svo->dyn.CountChangedPages(since_version) : 0;
}
int yv_svo_update(yv_svo *svo, int device, uint64_t *bytes_transferred) {
  if (!svo) return fail(YV_ERR_ARG, "null scene");
  return sync_raw(svo, device, nullptr, bytes_transferred);
}

int yv_device_alloc(int device, size_t bytes, void **d_ptr) {
  if (!d_ptr || bytes == 0) return fail(YV_ERR_ARG, "bad argument");
  YV_CUDA(cudaSetDevice(device));
  YV_CUDA(cudaMalloc(d_ptr, bytes));
  YV_CUDA(cudaMemset(*d_ptr, 0, bytes));
  return YV_OK;
}

int yv_device_free(int device, void *d_ptr) {
  if (!d_ptr) return YV_OK;
  YV_CUDA(cudaSetDevice(device));
  YV_CUDA(cudaFree(d_ptr));
  return YV_OK;
}

int yv_copy_to_host(int device, void *dst_host, const void *src_device, size_t bytes) {
  if (!dst_host || !src_device) return fail(YV_ERR_ARG, "null argument");
  YV_CUDA(cudaSetDevice(device));
  YV_CUDA(cudaMemcpy(dst_host, src_device, bytes, cudaMemcpyDeviceToHost));
  return YV_OK;
}

int yv_ipc_export(void *d_ptr, uint8_t handle[64]) {
  if (!d_ptr || !handle) return fail(YV_ERR_ARG, "null argument");
  cudaIpcMemHandle_t h;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  YV_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
  std::memcpy(handle, &h, 64);
  return YV_OK;
}

int yv_ipc_open(int device, const uint8_t handle[64], void **d_ptr) {
  if (!handle || !d_ptr) return fail(YV_ERR_ARG, "null argument");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, 64);
  YV_CUDA(cudaSetDevice(device));
  YV_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return YV_OK;
}

int yv_ipc_close(void *d_ptr) {
  if (!d_ptr) return YV_OK;
  YV_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return YV_OK;
}

int yv_host_register(int device, void *host_ptr, size_t bytes, void **device_ptr) {
  if (!host_ptr || !device_ptr || bytes == 0) return fail(YV_ERR_ARG, "null argument");
  YV_CUDA(cudaSetDevice(device));
  YV_CUDA(cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
  void *d = nullptr;
  const cudaError_t e = cudaHostGetDevicePointer(&d, host_ptr, 0);
  if (e != cudaSuccess) {
    cudaHostUnregister(host_ptr);
    return fail(YV_ERR_CUDA, std::string("cudaHostGetDevicePointer: ") + cudaGetErrorString(e));
  }
  *device_ptr = d;
  return YV_OK;
}

int yv_host_unregister(void *host_ptr) {
  if (!host_ptr) return YV_OK;
  YV_CUDA(cudaHostUnregister(host_ptr));
  return YV_OK;
}

}  // extern "C"
