// yv_multi.cu — several GPUs behind ONE renderer handle, and frames in flight.
//
// Reference precedent: SPURenderer (cell/spu_renderer.cpp:30-90) drives all of the machine's accelerators inside one
// RenderFrame(): one worker per SPE, worker i renders the 16x16-pixel blocks b with b % threadNum == i
// (blockStart = i, blockStride = threadNum, :80-83; the block loop is cell/spu/trace_spu.cpp:164), every worker DMAs its
// finished blocks into the one colour buffer the PPU hands back (trace_spu.cpp:171-176). Here:
//
//   yv_renderer_create_multi(mask)   one handle; a leader renderer on the first GPU of the mask and one peer renderer
//                                    (own stream, own events) per further GPU, all in this process
//   scene                            uploaded and re-packed once, on the leader's GPU, then copied to the other GPUs
//                                    over NVLink (cudaMemcpyPeerAsync) — not N uploads through the host
//   frame                            member k renders the blocks b (part_rows rows each) with b % n == k. Its kernel's
//                                    final RGBA8 stores go straight into the frame the caller reads — the leader's
//                                    pinned host frame (every GPU over its own PCIe link) or a buffer in the leader's
//                                    HBM (peer stores over NVLink); the peers' streams are joined into the leader's
//   frames in flight                 yv_render_frame_async / yv_wait_frame: 2..4 frame slots, so that the delivery of
//                                    frame k (posted stores or copy-engine transfers) overlaps the traversal of k+1
//
// No collective library is involved: the frame shards into disjoint pixels and the only exchange is the final store.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>

#include "yv_internal.h"

using namespace yvi;

namespace {

// One host thread per peer GPU, alive as long as the peer (the reference starts one thread per SPE for every frame,
// cell/spu_renderer.cpp:76-87; here the threads persist and a frame costs them one condition-variable round trip).
// Issued from one loop, the launches of 8 GPUs start ~18 us apart (cudaSetDevice, events, parameter set-up, launch),
// which at 8K / 8 GPUs is 15 % of a member's 0.84 ms share; issued by 8 threads they start together.
struct MemberWorker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int()> job;
  std::atomic<bool> posted{ false };      // a job is waiting (what a spinning worker polls)
  bool has_job = false, done = true, quit = false;
  int rc = 0;
  std::string err;
  // After a frame the worker polls for the next one for a while before it goes to sleep on the condition variable: in a
  // flythrough the next frame's share arrives within a millisecond, and a condition-variable wake-up (10-30 us) would be
  // most of the launch skew between the members of a group. YV_WORKER_SPIN_US sets the polling time (default 2000, 0 = off).
  static long spin_us() {
    static const long v = [] { const char *e = std::getenv("YV_WORKER_SPIN_US"); return e ? std::max(0l, std::atol(e)) : 2000l; }();
    return v;
  }
  MemberWorker() {
    th = std::thread([this] {
      for (;;) {
        std::function<int()> j;
        if (spin_us() > 0) {
          const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(spin_us());
          while (!posted.load(std::memory_order_acquire) && std::chrono::steady_clock::now() < until) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
          }
        }
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [this] { return has_job || quit; });
          if (quit) return;
          j = std::move(job); has_job = false; posted.store(false, std::memory_order_relaxed);
        }
        int r;
        std::string e;
        try { r = j(); if (r) e = yv_last_error(); }
        catch (const std::exception &ex) { r = YV_ERR_CUDA; e = ex.what(); }
        catch (...) { r = YV_ERR_CUDA; e = "exception in a member launch"; }
        { std::lock_guard<std::mutex> lk(mu); rc = r; err = e; done = true; }
        cv.notify_all();
      }
    });
  }
  void submit(std::function<int()> j) {
    { std::lock_guard<std::mutex> lk(mu); job = std::move(j); has_job = true; done = false; posted.store(true, std::memory_order_release); }
    cv.notify_all();
  }
  int wait(std::string &why) {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [this] { return done; });
    why = err;
    return rc;
  }
  ~MemberWorker() {
    { std::lock_guard<std::mutex> lk(mu); quit = true; }
    cv.notify_all();
    if (th.joinable()) th.join();
  }
};

// the partition of member k of n: interleaved blocks (default) or contiguous bands of rows rounded to the tile height
void set_partition(yv_renderer *lead, yv_renderer *m, int k, int n) {
  if (lead->part_mode == 0) {
    m->il_rows = lead->part_rows; m->il_stride = n; m->il_phase = k; m->rows_set = false;
  } else {
    int per = (lead->height + n - 1) / n;
    per = (per + 7) / 8 * 8;
    m->il_stride = 1; m->rows_set = true;
    m->y0 = std::min(lead->height, k * per); m->y1 = std::min(lead->height, (k + 1) * per);
  }
}

// camera, shading and kernel options travel from the leader to the peers before every frame
void copy_state(const yv_renderer *lead, yv_renderer *m) {
  std::memcpy(m->pos, lead->pos, sizeof m->pos); std::memcpy(m->dir, lead->dir, sizeof m->dir);
  std::memcpy(m->up, lead->up, sizeof m->up);
  m->fov = lead->fov;
  std::memcpy(m->lights, lead->lights, sizeof m->lights);
  m->show_normals = lead->show_normals; m->ssna = false; m->ssna_voxel_size = lead->ssna_voxel_size;
  m->jitter_amp = lead->jitter_amp; m->jitter_seed = lead->jitter_seed;
  m->detail_coef = lead->detail_coef;
  m->width = lead->width; m->height = lead->height;
  m->shadow = lead->shadow; m->ao_samples = lead->ao_samples; m->seed = lead->seed;
  std::memcpy(m->light, lead->light, sizeof m->light);
  m->voxel_size = lead->voxel_size; m->ao_max_t = lead->ao_max_t;
  m->hits = lead->hits; m->counters = lead->counters;
  m->opt_smem_nodes = lead->opt_smem_nodes; m->opt_persistent = lead->opt_persistent; m->opt_refill = lead->opt_refill;
  m->opt_sec_threshold = lead->opt_sec_threshold; m->opt_sec_queue = lead->opt_sec_queue;
  m->opt_layout = lead->opt_layout; m->opt_stack = lead->opt_stack;
  m->opt_slots = lead->opt_slots; m->opt_cull = lead->opt_cull;
}

// Copy the packed pool of `svo` from src_dev to dst_dev with peer copies, asynchronously on `st` (a stream of dst_dev).
// Two steps so that a group can allocate on every peer first and then run (and time) all the copies concurrently:
// replicate_alloc returns 1 when a copy is needed, 0 when dst already holds this version.
int replicate_alloc(yv_svo *svo, int src_dev, int dst_dev, bool *needed) {
  std::lock_guard<std::mutex> lock(svo->mu);
  *needed = false;
  auto it = svo->dev.find(src_dev);
  if (it == svo->dev.end() || !it->second.recs) return fail(YV_ERR_ARG, "replicate: the scene is not resident on the source device");
  const DeviceSVO src = it->second;
  DeviceSVO &d = svo->dev[dst_dev];
  if (d.recs && d.packed_version == src.packed_version && d.n_recs == src.n_recs) return YV_OK;
  YV_CUDA(cudaSetDevice(dst_dev));
  YV_CUDA(cudaDeviceSynchronize());                          // nothing still reads the copy that is replaced
  cudaFree(d.recs); cudaFree(d.octs); cudaFree(d.leaves); cudaFree(d.node_data);
  d.recs = nullptr; d.octs = nullptr; d.leaves = nullptr; d.node_data = nullptr;
  d.packed_version = 0; d.n_recs = 0;
  const size_t rb = std::max<size_t>(1, src.n_recs) * sizeof(uint4), lb = std::max<size_t>(1, src.n_leaves) * sizeof(uint32_t);
  YV_CUDA(cudaMalloc(&d.recs, rb));
  YV_CUDA(cudaMalloc(&d.octs, std::max<size_t>(1, src.n_recs) * sizeof(uint2)));
  YV_CUDA(cudaMalloc(&d.leaves, lb));
  if (src.node_data && src.n_recs) YV_CUDA(cudaMalloc(&d.node_data, src.n_recs * sizeof(uint32_t)));
  *needed = true;
  return YV_OK;
}

int replicate_copy_async(yv_svo *svo, int src_dev, int dst_dev, cudaStream_t st, uint64_t *bytes) {
  std::lock_guard<std::mutex> lock(svo->mu);
  const DeviceSVO src = svo->dev[src_dev];
  DeviceSVO &d = svo->dev[dst_dev];
  YV_CUDA(cudaSetDevice(dst_dev));
  uint64_t moved = 0;
  if (src.n_recs) { YV_CUDA(cudaMemcpyPeerAsync(d.recs, dst_dev, src.recs, src_dev, src.n_recs * sizeof(uint4), st)); moved += src.n_recs * sizeof(uint4); }
  if (src.n_recs) { YV_CUDA(cudaMemcpyPeerAsync(d.octs, dst_dev, src.octs, src_dev, src.n_recs * sizeof(uint2), st)); moved += src.n_recs * sizeof(uint2); }
  if (src.n_leaves) { YV_CUDA(cudaMemcpyPeerAsync(d.leaves, dst_dev, src.leaves, src_dev, src.n_leaves * sizeof(uint32_t), st)); moved += src.n_leaves * sizeof(uint32_t); }
  if (d.node_data && src.node_data && src.n_recs) {
    YV_CUDA(cudaMemcpyPeerAsync(d.node_data, dst_dev, src.node_data, src_dev, src.n_recs * sizeof(uint32_t), st));
    moved += src.n_recs * sizeof(uint32_t);
  }
  d.n_recs = src.n_recs; d.n_leaves = src.n_leaves; d.root_null = src.root_null; d.levels = src.levels;
  d.packed_version = src.packed_version;
  if (bytes) *bytes = moved;
  return YV_OK;
}

// scene replicas, frame buffers, state and partition on every member
int group_prepare(yv_renderer *r, bool need_local_fb) {
  const int n = group_size(r);
  if (!r->svo) return fail(YV_ERR_NOSCENE, "no scene set");
  if (single_pass_ssna(r))
    return fail(YV_ERR_ARG, "SSNA needs the whole frame on one device: use a single-device renderer");
  if (r->opt_layout == 0) {
    DeviceSVO *d0 = nullptr;
    int rc = ensure_uploaded(r->svo, r->device, &d0);
    if (rc) return rc;
    // allocate on every peer, then run all the peer copies concurrently (every destination pulls over its own NVLink
    // ports through the switch) and time just them
    std::vector<yv_renderer *> todo;
    const auto ta = std::chrono::steady_clock::now();
    for (yv_renderer *p : r->peers) {
      if (p->device == r->device) continue;
      bool dup = false;
      for (yv_renderer *q : todo) dup = dup || q->device == p->device;
      if (dup) continue;
      bool needed = false;
      rc = replicate_alloc(r->svo, r->device, p->device, &needed);
      if (rc) return rc;
      if (needed) todo.push_back(p);
    }
    if (!todo.empty()) {
      uint64_t total = 0;
      const auto t0 = std::chrono::steady_clock::now();
      for (yv_renderer *p : todo) {
        uint64_t b = 0;
        rc = replicate_copy_async(r->svo, r->device, p->device, p->own_stream, &b);
        if (rc) return rc;
        total += b;
      }
      for (yv_renderer *p : todo) { YV_CUDA(cudaSetDevice(p->device)); YV_CUDA(cudaStreamSynchronize(p->own_stream)); }
      const auto t1 = std::chrono::steady_clock::now();
      r->replicate_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
      r->replicate_alloc_ms = std::chrono::duration<double, std::milli>(t0 - ta).count();
      r->replicate_bytes = total;
    }
  }
  for (int k = 0; k < n; ++k) {
    yv_renderer *m = group_member(r, k);
    if (k > 0) { copy_state(r, m); if (m->svo != r->svo) bind_scene(m, r->svo); }
    set_partition(r, m, k, n);
    if (k > 0 && (need_local_fb || m->hits || m->counters)) { int rc = ensure_frame_buffers(m); if (rc) return rc; }
  }
  return ensure_frame_buffers(r);
}

void clear_partition(yv_renderer *r) { r->il_stride = 1; r->rows_set = false; }

// member k's rows of the frame at `src` -> the same rows of the frame at `dst`, asynchronously on `st`
int copy_member_rows(yv_renderer *lead, int k, int n, void *dst, const void *src, cudaStream_t st) {
  const size_t row = (size_t)lead->width * 4;
  const int H = lead->height;
  if (lead->part_mode != 0) {
    int per = (H + n - 1) / n; per = (per + 7) / 8 * 8;
    const int y0 = std::min(H, k * per), y1 = std::min(H, (k + 1) * per);
    if (y1 > y0) YV_CUDA(cudaMemcpyAsync((uint8_t *)dst + y0 * row, (const uint8_t *)src + y0 * row, (size_t)(y1 - y0) * row, cudaMemcpyDefault, st));
    return YV_OK;
  }
  const int R = lead->part_rows;
  const size_t block = (size_t)R * row;
  const int blocks_total = (H + R - 1) / R;
  const int mine = blocks_total > k ? (blocks_total - k + n - 1) / n : 0;
  if (mine == 0) return YV_OK;
  const int last_block = k + (mine - 1) * n;                        // may be the frame's partial last block
  const bool last_partial = (last_block + 1) * R > H;
  const int full = last_partial ? mine - 1 : mine;
  const size_t off = (size_t)k * block, pitch = (size_t)n * block;
  if (full > 0)
    YV_CUDA(cudaMemcpy2DAsync((uint8_t *)dst + off, pitch, (const uint8_t *)src + off, pitch, block, (size_t)full, cudaMemcpyDefault, st));
  if (last_partial) {
    const size_t o = (size_t)last_block * block, bytes = (size_t)(H - last_block * R) * row;
    YV_CUDA(cudaMemcpyAsync((uint8_t *)dst + o, (const uint8_t *)src + o, bytes, cudaMemcpyDefault, st));
  }
  return YV_OK;
}

int ensure_slots(yv_renderer *r, bool staged, bool host_frames) {
  const size_t n = (size_t)r->width * (size_t)r->height;
  if (r->slot_pixels != n) {
    for (int s = 0; s < yv_renderer::kSlots; ++s)
      if (r->slots[s].ticket >= 0) return fail(YV_ERR_ARG, "resolution changed while frames are in flight: yv_wait_frame first");
    free_slots(r);
    r->slot_pixels = n;
  }
  YV_CUDA(cudaSetDevice(r->device));
  for (int s = 0; s < r->opt_slots; ++s) {
    yv_frame_slot &sl = r->slots[s];
    if (host_frames && !sl.h_fb) YV_CUDA(cudaHostAlloc(&sl.h_fb, std::max<size_t>(1, n) * 4, cudaHostAllocPortable | cudaHostAllocMapped));
    if (staged && !sl.d_fb) YV_CUDA(cudaMalloc(&sl.d_fb, std::max<size_t>(1, n) * 4));
    if (!sl.ev_begin) YV_CUDA(cudaEventCreate(&sl.ev_begin));
    if (!sl.ev_done) YV_CUDA(cudaEventCreate(&sl.ev_done));
  }
  return YV_OK;
}

}  // namespace

namespace yvi {

bool member_owns_row(const yv_renderer *lead, int k, int n, int y) {
  if (n <= 1) return true;
  if (lead->part_mode == 0) return (y / lead->part_rows) % n == k;
  int per = (lead->height + n - 1) / n; per = (per + 7) / 8 * 8;
  return y >= k * per && y < (k + 1) * per;
}

void free_slots(yv_renderer *r) {
  cudaSetDevice(r->device);
  for (int s = 0; s < yv_renderer::kSlots; ++s) {
    yv_frame_slot &sl = r->slots[s];
    if (sl.ev_done && sl.ticket >= 0) cudaEventSynchronize(sl.ev_done);
    cudaFreeHost(sl.h_fb); cudaFree(sl.d_fb);
    if (sl.ev_begin) cudaEventDestroy(sl.ev_begin);
    if (sl.ev_done) cudaEventDestroy(sl.ev_done);
    sl = yv_frame_slot();
  }
  r->slot_pixels = 0;
}

void group_destroy_peers(yv_renderer *r) {
  for (yv_renderer *p : r->peers) { delete static_cast<MemberWorker *>(p->worker); p->worker = nullptr; }
  for (yv_renderer *p : r->peers) { p->leader = nullptr; yv_renderer_destroy(p); }
  r->peers.clear();
}

int group_launch(yv_renderer *r, void *target, bool staged, int slot, cudaStream_t *done_stream) {
  const int n = group_size(r);
  int rc = group_prepare(r, staged);
  if (rc) return rc;
  if (staged && slot >= 0)
    for (int k = 1; k < n; ++k) { rc = ensure_slots(group_member(r, k), true, false); if (rc) return rc; }
  // the frame is complete on `done`: with staged delivery that is the leader's copy stream, so that the render streams
  // are free for the next frame while the copy engines still move this one
  cudaStream_t done = staged ? r->copy_stream : r->stream;
  YV_CUDA(cudaSetDevice(r->device));
  YV_CUDA(cudaEventRecord(r->ev0, r->stream));
  YV_CUDA(cudaEventRecord(r->ev_fork, r->stream));          // peers start after whatever precedes the frame on the leader
  // member k's share: its kernel(s) on its own stream, then (staged) its rows on its copy engine; ev_join marks the end
  std::vector<int> member_launches((size_t)n, 0);
  auto member_job = [&](int k) -> int {
    yv_renderer *m = group_member(r, k);
    YV_CUDA(cudaSetDevice(m->device));
    if (k > 0) YV_CUDA(cudaStreamWaitEvent(m->stream, r->ev_fork, 0));
    uint32_t *local = slot >= 0 ? m->slots[slot].d_fb : m->d_fb;
    m->suppress_events = true;
    YV_CUDA(cudaEventRecord(m->ev_own0, m->stream));
    const int mrc = launch_frame(m, staged ? (void *)local : target);
    m->suppress_events = false;
    if (mrc) return mrc;
    YV_CUDA(cudaEventRecord(m->ev_own1, m->stream));
    m->own_timed = true;
    member_launches[(size_t)k] = m->last_launches;
    cudaStream_t tail = m->stream;
    if (staged) {                                            // this member's rows: its HBM -> the frame, on its copy engine
      YV_CUDA(cudaEventRecord(m->ev_copy, m->stream));
      YV_CUDA(cudaStreamWaitEvent(m->copy_stream, m->ev_copy, 0));
      const int crc = copy_member_rows(r, k, n, target, local, m->copy_stream);
      if (crc) return crc;
      tail = m->copy_stream;
    }
    if (k > 0) YV_CUDA(cudaEventRecord(m->ev_join, tail));
    return YV_OK;
  };
  const bool threaded = r->opt_group_threads != 0 && n > 1;
  std::string why;
  if (threaded) {                                            // peers on their own threads, the leader's share on this one
    for (int k = 1; k < n; ++k) {
      yv_renderer *m = group_member(r, k);
      if (!m->worker) m->worker = new MemberWorker();
      static_cast<MemberWorker *>(m->worker)->submit([&member_job, k] { return member_job(k); });
    }
    rc = member_job(0);
    if (rc) why = yv_last_error();
    for (int k = 1; k < n; ++k) {
      std::string w;
      const int wrc = static_cast<MemberWorker *>(group_member(r, k)->worker)->wait(w);
      if (wrc && !rc) { rc = wrc; why = w; }
    }
    if (rc) fail(rc, why);                                   // the message of the member that failed, on this thread
  } else {
    for (int k = 0; k < n && rc == YV_OK; ++k) rc = member_job(k);
  }
  int launches = 0;
  for (int k = 0; k < n; ++k) launches += member_launches[(size_t)k];
  for (int k = 0; k < n; ++k) clear_partition(group_member(r, k));
  if (rc) { for (int k = 0; k < n; ++k) { cudaSetDevice(group_member(r, k)->device); cudaDeviceSynchronize(); } return rc; }
  YV_CUDA(cudaSetDevice(r->device));
  for (yv_renderer *p : r->peers) YV_CUDA(cudaStreamWaitEvent(done, p->ev_join, 0));
  YV_CUDA(cudaEventRecord(r->ev1, done));
  r->timed = true; r->launches = launches; r->last_ms = -1.0f;
  if (done_stream) *done_stream = done;
  return YV_OK;
}

int group_render_frame(yv_renderer *r, const uint8_t **rgba) {
  int rc = ensure_frame_buffers(r);
  if (rc) return rc;
  const bool staged = !r->opt_zero_copy || needs_second_pass(r);
  cudaStream_t done = nullptr;
  rc = group_launch(r, r->h_fb, staged, -1, &done);
  if (rc) return rc;
  YV_CUDA(cudaStreamSynchronize(done));
  *rgba = r->h_fb;
  return YV_OK;
}

int group_render_device(yv_renderer *r, void *d_rgba) {
  // direct peer stores need every member to address the target; frames with a second pass are assembled by peer copies
  const bool staged = needs_second_pass(r);
  if (!staged && !r->peer_access) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, d_rgba) != cudaSuccess || at.type == cudaMemoryTypeDevice) {
      cudaGetLastError();
      return fail(YV_ERR_CUDA, "the GPUs of this group cannot address each other's memory (no P2P): render into host memory");
    }
  }
  cudaStream_t done = nullptr;
  int rc = group_launch(r, d_rgba, staged, -1, &done);
  if (rc) return rc;
  if (done != r->stream) {                                   // callers synchronise r->stream (yv_sync)
    YV_CUDA(cudaEventRecord(r->ev_copy, done));
    YV_CUDA(cudaStreamWaitEvent(r->stream, r->ev_copy, 0));
  }
  return YV_OK;
}

}  // namespace yvi

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int yv_renderer_create_group(const int *devices, int count, yv_renderer **out) {
  if (!out || !devices || count < 1) return fail(YV_ERR_ARG, "bad device list");
  if (count > 64) return fail(YV_ERR_ARG, "at most 64 members");
  yv_renderer *lead = nullptr;
  int rc = yv_renderer_create(devices[0], &lead);
  if (rc) return rc;
  for (int i = 1; i < count; ++i) {
    yv_renderer *p = nullptr;
    rc = yv_renderer_create(devices[i], &p);
    if (rc) { const std::string why = yv_last_error(); yv_renderer_destroy(lead); return fail(rc, why); }
    p->leader = lead;
    lead->peers.push_back(p);
    cudaError_t e = cudaSetDevice(p->device);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming);
    if (e != cudaSuccess) { const std::string why = cudaGetErrorString(e); yv_renderer_destroy(lead); return fail(YV_ERR_CUDA, why); }
    if (p->device == lead->device) continue;                 // a GPU listed twice shares the leader's memory anyway
    // stores of this GPU's kernel into the leader's HBM, and peer copies of the pool in the other direction
    int can = 0;
    cudaDeviceCanAccessPeer(&can, p->device, lead->device);
    if (can) {
      e = cudaDeviceEnablePeerAccess(lead->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
      cudaGetLastError();
      cudaSetDevice(lead->device);
      cudaDeviceEnablePeerAccess(p->device, 0);
      cudaGetLastError();
    }
    if (!can) lead->peer_access = false;
  }
  *out = lead;
  return YV_OK;
}

int yv_renderer_create_multi(uint64_t device_mask, yv_renderer **out) {
  if (!out) return fail(YV_ERR_ARG, "null argument");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    cudaGetLastError();
    return fail(YV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  }
  if (device_mask == YV_ALL_DEVICES) device_mask = count >= 64 ? ~0ull : ((1ull << count) - 1ull);   // SPE_COUNT_USABLE_SPES
  if (device_mask == 0) return fail(YV_ERR_ARG, "empty device mask");
  std::vector<int> devs;
  for (int d = 0; d < 64; ++d) if ((device_mask >> d) & 1ull) devs.push_back(d);
  if (devs.back() >= count) return fail(YV_ERR_ARG, "device mask names a device that does not exist");
  return yv_renderer_create_group(devs.data(), (int)devs.size(), out);
}

int yv_renderer_device_count(const yv_renderer *r) { return r ? group_size(r) : 0; }

int yv_renderer_device(const yv_renderer *r, int k) {
  if (!r || k < 0 || k >= group_size(r)) return -1;
  return k == 0 ? r->device : r->peers[(size_t)k - 1]->device;
}

int yv_set_partition(yv_renderer *r, int mode, int band_rows) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  if (mode != 0 && mode != 1) return fail(YV_ERR_ARG, "partition mode must be 0 (interleaved blocks) or 1 (contiguous bands)");
  if (mode == 0 && (band_rows < 16 || band_rows % 16 != 0)) return fail(YV_ERR_ARG, "band_rows must be a multiple of 16");
  r->part_mode = mode;
  if (mode == 0) r->part_rows = band_rows;
  return YV_OK;
}

float yv_member_frame_ms(const yv_renderer *r, int k) {
  if (!r || k < 0 || k >= group_size(r)) return -1.0f;
  const yv_renderer *m = k == 0 ? r : r->peers[(size_t)k - 1];
  if (!m->own_timed) return -1.0f;
  cudaSetDevice(m->device);
  if (cudaEventSynchronize(m->ev_own1) != cudaSuccess) return -1.0f;
  float ms = -1.0f;
  if (cudaEventElapsedTime(&ms, m->ev_own0, m->ev_own1) != cudaSuccess) return -1.0f;
  return ms;
}

int yv_replicate_stats(const yv_renderer *r, double *ms, uint64_t *bytes) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  if (ms) *ms = r->replicate_ms;
  if (bytes) *bytes = r->replicate_bytes;
  return YV_OK;
}

int yv_svo_replicate(yv_svo *svo, int src_device, int dst_device) {
  if (!svo) return fail(YV_ERR_ARG, "null scene");
  if (src_device == dst_device) return YV_OK;
  return guarded([&]() -> int {
    int rc = ensure_uploaded(svo, src_device, nullptr);
    if (rc) return rc;
    YV_CUDA(cudaSetDevice(dst_device));
    int can = 0;
    cudaDeviceCanAccessPeer(&can, dst_device, src_device);
    if (can) { cudaDeviceEnablePeerAccess(src_device, 0); cudaGetLastError(); }     // without it the copy is staged through the host
    bool needed = false;
    rc = replicate_alloc(svo, src_device, dst_device, &needed);
    if (rc) return rc;
    if (needed) { rc = replicate_copy_async(svo, src_device, dst_device, nullptr, nullptr); if (rc) return rc; }
    YV_CUDA(cudaSetDevice(dst_device));
    YV_CUDA(cudaDeviceSynchronize());
    return YV_OK;
  });
}

// ---- frames in flight --------------------------------------------------------------------------------------------

int yv_render_frame_async(yv_renderer *r, void *dst, int *ticket) {
  if (!r || !ticket) return fail(YV_ERR_ARG, "null argument");
  if (!r->svo) return fail(YV_ERR_NOSCENE, "no scene set");
  if (r->width <= 0 || r->height <= 0) return fail(YV_ERR_ARG, "resolution not set");
  if (single_pass_ssna(r) && (group_size(r) > 1))
    return fail(YV_ERR_ARG, "SSNA needs the whole frame on one device: use a single-device renderer");
  return guarded([&]() -> int {
    const int s = (int)(r->next_ticket % r->opt_slots);
    yv_frame_slot &sl = r->slots[s];
    if (sl.ticket >= 0) return fail(YV_ERR_ARG, "all frame slots are in flight: yv_wait_frame the oldest ticket first");
    // delivery: the kernel stores into the target (zero_copy, frames drawn in one pass) or the frame is rendered into
    // HBM and moved by the copy engine while the next frame traverses
    const bool second = needs_second_pass(r);
    bool staged = !r->opt_zero_copy || second;
    if (dst && !second && group_size(r) == 1) {
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, dst) == cudaSuccess && at.type == cudaMemoryTypeDevice && at.device == r->device) staged = false;
      cudaGetLastError();
    }
    int rc = ensure_slots(r, staged, dst == nullptr);
    if (rc) return rc;
    void *target = dst ? dst : (void *)sl.h_fb;
    YV_CUDA(cudaSetDevice(r->device));
    YV_CUDA(cudaEventRecord(sl.ev_begin, r->stream));
    cudaStream_t done = r->stream;
    if (group_size(r) > 1) {
      rc = group_launch(r, target, staged, s, &done);
      if (rc) return rc;
      sl.launches = r->launches;
    } else {
      rc = ensure_frame_buffers(r);
      if (rc) return rc;
      const bool keep = r->suppress_events;
      r->suppress_events = true;
      rc = launch_frame(r, staged ? (void *)sl.d_fb : target);
      r->suppress_events = keep;
      if (rc) return rc;
      sl.launches = r->last_launches;
      if (staged) {
        YV_CUDA(cudaEventRecord(r->ev_copy, r->stream));
        YV_CUDA(cudaStreamWaitEvent(r->copy_stream, r->ev_copy, 0));
        // the rows this renderer draws (a band set with yv_set_rows travels alone; otherwise the whole frame)
        const int y0 = r->rows_set ? std::max(0, r->y0) : 0, y1 = r->rows_set ? std::min(r->height, r->y1) : r->height;
        if (y1 > y0) {
          const size_t off = (size_t)y0 * r->width * 4, bytes = (size_t)(y1 - y0) * r->width * 4;
          YV_CUDA(cudaMemcpyAsync((uint8_t *)target + off, (const uint8_t *)sl.d_fb + off, bytes, cudaMemcpyDefault, r->copy_stream));
        }
        done = r->copy_stream;
      }
    }
    YV_CUDA(cudaEventRecord(sl.ev_done, done));
    sl.target = target;
    sl.ticket = r->next_ticket;
    *ticket = (int)(r->next_ticket & 0x7fffffff);
    ++r->next_ticket;
    return YV_OK;
  });
}

int yv_wait_frame(yv_renderer *r, int ticket, const uint8_t **rgba) {
  if (!r) return fail(YV_ERR_ARG, "null renderer");
  if (rgba) *rgba = nullptr;
  for (int s = 0; s < r->opt_slots; ++s) {
    yv_frame_slot &sl = r->slots[s];
    if (sl.ticket < 0 || (int)(sl.ticket & 0x7fffffff) != ticket) continue;
    YV_CUDA(cudaSetDevice(r->device));
    YV_CUDA(cudaEventSynchronize(sl.ev_done));
    float ms = -1.0f;
    if (cudaEventElapsedTime(&ms, sl.ev_begin, sl.ev_done) != cudaSuccess) { cudaGetLastError(); ms = -1.0f; }
    r->last_ms = ms; r->timed = true; r->launches = sl.launches;
    if (rgba) *rgba = (const uint8_t *)sl.target;
    sl.ticket = -1;
    return YV_OK;
  }
  return fail(YV_ERR_ARG, "unknown frame ticket");
}

}  // extern "C"
