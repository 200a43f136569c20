// svo_host.cpp — .vox I/O, validation and the top-down procedural SVO builders.
//
// The reference's builder (cpp/DynamicSVO, cpp/builders.h) is not in the snapshot; what is
// documented is its contract: VoxelSource::TryRange classifies a cube as empty / full /
// surface voxel / needs subdivision (reaction/report/main.tex:88-94) and DynamicSVO merges
// that into the tree (GROW = union, ore/src/main.cpp:101-103). This file implements that
// contract as a one-pass parallel top-down build (no incremental editing): a Source
// classifies cubes, the builder recurses on Mixed cubes, collapses uniform octets and emits
// reference-layout VoxNode records (main.tex:38-55) that the .vox writer stores the way
// SVOData::Load reads them (cell/svodata.h:31-50).
#include "svo_host.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>

namespace yv {

// ---------------------------------------------------------------------------------------------
// .vox container
// ---------------------------------------------------------------------------------------------

int load_vox(const char *path, HostSVO &out, std::string &err) {
  FILE *f = std::fopen(path, "rb");
  if (!f) { err = std::string("cannot open ") + path; return -1; }
  uint32_t hdr[4];
  if (std::fread(hdr, sizeof(uint32_t), 4, f) != 4) { std::fclose(f); err = "short .vox header"; return -2; }
  out.root = hdr[0];
  out.depth = (hdr[1] == YV_VOX_MAGIC) ? hdr[2] : 0;   // the reference discards words 1,2 (svodata.h:38-39)
  const uint32_t count = hdr[3];
  // the header word is untrusted: check it against what the file holds before allocating count * 40 bytes
  {
    const long here = std::ftell(f);
    long size = -1;
    if (here >= 0 && std::fseek(f, 0, SEEK_END) == 0) { size = std::ftell(f); std::fseek(f, here, SEEK_SET); }
    if (size >= 0 && (uint64_t)count * sizeof(yv_vox_node) > (uint64_t)(size - here)) {
      std::fclose(f); err = "truncated .vox node array"; out.nodes.clear(); return -3;
    }
  }
  out.nodes.resize(count);
  size_t got = count ? std::fread(out.nodes.data(), sizeof(yv_vox_node), count, f) : 0;
  std::fclose(f);
  if (got != count) { err = "truncated .vox node array"; out.nodes.clear(); return -3; }
  normalize_flags(out);
  return validate(out, err);
}

int save_vox(const char *path, const HostSVO &svo, std::string &err) {
  FILE *f = std::fopen(path, "wb");
  if (!f) { err = std::string("cannot create ") + path; return -1; }
  uint32_t hdr[4] = { svo.root, YV_VOX_MAGIC, svo.depth, (uint32_t)svo.nodes.size() };
  bool ok = std::fwrite(hdr, sizeof(uint32_t), 4, f) == 4;
  if (ok && !svo.nodes.empty())
    ok = std::fwrite(svo.nodes.data(), sizeof(yv_vox_node), svo.nodes.size(), f) == svo.nodes.size();
  ok = (std::fclose(f) == 0) && ok;
  if (!ok) { err = "write failed"; return -2; }
  return 0;
}

// The null flags (bits 8..15) are derived data; the raw-layout kernel trusts them, so pools that come from a file or
// from the caller get them recomputed from the child words (top bit set and not a leaf = null, main.tex:40-42,62).
void normalize_flags(HostSVO &svo) {
  for (yv_vox_node &nd : svo.nodes) {
    uint32_t nulls = 0;
    for (int c = 0; c < 8; ++c)
      if (!YV_LEAF_FLAG(nd.flags, c) && YV_IS_NULL(nd.child[c])) nulls |= 1u << (8 + c);
    nd.flags = (nd.flags & ~0xff00u) | nulls;
  }
}

int validate(const HostSVO &svo, std::string &err) {
  const uint64_t n = svo.nodes.size();
  if (n >= 0x80000000ull) { err = "node pool exceeds 2^31 entries"; return -10; }
  if (!YV_IS_NULL(svo.root) && svo.root >= n) { err = "root id outside node pool"; return -11; }
  for (uint64_t i = 0; i < n; ++i) {
    const yv_vox_node &nd = svo.nodes[i];
    for (int c = 0; c < 8; ++c) {
      if (YV_LEAF_FLAG(nd.flags, c)) continue;
      const uint32_t id = nd.child[c];
      if (!YV_IS_NULL(id) && id >= n) {
        err = "child id outside node pool at node " + std::to_string(i);
        return -12;
      }
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// VoxData packing (spec: include/yv_format.h)
// ---------------------------------------------------------------------------------------------

uint32_t pack_voxdata(uint8_t r, uint8_t g, uint8_t b, float nx, float ny, float nz) {
  double l1 = std::fabs((double)nx) + std::fabs((double)ny) + std::fabs((double)nz);
  double px = 0.0, py = 0.0;
  if (l1 > 0.0) { px = nx / l1; py = ny / l1; }
  if (nz < 0.0f) {
    double ox = (1.0 - std::fabs(py)) * (px >= 0.0 ? 1.0 : -1.0);
    double oy = (1.0 - std::fabs(px)) * (py >= 0.0 ? 1.0 : -1.0);
    px = ox; py = oy;
  }
  long u = std::lround((px * 0.5 + 0.5) * 255.0);
  long v = std::lround((py * 0.5 + 0.5) * 255.0);
  u = std::min(255l, std::max(0l, u));
  v = std::min(255l, std::max(0l, v));
  return YV_PACK_RGB565((uint32_t)r, (uint32_t)g, (uint32_t)b) | ((uint32_t)u << 16) | ((uint32_t)v << 24);
}

namespace {

// decode used only to average sub-tree attributes into VoxNode::data (main.tex:59)
struct Accum {
  double r = 0, g = 0, b = 0, nx = 0, ny = 0, nz = 0;
  uint64_t count = 0;
  void add(const Accum &o) { r += o.r; g += o.g; b += o.b; nx += o.nx; ny += o.ny; nz += o.nz; count += o.count; }
};

void accum_voxel(Accum &a, uint32_t d) {
  uint32_t r5 = (d >> 11) & 31u, g6 = (d >> 5) & 63u, b5 = d & 31u;
  a.r += (r5 << 3) | (r5 >> 2); a.g += (g6 << 2) | (g6 >> 4); a.b += (b5 << 3) | (b5 >> 2);
  double fx = ((d >> 16) & 255u) / 127.5 - 1.0, fy = ((d >> 24) & 255u) / 127.5 - 1.0;
  double fz = 1.0 - std::fabs(fx) - std::fabs(fy);
  if (fz < 0) {
    double ox = (1.0 - std::fabs(fy)) * (fx >= 0 ? 1.0 : -1.0), oy = (1.0 - std::fabs(fx)) * (fy >= 0 ? 1.0 : -1.0);
    fx = ox; fy = oy;
  }
  double l = std::sqrt(fx * fx + fy * fy + fz * fz);
  a.nx += fx / l; a.ny += fy / l; a.nz += fz / l;
  a.count += 1;
}

uint32_t accum_pack(const Accum &a) {
  if (!a.count) return 0;
  double inv = 1.0 / (double)a.count;
  auto c8 = [&](double v) { long q = std::lround(v * inv); return (uint8_t)std::min(255l, std::max(0l, q)); };
  return pack_voxdata(c8(a.r), c8(a.g), c8(a.b), (float)a.nx, (float)a.ny, (float)a.nz);
}

// result of building one cube
struct Built {
  RangeClass cls = RangeClass::Empty;   // Mixed here means "a real node", id in `ref`
  uint32_t ref = YV_EMPTY_NODE;         // node id (Mixed) or VoxData (Voxel)
  Accum acc;
};

// Generic top-down builder over a Source with per-cube context:
//   struct Source { using Ctx = ...; Ctx root_ctx() const;
//                   RangeClass classify(const Ctx &parent, int x, int y, int z, int size,
//                                       Ctx &ctx, uint32_t &voxdata) const; };
template <class Source>
class TopDownBuilder {
 public:
  TopDownBuilder(const Source &src, int depth, int threads)
      : src_(src), depth_(depth), threads_(std::max(1, threads)) {}

  void run(HostSVO &out) {
    using Ctx = typename Source::Ctx;
    const int split = std::min(depth_, depth_ >= 9 ? 4 : (depth_ >= 6 ? 2 : 0));
    const int full = 1 << depth_;

    // phase 1: serial descent to the split level, collecting Mixed cubes as tasks
    struct Task { int x, y, z, size; Ctx ctx; std::vector<yv_vox_node> nodes; Built res; };
    std::vector<Task> tasks;
    struct Top { RangeClass cls; uint32_t vox; int task; int child[8]; };   // top-part tree
    std::vector<Top> tops;

    std::function<int(const Ctx &, int, int, int, int, int)> descend =
        [&](const Ctx &pctx, int x, int y, int z, int size, int level) -> int {
      Ctx ctx; uint32_t vox = 0;
      RangeClass cls = src_.classify(pctx, x, y, z, size, ctx, vox);
      int me = (int)tops.size();
      tops.push_back(Top{cls, vox, -1, {-1, -1, -1, -1, -1, -1, -1, -1}});
      if (cls != RangeClass::Mixed) return me;
      if (level == split) {
        tops[me].task = (int)tasks.size();
        tasks.push_back(Task{x, y, z, size, ctx, {}, {}});
        return me;
      }
      int h = size / 2;
      for (int c = 0; c < 8; ++c) {
        int id = descend(ctx, x + ((c & 1) ? h : 0), y + ((c & 2) ? h : 0), z + ((c & 4) ? h : 0), h, level + 1);
        tops[me].child[c] = id;
      }
      return me;
    };
    int root_top = descend(src_.root_ctx(), 0, 0, 0, full, 0);

    // phase 2: build every task's subtree into its own pool (ids local to the task)
    std::atomic<size_t> next{0};
    auto worker = [&]() {
      for (;;) {
        size_t i = next.fetch_add(1);
        if (i >= tasks.size()) break;
        Task &t = tasks[i];
        t.res = build_children(t.ctx, t.x, t.y, t.z, t.size, t.nodes);
      }
    };
    std::vector<std::thread> pool;
    int nthreads = (int)std::min<size_t>((size_t)threads_, std::max<size_t>(1, tasks.size()));
    for (int i = 1; i < nthreads; ++i) pool.emplace_back(worker);
    worker();
    for (auto &th : pool) th.join();

    // phase 3: concatenate task pools, rebasing the child ids
    std::vector<uint64_t> base(tasks.size() + 1, 0);
    for (size_t i = 0; i < tasks.size(); ++i) base[i + 1] = base[i] + tasks[i].nodes.size();
    out.nodes.clear();
    out.nodes.resize(base.back());
    next = 0;
    auto copier = [&]() {
      for (;;) {
        size_t i = next.fetch_add(1);
        if (i >= tasks.size()) break;
        Task &t = tasks[i];
        const uint32_t b = (uint32_t)base[i];
        yv_vox_node *dst = out.nodes.data() + base[i];
        for (size_t k = 0; k < t.nodes.size(); ++k) {
          yv_vox_node nd = t.nodes[k];
          for (int c = 0; c < 8; ++c)
            if (!YV_LEAF_FLAG(nd.flags, c) && !YV_IS_NULL(nd.child[c])) nd.child[c] += b;
          dst[k] = nd;
        }
        if (t.res.cls == RangeClass::Mixed) t.res.ref += b;
        std::vector<yv_vox_node>().swap(t.nodes);
      }
    };
    pool.clear();
    for (int i = 1; i < nthreads; ++i) pool.emplace_back(copier);
    copier();
    for (auto &th : pool) th.join();

    // phase 4: assemble the top part bottom-up
    std::function<Built(int)> assemble = [&](int ti) -> Built {
      const Top &tp = tops[ti];
      Built b;
      if (tp.cls == RangeClass::Empty) { b.cls = RangeClass::Empty; b.ref = YV_EMPTY_NODE; return b; }
      if (tp.cls == RangeClass::Full) { b.cls = RangeClass::Full; b.ref = YV_FULL_NODE; return b; }
      if (tp.cls == RangeClass::Voxel) { b.cls = RangeClass::Voxel; b.ref = tp.vox; accum_voxel(b.acc, tp.vox); return b; }
      if (tp.task >= 0) return tasks[tp.task].res;
      Built kids[8];
      for (int c = 0; c < 8; ++c) kids[c] = assemble(tp.child[c]);
      return emit(kids, out.nodes);
    };
    Built root = assemble(root_top);
    out.depth = (uint32_t)depth_;
    if (root.cls == RangeClass::Mixed) out.root = root.ref;
    else if (root.cls == RangeClass::Voxel || root.cls == RangeClass::Full) {
      // a uniformly solid scene: represent as a node whose 8 children are Full (invisible), like the
      // reference where FullNode is null to the tracer
      out.root = YV_FULL_NODE;
    } else out.root = YV_EMPTY_NODE;
  }

 private:
  // Emit one node from 8 built children, or collapse when they are uniformly empty / full.
  static Built emit(const Built kids[8], std::vector<yv_vox_node> &pool) {
    int n_empty = 0, n_full = 0;
    for (int c = 0; c < 8; ++c) { n_empty += kids[c].cls == RangeClass::Empty; n_full += kids[c].cls == RangeClass::Full; }
    Built b;
    if (n_empty == 8) { b.cls = RangeClass::Empty; b.ref = YV_EMPTY_NODE; return b; }
    if (n_full == 8) { b.cls = RangeClass::Full; b.ref = YV_FULL_NODE; return b; }
    yv_vox_node nd;
    std::memset(&nd, 0, sizeof nd);
    for (int c = 0; c < 8; ++c) {
      const Built &k = kids[c];
      switch (k.cls) {
        case RangeClass::Empty: nd.child[c] = YV_EMPTY_NODE; nd.flags |= 1u << (8 + c); break;
        case RangeClass::Full:  nd.child[c] = YV_FULL_NODE;  nd.flags |= 1u << (8 + c); break;
        case RangeClass::Voxel: nd.child[c] = k.ref; nd.flags |= 1u << c; break;
        case RangeClass::Mixed: nd.child[c] = k.ref; break;
      }
      b.acc.add(k.acc);
    }
    nd.data = accum_pack(b.acc);
    b.cls = RangeClass::Mixed;
    b.ref = (uint32_t)pool.size();
    pool.push_back(nd);
    return b;
  }

  // Build the 8 children of a cube already known to be Mixed, then emit its node.
  Built build_children(const typename Source::Ctx &ctx, int x, int y, int z, int size,
                       std::vector<yv_vox_node> &pool) const {
    Built kids[8];
    const int h = size / 2;
    for (int c = 0; c < 8; ++c) {
      const int cx = x + ((c & 1) ? h : 0), cy = y + ((c & 2) ? h : 0), cz = z + ((c & 4) ? h : 0);
      typename Source::Ctx cctx; uint32_t vox = 0;
      RangeClass cls = src_.classify(ctx, cx, cy, cz, h, cctx, vox);
      Built &k = kids[c];
      k.cls = cls;
      if (cls == RangeClass::Empty) k.ref = YV_EMPTY_NODE;
      else if (cls == RangeClass::Full) k.ref = YV_FULL_NODE;
      else if (cls == RangeClass::Voxel) { k.ref = vox; accum_voxel(k.acc, vox); }
      else {
        if (h == 1) { k.cls = RangeClass::Empty; k.ref = YV_EMPTY_NODE; }   // sources must resolve unit cubes
        else k = build_children(cctx, cx, cy, cz, h, pool);
      }
    }
    return emit(kids, pool);
  }

  const Source &src_;
  int depth_;
  int threads_;
};

// ---------------------------------------------------------------------------------------------
// Sphere sources
// ---------------------------------------------------------------------------------------------

struct Sphere { int64_t cx, cy, cz, r; uint8_t col[3]; };

// Union of solid spheres. A unit cube cut by some sphere's surface and not swallowed by another
// sphere is a surface voxel; its attributes come from the last such sphere in build order
// (the order in which gen_spheres.py issues BuildRange calls, GROW mode).
class SpheresSource {
 public:
  using Ctx = std::vector<uint32_t>;   // candidate sphere indices for this cube
  explicit SpheresSource(std::vector<Sphere> s) : spheres_(std::move(s)) {}
  Ctx root_ctx() const {
    Ctx c(spheres_.size());
    for (size_t i = 0; i < c.size(); ++i) c[i] = (uint32_t)i;
    return c;
  }
  RangeClass classify(const Ctx &parent, int x, int y, int z, int size, Ctx &ctx, uint32_t &vox) const {
    ctx.clear();
    for (uint32_t si : parent) {
      const Sphere &s = spheres_[si];
      // squared distance from the centre to the nearest / farthest point of the cube
      auto axis = [&](int64_t lo, int64_t c, int64_t &dmin, int64_t &dmax) {
        int64_t hi = lo + size;
        int64_t dn = c < lo ? lo - c : (c > hi ? c - hi : 0);
        int64_t df = std::max(std::llabs(c - lo), std::llabs(hi - c));
        dmin += dn * dn; dmax += df * df;
      };
      int64_t dmin = 0, dmax = 0;
      axis(x, s.cx, dmin, dmax); axis(y, s.cy, dmin, dmax); axis(z, s.cz, dmin, dmax);
      const int64_t r2 = s.r * s.r;
      if (dmax < r2) return RangeClass::Full;
      if (dmin <= r2) ctx.push_back(si);
    }
    if (ctx.empty()) return RangeClass::Empty;
    if (size > 1) return RangeClass::Mixed;
    const Sphere &s = spheres_[ctx.back()];
    vox = pack_voxdata(s.col[0], s.col[1], s.col[2],
                       (float)(x + 0.5 - (double)s.cx), (float)(y + 0.5 - (double)s.cy), (float)(z + 0.5 - (double)s.cz));
    return RangeClass::Voxel;
  }
 private:
  std::vector<Sphere> spheres_;
};

// Dense grid source (tests): VoxData != 0 is a surface voxel; there is no "full" class.
class DenseSource {
 public:
  struct Ctx {};
  DenseSource(int depth, const uint32_t *vox) : n_(1 << depth), vox_(vox) {
    // occupancy mip pyramid: level l holds (n>>l)^3 "any voxel set" bits
    int m = n_;
    const uint32_t *src = vox;
    std::vector<uint8_t> lvl((size_t)m * m * m);
    for (size_t i = 0; i < lvl.size(); ++i) lvl[i] = src[i] != 0;
    mips_.push_back(lvl);
    while (m > 1) {
      int h = m / 2;
      std::vector<uint8_t> up((size_t)h * h * h, 0);
      const std::vector<uint8_t> &lo = mips_.back();
      for (int z = 0; z < m; ++z) for (int y = 0; y < m; ++y) for (int x = 0; x < m; ++x)
        if (lo[((size_t)z * m + y) * m + x]) up[((size_t)(z / 2) * h + y / 2) * h + x / 2] = 1;
      mips_.push_back(up);
      m = h;
    }
  }
  Ctx root_ctx() const { return {}; }
  RangeClass classify(const Ctx &, int x, int y, int z, int size, Ctx &, uint32_t &vox) const {
    int l = 0; while ((1 << l) < size) ++l;
    int m = n_ >> l;
    if (!mips_[l][((size_t)(z >> l) * m + (y >> l)) * m + (x >> l)]) return RangeClass::Empty;
    if (size > 1) return RangeClass::Mixed;
    vox = vox_[((size_t)z * n_ + y) * n_ + x];
    return RangeClass::Voxel;
  }
 private:
  int n_;
  const uint32_t *vox_;
  std::vector<std::vector<uint8_t>> mips_;
};

// ---------------------------------------------------------------------------------------------
// Iso-volume source: sum of trilinear value-noise octaves + a height term
// ---------------------------------------------------------------------------------------------

inline uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// Within one lattice cell of every octave the field is trilinear, so its extrema over any
// aligned sub-cube are attained at the sub-cube's corners: corner evaluation gives exact bounds.
class IsoSource {
 public:
  // Per-cube context: the field at the cube's 8 corners once the cube is small enough for corner evaluation to be exact,
  // and — filled in when its first child is classified — at the 27 points of its 3x3x3 half-size lattice, so that the
  // eight children share their corners (19 new field evaluations per cube instead of 64). The field is a pure function
  // of the integer lattice point, so the tree is the one the unshared evaluation built.
  struct Ctx {
    int x = 0, y = 0, z = 0, size = 0;
    bool have8 = false;
    mutable bool have27 = false;
    float c8[8];
    mutable float g27[27];
  };
  IsoSource(int depth, uint32_t seed, int iso) : depth_(depth), n_(1 << depth), iso_((float)iso / 255.0f) {
    zmax_ = (n_ * 5) / 16;                       // slab: 640/2048 of the cube (gen_largevol.py:8-9: 5 of 16 brick layers)
    if (zmax_ < 2) zmax_ = std::min(n_, 2);
    const int cells[3] = { 8, 32, 128 };
    const float amps[3] = { 0.55f, 0.30f, 0.15f };
    for (int o = 0; o < 3; ++o) {
      Octave oc;
      oc.cells = std::min(cells[o], n_);          // lattice cells across the cube (power of two)
      oc.amp = amps[o];
      const int m = oc.cells + 1;
      oc.lat.resize((size_t)m * m * m);
      for (int k = 0; k < m; ++k) for (int j = 0; j < m; ++j) for (int i = 0; i < m; ++i) {
        uint32_t h = mix32(seed * 0x9e3779b9U + (uint32_t)o);
        h = mix32(h ^ (uint32_t)i); h = mix32(h ^ ((uint32_t)j * 0x85ebca6bU)); h = mix32(h ^ ((uint32_t)k * 0xc2b2ae35U));
        oc.lat[((size_t)k * m + j) * m + i] = (float)(h >> 8) * (1.0f / 16777216.0f);
      }
      // min/max pyramid over lattice cells (cell value range = range of its 8 corners)
      int c = oc.cells;
      std::vector<float> mn((size_t)c * c * c), mx((size_t)c * c * c);
      for (int k = 0; k < c; ++k) for (int j = 0; j < c; ++j) for (int i = 0; i < c; ++i) {
        float lo = 1e30f, hi = -1e30f;
        for (int d = 0; d < 8; ++d) {
          float v = oc.lat[((size_t)(k + ((d >> 2) & 1)) * m + (j + ((d >> 1) & 1))) * m + (i + (d & 1))];
          lo = std::min(lo, v); hi = std::max(hi, v);
        }
        mn[((size_t)k * c + j) * c + i] = lo; mx[((size_t)k * c + j) * c + i] = hi;
      }
      oc.mn.push_back(mn); oc.mx.push_back(mx);
      while (c > 1) {
        int h = c / 2;
        std::vector<float> mn2((size_t)h * h * h, 1e30f), mx2((size_t)h * h * h, -1e30f);
        const std::vector<float> &pmn = oc.mn.back(), &pmx = oc.mx.back();
        for (int k = 0; k < c; ++k) for (int j = 0; j < c; ++j) for (int i = 0; i < c; ++i) {
          size_t d = ((size_t)(k / 2) * h + j / 2) * h + i / 2, s = ((size_t)k * c + j) * c + i;
          mn2[d] = std::min(mn2[d], pmn[s]); mx2[d] = std::max(mx2[d], pmx[s]);
        }
        oc.mn.push_back(mn2); oc.mx.push_back(mx2);
        c = h;
      }
      oct_.push_back(std::move(oc));
    }
  }
  Ctx root_ctx() const { return {}; }

  float field(double x, double y, double z) const {
    double f = 0.0;
    for (const Octave &oc : oct_) {
      const double s = (double)oc.cells / (double)n_;
      double fx = x * s, fy = y * s, fz = z * s;
      int i = std::min((int)fx, oc.cells - 1), j = std::min((int)fy, oc.cells - 1), k = std::min((int)fz, oc.cells - 1);
      double u = fx - i, v = fy - j, w = fz - k;
      const int m = oc.cells + 1;
      auto L = [&](int a, int b, int c) { return (double)oc.lat[((size_t)(k + c) * m + (j + b)) * m + (i + a)]; };
      double c00 = L(0,0,0) + (L(1,0,0) - L(0,0,0)) * u, c10 = L(0,1,0) + (L(1,1,0) - L(0,1,0)) * u;
      double c01 = L(0,0,1) + (L(1,0,1) - L(0,0,1)) * u, c11 = L(0,1,1) + (L(1,1,1) - L(0,1,1)) * u;
      double c0 = c00 + (c10 - c00) * v, c1 = c01 + (c11 - c01) * v;
      f += oc.amp * (c0 + (c1 - c0) * w);
    }
    // height term: solid near the slab floor, open near its ceiling
    f = kNoiseW * f + kHeightW * (1.0 - z / (double)zmax_) + kBias;
    return (float)f;
  }

  RangeClass classify(const Ctx &parent, int x, int y, int z, int size, Ctx &ctx, uint32_t &vox) const {
    ctx.x = x; ctx.y = y; ctx.z = z; ctx.size = size; ctx.have8 = false; ctx.have27 = false;
    if (z >= zmax_) return RangeClass::Empty;
    float lo, hi;
    bool exact = true;
    for (const Octave &oc : oct_) if (size > n_ / oc.cells) exact = false;
    if (exact && z + size <= zmax_) {
      lo = 1e30f; hi = -1e30f;
      if (parent.have8 && parent.size == 2 * size) {
        if (!parent.have27) {                                    // the parent's 3x3x3 lattice, its own corners reused
          for (int k = 0; k < 3; ++k) for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) {
            float &g = parent.g27[(k * 3 + j) * 3 + i];
            if (i != 1 && j != 1 && k != 1) g = parent.c8[(i >> 1) | ((j >> 1) << 1) | ((k >> 1) << 2)];
            else g = field(parent.x + i * size, parent.y + j * size, parent.z + k * size);
          }
          parent.have27 = true;
        }
        const int ox = (x - parent.x) / size, oy = (y - parent.y) / size, oz = (z - parent.z) / size;
        for (int d = 0; d < 8; ++d)
          ctx.c8[d] = parent.g27[((oz + ((d >> 2) & 1)) * 3 + (oy + ((d >> 1) & 1))) * 3 + (ox + (d & 1))];
      } else {
        for (int d = 0; d < 8; ++d)
          ctx.c8[d] = field(x + ((d & 1) ? size : 0), y + ((d & 2) ? size : 0), z + ((d & 4) ? size : 0));
      }
      ctx.have8 = true;
      for (int d = 0; d < 8; ++d) { lo = std::min(lo, ctx.c8[d]); hi = std::max(hi, ctx.c8[d]); }
    } else {
      double nlo = 0, nhi = 0;
      for (const Octave &oc : oct_) {
        int cell = n_ / oc.cells;                 // voxels per lattice cell
        if (size >= cell) {
          int l = 0; while ((cell << l) < size) ++l;
          int c = oc.cells >> l;
          size_t idx = ((size_t)(z / size) * c + (y / size)) * c + (x / size);
          nlo += oc.amp * oc.mn[l][idx]; nhi += oc.amp * oc.mx[l][idx];
        } else {
          int c = oc.cells;
          size_t idx = ((size_t)(z / cell) * c + (y / cell)) * c + (x / cell);
          nlo += oc.amp * oc.mn[0][idx]; nhi += oc.amp * oc.mx[0][idx];
        }
      }
      int z1 = std::min(z + size, zmax_);
      lo = (float)(kNoiseW * nlo + kHeightW * (1.0 - (double)z1 / zmax_) + kBias) - 1e-4f;
      hi = (float)(kNoiseW * nhi + kHeightW * (1.0 - (double)z / zmax_) + kBias) + 1e-4f;
    }
    if (hi < iso_) return RangeClass::Empty;
    if (lo >= iso_ && z + size <= zmax_) return RangeClass::Full;
    if (size > 1) return RangeClass::Mixed;
    // surface voxel: the iso-surface crosses this unit cell. Normal = -gradient (central differences).
    double cx = x + 0.5, cy = y + 0.5, cz = z + 0.5;
    float gx = field(cx + 0.5, cy, cz) - field(cx - 0.5, cy, cz);
    float gy = field(cx, cy + 0.5, cz) - field(cx, cy - 0.5, cz);
    float gz = field(cx, cy, cz + 0.5) - field(cx, cy, cz - 0.5);
    float gl = std::sqrt(gx * gx + gy * gy + gz * gz);
    if (!(gl > 0)) { gx = 0; gy = 0; gz = -1; }
    // colour: height-banded rock / grass / snow palette perturbed by the lattice hash
    double hrel = cz / (double)zmax_;
    uint32_t hsh = mix32((uint32_t)(x >> 3) * 73856093u ^ (uint32_t)(y >> 3) * 19349663u ^ (uint32_t)(z >> 3) * 83492791u);
    int jit = (int)(hsh & 31u) - 16;
    int r, g, b;
    if (hrel < 0.35) { r = 120; g = 100; b = 80; } else if (hrel < 0.6) { r = 80; g = 150; b = 70; } else { r = 220; g = 220; b = 230; }
    auto cl = [](int v) { return (uint8_t)std::min(255, std::max(0, v)); };
    vox = pack_voxdata(cl(r + jit), cl(g + jit), cl(b + jit), -gx, -gy, -gz);
    return RangeClass::Voxel;
  }

 private:
  struct Octave { int cells; float amp; std::vector<float> lat; std::vector<std::vector<float>> mn, mx; };
  static constexpr double kNoiseW = 0.55, kHeightW = 0.45, kBias = 0.284;
  int depth_, n_, zmax_;
  float iso_;
  std::vector<Octave> oct_;
};

int check_depth(int depth, std::string &err) {
  if (depth < 1 || depth > 16) { err = "depth must be in 1..16"; return -1; }
  return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// public builders
// ---------------------------------------------------------------------------------------------

int build_sphere_fractal(int depth, int threads, HostSVO &out, std::string &err) {
  if (check_depth(depth, err)) return -1;
  if (depth < 4) { err = "sphere fractal needs depth >= 4"; return -1; }
  // gen_spheres.py:8-32, with the level-11 constants (1024, 256) scaled by 2^(depth-11)
  const int LevelNum = 8;
  const double BaseRadius = std::ldexp(1.0, depth - 3);
  std::vector<Sphere> spheres;
  struct V { double x, y, z; };
  std::function<void(int, V, V, V, V)> rec = [&](int lev, V pos, V x, V y, V z) {
    if (lev > 4) {
      int64_t r = (int64_t)std::floor(BaseRadius / (double)(1 << lev));      // Py2 integer division (:13)
      Sphere s;
      s.cx = (int64_t)pos.x; s.cy = (int64_t)pos.y; s.cz = (int64_t)pos.z;   // p3i(pos) truncation (:18)
      s.r = r;
      s.col[0] = 128; s.col[1] = 128; s.col[2] = (uint8_t)(lev * 255 / LevelNum);   // (:13)
      if (r >= 1) spheres.push_back(s);
    }
    if (lev < LevelNum - 1) {
      V x1{ x.x / 2, x.y / 2, x.z / 2 }, y1{ y.x / 2, y.y / 2, y.z / 2 }, z1{ z.x / 2, z.y / 2, z.z / 2 };
      auto add = [](V a, V b) { return V{ a.x + b.x, a.y + b.y, a.z + b.z }; };
      auto sub = [](V a, V b) { return V{ a.x - b.x, a.y - b.y, a.z - b.z }; };
      auto neg = [](V a) { return V{ -a.x, -a.y, -a.z }; };
      rec(lev + 1, add(pos, x), y1, z1, x1);           // (:23)
      rec(lev + 1, sub(pos, x), y1, z1, neg(x1));      // (:24)
      rec(lev + 1, add(pos, y), x1, z1, y1);           // (:26)
      rec(lev + 1, sub(pos, y), x1, z1, neg(y1));      // (:27)
      rec(lev + 1, add(pos, z), x1, y1, z1);           // (:29)
    }
  };
  const double c = std::ldexp(1.0, depth - 1);
  rec(0, V{ c, c, c }, V{ BaseRadius * 1.5, 0, 0 }, V{ 0, BaseRadius * 1.5, 0 }, V{ 0, 0, BaseRadius * 1.5 });   // (:32)
  SpheresSource src(std::move(spheres));
  TopDownBuilder<SpheresSource>(src, depth, threads).run(out);
  return validate(out, err);
}

int build_single_sphere(int depth, int cx, int cy, int cz, int radius,
                        uint8_t r, uint8_t g, uint8_t b, HostSVO &out, std::string &err) {
  if (check_depth(depth, err)) return -1;
  Sphere s; s.cx = cx; s.cy = cy; s.cz = cz; s.r = radius; s.col[0] = r; s.col[1] = g; s.col[2] = b;
  SpheresSource src(std::vector<Sphere>{ s });
  TopDownBuilder<SpheresSource>(src, depth, 1).run(out);
  return validate(out, err);
}

int build_from_dense(int depth, const uint32_t *vox, HostSVO &out, std::string &err) {
  if (check_depth(depth, err)) return -1;
  if (depth > 8) { err = "dense build limited to depth <= 8"; return -1; }
  DenseSource src(depth, vox);
  TopDownBuilder<DenseSource>(src, depth, 1).run(out);
  return validate(out, err);
}

int build_iso_volume(int depth, uint32_t seed, int iso, int threads, HostSVO &out, std::string &err) {
  if (check_depth(depth, err)) return -1;
  if (depth < 4) { err = "iso volume needs depth >= 4"; return -1; }
  IsoSource src(depth, seed, iso);
  TopDownBuilder<IsoSource>(src, depth, threads).run(out);
  return validate(out, err);
}

}  // namespace yv
