// dynamic_svo.cpp — see dynamic_svo.h
#include "dynamic_svo.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace yv {

// ---- sources ------------------------------------------------------------------------------------

SphereSource::SphereSource(int radius, uint8_t r, uint8_t g, uint8_t b, bool inverted)
    : radius_(std::max(0, radius)), inverted_(inverted) { col_[0] = r; col_[1] = g; col_[2] = b; }
void SphereSource::GetSize(int s[3]) const { s[0] = s[1] = s[2] = 2 * radius_ + 1; }
void SphereSource::GetPivot(int p[3]) const { p[0] = p[1] = p[2] = radius_; }

// same rule as the batch builder's sphere (svo_host.cpp): centre on the lattice point `radius` (the pivot),
// a unit cube cut by the surface is a surface voxel
RangeClass SphereSource::TryRange(const int p[3], int size, uint32_t &vox) const {
  int64_t dmin = 0, dmax = 0;
  for (int a = 0; a < 3; ++a) {
    const int64_t lo = p[a], hi = (int64_t)p[a] + size, c = radius_;
    const int64_t dn = c < lo ? lo - c : (c > hi ? c - hi : 0);
    const int64_t df = std::max(std::llabs(c - lo), std::llabs(hi - c));
    dmin += dn * dn; dmax += df * df;
  }
  const int64_t r2 = (int64_t)radius_ * radius_;
  if (dmax < r2) return RangeClass::Full;
  if (dmin > r2) return RangeClass::Empty;
  if (size > 1) return RangeClass::Mixed;
  float n[3] = { (float)(p[0] + 0.5 - radius_), (float)(p[1] + 0.5 - radius_), (float)(p[2] + 0.5 - radius_) };
  if (inverted_) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
  vox = pack_voxdata(col_[0], col_[1], col_[2], n[0], n[1], n[2]);
  return RangeClass::Voxel;
}

RawSource::RawSource(const int size[3], const uint32_t *v) {
  for (int a = 0; a < 3; ++a) size_[a] = std::max(1, size[a]);
  vox_.assign(v, v + (size_t)size_[0] * size_[1] * size_[2]);
  kind_.resize(vox_.size());
  for (size_t i = 0; i < vox_.size(); ++i) kind_[i] = vox_[i] ? 2 : 0;
}
RawSource::RawSource(const int size[3], const uint8_t *c, const int8_t *n) {
  for (int a = 0; a < 3; ++a) size_[a] = std::max(1, size[a]);
  const size_t count = (size_t)size_[0] * size_[1] * size_[2];
  vox_.assign(count, 0u);
  kind_.assign(count, 0);
  for (size_t i = 0; i < count; ++i) {
    const uint8_t alpha = c[4 * i + 3];
    if (alpha == 0) continue;
    if (alpha != 255) { kind_[i] = 1; continue; }
    kind_[i] = 2;
    vox_[i] = pack_voxdata(c[4 * i], c[4 * i + 1], c[4 * i + 2], (float)n[4 * i], (float)n[4 * i + 1], (float)n[4 * i + 2]);
  }
}
void RawSource::GetSize(int s[3]) const { for (int a = 0; a < 3; ++a) s[a] = size_[a]; }
void RawSource::GetPivot(int p[3]) const { p[0] = p[1] = p[2] = 0; }
RangeClass RawSource::TryRange(const int p[3], int size, uint32_t &vox) const {
  int lo[3], hi[3];
  bool clipped = false;
  for (int a = 0; a < 3; ++a) {
    lo[a] = std::max(0, p[a]); hi[a] = std::min(size_[a], p[a] + size);
    if (lo[a] >= hi[a]) return RangeClass::Empty;
    clipped = clipped || lo[a] != p[a] || hi[a] != p[a] + size;
  }
  if (size == 1) {
    const size_t i = ((size_t)lo[2] * size_[1] + lo[1]) * size_[0] + lo[0];
    vox = vox_[i];
    return kind_[i] == 2 ? RangeClass::Voxel : (kind_[i] == 1 ? RangeClass::Full : RangeClass::Empty);
  }
  bool any = false, all_buried = !clipped;      // a cube that sticks out of the brick is never Full
  for (int z = lo[2]; z < hi[2]; ++z)
    for (int y = lo[1]; y < hi[1]; ++y)
      for (int x = lo[0]; x < hi[0]; ++x) {
        const uint8_t k = kind_[((size_t)z * size_[1] + y) * size_[0] + x];
        any = any || k != 0;
        all_buried = all_buried && k == 1;
        if (any && !all_buried) return RangeClass::Mixed;
      }
  return any ? RangeClass::Full : RangeClass::Empty;
}

IsoBrickSource::IsoBrickSource(const int size[3], const uint8_t *d) {
  for (int a = 0; a < 3; ++a) size_[a] = std::max(1, size[a]);
  data_.assign(d, d + (size_t)size_[0] * size_[1] * size_[2]);
}
void IsoBrickSource::GetSize(int s[3]) const { for (int a = 0; a < 3; ++a) s[a] = size_[a]; }
void IsoBrickSource::GetPivot(int p[3]) const { p[0] = p[1] = p[2] = 0; }
int IsoBrickSource::at(int x, int y, int z) const {
  x = std::min(std::max(x, 0), size_[0] - 1); y = std::min(std::max(y, 0), size_[1] - 1); z = std::min(std::max(z, 0), size_[2] - 1);
  return data_[((size_t)z * size_[1] + y) * size_[0] + x];
}
// solid where the sample is >= iso (or < iso with SetInside(true)); a solid voxel with an open 6-neighbour
// is a surface voxel (normal = -gradient by central differences), a buried one is Full
RangeClass IsoBrickSource::TryRange(const int p[3], int size, uint32_t &vox) const {
  int lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    lo[a] = std::max(0, p[a]); hi[a] = std::min(size_[a], p[a] + size);
    if (lo[a] >= hi[a]) return RangeClass::Empty;
  }
  auto solid = [&](int x, int y, int z) {
    if (x < 0 || y < 0 || z < 0 || x >= size_[0] || y >= size_[1] || z >= size_[2]) return false;
    const bool s = at(x, y, z) >= iso_;
    return inside_ ? !s : s;
  };
  if (size == 1) {
    const int x = lo[0], y = lo[1], z = lo[2];
    if (!solid(x, y, z)) return RangeClass::Empty;
    const bool buried = solid(x - 1, y, z) && solid(x + 1, y, z) && solid(x, y - 1, z) && solid(x, y + 1, z) &&
                        solid(x, y, z - 1) && solid(x, y, z + 1);
    if (buried) return RangeClass::Full;
    float g[3] = { (float)(at(x + 1, y, z) - at(x - 1, y, z)), (float)(at(x, y + 1, z) - at(x, y - 1, z)),
                   (float)(at(x, y, z + 1) - at(x, y, z - 1)) };
    const float sgn = inside_ ? 1.0f : -1.0f;
    if (g[0] == 0 && g[1] == 0 && g[2] == 0) g[2] = -sgn;
    vox = pack_voxdata(col_[0], col_[1], col_[2], sgn * g[0], sgn * g[1], sgn * g[2]);
    return RangeClass::Voxel;
  }
  bool any = false, all = (hi[0] - lo[0] == size && hi[1] - lo[1] == size && hi[2] - lo[2] == size);
  for (int z = lo[2] - 1; z <= hi[2] && (all || !any); ++z)
    for (int y = lo[1] - 1; y <= hi[1] && (all || !any); ++y)
      for (int x = lo[0] - 1; x <= hi[0]; ++x) {
        const bool s = solid(x, y, z);
        const bool interior = x >= lo[0] && x < hi[0] && y >= lo[1] && y < hi[1] && z >= lo[2] && z < hi[2];
        if (interior && s) any = true;
        if (!s) all = false;          // an open sample in the cube or its 1-voxel shell -> not uniformly buried
        if (!all && any) break;
      }
  if (!any) return RangeClass::Empty;
  return all ? RangeClass::Full : RangeClass::Mixed;
}

// ---- pool bookkeeping ------------------------------------------------------------------------------

void DynamicSVO::touch(uint32_t id) {
  const size_t page = id / kPageNodes;
  if (page_version_.size() <= page) page_version_.resize(page + 1, 0u);
  page_version_[page] = version_;
}

void DynamicSVO::adopt_existing() {
  version_ = std::max(version_, 1u);
  page_version_.assign((svo_.nodes.size() + kPageNodes - 1) / kPageNodes, version_);
}

void DynamicSVO::reset_after_reload() {
  free_.clear();
  ++version_;
  page_version_.assign((svo_.nodes.size() + kPageNodes - 1) / kPageNodes, version_);
}

uint32_t DynamicSVO::alloc_node() {
  uint32_t id;
  if (!free_.empty()) { id = free_.back(); free_.pop_back(); }
  else { id = (uint32_t)svo_.nodes.size(); svo_.nodes.emplace_back(); }
  yv_vox_node &nd = svo_.nodes[id];
  std::memset(&nd, 0, sizeof nd);
  for (int c = 0; c < 8; ++c) nd.child[c] = YV_EMPTY_NODE;
  nd.flags = 0xffu << 8;
  touch(id);
  return id;
}

DynamicSVO::Ref DynamicSVO::child_ref(uint32_t id, int c) const {
  const yv_vox_node &nd = svo_.nodes[id];
  if (YV_LEAF_FLAG(nd.flags, c)) return Ref{ 2, nd.child[c] };
  if (nd.child[c] == YV_FULL_NODE) return Ref{ 1, 0 };
  if (YV_IS_NULL(nd.child[c])) return Ref{ 0, 0 };
  return Ref{ 3, nd.child[c] };
}

void DynamicSVO::set_child(uint32_t id, int c, Ref r) {
  yv_vox_node &nd = svo_.nodes[id];
  nd.flags &= ~((1u << c) | (1u << (8 + c)));
  switch (r.kind) {
    case 0: nd.child[c] = YV_EMPTY_NODE; nd.flags |= 1u << (8 + c); break;
    case 1: nd.child[c] = YV_FULL_NODE; nd.flags |= 1u << (8 + c); break;
    case 2: nd.child[c] = r.v; nd.flags |= 1u << c; break;
    default: nd.child[c] = r.v; break;
  }
}

void DynamicSVO::free_subtree(Ref r) {
  if (r.kind != 3) return;
  for (int c = 0; c < 8; ++c) free_subtree(child_ref(r.v, c));
  yv_vox_node &nd = svo_.nodes[r.v];
  std::memset(&nd, 0, sizeof nd);
  for (int c = 0; c < 8; ++c) nd.child[c] = YV_EMPTY_NODE;
  nd.flags = 0xffu << 8;
  touch(r.v);
  free_.push_back(r.v);
}

// VoxNode::data: mean colour / normal of the children's representatives (leaf voxels and child nodes' data)
uint32_t DynamicSVO::average_data(uint32_t id) const {
  double col[3] = { 0, 0, 0 }, nrm[3] = { 0, 0, 0 }; int n = 0;
  for (int c = 0; c < 8; ++c) {
    const Ref r = child_ref(id, c);
    uint32_t d;
    if (r.kind == 2) d = r.v; else if (r.kind == 3) d = svo_.nodes[r.v].data; else continue;
    const uint32_t r5 = (d >> 11) & 31u, g6 = (d >> 5) & 63u, b5 = d & 31u;
    col[0] += (r5 << 3) | (r5 >> 2); col[1] += (g6 << 2) | (g6 >> 4); col[2] += (b5 << 3) | (b5 >> 2);
    double fx = ((d >> 16) & 255u) / 127.5 - 1.0, fy = ((d >> 24) & 255u) / 127.5 - 1.0;
    double fz = 1.0 - std::fabs(fx) - std::fabs(fy);
    if (fz < 0) { const double ox = (1.0 - std::fabs(fy)) * (fx >= 0 ? 1 : -1), oy = (1.0 - std::fabs(fx)) * (fy >= 0 ? 1 : -1); fx = ox; fy = oy; }
    const double l = std::sqrt(fx * fx + fy * fy + fz * fz);
    nrm[0] += fx / l; nrm[1] += fy / l; nrm[2] += fz / l;
    ++n;
  }
  if (!n) return 0;
  auto c8 = [&](double v) { return (uint8_t)std::min(255l, std::max(0l, std::lround(v / n))); };
  return pack_voxdata(c8(col[0]), c8(col[1]), c8(col[2]), (float)nrm[0], (float)nrm[1], (float)nrm[2]);
}

// ---- BuildRange ---------------------------------------------------------------------------------------

DynamicSVO::Ref DynamicSVO::merge(Ref cur, int x, int y, int z, int size, BuildMode mode,
                                  const VoxelSource &src, const int org[3]) {
  const int p[3] = { x - org[0], y - org[1], z - org[2] };
  uint32_t vox = 0;
  const RangeClass cls = src.TryRange(p, size, vox);
  if (cls == RangeClass::Empty) return cur;                       // outside the source: untouched
  if (mode == BuildMode::Grow) {
    if (cur.kind == 1) return cur;                                // already solid
    if (cls == RangeClass::Full) { free_subtree(cur); return Ref{ 1, 0 }; }
    if (cls == RangeClass::Voxel) { free_subtree(cur); return Ref{ 2, vox }; }
  } else {
    if (cur.kind == 0) return cur;                                // nothing to carve
    if (cls == RangeClass::Full) { free_subtree(cur); return Ref{ 0, 0 }; }
    if (cls == RangeClass::Voxel) { free_subtree(cur); return Ref{ 2, vox }; }   // the cavity's wall
  }
  // Mixed: descend, expanding a uniform region into a node first
  if (size == 1) return cur;                                      // sources resolve unit cubes; defensive
  uint32_t id;
  if (cur.kind == 3) id = cur.v;
  else {
    id = alloc_node();
    if (cur.kind == 1) for (int c = 0; c < 8; ++c) set_child(id, c, Ref{ 1, 0 });
    // a coarser leaf voxel cannot be split meaningfully: its octants start out empty
  }
  const int h = size / 2;
  for (int c = 0; c < 8; ++c) {
    const Ref before = child_ref(id, c);
    const Ref after = merge(before, x + ((c & 1) ? h : 0), y + ((c & 2) ? h : 0), z + ((c & 4) ? h : 0), h, mode, src, org);
    if (after.kind != before.kind || after.v != before.v) { set_child(id, c, after); touch(id); }
  }
  int n_empty = 0, n_full = 0;
  for (int c = 0; c < 8; ++c) { const Ref r = child_ref(id, c); n_empty += r.kind == 0; n_full += r.kind == 1; }
  if (n_empty == 8 || n_full == 8) {
    for (int c = 0; c < 8; ++c) set_child(id, c, Ref{ 0, 0 });
    free_subtree(Ref{ 3, id });
    return Ref{ n_full == 8 ? 1 : 0, 0 };
  }
  const uint32_t avg = average_data(id);
  if (avg != svo_.nodes[id].data) { svo_.nodes[id].data = avg; touch(id); }
  return Ref{ 3, id };
}

int DynamicSVO::BuildRange(int level, const int pos[3], BuildMode mode, const VoxelSource &src, std::string &err) {
  if (level < 1 || level > 16) { err = "BuildRange: level must be in 1..16"; return -1; }
  if (page_version_.empty() && !svo_.nodes.empty()) adopt_existing();
  ++version_;
  int pivot[3]; src.GetPivot(pivot);
  const int org[3] = { pos[0] - pivot[0], pos[1] - pivot[1], pos[2] - pivot[2] };
  Ref root;
  if (svo_.root == YV_FULL_NODE) root = Ref{ 1, 0 };
  else if (YV_IS_NULL(svo_.root)) root = Ref{ 0, 0 };
  else root = Ref{ 3, svo_.root };
  Ref out = merge(root, 0, 0, 0, 1 << level, mode, src, org);
  if (out.kind == 3) svo_.root = out.v;
  else if (out.kind == 1) svo_.root = YV_FULL_NODE;
  else if (out.kind == 2) {                     // a single voxel filling the whole cube: keep it visible as a node
    const uint32_t id = alloc_node();
    for (int c = 0; c < 8; ++c) set_child(id, c, Ref{ 2, out.v });
    svo_.nodes[id].data = out.v;
    svo_.root = id;
  } else svo_.root = YV_EMPTY_NODE;
  svo_.depth = std::max<uint32_t>(svo_.depth, (uint32_t)level);
  return 0;
}

std::vector<int> DynamicSVO::GetNodeCountByLevel() const {
  std::vector<int> counts;
  if (YV_IS_NULL(svo_.root)) return counts;
  std::vector<uint32_t> cur{ svo_.root }, nxt;
  while (!cur.empty() && counts.size() < 40) {
    counts.push_back((int)cur.size());
    nxt.clear();
    for (uint32_t id : cur)
      for (int c = 0; c < 8; ++c) { const Ref r = child_ref(id, c); if (r.kind == 3) nxt.push_back(r.v); }
    cur.swap(nxt);
  }
  return counts;
}

int DynamicSVO::CountChangedPages(uint32_t since_version) const {
  int n = 0;
  for (uint32_t v : page_version_) n += v > since_version;
  return n;
}

}  // namespace yv
