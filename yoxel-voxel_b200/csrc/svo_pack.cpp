// svo_pack.cpp — breadth-first repack of the reference node pool (see svo_pack.h).
#include "svo_pack.h"

namespace yv {

namespace {

// A pool in which some node has two parents is a DAG (legal: shared sub-trees are duplicated by the repack) or cyclic
// (not). Decide which, and how many records the duplication produces, BEFORE the breadth-first expansion allocates
// anything: a small cyclic pool with branching would otherwise grow the frontier eight-fold per level until the host
// runs out of memory. Memoised depth-first walk: state 0 = new, 1 = on the walk's path, 2 = done.
int analyse_shared_pool(const HostSVO &svo, uint64_t limit, std::string &err) {
  const size_t n = svo.nodes.size();
  std::vector<uint8_t> state(n, 0);
  std::vector<uint64_t> expanded(n, 0);                    // records the sub-tree of a node expands to (saturating)
  struct Frame { uint32_t id; int next; };
  std::vector<Frame> path{ Frame{ svo.root, 0 } };
  state[svo.root] = 1; expanded[svo.root] = 1;
  while (!path.empty()) {
    Frame &f = path.back();
    if (f.next == 8) {
      state[f.id] = 2;
      const uint64_t mine = expanded[f.id];
      path.pop_back();
      if (!path.empty()) {
        uint64_t &up = expanded[path.back().id];
        up = (up + mine > limit) ? limit + 1 : up + mine;
      }
      continue;
    }
    const int c = f.next++;
    const yv_vox_node &nd = svo.nodes[f.id];
    if ((nd.flags >> c) & 1u) continue;
    const uint32_t v = nd.child[c];
    if (YV_IS_NULL(v)) continue;
    if (v >= n) { err = "child id outside node pool"; return -3; }
    if (state[v] == 1) { err = "node pool is cyclic (node " + std::to_string(v) + " is its own descendant)"; return -6; }
    if (state[v] == 2) {
      uint64_t &up = expanded[f.id];
      up = (up + expanded[v] > limit) ? limit + 1 : up + expanded[v];
      continue;
    }
    state[v] = 1; expanded[v] = 1;
    path.push_back(Frame{ v, 0 });
  }
  if (expanded[svo.root] > limit) { err = "packed pool exceeds 2^31 records (shared sub-trees are duplicated)"; return -4; }
  return 0;
}

}  // namespace

int pack_svo(const HostSVO &svo, PackedSVO &out, std::string &err) {
  out.records.clear(); out.leaves.clear(); out.level_start.clear(); out.node_data.clear(); out.octants.clear();
  out.root_null = YV_IS_NULL(svo.root);
  if (out.root_null) { out.level_start.push_back(0); return 0; }
  if (svo.root >= svo.nodes.size()) { err = "root id outside node pool"; return -1; }

  const uint64_t kMaxRecords = 0x7fffffffull;
  const int kMaxLevels = 32;
  std::vector<uint32_t> cur{ svo.root }, nxt;      // reference ids of the current / next level
  std::vector<bool> seen(svo.nodes.size(), false); // a node met twice: shared sub-tree or cycle, analysed before going on
  bool analysed = false;
  seen[svo.root] = true;
  out.records.reserve(svo.nodes.size());
  out.node_data.reserve(svo.nodes.size());
  uint64_t emitted = 0;
  for (int level = 0; !cur.empty(); ++level) {
    if (level >= kMaxLevels) { err = "node pool deeper than 32 levels (cycle?)"; return -2; }
    out.level_start.push_back((uint32_t)emitted);
    nxt.clear();
    // children of this level start right after the last record of this level
    uint64_t child_cursor = emitted + cur.size();
    for (uint32_t id : cur) {
      const yv_vox_node &nd = svo.nodes[id];
      PackedRecord r;
      uint32_t leaf_mask = nd.flags & 0xffu, child_mask = 0;
      r.leaf_base = (uint32_t)out.leaves.size();
      r.child_base = (uint32_t)child_cursor;
      for (int c = 0; c < 8; ++c) {
        const uint32_t v = nd.child[c];
        if ((leaf_mask >> c) & 1u) out.leaves.push_back(v);
        else if (!YV_IS_NULL(v)) {
          if (v >= svo.nodes.size()) { err = "child id outside node pool"; return -3; }
          if (seen[v] && !analysed) {
            const int rc = analyse_shared_pool(svo, kMaxRecords, err);
            if (rc) return rc;
            analysed = true;
          }
          seen[v] = true;
          child_mask |= 1u << c; nxt.push_back(v); ++child_cursor;
        }
      }
      r.masks = leaf_mask | (child_mask << 8);
      r.orig_id = id;
      out.records.push_back(r);
      out.node_data.push_back(nd.data);
    }
    emitted += cur.size();
    if (emitted + nxt.size() > kMaxRecords) { err = "packed pool exceeds 2^31 records"; return -4; }
    if (out.leaves.size() > 0xfffffff0ull) { err = "leaf array exceeds 2^32 entries"; return -5; }
    cur.swap(nxt);
  }
  out.level_start.push_back((uint32_t)emitted);
  // grandchild masks: byte c of a record = occupied octants (leaf | child flags) of its child node c
  out.octants.assign(out.records.size(), 0ull);
  for (size_t i = 0; i < out.records.size(); ++i) {
    const PackedRecord &r = out.records[i];
    uint64_t g = 0; uint32_t k = 0;
    for (uint32_t c = 0; c < 8; ++c)
      if ((r.masks >> (8 + c)) & 1u) {
        const uint32_t m = out.records[r.child_base + k++].masks;
        g |= (uint64_t)((m | (m >> 8)) & 0xffu) << (8 * c);
      }
    out.octants[i] = g;
  }
  return 0;
}

void device_layout(const PackedSVO &p, std::vector<DeviceRecord> &trav, std::vector<DeviceRecordOctants> &octs) {
  trav.resize(p.records.size()); octs.resize(p.records.size());
  for (size_t i = 0; i < p.records.size(); ++i) {
    const PackedRecord &r = p.records[i];
    const uint64_t g = i < p.octants.size() ? p.octants[i] : 0ull;
    trav[i] = DeviceRecord{ r.child_base, r.masks, r.leaf_base, r.orig_id };
    octs[i] = DeviceRecordOctants{ (uint32_t)g, (uint32_t)(g >> 32) };
  }
}

}  // namespace yv
