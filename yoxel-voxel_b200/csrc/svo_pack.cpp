// svo_pack.cpp — breadth-first repack of the reference node pool (see svo_pack.h).
#include "svo_pack.h"

namespace yv {

int pack_svo(const HostSVO &svo, PackedSVO &out, std::string &err) {
  out.records.clear(); out.leaves.clear(); out.level_start.clear(); out.node_data.clear();
  out.root_null = YV_IS_NULL(svo.root);
  if (out.root_null) { out.level_start.push_back(0); return 0; }
  if (svo.root >= svo.nodes.size()) { err = "root id outside node pool"; return -1; }

  const uint64_t kMaxRecords = 0x7fffffffull;
  const int kMaxLevels = 32;
  std::vector<uint32_t> cur{ svo.root }, nxt;      // reference ids of the current / next level
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
  return 0;
}

}  // namespace yv
