// svo_pack.h — repack of the reference node pool into the GPU-resident record form.
//
// Reference layout (reaction/report/main.tex:38-55): 40-byte AoS records, 4-byte aligned, eight
// 32-bit child slots that are a node id, an inline leaf VoxData or a null marker. A descent step
// needs only the two flag bytes and ONE child slot, but drags 2-3 32-byte sectors.
//
// Packed layout (this file): breadth-first renumbering so that the existing (non-null, non-leaf)
// children of a node are contiguous, which lets one 16-byte, 16-byte-aligned record replace the
// 40-byte node:
//     .x  child_base  packed index of the first existing child
//     .y  leaf_base   index of the first inline leaf of this node in the leaf array
//     .z  masks       bits 0..7 leaf flags (GetLeafFlag), bits 8..15 existing-child flags
//     .w  orig_id     the node's VoxNodeId in the reference pool (reported as the hit id)
// child c lives at child_base + popc(child_mask & ((1<<c)-1)); leaf c at leaf_base + popc(leaf_mask & ...).
// Next to the records: octants[i], the 64-bit "grandchild mask" of record i — byte c holds, for child NODE c, which of
// that child's eight octants contain anything (its leaf flags | child flags); 0 for leaf and empty slots. The traversal
// reads it with the record and uses it to skip child nodes the ray crosses through empty octants only (trace_core.cuh).
// On the device a record is { child_base, masks, leaf_base, orig_id }: the descent reads the first 8 bytes (LDG.64), a hit
// reads the other 8 of the same 32-byte sector (an L1 hit, the record was fetched a moment ago). The octant masks live in
// a side array that only the culling traversal (an ablation, off by default) reads. (Round 2 first put the octant words
// INTO the record and { leaf_base, orig_id } into the side array: every hit then paid a cold DRAM/L2 miss at the very end
// of its warp's life, +1.2 % frame time on config 2 — profiles/README.md.)
// Breadth-first order also puts the top of the tree at the lowest indices, so "stage the hot top
// levels in shared memory" (SPU software cache precedent: cell/spu/trace_spu.cpp:15-35) is the
// test `index < staged_count`.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "svo_host.h"

namespace yv {

struct PackedRecord { uint32_t child_base, leaf_base, masks, orig_id; };
static_assert(sizeof(PackedRecord) == 16, "record must be 16 bytes");

struct PackedSVO {
  std::vector<PackedRecord> records;     // records[0] is the root when !root_null
  std::vector<uint32_t> leaves;          // inline VoxData words, grouped per node
  std::vector<uint32_t> node_data;       // VoxNode::data (sub-tree average) per record, read only by LOD hits
  std::vector<uint64_t> octants;         // grandchild mask per record (see above)
  std::vector<uint32_t> level_start;     // first record index of each tree level (+ end sentinel)
  bool root_null = true;
};

// The device form of the records: trav[i] = { child_base, masks, leaf_base, orig_id }, octs[i] = { octants lo, octants hi }
struct DeviceRecord { uint32_t child_base, masks, leaf_base, orig_id; };
struct DeviceRecordOctants { uint32_t oct_lo, oct_hi; };
static_assert(sizeof(DeviceRecord) == 16 && sizeof(DeviceRecordOctants) == 8, "device record layout");
void device_layout(const PackedSVO &p, std::vector<DeviceRecord> &trav, std::vector<DeviceRecordOctants> &octs);

// Shared sub-trees (a DAG) are duplicated; cyclic pools are rejected via the level limit.
int pack_svo(const HostSVO &svo, PackedSVO &out, std::string &err);

}  // namespace yv
