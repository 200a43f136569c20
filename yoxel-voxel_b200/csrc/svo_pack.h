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
// On the device the four words a descent needs travel in one 16-byte load — { child_base, masks, octants lo, octants hi }
// — and the two words only a hit needs — { leaf_base, orig_id } — live in a side array (device_layout below).
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

// The device form of the records: trav[i] = { child_base, masks, octants lo, octants hi }, info[i] = { leaf_base, orig_id }
struct DeviceRecord { uint32_t child_base, masks, oct_lo, oct_hi; };
struct DeviceRecordInfo { uint32_t leaf_base, orig_id; };
static_assert(sizeof(DeviceRecord) == 16 && sizeof(DeviceRecordInfo) == 8, "device record layout");
void device_layout(const PackedSVO &p, std::vector<DeviceRecord> &trav, std::vector<DeviceRecordInfo> &info);

// Shared sub-trees (a DAG) are duplicated; cyclic pools are rejected via the level limit.
int pack_svo(const HostSVO &svo, PackedSVO &out, std::string &err);

}  // namespace yv
