// svo_pack_gpu.h — the breadth-first repack of svo_pack.h done on the GPU, level by level.
// Same output, bit for bit, as yv::pack_svo (tests compare the two); used for the device copy so that a
// 500 M-node pool is re-laid-out in well under a second instead of a minute of single-threaded host BFS,
// and so that an edited scene can go back to the packed layout without a host round trip.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>

#include "../../include/yv_format.h"

namespace yv {

struct DevicePacked {
  void *recs = nullptr;        // uint4[n_recs]: { child_base, masks, leaf_base, orig_id } (svo_pack.h, device form)
  void *octs = nullptr;        // uint2[n_recs]: { octants lo, octants hi } (culling traversal only)
  uint32_t *leaves = nullptr;  // [n_leaves]
  uint32_t *node_data = nullptr;  // [n_recs]
  size_t n_recs = 0, n_leaves = 0;
  int levels = 0;
};

// d_raw: the reference pool resident on the current device (count nodes). Allocates the outputs with cudaMalloc.
// Returns 0 on success; on failure `err` is set and nothing stays allocated.
int pack_svo_on_device(const yv_vox_node *d_raw, size_t count, yv_node_id root, DevicePacked &out, std::string &err);

}  // namespace yv
