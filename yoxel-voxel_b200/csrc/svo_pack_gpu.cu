// svo_pack_gpu.cu — see svo_pack_gpu.h. One level of the tree per iteration:
//   count : per frontier node, how many existing children and inline leaves it has
//   scan  : exclusive prefix sums of both counts (cub::DeviceScan — library code, set-up path only)
//   emit  : write the node's 16-byte record, its leaves, its data word, and its children into the next frontier
// and, once every level is out, one pass that fills in each record's grandchild mask from its children's records.
// The frontier order of level L+1 is the concatenation of the children of level L's nodes in node order and child
// order 0..7, which is exactly the order of the host BFS in svo_pack.cpp.
#include "svo_pack_gpu.h"

#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <vector>

namespace yv {
namespace {

__device__ __forceinline__ void node_masks(const yv_vox_node &nd, uint32_t &leaf_mask, uint32_t &child_mask) {
  leaf_mask = nd.flags & 0xffu;
  child_mask = 0u;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (!((leaf_mask >> c) & 1u) && !(nd.child[c] & YV_NULL_BIT)) child_mask |= 1u << c;
}

__global__ void count_kernel(const yv_vox_node *raw, const uint32_t *frontier, uint32_t n, uint32_t *cc, uint32_t *lc,
                             uint32_t pool_size, int *bad) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const yv_vox_node nd = raw[frontier[i]];
  uint32_t lm, cm;
  node_masks(nd, lm, cm);
  for (int c = 0; c < 8; ++c)
    if (((cm >> c) & 1u) && nd.child[c] >= pool_size) *bad = 1;       // dangling child id
  cc[i] = __popc(cm);
  lc[i] = __popc(lm);
}

__global__ void emit_kernel(const yv_vox_node *raw, const uint32_t *frontier, uint32_t n, const uint32_t *coff,
                            const uint32_t *loff, uint32_t base, uint32_t leaf_base, uint4 *recs, uint32_t *leaves,
                            uint32_t *node_data, uint32_t *next_frontier) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t id = frontier[i];
  const yv_vox_node nd = raw[id];
  uint32_t lm, cm;
  node_masks(nd, lm, cm);
  uint32_t ci = coff[i], li = leaf_base + loff[i];
  recs[base + i] = make_uint4(base + n + ci, lm | (cm << 8), li, id);
  node_data[base + i] = nd.data;
  for (int c = 0; c < 8; ++c) {
    if ((lm >> c) & 1u) leaves[li++] = nd.child[c];
    else if ((cm >> c) & 1u) next_frontier[ci++] = nd.child[c];
  }
}

// byte c of a record's grandchild mask = (leaf flags | child flags) of its child node c; children are contiguous
__global__ void octants_kernel(const uint4 *recs, uint2 *octs, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 r = recs[i];
  uint32_t lo = 0u, hi = 0u, k = 0u;
  for (uint32_t c = 0; c < 8; ++c)
    if ((r.y >> (8 + c)) & 1u) {
      const uint32_t m = recs[r.x + k++].y;
      const uint32_t occ = (m | (m >> 8)) & 0xffu;
      if (c < 4) lo |= occ << (8 * c); else hi |= occ << (8 * (c - 4));
    }
  octs[i] = make_uint2(lo, hi);
}

#define PK_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); goto fail; } \
  } while (0)

}  // namespace

int pack_svo_on_device(const yv_vox_node *d_raw, size_t count, yv_node_id root, DevicePacked &out, std::string &err) {
  out = DevicePacked();
  if (YV_IS_NULL(root)) return 0;
  if (root >= count) { err = "root id outside node pool"; return -1; }
  const uint32_t N = (uint32_t)count;
  uint32_t *front[2] = { nullptr, nullptr }, *cc = nullptr, *lc = nullptr, *coff = nullptr, *loff = nullptr;
  void *tmp = nullptr; size_t tmp_bytes = 0;
  int *d_bad = nullptr;
  uint4 *recs = nullptr; uint2 *octs = nullptr; uint32_t *leaves = nullptr, *node_data = nullptr;
  size_t leaves_cap = 0;
  // a tree has one record per reachable node: at most N records; leaves at most 8 per node, grown on demand
  PK_CUDA(cudaMalloc(&front[0], (size_t)N * 4)); PK_CUDA(cudaMalloc(&front[1], (size_t)N * 4));
  PK_CUDA(cudaMalloc(&cc, (size_t)N * 4)); PK_CUDA(cudaMalloc(&lc, (size_t)N * 4));
  PK_CUDA(cudaMalloc(&coff, (size_t)N * 4)); PK_CUDA(cudaMalloc(&loff, (size_t)N * 4));
  PK_CUDA(cudaMalloc(&d_bad, sizeof(int))); PK_CUDA(cudaMemset(d_bad, 0, sizeof(int)));
  PK_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cc, coff, (int)N));
  PK_CUDA(cudaMalloc(&tmp, tmp_bytes));
  PK_CUDA(cudaMalloc(&recs, (size_t)N * sizeof(uint4)));
  PK_CUDA(cudaMalloc(&octs, (size_t)N * sizeof(uint2)));
  PK_CUDA(cudaMalloc(&node_data, (size_t)N * 4));
  leaves_cap = std::max<size_t>((size_t)N * 3, 1024);
  PK_CUDA(cudaMalloc(&leaves, leaves_cap * 4));
  PK_CUDA(cudaMemcpy(front[0], &root, 4, cudaMemcpyHostToDevice));
  {
    uint32_t n = 1, base = 0, leaf_base = 0;
    int level = 0, cur = 0;
    while (n > 0) {
      if (level >= 32) { err = "node pool deeper than 32 levels (cycle?)"; goto fail; }
      if ((uint64_t)base + n > N) { err = "node pool is not a tree (shared sub-trees): use the host repack"; goto fail; }
      const unsigned grid = (n + 255) / 256;
      count_kernel<<<grid, 256>>>(d_raw, front[cur], n, cc, lc, N, d_bad);
      PK_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cc, coff, (int)n));
      PK_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, lc, loff, (int)n));
      uint32_t last[4];
      PK_CUDA(cudaMemcpy(&last[0], cc + (n - 1), 4, cudaMemcpyDeviceToHost));
      PK_CUDA(cudaMemcpy(&last[1], coff + (n - 1), 4, cudaMemcpyDeviceToHost));
      PK_CUDA(cudaMemcpy(&last[2], lc + (n - 1), 4, cudaMemcpyDeviceToHost));
      PK_CUDA(cudaMemcpy(&last[3], loff + (n - 1), 4, cudaMemcpyDeviceToHost));
      const uint64_t n_children = (uint64_t)last[0] + last[1], n_leaves = (uint64_t)last[2] + last[3];
      if ((uint64_t)base + n + n_children > N) { err = "node pool is not a tree (shared sub-trees): use the host repack"; goto fail; }
      if ((uint64_t)leaf_base + n_leaves > 0xfffffff0ull) { err = "leaf array exceeds 2^32 entries"; goto fail; }
      if ((size_t)leaf_base + n_leaves > leaves_cap) {                 // grow the leaf array
        const size_t cap2 = std::max<size_t>(leaves_cap * 2, (size_t)leaf_base + n_leaves);
        uint32_t *l2 = nullptr;
        PK_CUDA(cudaMalloc(&l2, cap2 * 4));
        PK_CUDA(cudaMemcpy(l2, leaves, (size_t)leaf_base * 4, cudaMemcpyDeviceToDevice));
        cudaFree(leaves); leaves = l2; leaves_cap = cap2;
      }
      emit_kernel<<<grid, 256>>>(d_raw, front[cur], n, coff, loff, base, leaf_base, recs, leaves, node_data, front[cur ^ 1]);
      PK_CUDA(cudaGetLastError());
      base += n; leaf_base += (uint32_t)n_leaves; n = (uint32_t)n_children; cur ^= 1; ++level;
    }
    int bad = 0;
    PK_CUDA(cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    if (bad) { err = "child id outside node pool"; goto fail; }
    octants_kernel<<<(base + 255) / 256, 256>>>(recs, octs, base);
    PK_CUDA(cudaGetLastError());
    PK_CUDA(cudaDeviceSynchronize());
    out.recs = recs; out.octs = octs; out.leaves = leaves; out.node_data = node_data;
    out.n_recs = base; out.n_leaves = leaf_base; out.levels = level;
  }
  cudaFree(front[0]); cudaFree(front[1]); cudaFree(cc); cudaFree(lc); cudaFree(coff); cudaFree(loff); cudaFree(tmp); cudaFree(d_bad);
  return 0;
fail:
  cudaFree(front[0]); cudaFree(front[1]); cudaFree(cc); cudaFree(lc); cudaFree(coff); cudaFree(loff); cudaFree(tmp); cudaFree(d_bad);
  cudaFree(recs); cudaFree(octs); cudaFree(leaves); cudaFree(node_data);
  out = DevicePacked();
  return -1;
}

}  // namespace yv
