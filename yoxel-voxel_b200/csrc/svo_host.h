// svo_host.h — host-side SVO container: the reference node pool (40-byte VoxNode records,
// include/yv_format.h) plus the .vox reader/writer and the procedural scene builders.
//
// Reference: cell/svodata.h:22-55 (SVOData: root id + node array, Load);
//            reaction/report/main.tex:38-55 (VoxNode), :88-96 (VoxelSource::TryRange contract);
//            gen_spheres.py:5-35, gen_largevol.py:8-40 (the two scene generators).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/yv_format.h"

namespace yv {

struct HostSVO {
  yv_node_id root = YV_EMPTY_NODE;
  uint32_t depth = 0;                 // levels below the root cube (informational)
  std::vector<yv_vox_node> nodes;     // reference layout, reference numbering
};

// SVOData::Load (cell/svodata.h:31-50). Unlike the reference, I/O failures are reported.
int load_vox(const char *path, HostSVO &out, std::string &err);
int save_vox(const char *path, const HostSVO &svo, std::string &err);

// Recompute the derived null flags (bits 8..15) from the child words.
void normalize_flags(HostSVO &svo);

// Structural validation: root and every non-leaf, non-null child id must index the pool.
int validate(const HostSVO &svo, std::string &err);

// TryRange result (main.tex:88-94)
enum class RangeClass : int { Empty = 0, Full = 1, Voxel = 2, Mixed = 3 };

// Scene #1 — sphere fractal (gen_spheres.py:8-32) scaled to `depth` levels:
// centre 2^(depth-1), BaseRadius 2^(depth-3), 8 recursion levels, spheres for lev > 4.
int build_sphere_fractal(int depth, int threads, HostSVO &out, std::string &err);

// Scene #2 — synthetic iso-volume standing in for gen_largevol.py:8-40 (dataset not shipped):
// seeded 3-octave value noise over a slab (x,y full extent, z extent 5/16 of the cube),
// iso level `iso`/255, built top-down at `depth` levels.
int build_iso_volume(int depth, uint32_t seed, int iso, int threads, HostSVO &out, std::string &err);

// A single solid sphere (MakeSphereSource + one BuildRange; ore/src/main.cpp:69,121):
// used by small tests. Centre/radius in finest-level voxel units.
int build_single_sphere(int depth, int cx, int cy, int cz, int radius,
                        uint8_t r, uint8_t g, uint8_t b, HostSVO &out, std::string &err);

// Dense occupancy grid -> SVO (RawSource-like; ore/src/main.cpp:37-52). `vox` holds
// (1<<depth)^3 VoxData words in x-fastest order; 0 = empty. Used by tests.
int build_from_dense(int depth, const uint32_t *vox, HostSVO &out, std::string &err);

uint32_t pack_voxdata(uint8_t r, uint8_t g, uint8_t b, float nx, float ny, float nz);

}  // namespace yv
