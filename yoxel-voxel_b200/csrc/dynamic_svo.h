// dynamic_svo.h — editable SVO: DynamicSVO::BuildRange with GROW / CLEAR and voxel sources, page versions.
//
// Reference (implementation absent from the snapshot; contract documented):
//   DynamicSVO API ........ ore/src/main.cpp:119-129 (BuildRange, Save, Load, TraceRay, nodecount,
//                           CountChangedPages, CountTransfrerSize, GetNodeCountByLevel)
//   VoxelSource ........... reaction/report/main.tex:88-94 (TryRange: empty / full / surface voxel / subdivide),
//                           ore/src/main.cpp:106-117 (GetSize, GetPivot; Raw, Sphere, Iso sources)
//   edit call sites ....... demo/Demo.cpp:82-114 (grow: SphereSource(4, colour, false) + BUILD_MODE_GROW;
//                           carve: SphereSource(4, colour, true) + BUILD_MODE_CLEAR), qtview.py:151-153,
//                           gen_spheres.py:18, gen_largevol.py:26-30, scene_gen.py:80-130
//   storage ............... main.tex:69-71: pool of equal-sized nodes with a free list; 256-node pages, each
//                           with a version number bumped on any write, used to ship only changed pages to the GPU
//
// The pool stays in the reference's 40-byte layout so that .vox files, the oracle and the raw-layout kernel
// all read it unchanged.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "svo_host.h"

namespace yv {

enum class BuildMode : int { Grow = 0, Clear = 1 };       // BUILD_MODE_GROW / BUILD_MODE_CLEAR

// VoxelSource: a finite voxel shape on the integer grid of the level it is built at
class VoxelSource {
 public:
  virtual ~VoxelSource() {}
  virtual void GetSize(int size[3]) const = 0;            // extent in voxels
  virtual void GetPivot(int pivot[3]) const = 0;          // voxel that BuildRange's `pos` refers to
  // classify the cube [p, p+size)^3 given in the source's own voxel coordinates
  virtual RangeClass TryRange(const int p[3], int size, uint32_t &voxdata) const = 0;
};

// MakeSphereSource(radius, colour, inverted) (ore/src/main.cpp:69): solid ball; `inverted` flips the
// surface normals (used with CLEAR so that a carved cavity shows its inside, demo/Demo.cpp:109)
class SphereSource : public VoxelSource {
 public:
  SphereSource(int radius, uint8_t r, uint8_t g, uint8_t b, bool inverted);
  void GetSize(int size[3]) const override;
  void GetPivot(int pivot[3]) const override;
  RangeClass TryRange(const int p[3], int size, uint32_t &voxdata) const override;
 private:
  int radius_; uint8_t col_[3]; bool inverted_;
};

// MakeRawSource(size, colours, normals) (ore/src/main.cpp:37-52): dense brick, x fastest. Two forms:
//   VoxData words, 0 = empty;
//   the reference's own pair of arrays — Color32 RGBA and Normal32 (int8 x,y,z,pad) per voxel — where the colour's
//   alpha tells the kind of voxel the way scene_gen.py:20-62 writes it: 0 empty, 255 surface, anything else buried
//   ("internal": no data, becomes part of a FullNode).
class RawSource : public VoxelSource {
 public:
  RawSource(const int size[3], const uint32_t *voxdata);   // copies
  RawSource(const int size[3], const uint8_t *colors_rgba, const int8_t *normals_xyzw);
  void GetSize(int size[3]) const override;
  void GetPivot(int pivot[3]) const override;
  RangeClass TryRange(const int p[3], int size, uint32_t &voxdata) const override;
 private:
  int size_[3]; std::vector<uint32_t> vox_; std::vector<uint8_t> kind_;   // kind: 0 empty, 1 buried, 2 surface voxel
};

// MakeIsoSource(size, uint8 data) + SetIsoLevel / SetInside / SetColor (ore/src/main.cpp:54-67,112-116)
class IsoBrickSource : public VoxelSource {
 public:
  IsoBrickSource(const int size[3], const uint8_t *data);  // x fastest; copies
  void SetIsoLevel(int level) { iso_ = level; }
  void SetInside(bool inside) { inside_ = inside; }
  void SetColor(uint8_t r, uint8_t g, uint8_t b) { col_[0] = r; col_[1] = g; col_[2] = b; }
  void GetSize(int size[3]) const override;
  void GetPivot(int pivot[3]) const override;
  RangeClass TryRange(const int p[3], int size, uint32_t &voxdata) const override;
 private:
  int at(int x, int y, int z) const;
  int size_[3]; std::vector<uint8_t> data_; int iso_ = 128; bool inside_ = false; uint8_t col_[3] = { 200, 200, 200 };
};

constexpr uint32_t kPageNodes = 256;                       // main.tex:71

// Editing state layered over a HostSVO (the pool itself lives in HostSVO::nodes)
class DynamicSVO {
 public:
  explicit DynamicSVO(HostSVO &svo) : svo_(svo) {}
  // BuildRange(level, pos, mode, src) (ore/src/main.cpp:121): merge `src`, placed with its pivot at voxel
  // `pos` of the 2^level grid, into the tree: GROW = union, CLEAR = subtraction.
  int BuildRange(int level, const int pos[3], BuildMode mode, const VoxelSource &src, std::string &err);
  uint32_t GetNodeCount() const { return (uint32_t)(svo_.nodes.size() - free_.size()); }
  std::vector<int> GetNodeCountByLevel() const;
  // pages written since `since_version`; CountTransfrerSize = pages * 256 * 40 bytes
  uint32_t version() const { return version_; }
  int CountChangedPages(uint32_t since_version) const;
  const std::vector<uint32_t> &page_versions() const { return page_version_; }
  void adopt_existing();     // after Load / a batch build: treat the current pool as version 1
  void reset_after_reload(); // the pool was replaced wholesale (SVOData::Load on a live scene): every page is new

 private:
  struct Ref { int kind; uint32_t v; };    // 0 empty, 1 full, 2 leaf(VoxData), 3 node(id)
  Ref merge(Ref cur, int x, int y, int z, int size, BuildMode mode, const VoxelSource &src, const int org[3]);
  uint32_t alloc_node();
  void free_subtree(Ref r);
  void touch(uint32_t id);
  Ref child_ref(uint32_t id, int c) const;
  void set_child(uint32_t id, int c, Ref r);
  uint32_t average_data(uint32_t id) const;

  HostSVO &svo_;
  std::vector<uint32_t> free_;
  std::vector<uint32_t> page_version_;
  uint32_t version_ = 0;
};

}  // namespace yv
