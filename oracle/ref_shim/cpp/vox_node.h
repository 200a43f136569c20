/* vox_node.h — stand-in for the reference's cpp/vox_node.h (absent).  TEST INFRASTRUCTURE ONLY.
 * The node record as the report documents it (reaction/report/main.tex:38-55) and as its users touch it
 * (cell/ppu_renderer.cpp:23,27,35,67; cell/spu/trace_spu.cpp:28,105-113; cell/svodata.h:26,42,54):
 * {flags, data, child[8]}, 40 bytes, 4-byte packed; ids with the top bit set are null (EmptyNode / FullNode);
 * leaf flag i = bit i of flags. */
#ifndef YV_REF_SHIM_VOX_NODE_H
#define YV_REF_SHIM_VOX_NODE_H

typedef unsigned int VoxNodeId;
typedef unsigned int VoxData;

#pragma pack(push, 4)
struct VoxNode {
  unsigned int flags;
  VoxData data;
  VoxNodeId child[8];
};
#pragma pack(pop)

const VoxNodeId EmptyNode = 0x80000000u;
const VoxNodeId FullNode = 0x80000001u;

/* Color32 as its users build and consume it (cell/ppu_renderer.cpp:54: four ints R,G,B,A; cell/main.cpp:36: "RGBA",
 * CharPixel). cell/svorenderer.h:23 names it after including only svodata.h -> vox_node.h, alignedarray.h, utils.h,
 * so in the reference it came from one of the shared headers; it lives here. */
struct Color32 {
  unsigned char r, g, b, a;
  Color32() {}
  Color32(int r_, int g_, int b_, int a_) : r((unsigned char)r_), g((unsigned char)g_), b((unsigned char)b_), a((unsigned char)a_) {}
};

inline bool IsNull(VoxNodeId id) { return (id & 0x80000000u) != 0u; }
inline bool GetLeafFlag(unsigned int flags, int i) { return ((flags >> i) & 1u) != 0u; }

#endif
