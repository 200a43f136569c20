// b200_renderer.cpp — the translation unit a maintainer adds to the reference's cell/ directory to put the B200
// renderer behind the tree's own interface. It is compiled against the reference's unmodified
// cell/svorenderer.h / cell/svodata.h; the adapter class itself is include/yv_renderer.hpp in reference-types mode.
//
//   CreateB200Renderer()  a fourth factory next to CreateSimpleRenderer / CreateThreadedRenderer / CreateSPURenderer
//                         (cell/svorenderer.h:26-30)
//   -DYV_B200_AS_SPU_RENDERER  additionally defines CreateSPURenderer() itself, so that cell/main.cpp — which calls
//                         exactly that factory (cell/main.cpp:51-53) — links against this file instead of
//                         cell/spu_renderer.cpp and runs unmodified with the B200 in the place of the Cell's SPEs
//                         (tests/test_in_tree_binding.py does that).
#include "stdafx.h"
#include "svorenderer.h"

#define YV_USE_REFERENCE_TYPES
#include "yv_renderer.hpp"

shared_ptr<ISVORenderer> CreateB200Renderer()
{
  return shared_ptr<ISVORenderer>(new yv::B200Renderer(0));
}

#ifdef YV_B200_AS_SPU_RENDERER
shared_ptr<ISVORenderer> CreateSPURenderer()
{
  return shared_ptr<ISVORenderer>(new yv::B200Renderer(0));
}
#endif
