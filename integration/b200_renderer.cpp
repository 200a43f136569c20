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
//
// Like SPURenderer, which takes every usable SPE of the machine (cell/spu_renderer.cpp:73), the renderer made here
// drives every GPU of the machine inside each RenderFrame(); the environment variable YV_B200_DEVICES (a bit mask,
// decimal or 0x-hex) narrows that down, e.g. YV_B200_DEVICES=1 for the first GPU only; YV_B200_DEVICE_LIST=0,1,1 names
// the members one by one (a GPU listed twice is driven by two members).
#include "stdafx.h"
#include "svorenderer.h"

#include <cstdlib>

#define YV_USE_REFERENCE_TYPES
#include "yv_renderer.hpp"

static yv::B200Renderer *NewB200Renderer()
{
  if (const char *list = std::getenv("YV_B200_DEVICE_LIST"))
    if (*list) {
      std::vector<int> ordinals;
      for (const char *p = list; *p;) { char *end; ordinals.push_back((int)std::strtol(p, &end, 10)); p = *end ? end + 1 : end; }
      return new yv::B200Renderer(ordinals);
    }
  uint64_t mask = YV_ALL_DEVICES;
  if (const char *env = std::getenv("YV_B200_DEVICES"))
    if (*env) mask = std::strtoull(env, NULL, 0);
  return new yv::B200Renderer(yv::B200Renderer::Devices{ mask });
}

shared_ptr<ISVORenderer> CreateB200Renderer()
{
  return shared_ptr<ISVORenderer>(NewB200Renderer());
}

#ifdef YV_B200_AS_SPU_RENDERER
shared_ptr<ISVORenderer> CreateSPURenderer()
{
  return shared_ptr<ISVORenderer>(NewB200Renderer());
}
#endif
