// render_main.cpp — the reference's one-shot driver (cell/main.cpp:21-56) against the B200 renderer:
// load (or build) a scene, 1024x768, eye (0.5,0.5,0.3), time one RenderFrame, write the frame.
// The JPEG writer of the reference (Magick++) is replaced by a binary PPM.
//   render_main [scene.vox | --fractal DEPTH] [out.ppm] [W H] [dx dy dz]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "../include/yv_renderer.hpp"

using namespace yv;

static double mytime() {                                         // main.cpp:9-18
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

static int testRenderer(std::shared_ptr<ISVORenderer> renderer, SVOData &scene, const char *outfn,
                        int w, int h, point_3f dir) {             // main.cpp:21-40
  renderer->SetScene(&scene);
  renderer->SetResolution(w, h);
  renderer->SetViewPos(point_3f(0.5f, 0.5f, 0.3f));
  renderer->SetViewDir(dir);

  double start = mytime();
  const Color32 *frameBuf = renderer->RenderFrame();
  double dt = (mytime() - start) * 1000.0;
  if (!frameBuf) { std::fprintf(stderr, "RenderFrame failed: %s\n", yv_last_error()); return 2; }
  std::printf("time: %f ms\n", (float)dt);
  double start2 = mytime();
  frameBuf = renderer->RenderFrame();
  std::printf("time (second frame): %f ms\n", (float)((mytime() - start2) * 1000.0));

  if (outfn) {
    point_2i size = renderer->GetResolution();
    FILE *f = std::fopen(outfn, "wb");
    if (!f) return 3;
    std::fprintf(f, "P6\n%d %d\n255\n", size.x, size.y);
    for (int i = 0; i < size.x * size.y; ++i) std::fwrite(&frameBuf[i], 1, 3, f);
    std::fclose(f);
  }
  return 0;
}

int main(int argc, char **argv) {
  SVOData scene;
  int arg = 1;
  if (argc > 2 && !std::strcmp(argv[1], "--fractal")) { scene.BuildSphereFractal(std::atoi(argv[2]), 8); arg = 3; }
  else if (argc > 1) { scene.Load(argv[1]); arg = 2; }
  else scene.Load("../data/scene.vox");                           // main.cpp:45
  if (scene.GetNodeCount() == 0) { std::fprintf(stderr, "no scene: %s\n", yv_last_error()); return 1; }
  const char *out = argc > arg ? argv[arg] : nullptr;
  int w = argc > arg + 2 ? std::atoi(argv[arg + 1]) : 1024, h = argc > arg + 2 ? std::atoi(argv[arg + 2]) : 768;
  point_3f dir(-1, -1, -1.5f);                                    // main.cpp:26
  if (argc > arg + 5) dir = point_3f((float)std::atof(argv[arg + 3]), (float)std::atof(argv[arg + 4]), (float)std::atof(argv[arg + 5]));
  std::shared_ptr<ISVORenderer> renderer = CreateB200Renderer();
  return testRenderer(renderer, scene, out, w, h, dir);
}
