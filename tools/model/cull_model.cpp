// cull_model.cpp — analysis tool (not product, not test): over BASELINE config 2, how many node visits of the reference
// traversal are "fruitless" (the ray crosses the node and finds nothing), and how many of those could be recognised
// from the parent with an occupancy mask of the node's octants (a 64-bit grandchild mask in the parent's record).
//   python tools/model/dump_scene.py 12 && g++ -O2 -std=c++17 -ffp-contract=off -DYV_TEST_HOST_BUILD -o tools/model/_data/cm tools/model/cull_model.cpp && tools/model/_data/cm
#include <cstdio>
#include <vector>
#include "../../yoxel-voxel_b200/csrc/trace_core.cuh"
using namespace yv;
static const Rec *recs;
static long visits, shallow_fruitless, fruitless, pushes, pops, steps, leafmost_push, culled, culled_wrong, cand, culled_k[4], on_path;
static bool g_culled_flag;
struct Res { bool hit; bool descended; };
// returns hit; `sub` = number of visits in this subtree
static bool rec(uint32_t idx, RayState s, long &sub) {
  ++visits; long mine = 1;
  const Rec r = recs[idx];
  find_first_child(s);
  bool any_desc = false;
  for (;;) {
    const uint32_t c = s.ch ^ s.flags, bit = 1u << c;
    if (r.masks & bit) { sub += mine; ++on_path; return true; }
    const float t2min = min3f(s.t2x, s.t2y, s.t2z);
    if (((r.masks >> 8) & bit) && t2min > 0.0f) {
      any_desc = true;
      long subc = 0;
      RayState cs = s; cs.idx = r.child_base + __builtin_popcount((r.masks >> 8) & (bit - 1u));
      // can the parent still advance? (push would happen)
      const uint32_t e0 = (s.t2x > s.t2y) ? ((s.t2y < s.t2z) ? 1u : 2u) : ((s.t2x < s.t2z) ? 0u : 2u);
      const bool can_adv0 = (s.ch & (1u << e0)) == 0u;
      if (can_adv0) ++pushes;
      // ---- conservative cull test from the parent: octants of c the ray can touch = box [ch0, chx] ----
      bool cull = false;
      {
        ++cand;
        const Rec cr = recs[cs.idx];
        const uint32_t occ = (cr.masks | (cr.masks >> 8)) & 0xffu;          // stored-space occupancy of c's octants
        const float tmx = 0.5f * (cs.t1x + cs.t2x), tmy = 0.5f * (cs.t1y + cs.t2y), tmz = 0.5f * (cs.t1z + cs.t2z);
        const float te = max3f(cs.t1x, cs.t1y, cs.t1z), tx = min3f(cs.t2x, cs.t2y, cs.t2z);
        const uint32_t ch0 = (te > tmx ? 1u : 0u) | (te > tmy ? 2u : 0u) | (te > tmz ? 4u : 0u);
        auto crossed = [&](float tm) { return !(tm > tx + 1e-6f * (fabsf(tx) + fabsf(tm))); };
        const uint32_t chx = ch0 | (crossed(tmx) ? 1u : 0u) | (crossed(tmy) ? 2u : 0u) | (crossed(tmz) ? 4u : 0u);
        uint32_t box = 0;
        for (uint32_t o = 0; o < 8; ++o) if ((o & ch0) == ch0 && (o | chx) == chx) box |= 1u << (o ^ s.flags);
        cull = (box & occ) == 0u;
        if (cull) { ++culled; ++culled_k[__builtin_popcount(chx ^ ch0)]; }
      }
      const long sf0 = shallow_fruitless;
      const bool h = rec(cs.idx, cs, subc);
      if (cull && (h || shallow_fruitless == sf0)) ++culled_wrong;
      mine += subc;
      if (h) { sub += mine; ++on_path; return true; }
      if (can_adv0) ++pops;
    }
    const uint32_t e = (s.t2x > s.t2y) ? ((s.t2y < s.t2z) ? 1u : 2u) : ((s.t2x < s.t2z) ? 0u : 2u);
    if (s.ch & (1u << e)) break;
    ++steps;
    float *t1 = e == 0 ? &s.t1x : (e == 1 ? &s.t1y : &s.t1z), *t2 = e == 0 ? &s.t2x : (e == 1 ? &s.t2y : &s.t2z);
    const float dt = *t2 - *t1; *t1 = *t2; *t2 = *t2 + dt; s.ch ^= 1u << e;
  }
  if (!any_desc) ++shallow_fruitless;
  sub += mine;
  return false;
}
int main() {
  FILE *f = fopen("tools/model/_data/recs.bin", "rb"); fseek(f, 0, SEEK_END); long n = ftell(f) / 16; fseek(f, 0, SEEK_SET);
  std::vector<Rec> R(n); if (fread(R.data(), 16, n, f) != (size_t)n) return 1; fclose(f); recs = R.data();
  float cam[9]; f = fopen("tools/model/_data/cam.bin", "rb"); if (fread(cam, 4, 9, f) != 9) return 1; fclose(f);
  const int W = 1920, H = 1080; const float pos[3] = { 0.5f, 0.5f, 0.3f };
  long rays = 0, hits = 0;
  for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
    float dx, dy, dz; primary_dir(cam, cam + 3, cam + 6, x, y, dx, dy, dz); dx = adjust_dir1(dx); dy = adjust_dir1(dy); dz = adjust_dir1(dz);
    RayState s; ++rays;
    if (!setup_trace(pos[0], pos[1], pos[2], dx, dy, dz, s)) continue;
    if (min3f(s.t2x, s.t2y, s.t2z) <= 0.0f) continue;
    long sub = 0; s.idx = 0;
    if (rec(0, s, sub)) ++hits;
  }
  const double r = (double)rays;
  fruitless = visits - on_path;
  printf("candidate descents %.2f/ray, culled by the box test %.2f/ray (wrongly: %ld), by crossed planes 0..3: %.2f %.2f %.2f %.2f\n", cand / r, culled / r, culled_wrong, culled_k[0] / r, culled_k[1] / r, culled_k[2] / r, culled_k[3] / r);
  printf("rays %ld hits %.3f\nper ray: visits %.2f  fruitless visits %.2f  of which recognisable from the parent's octant mask (no occupied octant on the ray) %.2f\n"
         "         pushes %.2f pops %.2f sibling steps %.2f\n", rays, hits / r, visits / r, fruitless / r, shallow_fruitless / r, pushes / r, pops / r, steps / r);
  return 0;
}
