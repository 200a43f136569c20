"""Generates tests/golden/reference_golden.npz from the REFERENCE's own code (not from the oracle).

oracle/_ref holds cell/ppu_renderer.cpp and cell/spu/trace_spu.cpp compiled unmodified from /root/reference
(oracle/Makefile, target `ref`; stand-ins for the absent cpp/*.h under oracle/ref_shim/). This script runs them on the
committed scenes and cameras and stores what they return, so that the oracle (CPU tests) and the CUDA path (GPU tests,
on a box where /root/reference does not exist) are checked against reference output — tests/test_reference_golden.py.

    python tests/golden/make_reference_golden.py          (needs /root/reference; run in the build container)

Contents, per scene/camera key:
  <scene>/<camera>/rgba ....... SimpleRenderer::RenderFrame, RGBA8 [H][W][4]
  <scene>/<camera>/tbits ...... bits of TraceResult::t per pixel (probe shader), 0 where the ray missed
  <scene>/<camera>/data ....... VoxData the frame was shaded with = nodes[res.node].child[res.child]
  <scene>/<camera>/spu_rgba ... the SPU program's frame (clear colour (0,0,0,255))
  <scene>/<camera>/spu_fetches  FetchNode calls of that frame (= node visits)
  tagged/<camera>/ids ......... fractal with leaf words node*8+child: the hit ids straight from RecTrace
  raydir/in, raydir/out ....... InitRayDir: (dir, up, fov, W, H) -> (dir0, du, dv) for 256 cameras
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import scenes  # noqa: E402
import yvo  # noqa: E402
import yvref  # noqa: E402
import yoxel_voxel_b200 as yv  # noqa: E402

W, H = 64, 48                       # the frame size of golden.npz; a multiple of the SPU block
SCENES = ("sphere6", "dense4", "two_level")
TAGGED_DEPTH = 8


def tagged_fractal():
    svo = scenes.fractal(TAGGED_DEPTH)
    nodes = svo.nodes().copy()
    ids = (np.arange(len(nodes), dtype=np.uint32)[:, None] * 8 + np.arange(8, dtype=np.uint32)[None, :])
    leaf = ((nodes["flags"][:, None] >> np.arange(8)[None, :]) & 1).astype(bool)
    nodes["child"] = np.where(leaf, ids, nodes["child"])
    return yv.SVOData.FromNodes(svo.GetRoot(), nodes)


def main():
    assert yvref.available(), "needs /root/reference (oracle/_ref)"
    out = {}
    for scene in SCENES:
        path = os.path.join(HERE, scene + ".vox")
        sc = yvref.Scene(path)
        root, nodes = yvo.load_vox(path)
        assert sc.root() == root
        for name, pos, d, up, fov in scenes.CAMERAS:
            key = "%s/%s" % (scene, name)
            rgba, tbits, data = yvref.ppu_result(sc, pos, d, up, fov, W, H)
            out[key + "/rgba"], out[key + "/tbits"], out[key + "/data"] = rgba, tbits, data
            d0, du, dv = yvref.init_ray_dir(d, up, fov, W, H)
            frame, fetches, _ = yvref.spu_frame(nodes, root, pos, d0, du, dv, W, H)
            out[key + "/spu_rgba"] = frame.view(np.uint8).reshape(H, W, 4)
            out[key + "/spu_fetches"] = np.int64(fetches)
        sc.close()
    tagged = tagged_fractal()
    path = os.path.join(HERE, "_tagged_tmp.vox")
    tagged.Save(path)
    sc = yvref.Scene(path)
    for name, pos, d, up, fov in scenes.CAMERAS:
        out["tagged/%s/ids" % name] = yvref.ppu_frame(sc, pos, d, up, fov, 96, 72, yvref.PROBE_DATA)
        out["tagged/%s/tbits" % name] = yvref.ppu_frame(sc, pos, d, up, fov, 96, 72, yvref.PROBE_T)
    sc.close()
    os.remove(path)
    rng = np.random.RandomState(11)
    cams = np.zeros((256, 9), np.float32)
    res = np.zeros((256, 9), np.float32)
    for i in range(256):
        d = rng.randn(3)
        up = rng.randn(3) if i % 2 else np.array([0.0, 0.0, 1.0])
        fov = rng.uniform(5, 170) if i % 3 else [70.0, 40.0, 55.0, 110.0, 90.0, 60.0][i % 6]
        w, h = rng.choice([64, 160, 333, 640, 1024, 1920, 3840, 7680]), rng.choice([48, 77, 480, 768, 1080, 2160, 4320])
        cams[i] = list(d) + list(up) + [fov, w, h]
        res[i] = np.concatenate(yvref.init_ray_dir(cams[i, 0:3], cams[i, 3:6], cams[i, 6], int(cams[i, 7]), int(cams[i, 8])))
    out["raydir/in"], out["raydir/out"] = cams, res
    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
