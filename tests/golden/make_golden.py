"""Generates the committed fixtures under tests/golden/ from the oracle.

The reference ships no scene files, images or known-answer rays for this path, and
its tracer cannot be built here, so these vectors are produced by the CPU oracle after it has been
pinned by tests/test_oracle.py (hand-computed rays, quirk tests, grid-marcher cross-check). They are
regression pins for the oracle and fixed inputs/outputs for the CUDA parity tests.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import scenes  # noqa: E402
import yvo  # noqa: E402
import yoxel_voxel_b200 as yv  # noqa: E402

W, H = 64, 48
SEC = dict(shadow=1, ao_samples=4, seed=1, light_pos=(0.6, 0.4, 1.2), voxel_size=1.0 / 64, ao_max_t=0.2)


def main():
    out = {}
    sph = scenes.single_sphere(6)
    sph.Save(os.path.join(HERE, "sphere6.vox"))
    dense, _ = scenes.dense_random(4, 0.08)
    dense.Save(os.path.join(HERE, "dense4.vox"))
    nodes, root, _ = scenes.two_level_tree()
    yv.SVOData.FromNodes(root, nodes).Save(os.path.join(HERE, "two_level.vox"))
    for scene in ("sphere6", "dense4", "two_level"):
        root, nodes = yvo.load_vox(os.path.join(HERE, scene + ".vox"))
        for ci, (name, pos, d, up, fov) in enumerate(scenes.CAMERAS):
            cam = yvo.camera(pos, d, up, fov, W, H)
            r = yvo.render(nodes, root, cam)
            key = "%s/%s" % (scene, name)
            out[key + "/node"] = r["node"]
            out[key + "/child"] = r["child"].astype(np.int8)
            out[key + "/t"] = r["t"]
            out[key + "/rgba"] = r["rgba"]
            if scene == "sphere6":
                r2 = yvo.render(nodes, root, cam, sec=yvo.secondary(**SEC))
                out[key + "/rgba_sec"] = r2["rgba"]
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
