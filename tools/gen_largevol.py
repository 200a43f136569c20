#!/usr/bin/env python
"""gen_largevol.py of the reference (gen_largevol.py:1-41) on this library. The original iso-surfaces 320 bricks of a
CT data set (data/VolumeData/d_0219_*, 256x256x128 uint8 each, not shipped) at level 11 with iso level 200. With
--bricks DIR it does exactly that through MakeIsoSource + BuildRange; without, it builds the seeded synthetic stand-in
of the same proportions (yv.SVOData.IsoVolume) that the benchmark uses.
    python tools/gen_largevol.py [--bricks DIR] [--level 11] [--out data/large_vol.vox]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import yoxel_voxel_b200 as yv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bricks", default=None)
ap.add_argument("--level", type=int, default=11)
ap.add_argument("--out", default="data/large_vol.vox")
a = ap.parse_args()

if a.bricks:
    bld = yv.DynamicSVO()
    start, end = (3, 0, 0), (8, 8, 8)                                  # gen_largevol.py:8-9
    for k in range(start[0], end[0]):
        for j in range(start[1], end[1]):
            for i in range(start[2], end[2]):
                fn = os.path.join(a.bricks, "d_0219_%04d" % (k * 64 + j * 8 + i))
                print("processing", k, i, j, end=" ")
                try:
                    data = np.fromfile(fn, np.uint8).reshape(128, 256, 256)
                except Exception:
                    print("error")
                    continue
                src = yv.MakeIsoSource(data, iso_level=200)             # gen_largevol.py:26-27
                bld.BuildRange(a.level, (i * 256, j * 256, k * 128), yv.BuildMode.GROW, src)
                print(bld.livenodes)
else:
    bld = yv.SVOData.IsoVolume(a.level, seed=219, iso_level=200)
    print("synthetic iso volume, nodes:", bld.nodecount)
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
bld.Save(a.out)
