#!/usr/bin/env python
"""gen_spheres.py of the reference (gen_spheres.py:1-36) on this library: the recursive fractal of spheres, built
with DynamicSVO.BuildRange exactly as the original script does, saved as a .vox file.
    python tools/gen_spheres.py [level=11] [out=data/spheres.vox]
(level 11 is the original; the batch builder yv.SVOData.SphereFractal(level) gives the same tree faster.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numpy import array  # noqa: E402

from yoxel_voxel_b200 import DynamicSVO, MakeSphereSource, BuildMode  # noqa: E402

level = int(sys.argv[1]) if len(sys.argv) > 1 else 11
out = sys.argv[2] if len(sys.argv) > 2 else "data/spheres.vox"

bld = DynamicSVO()

LevelNum = 8
BaseRadius = 256 * 2 ** (level - 11) if level >= 11 else 256 // 2 ** (11 - level)

lev_gens = []
for lev in range(LevelNum):
    lev_gens.append(MakeSphereSource(BaseRadius // (2 ** lev), (128, 128, lev * 255 // LevelNum), False))


def Rec(lev, pos, x, y, z):
    if lev > 4 and BaseRadius // (2 ** lev) >= 1:
        bld.BuildRange(level, [int(v) for v in pos], BuildMode.GROW, lev_gens[lev])
    if lev < LevelNum - 1:
        (x1, y1, z1) = [v / 2 for v in (x, y, z)]
        Rec(lev + 1, pos + x, y1, z1, x1)
        Rec(lev + 1, pos - x, y1, z1, -x1)
        Rec(lev + 1, pos + y, x1, z1, y1)
        Rec(lev + 1, pos - y, x1, z1, -y1)
        Rec(lev + 1, pos + z, x1, y1, z1)


c = 2.0 ** (level - 1)
Rec(0, array([c] * 3), array([BaseRadius * 1.5, 0, 0]), array([0, BaseRadius * 1.5, 0]), array([0, 0, BaseRadius * 1.5]))
print("nodes:", bld.livenodes)
print("saving...")
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
bld.Save(out)
