#!/usr/bin/env python
"""scene_gen.py of the reference (scene_gen.py:1-141) on this library: a height-map landscape built region by region
through MakeRawSource(size, colours, normals) + BuildRange(GROW), and iso-surfaced volumes ("trees") through
MakeIsoSource + SetIsoLevel. The original reads data/gcanyon_height.png, data/gcanyon_color_4k2k.png and
data/bonsai.raw (not shipped); without --height-map / --volume this script makes seeded stand-ins of the same kind.

    python tools/scene_gen.py [--size 256] [--level 9] [--height-map H.npy --texture T.npy] [--volume V.npy] [--out data/scene.vox]

build_region / build_heightmap follow buildRegion / buildHeightmap (scene_gen.py:8-84): per 8x8 region the column range
[h0, h1] of its one-voxel neighbourhood, a voxel at or below the terrain height, buried (alpha 1) when none of its
four neighbours' columns is lower, else a surface voxel (alpha 255) with the height field's normal and the texture's
colour."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import yoxel_voxel_b200 as yv  # noqa: E402

STEP = 8                                                               # scene_gen.py:75


def build_region(x0, y0, x1, y1, hmap, tex):
    """buildRegion (scene_gen.py:8-70) -> (cmap [dh][y][x][4] uint8, nmap int8, h0, dh)."""
    H, W = hmap.shape
    sx = slice(max(0, x0 - 1), min(W, x1 + 1))
    sy = slice(max(0, y0 - 1), min(H, y1 + 1))
    h0, h1 = int(hmap[sy, sx].min()), int(hmap[sy, sx].max())
    dh = h1 - h0 + 1
    cmap = np.zeros((dh, y1 - y0, x1 - x0, 4), np.uint8)
    nmap = np.zeros((dh, y1 - y0, x1 - x0, 4), np.int8)
    ya, yb = max(y0, 1), min(y1, H - 1)                                # the outermost ring of the map is left out (:26-27)
    xa, xb = max(x0, 1), min(x1, W - 1)
    if ya >= yb or xa >= xb:
        return cmap, nmap, h0, dh
    ch = hmap[ya:yb, xa:xb].astype(np.int64)                           # (int)hmap(y, x)
    nb = np.stack([hmap[ya:yb, xa - 1:xb - 1], hmap[ya - 1:yb - 1, xa:xb], hmap[ya:yb, xa + 1:xb + 1],
                   hmap[ya + 1:yb + 1, xa:xb]]).astype(np.int64)
    lowest_nb = nb.min(axis=0)
    dx = (hmap[ya:yb, xa + 1:xb + 1] - hmap[ya:yb, xa - 1:xb - 1]) * np.float32(0.5)
    dy = (hmap[ya + 1:yb + 1, xa:xb] - hmap[ya - 1:yb - 1, xa:xb]) * np.float32(0.5)
    n = np.stack([-dx, -dy, np.ones_like(dx)], axis=-1).astype(np.float32)
    n = n / np.linalg.norm(n, axis=-1, keepdims=True) * np.float32(127.0)
    hs = (h0 + np.arange(dh))[:, None, None]                           # [dh][1][1]
    solid = hs <= ch[None]
    buried = solid & (hs < ch[None]) & ~(lowest_nb[None] < hs)         # h < ch and no neighbour column below h (:37-52)
    surface = solid & ~buried
    ry, rx = slice(ya - y0, yb - y0), slice(xa - x0, xb - x0)
    c = cmap[:, ry, rx]
    c[..., 3] = np.where(surface, 255, np.where(buried, 1, 0))
    c[..., :3] = np.where(surface[..., None], tex[ya:yb, xa:xb, :3][None], 0)
    nmap[:, ry, rx, :3] = np.where(surface[..., None], n[None].astype(np.int8), 0)      # float -> signed char, truncating (:64-66)
    return cmap, nmap, h0, dh


def build_heightmap(bld, hmap, tex, level, pos, log=None):
    """buildHeightmap (scene_gen.py:73-84)."""
    for y in range(hmap.shape[0] // STEP):
        for x in range(hmap.shape[1] // STEP):
            cmap, nmap, h0, dh = build_region(x * STEP, y * STEP, (x + 1) * STEP, (y + 1) * STEP, hmap, tex)
            src = yv.MakeRawSource((STEP, STEP, dh), cmap, nmap)
            bld.BuildRange(level, (pos[0] + x * STEP, pos[1] + y * STEP, pos[2] + h0), yv.BuildMode.GROW, src)
        if log:
            log(y, bld.livenodes)


def synthetic_heightmap(n, seed=7, amplitude=None):
    """Seeded stand-in for the canyon height map: a few octaves of smoothed noise, heights 0 .. n/4, plus a texture."""
    rng = np.random.RandomState(seed)
    h = np.zeros((n, n), np.float32)
    for octave in range(2, 6):
        k = 1 << octave
        coarse = rng.rand(k + 1, k + 1).astype(np.float32)
        xs = np.linspace(0, k, n, endpoint=False)
        i = xs.astype(int); f = (xs - i).astype(np.float32); f = f * f * (3 - 2 * f)
        rows = coarse[i][:, None, :] * (1 - f)[:, None, None] + coarse[i + 1][:, None, :] * f[:, None, None]
        rows = rows[:, 0, :]
        h += (rows[:, i] * (1 - f)[None, :] + rows[:, i + 1] * f[None, :]) / k
    h = (h - h.min()) / (h.max() - h.min())
    h = h * np.float32(amplitude if amplitude is not None else n / 4.0)
    tex = np.zeros((n, n, 3), np.uint8)
    t = h / max(1e-6, float(h.max()))
    tex[..., 0] = (60 + 150 * t).astype(np.uint8)
    tex[..., 1] = (120 - 40 * t).astype(np.uint8)
    tex[..., 2] = (40 + 30 * t).astype(np.uint8)
    return h, tex


def synthetic_volume(n, seed=11):
    """Seeded stand-in for bonsai.raw: a blobby density, uint8 [z][y][x]."""
    rng = np.random.RandomState(seed)
    z, y, x = np.mgrid[0:n, 0:n, 0:n].astype(np.float32) / n - 0.5
    d = np.zeros((n, n, n), np.float32)
    for _ in range(6):
        c = rng.uniform(-0.25, 0.25, 3)
        r = rng.uniform(0.08, 0.2)
        d += np.exp(-((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) / (r * r))
    return np.clip(d * 90, 0, 255).astype(np.uint8)


def add_trees(bld, vol, level, places, log=None):
    """addTrees (scene_gen.py:110-128): one IsoSource, three iso levels, three places."""
    for (pos, iso) in places:
        src = yv.MakeIsoSource(vol, iso_level=iso)
        bld.BuildRange(level, pos, yv.BuildMode.GROW, src)
        if log:
            log("tree", bld.livenodes)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256, help="height map edge in voxels")
    ap.add_argument("--level", type=int, default=9)
    ap.add_argument("--height-map", default=None); ap.add_argument("--texture", default=None); ap.add_argument("--volume", default=None)
    ap.add_argument("--out", default="data/scene.vox")
    a = ap.parse_args()
    if a.height_map:
        hmap = np.load(a.height_map).astype(np.float32); tex = np.load(a.texture)
    else:
        hmap, tex = synthetic_heightmap(a.size)
    vol = np.load(a.volume) if a.volume else synthetic_volume(max(16, a.size // 4))
    bld = yv.DynamicSVO()
    build_heightmap(bld, hmap, tex, a.level, (0, 0, 0), log=lambda y, n: print(y, n))
    q = a.size // 4
    add_trees(bld, vol, a.level, [((q, q, a.size // 4), 40), ((3 * q - vol.shape[2], q, a.size // 4), 60),
                                  ((q, 3 * q - vol.shape[1], a.size // 4), 80)], log=print)
    print("saving tree...", bld.livenodes, "nodes")
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    bld.Save(a.out)


if __name__ == "__main__":
    main()
