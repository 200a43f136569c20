"""Shared scene / camera fixtures for the tests (small, deterministic, built through the C ABI host code)."""
import functools

import numpy as np

import yoxel_voxel_b200 as yv

# cameras: (name, pos, dir, up, fov)
CAMERAS = [
    ("main_cpp", (0.5, 0.5, 0.3), (-1, -1, -1.5), (0, 0, 1), 70.0),     # cell/main.cpp:25-26
    ("main_cpp_up", (0.5, 0.5, 0.3), (-1, -1, 1.5), (0, 0, 1), 70.0),   # same eye, looking at the fractal
    ("outside", (1.3, 1.2, 0.9), (-0.8, -0.7, -0.4), (0, 0, 1), 70.0),
    ("demo_cpp", (0.16560096, 0.46532935, 0.11644295),                  # demo/Demo.cpp:11-13: course 281, pitch -11
     (np.cos(np.radians(281)) * np.cos(np.radians(-11)), np.sin(np.radians(281)) * np.cos(np.radians(-11)),
      np.sin(np.radians(-11))), (0, 0, 1), 70.0),
    ("inside", (0.52, 0.47, 0.61), (0.2, -1, 0.1), (0, 0, 1), 55.0),
    ("axis", (0.5, 0.5, -0.7), (0, 0, 1), (0, 1, 0), 40.0),             # exercises AdjustDir on the centre column
    ("wide_up", (0.3, 0.7, 0.2), (0.4, -0.3, 1.0), (1, 0, 0), 110.0),
]


@functools.lru_cache(maxsize=None)
def fractal(depth):
    return yv.SVOData.SphereFractal(depth, threads=8)


@functools.lru_cache(maxsize=None)
def single_sphere(depth=6):
    n = 1 << depth
    return yv.SVOData.SingleSphere(depth, (n // 2, n // 2 + 1, n // 2 - 2), int(n * 0.3), (200, 120, 40))


def dense_random_grid(depth=4, fill=0.08, seed=3):
    """[z][y][x] uint32 grid; every set voxel carries a unique id in the colour bits and a hashed normal."""
    n = 1 << depth
    rng = np.random.RandomState(seed)
    occ = rng.rand(n, n, n) < fill
    vox = np.zeros((n, n, n), np.uint32)
    ids = np.arange(1, occ.sum() + 1, dtype=np.uint32)
    # unique id -> low 16 bits (colour); normal bits from a hash so shading varies
    vox[occ] = (ids & 0xFFFF) | ((ids * 2654435761 & 0xFFFF).astype(np.uint32) << 16)
    return vox


@functools.lru_cache(maxsize=None)
def dense_random(depth=4, fill=0.08, seed=3):
    vox = dense_random_grid(depth, fill, seed)
    return yv.SVOData.FromDense(vox), vox


def two_level_tree():
    """Hand-built pool: root(id 1) -> child 0 = node(id 0) whose child 7 is a leaf.
    Leaf cube = [0.25,0.5]^3."""
    nodes = np.zeros(2, yv.NODE_DTYPE)
    leaf = yv.pack_voxdata(255, 0, 0, 0, 0, -1)
    inner = nodes[0]
    inner["child"][:] = yv.EMPTY_NODE
    inner["child"][7] = leaf
    inner["flags"] = (1 << 7) | (0x7F << 8)
    root = nodes[1]
    root["child"][:] = yv.EMPTY_NODE
    root["child"][0] = 0
    root["flags"] = (0xFE << 8)
    return nodes, 1, leaf
