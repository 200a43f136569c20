"""tools/scene_gen.py — the reference's landscape script (scene_gen.py:1-141) on this library: MakeRawSource in the
reference's own argument form (colours + normals, alpha = empty / surface / buried) and region-wise BuildRange."""
import importlib.util
import os

import numpy as np

import yvo
import yoxel_voxel_b200 as yv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("scene_gen_tool", os.path.join(ROOT, "tools", "scene_gen.py"))
sg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(sg)

CAM = ((0.25, 0.2, 0.9), (0.15, 0.25, -1.0), (0, 1, 0), 60.0)
LEVEL, N = 7, 64


def _img(svo, W=160, H=120):
    return yvo.render(svo.nodes(), svo.GetRoot(), yvo.camera(CAM[0], CAM[1], CAM[2], CAM[3], W, H), threads=4)


def test_regions_compose_the_landscape():
    """64 regions of 8x8 columns, each a brick over its own height range [h0, h1] (so the ground under a region's h0
    stays empty, as in the reference), show the same landscape from above as one brick of the whole map."""
    hmap, tex = sg.synthetic_heightmap(N, amplitude=20.0)
    a = yv.DynamicSVO()
    sg.build_heightmap(a, hmap, tex, LEVEL, (0, 0, 0))
    cmap, nmap, h0, dh = sg.build_region(0, 0, N, N, hmap, tex)
    b = yv.DynamicSVO()
    b.BuildRange(LEVEL, (0, 0, h0), yv.BuildMode.GROW, yv.MakeRawSource((N, N, dh), cmap, nmap))
    assert sum(a.GetNodeCountByLevel1()) == a.livenodes and a.livenodes > 1000 and b.livenodes > 1000
    ia, ib = _img(a), _img(b)
    assert (ia["rgba"] == ib["rgba"]).all() and ia["t"].tobytes() == ib["t"].tobytes()
    assert (ia["node"] != yvo.MISS_NODE).sum() > 3000


def test_buried_voxels_become_full_nodes_and_surface_voxels_carry_texture_and_normal():
    n = 16
    hmap = np.full((n, n), 9.0, np.float32)                   # a slab 10 voxels thick
    tex = np.zeros((n, n, 3), np.uint8); tex[..., 0] = 200; tex[..., 1] = 40; tex[..., 2] = 80
    cmap, nmap, h0, dh = sg.build_region(0, 0, n, n, hmap, tex)
    assert (h0, dh) == (9, 1)                                 # a flat region is one layer: its own top (scene_gen.py:11-13)
    assert (cmap[0, 1:-1, 1:-1, 3] == 255).all() and (cmap[0, 0, :, 3] == 0).all()      # border ring left out (:26-27)
    assert (nmap[0, 1:-1, 1:-1, :3] == (0, 0, 127)).all()
    # a thick brick: everything under the top layer is buried and collapses into FullNode slots
    col = np.zeros((10, n, n, 4), np.uint8); nrm = np.zeros((10, n, n, 4), np.int8)
    col[:9, ..., 3] = 1                                       # buried
    col[9, ..., :3] = (200, 40, 80); col[9, ..., 3] = 255; nrm[9, ..., 2] = 127
    bld = yv.DynamicSVO()
    bld.BuildRange(4, (0, 0, 0), yv.BuildMode.GROW, yv.MakeRawSource((n, n, 10), col, nrm))
    nodes = bld.nodes()
    assert (nodes["child"] == yv.FULL_NODE).any()
    # looking straight down: every pixel inside the slab's footprint hits a top-layer voxel with the texture's colour
    o = yvo.render(nodes, bld.GetRoot(), yvo.camera((0.5, 0.5, 1.8), (0, 0, -1), (0, 1, 0), 30.0, 64, 64))
    hit = o["node"] != yvo.MISS_NODE
    assert hit.sum() > 2000
    data = nodes["child"][o["node"][hit], o["child"][hit]]
    assert (data == yv.pack_voxdata(200, 40, 80, 0, 0, 1)).all()
    top = 10.0 / 16.0
    assert np.allclose(1.8 - o["t"][hit] * np.abs(_dirz(o, hit)), top, atol=1e-3)


def _dirz(o, hit):
    # |dz| of the unit view rays at the hit pixels of the 64x64, 30-degree, straight-down camera
    W = H = 64
    d0, du, dv = yv.init_ray_dir((0, 0, -1), (0, 1, 0), 30.0, W, H)
    ys, xs = np.nonzero(hit)
    d = d0[None, :] + du[None, :] * xs[:, None] + dv[None, :] * ys[:, None]
    return (d[:, 2] / np.linalg.norm(d, axis=1)).astype(np.float32)


def test_trees_and_save_load_roundtrip(tmp_path):
    hmap, tex = sg.synthetic_heightmap(32, amplitude=8.0)
    bld = yv.DynamicSVO()
    sg.build_heightmap(bld, hmap, tex, 6, (0, 0, 0))
    before = bld.livenodes
    vol = sg.synthetic_volume(16)
    sg.add_trees(bld, vol, 6, [((4, 4, 12), 40), ((12, 4, 12), 60), ((4, 12, 12), 80)])
    assert bld.livenodes > before
    fn = str(tmp_path / "scene.vox")
    bld.Save(fn)
    back = yv.SVOData().Load(fn)
    a, b = _img(bld, 96, 72), _img(back, 96, 72)
    assert (a["rgba"] == b["rgba"]).all() and (a["node"] != yvo.MISS_NODE).sum() > 500
