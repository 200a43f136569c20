"""DynamicSVO editing surface (ore/src/main.cpp:101-129; demo/Demo.cpp:82-114): BuildRange GROW / CLEAR with
voxel sources, node-count statistics, 256-node page versions, and (GPU) the page-wise device update with the
raw-layout kernel. The edited pool is the reference's 40-byte layout, so the oracle reads it directly."""
import numpy as np
import pytest

import scenes
import yvo
import yoxel_voxel_b200 as yv

CAM = ((0.5, 0.5, -1.2), (0.05, 0.1, 1.0), (0, 1, 0), 60.0)


def _img(svo, W=128, H=128, cam=CAM, **kw):
    return yvo.render(svo.nodes(), svo.GetRoot(), yvo.camera(cam[0], cam[1], cam[2], cam[3], W, H), **kw)


def test_grow_sphere_equals_batch_builder():
    bld = yv.DynamicSVO()
    src = yv.MakeSphereSource(19, (200, 120, 40), False)
    assert src.GetSize() == (39, 39, 39) and src.GetPivot() == (19, 19, 19)
    bld.BuildRange(6, (32, 33, 30), yv.BuildMode.GROW, src)
    ref = yv.SVOData.SingleSphere(6, (32, 33, 30), 19, (200, 120, 40))
    assert bld.livenodes == ref.nodecount
    a, b = _img(bld), _img(ref)
    assert (a["rgba"] == b["rgba"]).all() and a["t"].tobytes() == b["t"].tobytes()
    assert bld.GetNodeCountByLevel1()[0] == 1 and sum(bld.GetNodeCountByLevel1()) == bld.livenodes
    # growing the same shape again changes nothing
    v = bld.version
    bld.BuildRange(6, (32, 33, 30), yv.BuildMode.GROW, src)
    assert bld.CountChangedPages(v) == 0 and bld.livenodes == ref.nodecount


def test_gen_spheres_through_buildrange_equals_batch():
    """gen_spheres.py:16-32 issued as BuildRange calls (depth 8: 3125 spheres of radius 1)."""
    depth, LevelNum, BaseRadius = 8, 8, 32
    bld = yv.DynamicSVO()
    gens = [yv.MakeSphereSource(BaseRadius // (2 ** lev), (128, 128, lev * 255 // LevelNum), False) for lev in range(LevelNum)]

    def rec(lev, pos, x, y, z):
        if lev > 4 and BaseRadius // (2 ** lev) >= 1:
            bld.BuildRange(depth, [int(v) for v in pos], yv.BuildMode.GROW, gens[lev])
        if lev < LevelNum - 1:
            x1, y1, z1 = x / 2, y / 2, z / 2
            rec(lev + 1, pos + x, y1, z1, x1); rec(lev + 1, pos - x, y1, z1, -x1)
            rec(lev + 1, pos + y, x1, z1, y1); rec(lev + 1, pos - y, x1, z1, -y1)
            rec(lev + 1, pos + z, x1, y1, z1)

    c = 2.0 ** (depth - 1)
    rec(0, np.array([c, c, c]), np.array([BaseRadius * 1.5, 0, 0]), np.array([0, BaseRadius * 1.5, 0]), np.array([0, 0, BaseRadius * 1.5]))
    ref = yv.SVOData.SphereFractal(depth)
    assert bld.livenodes == ref.nodecount
    for cam in (scenes.CAMERAS[1], scenes.CAMERAS[2]):
        a = _img(bld, 160, 120, cam[1:])
        b = _img(ref, 160, 120, cam[1:])
        assert (a["rgba"] == b["rgba"]).all() and a["t"].tobytes() == b["t"].tobytes()


def test_clear_carves_a_walled_cavity_and_frees_nodes():
    bld = yv.DynamicSVO()
    bld.BuildRange(6, (32, 33, 30), yv.BuildMode.GROW, yv.MakeSphereSource(19, (200, 120, 40), False))
    before = _img(bld)
    v0 = bld.version
    bld.BuildRange(6, (32, 33, 12), yv.BuildMode.CLEAR, yv.MakeSphereSource(10, (192, 182, 128), True))   # Demo.cpp:109
    after = _img(bld)
    changed = (before["rgba"] != after["rgba"]).any(axis=2)
    assert 50 < changed.sum() < 2000
    # inside the cavity the ray travels farther and lands on the wall colour (RGB565 of 192,182,128)
    deeper = after["t"] > before["t"] + 1e-3
    assert deeper.sum() > 30
    nodes = bld.nodes()
    wall = nodes["child"][after["node"][deeper], after["child"][deeper]] & 0xFFFF
    assert (wall == (192 >> 3 << 11 | 182 >> 2 << 5 | 128 >> 3)).all()
    assert 0 < bld.CountChangedPages(v0) <= bld.CountChangedPages(0)
    # carving everything away empties the scene and returns every node to the free list
    bld.BuildRange(6, (32, 32, 32), yv.BuildMode.CLEAR, yv.MakeSphereSource(60, (1, 1, 1), True))
    assert bld.livenodes == 0 and bld.GetRoot() == yv.EMPTY_NODE and bld.GetNodeCountByLevel1() == []
    # freed nodes are reused
    pool = bld.nodecount
    bld.BuildRange(6, (20, 20, 20), yv.BuildMode.GROW, yv.MakeSphereSource(6, (10, 200, 10), False))
    assert bld.nodecount == pool and bld.livenodes > 0


def test_raw_and_iso_sources():
    rng = np.random.RandomState(2)
    brick = np.zeros((8, 8, 8), np.uint32)
    occ = rng.rand(8, 8, 8) < 0.2
    brick[occ] = [yv.pack_voxdata(50 + i % 200, 90, 30, 0, 0, 1) for i in range(occ.sum())]
    bld = yv.DynamicSVO()
    bld.BuildRange(5, (8, 12, 16), yv.BuildMode.GROW, yv.MakeRawSource(brick))
    nodes = bld.nodes()
    leaf = ((nodes["flags"][:, None] >> np.arange(8)) & 1).astype(bool)
    assert sorted(nodes["child"][leaf]) == sorted(brick[occ])            # every voxel landed, nothing else
    # iso source: a ball in a uint8 volume
    z, y, x = np.mgrid[0:24, 0:24, 0:24]
    vol = np.clip(255 - 14 * np.sqrt((x - 12) ** 2 + (y - 12) ** 2 + (z - 12) ** 2), 0, 255).astype(np.uint8)
    iso = yv.DynamicSVO()
    iso.BuildRange(6, (20, 20, 20), yv.BuildMode.GROW, yv.MakeIsoSource(vol, iso_level=150, color=(10, 220, 30)))
    r = _img(iso)
    assert (r["node"] != yvo.MISS_NODE).sum() > 100
    assert (r["rgba"][r["node"] != yvo.MISS_NODE][:, 1] > r["rgba"][r["node"] != yvo.MISS_NODE][:, 0]).all()   # green


def test_small_edit_touches_few_pages_of_a_large_scene():
    svo = yv.SVOData.SphereFractal(10)
    total_pages = (svo.nodecount + 255) // 256
    v0 = svo.version
    svo.BuildRange(10, (512, 512, 600), yv.BuildMode.GROW, yv.MakeSphereSource(4, (255, 0, 0), False))   # Demo.cpp:103-105
    changed = svo.CountChangedPages(max(v0, 1))
    assert 0 < changed < total_pages // 50
    assert svo.livenodes >= svo.nodecount - 1000


def test_scene_scripts(tmp_path):
    """tools/gen_spheres.py and tools/gen_largevol.py (the reference's scene scripts on this library)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "spheres.vox")
    subprocess.check_call([sys.executable, os.path.join(root, "tools", "gen_spheres.py"), "8", out], stdout=subprocess.DEVNULL)
    a, b = yv.SVOData().Load(out), yv.SVOData.SphereFractal(8)
    assert a.nodecount == b.nodecount
    assert (_img(a, 96, 96, scenes.CAMERAS[1][1:])["rgba"] == _img(b, 96, 96, scenes.CAMERAS[1][1:])["rgba"]).all()
    out2 = str(tmp_path / "vol.vox")
    subprocess.check_call([sys.executable, os.path.join(root, "tools", "gen_largevol.py"), "--level", "7", "--out", out2],
                          stdout=subprocess.DEVNULL)
    assert yv.SVOData().Load(out2).nodecount == yv.SVOData.IsoVolume(7).nodecount
    # and through MakeIsoSource bricks, as the original does
    bricks = tmp_path / "bricks"
    bricks.mkdir()
    z, y, x = np.mgrid[0:128, 0:256, 0:256]
    vol = np.clip(260 - 3 * np.sqrt((x - 128.0) ** 2 + (y - 128.0) ** 2 + (z - 64.0) ** 2), 0, 255).astype(np.uint8)
    vol.tofile(str(bricks / ("d_0219_%04d" % (3 * 64))))
    out3 = str(tmp_path / "ct.vox")
    subprocess.check_call([sys.executable, os.path.join(root, "tools", "gen_largevol.py"), "--bricks", str(bricks), "--out", out3],
                          stdout=subprocess.DEVNULL)
    ct = yv.SVOData().Load(out3)
    assert ct.nodecount > 100


def test_null_flags_are_recomputed_on_import():
    """A pool whose derived null flags are missing (e.g. written by another tool) is normalised on import, so the
    raw-layout kernel, which trusts them, cannot chase a null child id."""
    nodes, root, leaf = scenes.two_level_tree()
    nodes["flags"] &= 0xFF                       # drop every null flag
    svo = yv.SVOData.FromNodes(root, nodes)
    fixed = svo.nodes()
    assert (fixed["flags"] >> 8 & 0xFF).tolist() == [0x7F, 0xFE]


@pytest.mark.gpu
@pytest.mark.parametrize("sec", [None, dict(shadow=1, ao_samples=2, seed=5, light_pos=(0.6, 0.4, 1.2), voxel_size=1.0 / 1024, ao_max_t=0.05)],
                         ids=["primary", "secondary"])
def test_raw_layout_kernel_and_paged_update(sec):
    """Edit a scene while it is on the GPU: only dirty pages travel, and the raw-layout kernel matches the oracle
    before and after (ids are the reference ids in this layout, no remap)."""
    svo = yv.SVOData.SphereFractal(10)
    r = yv.SVORenderer(0)
    r.EnableHits(True)
    r.SetOption("layout", 1)
    r.SetScene(svo)
    cam = scenes.CAMERAS[1]

    def check(tag, detail=0.0):
        r.SetResolution(400, 300)
        r.SetViewPos(cam[1]); r.SetViewDir(cam[2]); r.SetViewUp(cam[3]); r.SetFOV(cam[4])
        r.SetDetailCoef(detail)
        r.SetSecondary(**sec) if sec else r.SetSecondary(0, 0)
        img = r.RenderFrame().copy()
        node, child, t = r.GetHits()
        o = yvo.render(svo.nodes(), svo.GetRoot(), yvo.camera(cam[1], cam[2], cam[3], cam[4], 400, 300, detail_coef=detail),
                       sec=yvo.secondary(**sec) if sec else None, threads=8)
        assert (node == o["node"]).all() and (child == o["child"]).all(), tag
        assert t.tobytes() == o["t"].tobytes() and (img == o["rgba"]).all(), tag
        return img

    first = svo.Update(0)
    assert first == svo.nodecount * 40                       # the whole pool on the first sync
    base = check("initial")
    assert svo.Update(0) == 0                                # nothing changed since
    # shoot spheres at what the centre ray sees, like Demo::DoEdit
    rng = np.random.RandomState(1)
    dirs = np.array(cam[2], np.float32) + 0.1 * rng.uniform(-1, 1, (12, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    node, child, t = r.TraceRays(np.tile(np.array(cam[1], np.float32), (12, 1)), dirs)      # DynamicSVO::TraceRay
    assert (node != yv.EMPTY_NODE).sum() >= 6
    grow, carve = yv.MakeSphereSource(4, (255, 40, 40), False), yv.MakeSphereSource(6, (192, 182, 128), True)
    for i in range(12):
        if node[i] == yv.EMPTY_NODE or t[i] <= 0:
            continue
        pt = (np.array(cam[1]) + dirs[i] * t[i]) * 1024
        svo.BuildRange(10, [int(v) for v in pt], yv.BuildMode.GROW if i % 2 else yv.BuildMode.CLEAR, grow if i % 2 else carve)
    sent = svo.Update(0)
    assert 0 < sent < first // 20                            # a few pages, not the pool
    edited = check("edited")
    assert (edited != base).any()
    check("edited+lod", detail=6.0)
    # the packed layout notices the edit, re-packs, and agrees as well
    r.SetOption("layout", 0)
    check("edited/packed")
    r.close()
