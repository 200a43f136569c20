"""Committed output of the REFERENCE's own code (tests/golden/reference_golden.npz, written by
make_reference_golden.py from cell/ppu_renderer.cpp and cell/spu/trace_spu.cpp compiled unmodified into oracle/_ref).
The oracle and the kernel's host build must reproduce it on the CPU; the CUDA path must reproduce it on the GPU, where
/root/reference does not exist. Everything is compared bit for bit: RGBA, the bits of the hit distance, the VoxData,
hit ids (through leaf words that name their node and child), the number of node fetches, InitRayDir."""
import os

import numpy as np
import pytest

import scenes
import yve
import yvo
import yoxel_voxel_b200 as yv

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
W, H = 64, 48
SCENES = ("sphere6", "dense4", "two_level")
TAGGED_DEPTH, TW, TH = 8, 96, 72


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(HERE, "reference_golden.npz"))


def _tagged():
    svo = scenes.fractal(TAGGED_DEPTH)
    nodes = svo.nodes().copy()
    ids = (np.arange(len(nodes), dtype=np.uint32)[:, None] * 8 + np.arange(8, dtype=np.uint32)[None, :])
    leaf = ((nodes["flags"][:, None] >> np.arange(8)[None, :]) & 1).astype(bool)
    nodes["child"] = np.where(leaf, ids, nodes["child"])
    return yv.SVOData.FromNodes(svo.GetRoot(), nodes)


def test_reference_golden_has_content(ref):
    hits = sum(int((ref[k] != 0).sum()) for k in ref.files if k.endswith("/tbits"))
    assert hits > 20000
    assert sum(int(ref[k]) for k in ref.files if k.endswith("/spu_fetches")) > 100000


@pytest.mark.parametrize("scene", SCENES)
def test_oracle_reproduces_the_reference_frames(ref, scene):
    root, nodes = yvo.load_vox(os.path.join(HERE, scene + ".vox"))
    for name, pos, d, up, fov in scenes.CAMERAS:
        key = "%s/%s" % (scene, name)
        o = yvo.render(nodes, root, yvo.camera(pos, d, up, fov, W, H))
        hit = o["node"] != yvo.MISS_NODE
        assert (o["rgba"] == ref[key + "/rgba"]).all(), key
        assert (np.where(hit, o["t"].view(np.uint32), 0) == ref[key + "/tbits"]).all(), key
        data = np.where(hit, nodes["child"][np.where(hit, o["node"], 0), np.where(hit, o["child"], 0)], 0)
        assert (data == ref[key + "/data"]).all(), key
        spu = o["rgba"].copy()
        spu[~hit] = (0, 0, 0, 255)                                   # trace_spu.cpp:127
        assert (spu == ref[key + "/spu_rgba"]).all(), key
        assert o["stats"]["node_visits"] == int(ref[key + "/spu_fetches"]), key      # FetchNode calls, trace_spu.cpp:33


def test_oracle_and_emu_reproduce_the_reference_hit_ids(ref):
    svo = _tagged()
    nodes, root = svo.nodes(), svo.GetRoot()
    recs, leaves = svo.packed()
    n = 0
    for name, pos, d, up, fov in scenes.CAMERAS:
        ids, tbits = ref["tagged/%s/ids" % name], ref["tagged/%s/tbits" % name]
        o = yvo.render(nodes, root, yvo.camera(pos, d, up, fov, TW, TH))
        d0, du, dv = yv.init_ray_dir(d, up, fov, TW, TH)
        e = yve.render(recs, leaves, 1, pos, d0, du, dv, pos, TW, TH)
        for r in (o, e):
            hit = r["node"] != yvo.MISS_NODE
            assert ((ids != 0) | (tbits != 0) == hit).all(), name
            assert ((ids >> 3)[hit] == r["node"][hit]).all() and ((ids & 7)[hit] == r["child"][hit].astype(np.uint32)).all(), name
            assert (tbits[hit] == r["t"].view(np.uint32)[hit]).all(), name
        n += int(hit.sum())
    assert n > 5000


def test_init_ray_dir_reproduces_the_reference(ref):
    cams, want = ref["raydir/in"], ref["raydir/out"]
    for c, w in zip(cams, want):
        d, up, fov, Wc, Hc = tuple(c[0:3]), tuple(c[3:6]), float(c[6]), int(c[7]), int(c[8])
        mine = np.concatenate(yv.init_ray_dir(d, up, fov, Wc, Hc)).astype(np.float32)
        orc = np.concatenate(yvo.init_ray_dir(yvo.camera(d, d, up, fov, Wc, Hc)))
        assert mine.tobytes() == w.tobytes() and orc.tobytes() == w.tobytes(), c


@pytest.mark.gpu
@pytest.mark.parametrize("schedule", [0, 1, 2])
def test_cuda_reproduces_the_reference_frames(ref, schedule):
    r = yv.SVORenderer(0)
    r.SetOption("persistent", schedule)
    r.EnableHits(True)
    try:
        for scene in SCENES:
            svo = yv.SVOData().Load(os.path.join(HERE, scene + ".vox"))
            nodes = svo.nodes()
            r.SetScene(svo)
            r.SetResolution(W, H)
            for name, pos, d, up, fov in scenes.CAMERAS:
                key = "%s/%s" % (scene, name)
                r.SetViewPos(pos); r.SetViewDir(d); r.SetViewUp(up); r.SetFOV(fov)
                img = r.RenderFrame().copy()
                node, child, t = r.GetHits()
                hit = node != yvo.MISS_NODE
                assert (img == ref[key + "/rgba"]).all(), key
                assert (np.where(hit, t.view(np.uint32), 0) == ref[key + "/tbits"]).all(), key
                data = np.where(hit, nodes["child"][np.where(hit, node, 0), np.where(hit, child, 0)], 0)
                assert (data == ref[key + "/data"]).all(), key
    finally:
        r.close()


@pytest.mark.gpu
@pytest.mark.parametrize("layout", [0, 1], ids=["packed", "raw"])
def test_cuda_reproduces_the_reference_hit_ids(ref, layout):
    svo = _tagged()
    r = yv.SVORenderer(0)
    r.SetOption("layout", layout)
    r.EnableHits(True)
    r.EnableCounters(True)
    try:
        r.SetScene(svo)
        r.SetResolution(TW, TH)
        n = 0
        for name, pos, d, up, fov in scenes.CAMERAS:
            ids, tbits = ref["tagged/%s/ids" % name], ref["tagged/%s/tbits" % name]
            r.SetViewPos(pos); r.SetViewDir(d); r.SetViewUp(up); r.SetFOV(fov)
            r.RenderFrame()
            node, child, t = r.GetHits()
            hit = node != yvo.MISS_NODE
            assert ((ids != 0) | (tbits != 0) == hit).all(), name
            assert ((ids >> 3)[hit] == node[hit]).all() and ((ids & 7)[hit] == child[hit].astype(np.uint32)).all(), name
            assert (tbits[hit] == t.view(np.uint32)[hit]).all(), name
            n += int(hit.sum())
        assert n > 5000
    finally:
        r.close()


@pytest.mark.gpu
def test_cuda_node_visits_equal_the_spu_programs_fetch_count(ref):
    """FetchNode calls of the reference's SPU program (trace_spu.cpp:33) == node visits counted by the kernel."""
    r = yv.SVORenderer(0)
    r.EnableCounters(True)
    r.SetOption("cull", 0)             # count the reference traversal's node fetches (the default; stated because the count depends on it)
    try:
        for scene in SCENES:
            r.SetScene(yv.SVOData().Load(os.path.join(HERE, scene + ".vox")))
            r.SetResolution(W, H)
            for name, pos, d, up, fov in scenes.CAMERAS:
                r.SetViewPos(pos); r.SetViewDir(d); r.SetViewUp(up); r.SetFOV(fov)
                r.RenderFrame()
                visits = r.GetCounters()[0]
                assert int(np.asarray(visits, np.int64).sum()) == int(ref["%s/%s/spu_fetches" % (scene, name)]), (scene, name)
    finally:
        r.close()
