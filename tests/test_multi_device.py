"""One renderer handle over several GPUs (yv_renderer_create_multi / _group), and frames in flight.

Reference: SPURenderer drives every SPE inside one RenderFrame() — worker i takes the blocks b with
b % threadNum == i (cell/spu_renderer.cpp:65-90, cell/spu/trace_spu.cpp:164) and all of them DMA into the one colour
buffer (trace_spu.cpp:171-176). SURVEY §4 item 5 / §8(e): the N-GPU frame must be byte-identical to the 1-GPU frame.

On a single-GPU box the group is built from the same ordinal listed several times (its members share the GPU), so the
whole machinery — partition, stream joins, shared host frame, staged copies, merged hit planes — runs in the driver's
1-GPU `pytest -m gpu` as well; with two or more GPUs the same tests also run over distinct devices (NVLink replication,
peer stores)."""
import os
import subprocess

import numpy as np
import pytest

import ctypes as C

import scenes
import yvo
import yoxel_voxel_b200 as yv
from yoxel_voxel_b200 import api as yvapi

pytestmark = pytest.mark.gpu

SEC = dict(shadow=1, ao_samples=4, seed=1, light_pos=(0.6, 0.4, 1.2), voxel_size=1.0 / 1024, ao_max_t=0.05)


def _device_lists():
    n = yv.device_count()
    lists = [[0, 0], [0, 0, 0]]
    if n >= 2:
        lists.append(list(range(min(n, 8))))
    if n >= 3:
        lists.append([1, 0])                      # leader is not device 0
    return lists


def _setup(r, svo, cam, W, H):
    _, pos, d, up, fov = cam
    r.SetScene(svo)
    r.SetResolution(W, H)
    r.SetViewPos(pos); r.SetViewDir(d); r.SetViewUp(up); r.SetFOV(fov)


@pytest.fixture(scope="module")
def single():
    r = yv.SVORenderer(0)
    r.EnableHits(True)
    yield r
    r.close()


@pytest.mark.parametrize("devs", _device_lists(), ids=lambda d: "gpus" + "".join(map(str, d)))
@pytest.mark.parametrize("size", [(640, 480), (333, 250), (1024, 768)])
def test_group_frame_equals_single_gpu_frame_and_oracle(single, devs, size):
    svo = scenes.fractal(10)
    cam = scenes.CAMERAS[1]
    W, H = size
    _setup(single, svo, cam, W, H)
    single.SetSecondary(0, 0)
    ref = single.RenderFrame().copy()
    rn, rc, rt = single.GetHits()
    o = yvo.render(svo.nodes(), svo.GetRoot(), yvo.camera(cam[1], cam[2], cam[3], cam[4], W, H), threads=8)
    assert (ref == o["rgba"]).all()
    g = yv.SVORenderer(devices=devs)
    try:
        assert g.DeviceCount() == len(devs) and g.Devices() == devs
        g.EnableHits(True)
        _setup(g, svo, cam, W, H)
        for zero_copy in (1, 0):                          # kernels store into the host frame / copy engines move the rows
            g.SetOption("zero_copy", zero_copy)
            for mode, rows in (("interleaved", 32), ("interleaved", 16), ("bands", 32)):
                g.SetPartition(mode, rows)
                img = g.RenderFrame().copy()
                assert (img == ref).all(), (devs, zero_copy, mode, rows)
                node, child, t = g.GetHits()
                assert (node == rn).all() and (child == rc).all() and t.tobytes() == rt.tobytes()
        ms = g.MemberFrameMs()
        assert len(ms) == len(devs) and all(m > 0 for m in ms) and g.LastFrameMs() > 0
        assert g.GetOption("group_threads") == 1          # every peer's share is launched by its own host thread ...
        g.SetOption("group_threads", 0)                   # ... or all of them by one loop on the calling thread
        assert (g.RenderFrame() == ref).all()
        g.SetOption("group_threads", 1)
        assert (g.RenderFrame() == ref).all()
    finally:
        g.close()


@pytest.mark.parametrize("devs", _device_lists()[:3], ids=lambda d: "gpus" + "".join(map(str, d)))
def test_group_secondary_lod_phong(single, devs):
    """Secondary rays and LOD are one-pass frames (direct stores); Phong / show-normals re-read the frame, so every
    member shades its rows in its own HBM and the rows are then copied into the frame."""
    svo = scenes.fractal(10)
    cam = scenes.CAMERAS[4]
    W, H = 397, 301
    g = yv.SVORenderer(devices=devs)
    try:
        _setup(g, svo, cam, W, H)
        _setup(single, svo, cam, W, H)
        for r in (single, g):
            r.SetSecondary(**SEC)
        assert (g.RenderFrame() == single.RenderFrame()).all()
        o = yvo.render(svo.nodes(), svo.GetRoot(), yvo.camera(cam[1], cam[2], cam[3], cam[4], W, H), sec=yvo.secondary(**SEC), threads=8)
        assert (g.RenderFrame() == o["rgba"]).all()
        for r in (single, g):
            r.SetSecondary(0, 0); r.SetDetailCoef(8.0)
        assert (g.RenderFrame() == single.RenderFrame()).all()
        for r in (single, g):
            r.SetDetailCoef(0.0)
            r.SetLigth(0, yv.LightParams(True, (0.9, 0.2, 0.8), (0.8, 0.7, 0.6), (0.4, 0.4, 0.4), (1, 0.2, 0.5)))
        a, b = g.RenderFrame().copy(), single.RenderFrame().copy()
        assert (a == b).all() and (a[..., 3] == 255).sum() > 1000
        for r in (single, g):
            r.SetLigth(0, yv.LightParams(False)); r.SetShowNormals(True)
        assert (g.RenderFrame() == single.RenderFrame()).all()
        g.SetSSNA(True)
        with pytest.raises(yv.YVError):                   # BlurZ taps reach across the members' rows
            g.RenderFrame()
    finally:
        for r in (single,):
            r.SetShowNormals(False); r.SetLigth(0, yv.LightParams(False)); r.SetSecondary(0, 0); r.SetDetailCoef(0.0)
        g.close()


@pytest.mark.parametrize("devs", _device_lists()[:3], ids=lambda d: "gpus" + "".join(map(str, d)))
def test_group_renders_into_a_device_pointer(single, devs):
    """SVORenderer::Render(void* d_dstBuf) (demo/SVORenderer.h:36) on a group: every member stores into the leader's HBM."""
    import torch
    svo = scenes.fractal(9)
    cam = scenes.CAMERAS[1]
    W, H = 500, 300
    _setup(single, svo, cam, W, H)
    single.SetSecondary(0, 0)
    ref = single.RenderFrame().copy()
    g = yv.SVORenderer(devices=devs)
    try:
        _setup(g, svo, cam, W, H)
        dst = torch.zeros(H, W, 4, dtype=torch.uint8, device="cuda:%d" % g.device)
        g.Render(dst.data_ptr(), sync=True)
        assert (dst.cpu().numpy() == ref).all()
        g.SetShowNormals(True); single.SetShowNormals(True)           # second pass: assembled with peer copies
        dst.zero_()
        g.Render(dst.data_ptr(), sync=True)
        assert (dst.cpu().numpy() == single.RenderFrame()).all()
    finally:
        single.SetShowNormals(False)
        g.close()


@pytest.mark.parametrize("devs", [[0]] + _device_lists()[:3], ids=lambda d: "gpus" + "".join(map(str, d)))
@pytest.mark.parametrize("zero_copy", [1, 0])
def test_frames_in_flight(single, devs, zero_copy):
    """yv_render_frame_async / yv_wait_frame: a flythrough with two (and three) frames outstanding delivers, ticket by
    ticket, the frames the synchronous call renders for the same cameras."""
    svo = scenes.fractal(9)
    W, H = 320, 208
    cams = [((0.5 + 0.02 * i, 0.5 - 0.01 * i, 0.3 + 0.01 * i), (-1, -1 + 0.1 * i, 1.5)) for i in range(7)]
    single.SetScene(svo); single.SetResolution(W, H); single.SetViewUp((0, 0, 1)); single.SetFOV(70.0); single.SetSecondary(0, 0)
    want = []
    for pos, d in cams:
        single.SetViewPos(pos); single.SetViewDir(d)
        want.append(single.RenderFrame().copy())
    assert not (want[0] == want[3]).all()
    g = yv.SVORenderer(devices=devs)
    try:
        g.SetScene(svo); g.SetResolution(W, H); g.SetViewUp((0, 0, 1)); g.SetFOV(70.0)
        g.SetOption("zero_copy", zero_copy)
        for slots in (2, 3):
            g.SetOption("slots", slots)
            pending, got = [], []
            for pos, d in cams:
                if len(pending) == slots:
                    got.append(g.WaitFrame(pending.pop(0)).copy())
                g.SetViewPos(pos); g.SetViewDir(d)
                pending.append(g.RenderFrameAsync())
            assert len(pending) == slots
            with pytest.raises(yv.YVError):               # every slot is busy
                g.RenderFrameAsync()
            while pending:
                got.append(g.WaitFrame(pending.pop(0)).copy())
            assert len(got) == len(want)
            for i, (a, b) in enumerate(zip(got, want)):
                assert (a == b).all(), (slots, i)
            assert g.LastFrameMs() > 0
        with pytest.raises(yv.YVError):
            g.WaitFrame(12345)
        # a caller-owned target: page-locked host memory registered with the leader's GPU
        host = np.zeros((H, W, 4), np.uint8)
        d_ptr = C.c_void_p()
        yvapi._check(yv.lib().yv_host_register(g.device, C.c_void_p(host.ctypes.data), host.nbytes, C.byref(d_ptr)))
        try:
            g.SetViewPos(cams[2][0]); g.SetViewDir(cams[2][1])
            tk = g.RenderFrameAsync(d_ptr.value)
            assert g.WaitFrame(tk, as_array=False) == d_ptr.value
            assert (host == want[2]).all()
        finally:
            yv.lib().yv_host_unregister(C.c_void_p(host.ctypes.data))
    finally:
        g.close()


def test_scene_replication_over_peer_copies():
    """The packed pool is uploaded and re-packed once and copied to the other GPUs (yv_svo_replicate; a group does it
    by itself): the replica equals the original bit for bit."""
    if yv.device_count() < 2:
        pytest.skip("needs two GPUs")
    svo = yv.SVOData.SphereFractal(10, threads=8)
    svo.Upload(0)
    svo.Replicate(0, 1)
    a, b = svo.device_packed(0), svo.device_packed(1)
    assert a[0].shape[0] > 100000
    for x, y in zip(a, b):
        assert x.tobytes() == y.tobytes()
    g = yv.SVORenderer(devices=[0, 1])
    try:
        fresh = yv.SVOData.SphereFractal(9, threads=8)
        _setup(g, fresh, scenes.CAMERAS[1], 320, 200)
        g.RenderFrame()
        ms, nbytes = g.ReplicateStats()
        recs, leaves, _ = fresh.device_packed(1)
        assert nbytes >= recs.nbytes + leaves.nbytes and ms > 0
    finally:
        g.close()


def test_scene_reload_and_free_with_renderers_bound(tmp_path):
    """SVOData::Load on a scene a renderer already holds reloads in place (cell/svodata.h:31-50; SetScene keeps the
    pointer, renderer_base.h:28); freeing a scene un-sets it on its renderers (RenderFrame -> NULL) instead of leaving
    them a dangling pointer."""
    a, b = scenes.fractal(8), scenes.single_sphere(6)
    fa, fb = str(tmp_path / "a.vox"), str(tmp_path / "b.vox")
    a.Save(fa); b.Save(fb)
    cam = scenes.CAMERAS[2]
    r = yv.SVORenderer(0)
    try:
        svo = yv.SVOData().Load(fa)
        _setup(r, svo, cam, 256, 192)
        ia = r.RenderFrame().copy()
        svo.Load(fb)                                      # same handle, new pool
        ib = r.RenderFrame().copy()
        oa = yvo.render(a.nodes(), a.GetRoot(), yvo.camera(cam[1], cam[2], cam[3], cam[4], 256, 192), threads=4)
        ob = yvo.render(b.nodes(), b.GetRoot(), yvo.camera(cam[1], cam[2], cam[3], cam[4], 256, 192), threads=4)
        assert (ia == oa["rgba"]).all() and (ib == ob["rgba"]).all() and not (ia == ib).all()
        svo._release()
        assert r.RenderFrame() is None
    finally:
        r.close()


def test_raw_layout_survives_deep_and_cyclic_pools():
    """The raw (reference-layout) pool is traversed as handed in; a pool deeper than the traversal stack or a cyclic
    one must neither write past the stack nor hang (the kernel bounds the descent). The frame of a legal pool is
    unaffected (tests/test_dynamic_svo.py covers parity of the raw layout)."""
    leaf = yv.pack_voxdata(200, 100, 50, 0, 0, 1)
    # a chain 40 levels deep: node i -> child 0 = node i+1, the last one holds a leaf
    n = 40
    nodes = np.zeros(n, yv.NODE_DTYPE)
    nodes["child"][:] = yv.EMPTY_NODE
    nodes["flags"] = 0xFF << 8
    for i in range(n - 1):
        nodes[i]["child"][0] = i + 1
        nodes[i]["flags"] = 0xFE << 8
    nodes[n - 1]["child"][0] = leaf
    nodes[n - 1]["flags"] = 1 | (0xFE << 8)
    # and a cycle with branching: two nodes that are each other's (and their own) children
    cyc = np.zeros(2, yv.NODE_DTYPE)
    cyc["flags"] = 0
    cyc[0]["child"][:] = [1, 0, 1, 0, 1, 0, 1, 0]
    cyc[1]["child"][:] = [0, 1, 0, 1, 0, 1, 0, 1]
    r = yv.SVORenderer(0)
    try:
        r.SetOption("layout", 1)
        for pool in (nodes, cyc):
            svo = yv.SVOData.FromNodes(0, pool)
            _setup(r, svo, scenes.CAMERAS[2], 128, 96)
            img = r.RenderFrame()                         # terminates; nothing to compare with
            assert img is not None and img.shape == (96, 128, 4)
            node, child, t = r.TraceRays([(0.01, 0.02, -0.5)], [(0.001, 0.001, 1.0)])
            assert node.shape == (1,)
        r.SetOption("layout", 0)
        for pool, what in ((nodes, "deeper"), (cyc, "cycl")):
            svo = yv.SVOData.FromNodes(0, pool)
            r.SetScene(svo)
            with pytest.raises(yv.YVError) as e:          # the packed layout refuses them with a clean error
                r.RenderFrame()
            assert e.value.code == -3, str(e.value)
    finally:
        r.close()


# ---- the reference's own driver, unmodified, on every GPU of the box ------------------------------------------------
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "cell_main_b200")


@pytest.mark.parametrize("env", [{"YV_B200_DEVICES": "1"}, {}, {"YV_B200_DEVICE_LIST": "0,0,0,0"}],
                         ids=["first-gpu", "all-gpus", "four-members"])
def test_reference_main_on_every_gpu(tmp_path, env):
    """cell/main.cpp (unmodified) asks for CreateSPURenderer(), which on the Cell takes every usable SPE
    (cell/spu_renderer.cpp:73); the B200 binding takes every GPU. The frame it writes is the same whatever the
    number of GPUs, and equals the oracle's."""
    import test_in_tree_binding as tb
    svo = tb._scene()
    w, h, body = tb._run(tmp_path, svo, BIN, env)
    assert (w, h) == (tb.W, tb.H) and len(body) == tb.W * tb.H * 4
    img = np.frombuffer(body, np.uint8).reshape(tb.H, tb.W, 4)
    o = yvo.render(svo.nodes(), svo.GetRoot(), yvo.camera(tb.POS, tb.DIR, (0, 0, 1), 70.0, tb.W, tb.H), threads=8)
    assert (img == o["rgba"]).all()
