"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): hit node/child ids bit-exact; hit distance t within 1e-4 relative;
per-pixel RGB within 1 LSB. The kernel uses the same float32 operations in the same order as the
oracle, so the tests additionally report (and require) exact equality of t and RGBA."""
import os

import numpy as np
import pytest

import scenes
import yvo
import yoxel_voxel_b200 as yv

pytestmark = pytest.mark.gpu

T_RTOL = 1e-4      # north_star tolerance for the hit distance
RGB_LSB = 1        # north_star tolerance for colour


def _check(o, img, node, child, t, tag):
    assert (node == o["node"]).all(), "%s: hit node ids differ in %d pixels" % (tag, (node != o["node"]).sum())
    assert (child == o["child"]).all(), "%s: hit child ids differ" % tag
    assert np.allclose(t, o["t"], rtol=T_RTOL, atol=0), tag
    assert np.abs(img.astype(int) - o["rgba"].astype(int)).max() <= RGB_LSB, tag
    # stronger than required: same arithmetic => identical bits
    assert t.tobytes() == o["t"].tobytes(), "%s: t not bit-identical" % tag
    assert (img == o["rgba"]).all(), "%s: rgba not identical" % tag


def _render_gpu(r, cam_spec, W, H, sec=None, detail=0.0):
    name, pos, d, up, fov = cam_spec
    r.SetResolution(W, H)
    r.SetDetailCoef(detail)
    r.SetViewPos(pos); r.SetViewDir(d); r.SetViewUp(up); r.SetFOV(fov)
    if sec:
        r.SetSecondary(**sec)
    else:
        r.SetSecondary(0, 0)
    img = r.RenderFrame().copy()
    node, child, t = r.GetHits()
    return img, node, child, t


def _render_cpu(svo, cam_spec, W, H, sec=None, visits=False, detail=0.0):
    name, pos, d, up, fov = cam_spec
    return yvo.render(svo.nodes(), svo.GetRoot(), yvo.camera(pos, d, up, fov, W, H, detail_coef=detail),
                      sec=yvo.secondary(**sec) if sec else None, threads=8, want_visits=visits)


@pytest.fixture(scope="module")
def renderer():
    r = yv.SVORenderer(0)
    r.EnableHits(True)
    yield r
    r.close()


@pytest.mark.parametrize("persistent", [0, 1, 2], ids=["tiles", "persistent", "queue"])
@pytest.mark.parametrize("cam", scenes.CAMERAS, ids=[c[0] for c in scenes.CAMERAS])
def test_fractal10_primary(renderer, cam, persistent):
    """BASELINE config 1: depth-10 sphere fractal, 512x512 primary rays (+ the other cameras)."""
    svo = scenes.fractal(10)
    renderer.SetOption("persistent", persistent)
    renderer.SetScene(svo)
    img, node, child, t = _render_gpu(renderer, cam, 512, 512)
    _check(_render_cpu(svo, cam, 512, 512), img, node, child, t, cam[0])


@pytest.mark.parametrize("smem_nodes", [0, 1, 73, 585, 4681])
@pytest.mark.parametrize("persistent", [0, 1], ids=["tiles", "persistent"])
def test_shared_memory_staging_sizes(renderer, smem_nodes, persistent):
    svo = scenes.fractal(9)
    renderer.SetOption("persistent", persistent)
    renderer.SetOption("smem_nodes", smem_nodes)
    renderer.SetScene(svo)
    cam = scenes.CAMERAS[1]
    img, node, child, t = _render_gpu(renderer, cam, 320, 200)
    _check(_render_cpu(svo, cam, 320, 200), img, node, child, t, "smem%d" % smem_nodes)
    renderer.SetOption("smem_nodes", 0)


@pytest.mark.parametrize("stack", [0, 4], ids=["local", "ring4"])
@pytest.mark.parametrize("persistent", [0, 1], ids=["tiles", "persistent"])
def test_stack_variants(renderer, stack, persistent):
    """The traversal stack in local memory, shared memory, or a shared ring spilling to local memory."""
    svo = scenes.fractal(10)
    renderer.SetOption("persistent", persistent)
    renderer.SetOption("stack", stack)
    renderer.SetScene(svo)
    sec = dict(shadow=1, ao_samples=4, seed=1, light_pos=(0.6, 0.4, 1.2), voxel_size=1.0 / 1024, ao_max_t=0.05)
    for cam in (scenes.CAMERAS[1], scenes.CAMERAS[4]):
        img, node, child, t = _render_gpu(renderer, cam, 333, 250)
        _check(_render_cpu(svo, cam, 333, 250), img, node, child, t, "stack%d" % stack)
        img, node, child, t = _render_gpu(renderer, cam, 333, 250, sec)
        _check(_render_cpu(svo, cam, 333, 250, sec), img, node, child, t, "stack%d/sec" % stack)
    renderer.SetOption("stack", 0)
    renderer.SetSecondary(0, 0)


@pytest.mark.parametrize("size", [(1, 1), (7, 5), (37, 23), (130, 67), (1024, 768)])
@pytest.mark.parametrize("persistent", [0, 1, 2], ids=["tiles", "persistent", "queue"])
def test_ragged_resolutions(renderer, size, persistent):
    """Partial tiles at the right / bottom edges; cell/main.cpp's 1024x768."""
    svo = scenes.fractal(9)
    renderer.SetOption("persistent", persistent)
    renderer.SetScene(svo)
    cam = scenes.CAMERAS[1]
    img, node, child, t = _render_gpu(renderer, cam, *size)
    assert img.shape == (size[1], size[0], 4)
    _check(_render_cpu(svo, cam, *size), img, node, child, t, "res%dx%d" % size)


@pytest.mark.parametrize("persistent", [0, 1], ids=["tiles", "persistent"])
@pytest.mark.parametrize("sec", [
    dict(shadow=1, ao_samples=0, seed=1, light_pos=(0.6, 0.4, 1.2), voxel_size=1.0 / 1024, ao_max_t=0.05),
    dict(shadow=0, ao_samples=4, seed=1, light_pos=(0.6, 0.4, 1.2), voxel_size=1.0 / 1024, ao_max_t=0.05),
    dict(shadow=1, ao_samples=4, seed=1, light_pos=(0.6, 0.4, 1.2), voxel_size=1.0 / 1024, ao_max_t=0.05),
    dict(shadow=1, ao_samples=16, seed=99, light_pos=(0.1, 0.9, 0.2), voxel_size=2.0 / 1024, ao_max_t=0.2),
], ids=["shadow", "ao4", "shadow+ao4", "shadow+ao16"])
def test_secondary_rays(renderer, sec, persistent):
    """BASELINE config 4 (at depth 10): primary + shadow + AO rays; stage machine and pooled-AO kernel."""
    svo = scenes.fractal(10)
    renderer.SetOption("persistent", persistent)
    renderer.SetScene(svo)
    renderer.EnableCounters(True)
    for cam in (scenes.CAMERAS[1], scenes.CAMERAS[4]):
        o = _render_cpu(svo, cam, 397, 301, sec, visits=True)
        for sec_queue, cull in ((0, 0), (1, 0), (0, 1)):
            renderer.SetOption("sec_queue", sec_queue)
            renderer.SetOption("cull", cull)
            img, node, child, t = _render_gpu(renderer, cam, 397, 301, sec)
            _check(o, img, node, child, t, "sec/q%d/cull%d" % (sec_queue, cull))
            visits, pops = renderer.GetCounters()
            if cull:                                  # octant culling: same pixels from fewer node fetches
                assert (visits <= o["visits"]).all() and visits.sum() < 0.85 * o["visits"].sum()
            else:
                assert (visits == o["visits"]).all(), "node visits differ (sec_queue=%d)" % sec_queue
    renderer.EnableCounters(False)
    renderer.SetOption("cull", 0)
    renderer.SetOption("sec_queue", 0)
    renderer.SetSecondary(0, 0)


@pytest.mark.parametrize("persistent", [0, 1, 2], ids=["tiles", "persistent", "queue"])
def test_other_scenes(renderer, persistent):
    renderer.SetOption("persistent", persistent)
    for svo in (scenes.single_sphere(6), scenes.dense_random(5, 0.03)[0], scenes.dense_random(4, 0.08)[0],
                yv.SVOData.IsoVolume(8, threads=8)):
        renderer.SetScene(svo)
        for cam in scenes.CAMERAS:
            img, node, child, t = _render_gpu(renderer, cam, 200, 160)
            _check(_render_cpu(svo, cam, 200, 160), img, node, child, t, cam[0])


@pytest.mark.parametrize("persistent", [0, 1], ids=["tiles", "persistent"])
@pytest.mark.parametrize("coef", [1.0, 4.0, 12.0])
def test_lod_detail_coef(renderer, coef, persistent):
    """SVORenderer::SetDetailCoef (demo/SVORenderer.h:25-26): LOD-terminated traversal, primary and secondary."""
    svo = scenes.fractal(11)
    renderer.SetOption("schedule", persistent)
    renderer.SetScene(svo)
    sec = dict(shadow=1, ao_samples=2, seed=3, light_pos=(0.6, 0.4, 1.2), voxel_size=1.0 / 2048, ao_max_t=0.05)
    n_lod = 0
    for cam in (scenes.CAMERAS[1], scenes.CAMERAS[2], scenes.CAMERAS[4]):
        img, node, child, t = _render_gpu(renderer, cam, 640, 400, detail=coef)
        o = _render_cpu(svo, cam, 640, 400, detail=coef)
        _check(o, img, node, child, t, "lod%g" % coef)
        n_lod += int(((child == -1) & (node != yvo.MISS_NODE)).sum())
        img, node, child, t = _render_gpu(renderer, cam, 320, 200, sec, detail=coef)
        _check(_render_cpu(svo, cam, 320, 200, sec, detail=coef), img, node, child, t, "lod%g/sec" % coef)
    assert renderer.GetDetailCoef() == coef
    if coef >= 4.0:
        assert n_lod > 1000
    renderer.SetDetailCoef(0.0)
    renderer.SetSecondary(0, 0)
    renderer.SetOption("schedule", 0)


def test_empty_scene_and_no_scene():
    r = yv.SVORenderer(0)
    assert r.RenderFrame() is None                                  # reference: NULL (ppu_renderer.cpp:78-79)
    assert r.GetResolution() == (640, 480) and r.GetFOV() == 70.0   # renderer_base.h:25
    empty = yv.SVOData.FromNodes(yv.EMPTY_NODE, np.zeros(0, yv.NODE_DTYPE))
    r.SetScene(empty)
    r.SetResolution(64, 64)
    r.SetViewPos((0.5, 0.5, -1)); r.SetViewDir((0, 0, 1))
    img = r.RenderFrame()
    assert img.shape == (64, 64, 4) and (img == 0).all()
    r.close()


def test_row_bands_compose_to_the_full_frame(renderer):
    """Multi-GPU partition: bands rendered separately are byte-identical to the full frame."""
    svo = scenes.fractal(10)
    renderer.SetOption("persistent", 0)
    renderer.SetScene(svo)
    cam = scenes.CAMERAS[1]
    full, *_ = _render_gpu(renderer, cam, 640, 363)
    out = np.zeros_like(full)
    for (y0, y1) in [(0, 91), (91, 182), (182, 300), (300, 363)]:
        for persistent in (0, 1, 2):
            renderer.SetOption("persistent", persistent)
            renderer.SetRows(y0, y1)
            band = renderer.RenderFrame()
            assert (band[y0:y1] == full[y0:y1]).all()
        out[y0:y1] = band[y0:y1]
    assert (out == full).all()
    renderer.SetResolution(640, 363)      # resets the band


@pytest.mark.parametrize("schedule", [0, 1, 2], ids=["tiles", "persistent", "queue"])
@pytest.mark.parametrize("world,band", [(2, 16), (3, 32), (8, 32)])
def test_interleaved_blocks_compose_to_the_full_frame(renderer, schedule, world, band):
    """Round-robin row blocks (yv_set_interleave) over `world` ranks reproduce the 1-GPU frame byte for byte."""
    from yoxel_voxel_b200 import multigpu
    svo = scenes.fractal(10)
    renderer.SetOption("schedule", 0)
    renderer.SetScene(svo)
    cam = scenes.CAMERAS[1]
    W, H = 640, 363
    full, *_ = _render_gpu(renderer, cam, W, H)
    renderer.SetOption("schedule", schedule)
    out = np.zeros_like(full)
    for rank in range(world):
        renderer.SetInterleave(band, world, rank)
        part = renderer.RenderFrame()
        rows = multigpu.interleaved_rows(rank, world, H, band)
        out[rows] = part[rows]
    renderer.SetInterleave(16, 1, 0)
    renderer.SetOption("schedule", 0)
    assert (out == full).all()


def test_trace_rays_matches_oracle(renderer):
    """DynamicSVO::TraceRay (ore/src/main.cpp:125) batched on the device."""
    svo = scenes.fractal(9)
    renderer.SetScene(svo)
    nodes = svo.nodes()
    rng = np.random.RandomState(5)
    pos = rng.rand(4000, 3).astype(np.float32) * 1.6 - 0.3
    dirs = rng.randn(4000, 3).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dirs[::7, 0] = 0.0                      # exercise AdjustDir
    node, child, t = renderer.TraceRays(pos, dirs)
    nh = 0
    for i in range(len(pos)):
        hit, n, c, tt = yvo.trace_ray(nodes, svo.GetRoot(), pos[i], dirs[i])
        assert (node[i], child[i]) == (n, c), i
        assert np.float32(tt).tobytes() == t[i].tobytes()
        nh += hit
    assert nh > 100


def test_counters_equal_oracle_visits(renderer):
    """With the octant culling off the kernel dereferences exactly the nodes the oracle does (basis of the roofline's
    V-bar); with it on (option "cull") it produces the same hits from fewer dereferences."""
    svo = scenes.fractal(10)
    renderer.SetScene(svo)
    renderer.EnableCounters(True)
    cam = scenes.CAMERAS[1]
    o = _render_cpu(svo, cam, 256, 256, visits=True)
    for persistent in (0, 1, 2):
        renderer.SetOption("persistent", persistent)
        renderer.SetOption("cull", 0)
        _render_gpu(renderer, cam, 256, 256)
        visits, pops = renderer.GetCounters()
        assert (visits == o["visits"]).all()
        assert pops.sum() > 0
        renderer.SetOption("cull", 1)
        img, node, child, t = _render_gpu(renderer, cam, 256, 256)
        _check(o, img, node, child, t, "cull/%d" % persistent)
        culled, cpops = renderer.GetCounters()
        assert (culled <= visits).all()
        if persistent != 2:                           # (the queue schedule has no culling variant)
            assert culled.sum() < 0.8 * visits.sum() and cpops.sum() < pops.sum()
        renderer.SetOption("cull", 0)
    renderer.SetOption("persistent", 0)
    renderer.EnableCounters(False)


def test_render_into_a_registered_host_frame(renderer):
    """yv_host_register: a caller-owned (here: shared-memory) host frame as the render target; two interleaved halves
    written by two launches compose the frame the renderer itself returns."""
    from yoxel_voxel_b200 import multigpu
    svo = scenes.fractal(10)
    renderer.SetOption("persistent", 0)
    renderer.SetScene(svo)
    cam = scenes.CAMERAS[1]
    W, H = 640, 400
    img, *_ = _render_gpu(renderer, cam, W, H)
    shared = multigpu.SharedHostFrame(None, 0, 1, 0, W * H * 4, "pytest%d" % os.getpid())
    try:
        assert shared.registered and not os.path.exists(shared.path)
        shared.array[:] = 0x7f
        for phase in (0, 1):                          # what ranks 0 and 1 of a two-GPU frame would each store
            renderer.SetInterleave(32, 2, phase)
            renderer.Render(shared.ptr)
        renderer.SetInterleave(32, 1, 0)
        assert (shared.array.reshape(H, W, 4) == img).all()
    finally:
        renderer.SetInterleave(32, 1, 0)
        shared.close()


def test_device_pointer_render_and_timing(renderer):
    """SVORenderer::Render(void* d_dstBuf) (demo/SVORenderer.h:36) into a caller-owned device buffer."""
    import torch
    svo = scenes.fractal(10)
    renderer.SetOption("persistent", 0)
    renderer.SetScene(svo)
    cam = scenes.CAMERAS[1]
    img, *_ = _render_gpu(renderer, cam, 512, 384)
    buf = torch.zeros(384, 512, 4, dtype=torch.uint8, device="cuda:0")
    torch.cuda.synchronize()
    renderer.Render(buf.data_ptr())
    assert (buf.cpu().numpy() == img).all()
    assert 0 < renderer.LastFrameMs() < 1000 and renderer.LastFrameLaunches() == 1
    # RenderFrame's three ways of getting the frame to the host: the kernel stores into the pinned host frame
    # (default), 8 row chunks pipelined with their D2H copies, one launch + one copy; the pixels are the same
    assert renderer.GetOption("zero_copy") == 1
    z = renderer.RenderFrame().copy()
    assert renderer.LastFrameLaunches() == 1 and 0 < renderer.LastFrameMs() < 1000
    renderer.SetOption("zero_copy", 0)
    renderer.SetOption("pipeline", 8)
    a = renderer.RenderFrame().copy()
    assert renderer.LastFrameLaunches() == 8 and 0 < renderer.LastFrameMs() < 1000
    renderer.SetOption("pipeline", 0)
    b = renderer.RenderFrame().copy()
    assert renderer.LastFrameLaunches() == 1
    renderer.SetOption("pipeline", 4)
    renderer.SetOption("pipeline_taper", 60)
    c = renderer.RenderFrame().copy()
    renderer.SetOption("pipeline_taper", 100)
    renderer.SetOption("zero_copy", 1)
    assert (z == img).all() and (a == img).all() and (b == img).all() and (c == img).all()
    # on torch's current stream
    renderer.SetStream(torch.cuda.current_stream().cuda_stream)
    buf.zero_()
    renderer.Render(buf.data_ptr(), sync=False)
    torch.cuda.synchronize()
    assert (buf.cpu().numpy() == img).all()
    renderer.SetStream(0)


def test_full_size_config2_depth12_1080p():
    """BASELINE config 2 at full size: depth-12 sphere fractal, 1920x1080, primary + Lambert.
    The threaded oracle renders the whole frame in seconds, so the comparison is exhaustive."""
    svo = scenes.fractal(12)
    r = yv.SVORenderer(0)
    r.EnableHits(True)
    r.SetScene(svo)
    cam = scenes.CAMERAS[1]
    o = _render_cpu(svo, cam, 1920, 1080)
    assert (o["node"] != yvo.MISS_NODE).mean() > 0.3
    for persistent in (0, 1, 2):
        r.SetOption("persistent", persistent)
        img, node, child, t = _render_gpu(r, cam, 1920, 1080)
        _check(o, img, node, child, t, "config2/%d" % persistent)
    # idempotence: a second frame is byte-identical
    img2, *_ = _render_gpu(r, cam, 1920, 1080)
    assert (img2 == img).all()
    r.close()


def test_cpp_adapter_renders_like_cell_main(tmp_path):
    """The C++ ISVORenderer adapter driven like cell/main.cpp:21-40 (1024x768, eye (0.5,0.5,0.3))."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-s", "-C", os.path.join(root, "tools")])
    svo = scenes.fractal(9)
    vox = str(tmp_path / "scene.vox")
    svo.Save(vox)
    out = str(tmp_path / "frame.ppm")
    p = subprocess.run([os.path.join(root, "tools", "render_main"), vox, out, "1024", "768", "-1", "-1", "1.5"],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "time:" in p.stdout
    raw = open(out, "rb").read()
    hdr = b"P6\n1024 768\n255\n"
    assert raw.startswith(hdr)
    img = np.frombuffer(raw[len(hdr):], np.uint8).reshape(768, 1024, 3)
    o = _render_cpu(svo, ("m", (0.5, 0.5, 0.3), (-1, -1, 1.5), (0, 0, 1), 70.0), 1024, 768)
    assert (img == o["rgba"][:, :, :3]).all()


def test_dump_trace_data(renderer, tmp_path):
    """SVORenderer::DumpTraceData (demo/SVORenderer.cpp:158-192): .dist / .color / .normal of the last frame."""
    svo = scenes.fractal(9)
    renderer.SetOption("schedule", 0)
    renderer.SetScene(svo)
    cam = scenes.CAMERAS[1]
    img, node, child, t = _render_gpu(renderer, cam, 200, 120, detail=8.0)
    base = str(tmp_path / "dmp_0")
    renderer.DumpTraceData(base)
    dist = np.fromfile(base + "_200x120.dist", np.float32).reshape(120, 200)
    color = np.fromfile(base + "_200x120.color", np.uint8).reshape(120, 200, 4)
    normal = np.fromfile(base + "_200x120.normal", np.float32).reshape(120, 200, 3)
    assert dist.tobytes() == t.tobytes()
    hit = node != yvo.MISS_NODE
    assert (color[~hit] == 0).all() and (normal[~hit] == 0).all() and (color[hit][:, 3] == 255).all()
    nodes = svo.nodes()
    ys, xs = np.nonzero(hit)
    for y, x in list(zip(ys, xs))[::37]:
        d = nodes["data"][node[y, x]] if child[y, x] < 0 else nodes["child"][node[y, x], child[y, x]]
        assert np.allclose(normal[y, x], yvo.unpack_normal(int(d)), atol=1e-6)
        r5, g6, b5 = (d >> 11) & 31, (d >> 5) & 63, d & 31
        assert tuple(color[y, x][:3]) == ((r5 << 3) | (r5 >> 2), (g6 << 2) | (g6 >> 4), (b5 << 3) | (b5 >> 2))
    renderer.SetDetailCoef(0.0)


@pytest.mark.parametrize("schedule", [0, 2], ids=["tiles", "queue"])
def test_phong_lights_and_show_normals(renderer, schedule):
    """SetLigth / SetShowNormals (demo/SVORenderer.h:31-34): ShadeSimple's point-light Phong model."""
    svo = scenes.fractal(10)
    renderer.SetOption("schedule", schedule)
    renderer.SetScene(svo)
    cam = scenes.CAMERAS[1]
    lights = [dict(pos=cam[1], diffuse=(0.7, 0.7, 0.7), specular=(0.3, 0.3, 0.3), attenuation=(1, 0, 0.5)),       # Demo.cpp:141-147
              dict(pos=(0.45, 0.4, 0.55), diffuse=(1, 0.8, 0.6), specular=(0.3, 0.3, 0.3), attenuation=(1, 10, 400))]   # Demo.cpp:160-165
    for i, lt in enumerate(lights):
        renderer.SetLigth(i, yv.LightParams(True, lt["pos"], lt["diffuse"], lt["specular"], lt["attenuation"]))
    img, node, child, t = _render_gpu(renderer, cam, 480, 320)
    o = yvo.render(svo.nodes(), svo.GetRoot(), yvo.camera(cam[1], cam[2], cam[3], cam[4], 480, 320, lights=lights), threads=8)
    _check(o, img, node, child, t, "phong")
    lambert = _render_cpu(svo, cam, 480, 320)
    assert (o["rgba"] != lambert["rgba"]).any()
    renderer.SetShowNormals(True)
    assert renderer.GetShowNormals()
    img, node, child, t = _render_gpu(renderer, cam, 480, 320)
    o = yvo.render(svo.nodes(), svo.GetRoot(), yvo.camera(cam[1], cam[2], cam[3], cam[4], 480, 320, show_normals=True), threads=8)
    _check(o, img, node, child, t, "normals")
    renderer.SetShowNormals(False)
    for i in range(2):
        renderer.SetLigth(i, yv.LightParams(False))
    renderer.SetOption("schedule", 0)
    img, *_ = _render_gpu(renderer, cam, 480, 320)
    assert (img == lambert["rgba"]).all()


def test_8k_frame_sampled_bands():
    """BASELINE config 5 resolution (7680x4320): the whole frame is rendered on the GPU, the oracle checks
    bands of rows spread over the image (an exhaustive 33 M-ray CPU frame is not needed to catch an
    addressing or partition bug at this size)."""
    svo = scenes.fractal(10)
    r = yv.SVORenderer(0)
    r.EnableHits(True)
    r.SetScene(svo)
    cam = scenes.CAMERAS[1]
    W, H = 7680, 4320
    img, node, child, t = _render_gpu(r, cam, W, H)
    name, pos, d, up, fov = cam
    for y0 in (0, 1077, 2160, 3333, H - 24):
        o = yvo.render(svo.nodes(), svo.GetRoot(), yvo.camera(pos, d, up, fov, W, H), threads=8, rows=(y0, y0 + 24))
        sl = slice(y0, y0 + 24)
        assert (node[sl] == o["node"][sl]).all() and (child[sl] == o["child"][sl]).all()
        assert t[sl].tobytes() == o["t"][sl].tobytes() and (img[sl] == o["rgba"][sl]).all()
    # interleaved 8-way partition of the same frame reproduces it
    from yoxel_voxel_b200 import multigpu
    out = np.zeros_like(img)
    for rank in range(8):
        r.SetInterleave(32, 8, rank)
        part = r.RenderFrame()
        rows = multigpu.interleaved_rows(rank, 8, H, 32)
        out[rows] = part[rows]
    assert (out == img).all()
    r.close()


def test_cuda_renderer_mirror_of_trace_cuda_py():
    """CudaRenderer (trace_cuda.py:16-118): constructor, updateScene, setLightPos, render -> stats, getImage."""
    cr = yv.CudaRenderer(res=(320, 240))
    cr.updateScene(scenes.fractal(9))
    assert cr.getViewSize() == (320, 240) and cr.detailCoef == 10.0
    stat = cr.render((0.5, 0.5, 0.3), (-1, -1, 1.5))
    assert "gpu time:" in stat and "eye trace time:" in stat and "detailCoef: 10.0" in stat
    img = cr.getImage()
    assert img.shape == (240, 320, 3) and img.any()
    cr.detailCoef = 1.0                         # coarser LOD -> a different picture
    cr.render((0.5, 0.5, 0.3), (-1, -1, 1.5))
    assert (cr.getImage() != img).any()
