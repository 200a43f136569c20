"""SSNA — screen-space normal approximation (SetSSNA, demo/SVORenderer.h:28-29; BlurZ x5 + ShadeSimple,
demo/SVORenderer.cpp:55-79,126-147; normal reconstruction after demo/dumps/ztools.py:23-44).

The kernel bodies are absent from the snapshot, so include/yv_format.h "SSNA" is the written spec. The CPU tests pin
the oracle's statement of it to closed-form cases (taps, planes, a sphere); the GPU tests require the CUDA passes
to reproduce the oracle's frame bit for bit."""
import os

import numpy as np
import pytest

import scenes
import yvo
import yoxel_voxel_b200 as yv

OUTSIDE = scenes.CAMERAS[2]
LIGHTS = [dict(pos=(1.3, 1.2, 0.9), diffuse=(0.7, 0.7, 0.7), specular=(0.3, 0.3, 0.3), attenuation=(1, 0, 0.5)),
          dict(pos=(0.45, 0.4, 0.95), diffuse=(1, 0.8, 0.6), specular=(0.3, 0.3, 0.3), attenuation=(1, 2, 4))]


def _cam(spec, W, H, **kw):
    return yvo.camera(spec[1], spec[2], spec[3], spec[4], W, H, **kw)


# ---- oracle: the written spec against closed forms ---------------------------------------------------------------

def test_blur_taps_follow_init_blur():
    k = yvo.blur_taps()
    assert k.shape == (7, 7)
    assert abs(float(k.sum()) - 1.0) < 1e-6
    assert np.array_equal(k, k.T) and np.array_equal(k, k[::-1]) and np.array_equal(k, k[:, ::-1])
    assert k.argmax() == 24                                            # centre tap
    # scale 2 over half-width 3: the corner is exp(-(4+4)) of the centre (demo/SVORenderer.cpp:59-71)
    assert np.isclose(k[0, 0] / k[3, 3], np.exp(-8.0), rtol=1e-5)
    assert np.isclose(k[3, 0] / k[3, 3], np.exp(-4.0), rtol=1e-5)


def test_blur_keeps_constant_depth_and_holes():
    cam = yvo.camera((0, 0, 0), (1, 0, 0), width=48, height=40)
    z = np.full((40, 48), 0.75, np.float32)
    z[10:14, 20:30] = 0.0                                              # misses stay misses and are never averaged in
    out = yvo.blur_z(cam, z)
    assert (out[z == 0] == 0).all()
    assert np.allclose(out[z != 0], 0.75, rtol=1e-6)


def test_blur_does_not_mix_across_depth_edges():
    """Taps farther than zlimit from the centre depth are rejected: a step edge survives five passes."""
    cam = yvo.camera((0, 0, 0), (1, 0, 0), width=64, height=32, ssna_voxel_size=1.0 / 2048)
    z = np.full((32, 64), 0.5, np.float32)
    z[:, 32:] = 1.5                                                    # jump of 1.0 >> zlimit (~0.04 for pass 0)
    out = yvo.blur_z(cam, z)
    assert np.allclose(out[:, :32], 0.5, rtol=1e-6) and np.allclose(out[:, 32:], 1.5, rtol=1e-6)


def test_blur_smooths_voxel_steps():
    """A staircase whose steps are below zlimit is flattened towards the underlying ramp."""
    cam = yvo.camera((0, 0, 0), (1, 0, 0), width=96, height=24, fov=70.0)
    x = np.arange(96, dtype=np.float32)
    ramp = 0.5 + x * 1e-4
    stairs = 0.5 + np.floor(x / 4) * 4e-4
    z = np.tile(stairs, (24, 1)).astype(np.float32)
    out = yvo.blur_z(cam, z)
    mid = slice(20, 76)
    err_in = np.abs(z[12, mid] - ramp[mid] - (z[12, mid] - ramp[mid]).mean()).max()
    err_out = np.abs(out[12, mid] - ramp[mid] - (out[12, mid] - ramp[mid]).mean()).max()
    assert err_out < 0.25 * err_in


def test_normal_of_fronto_parallel_plane_faces_the_camera():
    cam = yvo.camera((0.1, 0.2, 0.3), (0.3, -0.5, 0.2), up=(0, 0, 1), width=40, height=30)
    z = np.full((30, 40), 0.8, np.float32)
    fwd = np.array([0.3, -0.5, 0.2]); fwd /= np.linalg.norm(fwd)
    for (x, y) in [(0, 0), (20, 15), (39, 29), (39, 0)]:
        n = yvo.ssna_normal(cam, z, x, y)
        assert np.allclose(n, -fwd, atol=1e-6)
    z[15, 20] = 0.0
    assert yvo.ssna_normal(cam, z, 20, 15) is None                      # invalid pixel: caller keeps the voxel normal


def test_normal_of_tilted_plane_near_the_axis():
    """z(x) of a plane tilted about the view's vertical axis: near the optical axis the reconstruction is
    the plane's normal (the prototype drops the off-axis terms, ztools.py:36-38)."""
    W, H, fov = 200, 100, 40.0
    cam = yvo.camera((0, 0, 0), (1, 0, 0), up=(0, 0, 1), fov=fov, width=W, height=H)
    d2 = 2 * np.tan(np.radians(fov / 2)) / W
    # view space: x_v right, y_v down, z_v forward; plane  z_v = z0 + s * x_v  =>  z = z0 / (1 - s * u), u = (x - W/2) * d2
    z0, s = 2.0, 0.5
    u = (np.arange(W) - W / 2) * d2
    z = np.tile(z0 / (1 - s * u), (H, 1)).astype(np.float32)
    n = yvo.ssna_normal(cam, z, W // 2, H // 2)
    # world frame of this camera: fwd = +x, right = fwd x up = (0,-1,0), down = (0,0,-1)
    nv = np.array([s, 0.0, -1.0]); nv /= np.linalg.norm(nv)            # camera-facing normal in view space
    expect = nv[0] * np.array([0, -1, 0]) + nv[1] * np.array([0, 0, -1]) + nv[2] * np.array([1, 0, 0])
    assert np.allclose(n, expect, atol=2e-3)


def test_one_sided_difference_at_silhouettes():
    cam = yvo.camera((0, 0, 0), (1, 0, 0), width=16, height=16)
    z = np.zeros((16, 16), np.float32)
    z[8, 8] = 1.0                                                      # an isolated hit pixel: no valid neighbour
    fwd = np.array([1.0, 0, 0])
    assert np.allclose(yvo.ssna_normal(cam, z, 8, 8), -fwd, atol=1e-6)
    z[8, 9] = 1.01                                                     # only the forward difference exists
    n = yvo.ssna_normal(cam, z, 8, 8)
    assert n is not None and abs(np.linalg.norm(n) - 1) < 1e-5 and not np.allclose(n, -fwd, atol=1e-4)
    z[8, 7] = 1.0                                                      # backward difference 0 has the smaller magnitude
    assert np.allclose(yvo.ssna_normal(cam, z, 8, 8), -fwd, atol=1e-6)


def _sphere_normals(cam_spec, W, H, depth, o):
    n = 1 << depth
    ctr = np.array([n // 2, n // 2 + 1, n // 2 - 2]) / n                # scenes.single_sphere
    d0, du, dv = yvo.init_ray_dir(_cam(cam_spec, W, H))
    ys, xs = np.mgrid[:H, :W]
    dirs = d0 + du * xs[..., None] + dv * ys[..., None]
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    P = np.array(cam_spec[1]) + dirs * o["t"][..., None]
    N = P - ctr
    return N / np.linalg.norm(N, axis=-1, keepdims=True)


def test_oracle_ssna_frame_on_a_sphere():
    depth, W, H = 8, 256, 192
    svo = scenes.single_sphere(depth)
    plain = yvo.render(svo.nodes(), svo.GetRoot(), _cam(OUTSIDE, W, H, show_normals=True), threads=8)
    ssna = yvo.render(svo.nodes(), svo.GetRoot(),
                      _cam(OUTSIDE, W, H, show_normals=True, ssna=True, ssna_voxel_size=2.0 ** -depth), threads=8)
    hit = plain["rgba"][..., 3] > 0
    assert hit.sum() > 5000
    assert np.array_equal(ssna["rgba"][..., 3], plain["rgba"][..., 3])          # same coverage; misses stay (0,0,0,0)
    assert (ssna["rgba"][~hit] == 0).all()
    assert np.array_equal(ssna["node"], plain["node"]) and np.array_equal(ssna["t"], plain["t"])   # tracing untouched
    assert (ssna["rgba"] != plain["rgba"]).any()
    N = _sphere_normals(OUTSIDE, W, H, depth, plain)
    dec = ssna["rgba"][..., :3].astype(np.float32) / 255 * 2 - 1
    dec /= np.maximum(1e-6, np.linalg.norm(dec, axis=-1, keepdims=True))
    cosang = (dec * N).sum(-1)[hit]
    assert cosang.mean() > 0.97 and np.percentile(cosang, 5) > 0.85


def test_oracle_ssna_needs_the_whole_frame_and_primary_rays():
    svo = scenes.single_sphere(6)
    cam = _cam(OUTSIDE, 64, 48, ssna=True)
    with pytest.raises(AssertionError):
        yvo.render(svo.nodes(), svo.GetRoot(), cam, rows=(0, 24))
    sec = yvo.secondary(shadow=1, ao_samples=2, voxel_size=2.0 ** -6)
    a = yvo.render(svo.nodes(), svo.GetRoot(), cam, sec=sec)
    b = yvo.render(svo.nodes(), svo.GetRoot(), _cam(OUTSIDE, 64, 48), sec=sec)
    assert np.array_equal(a["rgba"], b["rgba"])                         # secondary-ray frames ignore SSNA


# ---- GPU: the CUDA passes against the oracle ---------------------------------------------------------------------

def _gpu_frame(r, spec, W, H, detail=0.0):
    r.SetResolution(W, H)
    r.SetDetailCoef(detail)
    r.SetViewPos(spec[1]); r.SetViewDir(spec[2]); r.SetViewUp(spec[3]); r.SetFOV(spec[4])
    r.SetSecondary(0, 0)
    return r.RenderFrame().copy()


@pytest.fixture(scope="module")
def renderer():
    r = yv.SVORenderer(0)
    yield r
    r.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["lambert", "phong", "normals"])
@pytest.mark.parametrize("size", [(480, 320), (333, 217)], ids=["480x320", "333x217"])
def test_gpu_ssna_matches_oracle(renderer, mode, size):
    W, H = size
    svo = scenes.fractal(10)
    spec = scenes.CAMERAS[1]
    renderer.SetScene(svo)
    kw = dict(ssna=True, ssna_voxel_size=2.0 ** -10)
    if mode == "phong":
        kw["lights"] = LIGHTS
        for i, lt in enumerate(LIGHTS):
            renderer.SetLigth(i, yv.LightParams(True, lt["pos"], lt["diffuse"], lt["specular"], lt["attenuation"]))
    renderer.SetShowNormals(mode == "normals")
    kw["show_normals"] = mode == "normals"
    renderer.SetSSNA(True, 2.0 ** -10)
    assert renderer.GetSSNA()
    try:
        img = _gpu_frame(renderer, spec, W, H)
        o = yvo.render(svo.nodes(), svo.GetRoot(), _cam(spec, W, H, **kw), threads=8)
        assert np.abs(img.astype(int) - o["rgba"].astype(int)).max() <= 1
        assert (img == o["rgba"]).all(), "%d pixels differ" % (img != o["rgba"]).any(-1).sum()
        for fused in (1, 2):                                            # ssna_post: TMA-prefetched tiles / plain staging
            renderer.SetOption("ssna_fused", fused)
            again = _gpu_frame(renderer, spec, W, H)
            renderer.SetOption("ssna_fused", 0)
            assert (again == img).all(), "ssna_fused %d: %d pixels differ" % (fused, (again != img).any(-1).sum())
        renderer.SetSSNA(False)
        plain = _gpu_frame(renderer, spec, W, H)
        assert (plain != img).any() and np.array_equal(plain[..., 3], img[..., 3])
    finally:
        renderer.SetSSNA(False)
        renderer.SetShowNormals(False)
        for i in range(2):
            renderer.SetLigth(i, yv.LightParams(False))


@pytest.mark.gpu
@pytest.mark.parametrize("spec", [scenes.CAMERAS[2], scenes.CAMERAS[4], scenes.CAMERAS[6]], ids=lambda s: s[0])
def test_gpu_ssna_cameras_default_voxel_size_and_lod(renderer, spec):
    """Reference voxSize (1/2048), other cameras (inside the scene: t < 0 quirk hits included), LOD hits."""
    svo = scenes.fractal(9)
    renderer.SetScene(svo)
    renderer.SetSSNA(True, 0.0)
    try:
        for detail in (0.0, 2.0):
            img = _gpu_frame(renderer, spec, 400, 300, detail)
            o = yvo.render(svo.nodes(), svo.GetRoot(), _cam(spec, 400, 300, ssna=True, detail_coef=detail), threads=8)
            assert (img == o["rgba"]).all(), "%s detail %g: %d pixels differ" % (spec[0], detail, (img != o["rgba"]).any(-1).sum())
    finally:
        renderer.SetSSNA(False)
        renderer.SetDetailCoef(0.0)


@pytest.mark.gpu
def test_gpu_ssna_device_render_and_partition_errors():
    import torch
    renderer = yv.SVORenderer(0)
    svo = scenes.single_sphere(7)
    renderer.SetScene(svo)
    W, H = 320, 256
    renderer.SetSSNA(True, 2.0 ** -7)
    try:
        host = _gpu_frame(renderer, OUTSIDE, W, H)
        dst = torch.zeros(H, W, 4, dtype=torch.uint8, device="cuda:0")
        renderer.Render(dst.data_ptr())                                 # SVORenderer::Render(void* d_dstBuf)
        assert np.array_equal(dst.cpu().numpy(), host)
        assert renderer.LastFrameLaunches() == 7                        # trace (writes z itself) + 5 x BlurZ + ShadeSimple
        try:
            for fused in (1, 2):                                        # the same passes as one persistent cooperative launch:
                renderer.SetOption("ssna_fused", fused)                 # 1 = tiles prefetched by TMA, 2 = plain staging loads
                assert np.array_equal(_gpu_frame(renderer, OUTSIDE, W, H), host)
                assert renderer.LastFrameLaunches() == 2                # trace + ssna_post
                assert np.array_equal(_gpu_frame(renderer, OUTSIDE, W, H), host)     # ... again: the tile counters were reset
            renderer.SetOption("ssna_fused", 1)
            for mode in (1, 2):                                         # ... on the other schedules (ssna_z_pass runs: + 1)
                renderer.SetOption("persistent", mode)
                assert np.array_equal(_gpu_frame(renderer, OUTSIDE, W, H), host)
                assert renderer.LastFrameLaunches() == 3
            renderer.SetOption("persistent", 0)
            for (w2, h2) in ((333, 217), (644, 36), (1000, 500)):       # a row pitch that is no 16-byte multiple (no TMA), odd shapes
                renderer.SetOption("ssna_fused", 0)
                want = _gpu_frame(renderer, OUTSIDE, w2, h2)
                renderer.SetOption("ssna_fused", 1)
                assert np.array_equal(_gpu_frame(renderer, OUTSIDE, w2, h2), want), (w2, h2)
        finally:
            renderer.SetOption("ssna_fused", 0)
            renderer.SetOption("persistent", 0)
        renderer.SetOption("persistent", 1)                             # a schedule without the z epilogue: ssna_z_pass runs
        try:
            assert np.array_equal(_gpu_frame(renderer, OUTSIDE, W, H), host)
            assert renderer.LastFrameLaunches() == 8                    # trace + z + 5 x BlurZ + ShadeSimple
        finally:
            renderer.SetOption("persistent", 0)
        renderer.SetRows(0, H // 2)
        with pytest.raises(yv.YVError):
            renderer.RenderFrame()
        renderer.SetRows(0, H)
        renderer.SetInterleave(32, 2, 0)
        with pytest.raises(yv.YVError):
            renderer.RenderFrame()
        renderer.SetInterleave(32, 1, 0)                                # stride 1: contiguous again
        renderer.SetSSNA(False)
        assert np.array_equal(_gpu_frame(renderer, OUTSIDE, W, H)[..., 3], host[..., 3])
    finally:
        renderer.close()


def _view_space(cam_dir, up, n_world):
    fwd = np.array(cam_dir, np.float64); fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.array(up, np.float64)); right /= np.linalg.norm(right)
    down = -np.cross(right, fwd)
    return np.array([n_world @ right, n_world @ down, n_world @ fwd])


def _check_against_prototype(z, proto):
    """The oracle's SSNA normal, taken back to view space, is minus the prototype's (which points away from the
    camera: nz = +d^2 z^2, ztools.py:38) at every pixel the prototype defines (it drops the one-pixel border)."""
    H, W = z.shape
    cam_dir, up = (0.3, -0.5, 0.2), (0, 0, 1)
    cam = yvo.camera((0.1, 0.2, 0.3), cam_dir, up=up, fov=70.0, width=W, height=H)         # ztools.py:16: fov = 70
    worst = 0.0
    for y in range(1, H - 1, 3):
        for x in range(1, W - 1, 2):
            n = yvo.ssna_normal(cam, z, x, y)
            nv = _view_space(cam_dir, up, n.astype(np.float64))
            worst = max(worst, float(np.abs(nv + proto[y - 1, x - 1]).max()))
    assert worst < 2e-4, worst


def test_normals_equal_the_reference_prototype_golden():
    """calcNormals of demo/dumps/ztools.py:23-44, run by tests/golden/make_ztools_golden.py, against the oracle."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ztools_normals.npz"))
    for i in range(2):
        _check_against_prototype(g["z%d" % i], g["n%d" % i])


def test_normals_equal_the_reference_prototype_live():
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_ztools_golden.py")
    spec = importlib.util.spec_from_file_location("make_ztools_golden", path)
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    if not os.path.exists(mod.ZTOOLS):
        pytest.skip("the prototype lives in /root/reference")
    calc = mod.load_calc_normals()
    rng = np.random.RandomState(8)
    y, x = np.mgrid[:40, :56].astype(np.float64)
    z = (0.9 + 0.2 * np.sin(x / 5.0) * np.sin(y / 6.0) + rng.rand(40, 56) * 2e-3).astype(np.float32)
    _check_against_prototype(z, calc(z.astype(np.float64)))


def test_blur_taps_do_not_depend_on_the_exp_overload():
    """`float v = exp(-(tx + ty))` (demo/SVORenderer.cpp:70) picks the float overload in C++; the spec evaluates it in double
    and rounds once. For the 49 taps of the 7x7 kernel the two give the same floats, so the table is pinned either way."""
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.expf.restype = ctypes.c_float; libm.expf.argtypes = [ctypes.c_float]
    K, f = 7, np.float32
    h, v = f(K // 2), np.zeros((K, K), np.float32)
    for y in range(K):
        for x in range(K):
            tx = f(f(f(2.0) * f(f(x) - h)) / h); ty = f(f(f(2.0) * f(f(y) - h)) / h)
            v[y, x] = libm.expf(float(-f(f(tx * tx) + f(ty * ty))))
    s = f(0)
    for y in range(K):
        for x in range(K):
            s = f(s + v[y, x])
    assert (yvo.blur_taps() == (v / s).astype(np.float32)).all()
