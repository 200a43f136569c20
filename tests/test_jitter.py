"""Hiding voxelisation artefacts (reaction/report/main.tex:107-114): displaced ray origins and frame averaging.

The report describes the two techniques in prose only, so include/yv_format.h states the arithmetic. The CPU tests
check the oracle's statement of it (displacement size, determinism, independence from the traversal); the GPU tests
require the CUDA frames, and the integer mean of several of them, to equal the oracle's."""
import numpy as np
import pytest

import scenes
import yvo
import yoxel_voxel_b200 as yv

CAM = scenes.CAMERAS[1]
W, H = 320, 200


def _cam(spec=CAM, w=W, h=H, **kw):
    return yvo.camera(spec[1], spec[2], spec[3], spec[4], w, h, **kw)


def _mean(frames):
    n = len(frames)
    return ((np.sum([f.astype(np.uint32) for f in frames], axis=0) + n // 2) // n).astype(np.uint8)


def _hash_u32(x):
    x &= 0xFFFFFFFF
    x ^= x >> 16; x = (x * 0x7feb352d) & 0xFFFFFFFF
    x ^= x >> 15; x = (x * 0x846ca68b) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def _unit_vector(key):
    """The lattice-rejection unit vector of include/yv_format.h, restated in numpy float32."""
    f = np.float32
    for k in range(8):
        h = _hash_u32((key + 0x9e3779b9 * k) & 0xFFFFFFFF)
        v = np.array([f(h & 1023) - f(511.5), f((h >> 10) & 1023) - f(511.5), f((h >> 20) & 1023) - f(511.5)], f)
        l2 = f(f(v[0] * v[0]) + f(v[1] * v[1])) + f(v[2] * v[2])
        if 1.0 <= l2 <= 261632.25:
            return (v / np.sqrt(l2, dtype=f)).astype(f)
    return np.array([0, 0, 1], f)


def test_oracle_jitter_moves_origins_by_the_amplitude():
    """Each pixel's ray is the ordinary pixel direction from eye + amplitude * U(pixel, seed): re-derive the origin
    in numpy from the written spec and trace single rays from it."""
    svo = scenes.single_sphere(7)
    spec = scenes.CAMERAS[2]
    amp, seed = np.float32(2.0 ** -7), 5
    base = yvo.render(svo.nodes(), svo.GetRoot(), _cam(spec), threads=4)
    jit = yvo.render(svo.nodes(), svo.GetRoot(), _cam(spec, jitter_amp=amp, jitter_seed=seed), threads=4)
    both = (base["rgba"][..., 3] > 0) & (jit["rgba"][..., 3] > 0)
    assert both.sum() > 3000
    dt = np.abs(jit["t"][both] - base["t"][both])
    assert dt.max() < 40 * amp and np.median(dt) < 2 * amp and (dt > 0).mean() > 0.5
    d0, du, dv = yvo.init_ray_dir(_cam(spec))
    eye = np.array(spec[1], np.float32)
    ys, xs = np.nonzero(jit["rgba"][..., 3] > 0)
    f = np.float32
    for i in range(0, len(ys), max(1, len(ys) // 40)):
        x, y = int(xs[i]), int(ys[i])
        U = _unit_vector(_hash_u32(y * W + x) ^ _hash_u32(seed ^ 0x6a09e667))
        assert abs(float(np.linalg.norm(U)) - 1) < 1e-6
        org = (eye + amp * U).astype(f)
        d = ((d0 + du * f(x)).astype(f) + dv * f(y)).astype(f)
        n2 = f(f(f(d[0] * d[0]) + f(d[1] * d[1])) + f(d[2] * d[2]))
        d = (d / np.sqrt(n2, dtype=f)).astype(f)
        hit, node, child, t = yvo.trace_ray(svo.nodes(), svo.GetRoot(), org, d)
        assert hit and node == jit["node"][y, x] and child == jit["child"][y, x]
        assert np.float32(t).tobytes() == jit["t"][y, x].tobytes()


def test_oracle_jitter_is_deterministic_and_seeded():
    svo = scenes.fractal(8)
    a = yvo.render(svo.nodes(), svo.GetRoot(), _cam(jitter_amp=1e-3, jitter_seed=3), threads=4)
    b = yvo.render(svo.nodes(), svo.GetRoot(), _cam(jitter_amp=1e-3, jitter_seed=3), threads=2)
    c = yvo.render(svo.nodes(), svo.GetRoot(), _cam(jitter_amp=1e-3, jitter_seed=4), threads=4)
    z = yvo.render(svo.nodes(), svo.GetRoot(), _cam(jitter_amp=0.0, jitter_seed=3), threads=4)
    plain = yvo.render(svo.nodes(), svo.GetRoot(), _cam(), threads=4)
    assert np.array_equal(a["rgba"], b["rgba"]) and a["t"].tobytes() == b["t"].tobytes()
    assert (a["t"] != c["t"]).any()
    assert np.array_equal(z["rgba"], plain["rgba"]) and z["t"].tobytes() == plain["t"].tobytes()


def test_averaging_softens_voxel_edges():
    """The mean of jittered frames has partial coverage along silhouettes (the report's "blurry spots")."""
    svo = scenes.single_sphere(5)
    spec = scenes.CAMERAS[2]
    frames = [yvo.render(svo.nodes(), svo.GetRoot(), _cam(spec, jitter_amp=2.0 ** -5, jitter_seed=1 + k), threads=4)["rgba"]
              for k in range(8)]
    m = _mean(frames)
    alpha = m[..., 3]
    assert ((alpha > 0) & (alpha < 255)).sum() > 50
    assert (alpha == 255).sum() > 1000


# ---- GPU -----------------------------------------------------------------------------------------------------------

def _setup(r, spec, w, h):
    r.SetResolution(w, h)
    r.SetViewPos(spec[1]); r.SetViewDir(spec[2]); r.SetViewUp(spec[3]); r.SetFOV(spec[4])
    r.SetSecondary(0, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("detail", [0.0, 2.0], ids=["full", "lod"])
def test_gpu_jitter_matches_oracle(detail):
    svo = scenes.fractal(10)
    r = yv.SVORenderer(0)
    try:
        r.EnableHits(True)
        r.SetScene(svo)
        _setup(r, CAM, 400, 304)
        r.SetDetailCoef(detail)
        amp = 2.0 ** -10
        r.SetJitter(amp, 7)
        img = r.RenderFrame().copy()
        node, child, t = r.GetHits()
        o = yvo.render(svo.nodes(), svo.GetRoot(), _cam(CAM, 400, 304, jitter_amp=amp, jitter_seed=7, detail_coef=detail), threads=8)
        assert (node == o["node"]).all() and (child == o["child"]).all()
        assert t.tobytes() == o["t"].tobytes()
        assert (img == o["rgba"]).all()
        plain = yvo.render(svo.nodes(), svo.GetRoot(), _cam(CAM, 400, 304, detail_coef=detail), threads=8)
        assert (plain["t"] != o["t"]).any()
        r.SetJitter(0.0)
        assert (r.RenderFrame() == plain["rgba"]).all()
    finally:
        r.close()


@pytest.mark.gpu
def test_gpu_jitter_with_phong_and_ssna_passes():
    """The ShadeSimple / SSNA passes rebuild the shaded point from the displaced origin as well."""
    svo = scenes.fractal(9)
    lights = [dict(pos=(0.45, 0.4, 0.55), diffuse=(1, 0.8, 0.6), specular=(0.3, 0.3, 0.3), attenuation=(1, 2, 4))]
    r = yv.SVORenderer(0)
    try:
        r.SetScene(svo)
        _setup(r, CAM, 352, 240)
        r.SetJitter(1e-3, 2)
        r.SetLigth(0, yv.LightParams(True, lights[0]["pos"], lights[0]["diffuse"], lights[0]["specular"], lights[0]["attenuation"]))
        img = r.RenderFrame().copy()
        o = yvo.render(svo.nodes(), svo.GetRoot(), _cam(CAM, 352, 240, jitter_amp=1e-3, jitter_seed=2, lights=lights), threads=8)
        assert (img == o["rgba"]).all()
        r.SetSSNA(True, 2.0 ** -9)
        img = r.RenderFrame().copy()
        o = yvo.render(svo.nodes(), svo.GetRoot(), _cam(CAM, 352, 240, jitter_amp=1e-3, jitter_seed=2, lights=lights,
                                                         ssna=True, ssna_voxel_size=2.0 ** -9), threads=8)
        assert (img == o["rgba"]).all()
    finally:
        r.close()


@pytest.mark.gpu
def test_gpu_accumulated_frames_equal_the_integer_mean():
    svo = scenes.fractal(9)
    r = yv.SVORenderer(0)
    try:
        r.SetScene(svo)
        _setup(r, CAM, W, H)
        amp = 2.0 ** -9
        r.SetJitter(amp, 11)
        got = r.RenderAccumulated(6).copy()
        frames = [yvo.render(svo.nodes(), svo.GetRoot(), _cam(jitter_amp=amp, jitter_seed=11 + k), threads=8)["rgba"] for k in range(6)]
        assert (got == _mean(frames)).all()
        assert r.LastFrameLaunches() == 6 * 2 + 1                       # (trace + accumulate) x 6 + resolve
        one = r.RenderAccumulated(1).copy()                             # n = 1: the frame itself; the seed is restored
        assert (one == frames[0]).all()
        with pytest.raises(yv.YVError):
            r.RenderAccumulated(0)
    finally:
        r.close()


@pytest.mark.gpu
def test_gpu_jitter_rejects_unsupported_combinations():
    svo = scenes.single_sphere(6)
    r = yv.SVORenderer(0)
    try:
        r.SetScene(svo)
        _setup(r, scenes.CAMERAS[2], 128, 96)
        r.SetJitter(1e-3, 1)
        r.SetSecondary(shadow=1, ao_samples=2, voxel_size=2.0 ** -6)
        with pytest.raises(yv.YVError):
            r.RenderFrame()
        r.SetSecondary(0, 0)
        r.SetOption("schedule", 1)
        with pytest.raises(yv.YVError):
            r.RenderFrame()
        r.SetOption("schedule", 0)
        assert r.RenderFrame() is not None
    finally:
        r.close()
