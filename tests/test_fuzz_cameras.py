"""Seeded random cameras (eye inside / outside / on faces of the cube, axis-parallel views, tiny and huge fields of
view, odd aspect ratios): the kernel's per-ray code must agree with the oracle bit for bit on every one.
CPU: the host build of trace_core.cuh (tests/emu); GPU: the CUDA path."""
import numpy as np
import pytest

import scenes
import yve
import yvo
import yoxel_voxel_b200 as yv


def _cameras(n, seed):
    rng = np.random.RandomState(seed)
    cams = []
    for i in range(n):
        kind = i % 6
        if kind == 0:      # anywhere around the cube
            pos = rng.uniform(-1.5, 2.5, 3)
        elif kind == 1:    # inside the cube
            pos = rng.uniform(0.05, 0.95, 3)
        elif kind == 2:    # exactly on a face / grid plane
            pos = rng.uniform(0, 1, 3); pos[rng.randint(3)] = rng.choice([0.0, 0.5, 1.0, 0.25])
        elif kind == 3:    # far away
            pos = rng.uniform(-30, 30, 3)
        else:
            pos = rng.uniform(-0.5, 1.5, 3)
        if kind == 4:      # axis-parallel view (AdjustDir on whole rows / columns)
            d = np.zeros(3); d[rng.randint(3)] = rng.choice([-1.0, 1.0])
        else:
            d = rng.uniform(0, 1, 3) - pos + rng.normal(0, 0.2, 3)      # roughly towards the cube
            if np.linalg.norm(d) < 1e-3:
                d = np.array([1.0, 0.3, 0.2])
        up = np.array([0.0, 0.0, 1.0]) if abs(d[2]) < 0.9 * np.linalg.norm(d) else np.array([0.0, 1.0, 0.0])
        fov = float(rng.choice([5.0, 30.0, 70.0, 110.0, 150.0]))
        cams.append((tuple(float(np.float32(v)) for v in pos), tuple(float(np.float32(v)) for v in d), tuple(up), fov))
    return cams


@pytest.mark.parametrize("scene", ["fractal9", "dense5", "iso8"])
def test_emu_matches_oracle_on_random_cameras(scene):
    svo = {"fractal9": lambda: scenes.fractal(9), "dense5": lambda: scenes.dense_random(5, 0.03)[0],
           "iso8": lambda: yv.SVOData.IsoVolume(8, threads=4)}[scene]()
    nodes = svo.nodes()
    recs, leaves = svo.packed()
    node_data = nodes["data"][recs[:, 3]]
    hits = 0
    for i, (pos, d, up, fov) in enumerate(_cameras(36, seed=11)):
        W, H = [(64, 48), (37, 53), (96, 16)][i % 3]
        detail = [0.0, 0.0, 5.0][i % 3]
        o = yvo.render(nodes, svo.GetRoot(), yvo.camera(pos, d, up, fov, W, H, detail_coef=detail), threads=2)
        d0, du, dv = yv.init_ray_dir(d, up, fov, W, H)
        half_rad = np.float32(np.float32(fov) / np.float32(2)) * np.float32(np.pi / 180.0)
        det = float(np.float32(np.float32(detail) * half_rad) / np.float32(W))
        e = yve.render(recs, leaves, 1, pos, d0, du, dv, pos, W, H, detail=det, node_data=node_data)
        assert (o["node"] == e["node"]).all() and (o["child"] == e["child"]).all(), (scene, i)
        assert o["t"].tobytes() == e["t"].tobytes() and (o["rgba"] == e["rgba"]).all(), (scene, i)
        hits += int((o["node"] != yvo.MISS_NODE).sum())
    assert hits > 5000


@pytest.mark.gpu
def test_cuda_matches_oracle_on_random_cameras():
    r = yv.SVORenderer(0)
    r.EnableHits(True)
    sec = dict(shadow=1, ao_samples=3, seed=9, light_pos=(0.6, 0.4, 1.2), voxel_size=1.0 / 512, ao_max_t=0.08)
    for svo in (scenes.fractal(10), yv.SVOData.IsoVolume(9, threads=8)):
        r.SetScene(svo)
        nodes = svo.nodes()
        for i, (pos, d, up, fov) in enumerate(_cameras(48, seed=23)):
            W, H = [(160, 120), (97, 131), (256, 40)][i % 3]
            detail = [0.0, 0.0, 5.0][i % 3]
            use_sec = i % 4 == 3
            r.SetOption("schedule", i % 3 if not use_sec else i % 2)
            r.SetOption("layout", (i // 3) % 2)
            r.SetResolution(W, H)
            r.SetViewPos(pos); r.SetViewDir(d); r.SetViewUp(up); r.SetFOV(fov); r.SetDetailCoef(detail)
            r.SetSecondary(**sec) if use_sec else r.SetSecondary(0, 0)
            img = r.RenderFrame().copy()
            node, child, t = r.GetHits()
            o = yvo.render(nodes, svo.GetRoot(), yvo.camera(pos, d, up, fov, W, H, detail_coef=detail),
                           sec=yvo.secondary(**sec) if use_sec else None, threads=8)
            assert (node == o["node"]).all() and (child == o["child"]).all(), i
            assert np.allclose(t, o["t"], rtol=1e-4, atol=0) and t.tobytes() == o["t"].tobytes(), i
            assert np.abs(img.astype(int) - o["rgba"].astype(int)).max() <= 1 and (img == o["rgba"]).all(), i
    r.close()
