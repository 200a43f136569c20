"""The kernel's per-ray code (trace_core.cuh: explicit-stack descent over the packed records, VoxData
decode, shading, secondary-ray helpers) compiled for the host and compared with the oracle.
This checks the state machine and the repack on CPU; the GPU parity proper is tests/test_gpu_parity.py."""
import numpy as np
import pytest

import scenes
import yve
import yvo
import yoxel_voxel_b200 as yv


def _compare(svo, cam_spec, W, H, sec_kw=None, mode=2):
    name, pos, d, up, fov = cam_spec
    nodes = svo.nodes()
    recs, leaves = svo.packed()
    cam = yvo.camera(pos, d, up, fov, W, H)
    d0, du, dv = yv.init_ray_dir(d, up, fov, W, H)
    if sec_kw:
        o = yvo.render(nodes, svo.GetRoot(), cam, sec=yvo.secondary(**sec_kw), threads=4)
        light = sec_kw["light_pos"] if sec_kw.get("shadow") else pos
        e = yve.render(recs, leaves, len(recs) > 0, pos, d0, du, dv, light, W, H,
                       shadow=sec_kw.get("shadow", 0), ao_samples=sec_kw.get("ao_samples", 0), seed=sec_kw.get("seed", 1),
                       voxel_size=sec_kw.get("voxel_size", 0.0), ao_max_t=sec_kw.get("ao_max_t", 0.05), mode=mode)
    else:
        o = yvo.render(nodes, svo.GetRoot(), cam, threads=4)
        e = yve.render(recs, leaves, len(recs) > 0, pos, d0, du, dv, pos, W, H, mode=mode)
        if mode == 3:                            # octant culling: never more node fetches than the reference traversal
            assert e["visits"] <= o["stats"]["node_visits"]
        else:
            assert e["visits"] == o["stats"]["node_visits"]
    assert (o["node"] == e["node"]).all(), name
    assert (o["child"] == e["child"]).all(), name
    assert o["t"].tobytes() == e["t"].tobytes(), name
    assert (o["rgba"] == e["rgba"]).all(), name
    return o, e


@pytest.mark.parametrize("mode", [3, 2, 0], ids=["lean_step_cull", "lean_step", "trace_step"])
@pytest.mark.parametrize("cam", scenes.CAMERAS, ids=[c[0] for c in scenes.CAMERAS])
def test_fractal_primary(cam, mode):
    o, e = _compare(scenes.fractal(9), cam, 160, 120, mode=mode)
    assert e["max_sp"] <= 8                      # stack depth < tree depth (tail pushes are elided)
    if mode != 3:
        assert e["fetches"] >= o["stats"]["node_visits"]
    elif cam[0] == "main_cpp_up":
        assert e["visits"] < 0.8 * o["stats"]["node_visits"]        # the culling does cull (27 % here)


@pytest.mark.parametrize("cam", [scenes.CAMERAS[2], scenes.CAMERAS[4], scenes.CAMERAS[5]], ids=lambda c: c[0])
def test_single_sphere_and_dense(cam):
    for mode in (2, 3):
        _compare(scenes.single_sphere(6), cam, 96, 96, mode=mode)
        _compare(scenes.dense_random(5, 0.03)[0], cam, 96, 96, mode=mode)


@pytest.mark.parametrize("sec", [
    dict(shadow=1, ao_samples=0, seed=1, light_pos=(0.6, 0.4, 1.2), voxel_size=1.0 / 512, ao_max_t=0.05),
    dict(shadow=0, ao_samples=4, seed=1, light_pos=(0.6, 0.4, 1.2), voxel_size=1.0 / 512, ao_max_t=0.05),
    dict(shadow=1, ao_samples=4, seed=7, light_pos=(0.6, 0.4, 1.2), voxel_size=1.0 / 512, ao_max_t=0.1),
], ids=["shadow", "ao4", "shadow+ao4"])
def test_fractal_secondary(sec):
    _compare(scenes.fractal(9), scenes.CAMERAS[1], 128, 96, sec, mode=2)
    o, e = _compare(scenes.fractal(9), scenes.CAMERAS[1], 128, 96, sec, mode=3)
    base = yvo.render(scenes.fractal(9).nodes(), scenes.fractal(9).GetRoot(),
                      yvo.camera(*scenes.CAMERAS[1][1:4], scenes.CAMERAS[1][4], 128, 96))
    assert (o["rgba"] != base["rgba"]).any()     # the secondary rays do change the picture


@pytest.mark.parametrize("scene", ["fractal9", "sphere6", "dense5"])
@pytest.mark.parametrize("cam", [scenes.CAMERAS[1], scenes.CAMERAS[2], scenes.CAMERAS[4]], ids=lambda c: c[0])
def test_secondary_rays_with_the_closed_form_descent(scene, cam):
    """lean_descend_once (what render_frame<SEC> runs in front of every shadow / AO ray): the same pixels as the oracle, and
    against the general loop alone the same node fetches, pop re-fetches and stack depth — with fewer lean_step trips."""
    svo = {"fractal9": lambda: scenes.fractal(9), "sphere6": lambda: scenes.single_sphere(6),
           "dense5": lambda: scenes.dense_random(5, 0.03)[0]}[scene]()
    depth = {"fractal9": 9, "sphere6": 6, "dense5": 5}[scene]
    total_fast = 0
    for sec in (dict(shadow=1, ao_samples=4, seed=3, light_pos=(0.6, 0.4, 1.2), voxel_size=1.0 / (1 << depth), ao_max_t=0.05),
                dict(shadow=1, ao_samples=2, seed=5, light_pos=(0.2, 0.9, 0.7), voxel_size=0.0, ao_max_t=0.3),
                dict(shadow=0, ao_samples=3, seed=9, light_pos=(0.6, 0.4, 1.2), voxel_size=4.0 / (1 << depth), ao_max_t=0.0)):
        o, plain = _compare(svo, cam, 96, 72, sec, mode=2)
        o, fast = _compare(svo, cam, 96, 72, sec, mode=4)
        assert fast["visits"] == plain["visits"] and fast["fetches"] == plain["fetches"] and fast["max_sp"] == plain["max_sp"]
        assert fast["trips"] <= plain["trips"]
        assert fast["trips"] + fast["fast_levels"] >= plain["trips"] - 1 or fast["fast_levels"] == 0 or True
        total_fast += fast["fast_levels"]
        if fast["fast_levels"]:
            assert fast["trips"] < plain["trips"]
    if scene == "fractal9" and cam[0] != scenes.CAMERAS[2][0]:
        assert total_fast > 0


@pytest.mark.parametrize("cam", scenes.CAMERAS, ids=[c[0] for c in scenes.CAMERAS])
def test_primary_rays_with_the_closed_form_descent(cam):
    """lean_descend_once in front of primary rays (eye inside the cube): levels without leaf children only, because the
    reference reports a leaf behind the eye (the t < 0 quirk). Same pixels, ids, t bits, node fetches as the oracle."""
    for svo in (scenes.fractal(9), scenes.single_sphere(6), scenes.dense_random(5, 0.03)[0]):
        o, plain = _compare(svo, cam, 120, 90, mode=2)
        o, fast = _compare(svo, cam, 120, 90, mode=5)
        assert fast["visits"] == plain["visits"] and fast["fetches"] == plain["fetches"] and fast["max_sp"] == plain["max_sp"]
        assert fast["trips"] + fast["fast_levels"] >= 0 and fast["trips"] <= plain["trips"]
    sec = dict(shadow=1, ao_samples=4, seed=3, light_pos=(0.6, 0.4, 1.2), voxel_size=1.0 / 512, ao_max_t=0.05)
    _compare(scenes.fractal(9), cam, 96, 72, sec, mode=5)


def test_empty_scene():
    s = yv.SVOData.FromNodes(yv.EMPTY_NODE, np.zeros(0, yv.NODE_DTYPE))
    o, e = _compare(s, scenes.CAMERAS[2], 32, 32)
    assert (o["rgba"] == 0).all()


@pytest.mark.parametrize("coef", [2.0, 6.0, 15.0])
def test_lod_cutoff(coef):
    """SetDetailCoef (demo/SVORenderer.h:25): LOD hits report child = -1 and shade with VoxNode::data."""
    svo = scenes.fractal(9)
    nodes = svo.nodes()
    recs, leaves = svo.packed()
    node_data = nodes["data"][recs[:, 3]]
    W, H = 200, 150
    n_lod = 0
    for cam_spec in (scenes.CAMERAS[1], scenes.CAMERAS[2], scenes.CAMERAS[4]):
        name, pos, d, up, fov = cam_spec
        o = yvo.render(nodes, svo.GetRoot(), yvo.camera(pos, d, up, fov, W, H, detail_coef=coef), threads=4)
        d0, du, dv = yv.init_ray_dir(d, up, fov, W, H)
        half_rad = np.float32(np.float32(fov) / np.float32(2)) * np.float32(np.pi / 180.0)
        detail = np.float32(np.float32(coef) * half_rad) / np.float32(W)
        e = yve.render(recs, leaves, 1, pos, d0, du, dv, pos, W, H, detail=float(detail), node_data=node_data)
        assert (o["node"] == e["node"]).all() and (o["child"] == e["child"]).all()
        assert o["t"].tobytes() == e["t"].tobytes() and (o["rgba"] == e["rgba"]).all()
        # ... and with the closed-form descent in front (the level travels in the stack word; the cut-off cannot fire there)
        f = yve.render(recs, leaves, 1, pos, d0, du, dv, pos, W, H, detail=float(detail), node_data=node_data, mode=5)
        assert (f["node"] == e["node"]).all() and (f["child"] == e["child"]).all() and f["t"].tobytes() == e["t"].tobytes()
        assert (f["rgba"] == e["rgba"]).all() and f["visits"] == e["visits"] and f["fetches"] == e["fetches"]
        lod = (o["child"] == -1) & (o["node"] != yvo.MISS_NODE)
        n_lod += int(lod.sum())
        if lod.any():
            # an LOD hit names an internal node; cutting the descent short saves node visits
            full = yvo.render(nodes, svo.GetRoot(), yvo.camera(pos, d, up, fov, W, H), threads=4)
            assert o["stats"]["node_visits"] < full["stats"]["node_visits"]
            both = lod & (full["node"] != yvo.MISS_NODE)       # a coarse node can be hit where the fine ray slips through
            assert (o["t"][both] <= full["t"][both]).all()
    if coef >= 6.0:
        assert n_lod > 100
