"""Committed fixtures (tests/golden/, produced by make_golden.py): the oracle must keep reproducing
them (CPU), and the CUDA path must match them (GPU)."""
import os

import numpy as np
import pytest

import scenes
import yvo
import yoxel_voxel_b200 as yv

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
W, H = 64, 48
SEC = dict(shadow=1, ao_samples=4, seed=1, light_pos=(0.6, 0.4, 1.2), voxel_size=1.0 / 64, ao_max_t=0.2)
SCENES = ("sphere6", "dense4", "two_level")


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden.npz"))


@pytest.mark.parametrize("scene", SCENES)
def test_oracle_reproduces_golden(golden, scene):
    root, nodes = yvo.load_vox(os.path.join(HERE, scene + ".vox"))
    for name, pos, d, up, fov in scenes.CAMERAS:
        r = yvo.render(nodes, root, yvo.camera(pos, d, up, fov, W, H))
        key = "%s/%s" % (scene, name)
        assert (r["node"] == golden[key + "/node"]).all()
        assert (r["child"] == golden[key + "/child"]).all()
        assert r["t"].tobytes() == golden[key + "/t"].tobytes()
        assert (r["rgba"] == golden[key + "/rgba"]).all()
        if scene == "sphere6":
            r2 = yvo.render(nodes, root, yvo.camera(pos, d, up, fov, W, H), sec=yvo.secondary(**SEC))
            assert (r2["rgba"] == golden[key + "/rgba_sec"]).all()


def test_golden_has_content(golden):
    hits = sum(int((golden[k] != yvo.MISS_NODE).sum()) for k in golden.files if k.endswith("/node"))
    assert hits > 20000


@pytest.mark.gpu
@pytest.mark.parametrize("persistent", [0, 1, 2])
@pytest.mark.parametrize("scene", SCENES)
def test_cuda_matches_golden(golden, scene, persistent):
    svo = yv.SVOData().Load(os.path.join(HERE, scene + ".vox"))
    r = yv.SVORenderer(0)
    r.SetOption("persistent", persistent)
    r.SetScene(svo)
    r.SetResolution(W, H)
    r.EnableHits(True)
    for name, pos, d, up, fov in scenes.CAMERAS:
        r.SetViewPos(pos); r.SetViewDir(d); r.SetViewUp(up); r.SetFOV(fov)
        r.SetSecondary(0, 0)
        img = r.RenderFrame().copy()
        node, child, t = r.GetHits()
        key = "%s/%s" % (scene, name)
        assert (node == golden[key + "/node"]).all(), key
        assert (child == golden[key + "/child"]).all(), key
        assert t.tobytes() == golden[key + "/t"].tobytes(), key
        assert (img == golden[key + "/rgba"]).all(), key
        if scene == "sphere6":
            r.SetSecondary(**SEC)
            img2 = r.RenderFrame().copy()
            assert (img2 == golden[key + "/rgba_sec"]).all(), key
    r.close()
