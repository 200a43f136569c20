"""Pins the CPU oracle independently of the CUDA path (the reference ships no golden vectors for
this path: SURVEY.md §8c). Hand-computed rays, reference quirks, structural properties and a
brute-force voxel-grid marcher (a different algorithm) on small scenes."""
import numpy as np
import pytest

import scenes
import yvo
import yoxel_voxel_b200 as yv


# ---- known-answer rays through a hand-built two-level tree ---------------------------------------

def test_kat_axis_ray_hits_leaf():
    nodes, root, leaf = scenes.two_level_tree()
    hit, node, child, t = yvo.trace_ray(nodes, root, (0.375, 0.375, -1.0), (0, 0, 1))
    assert hit and node == 0 and child == 7
    assert t == 1.25           # enters the leaf cube at z = 0.25; all midpoints are exact binary fractions


def test_kat_mirrored_ray_uses_dir_flags():
    nodes, root, leaf = scenes.two_level_tree()
    hit, node, child, t = yvo.trace_ray(nodes, root, (0.375, 0.375, 2.0), (0, 0, -1))
    assert hit and node == 0 and child == 7
    assert t == 1.5            # enters through the z = 0.5 face


def test_kat_diagonal_and_miss():
    nodes, root, leaf = scenes.two_level_tree()
    d = np.float32(1.0) / np.sqrt(np.float32(3.0))
    hit, node, child, t = yvo.trace_ray(nodes, root, (-0.5, -0.5, -0.5), (d, d, d))
    assert hit and node == 0 and child == 7
    assert abs(t - 0.75 * np.sqrt(3.0)) < 1e-5
    # passes beside the leaf cube
    assert not yvo.trace_ray(nodes, root, (0.75, 0.75, -1.0), (0, 0, 1))[0]
    # never reaches the unit cube
    assert not yvo.trace_ray(nodes, root, (2.0, 2.0, 2.0), (1, 0, 0))[0]


def test_kat_null_root_and_full_nodes_are_invisible():
    nodes, root, leaf = scenes.two_level_tree()
    assert not yvo.trace_ray(nodes, yv.EMPTY_NODE, (0.375, 0.375, -1.0), (0, 0, 1))[0]
    nodes2 = nodes.copy()
    nodes2[1]["child"][1] = yv.FULL_NODE      # FullNode is null to the tracer
    hit, node, child, t = yvo.trace_ray(nodes2, root, (0.75, 0.25, -1.0), (0, 0, 1))
    assert not hit


# ---- reference quirks (SURVEY §7 "hard parts", §8a) ----------------------------------------------

def test_quirk_argmin_tie_order():
    """On an exact three-way tie GoNext leaves through z first, then x, then y
    (cell/spu/trace_spu.cpp:75-78): the diagonal visits children 0 -> 4 -> 5 -> 7."""
    leaf = lambda r: yv.pack_voxdata(r, 0, 0, 0, 0, 1)
    def pool(leaf_children):
        nodes = np.zeros(1, yv.NODE_DTYPE)
        nodes[0]["child"][:] = yv.EMPTY_NODE
        flags = 0
        for c in range(8):
            if c in leaf_children:
                nodes[0]["child"][c] = leaf(10 * c)
                flags |= 1 << c
            else:
                flags |= 1 << (8 + c)
        nodes[0]["flags"] = flags
        return nodes
    d = np.float32(1.0) / np.sqrt(np.float32(3.0))
    o = (-1.0, -1.0, -1.0)
    assert yvo.trace_ray(pool({1, 2, 4}), 0, o, (d, d, d))[2] == 4     # 1 and 2 are never visited
    assert yvo.trace_ray(pool({1, 2, 5, 6}), 0, o, (d, d, d))[2] == 5
    assert yvo.trace_ray(pool({1, 2, 3, 6, 7}), 0, o, (d, d, d))[2] == 7
    assert not yvo.trace_ray(pool({1, 2, 3, 6}), 0, o, (d, d, d))[0]


def test_quirk_leaf_behind_eye_reports_negative_t():
    """The leaf test precedes the child's t2 > 0 test (ppu_renderer.cpp:27 vs :20), so a leaf behind an
    eye that sits inside a straddling parent is reported with t < 0 (demo/Demo.cpp:95 guards for it)."""
    nodes = np.zeros(1, yv.NODE_DTYPE)
    nodes[0]["child"][:] = yv.EMPTY_NODE
    nodes[0]["child"][0] = yv.pack_voxdata(9, 9, 9, 0, 0, 1)
    nodes[0]["flags"] = 1 | (0xFE << 8)
    d = np.float32(1.0) / np.sqrt(np.float32(3.0))
    hit, node, child, t = yvo.trace_ray(nodes, 0, (0.6, 0.6, 0.6), (d, d, d))
    assert hit and child == 0 and t < 0


def test_quirk_pixel_corner_sampling_and_row0_top():
    cam = yvo.camera((0.5, 0.5, -1.0), (0, 0, 1), up=(0, 1, 0), fov=70, width=8, height=6)
    d0, du, dv = yvo.init_ray_dir(cam)
    # the centre of the image is the corner shared by pixels (3,2),(4,2),(3,3),(4,3): pixel (W/2,H/2) looks forward
    centre = d0 + du * np.float32(4) + dv * np.float32(3)
    assert np.allclose(centre / np.linalg.norm(centre), (0, 0, 1), atol=1e-6)
    # no +0.5: pixel (0,0) is exactly dir0
    assert np.linalg.norm(d0) > 1.0
    # dv points down the image (row 0 on top): moving down decreases the 'up' component
    assert dv[1] < 0 and abs(du[1]) < 1e-7


def test_quirk_miss_colour_and_alpha():
    s = scenes.single_sphere(6)
    cam = yvo.camera((0.5, 0.5, -1.2), (0, 0, 1), up=(0, 1, 0), width=48, height=48)
    r = yvo.render(s.nodes(), s.GetRoot(), cam)
    miss = r["node"] == yvo.MISS_NODE
    assert miss.any() and (~miss).any()
    assert (r["rgba"][miss] == 0).all()                       # Color32(0,0,0,0)  (ppu_renderer.cpp:54)
    assert (r["rgba"][~miss][:, 3] == 255).all()
    assert (r["child"][miss] == -1).all() and (r["t"][miss] == 0).all()


def test_quirk_threaded_renderer_leaves_remainder_rows():
    s = scenes.single_sphere(6)
    cam = yvo.camera((0.5, 0.5, -1.2), (0, 0, 1), up=(0, 1, 0), width=40, height=30)   # 30 % 4 = 2
    full = yvo.render(s.nodes(), s.GetRoot(), cam)["rgba"]
    quirk = yvo.render_threaded_ref(s.nodes(), s.GetRoot(), cam, prefill=7)
    assert (quirk[:28] == full[:28]).all()
    assert (quirk[28:] == 7).all()                            # rows 4*(H/4)..H-1 never rendered (ppu_renderer.cpp:130)


# ---- structural properties -----------------------------------------------------------------------

@pytest.mark.parametrize("threads", [2, 3, 8])
def test_threaded_equals_simple(threads):
    s = scenes.fractal(8)
    name, pos, d, up, fov = scenes.CAMERAS[1]
    cam = yvo.camera(pos, d, up, fov, 96, 70)
    a = yvo.render(s.nodes(), s.GetRoot(), cam, threads=1, want_visits=True)
    b = yvo.render(s.nodes(), s.GetRoot(), cam, threads=threads, want_visits=True)
    for k in ("node", "child", "t", "rgba", "visits"):
        assert (a[k] == b[k]).all()
    assert a["stats"] == b["stats"]


def test_row_bands_compose():
    s = scenes.fractal(8)
    name, pos, d, up, fov = scenes.CAMERAS[2]
    cam = yvo.camera(pos, d, up, fov, 64, 50)
    full = yvo.render(s.nodes(), s.GetRoot(), cam)
    top = yvo.render(s.nodes(), s.GetRoot(), cam, rows=(0, 21))
    bot = yvo.render(s.nodes(), s.GetRoot(), cam, rows=(21, 50))
    assert (top["rgba"][:21] == full["rgba"][:21]).all() and (bot["rgba"][21:] == full["rgba"][21:]).all()
    assert (top["rgba"][21:] == 0).all()


def test_hit_records_are_consistent_with_the_pool():
    s = scenes.fractal(8)
    nodes = s.nodes()
    name, pos, d, up, fov = scenes.CAMERAS[1]
    cam = yvo.camera(pos, d, up, fov, 80, 80)
    r = yvo.render(nodes, s.GetRoot(), cam, want_visits=True)
    hit = r["node"] != yvo.MISS_NODE
    assert hit.sum() > 500
    n, c = r["node"][hit], r["child"][hit]
    assert (n < len(nodes)).all() and ((c >= 0) & (c < 8)).all()
    assert ((nodes["flags"][n] >> c) & 1).all()                # every reported child is a leaf slot
    assert (r["t"][hit] > 0).all()                             # eye outside every sphere here
    assert (r["visits"][hit] >= 1).all()
    assert r["stats"]["node_visits"] == int(r["visits"].sum())


def test_unpack_normal_is_unit_and_roundtrips():
    rng = np.random.RandomState(0)
    for _ in range(200):
        n = rng.randn(3)
        n /= np.linalg.norm(n)
        d = yv.pack_voxdata(10, 20, 30, *n)
        m = yvo.unpack_normal(d)
        assert abs(np.linalg.norm(m) - 1) < 1e-6
        assert np.dot(m, n) > 0.999                            # 8+8-bit octahedral: < 2.6 degrees


def test_shade_lambert_head_light():
    d = yv.pack_voxdata(255, 255, 255, 0, 0, -1)
    # head-on: n.L = 1 -> k = 1.0 ; RGB565 white expands to 255
    assert yvo.shade(d, (0, 0, 1), 1.0, (0.5, 0.5, -1.0), (0.5, 0.5, -1.0)) == (255, 255, 255, 255)
    # facing away: ambient only -> floor(255*0.1+0.5) = 26
    d2 = yv.pack_voxdata(255, 255, 255, 0, 0, 1)
    assert yvo.shade(d2, (0, 0, 1), 1.0, (0.5, 0.5, -1.0), (0.5, 0.5, -1.0)) == (26, 26, 26, 255)
    # shadowed (visibility 0) equals ambient as well
    assert yvo.shade(d, (0, 0, 1), 1.0, (0.5, 0.5, -1.0), (0.5, 0.5, -1.0), 0.0) == (26, 26, 26, 255)


# ---- brute force: 3-D grid marcher over the dense voxel grid (independent algorithm) --------------

def _march(vox, origins, dirs, max_steps=400):
    """Amanatides-Woo grid traversal in float64 over a [z][y][x] grid spanning the unit cube.
    Returns (value, t_enter) of the first non-zero voxel per ray (0 when none)."""
    n = vox.shape[0]
    o = np.asarray(origins, np.float64)
    d = np.asarray(dirs, np.float64)
    d = np.where(np.abs(d) < 1e-6, np.copysign(1e-6, d), d)
    inv = 1.0 / d
    ta, tb = (0.0 - o) * inv, (1.0 - o) * inv
    tmin = np.minimum(ta, tb).max(axis=1)
    tmax = np.maximum(ta, tb).min(axis=1)
    alive = (tmin < tmax) & (tmax > 0)
    t = np.maximum(tmin, 0.0)
    p = o + d * (t[:, None] + 1e-9)
    cell = np.clip(np.floor(p * n).astype(np.int64), 0, n - 1)
    step = np.where(d > 0, 1, -1)
    nxt = (cell + (d > 0)) / n
    tnext = (nxt - o) * inv
    tdelta = np.abs(inv) / n
    val = np.zeros(len(o), np.uint32)
    tent = np.zeros(len(o), np.float64)
    for _ in range(max_steps):
        if not alive.any():
            break
        v = vox[cell[:, 2], cell[:, 1], cell[:, 0]]
        found = alive & (v != 0)
        val[found] = v[found]
        tent[found] = t[found]
        alive &= ~found
        ax = np.argmin(tnext, axis=1)
        idx = np.arange(len(o))
        t = np.where(alive, tnext[idx, ax], t)
        cell[idx, ax] += np.where(alive, step[idx, ax], 0)
        tnext[idx, ax] += tdelta[idx, ax]
        alive &= (cell[idx, ax] >= 0) & (cell[idx, ax] < n)
        cell = np.clip(cell, 0, n - 1)
    return val, tent


@pytest.mark.parametrize("cam_idx", [2, 5, 6])
@pytest.mark.parametrize("depth,fill", [(4, 0.08), (5, 0.03)])
def test_octree_descent_matches_grid_marcher(cam_idx, depth, fill):
    svo, vox = scenes.dense_random(depth, fill)
    nodes = svo.nodes()
    name, pos, d, up, fov = scenes.CAMERAS[cam_idx]
    if name == "wide_up":
        pos = (-0.3, 1.4, -0.4)                     # keep the eye outside the cube for this check
    if name == "axis":
        pos = (0.503, 0.4987, -0.7)                 # off the x = y = 0.5 voxel planes: the centre row/column
                                                    # (d = 0 -> AdjustDir) would otherwise graze them by construction
    W = H = 72
    cam = yvo.camera(pos, d, up, fov, W, H)
    r = yvo.render(nodes, svo.GetRoot(), cam)
    d0, du, dv = yvo.init_ray_dir(cam)
    ys, xs = np.mgrid[0:H, 0:W]
    dirs = d0[None, None, :] + du[None, None, :] * xs[..., None] + dv[None, None, :] * ys[..., None]
    dirs = dirs / np.linalg.norm(dirs, axis=2, keepdims=True)
    val, tent = _march(vox, np.tile(np.asarray(pos, np.float64), (W * H, 1)), dirs.reshape(-1, 3))
    val, tent = val.reshape(H, W), tent.reshape(H, W)
    hit = r["node"] != yvo.MISS_NODE
    oracle_val = np.zeros((H, W), np.uint32)
    oracle_val[hit] = nodes["child"][r["node"][hit], r["child"][hit]]
    agree = oracle_val == val
    # two float pipelines may disagree only on rays that graze a voxel edge
    assert agree.mean() > 0.995, (name, agree.mean())
    assert hit.sum() > 100
    both = hit & (val != 0) & agree
    assert np.allclose(r["t"][both], tent[both], rtol=1e-4, atol=1e-5)
