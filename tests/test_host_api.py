"""CPU-side checks of the product's host code: the C-ABI library loads and exports every symbol the
header declares, scene builders / .vox I/O / repack behave, and the renderer refuses to run without
a GPU (no CPU fallback). No kernel is launched here."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

import conftest
import scenes
import yvo
import yoxel_voxel_b200 as yv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "yv_b200.h")).read()
    declared = set(re.findall(r"\b(yv_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 45
    L = C.CDLL(yv.lib_path())
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    # and the Python mirror binds all of them
    assert declared <= set(yv.lib()._yv_signatures), declared - set(yv.lib()._yv_signatures)
    assert yv.lib().yv_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    if conftest.has_gpu():
        pytest.skip("GPU present")
    with pytest.raises(yv.YVError) as e:
        yv.SVORenderer(0)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "yoxel-voxel_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle/" not in text and "yv_oracle" not in text and "yvo_" not in text, f


def test_builder_is_deterministic_and_valid():
    a = yv.SVOData.SphereFractal(8, threads=1).nodes()
    b = yv.SVOData.SphereFractal(8, threads=8).nodes()
    assert a.tobytes() == b.tobytes()
    assert len(a) == 35645
    # flags are consistent with the child words (main.tex:40-42)
    leaf = (a["flags"][:, None] >> np.arange(8)) & 1
    null = (a["flags"][:, None] >> (8 + np.arange(8))) & 1
    topbit = (a["child"] >> 31) & 1
    assert ((null == 1) == ((leaf == 0) & (topbit == 1))).all()
    inner = (leaf == 0) & (null == 0)
    assert (a["child"][inner] < len(a)).all()
    # every internal node has at least one non-null child (uniform octets collapse)
    assert ((leaf | inner).sum(axis=1) > 0).all()


def test_sphere_fractal_follows_gen_spheres():
    """gen_spheres.py:8-32 at depth 8: only lev-5 spheres survive (radius 32/2^5 = 1), 5^5 of them,
    colour (128,128,5*255/8=159) -> RGB565."""
    s = yv.SVOData.SphereFractal(8)
    nodes = s.nodes()
    leaf = ((nodes["flags"][:, None] >> np.arange(8)) & 1).astype(bool)
    data = nodes["child"][leaf]
    assert len(np.unique(data & 0xFFFF)) == 1
    assert (data[0] & 0xFFFF) == ((128 >> 3) << 11 | (128 >> 2) << 5 | (159 >> 3))
    assert s.depth == 8


def test_vox_roundtrip_matches_svodata_layout(tmp_path):
    s = scenes.single_sphere(6)
    fn = str(tmp_path / "sphere.vox")
    s.Save(fn)
    raw = open(fn, "rb").read()
    root, w1, w2, count = np.frombuffer(raw[:16], "<u4")
    assert root == s.GetRoot() and count == s.nodecount and len(raw) == 16 + 40 * count     # svodata.h:36-42
    # the oracle's SVOData::Load restatement reads the same pool
    oroot, onodes = yvo.load_vox(fn)
    assert oroot == s.GetRoot() and onodes.tobytes() == s.nodes().tobytes()
    t = yv.SVOData().Load(fn)
    assert t.GetRoot() == s.GetRoot() and t.nodes().tobytes() == s.nodes().tobytes() and t.depth == 6


def test_load_errors_are_reported(tmp_path):
    with pytest.raises(yv.YVError) as e:
        yv.SVOData().Load(str(tmp_path / "missing.vox"))
    assert e.value.code == -2
    fn = tmp_path / "short.vox"
    fn.write_bytes(np.array([0, 0, 0, 5], "<u4").tobytes() + b"\0" * 40)
    with pytest.raises(yv.YVError):
        yv.SVOData().Load(str(fn))
    bad = np.zeros(1, yv.NODE_DTYPE)
    bad[0]["child"][:] = yv.EMPTY_NODE
    bad[0]["child"][3] = 77                                  # dangling child id
    with pytest.raises(yv.YVError) as e:
        yv.SVOData.FromNodes(0, bad)
    assert e.value.code == -3


def test_packed_pool_invariants():
    s = scenes.fractal(8)
    nodes = s.nodes()
    recs, leaves = s.packed()
    assert recs.dtype == np.uint32 and recs.shape[1] == 4              # 16-byte records
    assert len(recs) == len(nodes)                                     # a tree: no duplication
    assert sorted(recs[:, 3]) == list(range(len(nodes)))               # orig ids are a permutation
    assert recs[0, 3] == s.GetRoot()
    leaf_mask, child_mask = recs[:, 2] & 0xFF, (recs[:, 2] >> 8) & 0xFF
    assert (leaf_mask & child_mask == 0).all()
    pc = np.array([bin(i).count("1") for i in range(256)])
    # children of consecutive records are laid out back to back, breadth first
    assert recs[0, 0] == 1
    assert (np.diff(recs[:, 0].astype(np.int64)) == pc[child_mask][:-1]).all()
    assert (np.diff(recs[:, 1].astype(np.int64)) == pc[leaf_mask][:-1]).all()
    assert recs[-1, 1] + pc[leaf_mask][-1] == len(leaves)
    # spot-check content: leaf words and child links agree with the reference pool
    rng = np.random.RandomState(1)
    for i in rng.randint(0, len(recs), 300):
        nd = nodes[recs[i, 3]]
        k_leaf = k_child = 0
        for c in range(8):
            if (nd["flags"] >> c) & 1:
                assert leaves[recs[i, 1] + k_leaf] == nd["child"][c]; k_leaf += 1
                assert (leaf_mask[i] >> c) & 1
            elif not nd["child"][c] & 0x80000000:
                assert recs[recs[i, 0] + k_child, 3] == nd["child"][c]; k_child += 1
                assert (child_mask[i] >> c) & 1


def test_packed_pool_of_null_root_is_empty():
    s = yv.SVOData.FromNodes(yv.EMPTY_NODE, np.zeros(0, yv.NODE_DTYPE))
    recs, leaves = s.packed()
    assert len(recs) == 0 and len(leaves) == 0


def test_init_ray_dir_matches_oracle_bitwise():
    for name, pos, d, up, fov in scenes.CAMERAS:
        for (w, h) in [(512, 512), (1920, 1080), (37, 23), (1, 1)]:
            cam = yvo.camera(pos, d, up, fov, w, h)
            o = yvo.init_ray_dir(cam)
            p = yv.init_ray_dir(d, up, fov, w, h)
            for a, b in zip(o, p):
                assert a.tobytes() == b.tobytes(), (name, w, h)


def test_iso_volume_builder_small():
    s = yv.SVOData.IsoVolume(7, seed=219, iso_level=200, threads=4)
    nodes = s.nodes()
    assert 1000 < len(nodes) < 200000
    assert hashlib.sha256(nodes.tobytes()).hexdigest() == \
        hashlib.sha256(yv.SVOData.IsoVolume(7, seed=219, iso_level=200, threads=1).nodes().tobytes()).hexdigest()
    # a camera above the slab sees terrain
    cam = yvo.camera((0.5, 0.5, 0.6), (0.3, 0.4, -1), (0, 0, 1), 70, 64, 64)
    r = yvo.render(nodes, s.GetRoot(), cam)
    assert (r["node"] != yvo.MISS_NODE).mean() > 0.5


def test_cpp_adapter_builds_and_refuses_without_gpu(tmp_path):
    """include/yv_renderer.hpp (ISVORenderer-shaped adapter) + tools/render_main.cpp (cell/main.cpp:21-56)."""
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tools")])
    exe = os.path.join(ROOT, "tools", "render_main")
    assert os.path.exists(exe)
    if conftest.has_gpu():
        pytest.skip("GPU present: covered by the gpu test")
    p = subprocess.run([exe, "--fractal", "8", str(tmp_path / "o.ppm"), "64", "48"], capture_output=True, text=True)
    assert p.returncode == 2 and "no CPU fallback" in p.stderr       # RenderFrame() == NULL, like the reference


def test_host_register_rejects_bad_arguments():
    import ctypes as C
    d = C.c_void_p()
    assert yv.lib().yv_host_register(0, None, 4096, C.byref(d)) == -1            # YV_ERR_ARG, before any CUDA call
    buf = (C.c_uint8 * 4096)()
    assert yv.lib().yv_host_register(0, C.cast(buf, C.c_void_p), 0, C.byref(d)) == -1
    assert yv.lib().yv_host_unregister(None) == 0


# ---- untrusted pools and files ------------------------------------------------------------------
def _cyclic_pool():
    cyc = np.zeros(2, yv.NODE_DTYPE)
    cyc[0]["child"][:] = [1, 0, 1, 0, 1, 0, 1, 0]
    cyc[1]["child"][:] = [0, 1, 0, 1, 0, 1, 0, 1]
    return cyc


def test_cyclic_pool_is_rejected_without_exhausting_memory():
    """A two-node pool whose nodes are each other's children would grow the breadth-first frontier eight-fold per level;
    the repack must find the cycle first and fail with a format error, quickly."""
    import time
    svo = yv.SVOData.FromNodes(0, _cyclic_pool())
    t0 = time.time()
    with pytest.raises(yv.YVError) as e:
        svo.packed()
    assert e.value.code == -3 and "cyclic" in str(e.value) and time.time() - t0 < 5.0


def test_shared_subtrees_are_still_duplicated():
    """A DAG (one sub-tree referenced twice) is legal: the repack expands it."""
    leaf = yv.pack_voxdata(10, 200, 30, 0, 1, 0)
    nodes = np.zeros(2, yv.NODE_DTYPE)
    nodes["child"][:] = yv.EMPTY_NODE
    nodes[0]["child"][3] = leaf; nodes[0]["flags"] = 1 << 3
    nodes[1]["child"][0] = 0; nodes[1]["child"][5] = 0
    svo = yv.SVOData.FromNodes(1, nodes)
    recs, leaves = svo.packed()
    assert recs.shape[0] == 3 and leaves.shape[0] == 2 and (recs[1:, 3] == 0).all()


def test_vox_header_count_is_checked_against_the_file(tmp_path):
    """svodata.h:40-42 trusts the header's node count; a 16-byte file claiming 4 G nodes must not allocate 160 GB."""
    fn = tmp_path / "liar.vox"
    np.array([0, 0x5956, 3, 0xFFFFFFF0], "<u4").tofile(str(fn))
    with pytest.raises(yv.YVError) as e:
        yv.SVOData().Load(str(fn))
    assert "truncated" in str(e.value)


def test_load_reloads_in_place(tmp_path):
    """SVOData::Load on a loaded object replaces the pool inside the same handle (cell/svodata.h:31-50)."""
    a, b = scenes.fractal(7), scenes.single_sphere(6)
    fa, fb = str(tmp_path / "a.vox"), str(tmp_path / "b.vox")
    a.Save(fa); b.Save(fb)
    svo = yv.SVOData().Load(fa)
    h, v0 = svo._h.value, svo.version
    assert svo.nodes().tobytes() == a.nodes().tobytes()
    svo.Load(fb)
    assert svo._h.value == h and svo.version > v0 and svo.nodes().tobytes() == b.nodes().tobytes()
    assert svo.packed()[0].shape[0] == b.packed()[0].shape[0]
    with pytest.raises(yv.YVError):
        svo.Load(str(tmp_path / "missing.vox"))
    assert svo.nodes().tobytes() == b.nodes().tobytes()           # a failed reload keeps the scene


def test_multi_device_renderer_needs_gpus():
    if conftest.has_gpu():
        pytest.skip("GPU present")
    for make in (lambda: yv.SVORenderer(devices="all"), lambda: yv.SVORenderer(devices=[0, 0])):
        with pytest.raises(yv.YVError) as e:
            make()
        assert e.value.code == -4 and "no CPU fallback" in str(e.value)


def test_octant_masks_follow_their_definition():
    """Byte c of a record's 64-bit octant mask = (leaf flags | child flags) of its child NODE c, 0 for leaf / empty slots
    (svo_pack.h) — what the culling traversal relies on."""
    for svo in (scenes.fractal(8), scenes.dense_random(5, 0.05)[0], scenes.single_sphere(6)):
        recs, _ = svo.packed()
        g = svo.octant_masks()
        assert g.shape[0] == recs.shape[0]
        occ = ((recs[:, 2] | (recs[:, 2] >> 8)) & 0xFF).astype(np.uint64)
        child_mask = (recs[:, 2] >> 8) & 0xFF
        want = np.zeros(len(recs), np.uint64)
        rank = np.zeros(len(recs), np.int64)
        for c in range(8):
            has = ((child_mask >> c) & 1).astype(bool)
            idx = recs[:, 0].astype(np.int64) + rank
            want[has] |= occ[idx[has]] << np.uint64(8 * c)
            rank += has
        assert (g == want).all()
        assert (g != 0).sum() > 0 or len(recs) == 1
