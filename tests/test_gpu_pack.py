"""The GPU breadth-first repack (svo_pack_gpu.cu) must produce exactly the arrays of the host repack (svo_pack.cpp)."""
import os

import numpy as np
import pytest

import scenes
import yoxel_voxel_b200 as yv

pytestmark = pytest.mark.gpu


def _same(svo, check_data=True):
    recs, leaves = svo.packed()                 # host BFS
    drecs, dleaves, dnd = svo.device_packed(0)  # GPU BFS (default upload path)
    assert drecs.shape == recs.shape and dleaves.shape == leaves.shape
    assert (drecs == recs).all() and (dleaves == leaves).all()
    nodes = svo.nodes()
    assert (svo.octant_masks(0) == svo.octant_masks()).all()          # the culling traversal's occupancy words
    if len(recs) and check_data:       # (the host path uploads VoxNode::data only when the LOD cut-off is first used)
        assert (dnd == nodes["data"][recs[:, 3]]).all()


def test_gpu_pack_equals_host_pack():
    for svo in (scenes.fractal(10), scenes.single_sphere(6), scenes.dense_random(5, 0.03)[0],
                yv.SVOData.IsoVolume(9, threads=8), yv.SVOData.FromNodes(yv.EMPTY_NODE, np.zeros(0, yv.NODE_DTYPE))):
        _same(svo)


def test_gpu_pack_follows_edits_and_host_fallback(monkeypatch):
    svo = yv.SVOData.SphereFractal(9)
    _same(svo)
    svo.BuildRange(9, (256, 256, 300), yv.BuildMode.CLEAR, yv.MakeSphereSource(30, (200, 180, 120), True))
    svo.BuildRange(9, (200, 260, 280), yv.BuildMode.GROW, yv.MakeSphereSource(12, (20, 220, 40), False))
    _same(svo)                                   # version changed -> re-packed on the device
    # a pool with a shared sub-tree (a DAG) is not a tree: the device repack declines, the host repack duplicates it
    nodes, root, leaf = scenes.two_level_tree()
    nodes[1]["child"][3] = 0                     # node 0 referenced twice
    nodes[1]["flags"] &= ~np.uint32(1 << (8 + 3))
    dag = yv.SVOData.FromNodes(root, nodes)
    drecs, dleaves, _ = dag.device_packed(0)
    recs, leaves = dag.packed()
    assert len(recs) == 3 and (drecs == recs).all() and (dleaves == leaves).all()
    # forced host path gives the same device arrays
    monkeypatch.setenv("YV_HOST_PACK", "1")
    svo2 = yv.SVOData.SphereFractal(9)
    _same(svo2, check_data=False)
