"""Multi-GPU host logic on CPU: the screen partition and the tile gather, world_size 2 over gloo.
(The data path has no collective besides the gather of disjoint pixels — SURVEY §8e; on the GPU box
the gather is fused into the render kernel's stores, tests/test_gpu_parity.py covers band rendering.)"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import scenes
import yvo
from yoxel_voxel_b200 import multigpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("height", [1, 7, 8, 67, 1080, 2160, 4320])
def test_row_bands_are_a_partition(world, height):
    bands = [multigpu.row_band(r, world, height) for r in range(world)]
    covered = np.zeros(height, int)
    for y0, y1 in bands:
        assert 0 <= y0 <= y1 <= height
        assert y0 % 8 == 0 or y0 == height          # bands start on the kernel's tile rows
        covered[y0:y1] += 1
    assert (covered == 1).all()


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("height,band", [(1080, 32), (4320, 64), (67, 16), (16, 16)])
def test_interleaved_blocks_are_a_partition(world, height, band):
    covered = np.zeros(height, int)
    for r in range(world):
        rows = multigpu.interleaved_rows(r, world, height, band)
        covered[rows] += 1
    assert (covered == 1).all()


def _worker(rank, world, port, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    svo = scenes.fractal(8)
    name, pos, d, up, fov = scenes.CAMERAS[1]
    W, H = 96, 70
    cam = yvo.camera(pos, d, up, fov, W, H)
    # every rank holds a replica of the scene and renders only its band (oracle stands in for the GPU here)
    y0, y1 = multigpu.row_band(rank, world, H)
    band = yvo.render(svo.nodes(), svo.GetRoot(), cam, rows=(y0, y1))["rgba"]
    local = torch.from_numpy(band.copy())
    gathered = [torch.zeros_like(local) for _ in range(world)] if rank == 0 else None
    dist.gather(local, gathered, dst=0)
    if rank == 0:
        frame = np.zeros((H, W, 4), np.uint8)
        for r in range(world):
            a, b = multigpu.row_band(r, world, H)
            frame[a:b] = gathered[r].numpy()[a:b]
        full = yvo.render(svo.nodes(), svo.GetRoot(), cam)["rgba"]
        np.save(os.path.join(tmpdir, "ok.npy"), np.array([(frame == full).all()]))
    dist.barrier()
    dist.destroy_process_group()


def test_band_gather_world2_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert np.load(tmp_path / "ok.npy")[0]


def _shared_frame_worker(rank, world, port, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    svo = scenes.fractal(8)
    name, pos, d, up, fov = scenes.CAMERAS[1]
    W, H, band = 96, 70, 16
    cam = yvo.camera(pos, d, up, fov, W, H)
    # the host side of the multi-GPU e2e path: one frame in shared memory, every rank writes its own row blocks
    # straight into it (on the GPU box the writer is the render kernel, through yv_host_register; here the oracle)
    shared = multigpu.SharedHostFrame(dist, rank, world, 0, W * H * 4, "test%d" % port, register=False)
    frame = shared.array.reshape(H, W, 4)
    rows = multigpu.interleaved_rows(rank, world, H, band)
    full = yvo.render(svo.nodes(), svo.GetRoot(), cam)["rgba"]
    frame[rows] = full[rows]
    dist.barrier()
    if rank == 0:
        np.save(os.path.join(tmpdir, "shared_ok.npy"), np.array([(frame == full).all(), not os.path.exists(shared.path)]))
    dist.barrier()
    del frame
    shared.close()
    dist.destroy_process_group()


def test_shared_host_frame_world2_gloo(tmp_path):
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_shared_frame_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok = np.load(tmp_path / "shared_ok.npy")
    assert ok[0] and ok[1]
