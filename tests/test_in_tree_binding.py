"""The drop-in claim, executed: the reference's own driver, cell/main.cpp, compiled UNMODIFIED and linked against
integration/b200_renderer.cpp + libyv_b200.so (oracle/Makefile, target `ref` -> oracle/_ref/cell_main_b200).

main.cpp loads "../data/scene.vox" with the tree's SVOData::Load, asks for `CreateSPURenderer()` (the factory it calls
when built for the Cell's PPU, cell/main.cpp:51-53 — the binding supplies it, so the B200 takes the SPEs' place), runs
testRenderer (cell/main.cpp:21-40: SetScene, 1024x768, eye (0.5,0.5,0.3), direction (-1,-1,-1.5), RenderFrame) and writes
the frame through Magick++ (here a raw-pixel stand-in). The frame must equal the oracle's bit for bit.

Built only where /root/reference exists; the binary travels to the GPU box with the repo."""
import os
import subprocess

import numpy as np
import pytest

import conftest
import yvo
import yoxel_voxel_b200 as yv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "cell_main_b200")
W, H = 1024, 768                                         # cell/main.cpp:24
POS, DIR = (0.5, 0.5, 0.3), (-1.0, -1.0, -1.5)           # cell/main.cpp:25-26


CELL_APP = os.path.join(ROOT, "oracle", "_ref", "cell_main_spu")


def _binary(path=BIN):
    if not os.path.exists(path):
        if not os.path.exists("/root/reference/cell/main.cpp"):
            pytest.skip("oracle/_ref binaries are built only where /root/reference exists")
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/" + os.path.basename(path)])
    return path


def _run(tmp_path, svo, binary=BIN, env=None):
    cell, data = tmp_path / "cell", tmp_path / "data"
    cell.mkdir(); data.mkdir()
    le = str(data / "scene_le.vox")
    svo.Save(le)
    # built with TARGET_PPU, svodata.h byte-swaps every word it reads (:44-47): give it the big-endian file a PPU expects
    np.fromfile(le, dtype="<u4").astype(">u4").tofile(str(data / "scene.vox"))
    p = subprocess.run([_binary(binary)], cwd=str(cell), capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, **(env or {})))
    _run.stdout = p.stdout
    assert p.returncode == 0, p.stderr
    assert "Loading ../data/scene.vox" in p.stdout and "time:" in p.stdout          # svodata.h:33, main.cpp:31
    raw = open(str(cell / "test_spu.jpg"), "rb").read()                             # main.cpp:53
    head, _, body = raw.partition(b"\n")
    tag, w, h = head.split()
    assert tag == b"YVRGBA"
    return int(w), int(h), body


def _scene():
    return yv.SVOData.SingleSphere(7, (30, 30, 8), 20)


def test_reference_main_runs_and_reports_no_frame_without_a_gpu(tmp_path):
    if conftest.has_gpu():
        pytest.skip("GPU present")
    w, h, body = _run(tmp_path, _scene())
    assert (w, h, len(body)) == (0, 0, 0)            # no CPU fallback: RenderFrame() is NULL, main.cpp writes nothing


@pytest.mark.gpu
def test_reference_main_renders_through_the_b200(tmp_path):
    svo = _scene()
    w, h, body = _run(tmp_path, svo)
    assert (w, h) == (W, H) and len(body) == W * H * 4
    img = np.frombuffer(body, np.uint8).reshape(H, W, 4)
    o = yvo.render(svo.nodes(), svo.GetRoot(), yvo.camera(POS, DIR, (0, 0, 1), 70.0, W, H), threads=8)
    assert (o["node"] != yvo.MISS_NODE).sum() > 10000
    assert (img == o["rgba"]).all()


@pytest.mark.parametrize("spes", [1, 4, 6])
def test_the_whole_cell_application_on_the_host(tmp_path, spes):
    """oracle/_ref/cell_main_spu: cell/main.cpp + cell/spu_renderer.cpp + cell/spu/trace_spu.cpp, all unmodified — the
    reference's complete application, its SPEs played by the host CPU (oracle/ref_shim/spe/libspe2.h). It loads the
    scene, splits the 64x48 blocks of the 1024x768 frame over `spes` SPEs in an interleaved pattern
    (cell/spu_renderer.cpp:80-83, cell/spu/trace_spu.cpp:164) and writes the frame: identical to the oracle's, misses in
    the SPU program's clear colour (0,0,0,255) (:127); the node fetches it reports add up to the oracle's node visits."""
    svo = _scene()
    w, h, body = _run(tmp_path, svo, CELL_APP, {"YV_SHIM_SPES": str(spes)})
    assert (w, h) == (W, H) and len(body) == W * H * 4
    img = np.frombuffer(body, np.uint8).reshape(H, W, 4)
    o = yvo.render(svo.nodes(), svo.GetRoot(), yvo.camera(POS, DIR, (0, 0, 1), 70.0, W, H), threads=8)
    hit = o["node"] != yvo.MISS_NODE
    expect = o["rgba"].copy()
    expect[~hit] = (0, 0, 0, 255)
    assert hit.sum() > 10000 and (img == expect).all()
    fetch_lines = [l for l in _run.stdout.splitlines() if l.startswith("fetch:")]      # trace_spu.cpp:179, cumulative
    assert len(fetch_lines) == spes
    assert int(fetch_lines[-1].split()[1]) == o["stats"]["node_visits"]
