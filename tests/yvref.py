"""ctypes binding of oracle/_ref: the reference's own sources for this path, compiled unmodified (oracle/Makefile,
target `ref`; stand-ins for the headers the snapshot lacks under oracle/ref_shim/). TEST INFRASTRUCTURE ONLY.

  libppu_renderer_ref.so ..... cell/ppu_renderer.cpp: SVOData::Load, InitRayDir, RenderRect, RecTrace,
                               SimpleRenderer / TreadedRenderer behind ISVORenderer
  libtrace_spu_f32_ref.so .... cell/spu/trace_spu.cpp: the SPU program (node cache, FindFirstChildSPU, GoNextSPU, ...)

They exist only where /root/reference does (this container); tests that need them skip elsewhere, and what they
produced is committed as tests/golden/reference_golden.npz (make_reference_golden.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
PPU_SO = os.path.join(REF_DIR, "libppu_renderer_ref.so")
SPU_SO = os.path.join(REF_DIR, "libtrace_spu_f32_ref.so")
PROBE_SHADE, PROBE_T, PROBE_DATA = 0, 1, 2          # oracle/ref_shim/cpp/shader.h
SPU_BLOCK = 16                                      # BlockSize, cell/spu/trace_spu.h:5

_ppu = _spu = None
_f3 = C.POINTER(C.c_float)


def available():
    if os.path.exists(PPU_SO) and os.path.exists(SPU_SO):
        return True
    if not os.path.exists("/root/reference/cell/ppu_renderer.cpp"):
        return False
    try:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"),
                               "_ref/libppu_renderer_ref.so", "_ref/libtrace_spu_f32_ref.so"])
    except (subprocess.CalledProcessError, OSError):
        return False
    return os.path.exists(PPU_SO) and os.path.exists(SPU_SO)


def _v(a):
    return (C.c_float * 3)(*[float(x) for x in a]) if a is not None else None


def ppu():
    global _ppu
    if _ppu is None:
        L = C.CDLL(PPU_SO)
        L.yv_ref_scene_load.restype = C.c_void_p
        L.yv_ref_scene_load.argtypes = [C.c_char_p]
        L.yv_ref_scene_root.restype = C.c_uint
        L.yv_ref_scene_root.argtypes = [C.c_void_p]
        L.yv_ref_scene_free.argtypes = [C.c_void_p]
        L.yv_ref_ppu_render.restype = C.c_int
        L.yv_ref_ppu_render.argtypes = [C.c_void_p, C.c_int, _f3, _f3, _f3, C.c_float, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.yv_ref_init_ray_dir.argtypes = [_f3, _f3, C.c_float, C.c_int, C.c_int, _f3]
        _ppu = L
    return _ppu


def spu():
    global _spu
    if _spu is None:
        L = C.CDLL(SPU_SO)
        L.yv_ref_spu_render.restype = C.c_int
        L.yv_ref_spu_render.argtypes = [C.c_void_p, C.c_uint, C.c_uint, _f3, _f3, _f3, _f3, C.c_int, C.c_int, C.c_int,
                                        C.c_uint, C.c_void_p, C.POINTER(C.c_int)]
        _spu = L
    return _spu


class Scene:
    """SVOData::Load (cell/svodata.h:31-50) of a .vox file."""

    def __init__(self, path):
        self.h = ppu().yv_ref_scene_load(os.fsencode(path))

    def root(self):
        return ppu().yv_ref_scene_root(self.h)

    def close(self):
        if self.h:
            ppu().yv_ref_scene_free(self.h)
            self.h = None

    def __del__(self):
        self.close()


def init_ray_dir(d, up, fov, width, height):
    """RendererBase::InitRayDir (cell/renderer_base.h:50-61) -> (dir0, du, dv)."""
    out = (C.c_float * 9)()
    ppu().yv_ref_init_ray_dir(_v(d), _v(up), float(fov), int(width), int(height), out)
    a = np.array(out[:], np.float32)
    return a[0:3], a[3:6], a[6:9]


def ppu_frame(scene, pos, d, up, fov, width, height, probe=PROBE_SHADE, threaded=False):
    """One RenderFrame of SimpleRenderer / TreadedRenderer. Returns uint32 [H][W] (R in the low byte), or None when
    RenderFrame returned NULL. up=None / fov=0 / width=0 leave the renderer's defaults in place."""
    w, h = C.c_int(), C.c_int()
    n = (width * height) if width > 0 else 640 * 480
    out = np.zeros(n, np.uint32)
    ok = ppu().yv_ref_ppu_render(scene.h if scene is not None else None, int(threaded), _v(pos), _v(d), _v(up), float(fov),
                                 int(width), int(height), int(probe), out.ctypes.data_as(C.c_void_p), C.byref(w), C.byref(h))
    if not ok:
        return None
    return out.reshape(h.value, w.value)


def ppu_result(scene, pos, d, up, fov, width, height, threaded=False):
    """The three frames that carry a TraceResult out of the reference's loop: shaded RGBA8 [H][W][4], bits of t, VoxData."""
    rgba = ppu_frame(scene, pos, d, up, fov, width, height, PROBE_SHADE, threaded)
    t = ppu_frame(scene, pos, d, up, fov, width, height, PROBE_T, threaded)
    data = ppu_frame(scene, pos, d, up, fov, width, height, PROBE_DATA, threaded)
    return rgba.view(np.uint8).reshape(rgba.shape[0], rgba.shape[1], 4), t, data


def spu_frame(nodes, root, pos, dir0, du, dv, width, height, probe=PROBE_SHADE, fill=0):
    """One run of the SPU program over the whole frame (blockStart 0, blockStride 1). Returns (uint32 [H][W],
    node fetches, cache misses); pixels outside the 16x16 blocks it covers keep `fill`."""
    nodes = np.ascontiguousarray(nodes)
    out = np.zeros(width * height, np.uint32)
    st = (C.c_int * 2)()
    ok = spu().yv_ref_spu_render(nodes.ctypes.data_as(C.c_void_p), len(nodes), int(root), _v(pos), _v(dir0), _v(du), _v(dv),
                                 int(width), int(height), int(probe), int(fill), out.ctypes.data_as(C.c_void_p), st)
    assert ok
    return out.reshape(height, width), st[0], st[1]
