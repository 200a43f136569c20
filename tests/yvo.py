"""ctypes binding of the CPU oracle (oracle/libyv_oracle.so). Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NODE_DTYPE = np.dtype([("flags", "<u4"), ("data", "<u4"), ("child", "<u4", (8,))])
MISS_NODE = 0x80000000


class Light(C.Structure):
    _fields_ = [("enabled", C.c_int32), ("pos", C.c_float * 3), ("diffuse", C.c_float * 3),
                ("specular", C.c_float * 3), ("attenuation", C.c_float * 3)]


class Camera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("dir", C.c_float * 3), ("up", C.c_float * 3),
                ("fov_deg", C.c_float), ("width", C.c_int32), ("height", C.c_int32), ("detail_coef", C.c_float),
                ("show_normals", C.c_int32), ("lights", Light * 4),
                ("ssna", C.c_int32), ("ssna_voxel_size", C.c_float),
                ("jitter_amp", C.c_float), ("jitter_seed", C.c_uint32)]


class RayDir(C.Structure):
    _fields_ = [("dir0", C.c_float * 3), ("du", C.c_float * 3), ("dv", C.c_float * 3)]


class Secondary(C.Structure):
    _fields_ = [("shadow", C.c_int32), ("ao_samples", C.c_int32), ("seed", C.c_uint32),
                ("light_pos", C.c_float * 3), ("voxel_size", C.c_float), ("ao_max_t", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("node_visits", C.c_uint64), ("iterations", C.c_uint64), ("hits", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ROOT, "oracle", "libyv_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
        L = C.CDLL(path)
        vp = C.c_void_p
        L.yvo_init_ray_dir.argtypes = [C.POINTER(Camera), C.POINTER(RayDir)]
        L.yvo_init_ray_dir.restype = None
        L.yvo_render.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(Camera), C.POINTER(Secondary),
                                 C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, C.POINTER(Stats)]
        L.yvo_render.restype = C.c_int
        L.yvo_render_threaded_ref.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(Camera), vp]
        L.yvo_render_threaded_ref.restype = C.c_int
        L.yvo_trace_ray.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                    C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_float)]
        L.yvo_trace_ray.restype = C.c_int
        L.yvo_shade.argtypes = [C.c_uint32, C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float),
                                C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_uint8)]
        L.yvo_shade.restype = None
        L.yvo_unpack_normal.argtypes = [C.c_uint32, C.POINTER(C.c_float)]
        L.yvo_unpack_normal.restype = None
        L.yvo_set_tie_order.argtypes = [C.c_int]
        L.yvo_set_tie_order.restype = None
        L.yvo_blur_taps.argtypes = [C.POINTER(C.c_float)]
        L.yvo_blur_taps.restype = None
        L.yvo_blur_z.argtypes = [C.POINTER(Camera), vp, vp]
        L.yvo_blur_z.restype = C.c_int
        L.yvo_ssna_normal.argtypes = [C.POINTER(Camera), vp, C.c_int32, C.c_int32, C.POINTER(C.c_float)]
        L.yvo_ssna_normal.restype = C.c_int
        L.yvo_load_vox.argtypes = [C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(vp)]
        L.yvo_load_vox.restype = C.c_int
        L.yvo_free.argtypes = [vp]
        L.yvo_free.restype = None
        _lib = L
    return _lib


def camera(pos, dir, up=(0, 0, 1), fov=70.0, width=64, height=64, detail_coef=0.0, lights=None, show_normals=False,
           ssna=False, ssna_voxel_size=0.0, jitter_amp=0.0, jitter_seed=1):
    c = Camera()
    c.pos[:] = [float(v) for v in pos]
    c.dir[:] = [float(v) for v in dir]
    c.up[:] = [float(v) for v in up]
    c.fov_deg = float(fov)
    c.width = int(width)
    c.height = int(height)
    c.detail_coef = float(detail_coef)
    c.show_normals = 1 if show_normals else 0
    c.ssna = 1 if ssna else 0
    c.ssna_voxel_size = float(ssna_voxel_size)
    c.jitter_amp = float(jitter_amp)
    c.jitter_seed = int(jitter_seed)
    for i, lt in enumerate(lights or []):       # dicts: pos, diffuse, specular, attenuation (enabled implied)
        c.lights[i].enabled = 1 if lt.get("enabled", True) else 0
        c.lights[i].pos[:] = [float(v) for v in lt["pos"]]
        c.lights[i].diffuse[:] = [float(v) for v in lt.get("diffuse", (0.7, 0.7, 0.7))]
        c.lights[i].specular[:] = [float(v) for v in lt.get("specular", (0.3, 0.3, 0.3))]
        c.lights[i].attenuation[:] = [float(v) for v in lt.get("attenuation", (1, 0, 0.5))]
    return c


def secondary(shadow=0, ao_samples=0, seed=1, light_pos=(0.5, 0.5, 1.0), voxel_size=0.0, ao_max_t=0.05):
    s = Secondary()
    s.shadow = int(shadow)
    s.ao_samples = int(ao_samples)
    s.seed = int(seed)
    s.light_pos[:] = [float(v) for v in light_pos]
    s.voxel_size = float(voxel_size)
    s.ao_max_t = float(ao_max_t)
    return s


def init_ray_dir(cam):
    rd = RayDir()
    lib().yvo_init_ray_dir(C.byref(cam), C.byref(rd))
    return (np.array(rd.dir0[:], np.float32), np.array(rd.du[:], np.float32), np.array(rd.dv[:], np.float32))


def render(nodes, root, cam, sec=None, threads=1, rows=None, want_visits=False):
    """Returns dict(node, child, t, rgba, visits, stats) for the full frame (rows outside `rows` untouched)."""
    nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
    W, H = cam.width, cam.height
    node = np.full(W * H, MISS_NODE, np.uint32)
    child = np.full(W * H, -1, np.int32)
    t = np.zeros(W * H, np.float32)
    rgba = np.zeros(W * H * 4, np.uint8)
    visits = np.zeros(W * H, np.uint32) if want_visits else None
    st = Stats()
    y0, y1 = rows if rows else (0, H)
    rc = lib().yvo_render(nodes.ctypes.data_as(C.c_void_p), len(nodes), int(root), C.byref(cam),
                          C.byref(sec) if sec is not None else None, y0, y1, int(threads),
                          node.ctypes.data_as(C.c_void_p), child.ctypes.data_as(C.c_void_p),
                          t.ctypes.data_as(C.c_void_p), rgba.ctypes.data_as(C.c_void_p),
                          visits.ctypes.data_as(C.c_void_p) if want_visits else None, C.byref(st))
    assert rc == 0
    return dict(node=node.reshape(H, W), child=child.reshape(H, W), t=t.reshape(H, W),
                rgba=rgba.reshape(H, W, 4), visits=visits.reshape(H, W) if want_visits else None,
                stats=dict(rays=st.rays, node_visits=st.node_visits, iterations=st.iterations, hits=st.hits))


def render_threaded_ref(nodes, root, cam, prefill=0):
    nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
    rgba = np.full(cam.width * cam.height * 4, prefill, np.uint8)
    rc = lib().yvo_render_threaded_ref(nodes.ctypes.data_as(C.c_void_p), len(nodes), int(root), C.byref(cam),
                                       rgba.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return rgba.reshape(cam.height, cam.width, 4)


def trace_ray(nodes, root, pos, d):
    nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
    p = (C.c_float * 3)(*[float(v) for v in pos])
    dd = (C.c_float * 3)(*[float(v) for v in d])
    n, c, t = C.c_uint32(), C.c_int32(), C.c_float()
    hit = lib().yvo_trace_ray(nodes.ctypes.data_as(C.c_void_p), len(nodes), int(root), p, dd,
                              C.byref(n), C.byref(c), C.byref(t))
    return bool(hit), n.value, c.value, t.value


def spu_cache_model(nodes, root, cam):
    """(fetches, misses) of the SPU program's 2048-entry direct-mapped node cache over one run (trace_spu.cpp:15-35)."""
    nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
    f, m = C.c_uint64(), C.c_uint64()
    L = lib()
    L.yvo_spu_cache_model.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    rc = L.yvo_spu_cache_model(nodes.ctypes.data_as(C.c_void_p), len(nodes), int(root), C.byref(cam), C.byref(f), C.byref(m))
    assert rc == 0
    return f.value, m.value


def set_tie_order(order):
    """Test-only: 0 = the path's GoNext tie order (default), 1 = the scalar prototype's (cell/spu/vector.h:45-59)."""
    lib().yvo_set_tie_order(int(order))


def shade(data, d, t, viewer, light, visibility=1.0):
    out = (C.c_uint8 * 4)()
    lib().yvo_shade(int(data), (C.c_float * 3)(*d), float(t), (C.c_float * 3)(*viewer), (C.c_float * 3)(*light),
                    float(visibility), out)
    return tuple(out[:])


def unpack_normal(data):
    n = (C.c_float * 3)()
    lib().yvo_unpack_normal(int(data), n)
    return np.array(n[:], np.float32)


def blur_taps():
    k = (C.c_float * 49)()
    lib().yvo_blur_taps(k)
    return np.array(k[:], np.float32).reshape(7, 7)


def blur_z(cam, z0):
    z0 = np.ascontiguousarray(z0, np.float32)
    out = np.zeros_like(z0)
    rc = lib().yvo_blur_z(C.byref(cam), z0.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


def ssna_normal(cam, z, x, y):
    """World-space SSNA normal at (x, y) of a blurred z-buffer, or None where it is undefined."""
    z = np.ascontiguousarray(z, np.float32)
    n = (C.c_float * 3)()
    ok = lib().yvo_ssna_normal(C.byref(cam), z.ctypes.data_as(C.c_void_p), int(x), int(y), n)
    return np.array(n[:], np.float32) if ok else None


def load_vox(path):
    root, count, ptr = C.c_uint32(), C.c_uint32(), C.c_void_p()
    rc = lib().yvo_load_vox(os.fsencode(path), C.byref(root), C.byref(count), C.byref(ptr))
    if rc != 0:
        raise IOError("yvo_load_vox failed: %d" % rc)
    arr = np.zeros(count.value, NODE_DTYPE)
    if count.value:
        C.memmove(arr.ctypes.data, ptr.value, count.value * 40)
    lib().yvo_free(ptr)
    return root.value, arr
