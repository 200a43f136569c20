"""ctypes binding of tests/emu/libyv_emu.so: the kernel's per-ray header compiled for the host.
Test infrastructure only — used to check the explicit-stack state machine without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def lib():
    global _lib
    if _lib is None:
        d = os.path.join(ROOT, "tests", "emu")
        subprocess.check_call(["make", "-s", "-C", d])
        L = C.CDLL(os.path.join(d, "libyv_emu%s.so" % os.environ.get("YVE_VARIANT", "")))
        vp, f3 = C.c_void_p, C.POINTER(C.c_float)
        L.yve_render.argtypes = [vp, vp, C.c_int, f3, f3, f3, f3, f3, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_uint32, C.c_float, C.c_float, vp, vp, vp, vp,
                                 C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
        L.yve_set_mode.argtypes = [C.c_int]
        L.yve_set_lod.argtypes = [C.c_float, C.c_void_p]
        L.yve_render.restype = C.c_int
        L.yve_trips.restype = C.c_uint64
        L.yve_fast_levels.restype = C.c_uint64
        _lib = L
    return _lib


def render(records, leaves, root_valid, pos, dir0, du, dv, light, width, height,
           shadow=0, ao_samples=0, seed=1, voxel_size=0.0, ao_max_t=0.05, mode=2, detail=0.0, node_data=None):
    records = np.ascontiguousarray(records, np.uint32)
    leaves = np.ascontiguousarray(leaves, np.uint32)
    n = width * height
    node = np.zeros(n, np.uint32); child = np.zeros(n, np.int32); t = np.zeros(n, np.float32)
    rgba = np.zeros(n, np.uint32)
    fetches, max_sp, visits = C.c_uint64(), C.c_int(), C.c_uint64()
    lib().yve_set_mode(int(mode))
    nd = np.ascontiguousarray(node_data, np.uint32) if node_data is not None else None
    lib().yve_set_lod(float(detail), nd.ctypes.data_as(C.c_void_p) if nd is not None else None)
    v = lambda a: (C.c_float * 3)(*[float(x) for x in a])
    rc = lib().yve_render(records.ctypes.data_as(C.c_void_p), leaves.ctypes.data_as(C.c_void_p), int(root_valid),
                          v(pos), v(dir0), v(du), v(dv), v(light), width, height, shadow, ao_samples, seed,
                          voxel_size, ao_max_t, node.ctypes.data_as(C.c_void_p), child.ctypes.data_as(C.c_void_p),
                          t.ctypes.data_as(C.c_void_p), rgba.ctypes.data_as(C.c_void_p),
                          C.byref(fetches), C.byref(max_sp), C.byref(visits))
    assert rc == 0
    return dict(node=node.reshape(height, width), child=child.reshape(height, width), t=t.reshape(height, width),
                rgba=rgba.view(np.uint8).reshape(height, width, 4), fetches=fetches.value, max_sp=max_sp.value, visits=visits.value,
                trips=int(lib().yve_trips()), fast_levels=int(lib().yve_fast_levels()))
