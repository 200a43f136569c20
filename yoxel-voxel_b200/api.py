"""Host-side mirror of the reference renderer interface over the C ABI (include/yv_b200.h).

Class and method names follow the reference: ``SVOData`` (cell/svodata.h:22-55),
``ISVORenderer`` setters and ``RenderFrame`` (cell/svorenderer.h:5-24), the CUDA renderer's
``Render(d_dstBuf)`` / ``SetViewSize`` (demo/SVORenderer.h:8-64) and ``DynamicSVO.TraceRay`` /
``Save`` / ``nodecount`` (ore/src/main.cpp:119-129). Everything that computes goes through
``libyv_b200.so``; there is no Python or CPU implementation behind these classes, and loading
fails loudly when the library has not been built.
"""
import ctypes as C
import os

import numpy as np

__all__ = [
    "YVError", "lib", "lib_path", "SVOData", "SVORenderer", "CreateB200Renderer",
    "pack_voxdata", "device_count", "init_ray_dir", "BuildMode", "VoxelSource",
    "MakeSphereSource", "MakeRawSource", "MakeIsoSource", "DynamicSVO", "LightParams", "CudaRenderer", "NODE_DTYPE", "EMPTY_NODE", "FULL_NODE", "ALL_DEVICES",
]

EMPTY_NODE = 0x80000000
FULL_NODE = 0x80000001
ALL_DEVICES = 0xFFFFFFFFFFFFFFFF     # YV_ALL_DEVICES

# VoxNode (reaction/report/main.tex:46-51): 40 bytes
NODE_DTYPE = np.dtype([("flags", "<u4"), ("data", "<u4"), ("child", "<u4", (8,))])
assert NODE_DTYPE.itemsize == 40


class LightParams(C.Structure):
    """LightParams (demo/Demo.cpp:141-147) == yv_light."""
    _fields_ = [("enabled", C.c_int32), ("pos", C.c_float * 3), ("diffuse", C.c_float * 3),
                ("specular", C.c_float * 3), ("attenuationCoefs", C.c_float * 3)]

    def __init__(self, enabled=True, pos=(0, 0, 0), diffuse=(0.7, 0.7, 0.7), specular=(0.3, 0.3, 0.3),
                 attenuationCoefs=(1, 0, 0.5)):
        super().__init__()
        self.enabled = 1 if enabled else 0
        self.pos[:] = [float(v) for v in pos]
        self.diffuse[:] = [float(v) for v in diffuse]
        self.specular[:] = [float(v) for v in specular]
        self.attenuationCoefs[:] = [float(v) for v in attenuationCoefs]


class YVError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("yv_b200 error %d: %s" % (code, msg))
        self.code = code


def lib_path():
    """In-tree library; YV_B200_LIB selects an ablation build (csrc/Makefile `variant`) for sweeps."""
    override = os.environ.get("YV_B200_LIB")
    if override:
        return override if os.path.isabs(override) else os.path.join(os.path.dirname(os.path.abspath(__file__)), override)
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libyv_b200.so")


_lib = None


def lib():
    """Load libyv_b200.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise YVError(-100, "libyv_b200.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "or `make -C yoxel-voxel_b200/csrc` (there is no fallback implementation)")
    L = C.CDLL(path)
    vp, i32, u32, f32 = C.c_void_p, C.c_int, C.c_uint32, C.c_float
    P = C.POINTER
    sig = {
        "yv_last_error": (C.c_char_p, []),
        "yv_abi_version": (i32, []),
        "yv_svo_load": (i32, [C.c_char_p, P(vp)]),
        "yv_svo_load_into": (i32, [vp, C.c_char_p]),
        "yv_svo_replicate": (i32, [vp, i32, i32]),
        "yv_renderer_create_multi": (i32, [C.c_uint64, P(vp)]),
        "yv_renderer_create_group": (i32, [P(i32), i32, P(vp)]),
        "yv_renderer_device_count": (i32, [vp]),
        "yv_renderer_device": (i32, [vp, i32]),
        "yv_set_partition": (i32, [vp, i32, i32]),
        "yv_member_frame_ms": (f32, [vp, i32]),
        "yv_replicate_stats": (i32, [vp, P(C.c_double), P(C.c_uint64)]),
        "yv_render_frame_async": (i32, [vp, vp, P(i32)]),
        "yv_wait_frame": (i32, [vp, i32, P(vp)]),
        "yv_svo_from_memory": (i32, [u32, vp, u32, P(vp)]),
        "yv_svo_save": (i32, [vp, C.c_char_p]),
        "yv_svo_free": (None, [vp]),
        "yv_svo_root": (u32, [vp]),
        "yv_svo_node_count": (u32, [vp]),
        "yv_svo_depth": (u32, [vp]),
        "yv_svo_nodes": (vp, [vp]),
        "yv_svo_build_sphere_fractal": (i32, [i32, i32, P(vp)]),
        "yv_svo_build_iso_volume": (i32, [i32, u32, i32, i32, P(vp)]),
        "yv_svo_build_single_sphere": (i32, [i32, i32, i32, i32, i32, C.c_uint8, C.c_uint8, C.c_uint8, P(vp)]),
        "yv_svo_build_from_dense": (i32, [i32, vp, P(vp)]),
        "yv_pack_voxdata": (u32, [C.c_uint8, C.c_uint8, C.c_uint8, f32, f32, f32]),
        "yv_svo_create": (i32, [P(vp)]),
        "yv_source_sphere": (i32, [i32, C.c_uint8, C.c_uint8, C.c_uint8, i32, P(vp)]),
        "yv_source_raw": (i32, [P(i32), vp, P(vp)]),
        "yv_source_raw_colors_normals": (i32, [P(i32), vp, vp, P(vp)]),
        "yv_source_iso": (i32, [P(i32), vp, i32, i32, C.c_uint8, C.c_uint8, C.c_uint8, P(vp)]),
        "yv_source_free": (None, [vp]),
        "yv_source_size": (i32, [vp, P(i32), P(i32)]),
        "yv_svo_build_range": (i32, [vp, i32, P(i32), i32, vp]),
        "yv_svo_live_node_count": (u32, [vp]),
        "yv_svo_node_count_by_level": (i32, [vp, P(i32), i32]),
        "yv_svo_version": (u32, [vp]),
        "yv_svo_count_changed_pages": (i32, [vp, u32]),
        "yv_svo_update": (i32, [vp, i32, P(C.c_uint64)]),
        "yv_svo_upload": (i32, [vp, i32]),
        "yv_svo_device_bytes": (C.c_uint64, [vp, i32]),
        "yv_svo_device_packed_copy": (i32, [vp, i32, P(u32), P(u32), vp, vp, vp]),
        "yv_svo_packed_counts": (i32, [vp, P(u32), P(u32)]),
        "yv_svo_packed_copy": (i32, [vp, vp, vp]),
        "yv_svo_octant_masks": (i32, [vp, vp]),
        "yv_svo_device_octant_masks": (i32, [vp, i32, vp]),
        "yv_renderer_create": (i32, [i32, P(vp)]),
        "yv_renderer_destroy": (None, [vp]),
        "yv_set_scene": (i32, [vp, vp]),
        "yv_set_view_pos": (i32, [vp, P(f32)]),
        "yv_set_view_dir": (i32, [vp, P(f32)]),
        "yv_set_view_up": (i32, [vp, P(f32)]),
        "yv_set_resolution": (i32, [vp, i32, i32]),
        "yv_get_resolution": (i32, [vp, P(i32), P(i32)]),
        "yv_set_fov": (i32, [vp, f32]),
        "yv_get_fov": (i32, [vp, P(f32)]),
        "yv_set_light": (i32, [vp, i32, vp]),
        "yv_set_show_normals": (i32, [vp, i32]),
        "yv_get_show_normals": (i32, [vp, P(i32)]),
        "yv_set_ssna": (i32, [vp, i32]),
        "yv_get_ssna": (i32, [vp, P(i32)]),
        "yv_set_ssna_voxel_size": (i32, [vp, f32]),
        "yv_set_jitter": (i32, [vp, f32, u32]),
        "yv_render_accumulated": (i32, [vp, i32, P(vp)]),
        "yv_set_detail_coef": (i32, [vp, f32]),
        "yv_get_detail_coef": (i32, [vp, P(f32)]),
        "yv_render_frame": (i32, [vp, P(vp)]),
        "yv_render_frame_device": (i32, [vp, vp]),
        "yv_render_frame_device_async": (i32, [vp, vp]),
        "yv_sync": (i32, [vp]),
        "yv_device_framebuffer": (i32, [vp, P(vp)]),
        "yv_set_rows": (i32, [vp, i32, i32]),
        "yv_set_interleave": (i32, [vp, i32, i32, i32]),
        "yv_set_secondary": (i32, [vp, i32, i32, u32, P(f32), f32, f32]),
        "yv_enable_hits": (i32, [vp, i32]),
        "yv_get_hits": (i32, [vp, vp, vp, vp]),
        "yv_dump_trace_data": (i32, [vp, C.c_char_p]),
        "yv_enable_counters": (i32, [vp, i32]),
        "yv_get_counters": (i32, [vp, vp]),
        "yv_last_frame_ms": (f32, [vp]),
        "yv_last_frame_launches": (i32, [vp]),
        "yv_set_stream": (i32, [vp, vp]),
        "yv_set_option": (i32, [vp, C.c_char_p, i32]),
        "yv_get_option": (i32, [vp, C.c_char_p, P(i32)]),
        "yv_trace_rays": (i32, [vp, vp, vp, u32, vp, vp, vp]),
        "yv_device_alloc": (i32, [i32, C.c_size_t, P(vp)]),
        "yv_device_free": (i32, [i32, vp]),
        "yv_copy_to_host": (i32, [i32, vp, vp, C.c_size_t]),
        "yv_ipc_export": (i32, [vp, vp]),
        "yv_ipc_open": (i32, [i32, vp, P(vp)]),
        "yv_ipc_close": (i32, [vp]),
        "yv_host_register": (i32, [i32, vp, C.c_size_t, P(vp)]),
        "yv_host_unregister": (i32, [vp]),
        "yv_init_ray_dir": (i32, [P(f32), P(f32), f32, i32, i32, P(f32), P(f32), P(f32)]),
        "yv_device_count": (i32, []),
        "yv_device_name": (i32, [i32, C.c_char_p, C.c_size_t]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._yv_signatures = sig
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise YVError(rc, lib().yv_last_error().decode("utf-8", "replace"))


def _vec3(v):
    a = (C.c_float * 3)(*[float(x) for x in v])
    return a


def pack_voxdata(r, g, b, nx, ny, nz):
    return int(lib().yv_pack_voxdata(int(r), int(g), int(b), float(nx), float(ny), float(nz)))


def init_ray_dir(view_dir, up, fov, width, height):
    """RendererBase::InitRayDir (cell/renderer_base.h:50-61) -> (dir0, du, dv) float32 arrays."""
    d0, du, dv = (C.c_float * 3)(), (C.c_float * 3)(), (C.c_float * 3)()
    _check(lib().yv_init_ray_dir(_vec3(view_dir), _vec3(up), float(fov), int(width), int(height), d0, du, dv))
    return (np.array(d0[:], np.float32), np.array(du[:], np.float32), np.array(dv[:], np.float32))


def device_count():
    return int(lib().yv_device_count())


class BuildMode:
    """enum BuildMode (ore/src/main.cpp:101-103)"""
    GROW = 0
    CLEAR = 1


class VoxelSource:
    """VoxelSource handle (ore/src/main.cpp:106-117)."""

    def __init__(self, handle, keepalive=None):
        self._h = handle
        self._keep = keepalive

    def _sizes(self):
        size, pivot = (C.c_int * 3)(), (C.c_int * 3)()
        _check(lib().yv_source_size(self._h, size, pivot))
        return tuple(size), tuple(pivot)

    def GetSize(self):
        return self._sizes()[0]

    def GetPivot(self):
        return self._sizes()[1]

    def __del__(self):
        try:
            if self._h and self._h.value:
                lib().yv_source_free(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


def MakeSphereSource(radius, color, inverted=False):          # ore/src/main.cpp:69
    h = C.c_void_p()
    _check(lib().yv_source_sphere(int(radius), int(color[0]), int(color[1]), int(color[2]), 1 if inverted else 0, C.byref(h)))
    return VoxelSource(h)


def MakeRawSource(voxdata, colors=None, normals=None):         # ore/src/main.cpp:37-52
    """MakeRawSource(voxdata): array [z][y][x] of VoxData words, 0 = empty; or, as the reference's scripts call it
    (scene_gen.py:79), MakeRawSource(size, colors, normals) with size = (x, y, z), colors uint8 [z][y][x][4] (alpha 0 empty,
    255 surface, else buried) and normals int8 [z][y][x][4]."""
    if colors is not None:
        sx, sy, sz = (int(v) for v in voxdata)
        col = np.ascontiguousarray(colors, dtype=np.uint8).reshape(sz, sy, sx, 4)
        nrm = np.ascontiguousarray(normals, dtype=np.int8).reshape(sz, sy, sx, 4)
        h = C.c_void_p()
        _check(lib().yv_source_raw_colors_normals((C.c_int * 3)(sx, sy, sz), col.ctypes.data_as(C.c_void_p),
                                                  nrm.ctypes.data_as(C.c_void_p), C.byref(h)))
        return VoxelSource(h)
    vox = np.ascontiguousarray(voxdata, dtype=np.uint32)
    size = (C.c_int * 3)(vox.shape[2], vox.shape[1], vox.shape[0])
    h = C.c_void_p()
    _check(lib().yv_source_raw(size, vox.ctypes.data_as(C.c_void_p), C.byref(h)))
    return VoxelSource(h)


def MakeIsoSource(data, iso_level=128, inside=False, color=(200, 200, 200)):   # ore/src/main.cpp:54-67; uint8 [z][y][x]
    d = np.ascontiguousarray(data, dtype=np.uint8)
    size = (C.c_int * 3)(d.shape[2], d.shape[1], d.shape[0])
    h = C.c_void_p()
    _check(lib().yv_source_iso(size, d.ctypes.data_as(C.c_void_p), int(iso_level), 1 if inside else 0,
                               int(color[0]), int(color[1]), int(color[2]), C.byref(h)))
    return VoxelSource(h)


class SVOData:
    """SVOData (cell/svodata.h:22-55) plus the DynamicSVO surface the scene scripts use
    (ore/src/main.cpp:119-129: BuildRange, Save, Load, nodecount, CountChangedPages, ...)."""

    def __init__(self, handle=None):
        self._h = C.c_void_p(handle) if handle else C.c_void_p()
        if not handle:
            _check(lib().yv_svo_create(C.byref(self._h)))

    # -- editing (DynamicSVO) ----------------------------------------------------------------
    def BuildRange(self, level, pos, mode, src):          # ore/src/main.cpp:121
        p = (C.c_int * 3)(int(pos[0]), int(pos[1]), int(pos[2]))
        _check(lib().yv_svo_build_range(self._h, int(level), p, int(mode), src._h))

    def GetNodeCountByLevel1(self):                       # ore/src/main.cpp:129
        buf = (C.c_int * 40)()
        n = lib().yv_svo_node_count_by_level(self._h, buf, 40)
        return list(buf[:min(n, 40)])

    @property
    def version(self):
        return int(lib().yv_svo_version(self._h))

    def CountChangedPages(self, since_version=0):         # ore/src/main.cpp:127
        return int(lib().yv_svo_count_changed_pages(self._h, int(since_version)))

    def Update(self, device=0):
        """CudaSVO::Update: ship the pages written since the last call; returns the bytes transferred
        (CountTransfrerSize, ore/src/main.cpp:128)."""
        b = C.c_uint64()
        _check(lib().yv_svo_update(self._h, int(device), C.byref(b)))
        return int(b.value)

    @property
    def livenodes(self):
        return int(lib().yv_svo_live_node_count(self._h))

    # -- construction ------------------------------------------------------------------------
    def Load(self, fn):                                   # svodata.h:31 — reloads in place: renderers that hold this
        if self._h.value:                                 # scene (SetScene keeps the pointer) stay valid
            _check(lib().yv_svo_load_into(self._h, os.fsencode(fn)))
        else:
            _check(lib().yv_svo_load(os.fsencode(fn), C.byref(self._h)))
        return self

    def Replicate(self, src_device, dst_device):
        """Copy the packed pool from one GPU to another over NVLink instead of uploading it again."""
        _check(lib().yv_svo_replicate(self._h, int(src_device), int(dst_device)))

    @classmethod
    def _adopt(cls, handle):
        s = cls.__new__(cls)
        s._h = handle
        return s

    def Save(self, fn):                                   # ore/src/main.cpp:123
        _check(lib().yv_svo_save(self._h, os.fsencode(fn)))

    @classmethod
    def FromNodes(cls, root, nodes):
        nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
        h = C.c_void_p()
        _check(lib().yv_svo_from_memory(int(root), nodes.ctypes.data_as(C.c_void_p), len(nodes), C.byref(h)))
        return cls._adopt(h)

    @classmethod
    def SphereFractal(cls, depth, threads=0):             # gen_spheres.py
        h = C.c_void_p()
        _check(lib().yv_svo_build_sphere_fractal(int(depth), int(threads or os.cpu_count() or 1), C.byref(h)))
        return cls._adopt(h)

    @classmethod
    def IsoVolume(cls, depth, seed=219, iso_level=200, threads=0):   # gen_largevol.py
        h = C.c_void_p()
        _check(lib().yv_svo_build_iso_volume(int(depth), int(seed), int(iso_level),
                                             int(threads or os.cpu_count() or 1), C.byref(h)))
        return cls._adopt(h)

    @classmethod
    def SingleSphere(cls, depth, center, radius, color=(128, 128, 255)):
        h = C.c_void_p()
        _check(lib().yv_svo_build_single_sphere(int(depth), int(center[0]), int(center[1]), int(center[2]),
                                                int(radius), int(color[0]), int(color[1]), int(color[2]),
                                                C.byref(h)))
        return cls._adopt(h)

    @classmethod
    def FromDense(cls, vox):
        vox = np.ascontiguousarray(vox, dtype=np.uint32)
        n = vox.shape[0]
        assert vox.shape == (n, n, n) and n & (n - 1) == 0, "dense grid must be a power-of-two cube [z][y][x]"
        h = C.c_void_p()
        _check(lib().yv_svo_build_from_dense(int(n).bit_length() - 1, vox.ctypes.data_as(C.c_void_p), C.byref(h)))
        return cls._adopt(h)

    # -- accessors ---------------------------------------------------------------------------
    def GetRoot(self):                                    # svodata.h:52
        return int(lib().yv_svo_root(self._h))

    @property
    def nodecount(self):                                  # ore/src/main.cpp:126
        return int(lib().yv_svo_node_count(self._h))

    @property
    def depth(self):
        return int(lib().yv_svo_depth(self._h))

    def nodes(self, copy=True):
        """The host node pool as a structured array (reference layout). copy=False: a read-only view of the library's
        own pool (no second 20 GB for a depth-14 volume), valid until the scene is edited, reloaded or freed."""
        n = self.nodecount
        if not copy and n:
            buf = (C.c_uint8 * (n * 40)).from_address(lib().yv_svo_nodes(self._h))
            view = np.frombuffer(buf, dtype=NODE_DTYPE)
            view.flags.writeable = False
            return view
        out = np.zeros(n, dtype=NODE_DTYPE)
        if n:
            C.memmove(out.ctypes.data, lib().yv_svo_nodes(self._h), n * 40)
        return out

    def packed(self):
        nr, nl = C.c_uint32(), C.c_uint32()
        _check(lib().yv_svo_packed_counts(self._h, C.byref(nr), C.byref(nl)))
        recs = np.zeros((nr.value, 4), dtype=np.uint32)
        leaves = np.zeros(nl.value, dtype=np.uint32)
        _check(lib().yv_svo_packed_copy(self._h, recs.ctypes.data_as(C.c_void_p), leaves.ctypes.data_as(C.c_void_p)))
        return recs, leaves

    def octant_masks(self, device=None):
        """One uint64 per packed record: byte c = occupied octants of child node c (host repack, or the device's copy)."""
        nr = C.c_uint32()
        _check(lib().yv_svo_packed_counts(self._h, C.byref(nr), None))
        out = np.zeros(nr.value, np.uint64)
        if device is None:
            _check(lib().yv_svo_octant_masks(self._h, out.ctypes.data_as(C.c_void_p)))
        else:
            _check(lib().yv_svo_device_octant_masks(self._h, int(device), out.ctypes.data_as(C.c_void_p)))
        return out

    def device_packed(self, device=0):
        """The packed arrays as they sit on the device (records (n,4) u32, leaves, node_data)."""
        nr, nl = C.c_uint32(), C.c_uint32()
        _check(lib().yv_svo_device_packed_copy(self._h, int(device), C.byref(nr), C.byref(nl), None, None, None))
        recs = np.zeros((nr.value, 4), np.uint32); leaves = np.zeros(nl.value, np.uint32); nd = np.zeros(nr.value, np.uint32)
        _check(lib().yv_svo_device_packed_copy(self._h, int(device), None, None, recs.ctypes.data_as(C.c_void_p),
                                               leaves.ctypes.data_as(C.c_void_p), nd.ctypes.data_as(C.c_void_p)))
        return recs, leaves, nd

    def Upload(self, device=0):                           # CudaSVO::Update
        _check(lib().yv_svo_upload(self._h, int(device)))
        return int(lib().yv_svo_device_bytes(self._h, int(device)))

    def _release(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().yv_svo_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass


class SVORenderer:
    """ISVORenderer (cell/svorenderer.h:5-24) + the CUDA SVORenderer extras (demo/SVORenderer.h)."""

    def __init__(self, device=0, devices=None):
        """device: one GPU. devices: a list of CUDA ordinals, or "all" — ONE renderer that splits every frame's blocks
        over those GPUs (SPURenderer over every usable SPE, cell/spu_renderer.cpp:65-90)."""
        self._h = C.c_void_p()
        self._scene = None
        if devices is None:
            self.device = int(device)
            _check(lib().yv_renderer_create(self.device, C.byref(self._h)))
        else:
            if isinstance(devices, str) and devices == "all":
                _check(lib().yv_renderer_create_multi(C.c_uint64(ALL_DEVICES), C.byref(self._h)))
            else:                                          # explicit ordinals; one listed twice is driven by two members
                devs = [int(d) for d in devices]
                _check(lib().yv_renderer_create_group((C.c_int * len(devs))(*devs), len(devs), C.byref(self._h)))
            self.device = int(lib().yv_renderer_device(self._h, 0))

    def DeviceCount(self):
        return int(lib().yv_renderer_device_count(self._h))

    def Devices(self):
        return [int(lib().yv_renderer_device(self._h, k)) for k in range(self.DeviceCount())]

    def SetPartition(self, mode="interleaved", band_rows=32):
        """How a multi-device renderer splits a frame: blocks of band_rows rows dealt round-robin (the SPU program's
        block stride, cell/spu/trace_spu.cpp:164) or contiguous "bands"."""
        _check(lib().yv_set_partition(self._h, {"interleaved": 0, "bands": 1}[mode], int(band_rows)))

    def MemberFrameMs(self):
        """Device time every GPU of the group spent on its own share of the last frame."""
        return [float(lib().yv_member_frame_ms(self._h, k)) for k in range(self.DeviceCount())]

    def ReplicateStats(self):
        ms, b = C.c_double(), C.c_uint64()
        _check(lib().yv_replicate_stats(self._h, C.byref(ms), C.byref(b)))
        return float(ms.value), int(b.value)

    def RenderFrameAsync(self, dst_ptr=None):
        """Start a frame (camera as set) and return its ticket; see WaitFrame."""
        t = C.c_int()
        _check(lib().yv_render_frame_async(self._h, C.c_void_p(int(dst_ptr)) if dst_ptr else None, C.byref(t)))
        return t.value

    def WaitFrame(self, ticket, as_array=True):
        """Block until frame `ticket` is complete; returns the (H, W, 4) uint8 view of where it was delivered
        (as_array=False: the address)."""
        p = C.c_void_p()
        _check(lib().yv_wait_frame(self._h, int(ticket), C.byref(p)))
        if not as_array:
            return p.value
        w, h = self.GetResolution()
        buf = (C.c_uint8 * (w * h * 4)).from_address(p.value)
        return np.frombuffer(buf, dtype=np.uint8).reshape(h, w, 4)

    def SetScene(self, svo):                              # svorenderer.h:12
        self._scene = svo
        _check(lib().yv_set_scene(self._h, svo._h if svo is not None else None))

    def SetViewPos(self, pos):                            # :14
        _check(lib().yv_set_view_pos(self._h, _vec3(pos)))

    def SetViewDir(self, d):                              # :15
        _check(lib().yv_set_view_dir(self._h, _vec3(d)))

    def SetViewUp(self, up):                              # :16
        _check(lib().yv_set_view_up(self._h, _vec3(up)))

    def SetResolution(self, width, height):               # :18
        _check(lib().yv_set_resolution(self._h, int(width), int(height)))

    SetViewSize = SetResolution                           # demo/SVORenderer.h:20

    def GetResolution(self):                              # :19
        w, h = C.c_int(), C.c_int()
        _check(lib().yv_get_resolution(self._h, C.byref(w), C.byref(h)))
        return w.value, h.value

    def SetFOV(self, fov):                                # :21
        _check(lib().yv_set_fov(self._h, float(fov)))

    def GetFOV(self):                                     # demo/SVORenderer.h:23
        f = C.c_float()
        _check(lib().yv_get_fov(self._h, C.byref(f)))
        return f.value

    def SetLigth(self, i, lp):                            # demo/SVORenderer.h:34 (the reference's spelling)
        _check(lib().yv_set_light(self._h, int(i), C.byref(lp)))

    SetLight = SetLigth

    def SetShowNormals(self, enable):                     # demo/SVORenderer.h:31
        _check(lib().yv_set_show_normals(self._h, 1 if enable else 0))

    def GetShowNormals(self):                             # demo/SVORenderer.h:32
        v = C.c_int()
        _check(lib().yv_get_show_normals(self._h, C.byref(v)))
        return bool(v.value)

    def SetSSNA(self, enable, voxel_size=None):           # demo/SVORenderer.h:28 (voxSize: SVORenderer.cpp:129)
        _check(lib().yv_set_ssna(self._h, 1 if enable else 0))
        if voxel_size is not None:
            _check(lib().yv_set_ssna_voxel_size(self._h, float(voxel_size)))

    def GetSSNA(self):                                    # demo/SVORenderer.h:29
        v = C.c_int()
        _check(lib().yv_get_ssna(self._h, C.byref(v)))
        return bool(v.value)

    def SetJitter(self, amplitude, seed=1):               # reaction/report/main.tex:109 (displaced ray origins)
        _check(lib().yv_set_jitter(self._h, float(amplitude), int(seed)))

    def RenderAccumulated(self, frames):                  # reaction/report/main.tex:111 (mean of jittered frames)
        """HxWx4 uint8 mean of `frames` frames drawn with seeds seed, seed+1, ... (view of renderer-owned memory)."""
        px = C.c_void_p()
        _check(lib().yv_render_accumulated(self._h, int(frames), C.byref(px)))
        w, h = self.GetResolution()
        return np.ctypeslib.as_array(C.cast(px, C.POINTER(C.c_uint8)), shape=(h, w, 4))

    def SetDetailCoef(self, coef):                        # demo/SVORenderer.h:25
        _check(lib().yv_set_detail_coef(self._h, float(coef)))

    def GetDetailCoef(self):                              # demo/SVORenderer.h:26
        f = C.c_float()
        _check(lib().yv_get_detail_coef(self._h, C.byref(f)))
        return f.value

    def RenderFrame(self):
        """const Color32* RenderFrame() (:23): returns an (H, W, 4) uint8 view of the renderer-owned
        pinned host buffer, or None when no scene is set (the reference returns NULL)."""
        p = C.c_void_p()
        rc = lib().yv_render_frame(self._h, C.byref(p))
        if rc == -5:
            return None
        _check(rc)
        w, h = self.GetResolution()
        buf = (C.c_uint8 * (w * h * 4)).from_address(p.value)
        return np.frombuffer(buf, dtype=np.uint8).reshape(h, w, 4)

    def Render(self, d_dst_ptr, sync=True):               # demo/SVORenderer.h:36
        fn = lib().yv_render_frame_device if sync else lib().yv_render_frame_device_async
        _check(fn(self._h, C.c_void_p(int(d_dst_ptr))))

    def Sync(self):
        _check(lib().yv_sync(self._h))

    def DeviceFramebuffer(self):
        p = C.c_void_p()
        _check(lib().yv_device_framebuffer(self._h, C.byref(p)))
        return p.value

    def SetRows(self, y0, y1):
        _check(lib().yv_set_rows(self._h, int(y0), int(y1)))

    def SetInterleave(self, band_rows, stride, phase):
        _check(lib().yv_set_interleave(self._h, int(band_rows), int(stride), int(phase)))

    def SetSecondary(self, shadow=0, ao_samples=0, seed=1, light_pos=(0.5, 0.5, 1.0), voxel_size=0.0, ao_max_t=0.05):
        _check(lib().yv_set_secondary(self._h, int(shadow), int(ao_samples), int(seed), _vec3(light_pos),
                                      float(voxel_size), float(ao_max_t)))

    def EnableHits(self, on=True):
        _check(lib().yv_enable_hits(self._h, 1 if on else 0))

    def GetHits(self):
        w, h = self.GetResolution()
        node = np.zeros(w * h, dtype=np.uint32)
        child = np.zeros(w * h, dtype=np.int32)
        t = np.zeros(w * h, dtype=np.float32)
        _check(lib().yv_get_hits(self._h, node.ctypes.data_as(C.c_void_p), child.ctypes.data_as(C.c_void_p),
                                 t.ctypes.data_as(C.c_void_p)))
        return node.reshape(h, w), child.reshape(h, w), t.reshape(h, w)

    def DumpTraceData(self, fnbase):                      # demo/SVORenderer.cpp:158
        _check(lib().yv_dump_trace_data(self._h, os.fsencode(fnbase)))

    def EnableCounters(self, on=True):
        _check(lib().yv_enable_counters(self._h, 1 if on else 0))

    def GetCounters(self):
        w, h = self.GetResolution()
        c = np.zeros(w * h, dtype=np.uint32)
        _check(lib().yv_get_counters(self._h, c.ctypes.data_as(C.c_void_p)))
        c = c.reshape(h, w)
        return c & 0xFFFF, c >> 16

    def LastFrameMs(self):
        return float(lib().yv_last_frame_ms(self._h))

    def LastFrameLaunches(self):
        return int(lib().yv_last_frame_launches(self._h))

    def SetStream(self, cuda_stream_ptr):
        _check(lib().yv_set_stream(self._h, C.c_void_p(int(cuda_stream_ptr)) if cuda_stream_ptr else None))

    def SetOption(self, name, value):
        _check(lib().yv_set_option(self._h, name.encode(), int(value)))

    def GetOption(self, name):
        v = C.c_int()
        _check(lib().yv_get_option(self._h, name.encode(), C.byref(v)))
        return v.value

    def TraceRays(self, pos, dirs):                       # DynamicSVO::TraceRay, batched
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        n = len(pos)
        node = np.zeros(n, dtype=np.uint32)
        child = np.zeros(n, dtype=np.int32)
        t = np.zeros(n, dtype=np.float32)
        _check(lib().yv_trace_rays(self._h, pos.ctypes.data_as(C.c_void_p), dirs.ctypes.data_as(C.c_void_p), n,
                                   node.ctypes.data_as(C.c_void_p), child.ctypes.data_as(C.c_void_p),
                                   t.ctypes.data_as(C.c_void_p)))
        return node, child, t

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().yv_renderer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


DynamicSVO = SVOData     # the scripts' name for the editable scene (ore/src/main.cpp:119)


class CudaRenderer:
    """The pycuda host class of the reference (trace_cuda.py:16-118) over this library: same constructor, methods
    and `detailCoef` / `lightPos` attributes, so qtview.py-style callers keep working. `render` returns the same kind
    of statistics string; the three kernels it used to launch (InitEyeRays, Trace, ShadeSimple) are one fused launch."""

    FOV = 70.0

    def __init__(self, res=(640, 480), device=0):
        self._r = SVORenderer(device)
        self.resx, self.resy = int(res[0]), int(res[1])          # (trace_cuda.py rounds to its block size; not needed here)
        self._r.SetResolution(self.resx, self.resy)
        self._r.SetFOV(self.FOV)
        self.setLightPos((0.5, 0.5, 1))                         # trace_cuda.py:48
        self.detailCoef = 10.0                                   # trace_cuda.py:49
        self._img = None

    def updateScene(self, scene):                                # trace_cuda.py:51
        self._scene = scene
        self._r.SetScene(scene)

    def setLightPos(self, pos):                                  # trace_cuda.py:63
        self.lightPos = tuple(float(v) for v in pos)

    def getViewSize(self):                                       # trace_cuda.py:66
        return (self.resx, self.resy)

    def render(self, eyePos, viewDir, first=False):              # trace_cuda.py:69
        import time
        self._r.SetViewPos(eyePos); self._r.SetViewDir(viewDir); self._r.SetViewUp((0, 0, 1))
        # trace_cuda.py:93 passes an angular threshold of (1 degree / detailCoef) per node; the same angle in
        # SetDetailCoef's unit (rad(fov/2)/width per unit coefficient):
        ang = np.radians(1.0) / max(self.detailCoef, 1e-6)
        self._r.SetDetailCoef(float(ang * self.resx / np.radians(self.FOV / 2)))
        self._r.SetLigth(0, LightParams(True, self.lightPos, (0.9, 0.9, 0.9), (0.0, 0.0, 0.0), (1, 0, 0)))
        t = time.perf_counter()
        self._img = self._r.RenderFrame()
        gpu = time.perf_counter() - t
        stat = "gpu time: %.2f ms\n" % (gpu * 1000)
        stat += "eye trace time: %.2f ms\n" % self._r.LastFrameMs()
        stat += "detailCoef: %f\n" % (self.detailCoef)
        return stat

    def getImage(self):                                          # trace_cuda.py:117
        return None if self._img is None else self._img[..., :3].copy()      # (d_img.get() copies as well)


def CreateB200Renderer(device=0, devices=None):
    """Factory in the style of CreateSimpleRenderer / CreateThreadedRenderer / CreateSPURenderer
    (cell/svorenderer.h:26-30); devices="all" or a list = one renderer over several GPUs."""
    return SVORenderer(device, devices)
