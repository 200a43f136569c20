"""_ore — the reference's Python extension module, on libyv_b200.

The reference builds `_ore` from ore/src/main.cpp with Boost.Python over C++ sources that are not in the snapshot
(DynamicSVO.h, builders.h). This module exports the same names with the same call signatures (main.cpp:37-75 the
Make*Source factories, :84-130 the module definition), implemented over this repo's C ABI, so the reference's scene
scripts run against the library unchanged apart from their Python-2 print statements
(tests/test_ore_compat.py runs gen_spheres.py that way).

  point_3i, point_3f ................. main.cpp:84-93
  BuildMode.GROW / CLEAR ............. :101-103
  VoxelSource.GetSize / GetPivot ..... :105-107
  MakeRawSource(size, colors, normals) :37-52, 109-110
  MakeSphereSource(radius, color, inverted) :69-72, 112-113
  MakeIsoSource(size, data) + SetIsoLevel / SetInside / SetColor :54-67, 115-119
  DynamicSVO: BuildRange, Save, Load, TraceRay, nodecount, CountChangedPages, CountTransfrerSize,
              GetNodeCountByLevel1 ... :121-129
"""
import numpy as np

import yoxel_voxel_b200 as _yv

__all__ = ["point_3i", "point_3f", "BuildMode", "VoxelSource", "RawSource", "SphereSource", "IsoSource",
           "MakeRawSource", "MakeSphereSource", "MakeIsoSource", "DynamicSVO"]

_PAGE_BYTES = 256 * 40          # 256-node pages of 40-byte nodes (reaction/report/main.tex:71)


class point_3i:
    def __init__(self, x, y, z):
        self.x, self.y, self.z = int(x), int(y), int(z)

    def __iter__(self):
        return iter((self.x, self.y, self.z))

    def __repr__(self):
        return "point_3i(%d, %d, %d)" % (self.x, self.y, self.z)


class point_3f:
    def __init__(self, x, y, z):
        self.x, self.y, self.z = float(x), float(y), float(z)

    def __iter__(self):
        return iter((self.x, self.y, self.z))

    def __repr__(self):
        return "point_3f(%g, %g, %g)" % (self.x, self.y, self.z)


BuildMode = _yv.BuildMode


class VoxelSource:
    """Base of the three sources; `_source()` is the library-side object handed to BuildRange."""

    def _source(self):
        raise NotImplementedError

    def GetSize(self):
        return point_3i(*self._source().GetSize())

    def GetPivot(self):
        return point_3i(*self._source().GetPivot())


class RawSource(VoxelSource):
    def __init__(self, size, colors, normals):
        n = size.x * size.y * size.z * 4
        col = np.frombuffer(colors, dtype=np.uint8) if not isinstance(colors, np.ndarray) else colors
        nrm = np.frombuffer(normals, dtype=np.int8) if not isinstance(normals, np.ndarray) else normals
        if col.size != n or nrm.size != n:
            raise ValueError("incorrect data buffer size")                    # main.cpp:48-49
        self._src = _yv.MakeRawSource((size.x, size.y, size.z), col, nrm)

    def _source(self):
        return self._src


class SphereSource(VoxelSource):
    def __init__(self, radius, color, inverted):
        self._src = _yv.MakeSphereSource(int(radius), tuple(int(c) for c in color), bool(inverted))

    def _source(self):
        return self._src


class IsoSource(VoxelSource):
    """The reference's IsoSource is configured after construction (SetIsoLevel / SetInside / SetColor); the library's is
    configured at construction, so the object is (re)made when a setting has changed since its last use."""

    def __init__(self, size, data):
        d = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        if d.size != size.x * size.y * size.z:
            raise ValueError("incorrect data buffer size")                    # main.cpp:62-63
        self._data = np.ascontiguousarray(d, np.uint8).reshape(size.z, size.y, size.x)
        self._iso, self._inside, self._color = 128, False, (200, 200, 200)
        self._src = None

    def SetIsoLevel(self, level):
        self._iso, self._src = int(level), None

    def SetInside(self, inside):
        self._inside, self._src = bool(inside), None

    def SetColor(self, color):
        self._color, self._src = tuple(int(c) for c in color), None

    def _source(self):
        if self._src is None:
            self._src = _yv.MakeIsoSource(self._data, iso_level=self._iso, inside=self._inside, color=self._color)
        return self._src


def MakeRawSource(size, colors, normals):
    return RawSource(size, colors, normals)


def MakeSphereSource(radius, color, inverted):
    return SphereSource(radius, color, inverted)


def MakeIsoSource(size, data):
    return IsoSource(size, data)


class DynamicSVO:
    def __init__(self):
        self._svo = _yv.DynamicSVO()
        self._seen_version = 0
        self._renderer = None

    def BuildRange(self, level, pos, mode, src):
        self._svo.BuildRange(int(level), (pos.x, pos.y, pos.z), mode, src._source())

    def Save(self, fn):
        self._svo.Save(fn)

    def Load(self, fn):
        self._svo.Load(fn)
        self._seen_version = 0
        return True

    @property
    def nodecount(self):
        return self._svo.livenodes

    def GetNodeCountByLevel1(self):
        return self._svo.GetNodeCountByLevel1()

    def CountChangedPages(self):
        """Pages written since the previous call of this method (qtview.py:72 polls it after every edit)."""
        n = self._svo.CountChangedPages(self._seen_version)
        self._pending_pages = n
        self._seen_version = self._svo.version
        return n

    def CountTransfrerSize(self):                       # the reference's spelling (main.cpp:128)
        return getattr(self, "_pending_pages", 0) * _PAGE_BYTES

    def TraceRay(self, pos, direction):
        """Distance to the first voxel along the ray (qtview.py:66-67 uses it as pos + dir * t); runs on the GPU —
        there is no CPU tracer in the product, so this raises without one."""
        if self._renderer is None:
            self._renderer = _yv.SVORenderer(0)
        self._renderer.SetScene(self._svo)
        node, child, t = self._renderer.TraceRays([tuple(pos)], [tuple(direction)])
        return float(t[0])

    # what this repo's own tools need from the object
    def nodes(self):
        return self._svo.nodes()

    def GetRoot(self):
        return self._svo.GetRoot()
