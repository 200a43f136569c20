"""Multi-GPU frame partitioning and gather — one process per GPU.

The read-only SVO is replicated on every GPU; the image (or a batch of flythrough frames) is
partitioned across ranks and the pixels land in a buffer on GPU 0. The Cell renderer's split of
16x16 blocks over SPEs (cell/spu_renderer.cpp:73-87, cell/spu/trace_spu.cpp:162-177) is the
precedent; there is no reduction and no halo, only the gather of disjoint pixels.

Two gather paths:
  * "p2p": rank 0 allocates the destination with cudaMalloc and exports a CUDA IPC handle; every
    other rank maps it and its render kernel stores final RGBA8 pixels straight into GPU 0's memory
    over NVLink (the "collective" is fused into the kernel's last store).
  * "nccl": every rank renders locally and torch.distributed.gather moves the tiles (baseline).
torch.distributed is used only for the handle exchange, barriers and the NCCL baseline.
"""
import ctypes as C
import os

import numpy as np

from . import api


def row_band(rank, world, height, align=8):
    """Rows [y0, y1) of rank `rank`: bands of equal size rounded up to `align` rows (the kernel's
    tile height), the last band(s) absorbing the remainder. Bands are disjoint and cover [0,height)."""
    per = -(-height // world)
    per = -(-per // align) * align
    y0 = min(height, rank * per)
    y1 = min(height, (rank + 1) * per)
    return y0, y1


def interleaved_rows(rank, world, height, band_rows=32):
    """Row indices rank `rank` renders under yv_set_interleave(band_rows, world, rank): blocks of
    `band_rows` rows dealt round-robin (load balance when cost varies down the image)."""
    rows = np.arange(height)
    return rows[(rows // band_rows) % world == rank]


class DeviceBuffer:
    """A cudaMalloc'ed buffer owned by libyv_b200 (exportable over CUDA IPC)."""

    def __init__(self, device, nbytes):
        self.device, self.nbytes = int(device), int(nbytes)
        p = C.c_void_p()
        api._check(api.lib().yv_device_alloc(self.device, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def export(self):
        h = (C.c_uint8 * 64)()
        api._check(api.lib().yv_ipc_export(C.c_void_p(self.ptr), h))
        return bytes(h)

    def to_host(self, offset=0, nbytes=None):
        nbytes = self.nbytes - offset if nbytes is None else nbytes
        out = np.empty(nbytes, np.uint8)
        api._check(api.lib().yv_copy_to_host(self.device, out.ctypes.data_as(C.c_void_p),
                                             C.c_void_p(self.ptr + offset), nbytes))
        return out

    def free(self):
        if self.ptr:
            api.lib().yv_device_free(self.device, C.c_void_p(self.ptr))
            self.ptr = 0


class PeerMapping:
    """Another process's DeviceBuffer mapped into this process (cudaIpcOpenMemHandle)."""

    def __init__(self, device, handle):
        p = C.c_void_p()
        api._check(api.lib().yv_ipc_open(int(device), (C.c_uint8 * 64).from_buffer_copy(handle), C.byref(p)))
        self.ptr = p.value

    def close(self):
        if self.ptr:
            api.lib().yv_ipc_close(C.c_void_p(self.ptr))
            self.ptr = 0


def open_gather_target(dist, rank, world, device, nbytes):
    """Collective: rank 0 allocates `nbytes` on its GPU, everyone gets a device pointer to it.
    Returns (ptr, owner_or_mapping). Raises YVError if IPC is not available."""
    if rank == 0:
        buf = DeviceBuffer(device, nbytes)
        payload = [buf.export()]
    else:
        buf, payload = None, [None]
    if world > 1:
        dist.broadcast_object_list(payload, src=0)
    if rank == 0:
        return buf.ptr, buf
    m = PeerMapping(device, payload[0])
    return m.ptr, m


class SharedHostFrame:
    """One host frame (or batch of frames) shared by the ranks of a node: a /dev/shm segment mapped by every process and
    registered with its GPU (yv_host_register), so that each GPU stores its share of the pixels straight into host
    memory over its own PCIe link — the counterpart, for a host consumer, of the NVLink gather into GPU 0
    (the SPEs' DMA of finished blocks into the PPU's colour buffer, cell/spu/trace_spu.cpp:171-176).
    Collective over `dist` (any backend; dist=None for a single process). `register=False` maps without a GPU."""

    def __init__(self, dist, rank, world, device, nbytes, tag, register=True):
        import mmap
        self.path = "/dev/shm/yv_frame_%s.bin" % tag
        self.nbytes, self.rank, self.registered = int(nbytes), rank, False
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(self.nbytes)
        if dist is not None and world > 1:
            dist.barrier()
        self._f = open(self.path, "r+b")
        self._mm = mmap.mmap(self._f.fileno(), self.nbytes)
        self.array = np.frombuffer(self._mm, dtype=np.uint8)
        self.host_ptr = self.array.ctypes.data
        self.ptr = self.host_ptr
        if register:
            d = C.c_void_p()
            api._check(api.lib().yv_host_register(int(device), C.c_void_p(self.host_ptr), self.nbytes, C.byref(d)))
            self.ptr, self.registered = d.value, True
        if dist is not None and world > 1:
            dist.barrier()
        if rank == 0:
            os.unlink(self.path)              # every rank holds its mapping; the name is no longer needed

    def close(self):
        if self.registered:
            api.lib().yv_host_unregister(C.c_void_p(self.host_ptr))
            self.registered = False
        self.array = None
        try:
            self._mm.close()
        except BufferError:
            pass
        self._f.close()
