"""Frame sink: the raw stream writer of the C++ host layer (host/ckd_sink.cpp) and a reader.

File format: 32-byte header ("CKDF", version 1, resX, resY, numFrames, 3 reserved u32), then numFrames frames of
resX*resY little-endian 0xAARRGGBB pixels; frame i starts at 32 + i*resX*resY*4."""
import ctypes as C
import struct

import numpy as np

from . import capi

HEADER_BYTES = 32
MAGIC = 0x46444B43  # "CKDF"


def _lib():
    L = capi.load()
    if not getattr(L, "_sink_bound", False):
        L.ckdsink_open.argtypes = [C.c_char_p, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_int, C.c_int]
        L.ckdsink_acquire.restype = C.c_void_p
        L.ckdsink_commit.argtypes = [C.c_void_p, C.c_uint]
        L.ckdhost_last_error.restype = C.c_char_p
        L._sink_bound = True
    return L


class Sink:
    """one per process: a ring of host frame buffers drained by a writer thread"""

    def __init__(self, path, res_x, res_y, num_frames, ring_frames=4, pinned=True, create=True):
        self.L = _lib()
        self.res_x, self.res_y = res_x, res_y
        if self.L.ckdsink_open(str(path).encode(), res_x, res_y, num_frames, ring_frames, int(pinned), int(create)) != 0:
            raise capi.CkdError(self.L.ckdhost_last_error().decode())

    def acquire(self):
        """-> address of the next free frame buffer (blocks while the ring is full)"""
        ptr = self.L.ckdsink_acquire()
        if not ptr:
            raise capi.CkdError(self.L.ckdhost_last_error().decode())
        return ptr

    def view(self, ptr):
        """numpy view (res_y, res_x) uint32 of an acquired buffer"""
        buf = (C.c_uint32 * (self.res_x * self.res_y)).from_address(ptr)
        return np.frombuffer(buf, dtype=np.uint32).reshape(self.res_y, self.res_x)

    def commit(self, ptr, frame_index):
        if self.L.ckdsink_commit(C.c_void_p(ptr), frame_index) != 0:
            raise capi.CkdError(self.L.ckdhost_last_error().decode())

    def close(self):
        if self.L.ckdsink_close() != 0:
            raise capi.CkdError(self.L.ckdhost_last_error().decode())


def read_header(path):
    with open(path, "rb") as f:
        magic, version, res_x, res_y, num_frames = struct.unpack("<5I", f.read(20))
    if magic != MAGIC or version != 1:
        raise ValueError(f"{path}: not a CKDF stream")
    return res_x, res_y, num_frames


def read_frame(path, index):
    res_x, res_y, num_frames = read_header(path)
    if not (0 <= index < num_frames):
        raise IndexError(index)
    return np.fromfile(path, dtype="<u4", count=res_x * res_y, offset=HEADER_BYTES + index * res_x * res_y * 4).reshape(res_y, res_x)
