"""Frame sink, host side only: the writer thread places frames by index, out of order and from
two attached writers, and the reader gets them back bit for bit.  No GPU: the ring uses plain host memory here."""
import os
import subprocess
import sys

import numpy as np

from conftest import REPO


def _frame(i, res_x, res_y):
    return ((np.arange(res_x * res_y, dtype=np.uint32) * np.uint32(2654435761)) ^ np.uint32(i * 0x01010101)).reshape(res_y, res_x)


def test_sink_round_trip_out_of_order(tmp_path):
    from cookiedough_b200 import sink
    path = tmp_path / "frames.ckdf"
    res_x, res_y, n = 320, 180, 12
    out = sink.Sink(path, res_x, res_y, n, ring_frames=3, pinned=False, create=True)
    order = [5, 0, 11, 3, 1, 2, 4, 10, 9, 8, 7, 6]
    for i in order:
        ptr = out.acquire()
        out.view(ptr)[:] = _frame(i, res_x, res_y)
        out.commit(ptr, i)
    out.close()
    assert sink.read_header(path) == (res_x, res_y, n)
    assert os.path.getsize(path) == sink.HEADER_BYTES + n * res_x * res_y * 4
    for i in range(n):
        assert np.array_equal(sink.read_frame(path, i), _frame(i, res_x, res_y)), i


def test_sink_two_writers_share_a_file(tmp_path):
    """one process creates the stream, a second one attaches and writes the odd frames (frame i -> rank i mod 2)"""
    from cookiedough_b200 import sink
    path = tmp_path / "shared.ckdf"
    res_x, res_y, n = 256, 64, 8
    out = sink.Sink(path, res_x, res_y, n, ring_frames=2, pinned=False, create=True)
    child = ("import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
             "from cookiedough_b200 import sink\n"
             "from test_sink import _frame\n"
             "out = sink.Sink(%r, %d, %d, %d, ring_frames=2, pinned=False, create=False)\n"
             "for i in range(1, %d, 2):\n"
             "    p = out.acquire(); out.view(p)[:] = _frame(i, %d, %d); out.commit(p, i)\n"
             "out.close()\n") % (REPO, os.path.join(REPO, "tests"), str(path), res_x, res_y, n, n, res_x, res_y)
    proc = subprocess.Popen([sys.executable, "-c", child])
    for i in range(0, n, 2):
        ptr = out.acquire()
        out.view(ptr)[:] = _frame(i, res_x, res_y)
        out.commit(ptr, i)
    out.close()
    assert proc.wait(timeout=120) == 0
    for i in range(n):
        assert np.array_equal(sink.read_frame(path, i), _frame(i, res_x, res_y)), i


def test_sink_rejects_bad_use(tmp_path):
    import pytest
    from cookiedough_b200 import capi, sink
    with pytest.raises(capi.CkdError):
        sink.Sink(tmp_path / "missing_dir" / "x.ckdf", 16, 16, 1, pinned=False)
    out = sink.Sink(tmp_path / "ok.ckdf", 16, 16, 2, ring_frames=1, pinned=False)
    ptr = out.acquire()
    with pytest.raises(capi.CkdError):
        out.commit(ptr, 2)  # frame index out of range
    # the rejected buffer went back to the one-frame ring: acquiring again does not block
    ptr = out.acquire()
    out.view(ptr)[:] = _frame(1, 16, 16)
    out.commit(ptr, 1)
    out.close()
    assert np.array_equal(sink.read_frame(tmp_path / "ok.ckdf", 1), _frame(1, 16, 16))
    # attaching checks the header of the existing stream: resolution and frame count must match
    with pytest.raises(capi.CkdError):
        sink.Sink(tmp_path / "ok.ckdf", 32, 16, 2, ring_frames=1, pinned=False, create=False)
    with pytest.raises(capi.CkdError):
        sink.Sink(tmp_path / "ok.ckdf", 16, 16, 3, ring_frames=1, pinned=False, create=False)
    again = sink.Sink(tmp_path / "ok.ckdf", 16, 16, 2, ring_frames=1, pinned=False, create=False)
    again.close()
