"""ctypes mirror of include/hap_mov.h (QuickTime sample tables around Hap frames; host code of libhap_b200.so)."""
from __future__ import annotations

import ctypes as C

from .lib import load


def fourcc(s: str) -> int:
    b = s.encode("ascii")
    assert len(b) == 4
    return (b[0] << 24) | (b[1] << 16) | (b[2] << 8) | b[3]


def fourcc_str(v: int) -> str:
    return bytes([(v >> 24) & 255, (v >> 16) & 255, (v >> 8) & 255, v & 255]).decode("ascii", "replace")


def _lib():
    L = load().lib
    if getattr(L, "_mov_ready", False):
        return L
    vp, u, ul = C.c_void_p, C.c_uint, C.c_ulong
    L.HapB200MovFourCCForFrame.restype = u
    L.HapB200MovFourCCForFrame.argtypes = [vp, ul, C.POINTER(u)]
    L.HapB200MovOpen.restype = vp
    L.HapB200MovOpen.argtypes = [C.c_char_p]
    L.HapB200MovInfo.restype = u
    L.HapB200MovInfo.argtypes = [vp, C.POINTER(u), C.POINTER(u), C.POINTER(u), C.POINTER(ul), C.POINTER(u), C.POINTER(ul)]
    L.HapB200MovFrameBytes.restype = ul
    L.HapB200MovFrameBytes.argtypes = [vp, ul]
    L.HapB200MovReadFrame.restype = u
    L.HapB200MovReadFrame.argtypes = [vp, ul, vp, ul, C.POINTER(ul), C.POINTER(u)]
    L.HapB200MovCreate.restype = vp
    L.HapB200MovCreate.argtypes = [C.c_char_p, u, u, u, u]
    L.HapB200MovWriteFrame.restype = u
    L.HapB200MovWriteFrame.argtypes = [vp, vp, ul, u]
    L.HapB200MovClose.restype = u
    L.HapB200MovClose.argtypes = [vp]
    L._mov_ready = True
    return L


def fourcc_for_frame(frame: bytes):
    """(result, 'HapY' | None)"""
    v = C.c_uint(0)
    buf = (C.c_char * len(frame)).from_buffer_copy(frame)
    r = _lib().HapB200MovFourCCForFrame(C.addressof(buf), len(frame), C.byref(v))
    return int(r), (fourcc_str(v.value) if r == 0 else None)


class MovReader:
    def __init__(self, path: str):
        self._h = _lib().HapB200MovOpen(path.encode())
        if not self._h:
            raise ValueError(f"{path}: not a QuickTime movie with a Hap video track")
        f, w, h, ts = C.c_uint(0), C.c_uint(0), C.c_uint(0), C.c_uint(0)
        n, d = C.c_ulong(0), C.c_ulong(0)
        _lib().HapB200MovInfo(self._h, C.byref(f), C.byref(w), C.byref(h), C.byref(n), C.byref(ts), C.byref(d))
        self.fourcc, self.width, self.height = fourcc_str(f.value), w.value, h.value
        self.frames, self.timescale, self.duration = n.value, ts.value, d.value

    def frame_bytes(self, i: int) -> int:
        return int(_lib().HapB200MovFrameBytes(self._h, i))

    def read(self, i: int, capacity: int | None = None):
        """(result, frame bytes | None, duration ticks)"""
        cap = self.frame_bytes(i) if capacity is None else capacity
        buf = (C.c_char * max(cap, 1))()
        used, dt = C.c_ulong(0), C.c_uint(0)
        r = _lib().HapB200MovReadFrame(self._h, i, C.addressof(buf), cap, C.byref(used), C.byref(dt))
        return int(r), (bytes(buf[: used.value]) if r == 0 else None), dt.value

    def close(self):
        if self._h:
            _lib().HapB200MovClose(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class MovWriter:
    def __init__(self, path: str, codec: str, width: int, height: int, timescale: int):
        self._h = _lib().HapB200MovCreate(path.encode(), fourcc(codec), width, height, timescale)
        if not self._h:
            raise ValueError("HapB200MovCreate refused the arguments or could not create the file")

    def write(self, frame: bytes, ticks: int) -> int:
        buf = (C.c_char * len(frame)).from_buffer_copy(frame) if frame else (C.c_char * 1)()
        return int(_lib().HapB200MovWriteFrame(self._h, C.addressof(buf), len(frame), ticks))

    def close(self) -> int:
        r = 0
        if self._h:
            r = int(_lib().HapB200MovClose(self._h))
            self._h = None
        return r

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
