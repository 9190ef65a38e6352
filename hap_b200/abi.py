"""ctypes mirror of the Hap C API (`include/hap.h`, same ABI as /root/reference/source/hap.h:40-152).

`HapABI` binds the six `hap.h` entry points of ANY shared library that exports them (optionally
behind a symbol prefix), so the very same Python calls drive libhap_b200.so (the product), the
unmodified reference build and the CPU oracle in the parity tests.  Nothing here computes anything:
it marshals pointers and sizes.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Sequence

# hap.h:40-48
HapTextureFormat_RGB_DXT1 = 0x83F0
HapTextureFormat_RGBA_DXT5 = 0x83F3
HapTextureFormat_YCoCg_DXT5 = 0x01
HapTextureFormat_A_RGTC1 = 0x8DBB
HapTextureFormat_RGBA_BPTC_UNORM = 0x8E8C
HapTextureFormat_RGB_BPTC_UNSIGNED_FLOAT = 0x8E8F
HapTextureFormat_RGB_BPTC_SIGNED_FLOAT = 0x8E8E
# hap.h:50-53
HapCompressorNone = 0
HapCompressorSnappy = 1
# hap.h:55-61
HapResult_No_Error = 0
HapResult_Bad_Arguments = 1
HapResult_Buffer_Too_Small = 2
HapResult_Bad_Frame = 3
HapResult_Internal_Error = 4

WORK_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint)                       # hap.h:66
DECODE_CB = C.CFUNCTYPE(None, WORK_FN, C.c_void_p, C.c_uint, C.c_void_p)  # hap.h:67


def _as_ptr(buf) -> tuple[int, int]:
    """(address, nbytes) of bytes / bytearray / numpy array / (addr, n) tuple."""
    if isinstance(buf, tuple):
        return int(buf[0]), int(buf[1])
    if isinstance(buf, (bytes, bytearray)):
        n = len(buf)
        if isinstance(buf, bytes):
            return C.cast(C.c_char_p(buf), C.c_void_p).value or 0, n
        return C.addressof((C.c_char * n).from_buffer(buf)) if n else 0, n
    # numpy / torch-cpu style
    if hasattr(buf, "ctypes"):
        return int(buf.ctypes.data), int(buf.nbytes)
    if hasattr(buf, "data_ptr"):
        return int(buf.data_ptr()), int(buf.numel() * buf.element_size())
    raise TypeError(f"unsupported buffer type {type(buf)!r}")


def serial_callback(function, p, count, info):
    """The example callback of hap.h:118-126: run every chunk on the calling thread."""
    for i in range(count):
        function(p, i)


class HapABI:
    def __init__(self, path: str, prefix: str = ""):
        self.path = path
        self.prefix = prefix
        self.lib = C.CDLL(path, mode=getattr(C, "RTLD_LOCAL", 0))
        L = self.lib

        def sym(name):
            return getattr(L, prefix + name)

        self._max = sym("HapMaxEncodedLength")
        self._max.restype = C.c_ulong
        self._max.argtypes = [C.c_uint, C.POINTER(C.c_ulong), C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        self._enc = sym("HapEncode")
        self._enc.restype = C.c_uint
        self._enc.argtypes = [C.c_uint, C.POINTER(C.c_void_p), C.POINTER(C.c_ulong), C.POINTER(C.c_uint),
                              C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.c_void_p, C.c_ulong,
                              C.POINTER(C.c_ulong)]
        self._dec = sym("HapDecode")
        self._dec.restype = C.c_uint
        self._dec.argtypes = [C.c_void_p, C.c_ulong, C.c_uint, DECODE_CB, C.c_void_p, C.c_void_p, C.c_ulong,
                              C.POINTER(C.c_ulong), C.POINTER(C.c_uint)]
        self._cnt = sym("HapGetFrameTextureCount")
        self._cnt.restype = C.c_uint
        self._cnt.argtypes = [C.c_void_p, C.c_ulong, C.POINTER(C.c_uint)]
        self._fmt = sym("HapGetFrameTextureFormat")
        self._fmt.restype = C.c_uint
        self._fmt.argtypes = [C.c_void_p, C.c_ulong, C.c_uint, C.POINTER(C.c_uint)]
        self._chk = sym("HapGetFrameTextureChunkCount")
        self._chk.restype = C.c_uint
        self._chk.argtypes = [C.c_void_p, C.c_ulong, C.c_uint, C.POINTER(C.c_int)]

    # -- hap.h:76-79
    def max_encoded_length(self, lengths: Sequence[int], formats: Sequence[int], chunks: Sequence[int],
                           count: Optional[int] = None) -> int:
        n = len(lengths) if count is None else count
        m = max(len(lengths), 1)
        return int(self._max(n, (C.c_ulong * m)(*lengths), (C.c_uint * m)(*formats), (C.c_uint * m)(*chunks)))

    # -- hap.h:98-104; returns (result, frame bytes or None)
    def encode(self, textures: Sequence, formats: Sequence[int], compressors: Sequence[int],
               chunks: Sequence[int], out_capacity: Optional[int] = None, out=None):
        n = len(textures)
        ptrs, lens, keep = [], [], []
        for t in textures:
            a, nb = _as_ptr(t)
            keep.append(t)
            ptrs.append(a)
            lens.append(nb)
        if out_capacity is None:
            out_capacity = self.max_encoded_length(lens, formats, chunks)
        if out is None:
            out = bytearray(max(out_capacity, 1))
        oaddr, _ = _as_ptr(out)
        used = C.c_ulong(0)
        r = self._enc(n, (C.c_void_p * n)(*ptrs), (C.c_ulong * n)(*lens), (C.c_uint * n)(*formats),
                      (C.c_uint * n)(*compressors), (C.c_uint * n)(*chunks), oaddr, out_capacity, C.byref(used))
        if r != HapResult_No_Error:
            return int(r), None
        return 0, bytes(memoryview(out)[: used.value]) if isinstance(out, bytearray) else used.value

    # -- hap.h:132-137; returns (result, texture bytes or None, format, callback invocations [(count)])
    def decode(self, frame, index: int = 0, out_capacity: Optional[int] = None,
               callback: Optional[Callable] = serial_callback, out=None, frame_bytes: Optional[int] = None):
        faddr, fn = _as_ptr(frame)
        if frame_bytes is not None:
            fn = frame_bytes
        if out is None:
            out = bytearray(max(out_capacity if out_capacity is not None else 1, 1))
        oaddr, on = _as_ptr(out)
        if out_capacity is None:
            out_capacity = on
        calls = []

        def _cb(function, p, count, info):
            calls.append(int(count))
            callback(function, p, count, info)

        cb = DECODE_CB(_cb) if callback is not None else C.cast(None, DECODE_CB)
        used = C.c_ulong(0)
        fmt = C.c_uint(0)
        r = self._dec(faddr, fn, index, cb, None, oaddr, out_capacity, C.byref(used), C.byref(fmt))
        if r != HapResult_No_Error:
            return int(r), None, int(fmt.value), calls
        data = bytes(memoryview(out)[: used.value]) if isinstance(out, bytearray) else used.value
        return 0, data, int(fmt.value), calls

    # -- hap.h:142-152
    def texture_count(self, frame):
        a, n = _as_ptr(frame)
        c = C.c_uint(0)
        r = self._cnt(a, n, C.byref(c))
        return int(r), int(c.value)

    def texture_format(self, frame, index: int):
        a, n = _as_ptr(frame)
        f = C.c_uint(0)
        r = self._fmt(a, n, index, C.byref(f))
        return int(r), int(f.value)

    def chunk_count(self, frame, index: int):
        a, n = _as_ptr(frame)
        c = C.c_int(0)
        r = self._chk(a, n, index, C.byref(c))
        return int(r), int(c.value)
