"""Test-only bindings to the CPU oracle (oracle/liboracle.so) and the unmodified reference build
(oracle/_ref/libhap_ref.so).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
may import this module; the product package never does."""
from __future__ import annotations

import ctypes as C
import functools
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libhap_ref.so")


def _ensure_built():
    if not os.path.exists(ORACLE_SO) or (os.path.isdir("/root/reference/source") and not os.path.exists(REF_SO)):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all"], check=True, capture_output=True)


@functools.lru_cache(None)
def oracle_abi():
    from hap_b200.abi import HapABI
    _ensure_built()
    return HapABI(ORACLE_SO, prefix="orc_")


@functools.lru_cache(None)
def ref_abi():
    """None when the reference build is unavailable."""
    from hap_b200.abi import HapABI
    _ensure_built()
    if not os.path.exists(REF_SO):
        return None
    try:
        return HapABI(REF_SO)
    except OSError:
        return None


@functools.lru_cache(None)
def _olib():
    _ensure_built()
    L = C.CDLL(ORACLE_SO)
    L.orc_snappy_max_compressed_length.restype = C.c_size_t
    L.orc_snappy_max_compressed_length.argtypes = [C.c_size_t]
    for name in ("orc_snappy_compress", "orc_snappy_uncompress"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)]
    L.orc_snappy_uncompressed_length.restype = C.c_int
    L.orc_snappy_uncompressed_length.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.orc_snappy_scan.restype = C.c_int
    L.orc_snappy_scan.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    L.orc_hap_limited_chunk_count.restype = C.c_uint
    L.orc_hap_limited_chunk_count.argtypes = [C.c_ulong, C.c_uint, C.c_uint]
    L.orc_mse_rgba.restype = C.c_double
    L.orc_mse_rgba.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint]
    return L


def _np(buf):
    return np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else np.ascontiguousarray(buf)


# ---- Snappy ----------------------------------------------------------------------------------------

def snappy_compress(data) -> bytes:
    a = _np(data)
    L = _olib()
    cap = L.orc_snappy_max_compressed_length(a.size)
    out = np.empty(cap, np.uint8)
    n = C.c_size_t(cap)
    r = L.orc_snappy_compress(a.ctypes.data, a.size, out.ctypes.data, C.byref(n))
    assert r == 0, r
    return out[: n.value].tobytes()


def snappy_uncompress(data, cap=None):
    """(status, bytes|None)"""
    a = _np(data)
    L = _olib()
    want = C.c_size_t(0)
    if L.orc_snappy_uncompressed_length(a.ctypes.data, a.size, C.byref(want)) != 0:
        return 1, None
    if cap is None:
        cap = want.value
    out = np.empty(max(cap, 1), np.uint8)
    n = C.c_size_t(cap)
    r = L.orc_snappy_uncompress(a.ctypes.data, a.size, out.ctypes.data, C.byref(n))
    if r != 0:
        return r, None
    return 0, out[: n.value].tobytes()


class SnappyStats(C.Structure):
    _fields_ = [("literals", C.c_uint64), ("literal_bytes", C.c_uint64), ("copy1", C.c_uint64),
                ("copy2", C.c_uint64), ("copy4", C.c_uint64), ("copy_bytes", C.c_uint64),
                ("overlapping", C.c_uint64), ("max_offset", C.c_uint32)]


def snappy_scan(data):
    a = _np(data)
    st = SnappyStats()
    r = _olib().orc_snappy_scan(a.ctypes.data, a.size, C.byref(st))
    return r, st


def limited_chunk_count(nbytes, fmt, k):
    return int(_olib().orc_hap_limited_chunk_count(nbytes, fmt, k))


# ---- block codecs ----------------------------------------------------------------------------------

def bc_decode(kind: str, blocks, w: int, h: int) -> np.ndarray:
    """kind in bc1|bc3|bc4|ycocg -> (h,w,4) uint8 (bc4: (h,w))"""
    L = _olib()
    a = _np(blocks)
    if kind == "bc4":
        out = np.empty((h, w), np.uint8)
        L.orc_bc4_decode(C.c_void_p(a.ctypes.data), w, h, C.c_void_p(out.ctypes.data))
        return out
    out = np.empty((h, w, 4), np.uint8)
    fn = {"bc1": L.orc_bc1_decode, "bc3": L.orc_bc3_decode, "ycocg": L.orc_ycocg_dxt5_decode}[kind]
    fn(C.c_void_p(a.ctypes.data), w, h, C.c_void_p(out.ctypes.data))
    return out


def bc_encode_clusterfit(kind: str, rgba: np.ndarray, iterations: int = 8, channel: int = 3) -> bytes:
    """kind in bc1|bc3|ycocg|bc4 ; rgba (h,w,4) uint8"""
    L = _olib()
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = rgba.shape[:2]
    nblk = (w // 4) * (h // 4)
    if kind == "bc4":
        out = np.empty(nblk * 8, np.uint8)
        L.orc_bc4_encode_squish(C.c_void_p(rgba.ctypes.data), w, h, channel, C.c_void_p(out.ctypes.data))
        return out.tobytes()
    out = np.empty(nblk * (8 if kind == "bc1" else 16), np.uint8)
    fn = {"bc1": L.orc_bc1_encode_clusterfit, "bc3": L.orc_bc3_encode_clusterfit,
          "ycocg": L.orc_ycocg_dxt5_encode_clusterfit}[kind]
    fn(C.c_void_p(rgba.ctypes.data), w, h, C.c_void_p(out.ctypes.data), iterations)
    return out.tobytes()


def ycocg_scaled_texels(rgba: np.ndarray) -> np.ndarray:
    L = _olib()
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = rgba.shape[:2]
    out = np.empty_like(rgba)
    L.orc_ycocg_scaled_texels(C.c_void_p(rgba.ctypes.data), w, h, C.c_void_p(out.ctypes.data))
    return out


def psnr(a: np.ndarray, b: np.ndarray, channels=(0, 1, 2)) -> float:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    d = a[..., list(channels)].astype(np.float64) - b[..., list(channels)].astype(np.float64)
    mse = float((d * d).mean())
    return 99.0 if mse == 0 else 10.0 * np.log10(255.0 * 255.0 / mse)


# ---- threaded drivers of the CPU baseline (oracle/mt_driver.c) ---------------------------------------

class MtDriver:
    """The reference's own parallel mechanisms, driven from C (oracle/mt_driver.c): frame-parallel HapEncode
    (hap.c:448-476 is serial inside a frame and has no callback), HapDecode with a pthread-pool HapDecodeCallback
    (hap.c:861) and frame-parallel HapDecode.  kind = "reference" (oracle/_ref, unmodified hap.c + Google Snappy) or
    "port" (oracle/liboracle.so).  Buffers are numpy uint8 arrays; nothing is copied inside the timed calls."""

    def __init__(self):
        _ensure_built()
        if os.path.exists(REF_SO):
            try:
                self.lib, self.prefix, self.kind = C.CDLL(REF_SO), "refdrv_", "reference"
            except OSError:
                self.lib, self.prefix, self.kind = C.CDLL(ORACLE_SO), "orcdrv_", "port"
        else:
            self.lib, self.prefix, self.kind = C.CDLL(ORACLE_SO), "orcdrv_", "port"
        vp, u, ul = C.c_void_p, C.c_uint, C.c_ulong
        self._enc = getattr(self.lib, self.prefix + "encode_frames_mt")
        self._enc.restype = u
        self._enc.argtypes = [u, u, C.POINTER(vp), C.POINTER(ul), C.POINTER(u), C.POINTER(u), C.POINTER(u), vp, ul, C.POINTER(ul), u]
        self._decf = getattr(self.lib, self.prefix + "decode_frames_mt")
        self._decf.restype = u
        self._decf.argtypes = [u, vp, ul, C.POINTER(ul), vp, ul, u]
        self._dec1 = getattr(self.lib, self.prefix + "decode_mt")
        self._dec1.restype = u
        self._dec1.argtypes = [vp, ul, u, vp, ul, C.POINTER(ul), C.POINTER(u), u]

    def encode_frames(self, textures, tex_bytes, fmt, compressor, chunks, out, out_stride, threads):
        """textures: list (one per frame) of numpy arrays; out: numpy uint8 [frames*out_stride].  Returns (result, used[])"""
        n = len(textures)
        ins = (C.c_void_p * n)(*[t.ctypes.data for t in textures])
        used = (C.c_ulong * n)()
        r = self._enc(n, 1, ins, (C.c_ulong * 1)(tex_bytes), (C.c_uint * 1)(fmt), (C.c_uint * 1)(compressor), (C.c_uint * 1)(chunks),
                      out.ctypes.data, out_stride, used, threads)
        return int(r), list(used)

    def decode_frames(self, frames_buf, in_stride, in_bytes, out, out_stride, threads):
        n = len(in_bytes)
        return int(self._decf(n, frames_buf.ctypes.data, in_stride, (C.c_ulong * n)(*in_bytes), out.ctypes.data, out_stride, threads))

    def decode_one_chunk_parallel(self, frame_addr, nbytes, out_addr, cap, threads):
        used, fmt = C.c_ulong(0), C.c_uint(0)
        r = self._dec1(frame_addr, nbytes, 0, out_addr, cap, C.byref(used), C.byref(fmt), threads)
        return int(r), int(used.value)
