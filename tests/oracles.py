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
