"""Host build of the block-encoder source (hap_b200/csrc/bc_block.cuh via tests/emu/bc_twin.cc)."""
import ctypes as C
import functools
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KIND = {"bc1": 0, "bc3": 1, "ycocg": 2, "bc4": 3, "ycocg_refine": 4}


@functools.lru_cache(None)
def _lib():
    out = os.path.join(ROOT, "tests", "emu", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libbc_twin.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-mfma", "-ffp-contract=off", "-fPIC", "-shared", "-w",
                    "-I", os.path.join(ROOT, "hap_b200", "csrc"), "-I", os.path.join(ROOT, "tests", "emu"),
                    os.path.join(ROOT, "tests", "emu", "bc_twin.cc"), "-o", so], check=True)
    return C.CDLL(so)


def encode(kind: str, rgba: np.ndarray) -> bytes:
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = rgba.shape[:2]
    n = (w // 4) * (h // 4) * (8 if kind in ("bc1", "bc4") else 16)   # "ycocg_refine": HAPB200_OPTION_CHROMA_REFINE on
    out = np.empty(n, np.uint8)
    _lib().twin_encode(C.c_void_p(rgba.ctypes.data), w, h, KIND[kind], C.c_void_p(out.ctypes.data))
    return out.tobytes()


def decode(kind: str, blocks: bytes, w: int, h: int) -> np.ndarray:
    """Host build of the block decoder the kernels compile (bc_decode.cuh) -> (h, w, 4) uint8."""
    b = np.frombuffer(blocks, dtype=np.uint8)
    out = np.empty((h, w, 4), np.uint8)
    _lib().twin_decode(C.c_void_p(b.ctypes.data), w, h, KIND[kind], C.c_void_p(out.ctypes.data))
    return out
