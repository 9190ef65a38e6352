"""Runs the CUDA kernels' logic under the fiber SIMT emulator (tests/emu/simt_emu.h) against the
oracle and the unmodified reference.  This is how kernel logic is debugged in a container without a
GPU; the -m gpu tests repeat the same checks on the real device."""
import os
import subprocess

import pytest

import oracles

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "emu", "_build")
INC = ["-I", os.path.join(ROOT, "hap_b200", "csrc"), "-I", os.path.join(ROOT, "tests", "emu"), "-I", os.path.join(ROOT, "oracle")]
ORC = [os.path.join(ROOT, "oracle", f) for f in ("snappy_oracle.c", "hap_oracle.c", "bc_oracle.c")]


def compile_emu(name, extra):
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, name)
    src = os.path.join(ROOT, "tests", "emu", name + ".cc")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-w"] + INC + [src] + extra + ["-ldl", "-lm", "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


def test_snappy_decode_kernel_emulated():
    exe = compile_emu("test_decode_emu", ORC[:1])
    p = subprocess.run([exe, "3"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr


def test_encode_path_emulated_and_decoded_by_reference():
    exe = compile_emu("test_encode_emu", ORC)
    oracles.ref_abi()  # builds oracle/_ref when the reference sources are present
    ref = oracles.REF_SO if os.path.exists(oracles.REF_SO) else ""
    p = subprocess.run([exe, "3"] + ([ref] if ref else []), capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    if ref:
        assert "reference decoder: loaded" in p.stdout


def test_kernels_under_address_sanitizer():
    """The same kernel sources under ASan with exact-size buffers: no read or write outside the chunk / frame."""
    for name, extra in (("test_decode_emu", ORC[:1]), ("test_encode_emu", ORC)):
        exe = os.path.join(BUILD, name + "_asan")
        src = os.path.join(ROOT, "tests", "emu", name + ".cc")
        subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-w", "-fsanitize=address"] + INC + [src] + extra + ["-ldl", "-lm", "-o", exe],
                       check=True)
        env = dict(os.environ, ASAN_OPTIONS="detect_stack_use_after_return=0:detect_leaks=0")
        p = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=900, env=env)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
