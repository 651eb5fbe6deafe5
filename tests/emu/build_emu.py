"""Builds tests/emu/liboc_emu.so (g++ compile of the kernel's per-lane code, see oc_emu.cpp)."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
SO = os.path.join(HERE, "liboc_emu.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(HERE, "oc_emu.cpp"), os.path.join(ROOT, "include", "ocb.h"),
            os.path.join(ROOT, "diverse_conventions_b200", "csrc", "oc_core.cuh"),
            os.path.join(ROOT, "diverse_conventions_b200", "csrc", "oc_tables.h"),
            os.path.join(ROOT, "diverse_conventions_b200", "csrc", "mixed_schedule.h")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                               "-I" + os.path.join(ROOT, "include"),
                               "-I" + os.path.join(ROOT, "diverse_conventions_b200", "csrc"), srcs[0], "-o", SO])
    return SO


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        vp = ctypes.c_void_p
        L.ocemu_rollout.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_uint64,
                                    ctypes.c_uint64, ctypes.c_uint32, vp, vp, vp, vp]
        L.ocemu_mix_forced.argtypes = [ctypes.c_int] * 3
        L.ocemu_mix_items.argtypes = [ctypes.c_int] * 4 + [vp, ctypes.c_int]
        L.ocemu_mix_draw.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64]
        _lib = L
    return _lib
