"""Host build of the resumable decoder step (csrc/inflate_resume_core.h) for CPU tests.

The core is plain C++ shared by the kernel (csrc/inflate_resume.cuh) and this harness: g++ compiles
the very same header, so the CPU tests exercise the code the GPU runs — one decoder state at a time
instead of one per lane.  `step` has the signature libdeflate_rsx_b200.stream.DeflateDecoder takes as
its backend.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CORE = os.path.join(ROOT, "libdeflate_rsx_b200", "csrc", "inflate_resume_core.h")
BUILD = os.path.join(HERE, "_build")

WRAPPER = r"""
#include <cstring>
#include "%s"
extern "C" int resume_host(bdf_inflate_state *S, const uint8_t *in, uint64_t in_len, int in_final, uint8_t *win,
                           uint64_t cap, uint64_t *pos, uint64_t *consumed)
{
    static bdf_rs::Tables T;
    memset(&T, 0xA5, sizeof(T));        // nothing may survive in the tables between two steps
    return bdf_rs::resume_step(*S, in, in_len, in_final != 0, win, cap, pos, consumed, T);
}
extern "C" unsigned resume_state_size(void) { return (unsigned)sizeof(bdf_inflate_state); }
"""

_lib = None


def build():
    global _lib
    if _lib is not None:
        return _lib
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(BUILD, "resume_host.cpp")
    so = os.path.join(BUILD, "libresume_host.so")
    text = WRAPPER % CORE
    stale = (not os.path.exists(so) or not os.path.exists(src) or open(src).read() != text or
             os.path.getmtime(CORE) > os.path.getmtime(so) or
             os.path.getmtime(os.path.join(ROOT, "include", "bdeflate.h")) > os.path.getmtime(so))
    if stale:
        with open(src, "w") as f:
            f.write(text)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-Wall", "-shared", "-fPIC", "-o", so, src])
    _lib = ctypes.CDLL(so)
    _lib.resume_host.restype = ctypes.c_int
    _lib.resume_host.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p,
                                 ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
    _lib.resume_state_size.restype = ctypes.c_uint
    return _lib


def step(state, data, in_final, window, write_pos):
    """state: writable buffer of 368 bytes; data: bytes; window: np.uint8 array -> (status, consumed, new write_pos)"""
    lib = build()
    st = (ctypes.c_char * len(state)).from_buffer(state)
    pos = ctypes.c_uint64(int(write_pos))
    used = ctypes.c_uint64(0)
    assert isinstance(window, np.ndarray) and window.dtype == np.uint8 and window.flags.c_contiguous
    rc = lib.resume_host(ctypes.addressof(st), bytes(data), len(data), int(bool(in_final)),
                         window.ctypes.data, window.size, ctypes.byref(pos), ctypes.byref(used))
    return rc, int(used.value), int(pos.value)
