"""TEST INFRASTRUCTURE ONLY: binds rootfinding_b200's host driver to the kernel-source emulator.

`emulated_engine()` builds tests/emul/libyr_emul.so (g++, pthreads) from the very headers the
CUDA kernels are compiled from and returns a SubdivisionEngine whose "device memory" is numpy.
It lets the CPU-only test tier exercise the real host code (engine, postpass, API mirrors) and
the real kernel logic.  Nothing in the product package imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from rootfinding_b200 import _lib
from rootfinding_b200.engine import SubdivisionEngine, _Buf

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libyr_emul.so")
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "rootfinding_b200", "csrc")


def build():
    srcs = [os.path.join(HERE, "emul.cpp")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
    srcs.append(os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "yroots_b200.h"))
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-shared", "-pthread", "-Wno-subobject-linkage",
                               "-o", SO, os.path.join(HERE, "emul.cpp")])
    return SO


class NumpyMemory:
    def __init__(self):
        self.h2d_bytes = self.d2h_bytes = 0

    def empty(self, nbytes):
        a = np.empty(max(1, (int(nbytes) + 7) // 8), dtype=np.int64)
        return _Buf(a.ctypes.data, a.nbytes, a)

    def zeros(self, nbytes):
        b = self.empty(nbytes)
        b.owner[:] = 0
        return b

    def zero_(self, buf):
        buf.owner[:] = 0

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        b = self.empty(arr.nbytes)
        b.owner.view(np.uint8)[:arr.nbytes] = arr.view(np.uint8).reshape(-1)
        return b

    def download(self, buf, nbytes, offset=0):
        return buf.owner.view(np.uint8)[offset:offset + nbytes].copy()

    def copy(self, dst, src, nbytes):
        dst.owner.view(np.uint8)[:nbytes] = src.owner.view(np.uint8)[:nbytes]

    def stream(self):
        return None

    def synchronize(self):
        pass


_engine = None


def emulated_engine(threads=8, lanes=4):
    global _engine
    os.environ["YR_EMUL_THREADS"] = str(threads)
    os.environ["YR_EMUL_LANES"] = str(lanes)
    if _engine is None:
        lib = _lib.bind(C.CDLL(build()))
        _engine = SubdivisionEngine(lib=lib, mem=NumpyMemory())
    return _engine
