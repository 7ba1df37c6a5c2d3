"""Host emulation of the CUDA library: TEST INFRASTRUCTURE ONLY.

``build()`` passes the product's own sources (safeincave_b200/csrc/*.cu, *.cuh, untouched) through
``translate.py`` (kernel-launch syntax and four inline-PTX loads -> plain C++), compiles them with g++
against ``cuda_emu.h`` (a single-threaded block/thread scheduler with fibers for __syncthreads and warp
shuffles) and links ``_build/libsic_hostemu.so``, which exports the single-GPU part of the C ABI of
include/safeincave_cuda.h operating on HOST pointers.

Why: the build container has no GPU.  With this, the CPU test-suite executes the very code the B200 runs
-- per-cell constitutive kernels, the matrix-free operator with its shared-memory scatter plan, the Krylov
and multigrid drivers -- and checks it against the oracle; `-m gpu` then only has to confirm that nvcc's
build of the same source behaves the same on the device.

Never imported by the product: ``safeincave_b200.engine.Engine`` loads libsafeincave_cuda.so and raises
without a CUDA device; ``EmuEngine`` below lives here, under tests/.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

from safeincave_b200 import _lib as L
from safeincave_b200.engine import Engine

from . import translate as _tr

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "safeincave_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libsic_hostemu.so")
# translation units of the single-GPU path (comm.cu = NCCL / CUDA IPC: not emulated)
UNITS = [("common.cu", []), ("constitutive.cu", ["-ffp-contract=off"]), ("fem.cu", []), ("solver.cu", []),
         ("mg.cu", []), ("heat.cu", []), ("fields.cu", [])]
CXXFLAGS = ["-O2", "-std=c++17", "-fPIC", "-mfma", "-w", "-x", "c++"]


def _stale(target, sources):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False):
    src_out = os.path.join(OUT, "src")
    os.makedirs(src_out, exist_ok=True)
    names = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")) and f != "ebe_tma.cuh")
    inputs = [os.path.join(CSRC, f) for f in names]
    support = [os.path.join(HERE, f) for f in ("cuda_emu.h", "cuda_runtime.h", "emu_runtime.cpp", "translate.py")]
    support.append(os.path.join(ROOT, "include", "safeincave_cuda.h"))
    if not force and not _stale(LIB, inputs + support):
        return LIB
    for f, p in zip(names, inputs):
        with open(p) as fh:
            text = fh.read()
        with open(os.path.join(src_out, f), "w") as fh:
            text = text.replace('"../../include/safeincave_cuda.h"', '"safeincave_cuda.h"')
            fh.write(_tr.translate(text) if f.endswith((".cu", ".cuh")) else text)
    objs = []
    inc = ["-I", HERE, "-I", os.path.join(ROOT, "include")]
    units = [(u, fl) for u, fl in UNITS if os.path.isfile(os.path.join(CSRC, u))]
    procs = []
    for u, extra in units:
        o = os.path.join(OUT, u.replace(".cu", ".o"))
        objs.append(o)
        cmd = ["g++"] + CXXFLAGS + extra + inc + ["-I", src_out, "-include", "cuda_emu.h", "-c",
                                                    os.path.join(src_out, u), "-o", o]
        procs.append((u, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    o = os.path.join(OUT, "emu_runtime.o")
    objs.append(o)
    procs.append(("emu_runtime.cpp", subprocess.Popen(
        ["g++", "-O2", "-std=c++17", "-fPIC", "-w"] + inc + ["-c", os.path.join(HERE, "emu_runtime.cpp"), "-o", o],
        stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for name, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"hostemu: compiling {name} failed:\n{out.decode()[-6000:]}")
    subprocess.check_call(["g++", "-shared", "-o", LIB] + objs + ["-lm", "-lpthread"])
    return LIB


_emu = None


def load():
    """ctypes handle of the emulation library with the product's prototypes (safeincave_b200._lib.declare)."""
    global _emu
    if _emu is None:
        lib = ctypes.CDLL(build())
        L.declare(lib, single_gpu_only=True)
        PH = ctypes.POINTER(L.SicHalo)
        lib.sic_exchange.argtypes = [PH, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        lib.sic_halo_sum.argtypes = [PH, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        lib.sic_allreduce_sum.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        lib.sic_emu_make_comm.argtypes = [ctypes.c_int, ctypes.c_int]
        lib.sic_emu_make_comm.restype = ctypes.c_void_p
        lib.sic_emu_set_timeout.argtypes = [ctypes.c_int]
        _emu = lib
    return _emu


class EmuEngine(Engine):
    """safeincave_b200.engine.Engine over host memory and the emulation library (tests only)."""

    def __init__(self, coords, cells, device="cpu", **kw):
        self.lib = load()
        self._setup(coords, cells, "cpu", **kw)

    def _stream(self):
        return ctypes.c_void_p(0)

    def _tic(self, name):
        return None
