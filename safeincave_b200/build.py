"""Build libsafeincave_cuda.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python -m safeincave_b200.build [--force]

constitutive.cu is compiled with -fmad=false: its finite-difference tangents must agree bit for
bit with the CPU oracle (see csrc/sic_math.h).  Everything else uses the default contraction.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsafeincave_cuda.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "--expt-extended-lambda", "-Xcompiler", "-fPIC"]
UNITS = [
    ("common.cu", []),
    ("constitutive.cu", ["-fmad=false"]),
    ("fem.cu", []),
    ("solver.cu", []),
    ("mg.cu", []),
    ("heat.cu", []),
    ("fields.cu", []),
    ("comm.cu", []),
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, sources):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False, defines=(), suffix=""):
    """defines/suffix: build an experimental variant (e.g. -DSIC_EBE_IMPL=2) as libsafeincave_cuda<suffix>.so."""
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build" + suffix)
    lib_path = LIB.replace(".so", suffix + ".so")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "safeincave_cuda.h"))
    objs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + ARCH + COMMON + extra + list(defines) + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    if force or _stale(lib_path, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", lib_path] + objs + ["-lcudart", "-ldl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
