"""In-tree build of the native pieces (called by __graft_entry__.build()):

  lib/libsio_b200.so      CUDA kernels + C-ABI (nvcc, sm_100a only, -lineinfo)
  lib/libsio_hostgeom.so  host-side mesh -> SDF preprocessing (gcc, OpenMP)
  lib/libsio_hostqp.so    batched host QP of the SIP step (gcc, OpenMP)

nvcc cross-compiles without a GPU.  Objects go to build/ (git-ignored); the .so files are
git-ignored too but travel to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
_LIB = os.path.join(_PKG, "lib")
_OBJ = os.path.join(_PKG, "build")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]

# (source, extra flags)
CUDA_UNITS = [
    ("sio_kernels.cu", []),
    ("sio_fp64.cu", ["--fmad=false"]),      # plain IEEE float64 sequence, no contraction
    ("sio_qp.cu", []),                      # float64 with FMA: agrees with the host solver to its tolerance, not bit for bit
    ("sio_lockstep.cu", ["--fmad=false"]),  # the host driver's float64 sequence
    ("sio_sdf.cu", ["--fmad=false"]),       # identical to hostgeom.c's float64 distances
    ("sio_abi.cu", []),
]


def _newer(src_list, out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in src_list)


def _run(cmd, log):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))


def build_cuda(force=False, verbose=False):
    os.makedirs(_LIB, exist_ok=True)
    os.makedirs(_OBJ, exist_ok=True)
    log = []
    hdrs = [os.path.join(_CSRC, "sio_common.cuh"), os.path.join(_CSRC, "sio_eval64.cuh"), os.path.join(_CSRC, "sio_lockstep.cuh"), os.path.join(_PKG, "..", "include", "sio_abi.h")]
    objs = []
    for src, extra in CUDA_UNITS:
        s = os.path.join(_CSRC, src)
        o = os.path.join(_OBJ, src.replace(".cu", ".o"))
        if force or _newer([s] + hdrs, o):
            _run([NVCC] + ARCH + COMMON + extra + ["-c", s, "-o", o], log)
        objs.append(o)
    out = os.path.join(_LIB, "libsio_b200.so")
    if force or _newer(objs, out):
        _run([NVCC] + ARCH + ["-shared", "-o", out] + objs, log)
    if verbose:
        print("\n".join(log))
    open(os.path.join(_OBJ, "build.log"), "a").write("\n".join(log))
    return out


def build_hostgeom(force=False):
    os.makedirs(_LIB, exist_ok=True)
    src = os.path.join(_PKG, "_host", "csrc", "hostgeom.c")
    out = os.path.join(_LIB, "libsio_hostgeom.so")
    if force or _newer([src], out):
        # x86-64-v3 (AVX2): the .so is built here and runs on the GPU box's host CPU.  -ffp-contract=off: the float64
        # distances are then the plain IEEE sequence, which the device builder (sio_sdf.cu, --fmad=false) reproduces
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", src, "-o", out, "-lm"])
    return out


def build_hostqp(force=False):
    os.makedirs(_LIB, exist_ok=True)
    src = os.path.join(_PKG, "_host", "csrc", "hostqp.c")
    out = os.path.join(_LIB, "libsio_hostqp.so")
    if force or _newer([src], out):
        # -ffp-contract=off: the same float64 sequence on every host
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-Wall",
                               src, "-o", out, "-lm"])
    return out


def build_all(force=False, verbose=False):
    return build_cuda(force, verbose), build_hostgeom(force), build_hostqp(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose=True))
