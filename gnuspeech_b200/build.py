"""Builds the native libraries in-tree (gnuspeech_b200/lib/):

  libtrm_cuda.so  -- sm_100a kernels + C-ABI shim (nvcc; cross-compiles without a GPU)
  libtrm.so       -- C host library behind include/trm.h (gcc), linked against libtrm_cuda.so

The FP64 strict kernels (kernels_f64s.cu) are compiled with -fmad=false (the reference build has no FMA
contraction); the FP64 conformance and FP32 fast-mode kernels with the default contraction.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gnuspeech_b200", "csrc")
LIB = os.path.join(ROOT, "gnuspeech_b200", "lib")
INC = os.path.join(ROOT, "include")
OBJ = os.path.join(ROOT, "build", "obj")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
GENCODE = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_COMMON = GENCODE + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-I" + INC, "-I" + CSRC]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    if verbose and r.stdout.strip():
        print(r.stdout)


def build(verbose=False, force=False, defines=(), tag=None):
    """Builds the libraries.  `defines` (e.g. ["-DTRM_PROFILE_SKIP=1"]) with a `tag` builds a profiling variant into
    gnuspeech_b200/lib_<tag>/ (selected at import time with TRM_LIB_DIR); the shipped library is built without either."""
    global LIB, OBJ
    if tag:
        LIB = os.path.join(ROOT, "gnuspeech_b200", "lib_" + tag)
        OBJ = os.path.join(ROOT, "build", "obj_" + tag)
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in ("tube_common.cuh", "tube_wide.cuh", "src_kernel.cuh", "framegen_kernel.cuh",
                                               "launch.cuh", "kernel_args.h")]
    headers += [os.path.join(INC, h) for h in ("trm.h", "trm_cuda.h", "trm_workload.h")]
    cu = [("kernels_f64s", ["-fmad=false"]), ("kernels_f64", []), ("kernels_f32", []), ("kernels_aux", ["-fmad=false"]), ("trm_cuda", [])]
    objs = []
    for name, extra in cu:
        src = os.path.join(CSRC, name + ".cu")
        obj = os.path.join(OBJ, name + ".o")
        if force or _newer(obj, [src] + headers):
            _run([NVCC] + NVCC_COMMON + list(defines) + extra + ["-c", src, "-o", obj], verbose)
        objs.append(obj)
    cuda_so = os.path.join(LIB, "libtrm_cuda.so")
    if force or _newer(cuda_so, objs):
        _run([NVCC] + GENCODE + ["-shared", "-o", cuda_so] + objs + ["-cudart", "static"], verbose)
    host_src = [os.path.join(CSRC, f) for f in ("trm_host.c", "trm_workload.c")]
    host_so = os.path.join(LIB, "libtrm.so")
    if force or _newer(host_so, host_src + headers + [cuda_so]):
        _run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=gnu99", "-Wall", "-fPIC", "-shared", "-I" + INC, "-o", host_so] + host_src +
             ["-L" + LIB, "-ltrm_cuda", "-Wl,-rpath,$ORIGIN", "-lm", "-lpthread"], verbose)
    return host_so, cuda_so


def build_oracle(verbose=False):
    """Builds the CPU oracle (test infrastructure) and, when /root/reference is mounted, oracle/_ref."""
    _run(["make", "-C", os.path.join(ROOT, "oracle")], verbose)


if __name__ == "__main__":
    if "--variant" in sys.argv:       # python -m gnuspeech_b200.build --variant <tag> -DX=1 ...
        i = sys.argv.index("--variant")
        build(verbose=True, force=True, defines=[a for a in sys.argv[i + 2:] if a.startswith("-D")], tag=sys.argv[i + 1])
    else:
        build(verbose=True, force="--force" in sys.argv)
        build_oracle(verbose=True)
