"""Builds the native libraries in-tree (gnuspeech_b200/lib/):

  libtrm_cuda.so  -- sm_100a kernels + C-ABI shim (nvcc; cross-compiles without a GPU)
  libtrm.so       -- C host library behind include/trm.h (gcc), linked against libtrm_cuda.so

The FP64 conformance kernels are compiled with -fmad=false (the reference build has no FMA
contraction); the FP32 fast-mode kernels with the default contraction.
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


def build(verbose=False, force=False):
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in ("tube_kernel.cuh", "tube_wide.cuh", "src_kernel.cuh", "framegen_kernel.cuh",
                                               "launch.cuh", "kernel_args.h")]
    headers += [os.path.join(INC, h) for h in ("trm.h", "trm_cuda.h", "trm_workload.h")]
    cu = [("kernels_f64", ["-fmad=false"]), ("kernels_f32", []), ("kernels_aux", ["-fmad=false"]), ("trm_cuda", [])]
    objs = []
    for name, extra in cu:
        src = os.path.join(CSRC, name + ".cu")
        obj = os.path.join(OBJ, name + ".o")
        if force or _newer(obj, [src] + headers):
            _run([NVCC] + NVCC_COMMON + extra + ["-c", src, "-o", obj], verbose)
        objs.append(obj)
    cuda_so = os.path.join(LIB, "libtrm_cuda.so")
    if force or _newer(cuda_so, objs):
        _run([NVCC] + GENCODE + ["-shared", "-o", cuda_so] + objs + ["-cudart", "static"], verbose)
    host_src = [os.path.join(CSRC, f) for f in ("trm_host.c", "trm_workload.c")]
    host_so = os.path.join(LIB, "libtrm.so")
    if force or _newer(host_so, host_src + headers + [cuda_so]):
        _run(["gcc", "-O2", "-std=gnu99", "-Wall", "-fPIC", "-shared", "-I" + INC, "-o", host_so] + host_src +
             ["-L" + LIB, "-ltrm_cuda", "-Wl,-rpath,$ORIGIN", "-lm", "-lpthread"], verbose)
    return host_so, cuda_so


def build_oracle(verbose=False):
    """Builds the CPU oracle (test infrastructure) and, when /root/reference is mounted, oracle/_ref."""
    _run(["make", "-C", os.path.join(ROOT, "oracle")], verbose)


if __name__ == "__main__":
    build(verbose=True, force="--force" in sys.argv)
    build_oracle(verbose=True)
