"""Builds libiivision_b200.so in-tree with nvcc for sm_100a (no JIT cache).

    python -m iivision_b200._build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""

import os
import shlex
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libiivision_b200.so")
SOURCES = ["iiv_core.cu", "iiv_lut.cu", "iiv_tables.cu", "iiv_scorer.cu",
           "iiv_encoder.cu", "iiv_stream.cu", "iiv_deflate.cu", "iiv_factored.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "-Xcompiler", "-Wall"]
PER_FILE = {"iiv_lut.cu": ["-fmad=false"]}
# development only: extra nvcc flags (e.g. -DIIV_... experiment switches) for a --force build
EXTRA = shlex.split(os.environ.get("IIV_NVCC_FLAGS", ""))


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC)
               if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "iivision_b200.h"))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        op = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(op)
        if force or _stale(op, [sp] + headers):
            cmd = [nvcc, *ARCH, *COMMON, *PER_FILE.get(src, []), *EXTRA, "-c", sp, "-o", op]
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
                print(" ".join(cmd))
            procs.append((src, subprocess.Popen(
                cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(out)
        if p.returncode != 0:
            failed = True
            print("nvcc failed on %s" % src, file=sys.stderr)
    if failed:
        raise RuntimeError("CUDA build failed")
    if force or _stale(OUT, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", OUT, *objs, "-lcudart"]
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
