"""Builds the sm_100a shared library in-tree with nvcc (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG_DIR)
SRC = os.path.join(PKG_DIR, "csrc", "hs_kernels.cu")
# the tick kernel once more with IEEE arithmetic (HS_OPT_EXACT_MATH, csrc/hs_tick_exact.cu): no FMA contraction
SRC_EXACT = os.path.join(PKG_DIR, "csrc", "hs_tick_exact.cu")
LIB = os.path.join(PKG_DIR, "libhs_b200.so")
OBJ_DIR = os.path.join(PKG_DIR, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
EXACT_FLAGS = ["-fmad=false", "-prec-div=true", "-prec-sqrt=true"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; the CUDA toolkit is required to build libhs_b200.so")
    return p


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    csrc = os.path.dirname(SRC)
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh"))]     # unity build: one TU
    deps += [os.path.join(REPO, "include", "hs_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=(), out=None):
    if out is None and not force and not needs_build():
        return LIB
    lib_out = out or LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    inc = ["-I", os.path.join(REPO, "include")]
    tag = "" if out is None else "_" + os.path.splitext(os.path.basename(out))[0]
    objs = [os.path.join(OBJ_DIR, f"hs_kernels{tag}.o"), os.path.join(OBJ_DIR, f"hs_tick_exact{tag}.o")]
    cmds = [[nvcc_path(), *NVCC_FLAGS, *extra_flags, *inc, "-c", "-o", objs[0], SRC],
            [nvcc_path(), *NVCC_FLAGS, *EXACT_FLAGS, *extra_flags, *inc, "-c", "-o", objs[1], SRC_EXACT]]
    if verbose:
        for cmd in cmds:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd))
    procs = [subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for cmd in cmds]   # in parallel
    outs = [p.communicate()[0] for p in procs]
    if any(p.returncode != 0 for p in procs):
        raise RuntimeError("nvcc failed:\n" + "\n".join(outs))
    link = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", lib_out, *objs]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + r.stdout + r.stderr)
    if verbose:
        print("\n".join(outs) + r.stdout + r.stderr)
    return lib_out


if __name__ == "__main__":
    defs = [f"-D{sys.argv[i + 1]}" for i, a in enumerate(sys.argv[:-1]) if a == "--define"]     # debug builds (tools/)
    out = next((sys.argv[i + 1] for i, a in enumerate(sys.argv[:-1]) if a == "--out"), None)        # e.g. a timing build for tools/
    print(build(force="--force" in sys.argv, verbose=True, extra_flags=defs, out=out))
