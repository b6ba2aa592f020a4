"""
tess_b200.build -- compiles libtess_b200.so (CUDA, sm_100a only) in-tree with nvcc.

    python -m tess_b200.build [--force]

nvcc cross-compiles without a GPU; the built library sits next to this file
(tess_b200/libtess_b200.so), is git-ignored and travels to the GPU box with the repository
snapshot.  There is no CPU build of the product: tess_b200 fails loudly when the library is
missing (tess_b200/_lib.py).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libtess_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-I", INCLUDE,
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def sources():
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(INCLUDE, "tess_b200.h"))
    return deps


def up_to_date():
    return os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in sources())


def build(force=False, verbose=False, out=None, defines=()):
    """Build tess_b200/libtess_b200.so if it is older than its sources.  Returns its path.
    `out` / `defines` build a tuning variant elsewhere (load it with TESS_B200_LIB=<path>)."""
    if out is None and not defines and not force and up_to_date():
        return LIB
    out = out or LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] + \
          ["-o", out, os.path.join(CSRC, "tess_b200.cu")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stdout))
    if verbose:
        print(r.stdout)
    return out


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("-v", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("-D", action="append", default=[])
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.v, out=a.out, defines=a.D))
