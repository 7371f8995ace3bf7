"""Build libldpc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m short_ldpc_decoding_osd_b200.build [--force]

The library is written next to this file so that it travels with the source tree.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_PATH = os.path.join(HERE, "libldpc_b200.so")
BUILD_DIR = os.path.join(ROOT, "build", "ldpc_b200")
SOURCES = ["handle.cu", "nms.cu", "nms_qc.cu", "osd.cu", "osd_pair.cu", "osd3.cu", "osd_blocks.cu", "osd_pb.cu", "aux.cu", "framegen.cu", "pipeline.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", INCLUDE, "-I", CSRC,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libldpc_b200.so cannot be built")


def _stamp() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)) + ["../../include/ldpc_b200.h"]:
        p = os.path.normpath(os.path.join(CSRC, name))
        if os.path.isfile(p):
            with open(p, "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    # flags without the include paths: the hash must be the same wherever the tree is checked out (the GPU box runs a copy
    # under another root, and profiles/r02_traffic.json carries the stamp of the library it was measured on)
    h.update(" ".join(f for f in NVCC_FLAGS if f not in (INCLUDE, CSRC)).encode())
    return h.hexdigest()


BOUNDS_LIB_PATH = os.path.join(HERE, "libldpc_b200_bounds.so")


def build(force: bool = False, verbose: bool = False, bounds: bool = False) -> str:
    """bounds=True: the same sources with -DLDPCB_BOUNDS (device-side asserts on every data-dependent index) into
    libldpc_b200_bounds.so; load it with LDPCB_B200_LIB=<path> (tests run scripts/sanitize_case.py against it)."""
    lib_path = BOUNDS_LIB_PATH if bounds else LIB_PATH
    build_dir = BUILD_DIR + ("_bounds" if bounds else "")
    os.makedirs(build_dir, exist_ok=True)
    stamp_file = os.path.join(build_dir, "stamp")
    stamp = _stamp()
    if not force and os.path.exists(lib_path) and os.path.exists(stamp_file):
        with open(stamp_file) as f:
            if f.read().strip() == stamp:
                return lib_path
    nvcc = _nvcc()
    extra = (["-Xptxas", "-v"] if verbose else []) + (["-DLDPCB_BOUNDS"] if bounds else [])

    def compile_one(src: str) -> str:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(6, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", lib_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return lib_path


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv, bounds="--bounds" in sys.argv)
    print(path)
