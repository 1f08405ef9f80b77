"""In-tree build of libsp_b200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsp_b200.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["sp_system.cu", "sp_cells.cu", "sp_sweep.cu", "sp_reduce.cu", "sp_isph.cu", "sp_program.cu", "sp_slab.cu", "sp_generate.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(src_paths, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(p) > t for p in src_paths)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "sp_b200.h"))
    env = dict(os.environ)
    # the image exports CXX/CC=/opt/gcc wrappers; nvcc should use the system g++
    host = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if not force and not _newer([path] + headers, obj):
            return obj, ""
        cmd = [nvcc, *ARCH, *FLAGS, "-c", path, "-o", obj]
        if host:
            cmd += ["-ccbin", host]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in results]
    log = "".join(l for _, l in results)
    if log:
        with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
            f.write(log)
    if force or _newer(objs, OUT):
        cmd = [nvcc, *ARCH, "-shared", "-o", OUT, *objs]  # cudart is linked statically (nvcc default)
        if host:
            cmd += ["-ccbin", host]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose and log:
        sys.stderr.write(log)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
