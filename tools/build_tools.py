"""Builds the small CUDA micro-benchmarks under tools/ (sm_100a)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def build():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = os.path.join(HERE, "fp64_peak.cu")
    out = os.path.join(HERE, "fp64_peak")
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", src, "-o", out]
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(build())
