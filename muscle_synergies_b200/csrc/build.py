"""Builds libms_b200.so (all CUDA kernels + the C ABI) for sm_100a with nvcc.

    python muscle_synergies_b200/csrc/build.py [--force] [--verbose]

The library is written next to the package (muscle_synergies_b200/libms_b200.so) so that it
travels with the source tree; cudart is linked statically.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "libms_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-shared", "-Xcompiler", "-fPIC", "-cudart", "static",
]


def sources():
    return sorted(glob.glob(os.path.join(HERE, "*.cu")))


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(HERE, "*.cuh")) + glob.glob(os.path.join(HERE, "*.inc"))
    deps.append(os.path.join(HERE, "..", "..", "include", "ms_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False) -> str:
    if not force and not stale():
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + sources()
    res = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libms_b200.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
