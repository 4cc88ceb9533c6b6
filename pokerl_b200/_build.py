"""In-tree build of the CUDA extension (sm_100a only; no multi-arch, no JIT cache).

`python -m pokerl_b200._build` or `__graft_entry__.build()` compiles
pokerl_b200/csrc/pkb_kernels.cu into pokerl_b200/_lib/libpokerl_b200.so with nvcc.
nvcc cross-compiles without a GPU, so this also runs in the CPU-only build container;
the built .so travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "_lib")
LIB = os.path.join(LIB_DIR, "libpokerl_b200.so")
SOURCES = [os.path.join(CSRC, "pkb_kernels.cu")]
HEADERS = [os.path.join(CSRC, f) for f in ("pkb_common.cuh", "pkb_philox.cuh", "pkb_eval.cuh", "pkb_game.cuh")] + [
    os.path.join(os.path.dirname(PKG), "include", "pokerl_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # the reference rounds a*b and +c separately (game.py:370-371, 675-678)
    "-shared", "-Xcompiler", "-fPIC",
    "-cudart", "static",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile the extension if it is missing or older than its sources; returns the .so path."""
    if not force and not is_stale():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB + ".tmp"] + SOURCES
    # nvcc needs a host compiler with a full libstdc++; prefer the system g++
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
