"""Builds csrc/*.cu into lib/libyolox_b200.so for sm_100a with nvcc (in-tree, no JIT cache)."""
import glob
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "lib", "libyolox_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--cudart", "static"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(os.path.dirname(_HERE), "include", "yolox_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libyolox_b200.so (there is no fallback path)")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    objs = []
    build_dir = os.path.join(_HERE, "lib", "obj")
    os.makedirs(build_dir, exist_ok=True)
    def compile_one(src):
        obj = os.path.join(build_dir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        subprocess.check_call(cmd)
        return obj

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:   # translation units are independent
        objs = list(pool.map(compile_one, sources()))
    subprocess.check_call([nvcc, "-shared", "--cudart", "static", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
