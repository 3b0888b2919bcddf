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


def kernel_rev() -> str:
    """Revision of the conv kernels and their planner: launch shapes persisted by the tuner (plan.TuneCache) are valid
    for exactly this revision."""
    import hashlib
    h = hashlib.sha1()
    for name in ("yx_conv.cu", "yx_internal.h", "yx_ptx.cuh"):
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:12]


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
    """Compiles csrc/*.cu and links lib/libyolox_b200.so.  Safe under `torchrun --nproc-per-node N` on a fresh checkout:
    an inter-process file lock serialises the build, staleness is re-checked once the lock is held (the other ranks find
    the finished library), objects go to a per-process directory and the library is moved into place atomically, so no
    process can ever dlopen a half-written file."""
    if not force and not _stale():
        return LIB
    import fcntl
    import tempfile
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    with open(os.path.join(os.path.dirname(LIB), ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():       # another process built it while this one waited for the lock
                return LIB
            return _build_locked(verbose, tempfile.mkdtemp(prefix=f"obj.{os.getpid()}.", dir=os.path.dirname(LIB)))
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool, build_dir: str) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        shutil.rmtree(build_dir, ignore_errors=True)
        raise RuntimeError("nvcc not found: cannot build libyolox_b200.so (there is no fallback path)")
    # (source, extra defines, object name): yx_conv.cu is compiled once for planning / dispatch and once per activation
    # (-DYX_CONV_ACT_SLICE=<yx_act>), so its ~100 kernel instantiations build in parallel
    units = []
    for src in sources():
        base = os.path.basename(src)[:-3]
        units.append((src, [], base))
        if base == "yx_conv":
            units += [(src, [f"-DYX_CONV_ACT_SLICE={a}", "-diag-suppress", "177"], f"{base}_act{a}") for a in range(5)]

    def compile_one(unit):
        src, defs, name = unit
        obj = os.path.join(build_dir, name + ".o")
        cmd = [nvcc] + NVCC_FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + \
            (["-DYX_DEBUG_TRAP"] if os.environ.get("YX_DEBUG_TRAP") else []) + os.environ.get("YX_NVCC_DEFS", "").split() + \
            ["-c", src, "-o", obj]
        subprocess.check_call(cmd)
        return obj

    from concurrent.futures import ThreadPoolExecutor
    try:
        with ThreadPoolExecutor(max_workers=min(12, os.cpu_count() or 1)) as pool:   # translation units are independent
            objs = list(pool.map(compile_one, units))
        tmp_lib = os.path.join(build_dir, "libyolox_b200.so.tmp")
        subprocess.check_call([nvcc, "-shared", "--cudart", "static", "-o", tmp_lib] + objs)
        os.replace(tmp_lib, LIB)                 # atomic: readers see the old library or the complete new one
    finally:
        shutil.rmtree(build_dir, ignore_errors=True)   # only the .so needs to travel with the tree
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
