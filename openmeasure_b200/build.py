"""Build libomb200.so (hand-written sm_100a kernels + the C ABI of include/omb200.h) in-tree."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libomb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # every FMA in the library is an explicit fma(); see DESIGN.md
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--shared", "-cudart", "shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [
        os.path.join(os.path.dirname(HERE), "include", "omb200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """nvcc every translation unit for sm_100a (in parallel: no device symbol crosses a file) and link
    libomb200.so in-tree."""
    if not force and not stale():
        return LIB
    import concurrent.futures
    import tempfile
    cflags = [f for f in FLAGS if f != "--shared"] + (["-Xptxas", "-v"] if verbose else [])
    with tempfile.TemporaryDirectory(prefix="omb200_build_") as tmp:
        def compile_one(src):
            obj = os.path.join(tmp, os.path.basename(src)[:-3] + ".o")
            out = subprocess.run([NVCC] + cflags + ["-c", "-o", obj, src], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                 text=True)
            return obj, out.returncode, out.stdout
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
            results = list(pool.map(compile_one, sources()))
        for obj, rc, log in results:
            if verbose or rc:
                print(log, end="")
            if rc:
                raise subprocess.CalledProcessError(rc, "nvcc -c " + obj)
        subprocess.check_call([NVCC, "--shared", "-cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a",
                               "-o", LIB] + [obj for obj, _, _ in results])
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
