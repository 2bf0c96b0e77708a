"""In-tree build of libclsph_cuda.so (sm_100a only) with nvcc.

nvcc cross-compiles without a GPU, so this runs in the build container; the resulting .so is
git-ignored but travels to the GPU box with the repository snapshot.
"""
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libclsph_cuda.so")
SOURCES = ["context.cu", "sort.cu", "grid.cu", "neighbors.cu", "subgrid.cu", "integrate.cu", "dist.cu"]
HEADERS = ["common.cuh", "kernels.cuh", "dist.cuh", "pair_terms.cuh", "subview.cuh"]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path():
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else shutil.which("nvcc")


def _newest_input():
    files = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    files += [os.path.join(_ROOT, "include", "clsph_cuda.h"), os.path.join(_ROOT, "include", "clsph", "clsph_types.h")]
    return max(os.path.getmtime(f) for f in files)


def build(force=False, verbose=False, extra_flags=()):
    """Compile every CUDA source for sm_100a into libclsph_b200/libclsph_cuda.so."""
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_input():
        return LIB_PATH
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found: libclsph_cuda.so cannot be built (and there is no CPU fallback)")
    obj_dir = os.path.join(_ROOT, "build", "obj")
    os.makedirs(obj_dir, exist_ok=True)
    common = [nvcc, *ARCH_FLAGS, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
              # the system g++ (this image exports CXX=/opt/gcc/bin/g++, which lacks pieces)
              "-ccbin", shutil.which("g++") or "g++",
              "-I" + os.path.join(_ROOT, "include"), "-I" + CSRC, *extra_flags]
    if verbose:
        common += ["-Xptxas", "-v"]
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        procs.append((src, subprocess.Popen(common + ["-c", os.path.join(CSRC, src), "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write("---- %s\n%s" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.run([nvcc, *ARCH_FLAGS, "-shared", "-ccbin", shutil.which("g++") or "g++", "-o", LIB_PATH, *objs,
                    "-lcudart", "-ldl"], check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
