"""Builds tests/emu/_build/libclsph_emu.so: the CUDA sources of libclsph_b200/csrc compiled for the
CPU against the test-only emulator (cuda_emu.h). TEST INFRASTRUCTURE -- the product never loads it.

The sources are used as they are; a small textual pass turns the CUDA-only syntax into C++:
  kernel<<<grid, block, smem, stream>>>(args);   ->  emu::launch(dim3(grid), dim3(block), smem, [&]() { kernel(args); });
  extern __shared__ T name[];                    ->  T* name = reinterpret_cast<T*>(emu::dyn_smem());
  __shared__                                     ->  static thread_local
  #include <cuda_runtime.h> / <nccl.h>           ->  "cuda_emu.h" / "nccl_emu.h"
  dlopen / dlsym                                 ->  emu::fake_dlopen / emu::fake_dlsym  (NCCL stand-in)
The two inline-PTX helpers (lanemask, cp.async) have `#ifdef CLSPH_EMU` branches in the sources.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "libclsph_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(OUT, "libclsph_emu.so")
EMU_SOURCES = ["cuda_emu.cpp", "nccl_emu.cpp"]


def _match_forward(text, pos, open_ch, close_ch):
    """Index just past the bracket that closes the one at text[pos]."""
    depth = 0
    i = pos
    while i < len(text):
        c = text[i]
        if c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced %s at %d" % (open_ch, pos))


def _split_top_level(s):
    parts, depth, cur = [], 0, []
    for c in s:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(c)
    parts.append("".join(cur).strip())
    return parts


def rewrite_launches(text):
    out = []
    pos = 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            out.append(text[pos:])
            break
        # kernel expression: identifier, optionally followed by <template args>
        j = k
        while j > 0 and text[j - 1].isspace():
            j -= 1
        if text[j - 1] == ">":
            depth, i = 0, j - 1
            while i >= 0:
                if text[i] == ">":
                    depth += 1
                elif text[i] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                i -= 1
            j = i
        while j > 0 and (text[j - 1].isalnum() or text[j - 1] in "_:"):
            j -= 1
        kernel = text[j:k].strip()
        e = text.find(">>>", k)
        cfg = _split_top_level(text[k + 3:e])
        a = e + 3
        while text[a].isspace():
            a += 1
        assert text[a] == "(", "launch of %s without an argument list" % kernel
        b = _match_forward(text, a, "(", ")")
        args = text[a + 1:b - 1]
        c = b
        while text[c].isspace():
            c += 1
        assert text[c] == ";", "launch of %s is not a statement" % kernel
        grid, block = cfg[0], cfg[1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        out.append(text[pos:j])
        out.append("emu::launch(dim3(%s), dim3(%s), (size_t)(%s), [&]() { %s(%s); });" % (grid, block, smem, kernel, args))
        pos = c + 1
    return "".join(out)


def preprocess(text):
    text = text.replace("#include <cuda_runtime.h>", '#include "cuda_emu.h"')
    text = text.replace("#include <nccl.h>", '#include "nccl_emu.h"')
    text = re.sub(r"\bdlopen\(", "emu::fake_dlopen(", text)
    text = re.sub(r"\bdlsym\(", "emu::fake_dlsym(", text)
    text = re.sub(r"extern\s+__shared__\s+([A-Za-z_][A-Za-z_0-9]*)\s+([A-Za-z_][A-Za-z_0-9]*)\s*\[\s*\]\s*;",
                  r"\1* \2 = reinterpret_cast<\1*>(emu::dyn_smem());", text)
    text = re.sub(r"\b__shared__\b", "static thread_local", text)
    return rewrite_launches(text)


def _inputs():
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    files += [os.path.join(HERE, f) for f in EMU_SOURCES + ["cuda_emu.h", "nccl_emu.h", "build_emu.py"]]
    files += [os.path.join(ROOT, "include", "clsph_cuda.h"), os.path.join(ROOT, "include", "clsph", "clsph_types.h")]
    return files


def build(force=False, verbose=False):
    """EMU_ASAN=1 in the environment builds libclsph_emu_asan.so instead: AddressSanitizer with plain
    malloc'ed device memory, i.e. red zones around every device allocation (a memcheck of the kernels). Run as
      EMU_ASAN=1 LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 \
          python -m pytest tests/test_emu_kernel_logic.py tests/test_emu_multi_rank.py"""
    asan = os.environ.get("EMU_ASAN") == "1"
    lib_path = LIB_PATH.replace(".so", "_asan.so") if asan else LIB_PATH
    if not force and os.path.exists(lib_path) and os.path.getmtime(lib_path) >= max(os.path.getmtime(f) for f in _inputs()):
        return lib_path
    gen = os.path.join(OUT, "gen")
    os.makedirs(gen, exist_ok=True)
    sources = []
    for f in sorted(os.listdir(CSRC)):
        if not f.endswith((".cu", ".cuh")):
            continue
        with open(os.path.join(CSRC, f)) as fh:
            text = preprocess(fh.read())
        name = f.replace(".cu", ".cpp") if f.endswith(".cu") else f
        with open(os.path.join(gen, name), "w") as fh:
            fh.write(text)
        if f.endswith(".cu"):
            sources.append(os.path.join(gen, name))
    sources += [os.path.join(HERE, f) for f in EMU_SOURCES]
    flags = ["-O2", "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-strict-aliasing", "-pthread",
             "-Wall", "-Wno-unused-function", "-Wno-unused-variable", "-Wno-unknown-pragmas", "-Wno-sign-compare",
             "-I" + HERE, "-I" + gen, "-I" + os.path.join(ROOT, "include")]
    if asan:
        flags += ["-fsanitize=address", "-fno-omit-frame-pointer", "-DEMU_ASAN=1", "-O1"]
    try:
        with open("/proc/cpuinfo") as fh:
            if " fma " in fh.read():
                flags.append("-mfma")
    except OSError:
        pass
    objs, procs = [], []
    for s in sources:
        o = os.path.join(OUT, os.path.basename(s) + (".asan.o" if asan else ".o"))
        objs.append(o)
        procs.append((s, subprocess.Popen(["g++", *flags, "-c", s, "-o", o], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or (verbose and out):
            sys.stderr.write("---- %s\n%s" % (s, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("emulator build failed")
    subprocess.run(["g++", "-shared", "-pthread", *(["-fsanitize=address"] if asan else []), "-o", lib_path, *objs, "-ldl"], check=True)
    return lib_path


if __name__ == "__main__":
    print(build(force=True, verbose=True))
