"""CPU tests: binary records, the C-ABI library loads and exports what the header declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from libclsph_b200 import abi, build, capi
from tests.helpers import ROOT


def test_record_sizes():
    assert abi.PARTICLE.itemsize == 80
    assert abi.PARTICLE.fields["density"][1] == 64
    assert abi.PARTICLE.fields["grid_index"][1] == 72
    assert ctypes.sizeof(abi.SimulationParameters) == 128
    assert ctypes.sizeof(abi.PrecomputedKernelValues) == 20


def test_header_compiles_as_c_and_cxx(tmp_path):
    import subprocess
    src = '#include "clsph_cuda.h"\nint main(void){return sizeof(particle)==80?0:1;}\n'
    for comp, name in (("gcc", "t.c"), ("g++", "t.cpp")):
        f = tmp_path / name
        f.write_text(src)
        exe = tmp_path / (name + ".out")
        subprocess.run([comp, "-I", os.path.join(ROOT, "include"), str(f), "-o", str(exe)], check=True)
        assert subprocess.run([str(exe)]).returncode == 0


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "clsph_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(clsph_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    build.build()
    lib = ctypes.CDLL(build.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), "libclsph_cuda.so does not export %s" % name
    assert sorted(capi.SYMBOLS) == names


def test_no_cpu_fallback_without_gpu():
    lib = capi.load_library()
    if lib.clsph_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.ClsphError) as e:
        capi.Context(1024)
    assert e.value.code == capi.E_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/ (test infrastructure)."""
    pkg = os.path.join(ROOT, "libclsph_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "oracle/" not in text and "liboracle" not in text, f


def test_product_never_loads_the_emulator_build():
    """tests/emu/ (the CUDA sources compiled for the CPU) is test infrastructure: nothing in the product
    package or in bench.py may import it or name its library; the only trace in the product is the pair
    of `#ifdef CLSPH_EMU` branches around inline PTX in csrc/."""
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    for dirpath, _, names in os.walk(os.path.join(ROOT, "libclsph_b200")):
        files += [os.path.join(dirpath, f) for f in names if f.endswith((".py", ".cpp", ".h", ".hpp"))]
    for path in files:
        text = open(path, errors="ignore").read()
        assert "tests.emu" not in text and "libclsph_emu" not in text and "build_emu" not in text and "cuda_emu" not in text, path
