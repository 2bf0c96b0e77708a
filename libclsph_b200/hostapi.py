"""ctypes binding of libclsph_host.so: the C++ drop-in classes (sph_simulation, scene,
houdini_file_saver) through the C wrappers of libclsph_b200/host/host_capi.cpp."""
import ctypes
import os
import subprocess

import numpy as np

from . import build as _build
from .abi import PARTICLE, PrecomputedKernelValues, SimulationParameters, particle_ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libclsph_host.so")
CLI_PATH = os.path.join(_HERE, "clsphparticles")
_lib = None


def build(force=False):
    """make -C libclsph_b200/host (g++): libclsph_host.so and the clsphparticles driver."""
    _build.build()
    srcs = [os.path.join(_HERE, "host", f) for f in os.listdir(os.path.join(_HERE, "host"))]
    for folder, _, files in os.walk(os.path.join(os.path.dirname(_HERE), "include", "clsph")):
        srcs += [os.path.join(folder, f) for f in files]
    newest = max(os.path.getmtime(f) for f in srcs)
    if force or not os.path.exists(LIB_PATH) or not os.path.exists(CLI_PATH) or os.path.getmtime(LIB_PATH) < newest:
        subprocess.run(["make", "-C", os.path.join(_HERE, "host"), "-B"], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = ctypes.CDLL(LIB_PATH)
        vp = ctypes.c_void_p
        L.clsph_host_load_settings.argtypes = [ctypes.c_char_p, ctypes.c_char_p, vp, vp, vp, vp, vp, ctypes.c_char_p]
        L.clsph_host_scene_load.argtypes = [ctypes.c_char_p, vp, vp, vp, vp, vp, vp]
        L.clsph_host_write_frames.argtypes = [ctypes.c_char_p, vp, vp, ctypes.c_int]
        L.clsph_host_write_frames_as.argtypes = [ctypes.c_char_p, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.clsph_host_simulate.argtypes = [vp, vp, ctypes.c_float, vp, vp, ctypes.c_size_t, vp, ctypes.c_uint32,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp]
        _lib = L
    return _lib


def _vp(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def load_settings(fluid_json, sim_json):
    p, t = SimulationParameters(), PrecomputedKernelValues()
    vol, wa, se = ctypes.c_float(), ctypes.c_int(), ctypes.c_int()
    err = ctypes.create_string_buffer(256)
    rc = lib().clsph_host_load_settings(fluid_json.encode(), sim_json.encode(), ctypes.byref(p), ctypes.byref(t),
                                        ctypes.byref(vol), ctypes.byref(wa), ctypes.byref(se), err)
    if rc:
        raise RuntimeError(err.value.decode())
    return p, t, vol.value, dict(write_all_frames=bool(wa.value), serialize=bool(se.value))


def scene_load(name, cwd):
    """scene::load(name) with `cwd` as working directory (it must contain scenes/)."""
    old = os.getcwd()
    os.chdir(cwd)
    try:
        fc, nv, ni = ctypes.c_uint32(), ctypes.c_size_t(), ctypes.c_size_t()
        if lib().clsph_host_scene_load(name.encode(), ctypes.byref(fc), ctypes.byref(nv), ctypes.byref(ni), None, None, None):
            raise RuntimeError("scene::load(%s) failed" % name)
        normals = np.zeros(3 * fc.value, dtype=np.float32)
        vertices = np.zeros(nv.value, dtype=np.float32)
        indices = np.zeros(ni.value, dtype=np.uint32)
        lib().clsph_host_scene_load(name.encode(), ctypes.byref(fc), ctypes.byref(nv), ctypes.byref(ni), _vp(normals),
                                    _vp(vertices), _vp(indices))
        return normals, vertices, indices
    finally:
        os.chdir(old)


def write_frames(prefix, particles, params, frames=1, fmt=None, packed=False):
    """houdini_file_saver::writeFrameToFile `frames` times. fmt "geo" / "bgeo" sets houdini_file_saver::format;
    packed=True goes through writeFramePoints (the seven floats per particle that clsph_frame_begin packs)."""
    if fmt is None and not packed:
        lib().clsph_host_write_frames(prefix.encode(), particle_ptr(particles), ctypes.byref(params), frames)
    else:
        lib().clsph_host_write_frames_as(prefix.encode(), particle_ptr(particles), ctypes.byref(params), frames,
                                         1 if fmt == "bgeo" else 0, 1 if packed else 0)


def simulate(params, terms, initial_volume, normals, vertices, indices, frames, policy=0, callbacks=True, cwd=None):
    """sph_simulation::simulate. Returns (particles, params_after, callback_calls)."""
    n = params.particles_count
    out = np.zeros(n, dtype=PARTICLE)
    p_out = SimulationParameters()
    calls = ctypes.c_long()
    old = os.getcwd()
    if cwd:
        os.chdir(cwd)
    try:
        rc = lib().clsph_host_simulate(ctypes.byref(params), ctypes.byref(terms), ctypes.c_float(initial_volume),
                                       _vp(normals), _vp(vertices), vertices.size, _vp(indices), indices.size // 3,
                                       frames, policy, 1 if callbacks else 0, particle_ptr(out), ctypes.byref(p_out),
                                       ctypes.byref(calls))
    finally:
        os.chdir(old)
    if rc:
        raise RuntimeError("clsph_host_simulate rc=%d" % rc)
    return out, p_out, calls.value
