"""ctypes binding of oracle/_ref/libclsph_ref.so -- the reference's own sources compiled by
oracle/build_ref.sh behind the OpenCL shim. TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).

The library exists only after build_ref.sh ran in a container that has /root/reference; it is
git-ignored but travels to the GPU box with the snapshot. `available()` says whether it is
there; nothing in the product depends on it.
"""
import ctypes
import os
import subprocess

import numpy as np

from libclsph_b200.abi import PARTICLE, PrecomputedKernelValues, SimulationParameters, particle_ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libclsph_ref.so")
WORK_DIR = os.path.join(_HERE, "_ref", "gen")  # holds kernels/sph.cl, which simulate() must find
REFERENCE = os.environ.get("REFERENCE", "/root/reference")

_lib = None


def can_build():
    return os.path.isfile(os.path.join(REFERENCE, "libclsph", "kernels", "sph.cl"))


def build(force=False):
    """Run oracle/build_ref.sh if the reference tree is present. Returns True if the .so exists."""
    if can_build() and (force or not os.path.exists(LIB_PATH)):
        subprocess.run([os.path.join(_HERE, "build_ref.sh")], check=True, stdout=subprocess.DEVNULL)
    return os.path.exists(LIB_PATH)


def available():
    return os.path.exists(LIB_PATH) and os.path.isfile(os.path.join(WORK_DIR, "kernels", "sph.cl"))


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref is not built (run oracle/build_ref.sh where /root/reference exists)")
        L = ctypes.CDLL(LIB_PATH)
        vp = ctypes.c_void_p
        L.clsph_ref_load_settings.restype = ctypes.c_int
        L.clsph_ref_load_settings.argtypes = [ctypes.c_char_p, ctypes.c_char_p, vp, vp, vp, vp, vp]
        L.clsph_ref_scene_load.restype = ctypes.c_int
        L.clsph_ref_scene_load.argtypes = [ctypes.c_char_p, ctypes.c_char_p, vp, vp, vp, vp, vp, vp]
        L.clsph_ref_simulate.restype = ctypes.c_int
        L.clsph_ref_simulate.argtypes = [ctypes.c_char_p, vp, vp, ctypes.c_float, vp, vp, ctypes.c_size_t, vp,
                                         ctypes.c_uint32, vp, ctypes.c_int, ctypes.c_int, vp, vp, vp]
        L.clsph_ref_kernel_locate_in_grid.restype = ctypes.c_int
        L.clsph_ref_kernel_locate_in_grid.argtypes = [vp, vp, vp]
        L.clsph_ref_kernel_density_pressure.restype = ctypes.c_int
        L.clsph_ref_kernel_density_pressure.argtypes = [vp, vp, vp, vp, vp]
        L.clsph_ref_kernel_forces.restype = ctypes.c_int
        L.clsph_ref_kernel_forces.argtypes = [vp, vp, vp, vp, vp]
        L.clsph_ref_kernel_advection_collision.restype = ctypes.c_int
        L.clsph_ref_kernel_advection_collision.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp, ctypes.c_uint32]
        L.clsph_ref_write_frames.restype = ctypes.c_int
        L.clsph_ref_write_frames.argtypes = [ctypes.c_char_p, vp, vp, ctypes.c_int]
        L.clsph_ref_num_threads.restype = ctypes.c_int
        L.clsph_ref_set_num_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def _vp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def load_settings(fluid_json, sim_json):
    """The reference's sph_simulation::load_settings. Returns (params, terms, initial_volume, flags)."""
    p, t = SimulationParameters(), PrecomputedKernelValues()
    vol, wa, se = ctypes.c_float(), ctypes.c_int(), ctypes.c_int()
    rc = lib().clsph_ref_load_settings(fluid_json.encode(), sim_json.encode(), ctypes.byref(p), ctypes.byref(t),
                                       ctypes.byref(vol), ctypes.byref(wa), ctypes.byref(se))
    if rc:
        raise RuntimeError("clsph_ref_load_settings rc=%d" % rc)
    return p, t, vol.value, dict(write_all_frames=bool(wa.value), serialize=bool(se.value))


def scene_load(root_dir, name):
    """The reference's scene::load (tinyobj + face normals). Returns (normals, vertices, indices)."""
    fc, nv, ni = ctypes.c_uint32(), ctypes.c_size_t(), ctypes.c_size_t()
    rc = lib().clsph_ref_scene_load(root_dir.encode(), name.encode(), ctypes.byref(fc), ctypes.byref(nv),
                                    ctypes.byref(ni), None, None, None)
    if rc:
        raise RuntimeError("clsph_ref_scene_load rc=%d" % rc)
    normals = np.zeros(3 * fc.value, dtype=np.float32)
    vertices = np.zeros(nv.value, dtype=np.float32)
    indices = np.zeros(ni.value, dtype=np.uint32)
    lib().clsph_ref_scene_load(root_dir.encode(), name.encode(), ctypes.byref(fc), ctypes.byref(nv),
                               ctypes.byref(ni), _vp(normals), _vp(vertices), _vp(indices))
    return normals, vertices, indices


def simulate(params, terms, initial_volume, scene, initial=None, substeps=1, record_all=False):
    """The reference's sph_simulation::simulate for `substeps` sub-steps.

    Returns (states, params_after, seconds): states is [substeps, N] if record_all else [N]."""
    n = params.particles_count
    states = np.zeros((substeps, n) if record_all else (n,), dtype=PARTICLE)
    p_out = SimulationParameters()
    secs = np.zeros(substeps, dtype=np.float64)
    rc = lib().clsph_ref_simulate(WORK_DIR.encode(), ctypes.byref(params), ctypes.byref(terms),
                                  ctypes.c_float(initial_volume), _vp(scene.face_normals), _vp(scene.vertices),
                                  scene.vertices.size, _vp(scene.indices), scene.face_count,
                                  None if initial is None else particle_ptr(initial), substeps,
                                  1 if record_all else 0, particle_ptr(states.reshape(-1)),
                                  ctypes.byref(p_out), _vp(secs))
    if rc:
        raise RuntimeError("clsph_ref_simulate rc=%d" % rc)
    return states, p_out, secs


def kernel_locate_in_grid(particles, params):
    out = np.zeros_like(particles)
    rc = lib().clsph_ref_kernel_locate_in_grid(particle_ptr(particles), particle_ptr(out), ctypes.byref(params))
    assert rc == 0, rc
    return out


def kernel_density_pressure(particles, params, terms, table):
    out = np.zeros_like(particles)
    rc = lib().clsph_ref_kernel_density_pressure(particle_ptr(particles), particle_ptr(out), ctypes.byref(params),
                                                 ctypes.byref(terms), _vp(table))
    assert rc == 0, rc
    return out


def kernel_forces(particles, params, terms, table):
    out = np.zeros_like(particles)
    rc = lib().clsph_ref_kernel_forces(particle_ptr(particles), particle_ptr(out), ctypes.byref(params),
                                       ctypes.byref(terms), _vp(table))
    assert rc == 0, rc
    return out


def kernel_advection_collision(particles, params, terms, scene):
    out = np.zeros_like(particles)
    rc = lib().clsph_ref_kernel_advection_collision(particle_ptr(particles), particle_ptr(out), ctypes.byref(params),
                                                    ctypes.byref(terms), _vp(scene.face_normals), _vp(scene.vertices),
                                                    scene.vertices.size, _vp(scene.indices), scene.face_count)
    assert rc == 0, rc
    return out


def write_frames(prefix, particles, params, frames=1):
    """The reference's houdini_file_saver: writes <prefix>frames/frame000000K.geo, K = 1..frames."""
    rc = lib().clsph_ref_write_frames(prefix.encode(), particle_ptr(particles), ctypes.byref(params), frames)
    assert rc == 0, rc


def num_threads():
    return int(lib().clsph_ref_num_threads())


def set_num_threads(n):
    lib().clsph_ref_set_num_threads(int(n))
