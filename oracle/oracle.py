"""ctypes binding of the CPU oracle (oracle/liboracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this module. The product
(libclsph_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

from libclsph_b200.abi import PARTICLE, PrecomputedKernelValues, SimulationParameters, particle_ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

_u32p = ctypes.POINTER(ctypes.c_uint32)
_f32p = ctypes.POINTER(ctypes.c_float)


class _Taps(ctypes.Structure):
    _fields_ = [
        ("keys_input_order", _u32p),
        ("permutation", _u32p),
        ("cell_table", _u32p),
        ("cell_table_capacity", ctypes.c_uint32),
        ("candidate_count", _u32p),
        ("support_count", _u32p),
        ("density", _f32p),
        ("pressure", _f32p),
        ("acceleration", _f32p),
        ("collision_iters", _u32p),
    ]


def build(force=False):
    """Compile liboracle.so in place (gcc, seconds)."""
    src = os.path.join(_HERE, "oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src),
                                                   os.path.getmtime(os.path.join(_HERE, "oracle.h")))):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True,
                   stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.oracle_morton_encode.restype = ctypes.c_uint32
        L.oracle_morton_encode.argtypes = [ctypes.c_uint32] * 3
        L.oracle_morton_decode.argtypes = [ctypes.c_uint32, _u32p]
        L.oracle_derive_constants.restype = ctypes.c_float
        L.oracle_derive_constants.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.oracle_init_particles.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float]
        L.oracle_face_normals.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
        L.oracle_bounds_and_grid.restype = ctypes.c_int
        L.oracle_bounds_and_grid.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_locate_in_grid.argtypes = [ctypes.c_void_p] * 3
        L.oracle_sort_particles.restype = ctypes.c_int
        L.oracle_sort_particles.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
        L.oracle_cell_table.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
        L.oracle_density_pressure.argtypes = [ctypes.c_void_p] * 7
        L.oracle_forces.argtypes = [ctypes.c_void_p] * 5
        L.oracle_advection_collision.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
        L.oracle_step.restype = ctypes.c_int
        L.oracle_step.argtypes = [ctypes.c_void_p] * 7 + [ctypes.c_uint32, ctypes.c_void_p]
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_set_num_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def _vp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def morton_encode(x, y, z):
    return int(lib().oracle_morton_encode(x, y, z))


def morton_decode(key):
    out = (ctypes.c_uint32 * 3)()
    lib().oracle_morton_decode(key, out)
    return tuple(int(v) for v in out)


def derive_constants(params, n_influence):
    """Fill the derived fields of `params`; returns (terms, initial_volume)."""
    terms = PrecomputedKernelValues()
    vol = lib().oracle_derive_constants(ctypes.byref(params), ctypes.byref(terms), n_influence)
    return terms, float(vol)


def init_particles(params, initial_volume):
    buf = np.zeros(params.particles_count, dtype=PARTICLE)
    lib().oracle_init_particles(particle_ptr(buf), ctypes.byref(params), ctypes.c_float(initial_volume))
    return buf


def face_normals(vertices, indices):
    vertices = np.ascontiguousarray(vertices, dtype=np.float32)
    indices = np.ascontiguousarray(indices, dtype=np.uint32)
    nf = indices.size // 3
    out = np.zeros(3 * nf, dtype=np.float32)
    lib().oracle_face_normals(_vp(vertices), _vp(indices), nf, _vp(out))
    return out


def bounds_and_grid(particles, params):
    return int(lib().oracle_bounds_and_grid(particle_ptr(particles), ctypes.byref(params)))


def locate_in_grid(particles, params):
    out = np.empty_like(particles)
    lib().oracle_locate_in_grid(particle_ptr(particles), particle_ptr(out), ctypes.byref(params))
    return out


def sort_particles(particles):
    """Returns (sorted copy, permutation) following the reference's 4-pass radix structure."""
    a = particles.copy()
    scratch = np.empty_like(a)
    perm = np.empty(a.size, dtype=np.uint32)
    rc = lib().oracle_sort_particles(particle_ptr(a), particle_ptr(scratch), a.size, _vp(perm))
    if rc:
        raise ValueError("oracle_sort_particles failed rc=%d" % rc)
    return a, perm


def cell_table(sorted_particles, grid_cell_count):
    t = np.empty(max(int(grid_cell_count), 1), dtype=np.uint32)
    lib().oracle_cell_table(particle_ptr(sorted_particles), sorted_particles.size, grid_cell_count, _vp(t))
    return t[:grid_cell_count]


def density_pressure(particles, params, terms, table):
    out = np.empty_like(particles)
    cand = np.empty(particles.size, dtype=np.uint32)
    supp = np.empty(particles.size, dtype=np.uint32)
    lib().oracle_density_pressure(particle_ptr(particles), particle_ptr(out), ctypes.byref(params),
                                  ctypes.byref(terms), _vp(table), _vp(cand), _vp(supp))
    return out, cand, supp


def forces(particles, params, terms, table):
    out = np.empty_like(particles)
    lib().oracle_forces(particle_ptr(particles), particle_ptr(out), ctypes.byref(params),
                        ctypes.byref(terms), _vp(table))
    return out


def advection_collision(particles, params, scene, max_iters=64):
    out = np.empty_like(particles)
    iters = np.empty(particles.size, dtype=np.uint32)
    lib().oracle_advection_collision(particle_ptr(particles), particle_ptr(out), ctypes.byref(params),
                                     _vp(scene.face_normals), _vp(scene.vertices), _vp(scene.indices),
                                     scene.face_count, max_iters, _vp(iters))
    return out, iters


class StepResult:
    pass


def step(particles, params, terms, scene, taps=True):
    """One sub-step. `params` grid block is updated in place. Returns StepResult with
    .particles (sorted output AoS) and, when taps=True, every per-stage observation."""
    n = particles.size
    out = np.empty_like(particles)
    res = StepResult()
    t = None
    if taps:
        cap = 1 << 24
        res.keys = np.empty(n, dtype=np.uint32)
        res.permutation = np.empty(n, dtype=np.uint32)
        res.cell_table = np.empty(cap, dtype=np.uint32)
        res.candidate_count = np.empty(n, dtype=np.uint32)
        res.support_count = np.empty(n, dtype=np.uint32)
        res.density = np.empty(n, dtype=np.float32)
        res.pressure = np.empty(n, dtype=np.float32)
        res.acceleration = np.empty((n, 3), dtype=np.float32)
        res.collision_iters = np.empty(n, dtype=np.uint32)
        t = _Taps()
        t.keys_input_order = res.keys.ctypes.data_as(_u32p)
        t.permutation = res.permutation.ctypes.data_as(_u32p)
        t.cell_table = res.cell_table.ctypes.data_as(_u32p)
        t.cell_table_capacity = cap
        t.candidate_count = res.candidate_count.ctypes.data_as(_u32p)
        t.support_count = res.support_count.ctypes.data_as(_u32p)
        t.density = res.density.ctypes.data_as(_f32p)
        t.pressure = res.pressure.ctypes.data_as(_f32p)
        t.acceleration = res.acceleration.ctypes.data_as(_f32p)
        t.collision_iters = res.collision_iters.ctypes.data_as(_u32p)
    rc = lib().oracle_step(particle_ptr(particles), particle_ptr(out), ctypes.byref(params),
                           ctypes.byref(terms), _vp(scene.face_normals), _vp(scene.vertices),
                           _vp(scene.indices), scene.face_count,
                           ctypes.byref(t) if t is not None else None)
    if rc:
        raise ValueError("oracle_step failed rc=%d" % rc)
    if taps:
        res.cell_table = res.cell_table[: params.grid_cell_count].copy()
    res.particles = out
    return res


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


class Scene:
    """Triangle scene in the reference's three-array form (libclsph/scene.h:7-15)."""

    def __init__(self, vertices, indices, normals=None):
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1)
        self.indices = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        self.face_count = self.indices.size // 3
        self.face_normals = (face_normals(self.vertices, self.indices) if normals is None
                             else np.ascontiguousarray(normals, dtype=np.float32).reshape(-1))


def load_obj(path):
    """Minimal Wavefront reader for the test harness: `v` and `f` records, fan
    triangulation, one shape. Vertices are kept in file order (the reference's tinyobj
    re-indexes them by first use; the triangles, their order and their corner order are
    the same, which is all the collision code observes)."""
    verts, faces = [], []
    with open(path) as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                verts.append([float(tok[1]), float(tok[2]), float(tok[3])])
            elif tok[0] == "f":
                idx = []
                for t in tok[1:]:
                    k = int(t.split("/")[0])
                    idx.append(k - 1 if k > 0 else len(verts) + k)
                for j in range(1, len(idx) - 1):
                    faces.append([idx[0], idx[j], idx[j + 1]])
    return Scene(np.array(verts, dtype=np.float32), np.array(faces, dtype=np.uint32))
