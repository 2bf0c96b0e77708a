"""Synthetic inputs for the BASELINE configs: derived constants, lattice and jittered states.

Host-side mirror (numpy fp32 + glibc cbrtf/pow through ctypes) of what the reference's
load_settings (libclsph/sph_simulation.cpp:490-505) and init_particles (:48-94) compute; the
C++ host library (libclsph_b200/host/) does the same natively for the drop-in API. No SPH step
arithmetic happens here.
"""
import ctypes
import ctypes.util
import math
import os

import numpy as np

from .abi import PARTICLE, PrecomputedKernelValues, raw_parameters

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.cbrtf.restype = ctypes.c_float
_libm.cbrtf.argtypes = [ctypes.c_float]
_libm.pow.restype = ctypes.c_double
_libm.pow.argtypes = [ctypes.c_double, ctypes.c_double]

f32 = np.float32
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# name -> (fluid, particles_count, particle_mass, scene) as restated in BASELINE.md section 4
CONFIGS = {
    "config1_box_100k": ("water", 102400, 0.05, "box.obj"),
    "config2_dambreak_1m": ("water", 1048576, 0.005, "box.obj"),
    "config3_mucus_labyrinth_4m": ("mucus", 4194304, 0.05 * 32000 / 4194304, "labyrinth.obj"),
    "config4_river_16m": ("water", 16777216, 0.05, "river.obj"),
    # config 5: uniform-block scaling sweep (water, plane.obj, state S1), N = 2^20 .. 2^26
    "sweep_1m": ("water", 1 << 20, 0.05, "plane.obj"),
    "sweep_4m": ("water", 1 << 22, 0.05, "plane.obj"),
    "sweep_16m": ("water", 1 << 24, 0.05, "plane.obj"),
    "sweep_64m": ("water", 1 << 26, 0.05, "plane.obj"),
}


def derive_constants(params, n_influence):
    """Fills h, time_delta, max_velocity, total_mass; returns (terms, initial_volume)."""
    p = params
    p.total_mass = float(f32(p.particles_count) * f32(p.particle_mass))
    volume = f32(p.total_mass) / f32(p.fluid_density)
    per_particle = f32(volume) / f32(p.particles_count)
    numer = f32(3.0) * (f32(n_influence) * per_particle)
    p.h = _libm.cbrtf(f32(float(numer) / (float(f32(4.0)) * math.pi)))
    p.time_delta = float(f32(1.0) / f32(p.target_fps))
    p.max_velocity = float(f32(0.8) * f32(p.h) / f32(p.time_delta))
    h9 = _libm.pow(float(f32(p.h)), 9.0)
    h6 = _libm.pow(float(f32(p.h)), 6.0)
    t = PrecomputedKernelValues()
    t.poly_6 = 315.0 / (64.0 * math.pi * h9)
    t.poly_6_gradient = -945.0 / (32.0 * math.pi * h9)
    t.poly_6_laplacian = -945.0 / (32.0 * math.pi * h9)
    t.spiky = -45.0 / (math.pi * h6)
    t.viscosity = 45.0 / (math.pi * h6)
    return t, float(volume)


def lattice_geometry(params, initial_volume):
    per_side = int(math.ceil(_libm.cbrtf(f32(params.particles_count))))
    side = f32(_libm.cbrtf(f32(initial_volume)))
    spacing = side / f32(per_side)
    return per_side, side, spacing


def lattice_positions(params, initial_volume, index):
    """Lattice positions (reference init_particles, sph_simulation.cpp:71-92) of the given global indices."""
    per_side, side, spacing = lattice_geometry(params, initial_volume)
    i = np.asarray(index, dtype=np.uint32)
    ups = np.uint32(per_side)
    half = side / f32(2.0)
    pos = np.zeros((i.size, 4), dtype=f32)
    pos[:, 0] = (i % ups).astype(f32) * spacing - half
    pos[:, 1] = ((i // ups) % ups).astype(f32) * spacing
    pos[:, 2] = (i // (ups * ups)).astype(f32) * spacing - half
    return pos


def lattice_state(params, initial_volume):
    """State S0: the reference's cubic lattice, everything else zero."""
    buf = np.zeros(params.particles_count, dtype=PARTICLE)
    buf["position"] = lattice_positions(params, initial_volume, np.arange(params.particles_count, dtype=np.uint32))
    return buf


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform01(seed, counters):
    """u = (splitmix64(seed xor counter) >> 40) / 2^24, a counter-based generator."""
    with np.errstate(over="ignore"):
        z = _splitmix64(np.uint64(seed) ^ counters.astype(np.uint64))
    return (z >> np.uint64(40)).astype(np.float64) / float(1 << 24)


def jittered_state(params, initial_volume, seed=20261017, position_jitter=0.25, velocity_jitter=0.5, index=None):
    """State S1: S0 plus uniform offsets of +-position_jitter*spacing and +-velocity_jitter m/s.

    `index` (global particle indices) restricts the result to a subset; the jitter of a particle
    depends only on its global index, so subsets generated on different ranks agree."""
    if index is None:
        index = np.arange(params.particles_count, dtype=np.uint32)
    index = np.asarray(index, dtype=np.uint32)
    buf = np.zeros(index.size, dtype=PARTICLE)
    buf["position"] = lattice_positions(params, initial_volume, index)
    _, _, spacing = lattice_geometry(params, initial_volume)
    i = index.astype(np.uint64)
    for c in range(3):
        u = uniform01(seed, np.uint64(6) * i + np.uint64(c))
        buf["position"][:, c] += ((u - 0.5) * 2.0 * position_jitter * float(spacing)).astype(f32)
        w = uniform01(seed, np.uint64(6) * i + np.uint64(3 + c))
        vel = ((w - 0.5) * 2.0 * velocity_jitter).astype(f32)
        buf["velocity"][:, c] = vel
        buf["intermediate_velocity"][:, c] = vel
    return buf


def slab_indices(params, initial_volume, rank, world):
    """Global indices of the lattice columns that slab `rank` of `world` starts with, and the x
    planes (world+1 values, -inf/+inf at the ends) halfway between neighbouring slabs' columns."""
    per_side, side, spacing = lattice_geometry(params, initial_volume)
    n = params.particles_count
    cols = [(per_side * r) // world for r in range(world + 1)]
    planes = np.full(world + 1, np.inf, dtype=f32)
    planes[0] = -np.inf
    for r in range(1, world):
        planes[r] = f32(cols[r]) * spacing - side / f32(2.0) - spacing / f32(2.0)
    x = np.arange(cols[rank], cols[rank + 1], dtype=np.uint64)
    rows = np.arange((n + per_side - 1) // per_side, dtype=np.uint64)  # (y, z) pairs, row-major
    idx = (rows[:, None] * np.uint64(per_side) + x[None, :]).reshape(-1)
    idx = idx[idx < n]
    return idx.astype(np.uint32), planes


def make_config(name=None, fluid="water", particles_count=32000, particle_mass=0.05, **overrides):
    """(params, terms, initial_volume, scene_file) for a named BASELINE config or explicit values."""
    scene = overrides.pop("scene", "box.obj")
    if name is not None:
        fluid, particles_count, particle_mass, scene = CONFIGS[name]
    p, n_infl = raw_parameters(fluid, particles_count=particles_count, particle_mass=particle_mass, **overrides)
    terms, volume = derive_constants(p, n_infl)
    return p, terms, volume, scene


def load_obj_triangles(path):
    """(vertices[3V] f32, indices[3F] u32) of a Wavefront file: `v`/`f` records, fan triangulation.
    Face normals are NOT computed here; callers pass their own (scene.cpp:36-64 computes them on
    the host in the reference, libclsph_b200/host/ does in the product)."""
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
    return np.asarray(verts, dtype=f32).reshape(-1), np.asarray(faces, dtype=np.uint32).reshape(-1)


def face_normals(vertices, indices):
    """Unit face normals as libclsph/scene.cpp:36-64 computes them (fp32 cross product, length via
    a double sqrt of the fp32 sum, fp32 divide)."""
    v = np.asarray(vertices, dtype=f32).reshape(-1, 3)
    t = np.asarray(indices, dtype=np.uint32).reshape(-1, 3)
    a, b, c = v[t[:, 0]], v[t[:, 1]], v[t[:, 2]]
    u, w = b - a, c - a
    nx = u[:, 1] * w[:, 2] - u[:, 2] * w[:, 1]
    ny = u[:, 2] * w[:, 0] - u[:, 0] * w[:, 2]
    nz = u[:, 0] * w[:, 1] - u[:, 1] * w[:, 0]
    length = np.sqrt((nx * nx + ny * ny + nz * nz).astype(np.float64)).astype(f32)
    return np.stack([nx / length, ny / length, nz / length], axis=1).astype(f32).reshape(-1)


def scene_arrays(scene_file):
    """(face_normals, vertices, indices) for scenes/<scene_file> of this repository."""
    vertices, indices = load_obj_triangles(os.path.join(ROOT, "scenes", scene_file))
    return face_normals(vertices, indices), vertices, indices
