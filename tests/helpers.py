"""Shared helpers of the test-suite: workloads and comparison rules."""
import ctypes
import os

import numpy as np

from libclsph_b200 import abi, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS_XYZ = ("position", "velocity", "intermediate_velocity")


def struct_bytes(s):
    return bytes(ctypes.string_at(ctypes.addressof(s), ctypes.sizeof(s)))


def config(fluid="water", n=4096, mass=0.05, **kw):
    """(params, terms, volume) through the product's host-side derivation."""
    p, terms, vol, _ = workloads.make_config(fluid=fluid, particles_count=n, particle_mass=mass, **kw)
    return p, terms, vol


def state_s0(p, vol):
    return workloads.lattice_state(p, vol)


def state_s1(p, vol, **kw):
    return workloads.jittered_state(p, vol, **kw)


def drop_state(p, vol, scene_floor_y, speed=2.5, seed=7, slab=0.003):
    """All particles squeezed into a slab `slab` metres thick just above a floor and moving down
    fast enough to cross it within one sub-step: nearly every particle goes through
    detect/respond/time-splitting, and the crowded cells (hundreds of particles each) exercise the
    candidate-list chunking and neighbour-list flushing of the CUDA neighbour kernels."""
    s = workloads.jittered_state(p, vol, seed=seed)
    u = workloads.uniform01(seed + 1, np.arange(s.size, dtype=np.uint64))
    s["position"][:, 1] = (scene_floor_y + 0.0005 + slab * u).astype(np.float32)
    s["intermediate_velocity"][:, 1] = np.float32(-speed)
    s["velocity"][:, 1] = np.float32(-speed)
    return s


def rel_err(a, b):
    """max |a-b| / max |b| : error relative to the field's magnitude (north_star: 1e-4 in fp32)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.abs(b).max()
    return float(np.abs(a - b).max() / (scale if scale > 0 else 1.0))


def assert_close_fields(got, want, tol=1e-4, fields=FIELDS_XYZ + ("density", "pressure"), what=""):
    for f in fields:
        g = got[f][:, :3] if got[f].ndim == 2 else got[f]
        w = want[f][:, :3] if want[f].ndim == 2 else want[f]
        e = rel_err(g, w)
        assert e <= tol, "%s %s: relative error %.3e > %.1e" % (what, f, e, tol)


def surface_state(scene, n, reach, speed, seed):
    """Particles scattered within `reach` of random points on the scene's triangles, moving in random
    directions at up to `speed`: many of them hit a face (or several) during one sub-step."""
    rng = np.random.default_rng(seed)
    v = scene.vertices.reshape(-1, 3)
    t = scene.indices.reshape(-1, 3)
    ok = np.isfinite(scene.face_normals.reshape(-1, 3)).all(axis=1)
    f = rng.choice(np.nonzero(ok)[0], size=n)
    a, b = rng.random(n), rng.random(n)
    flip = a + b > 1
    a[flip], b[flip] = 1 - a[flip], 1 - b[flip]
    on = v[t[f, 0]] + a[:, None] * (v[t[f, 1]] - v[t[f, 0]]) + b[:, None] * (v[t[f, 2]] - v[t[f, 0]])
    s = np.zeros(n, dtype=abi.PARTICLE)
    s["position"][:, :3] = (on + rng.normal(0, reach, size=(n, 3))).astype(np.float32)
    d = rng.normal(0, 1, size=(n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    s["intermediate_velocity"][:, :3] = (d * rng.uniform(0, speed, size=(n, 1))).astype(np.float32)
    s["velocity"] = s["intermediate_velocity"]
    s["acceleration"][:, :3] = rng.normal(0, 10, size=(n, 3)).astype(np.float32)
    return s
