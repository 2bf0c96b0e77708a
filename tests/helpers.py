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
