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


def elem_err(a, b, atol_frac=1e-3):
    """Per-element relative error, max over elements of |a-b| / (|b| + floor). The floor is atol_frac of the
    field's RMS magnitude: an element much smaller than the field's typical size (a velocity component that
    happens to vanish, a pressure at its zero crossing) is held to an absolute bound instead of an empty
    relative one. This is what "1e-4 relative" means particle by particle; rel_err only bounds the error
    against the field's largest value."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    rms = float(np.sqrt(np.mean(b * b))) if b.size else 0.0
    floor = atol_frac * rms if rms > 0 else 1.0
    return float((np.abs(a - b) / (np.abs(b) + floor)).max()) if b.size else 0.0


def assert_close_elementwise(got, want, tol=1e-4, fields=("position", "density"), what=""):
    for f in fields:
        g = got[f][:, :3] if got[f].ndim == 2 else got[f]
        w = want[f][:, :3] if want[f].ndim == 2 else want[f]
        e = elem_err(g, w)
        assert e <= tol, "%s %s: per-element relative error %.3e > %.1e" % (what, f, e, tol)


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


def edge_state(kind, p, vol, seed=3):
    """Small states built to sit on the decisions of the path rather than around them.

    coincident   a quarter of the particles share their position exactly with another one:
                 spiky_gradient's |r| < 1e-7 branch (erratum E3), r = 0 in every kernel
    on_support   pairs placed at distance h (1 +- a few ulp) along an axis and along a diagonal:
                 the window test floor(r/h) < 1 at its threshold
    on_cells     positions snapped to multiples of the cell side and of h: cell / sub-cell coordinates
                 exactly at their boundaries, right after the AABB padding
    one_cell     the whole fluid inside a single grid cell (a ball of radius 0.6 h): one huge cell, every
                 list overflows
    sparse       spacing of 3 h: every particle alone in its support
    """
    rng = np.random.default_rng(seed)
    s = workloads.jittered_state(p, vol, seed=seed)
    n = s.size
    h = np.float32(p.h)
    if kind == "coincident":
        src = rng.integers(0, n, n // 4)
        dst = rng.permutation(n)[: n // 4]
        s["position"][dst] = s["position"][src]
    elif kind == "on_support":
        half = n // 2
        base = s["position"][:half, :3].copy()
        ulps = rng.integers(-3, 4, half)
        d = np.nextafter(h, np.float32(np.inf)) if False else h
        dist = (d * (np.float32(1) + ulps.astype(np.float32) * np.float32(2.0 ** -23))).astype(np.float32)
        direction = np.zeros((half, 3), dtype=np.float32)
        axis = rng.integers(0, 4, half)
        for a in range(3):
            direction[axis == a, a] = 1
        direction[axis == 3] = np.float32(1 / np.sqrt(3))
        s["position"][half:2 * half, :3] = base + direction * dist[:, None]
    elif kind == "on_cells":
        cell = np.float32(2) * h
        q = np.round(s["position"][:, :3] / h).astype(np.float32)
        s["position"][:, :3] = np.where(rng.random((n, 3)) < 0.5, q * h, np.round(q / 2) * cell)
    elif kind == "one_cell":
        d = rng.normal(0, 1, (n, 3))
        d /= np.linalg.norm(d, axis=1)[:, None]
        s["position"][:, :3] = (d * (rng.random((n, 1)) ** (1 / 3)) * 0.6 * float(h)).astype(np.float32)
        s["position"][:, 1] += np.float32(0.5)
    elif kind == "sparse":
        per_side = int(np.ceil(n ** (1 / 3)))
        i = np.arange(n)
        s["position"][:, 0] = (i % per_side) * 3 * h
        s["position"][:, 1] = ((i // per_side) % per_side) * 3 * h
        s["position"][:, 2] = (i // (per_side * per_side)) * 3 * h
    else:
        raise ValueError(kind)
    return s


EDGE_KINDS = ("coincident", "on_support", "on_cells", "one_cell", "sparse")


_DEVELOPED = {}


def developed_state(fluid="water", n=4096, substeps=200, drop=0.05):
    """State S2 of SURVEY 8(d) in small: the lattice released `drop` metres above the floor of box.obj and
    advanced `substeps` sub-steps by the ORACLE, so that the fluid has hit the floor, spread against the
    walls and formed a free surface (collisions, crowded and thin regions, real velocity field). Cached."""
    from oracle import oracle as O
    key = (fluid, n, substeps, drop)
    if key not in _DEVELOPED:
        p, terms, vol = config(fluid, n)
        scene = O.load_obj(os.path.join(ROOT, "scenes", "box.obj"))
        s = state_s0(p, vol)
        s["position"][:, 1] += np.float32(-2.0 + drop)
        po = p.copy()
        for _ in range(substeps):
            s = O.step(s, po, terms, scene, taps=False).particles
        _DEVELOPED[key] = (p, terms, scene, s)
    p, terms, scene, s = _DEVELOPED[key]
    return p.copy(), terms, scene, s.copy()
