"""CPU tests: known answers for the oracle, derived by hand from the reference source.

The reference ships no tests or golden vectors; these closed forms, the
stable-sort / lower_bound cross-checks and tests/test_oracle_vs_ref.py (the reference's own
sources compiled behind a shim) are what pins the oracle.
"""
import math

import numpy as np
import pytest

from libclsph_b200 import abi
from oracle import oracle as O
from tests import helpers as H


def test_morton_known_answers():
    # common/util.h:41-62
    assert O.morton_encode(1, 0, 0) == 1
    assert O.morton_encode(0, 1, 0) == 2
    assert O.morton_encode(0, 0, 1) == 4
    assert O.morton_encode(2, 3, 5) == 286
    assert O.morton_encode(17, 17, 17) == 28679
    assert O.morton_encode(1023, 1023, 1023) == 0x3FFFFFFF
    rng = np.random.default_rng(0)
    for x, y, z in rng.integers(0, 1024, size=(200, 3)):
        assert O.morton_decode(O.morton_encode(int(x), int(y), int(z))) == (x, y, z)


def test_morton_is_monotone_in_each_coordinate():
    """Why every key is < grid_cell_count = morton(gx, gy, gz)."""
    rng = np.random.default_rng(1)
    for _ in range(300):
        a = rng.integers(0, 1023, size=3)
        b = a + rng.integers(0, 3, size=3)
        b = np.minimum(b, 1023)
        if (a == b).all():
            continue
        assert O.morton_encode(*map(int, a)) < O.morton_encode(*map(int, b))


def test_derived_constants_water_default():
    p, n_infl = abi.raw_parameters("water")
    terms, vol = O.derive_constants(p, n_infl)
    assert p.h == pytest.approx(0.0620705, rel=1e-6)
    assert p.max_velocity == pytest.approx(2.97938, rel=1e-5)
    assert p.time_delta == np.float32(1.0) / np.float32(60.0)
    assert terms.poly_6 == pytest.approx(315.0 / (64 * math.pi * p.h ** 9), rel=1e-6)
    assert terms.spiky == pytest.approx(-45.0 / (math.pi * p.h ** 6), rel=1e-6)
    assert terms.viscosity == -terms.spiky
    assert terms.poly_6_gradient == terms.poly_6_laplacian


def test_lattice_geometry_config1():
    p, n_infl = abi.raw_parameters("water", particles_count=102400)
    terms, vol = O.derive_constants(p, n_infl)
    s = O.init_particles(p, vol)
    assert vol == pytest.approx(5.129, rel=1e-3)
    xs = np.unique(s["position"][:, 0])
    assert xs.size == 47  # ceil(cbrt(102400))
    assert s["position"][:, 1].min() == 0.0
    assert O.bounds_and_grid(s, p) == 0
    assert (p.grid_size_x, p.grid_size_y, p.grid_size_z) == (17, 17, 17)
    assert p.grid_cell_count == 28679


def _two_particles(p, d):
    """128 particles: two at distance d, the rest far away and far from each other's cells."""
    s = np.zeros(128, dtype=abi.PARTICLE)
    s["position"][:, 0] = 1.0 + 0.5 * np.arange(128)  # 4 cells apart, 64 m in total (< 1024 cells)
    s["position"][0, :3] = (0.0, 0.0, 0.0)
    s["position"][1, :3] = (d, 0.0, 0.0)
    return s


def test_isolated_particle_density_and_pressure():
    # rho = m * C6 * h^6 = 315 m / (64 pi h^3);  p = K ((rho/rho0)^7 - 1)
    p, n_infl = abi.raw_parameters("water", particles_count=128)
    terms, vol = O.derive_constants(p, n_infl)
    s = _two_particles(p, 50 * p.h)
    scene = O.Scene(np.zeros(0, np.float32), np.zeros(0, np.uint32), np.zeros(0, np.float32))
    r = O.step(s, p, terms, scene)
    rho = 315.0 * p.particle_mass / (64 * math.pi * p.h ** 3)
    assert np.allclose(r.density, rho, rtol=2e-6)
    assert np.allclose(r.pressure, p.K * ((rho / p.fluid_density) ** 7 - 1), rtol=2e-5)
    assert (r.support_count == 1).all()


@pytest.mark.parametrize("frac", [0.25, 0.6, 0.95])
def test_two_particles_closed_form(frac):
    p, n_infl = abi.raw_parameters("water", particles_count=128)
    p.constant_acceleration.s[1] = 0.0
    terms, vol = O.derive_constants(p, n_infl)
    h, m = p.h, p.particle_mass
    d = frac * h
    s = _two_particles(p, d)
    s["velocity"][1, 1] = 0.3  # relative velocity for the viscosity term
    scene = O.Scene(np.zeros(0, np.float32), np.zeros(0, np.uint32), np.zeros(0, np.float32))
    r = O.step(s, p, terms, scene)
    i0 = int(np.where(r.permutation == 0)[0][0])
    i1 = int(np.where(r.permutation == 1)[0][0])
    c6 = 315.0 / (64 * math.pi * h ** 9)
    rho = m * c6 * (h ** 6 + (h * h - d * d) ** 3)
    assert r.density[i0] == pytest.approx(rho, rel=5e-6)
    assert r.density[i1] == pytest.approx(rho, rel=5e-6)
    assert r.support_count[i0] == 2
    prs = p.K * ((rho / p.fluid_density) ** 7 - 1)
    # pressure: a_x = -(2 p / rho^2) m * Cs * (d/|d|) (h-|d|)^2 with d = x0 - x1 = (-d, 0, 0)
    cs, cv = -45.0 / (math.pi * h ** 6), 45.0 / (math.pi * h ** 6)
    ax_pressure = -(2 * prs / rho ** 2) * m * cs * (-1.0) * (h - d) ** 2
    # surface tension is off unless |n| > threshold; check and add if needed
    cg = -945.0 / (32 * math.pi * h ** 9)
    n_x = (m / rho) * cg * (-d) * (h * h - d * d) ** 2
    lap = (m / rho) * cg * ((h * h - d * d) * (3 * h * h - 7 * d * d) + h * h * 3 * h * h)
    ax_tension = (-p.surface_tension * lap * n_x / abs(n_x)) / rho if abs(n_x) > p.surface_tension_threshold else 0.0
    assert r.acceleration[i0, 0] == pytest.approx(ax_pressure + ax_tension, rel=2e-4)
    # equal and opposite
    assert r.acceleration[i1, 0] == pytest.approx(-r.acceleration[i0, 0], rel=1e-5)
    # viscosity: a_y(0) = mu * (v1 - v0) * (m / rho) * Cv (h - d) / rho
    ay_visc = p.dynamic_viscosity * 0.3 * (m / rho) * cv * (h - d) / rho
    assert r.acceleration[i0, 1] == pytest.approx(ay_visc, rel=2e-4)
    assert r.acceleration[i1, 1] == pytest.approx(-ay_visc, rel=2e-4)


def test_lattice_interior_density_close_to_rest_density():
    """The poly6 kernel integrates to one: an interior lattice particle sees about rho0."""
    p, n_infl = abi.raw_parameters("water", particles_count=32768)
    terms, vol = O.derive_constants(p, n_infl)
    s = O.init_particles(p, vol)
    scene = O.Scene(np.zeros(0, np.float32), np.zeros(0, np.uint32), np.zeros(0, np.float32))
    r = O.step(s, p, terms, scene)
    interior = r.support_count == r.support_count.max()
    assert abs(np.median(r.density[interior]) / p.fluid_density - 1.0) < 0.08


def test_particle_falling_onto_plane_closed_form(plane_scene):
    """plane.obj is the floor y = -1. A particle 1 mm above it moving down at 1 m/s:
    hit point on the plane, 1 mm push-back along the travel-oriented normal, normal velocity
    removed (restitution 0), remaining time spent sliding (collisions.cl:91-128)."""
    p, n_infl = abi.raw_parameters("water", particles_count=128)
    p.constant_acceleration.s[1] = 0.0
    terms, vol = O.derive_constants(p, n_infl)
    s = np.zeros(128, dtype=abi.PARTICLE)
    s["position"][:, 0] = np.linspace(-0.9, 0.9, 128)
    s["position"][:, 1] = 5.0
    s["position"][0, :3] = (0.1, -0.999, 0.2)
    s["intermediate_velocity"][0, :3] = (0.5, -1.0, 0.0)
    out, iters = O.advection_collision(s, p, plane_scene)
    assert iters[0] == 2 and (iters[1:] == 1).all()
    # travel direction is downward, so the oriented normal is (0,-1,0): push-back moves it UP by 1 mm
    assert out["position"][0, 1] == pytest.approx(-1.0 + 0.001, abs=2e-6)
    assert out["intermediate_velocity"][0, 1] == pytest.approx(0.0, abs=1e-6)
    assert out["intermediate_velocity"][0, 0] == pytest.approx(0.5, rel=1e-6)
    assert out["velocity"][0, 1] == pytest.approx(-0.5, rel=1e-5)  # (iv_old + iv_new) / 2
    # time bookkeeping of collisions.cl:123-125: the hit is at 0.6 of the segment, I = (0.1005, -1, 0.2);
    # after the push-back the particle is back at its starting height, so the "distance covered" is the
    # 0.0005 tangential part only and t_used = dt * 0.0005 / |dt * v|; the rest of dt is spent sliding.
    dt = p.time_delta * p.simulation_scale
    used = dt * 0.0005 / (dt * math.sqrt(1.25))
    assert out["position"][0, 0] == pytest.approx(0.1005 + 0.5 * (dt - used), rel=1e-5)
    assert not out["acceleration"].any()


def test_later_face_wins_distance_ties(plane_scene):
    """The two triangles of plane.obj share the diagonal x = z; a hit exactly on it belongs to
    both with equal distance, and `>` (collisions.cl:77-80) lets the later face overwrite.
    Observable only through the normal, which is the same here, so check it still collides once."""
    p, n_infl = abi.raw_parameters("water", particles_count=128)
    terms, vol = O.derive_constants(p, n_infl)
    s = np.zeros(128, dtype=abi.PARTICLE)
    s["position"][:, 1] = 5.0
    s["position"][0, :3] = (0.25, -0.9995, 0.25)
    s["intermediate_velocity"][0, 1] = -2.0
    out, iters = O.advection_collision(s, p, plane_scene)
    assert iters[0] == 2
    assert out["position"][0, 1] > -1.0


@pytest.mark.parametrize("n", [128, 129, 5000, 32768])
def test_sort_is_stable_and_table_is_lower_bound(n):
    rng = np.random.default_rng(n)
    s = np.zeros(n, dtype=abi.PARTICLE)
    s["grid_index"] = rng.integers(0, 300, size=n).astype(np.uint32) * rng.integers(1, 70000, size=n).astype(np.uint32)
    s["density"] = np.arange(n, dtype=np.float32)  # identity marker
    out, perm = O.sort_particles(s)
    want = np.argsort(s["grid_index"], kind="stable").astype(np.uint32)
    assert np.array_equal(perm, want)
    assert np.array_equal(out["density"], s["density"][want])
    count = int(s["grid_index"].max()) + 5
    if count < 2_000_000:
        table = O.cell_table(out, count)
        assert np.array_equal(table, np.searchsorted(out["grid_index"], np.arange(count), side="left"))


def test_sort_rejects_fewer_than_128_particles():
    s = np.zeros(100, dtype=abi.PARTICLE)
    with pytest.raises(ValueError):
        O.sort_particles(s)


def test_candidate_count_is_a_function_of_the_table():
    p, terms, vol = H.config("water", 4096)
    s = H.state_s1(p, vol)
    scene = O.Scene(np.zeros(0, np.float32), np.zeros(0, np.uint32), np.zeros(0, np.float32))
    r = O.step(s, p, terms, scene)
    table = np.append(r.cell_table, s.size).astype(np.int64)
    keys = r.particles["grid_index"]
    for i in (0, 17, 2048, 4095):
        cx, cy, cz = O.morton_decode(int(keys[i]))
        total = 0
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    c = O.morton_encode(cx + dx, cy + dy, cz + dz)
                    total += table[c + 1] - table[c]
        assert total == r.candidate_count[i]
    assert (r.support_count <= r.candidate_count).all() and (r.support_count >= 1).all()


def test_grid_overflow_is_reported():
    p, terms, vol = H.config("water", 256)
    s = H.state_s0(p, vol)
    s["position"][:128, 0] += np.float32(1100 * 2 * p.h)
    assert O.bounds_and_grid(s, p) == 1
