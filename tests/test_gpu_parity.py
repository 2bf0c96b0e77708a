"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Bar (BASELINE.json north_star): cell keys, sort permutation, cell table and neighbour-candidate
counts bit-exact; support counts bit-exact under the oracle's rounding contract; single-step
densities, pressures, accelerations, positions and velocities within 1e-4 relative in fp32.
"""
import os

import numpy as np
import pytest

from libclsph_b200 import abi, capi, workloads
from oracle import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

TOL = 1e-4  # fp32 relative tolerance stated by the north star


def make_ctx(n, scene, params, terms, debug=True, cell_table_capacity=0, options=None):
    ctx = capi.Context(n, cell_table_capacity=cell_table_capacity)
    for k, v in (options or {}).items():
        ctx.set_option(k, v)
    ctx.set_scene(scene.face_normals, scene.vertices, scene.indices)
    ctx.set_parameters(params, terms)
    ctx.set_debug(debug)
    return ctx


def gpu_step_with_taps(state, params, terms, scene, **kw):
    ctx = make_ctx(state.size, scene, params, terms, **kw)
    ctx.upload(state)
    ctx.step(1)
    ctx.synchronize()
    taps = dict(
        keys=ctx.fetch(capi.TAP_KEYS_INPUT), skeys=ctx.fetch(capi.TAP_SORTED_KEYS),
        permutation=ctx.fetch(capi.TAP_PERMUTATION), cell_table=ctx.fetch(capi.TAP_CELL_TABLE),
        candidate_count=ctx.fetch(capi.TAP_CANDIDATE_COUNT), support_count=ctx.fetch(capi.TAP_SUPPORT_COUNT),
        density=ctx.fetch(capi.TAP_DENSITY), pressure=ctx.fetch(capi.TAP_PRESSURE),
        acceleration=ctx.fetch(capi.TAP_ACCELERATION), collision_iters=ctx.fetch(capi.TAP_COLLISION_ITERS))
    out = ctx.download()
    p_after = ctx.parameters()
    ctx.close()
    return out, taps, p_after


def check_against_oracle(state, params, terms, scene, what, **kw):
    po = params.copy()
    want = O.step(state, po, terms, scene)
    got, taps, p_after = gpu_step_with_taps(state, params, terms, scene, **kw)
    # --- the grid block the callbacks observe (sph_simulation.cpp:229-252): bit-exact
    assert H.struct_bytes(p_after) == H.struct_bytes(po), what + ": grid block of simulation_parameters differs"
    # --- integer observables: bit-exact
    assert np.array_equal(taps["keys"], want.keys), what + ": cell keys"
    assert np.array_equal(taps["permutation"], want.permutation), what + ": sort permutation"
    assert np.array_equal(taps["skeys"], want.keys[want.permutation]), what + ": sorted keys"
    assert np.array_equal(taps["cell_table"], want.cell_table), what + ": cell table"
    assert np.array_equal(taps["candidate_count"], want.candidate_count), what + ": candidate counts"
    assert np.array_equal(taps["support_count"], want.support_count), what + ": support counts"
    assert np.array_equal(taps["collision_iters"], want.collision_iters), what + ": collision loop trips"
    assert np.array_equal(got["grid_index"], want.particles["grid_index"]), what + ": grid_index in the output"
    # --- fp32 observables: 1e-4 relative
    assert H.rel_err(taps["density"], want.density) <= TOL, what + ": density"
    assert H.rel_err(taps["pressure"], want.pressure) <= TOL, what + ": pressure"
    assert H.rel_err(taps["acceleration"], want.acceleration) <= TOL, what + ": acceleration"
    H.assert_close_fields(got, want.particles, tol=TOL, what=what)
    # ... and particle by particle (relative to each element, with a small absolute floor), not only
    # against the field's largest value
    assert H.elem_err(taps["density"], want.density) <= TOL, what + ": density, per element"
    H.assert_close_elementwise(got, want.particles, tol=TOL, fields=("position", "density"), what=what)
    assert not got["acceleration"].any(), what + ": the step must export acceleration = 0 (sph.cl:97-99)"
    return got, taps, want


def check_resident_steps_against_oracle(state, params, terms, scene, steps, what, options=None, tol=TOL):
    """Device-resident sub-steps (no re-upload), each checked against the oracle stepped from the
    PREVIOUS downloaded state: the permutation tap and the output order must be exactly the
    reference's at every step, which is what exercises the carried reference rank of the sub-cell
    order (the arrays in HBM are in another order there)."""
    ctx = make_ctx(state.size, scene, params, terms, options=options)
    ctx.upload(state)
    prev = state
    for k in range(steps):
        ctx.step(1)
        got = ctx.download()
        perm = ctx.fetch(capi.TAP_PERMUTATION)
        keys_in = ctx.fetch(capi.TAP_KEYS_INPUT)
        supp = ctx.fetch(capi.TAP_SUPPORT_COUNT)
        want = O.step(prev, params.copy(), terms, scene)
        tag = "%s, sub-step %d" % (what, k)
        assert np.array_equal(keys_in, want.keys), tag + ": cell keys in pre-step order"
        assert np.array_equal(perm, want.permutation), tag + ": sort permutation"
        assert np.array_equal(got["grid_index"], want.particles["grid_index"]), tag + ": grid_index"
        assert np.array_equal(supp, want.support_count), tag + ": support counts"
        H.assert_close_fields(got, want.particles, tol=tol, what=tag)
        H.assert_close_elementwise(got, want.particles, tol=TOL, fields=("position", "density"), what=tag)
        prev = got
    ctx.close()


def check_counting_sort_against_radix(fluid, n, scene, steps=4):
    """Sorting by counting on the sub-cell table (default while the grid fits it) against the radix passes: the order
    inside a sub-cell is set by the reference rank either way, so every byte of the state and of the taps is the same
    after several resident sub-steps; clsph_sort_passes tells which sort ran."""
    p, terms, vol = H.config(fluid, n)
    s = H.state_s1(p, vol)
    s["intermediate_velocity"][:, 0] += (2.5 * np.sign(s["position"][:, 2])).astype(np.float32)  # shear: sub-cells change
    outs, passes = [], []
    for counting in (1, 0):
        ctx = make_ctx(s.size, scene, p, terms, debug=True, options=dict(count_sort=counting))
        ctx.upload(s)
        ctx.step(steps)
        passes.append(ctx.sort_passes())
        outs.append((ctx.download().tobytes(),) + tuple(ctx.fetch(t).tobytes() for t in (
            capi.TAP_SORTED_KEYS, capi.TAP_PERMUTATION, capi.TAP_CELL_TABLE, capi.TAP_SUPPORT_COUNT, capi.TAP_CANDIDATE_COUNT,
            capi.TAP_ACCELERATION)))
        ctx.close()
    assert passes[0] == 0 and passes[1] >= 1, passes
    assert outs[0] == outs[1]


@pytest.mark.parametrize("fluid,n", [("water", 200000), ("mucus", 60000)])
def test_counting_sort_is_bitwise_the_radix_sort(fluid, n, box_scene):
    check_counting_sort_against_radix(fluid, n, box_scene)


def check_counting_sort_on_a_sparse_table(n, scene, steps=3):
    """Two halves of the fluid 40 cells apart: many table words per particle, so the device's rule picks the radix passes
    (and zeroes the table row by row afterwards); count_sort = 2 forces the counting sort, whose scan then runs over
    many chunks with a ragged last one and whose table is zeroed as a range. Same bytes either way, over several sub-steps."""
    p, terms, vol = H.config("water", n)
    s = H.state_s1(p, vol)
    s["position"][: s.size // 2, 2] += np.float32(40 * 2 * p.h)
    s["intermediate_velocity"][:, 0] += (2.5 * np.sign(s["position"][:, 1])).astype(np.float32)
    outs, passes = [], []
    for mode in (2, 1, 0):
        ctx = make_ctx(s.size, scene, p, terms, debug=True, options=dict(count_sort=mode))
        ctx.upload(s)
        ctx.step(steps)
        passes.append(ctx.sort_passes())
        outs.append((ctx.download().tobytes(),) + tuple(ctx.fetch(t).tobytes() for t in (
            capi.TAP_SORTED_KEYS, capi.TAP_PERMUTATION, capi.TAP_CELL_TABLE, capi.TAP_SUPPORT_COUNT, capi.TAP_CANDIDATE_COUNT)))
        assert ctx.parameters().grid_cell_count * 9 > 6 * s.size
        ctx.close()
    assert passes[0] == 0 and passes[1] >= 1 and passes[2] >= 1, passes
    assert outs[0] == outs[1] == outs[2]


def test_counting_sort_on_a_table_much_larger_than_the_fluid(box_scene):
    check_counting_sort_on_a_sparse_table(100000, box_scene)


@pytest.mark.parametrize("n", [128, 1000, 4096, 32000])
def test_lattice_state_s0(n, box_scene):
    p, terms, vol = H.config("water", n)
    check_against_oracle(H.state_s0(p, vol), p, terms, box_scene, "S0 n=%d" % n)


@pytest.mark.parametrize("fluid,n", [("water", 4096), ("water", 102400), ("mucus", 20000)])
def test_jittered_state_s1(fluid, n, box_scene):
    p, terms, vol = H.config(fluid, n)
    check_against_oracle(H.state_s1(p, vol), p, terms, box_scene, "S1 %s n=%d" % (fluid, n))


def test_ragged_count_not_multiple_of_anything(box_scene):
    p, terms, vol = H.config("water", 12345)
    check_against_oracle(H.state_s1(p, vol, seed=3), p, terms, box_scene, "ragged n=12345")


def test_collision_heavy_step(plane_scene):
    """Most particles cross the floor of plane.obj (y = -1) during the sub-step."""
    p, terms, vol = H.config("mucus", 8192)
    s = H.drop_state(p, vol, scene_floor_y=-1.0)
    got, taps, want = check_against_oracle(s, p, terms, plane_scene, "drop onto plane")
    assert (want.collision_iters > 1).mean() > 0.2


def test_labyrinth_scene_many_faces():
    scene = O.load_obj(os.path.join(H.ROOT, "scenes", "labyrinth.obj"))
    p, terms, vol = H.config("mucus", 16384, mass=0.05 * 32000 / 4194304 * 256)
    s = H.state_s1(p, vol)
    check_against_oracle(s, p, terms, scene, "labyrinth")


def test_advection_collision_kernel_bit_exact(box_scene):
    """With identical inputs (acceleration included) the integrator is bit-identical to the
    oracle's restatement of kernels/sph.cl:64-112 + collisions.cl."""
    p, terms, vol = H.config("water", 20000)
    s = H.drop_state(p, vol, scene_floor_y=-2.0, speed=2.9, slab=0.02)
    rng = np.random.default_rng(11)
    s["acceleration"][:, :3] = rng.normal(0, 30, size=(s.size, 3)).astype(np.float32)
    s["intermediate_velocity"][:, 0] = rng.uniform(-2, 2, s.size).astype(np.float32)
    want, iters = O.advection_collision(s, p, box_scene)
    ctx = make_ctx(s.size, box_scene, p, terms)
    got = ctx.kernel_advection_collision(s)
    got_iters = ctx.fetch(capi.TAP_COLLISION_ITERS)
    ctx.close()
    assert (iters > 1).sum() > 100
    assert np.array_equal(got_iters, iters)
    for f in H.FIELDS_XYZ:
        assert np.array_equal(got[f][:, :3], want[f][:, :3]), f


def test_binary_search_fallback_matches_dense_table(box_scene):
    """A cell table too small for the grid switches to lower_bound lookups: same results."""
    p, terms, vol = H.config("water", 8192)
    s = H.state_s1(p, vol)
    dense, taps_d, _ = gpu_step_with_taps(s, p, terms, box_scene)
    sparse, taps_s, _ = gpu_step_with_taps(s, p, terms, box_scene, cell_table_capacity=8)
    for k in taps_d:
        assert np.array_equal(taps_d[k], taps_s[k]), k
    assert dense.tobytes() == sparse.tobytes()


@pytest.mark.parametrize("options", [dict(sub_cell_order=0, neighbour_lists=0), dict(sub_cell_order=0, neighbour_lists=1), dict(sub_cell_order=0, neighbour_lists=1, list_rows=8),
                                     dict(sub_cell_order=0, neighbour_lists=1, list_rows=24)])
def test_neighbour_organisations_agree_with_the_oracle(options, box_scene, plane_scene):
    """Two-pass search, stored neighbour lists, and lists too short for most particles (which
    sends them through the fallback kernel) all meet the same bar."""
    p, terms, vol = H.config("water", 20000)
    check_against_oracle(H.state_s1(p, vol), p, terms, box_scene, "water %r" % (options,), options=options)
    p, terms, vol = H.config("mucus", 6000)
    check_against_oracle(H.drop_state(p, vol, scene_floor_y=-1.0), p, terms, plane_scene, "crowded %r" % (options,),
                         options=options)


def test_simulate_single_frame_host_in_host_out(box_scene):
    """clsph_simulate_single_frame(in, out) == upload + step + download, in place allowed."""
    p, terms, vol = H.config("water", 4096)
    s = H.state_s1(p, vol)
    ctx = make_ctx(s.size, box_scene, p, terms, debug=False)
    ctx.upload(s)
    ctx.step(1)
    want = ctx.download()
    p_after = ctx.parameters()
    buf = s.copy()
    p2 = p.copy()
    ctx.simulate_single_frame(buf, p2, terms, out=buf)
    ctx.close()
    assert buf.tobytes() == want.tobytes()
    assert H.struct_bytes(p2) == H.struct_bytes(p_after)


def test_device_resident_steps_equal_host_round_trips(box_scene):
    """k device-resident sub-steps == k calls with a download/upload in between (bitwise)."""
    p, terms, vol = H.config("water", 8192)
    s = H.state_s1(p, vol)
    ctx = make_ctx(s.size, box_scene, p, terms, debug=False)
    ctx.upload(s)
    ctx.step(4)
    resident = ctx.download()
    cur = s
    for _ in range(4):
        ctx.upload(cur)
        ctx.step(1)
        cur = ctx.download()
    ctx.close()
    assert resident.tobytes() == cur.tobytes()


def test_several_steps_track_the_oracle(box_scene):
    """Errors stay at rounding level over a few sub-steps (not part of the 1e-4 single-step bar)."""
    p, terms, vol = H.config("water", 8192)
    s = H.state_s1(p, vol)
    ctx = make_ctx(s.size, box_scene, p, terms, debug=False)
    ctx.upload(s)
    ctx.step(5)
    got = ctx.download()
    ctx.close()
    po = p.copy()
    cur = s
    for _ in range(5):
        cur = O.step(cur, po, terms, box_scene, taps=False).particles
    # a key can flip for a particle within rounding of a cell boundary, which reorders the
    # arrays; compare order-independent aggregates
    for f in ("position", "velocity"):
        g, w = got[f][:, :3].astype(np.float64), cur[f][:, :3].astype(np.float64)
        assert np.abs(g.mean(0) - w.mean(0)).max() <= 1e-5 * max(1.0, np.abs(w).max()), f
        assert abs(np.abs(g).max() - np.abs(w).max()) <= 1e-3 * np.abs(w).max(), f
    assert abs(got["density"].astype(np.float64).mean() - cur["density"].astype(np.float64).mean()) <= 1e-4 * cur["density"].mean()


def test_error_paths(box_scene):
    p, terms, vol = H.config("water", 4096)
    ctx = capi.Context(4096)
    with pytest.raises(capi.ClsphError) as e:
        ctx.step(1)
    assert e.value.code == capi.E_STATE
    s = H.state_s0(p, vol)
    with pytest.raises(capi.ClsphError) as e:
        ctx.upload(s[:100])  # fewer than 128 particles (erratum E8)
    assert e.value.code == capi.E_INVAL
    big = np.zeros(5000, dtype=abi.PARTICLE)
    with pytest.raises(capi.ClsphError) as e:
        ctx.upload(big)
    assert e.value.code == capi.E_INVAL
    # grid overflow: two clusters 1100 cells apart -> CLSPH_EGRID (the reference asserts)
    ctx.set_scene(box_scene.face_normals, box_scene.vertices, box_scene.indices)
    ctx.set_parameters(p, terms)
    far = s.copy()
    far["position"][: s.size // 2, 0] += np.float32(1100 * 2 * p.h)
    ctx.upload(far)
    ctx.step(1)
    with pytest.raises(capi.ClsphError) as e:
        ctx.synchronize()
    assert e.value.code == capi.E_GRID
    # the context stays usable
    ctx.upload(s)
    ctx.step(1)
    ctx.synchronize()
    ctx.close()


def test_empty_scene_no_faces():
    p, terms, vol = H.config("water", 2048)
    s = H.state_s1(p, vol)
    empty = O.Scene(np.zeros(0, np.float32), np.zeros(0, np.uint32), np.zeros(0, np.float32))
    check_against_oracle(s, p, terms, empty, "no faces")


def test_one_million_particles_config2(box_scene):
    """BASELINE config 2 (dam break, 1 Mi particles, state S1): full comparison with the oracle
    plus size-independent properties."""
    p, terms, vol, _ = workloads.make_config("config2_dambreak_1m")
    s = workloads.jittered_state(p, vol)
    got, taps, want = check_against_oracle(s, p, terms, box_scene, "config 2, 1 Mi")
    sk = taps["skeys"]
    assert np.all(sk[1:] >= sk[:-1])
    assert np.array_equal(np.sort(taps["permutation"]), np.arange(s.size, dtype=np.uint32))
    assert np.all(np.isfinite(got["position"])) and np.all(np.isfinite(got["density"]))


def test_four_million_mucus_properties():
    """BASELINE config 3 size (4 Mi, mucus, labyrinth): properties that do not need the oracle."""
    p, terms, vol, scene_file = workloads.make_config("config3_mucus_labyrinth_4m")
    normals, vertices, indices = workloads.scene_arrays(scene_file)
    s = workloads.jittered_state(p, vol)
    ctx = capi.Context(s.size)
    ctx.set_scene(normals, vertices, indices)
    ctx.set_parameters(p, terms)
    ctx.set_debug(True)
    ctx.upload(s)
    ctx.step(1)
    ctx.synchronize()
    keys_in, perm, skeys = ctx.fetch(capi.TAP_KEYS_INPUT), ctx.fetch(capi.TAP_PERMUTATION), ctx.fetch(capi.TAP_SORTED_KEYS)
    table, cand, supp = ctx.fetch(capi.TAP_CELL_TABLE), ctx.fetch(capi.TAP_CANDIDATE_COUNT), ctx.fetch(capi.TAP_SUPPORT_COUNT)
    rho = ctx.fetch(capi.TAP_DENSITY)
    ctx.close()
    # stable sort by key: sorted, a permutation, and ties keep their input order
    assert np.array_equal(keys_in[perm], skeys)
    assert np.all(skeys[1:] >= skeys[:-1])
    assert np.array_equal(perm, np.argsort(keys_in, kind="stable").astype(np.uint32))
    # cell table = lower_bound of every cell id
    assert np.array_equal(table, np.searchsorted(skeys, np.arange(table.size, dtype=np.uint32), side="left").astype(np.uint32))
    assert supp.min() >= 1 and np.all(supp <= cand)
    assert np.all(np.isfinite(rho)) and rho.min() > 0
