"""Kernel LOGIC tests on the CPU: the CUDA sources compiled against the test-only emulator
(tests/emu/) and driven through the same C ABI and the same checks as the GPU parity tests, at
small sizes. Not a product path, not a fallback -- capi.Context keeps loading libclsph_cuda.so and
fails without a GPU; this module swaps in tests/emu/_build/libclsph_emu.so for its own duration
only. It catches indexing, collective, barrier and bookkeeping errors before GPU time is spent;
the -m gpu suite stays the parity gate (the emulator knows nothing about races or performance).
"""
import os
import threading

import numpy as np
import pytest

from libclsph_b200 import capi, workloads
from oracle import oracle as O
from tests import helpers as H
from tests import test_gpu_parity as G
from tests.emu import build_emu


@pytest.fixture(scope="module", autouse=True)
def emulator_library():
    saved = capi._lib
    capi._lib = capi.load_library(build_emu.build())
    yield capi._lib
    capi._lib = saved


def test_every_abi_symbol_is_exported_by_the_emulator_build(emulator_library):
    for name in capi.SYMBOLS:
        assert hasattr(emulator_library, name), name


def test_smoke_entry_point_under_the_emulator(capsys):
    """__graft_entry__.smoke() is what the driver runs on the GPU box; here its logic (options, taps, oracle comparison)."""
    import __graft_entry__ as entry
    entry.smoke()
    assert "smoke ok" in capsys.readouterr().out


@pytest.mark.parametrize("n", [128, 1000, 4096])
def test_lattice_state_s0(n, box_scene):
    G.test_lattice_state_s0(n, box_scene)


@pytest.mark.parametrize("fluid,n", [("water", 4096), ("mucus", 3000)])
def test_jittered_state_s1(fluid, n, box_scene):
    G.test_jittered_state_s1(fluid, n, box_scene)


def test_ragged_count(box_scene):
    p, terms, vol = H.config("water", 2345)
    G.check_against_oracle(H.state_s1(p, vol, seed=3), p, terms, box_scene, "ragged n=2345")


def test_collision_heavy_step(plane_scene):
    p, terms, vol = H.config("mucus", 2048)
    s = H.drop_state(p, vol, scene_floor_y=-1.0)
    got, taps, want = G.check_against_oracle(s, p, terms, plane_scene, "drop onto plane")
    assert (want.collision_iters > 1).mean() > 0.2


def test_labyrinth_scene_many_faces():
    scene = O.load_obj(os.path.join(H.ROOT, "scenes", "labyrinth.obj"))
    p, terms, vol = H.config("mucus", 4096, mass=0.05 * 32000 / 4194304 * 256)
    G.check_against_oracle(H.state_s1(p, vol), p, terms, scene, "labyrinth")


def test_advection_collision_kernel_bit_exact(box_scene):
    p, terms, vol = H.config("water", 4000)
    s = H.drop_state(p, vol, scene_floor_y=-2.0, speed=2.9, slab=0.02)
    rng = np.random.default_rng(11)
    s["acceleration"][:, :3] = rng.normal(0, 30, size=(s.size, 3)).astype(np.float32)
    s["intermediate_velocity"][:, 0] = rng.uniform(-2, 2, s.size).astype(np.float32)
    want, iters = O.advection_collision(s, p, box_scene)
    ctx = G.make_ctx(s.size, box_scene, p, terms)
    got = ctx.kernel_advection_collision(s)
    got_iters = ctx.fetch(capi.TAP_COLLISION_ITERS)
    ctx.close()
    assert (iters > 1).sum() > 20
    assert np.array_equal(got_iters, iters)
    for f in H.FIELDS_XYZ:
        assert np.array_equal(got[f][:, :3], want[f][:, :3]), f


def test_binary_search_fallback_matches_dense_table(box_scene):
    p, terms, vol = H.config("water", 2048)
    s = H.state_s1(p, vol)
    dense, taps_d, _ = G.gpu_step_with_taps(s, p, terms, box_scene)
    sparse, taps_s, _ = G.gpu_step_with_taps(s, p, terms, box_scene, cell_table_capacity=8)
    for k in taps_d:
        assert np.array_equal(taps_d[k], taps_s[k]), k
    assert dense.tobytes() == sparse.tobytes()


# (the pair density kernel is the default from 160 000 particles; the small states here ask for it explicitly)
SUB = dict(sub_cell_order=1, pair_density=1)


@pytest.mark.parametrize("options", [dict(sub_cell_order=0, neighbour_lists=0), dict(sub_cell_order=0, neighbour_lists=1), dict(sub_cell_order=0, neighbour_lists=1, list_rows=8),
                                     SUB, dict(sub_cell_order=1, list_rows=8), dict(sub_cell_order=1, forces_blocks=4),
                                     dict(sub_cell_order=1, fast_pairs=1), dict(sub_cell_order=0, neighbour_lists=1, fast_pairs=1, forces_blocks=4, face_grid=1),
                                     dict(sub_cell_order=1, merged_rows=1), dict(sub_cell_order=1, merged_rows=1, list_rows=8),
                                     dict(factored_forces=0, pair_density=1), dict(factored_forces=0, list_rows=8), dict(pair_density=0), dict(pair_density=1, list_rows=8), dict(count_sort=0), dict()])
def test_neighbour_organisations(options, box_scene, plane_scene):
    p, terms, vol = H.config("water", 3000)
    G.check_against_oracle(H.state_s1(p, vol), p, terms, box_scene, "water %r" % (options,), options=options)
    p, terms, vol = H.config("mucus", 1500)
    G.check_against_oracle(H.drop_state(p, vol, scene_floor_y=-1.0), p, terms, plane_scene, "crowded %r" % (options,),
                           options=options)


def test_host_in_host_out_and_resident_steps(box_scene):
    G.test_simulate_single_frame_host_in_host_out(box_scene)
    p, terms, vol = H.config("water", 2048)
    s = H.state_s1(p, vol)
    ctx = G.make_ctx(s.size, box_scene, p, terms, debug=False)
    ctx.upload(s)
    ctx.step(3)
    resident = ctx.download()
    cur = s
    for _ in range(3):
        ctx.upload(cur)
        ctx.step(1)
        cur = ctx.download()
    ctx.close()
    assert resident.tobytes() == cur.tobytes()


def test_error_paths(box_scene):
    G.test_error_paths(box_scene)


def test_empty_scene_no_faces():
    G.test_empty_scene_no_faces()


# ---- sub-cell order (subgrid.cu): arrays sorted by (cell key << 3 | octant), reference order carried as a rank

@pytest.mark.parametrize("n", [128, 1000, 4096])
def test_sub_cell_order_lattice(n, box_scene):
    p, terms, vol = H.config("water", n)
    G.check_against_oracle(H.state_s0(p, vol), p, terms, box_scene, "sub S0 n=%d" % n, options=SUB)


def test_sub_cell_order_mucus_labyrinth():
    scene = O.load_obj(os.path.join(H.ROOT, "scenes", "labyrinth.obj"))
    p, terms, vol = H.config("mucus", 4096, mass=0.05 * 32000 / 4194304 * 256)
    G.check_against_oracle(H.state_s1(p, vol), p, terms, scene, "sub labyrinth", options=SUB)


def test_sub_cell_order_binary_search_fallback(box_scene):
    p, terms, vol = H.config("water", 2048)
    s = H.state_s1(p, vol)
    dense, taps_d, _ = G.gpu_step_with_taps(s, p, terms, box_scene, options=SUB)
    sparse, taps_s, _ = G.gpu_step_with_taps(s, p, terms, box_scene, cell_table_capacity=8, options=SUB)
    for k in taps_d:
        assert np.array_equal(taps_d[k], taps_s[k]), k
    assert dense.tobytes() == sparse.tobytes()
    G.check_against_oracle(s, p, terms, box_scene, "sub, binary search", cell_table_capacity=8, options=SUB)


@pytest.mark.parametrize("options", [dict(sub_cell_order=0, neighbour_lists=1), SUB, dict(count_sort=0)])
def test_resident_steps_keep_the_reference_order(options, box_scene, plane_scene):
    p, terms, vol = H.config("water", 3000)
    s = H.state_s1(p, vol)
    s["intermediate_velocity"][:, 0] += (2.5 * np.sign(s["position"][:, 2])).astype(np.float32)  # shear: cells change
    G.check_resident_steps_against_oracle(s, p, terms, box_scene, 5, "water %r" % (options,), options=options)
    p, terms, vol = H.config("mucus", 1500)
    G.check_resident_steps_against_oracle(H.drop_state(p, vol, scene_floor_y=-1.0), p, terms, plane_scene, 3,
                                          "crowded %r" % (options,), options=options)


@pytest.mark.parametrize("options", [SUB, dict(sub_cell_order=1, merged_rows=1, fast_pairs=1, face_grid=1)])
def test_sub_cell_order_resident_steps_equal_host_round_trips_bitwise(options, box_scene):
    """Inside a sub-cell the particles are kept in the reference's order, so the arrays depend on the state
    alone and not on its history: k resident sub-steps == k upload/step/download round trips, bit for bit."""
    p, terms, vol = H.config("water", 3000)
    s = H.state_s1(p, vol)
    s["intermediate_velocity"][:, 0] += (2.5 * np.sign(s["position"][:, 2])).astype(np.float32)
    ctx = G.make_ctx(s.size, box_scene, p, terms, debug=False, options=options)
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


@pytest.mark.parametrize("fluid,n", [("water", 6000), ("mucus", 3000)])
def test_pair_density_is_bitwise_the_per_particle_kernel(fluid, n, box_scene):
    """k_density_pairs (two particles of a sub-cell per thread, packed fp32) against k_density_sub<merged>: each packed
    lane rounds like the scalar code and every particle meets its candidates in the same order, so after several
    resident sub-steps every byte of the state and every tap is the same."""
    p, terms, vol = H.config(fluid, n)
    s = H.state_s1(p, vol)
    s["intermediate_velocity"][:, 0] += (2.5 * np.sign(s["position"][:, 2])).astype(np.float32)
    outs = []
    for pair in (1, 0):
        ctx = G.make_ctx(s.size, box_scene, p, terms, debug=True, options=dict(pair_density=pair, merged_rows=1, factored_forces=0))
        ctx.upload(s)
        ctx.step(3)
        outs.append((ctx.download().tobytes(), ctx.fetch(capi.TAP_SUPPORT_COUNT).tobytes(), ctx.fetch(capi.TAP_CANDIDATE_COUNT).tobytes(),
                     ctx.fetch(capi.TAP_ACCELERATION).tobytes()))
        ctx.close()
    assert outs[0] == outs[1]


@pytest.mark.parametrize("fluid,n", [("water", 6000), ("mucus", 3000)])
def test_counting_sort_is_bitwise_the_radix_sort(fluid, n, box_scene):
    G.check_counting_sort_against_radix(fluid, n, box_scene)


def test_counting_sort_on_a_table_much_larger_than_the_fluid(box_scene):
    G.check_counting_sort_on_a_sparse_table(3000, box_scene)


def test_sub_cell_order_host_round_trip_and_option_rules(box_scene):
    p, terms, vol = H.config("water", 2048)
    s = H.state_s1(p, vol)
    ctx = G.make_ctx(s.size, box_scene, p, terms, debug=False, options=SUB)
    ctx.upload(s)
    ctx.step(1)
    want = ctx.download()
    with pytest.raises(capi.ClsphError) as e:
        ctx.set_option("sub_cell_order", 0)  # the arrays are in another order: not while particles are held
    assert e.value.code == capi.E_STATE
    buf = s.copy()
    ctx.simulate_single_frame(buf, p.copy(), terms, out=buf)
    ctx.close()
    assert buf.tobytes() == want.tobytes()


# ---- face grid (integrate.cu): only the faces of the cells a segment touches are tested; bit-identical results

@pytest.mark.parametrize("scene_file", ["labyrinth.obj", "river.obj", "box.obj", "cone.obj", "shower.obj", "monkey.obj"])
def test_face_grid_is_bit_identical_to_testing_every_face(scene_file):
    scene = O.load_obj(os.path.join(H.ROOT, "scenes", scene_file))
    p, terms, vol = H.config("water", 4096)
    total_hits = 0
    # short segments (a sub-step's millimetres), segments crossing a few grid cells, and segments so
    # long that the kernel falls back to the full scan
    for reach, speed, vmax, seed in [(0.002, 3.0, None, 1), (0.05, 80.0, 80.0, 2), (0.5, 3000.0, 3000.0, 3)]:
        q = p.copy()
        q.particles_count = 1500  # the oracle takes the count from the parameters
        if vmax is not None:
            q.max_velocity = vmax
        s = H.surface_state(scene, q.particles_count, reach, speed, seed)
        want, iters = O.advection_collision(s, q, scene)
        for grid_on in (0, 1):
            ctx = G.make_ctx(s.size, scene, q, terms, options=dict(face_grid=grid_on))
            got = ctx.kernel_advection_collision(s)
            got_iters = ctx.fetch(capi.TAP_COLLISION_ITERS)
            ctx.close()
            what = "%s reach %g grid %d" % (scene_file, reach, grid_on)
            assert np.array_equal(got_iters, iters), what
            for f in H.FIELDS_XYZ:
                assert got[f].tobytes() == want[f].tobytes() or np.array_equal(got[f][:, :3], want[f][:, :3]), what + " " + f
        total_hits += int((iters > 1).sum())
    assert total_hits > 100, total_hits


def test_face_grid_in_the_full_step(plane_scene):
    scene = O.load_obj(os.path.join(H.ROOT, "scenes", "labyrinth.obj"))
    p, terms, vol = H.config("mucus", 4096, mass=0.05 * 32000 / 4194304 * 256)
    G.check_against_oracle(H.state_s1(p, vol), p, terms, scene, "labyrinth, face grid", options=dict(face_grid=1, sub_cell_order=1))
    p, terms, vol = H.config("mucus", 1500)
    G.check_resident_steps_against_oracle(H.drop_state(p, vol, scene_floor_y=-1.0), p, terms, plane_scene, 3,
                                          "crowded, face grid", options=dict(face_grid=1))


def test_sub_cell_order_reports_a_grid_too_large_for_its_keys(box_scene):
    """With 512 to 1023 cells along z the reference still works but (Morton cell key << 3 | octant)
    no longer fits 32 bits: CLSPH_EGRID with advice, and the context stays usable."""
    p, terms, vol = H.config("water", 4096)
    s = H.state_s0(p, vol)
    far = s.copy()
    far["position"][: s.size // 2, 2] += np.float32(700 * 2 * p.h)
    for sub, expect_error in ((1, True), (0, False)):
        ctx = G.make_ctx(s.size, box_scene, p, terms, debug=False, options=dict(sub_cell_order=sub))
        ctx.upload(far)
        ctx.step(1)
        if expect_error:
            with pytest.raises(capi.ClsphError) as e:
                ctx.synchronize()
            assert e.value.code == capi.E_GRID and "sub_cell_order" in str(e.value)
        else:
            ctx.synchronize()
        ctx.upload(s)
        ctx.step(1)
        ctx.synchronize()
        ctx.close()


@pytest.mark.parametrize("kind", H.EDGE_KINDS)
@pytest.mark.parametrize("options", [dict(sub_cell_order=0, neighbour_lists=0), dict(sub_cell_order=0, neighbour_lists=1), dict(sub_cell_order=1, face_grid=1),
                                     dict(sub_cell_order=1, face_grid=1, pair_density=0),
                                     dict(sub_cell_order=1, face_grid=1, fast_pairs=1), dict(sub_cell_order=1, merged_rows=1)])
def test_edge_states(kind, options, box_scene):
    """States sitting ON the path's decisions (tests/helpers.edge_state; the oracle is pinned against the
    reference's own kernels on the same states in test_oracle_vs_ref.py), in every organisation."""
    p, terms, vol = H.config("water", 1024)
    s = H.edge_state(kind, p, vol)
    G.check_against_oracle(s, p, terms, box_scene, "%s %r" % (kind, options), options=options)


@pytest.mark.parametrize("options", [dict(sub_cell_order=0, neighbour_lists=1), dict(sub_cell_order=1, face_grid=1)])
def test_blown_up_particles_do_not_hang_the_kernels(options, box_scene):
    """A particle at infinity makes the grid overflow (CLSPH_EGRID, as the reference's assert would);
    a NaN position falls into cell 0. Either way every kernel must terminate."""
    p, terms, vol = H.config("water", 2048)
    s = H.state_s1(p, vol)
    bad = s.copy()
    bad["position"][5, 0] = np.inf
    bad["position"][6, 1] = -np.inf
    bad["position"][7, 2] = np.float32(3e38)
    ctx = G.make_ctx(s.size, box_scene, p, terms, debug=False, options=options)
    ctx.upload(bad)
    ctx.step(2)
    with pytest.raises(capi.ClsphError) as e:
        ctx.synchronize()
    assert e.value.code == capi.E_GRID
    nan = s.copy()
    nan["position"][9, :3] = np.nan
    ctx.upload(nan)
    ctx.step(2)
    ctx.synchronize()
    out = ctx.download()
    assert np.isfinite(out["position"][:, :3]).sum() >= 3 * (s.size - 64)  # the NaN may spread to its neighbours, not further
    ctx.close()


@pytest.mark.parametrize("options", [dict(sub_cell_order=0, neighbour_lists=0), dict(sub_cell_order=0, neighbour_lists=1), dict(sub_cell_order=1, face_grid=1, fast_pairs=1),
                                     dict(sub_cell_order=1, face_grid=1, pair_density=0, forces_blocks=4),
                                     dict(sub_cell_order=1, face_grid=1, merged_rows=1, fast_pairs=1)])
def test_developed_state(options):
    """State S2: fluid that has hit the floor of the box and spread -- free surface, wall
    contacts, ~10 % of the particles colliding in the step."""
    p, terms, scene, s = H.developed_state()
    got, taps, want = G.check_against_oracle(s, p, terms, scene, "developed %r" % (options,), options=options)
    assert (want.collision_iters > 1).sum() > 100


def test_developed_state_resident_steps_in_sub_cell_order():
    p, terms, scene, s = H.developed_state()
    G.check_resident_steps_against_oracle(s, p, terms, scene, 4, "developed, resident", options=dict(sub_cell_order=1, face_grid=1, fast_pairs=1))


def test_option_validation():
    ctx = capi.Context(1024)
    for name, value in (("forces_blocks", 5), ("list_rows", 5000), ("no_such_option", 1)):
        with pytest.raises(capi.ClsphError) as e:
            ctx.set_option(name, value)
        assert e.value.code == capi.E_INVAL, name
    for name, value in (("forces_blocks", 4), ("forces_blocks", 3), ("fast_pairs", 1), ("merged_rows", 1), ("pair_density", 0), ("pair_variant", 3), ("factored_forces", 1), ("count_sort", 0), ("count_sort", 2),
                        ("face_grid", 1), ("sub_cell_order", 1), ("sub_cell_order", 0), ("neighbour_lists", 0), ("list_rows", 48)):
        ctx.set_option(name, value)
    ctx.close()


def test_frame_points_are_the_downloaded_fields(box_scene):
    """clsph_frame_begin / clsph_frame_end: the seven floats per particle a frame file needs, packed on the device in
    the reference's output order, equal the same fields of the full download bit for bit."""
    p, terms, vol = H.config("water", 3000)
    s = H.state_s1(p, vol)
    ctx = G.make_ctx(s.size, box_scene, p, terms, debug=False)
    ctx.upload(s)
    ctx.step(2)
    pts = ctx.frame_points()
    got = ctx.download()
    ctx.close()
    assert pts.shape == (s.size, 7)
    assert np.array_equal(pts[:, 0:3], got["position"][:, :3]) and np.array_equal(pts[:, 3:6], got["velocity"][:, :3])
    assert np.array_equal(pts[:, 6], got["density"])
