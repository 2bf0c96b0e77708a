"""Multi-rank slab decomposition on the CPU: the CUDA sources (dist.cu included) under the test-only
emulator, one OS thread per rank, NCCL replaced by tests/emu/nccl_emu.cpp. Same checks as
tests/dist_worker.py does on real GPUs: after every sub-step the union of the ranks' owned
particles is a permutation of the ids, keys equal the single-rank run's, values agree to rounding,
and the first sub-step matches the oracle. Not a product path (see tests/emu/cuda_emu.h)."""
import threading

import numpy as np
import pytest

from libclsph_b200 import abi, capi, slabs, workloads
from oracle import oracle as O
from tests import helpers as H
from tests.emu import build_emu


@pytest.fixture(scope="module", autouse=True)
def emulator_library():
    saved = capi._lib
    capi._lib = capi.load_library(build_emu.build())
    yield capi._lib
    capi._lib = saved


def elongated_state(copies, n_cube=8000):
    """`copies` jittered lattice cubes side by side along x (one fluid block, long in x), sheared so
    that particles cross the slab planes."""
    p, terms, vol, scene_file = workloads.make_config(fluid="water", particles_count=n_cube, particle_mass=0.05)
    cube = workloads.jittered_state(p, vol)
    per_side, side, spacing = workloads.lattice_geometry(p, vol)
    parts = []
    for c in range(copies):
        s = cube.copy()
        s["position"][:, 0] += np.float32(c * per_side) * spacing
        parts.append(s)
    state = np.concatenate(parts)
    state["intermediate_velocity"][:, 0] += (2.5 * np.sign(state["position"][:, 2])).astype(np.float32)
    state["velocity"][:, 0] = state["intermediate_velocity"][:, 0]
    p.particles_count = state.size  # h, mass and the kernel constants are those of the cube's spacing
    return p, terms, state, scene_file


def rel(a, b):
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / np.abs(b).max())


def by_id(parts, ids, n):
    out = np.empty(n, dtype=abi.PARTICLE)
    out[ids] = parts
    return out


# world 1: a slab that is not cut at all; transport: stores into the neighbours' mailboxes (default) or NCCL messages
# "peer-copy": the exchange that copies the owned particles into a fresh array (CLSPH_DIST_IN_PLACE=0) instead of leaving them in place
# "peer-radix" / "peer-count": the ranks sort by radix passes (sub-cell table zeroed row by row) / by counting on the table
# whatever its size (zeroed as a range), the single rank they are compared with bitwise keeps the default rule
@pytest.mark.parametrize("world,copies,sub,transport", [(2, 2, 0, "peer"), (3, 3, 1, "peer"), (1, 2, 1, "peer"), (3, 3, 1, "nccl"),
                                                        (3, 3, 1, "peer-copy"), (3, 3, 1, "peer-radix"), (3, 3, 1, "peer-count")])
def test_slab_decomposition_matches_single_rank_and_oracle(world, copies, sub, transport, monkeypatch):
    steps = 3
    sort_option = {"peer-radix": 0, "peer-count": 2}.get(transport)
    sort_seen = []
    if transport == "nccl":
        monkeypatch.setenv("CLSPH_DIST_TRANSPORT", "nccl")
    if transport == "peer-copy":
        monkeypatch.setenv("CLSPH_DIST_IN_PLACE", "0")
    p, terms, state, scene_file = elongated_state(copies)
    n = state.size
    normals, vertices, indices = workloads.scene_arrays(scene_file)
    planes = slabs.equal_count_planes(state["position"][:, 0], world)
    owner = slabs.slab_of(state["position"][:, 0], planes)
    uid = capi.comm_unique_id()

    results = [[None] * steps for _ in range(world)]
    errors = []
    barrier = threading.Barrier(world)

    def run_rank(rank):
        try:
            mine = np.nonzero(owner == rank)[0].astype(np.uint32)
            ctx = capi.Context(int(n * (1.0 / world + 0.5)) + 4096)
            ctx.set_option("sub_cell_order", sub)
            if sort_option is not None:
                ctx.set_option("count_sort", sort_option)
            ctx.set_scene(normals, vertices, indices)
            ctx.set_parameters(p, terms)
            ctx.dist_init(rank, world, uid, float(planes[rank]), float(planes[rank + 1]))
            assert ctx.dist_transport().startswith("nccl" if transport == "nccl" else "peer stores")
            ctx.dist_upload(np.ascontiguousarray(state[mine]), mine)
            for k in range(steps):
                ctx.step(1)
                ctx.synchronize()
                results[rank][k] = ctx.dist_download()
                sort_seen.append(ctx.sort_passes())
            barrier.wait(timeout=600)
            ctx.close()
        except BaseException as exc:  # noqa: BLE001 - reported by the main thread
            errors.append((rank, exc))
            barrier.abort()

    threads = [threading.Thread(target=run_rank, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()

    # meanwhile: the whole block on one emulated device, and the oracle for the first sub-step
    single = capi.Context(n)
    single.set_option("sub_cell_order", sub)
    single.set_scene(normals, vertices, indices)
    single.set_parameters(p, terms)
    single.upload(state)
    wants, orders, ref_ids = [], [], np.arange(n, dtype=np.uint32)
    for k in range(steps):
        single.step(1)
        want = single.download()
        ref_ids = ref_ids[single.fetch(capi.TAP_PERMUTATION)]
        wants.append(by_id(want, ref_ids, n))
        orders.append(ref_ids.copy())  # id of the particle at each position of the reference's array
    single.close()
    scene = O.Scene(vertices, indices, normals)
    r0 = O.step(state, p.copy(), terms, scene)
    oracle_by_id = by_id(r0.particles, r0.permutation, n)

    for t in threads:
        t.join(timeout=900)
    assert not errors, errors
    moved_total = 0
    for k in range(steps):
        parts = np.concatenate([results[r][k][0] for r in range(world)])
        ids = np.concatenate([results[r][k][1] for r in range(world)])
        assert parts.size == n, "step %d: particles lost or duplicated (%d of %d)" % (k, parts.size, n)
        assert np.array_equal(np.sort(ids), np.arange(n, dtype=np.uint32)), "step %d: ids are not a permutation" % k
        got = by_id(parts, ids, n)
        holder = np.empty(n, dtype=np.int64)
        holder[ids] = np.concatenate([np.full(results[r][k][1].size, r) for r in range(world)])
        moved_total = int((holder != owner).sum())
        if sub:
            # the reference's GLOBAL array order: merge the ranks' downloads by (grid_index, rank in cell),
            # the latter carried in the records' padding word (clsph_dist_download)
            order = np.lexsort((parts["_pad"], parts["grid_index"]))
            assert np.array_equal(ids[order], orders[k]), "step %d: merged order differs from the single-rank array order" % k
            # ... and, because every rank keeps the particles of a sub-cell in that order too (ghosts included),
            # all sums run in the order a single device uses: the decomposition is BITWISE transparent
            for f in ("position", "velocity", "intermediate_velocity", "density", "pressure", "grid_index"):
                assert got[f].tobytes() == wants[k][f].tobytes(), "step %d: %s differs bitwise from the single-rank run" % (k, f)
        if k == 0:
            assert np.array_equal(got["grid_index"], wants[0]["grid_index"]), "keys differ from the single-rank run"
            assert np.array_equal(got["grid_index"], oracle_by_id["grid_index"]), "keys differ from the oracle"
            for f in ("position", "velocity", "density", "pressure"):
                g = got[f][:, :3] if got[f].ndim == 2 else got[f]
                w = oracle_by_id[f][:, :3] if oracle_by_id[f].ndim == 2 else oracle_by_id[f]
                assert rel(g, w) <= 1e-4, (f, rel(g, w))
        for f in ("position", "velocity", "intermediate_velocity", "density", "pressure"):
            g = got[f][:, :3] if got[f].ndim == 2 else got[f]
            w = wants[k][f][:, :3] if wants[k][f].ndim == 2 else wants[k][f]
            assert rel(g, w) <= (1e-4 if k == 0 else 2e-3), (k, f, rel(g, w))
    assert moved_total > 0 or world == 1, "the shear should have moved particles across a slab plane"
    if sort_option == 0:
        assert all(v >= 1 for v in sort_seen), sort_seen
    if sort_option == 2:
        assert all(v == 0 for v in sort_seen), sort_seen


def test_buffer_overflow_is_reported_and_nobody_hangs():
    """Message buffers too small for the ghost layers (what killed the first 8-GPU run of round 1): the rank
    that overflows reports CLSPH_ECOMM at its next synchronising call, every rank finishes its sub-steps
    (fixed-size messages: nobody waits for data that never comes), contexts close cleanly."""
    world = 2
    p, terms, state, scene_file = elongated_state(2)
    n = state.size
    normals, vertices, indices = workloads.scene_arrays(scene_file)
    planes = slabs.equal_count_planes(state["position"][:, 0], world)
    owner = slabs.slab_of(state["position"][:, 0], planes)
    uid = capi.comm_unique_id()
    codes = [None] * world
    errors = []

    def run_rank(rank):
        try:
            mine = np.nonzero(owner == rank)[0].astype(np.uint32)
            ctx = capi.Context(n)
            ctx.set_scene(normals, vertices, indices)
            ctx.set_parameters(p, terms)
            ctx.dist_init(rank, world, uid, float(planes[rank]), float(planes[rank + 1]), emigrant_capacity=64, ghost_capacity=64)
            ctx.dist_upload(np.ascontiguousarray(state[mine]), mine)
            ctx.step(3)
            try:
                ctx.synchronize()
                codes[rank] = 0
            except capi.ClsphError as e:
                codes[rank] = e.code
            ctx.close()
        except BaseException as exc:  # noqa: BLE001
            errors.append((rank, exc))

    threads = [threading.Thread(target=run_rank, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert not errors, errors
    assert not any(t.is_alive() for t in threads), "a rank is still waiting"
    assert capi.E_COMM in codes, codes
