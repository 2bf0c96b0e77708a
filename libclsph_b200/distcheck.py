"""Bitwise parity of the slab decomposition against a single-GPU run, on the ranks of a running job.

    report = distcheck.bitwise_parity(dist, rank, world, local_rank, n_total=60000 * world, steps=4)

Every rank runs its slab of one fluid block through the CUDA library (clsph_dist_*); rank 0 also runs the
whole block on its own GPU without decomposition. After every sub-step the ranks' owned particles are
gathered on rank 0 and compared with the single-GPU array:
  * nothing lost, nothing duplicated (the persistent ids are a permutation of 0..n-1);
  * the ranks' downloads merged by (grid_index, rank inside the cell) reproduce the single-GPU array ORDER;
  * position, velocity, half-step velocity, density, pressure and grid_index of every particle are equal
    BIT FOR BIT (each rank keeps the particles of a sub-cell in the order a single GPU does, ghosts carry
    their order keys, so every floating-point sum runs in the same order).
The state is a sheared block (the upper half moves right, the lower half left, 2.5 m/s) so that particles
really cross the slab planes during the check.

Used by bench.py before it times a multi-GPU run (`multi_gpu_parity` in its JSON line) and by
tests/dist_worker.py. No CPU code takes part: the single-GPU run is the product's own kernels, which the GPU
parity suite pins to the oracle.
"""
import numpy as np

from . import abi, capi, slabs, workloads

FIELDS = ("position", "velocity", "intermediate_velocity", "density", "pressure", "grid_index")


def _gather_owned(dist, ctx, rank, world):
    """All ranks' owned particles and ids, concatenated in rank order (only rank 0 uses them)."""
    parts, ids = ctx.dist_download()
    got = [None] * world
    dist.all_gather_object(got, (parts.tobytes(), ids.tobytes()))
    counts = [len(i) // 4 for _, i in got]
    if rank != 0:
        return None, None, counts
    all_p = np.concatenate([np.frombuffer(p, dtype=abi.PARTICLE) for p, _ in got])
    all_i = np.concatenate([np.frombuffer(i, dtype=np.uint32) for _, i in got])
    return all_p, all_i, counts


def sheared_block(n_total):
    p, terms, vol, scene_file = workloads.make_config(fluid="water", particles_count=n_total, particle_mass=0.05)
    state = workloads.jittered_state(p, vol)
    state["intermediate_velocity"][:, 0] += (2.5 * np.sign(state["position"][:, 2])).astype(np.float32)
    state["velocity"][:, 0] = state["intermediate_velocity"][:, 0]
    return p, terms, vol, scene_file, state


def bitwise_parity(dist, rank, world, local_rank, n_total, steps=4, options=(), verbose=False):
    """Returns the report dict (identical on every rank). `options`: "name=value" strings for both runs."""
    p, terms, vol, scene_file, state = sheared_block(n_total)
    normals, vertices, indices = workloads.scene_arrays(scene_file)
    planes = slabs.equal_count_planes(state["position"][:, 0], world)
    owner = slabs.slab_of(state["position"][:, 0], planes)
    mine = np.nonzero(owner == rank)[0].astype(np.uint32)

    box = [capi.comm_unique_id() if rank == 0 else None]  # one NCCL id for the library's own communicator
    dist.broadcast_object_list(box, 0)
    uid = bytes(box[0])

    def make(capacity):
        ctx = capi.Context(capacity, device=local_rank)
        # the check runs on a small block; keep it on the density kernel the timed (large) run uses, whatever the size rule says
        if not any(opt.startswith("pair_density=") for opt in options):
            ctx.set_option("pair_density", 1)
        for opt in options:
            k, v = opt.split("=")
            ctx.set_option(k, int(v))
        ctx.set_scene(normals, vertices, indices)
        ctx.set_parameters(p, terms)
        return ctx

    ctx = make(int(n_total * (1.0 / world + 0.5)) + 4096)  # owned + ghost layers on both sides, with slack
    ctx.dist_init(rank, world, uid, float(planes[rank]), float(planes[rank + 1]))
    ctx.dist_upload(np.ascontiguousarray(state[mine]), mine)
    single = None
    if rank == 0:
        single = make(n_total)
        single.upload(state)

    report = {"particles": int(n_total), "world": int(world), "substeps": int(steps), "bitwise": True, "order": True,
              "ids_exact": True, "migrated": 0, "first_failure": None}
    ref_ids = np.arange(n_total, dtype=np.uint32)
    for k in range(steps):
        ctx.step(1)
        ctx.synchronize()
        got, ids, counts = _gather_owned(dist, ctx, rank, world)
        if rank != 0:
            continue
        single.step(1)
        want = single.download()
        ref_ids = ref_ids[single.fetch(capi.TAP_PERMUTATION)]  # id of the particle at each position of the single run
        exact = sum(counts) == n_total and np.array_equal(np.sort(ids), np.arange(n_total, dtype=np.uint32))
        report["ids_exact"] = report["ids_exact"] and bool(exact)
        if not exact:  # (no early exit: the other ranks keep calling the collectives of the remaining sub-steps)
            report["bitwise"] = report["order"] = False
            report["first_failure"] = report["first_failure"] or "sub-step %d: particles lost or duplicated, counts %r" % (k, counts)
            continue
        by_id_got = np.empty(n_total, dtype=abi.PARTICLE)
        by_id_got[ids] = got
        by_id_want = np.empty(n_total, dtype=abi.PARTICLE)
        by_id_want[ref_ids] = want
        merged = np.lexsort((got["_pad"], got["grid_index"]))
        same_order = bool(np.array_equal(ids[merged], ref_ids))
        bad = [f for f in FIELDS if by_id_got[f].tobytes() != by_id_want[f].tobytes()]
        report["order"] = report["order"] and same_order
        report["bitwise"] = report["bitwise"] and not bad
        if (bad or not same_order) and not report["first_failure"]:
            report["first_failure"] = "sub-step %d: %s" % (k, ("fields differ: " + ",".join(bad)) if bad else "merged order differs")
        holder = np.empty(n_total, dtype=np.int64)
        holder[ids] = np.concatenate([np.full(c, r) for r, c in enumerate(counts)])
        report["migrated"] = int((holder != owner).sum())
        if verbose:
            print("distcheck sub-step %d: counts %r order %s bitwise %s migrated %d" % (k, counts, same_order, not bad, report["migrated"]),
                  flush=True)
    ctx.close()
    if single is not None:
        single.close()
    box = [report if rank == 0 else None]  # the verdict travels to every rank
    dist.broadcast_object_list(box, 0)
    return box[0]
