"""CPU model of the multi-GPU slab protocol, run with world_size >= 2 over gloo:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/slab_protocol_worker.py

Each rank plays one GPU of libclsph_b200/csrc/dist.cu with numpy + the CPU oracle as its compute:
all-reduced AABB -> identical grid, planes snapped to cell boundaries (slabs.cell_planes),
emigrants sent to the neighbour and kept as ghost copies, two ghost cell layers per side, density
for owned + first ghost layer, forces and integration for owned cells only. The union of the
ranks' owned particles must reproduce a single-rank oracle run particle by particle (matched by
id): this checks the protocol rules themselves, independently of the CUDA implementation."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from libclsph_b200 import abi, slabs, workloads  # noqa: E402
from oracle import oracle as O  # noqa: E402


def exchange(rank, world, to_left, to_right):
    """Sends one numpy byte blob to each neighbour, returns (from_left, from_right)."""
    out = [None, None]
    for direction, peer, payload in ((0, rank - 1, to_left), (1, rank + 1, to_right)):
        if 0 <= peer < world:
            size = torch.tensor([payload.size], dtype=torch.int64)
            other = torch.zeros(1, dtype=torch.int64)
            ops = [dist.P2POp(dist.isend, size, peer), dist.P2POp(dist.irecv, other, peer)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            buf = torch.zeros(int(other.item()), dtype=torch.uint8)
            snd = torch.from_numpy(payload.copy()) if payload.size else torch.zeros(0, dtype=torch.uint8)
            ops = []
            if payload.size:
                ops.append(dist.P2POp(dist.isend, snd, peer))
            if buf.numel():
                ops.append(dist.P2POp(dist.irecv, buf, peer))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            out[direction] = buf.numpy()
    return out


REC = np.dtype([("p", abi.PARTICLE), ("id", "<u4"), ("emigrant", "<u4")])


def pack(parts, ids, emigrant):
    r = np.zeros(parts.size, dtype=REC)
    r["p"], r["id"], r["emigrant"] = parts, ids, emigrant
    return r.view(np.uint8).reshape(-1)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    n, steps = 6000, 3
    p, terms, vol, scene_file = workloads.make_config(fluid="water", particles_count=n, particle_mass=0.05)
    state = workloads.jittered_state(p, vol)
    state["position"][:, 0] *= np.float32(0.75 * world)  # stretch the block so every slab is > 4 cells thick
    # shear along x so that particles really cross the planes
    state["intermediate_velocity"][:, 0] += (2.5 * np.sign(state["position"][:, 2])).astype(np.float32)
    state["velocity"][:, 0] = state["intermediate_velocity"][:, 0]
    scene = O.load_obj(os.path.join(ROOT, "scenes", scene_file))
    planes = slabs.equal_count_planes(state["position"][:, 0], world)
    mine = slabs.slab_of(state["position"][:, 0], planes) == rank
    owned, ids = state[mine].copy(), np.nonzero(mine)[0].astype(np.uint32)

    single = state
    single_ids = np.arange(n, dtype=np.uint32)
    pg = p.copy()
    migrated_total = 0
    for step in range(steps):
        # 1. global AABB -> the same grid everywhere (sph_simulation.cpp:201-252 on the reduced bounds)
        lo = torch.from_numpy(owned["position"][:, :3].min(0).copy())
        hi = torch.from_numpy(owned["position"][:, :3].max(0).copy())
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        box = np.zeros(2, dtype=abi.PARTICLE)
        box["position"][0, :3], box["position"][1, :3] = lo.numpy(), hi.numpy()
        pl = p.copy()
        pl.particles_count = 2
        assert O.bounds_and_grid(box, pl) == 0
        cell = np.float32(p.h * 2)
        bx = slabs.cell_planes(planes, pl.min_point.s[0], cell, pl.grid_size_x)
        own_lo, own_hi = bx[rank], bx[rank + 1]
        assert own_hi - own_lo >= 4, "slabs must be at least four cells thick"
        # 2. classify by the new cell (the k_dist_classify rules)
        cx = ((owned["position"][:, 0] - np.float32(pl.min_point.s[0])) / cell).astype(np.uint32).astype(np.int64)
        go_left, go_right = (cx < own_lo) & (rank > 0), (cx >= own_hi) & (rank + 1 < world)
        stay = ~go_left & ~go_right
        ghost_left, ghost_right = stay & (cx < own_lo + 2) & (rank > 0), stay & (cx >= own_hi - 2) & (rank + 1 < world)
        migrated_total += int(go_left.sum() + go_right.sum())
        msg_l = np.concatenate([pack(owned[go_left], ids[go_left], 1), pack(owned[ghost_left], ids[ghost_left], 0)])
        msg_r = np.concatenate([pack(owned[go_right], ids[go_right], 1), pack(owned[ghost_right], ids[ghost_right], 0)])
        got = exchange(rank, world, msg_l, msg_r)
        local, local_ids = [owned], [ids]  # stayers are owned, emigrants stay as ghost copies: the new key tells
        for blob in got:
            if blob is not None and blob.size:
                r = blob.view(REC)
                local.append(r["p"].copy())
                local_ids.append(r["id"].copy())
        local, local_ids = np.concatenate(local), np.concatenate(local_ids)
        # 3. the ordinary sub-step on owned + ghosts with the GLOBAL grid
        pl.particles_count = local.size
        located = O.locate_in_grid(local, pl)
        srt, perm = O.sort_particles(located) if local.size >= 128 else (located[np.argsort(located["grid_index"], kind="stable")],
                                                                           np.argsort(located["grid_index"], kind="stable"))
        srt_ids = local_ids[perm]
        table = O.cell_table(srt, pl.grid_cell_count)
        dens, _, _ = O.density_pressure(srt, pl, terms, table)
        frc = O.forces(dens, pl, terms, table)
        adv, _ = O.advection_collision(frc, pl, scene)
        kx = np.array([O.morton_decode(int(k))[0] for k in srt["grid_index"]], dtype=np.int64)
        is_owned = (kx >= own_lo) & (kx < own_hi)
        owned, ids = adv[is_owned].copy(), srt_ids[is_owned].copy()
        # 4. the same step on one rank
        r1 = O.step(single, pg, terms, scene)
        single, single_ids = r1.particles, single_ids[r1.permutation]
        # 5. gather and compare by id
        blobs = [None] * world
        dist.all_gather_object(blobs, (owned, ids))
        if rank == 0:
            allp = np.concatenate([b[0] for b in blobs])
            alli = np.concatenate([b[1] for b in blobs])
            assert np.array_equal(np.sort(alli), np.arange(n, dtype=np.uint32)), "ids lost or duplicated"
            a, b = np.empty(n, dtype=abi.PARTICLE), np.empty(n, dtype=abi.PARTICLE)
            a[alli], b[single_ids] = allp, single
            assert np.array_equal(a["grid_index"], b["grid_index"]), "keys differ"
            for f in ("position", "velocity", "intermediate_velocity", "density", "pressure"):
                x, y = a[f].astype(np.float64), b[f].astype(np.float64)
                err = np.abs(x - y).max() / np.abs(y).max()
                assert err <= 2e-5, (step, f, err)
    total = torch.tensor([migrated_total])
    dist.all_reduce(total)
    if rank == 0:
        assert int(total.item()) > 0, "the scenario was meant to make particles migrate"
        print("PROTOCOL_OK migrated=%d" % int(total.item()), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
