"""Generates tests/golden/*.npz from the reference's own sources (oracle/_ref, built by
oracle/build_ref.sh from /root/reference). Runs only in the build container; the fixtures are
committed so the CPU and GPU suites can check against the reference where the tree is absent.

    python tests/golden/make_golden.py
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from libclsph_b200 import abi, workloads  # noqa: E402
from oracle import oracle as O, ref as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SCENES = ["box.obj", "cone.obj", "cube.obj", "labyrinth.obj", "monkey.obj", "plane.obj", "river.obj", "shower.obj"]


def struct_bytes(s):
    return np.frombuffer(ctypes.string_at(ctypes.addressof(s), ctypes.sizeof(s)), dtype=np.uint8).copy()


def ref_scene(name):
    n, v, i = R.scene_load(ROOT, name)
    return O.Scene(v, i, n)


def case(name, fluid, n, mass, scene_name, state_fn, substeps):
    p, terms, vol, _ = workloads.make_config(fluid=fluid, particles_count=n, particle_mass=mass)
    scene = ref_scene(scene_name)
    state = state_fn(p, vol)
    states, p_after, _ = R.simulate(p, terms, vol, scene, initial=state, substeps=substeps, record_all=True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), params=struct_bytes(p), terms=struct_bytes(terms),
                        volume=np.float32(vol), scene=np.array(scene_name), initial=state, states=states,
                        params_after=struct_bytes(p_after))
    print(name, "n=%d" % n, "steps=%d" % substeps, "grid", p_after.grid_size_x, p_after.grid_size_y, p_after.grid_size_z)


def drop(floor_y, speed):
    def fn(p, vol):
        s = workloads.jittered_state(p, vol, seed=7)
        s["position"][:, 1] += np.float32(floor_y + 0.002 - s["position"][:, 1].min())
        s["intermediate_velocity"][:, 1] = np.float32(-speed)
        s["velocity"][:, 1] = np.float32(-speed)
        return s
    return fn


def main():
    assert R.build(), "oracle/_ref could not be built (no reference tree?)"
    # 1. settings as the reference's load_settings derives them from the shipped JSON files
    out = {}
    for fluid in ("water", "mucus"):
        p, t, vol, flags = R.load_settings(os.path.join(ROOT, "fluid_properties", fluid + ".json"),
                                           os.path.join(ROOT, "simulation_properties", "default.json"))
        out[fluid + "_params"] = struct_bytes(p)
        out[fluid + "_terms"] = struct_bytes(t)
        out[fluid + "_volume"] = np.float32(vol)
    np.savez_compressed(os.path.join(HERE, "load_settings.npz"), **out)
    # 2. scenes as the reference's scene::load (tinyobj + face normals) returns them
    out = {}
    for name in SCENES:
        n, v, i = R.scene_load(ROOT, name)
        out[name + ":normals"], out[name + ":vertices"], out[name + ":indices"] = n, v, i
    np.savez_compressed(os.path.join(HERE, "scenes.npz"), **out)
    # 3. whole sub-steps through sph_simulation::simulate
    case("step_water_box_s1_n2048", "water", 2048, 0.05, "box.obj", lambda p, v: workloads.jittered_state(p, v), 3)
    case("step_water_box_s0_n1000", "water", 1000, 0.05, "box.obj", lambda p, v: workloads.lattice_state(p, v), 2)
    case("step_mucus_plane_drop_n1536", "mucus", 1536, 0.05, "plane.obj", drop(-1.0, 2.5), 3)
    case("step_mucus_labyrinth_n2048", "mucus", 2048, 0.05 * 32000 / 4194304 * 2048, "labyrinth.obj",
         lambda p, v: workloads.jittered_state(p, v), 2)
    # 4. the force kernel alone (its output, the acceleration, never leaves the step)
    p, terms, vol, _ = workloads.make_config(fluid="water", particles_count=2048, particle_mass=0.05)
    s = workloads.jittered_state(p, vol)
    O.bounds_and_grid(s, p)
    located = R.kernel_locate_in_grid(s, p)
    srt, perm = O.sort_particles(located)
    table = O.cell_table(srt, p.grid_cell_count)
    dens = R.kernel_density_pressure(srt, p, terms, table)
    frc = R.kernel_forces(dens, p, terms, table)
    np.savez_compressed(os.path.join(HERE, "kernel_forces_water_n2048.npz"), params=struct_bytes(p), terms=struct_bytes(terms),
                        volume=np.float32(vol), initial=s, permutation=perm, cell_table=table,
                        density=dens["density"], pressure=dens["pressure"], acceleration=frc["acceleration"][:, :3].copy())
    print("kernel_forces_water_n2048 done")


if __name__ == "__main__":
    main()
