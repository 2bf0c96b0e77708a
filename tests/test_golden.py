"""Golden vectors produced by the reference's own sources (tests/golden/make_golden.py, which
runs oracle/_ref = libclsph's .cl kernels and host code compiled behind an OpenCL shim).

CPU part: the oracle and the product's host-side derivations reproduce them bit for bit.
GPU part (-m gpu): the CUDA path matches them (integers bit-exact, fp32 within 1e-4)."""
import ctypes
import glob
import os

import numpy as np
import pytest

from libclsph_b200 import abi, workloads
from oracle import oracle as O
from tests import helpers as H

GOLDEN = os.path.join(H.ROOT, "tests", "golden")
STEP_CASES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "step_*.npz")))
EXACT_FIELDS = ("position", "velocity", "intermediate_velocity", "density", "pressure", "grid_index")


def from_bytes(cls, arr):
    obj = cls()
    ctypes.memmove(ctypes.addressof(obj), arr.tobytes(), ctypes.sizeof(obj))
    return obj


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    p = from_bytes(abi.SimulationParameters, z["params"])
    terms = from_bytes(abi.PrecomputedKernelValues, z["terms"])
    scene = O.load_obj(os.path.join(H.ROOT, "scenes", str(z["scene"])))
    return z, p, terms, scene


def test_fixtures_present():
    assert len(STEP_CASES) >= 4


@pytest.mark.parametrize("fluid", ["water", "mucus"])
def test_derived_constants_match_reference_load_settings(fluid):
    z = np.load(os.path.join(GOLDEN, "load_settings.npz"))
    # the oracle's restatement
    p, n_infl = abi.raw_parameters(fluid)
    terms, vol = O.derive_constants(p, n_infl)
    assert H.struct_bytes(p) == z[fluid + "_params"].tobytes()
    assert H.struct_bytes(terms) == z[fluid + "_terms"].tobytes()
    assert np.float32(vol) == z[fluid + "_volume"]
    # the product's host-side derivation
    p2, terms2, vol2, _ = workloads.make_config(fluid=fluid)
    assert H.struct_bytes(p2) == z[fluid + "_params"].tobytes()
    assert H.struct_bytes(terms2) == z[fluid + "_terms"].tobytes()
    assert np.float32(vol2) == z[fluid + "_volume"]


@pytest.mark.parametrize("scene_name", ["box.obj", "cone.obj", "cube.obj", "labyrinth.obj", "monkey.obj", "plane.obj",
                                        "river.obj", "shower.obj"])
def test_scene_triangles_and_normals_match_reference_scene_load(scene_name):
    z = np.load(os.path.join(GOLDEN, "scenes.npz"))
    ref_tri = z[scene_name + ":vertices"].reshape(-1, 3)[z[scene_name + ":indices"]].reshape(-1, 9)
    for loader in ("oracle", "product"):
        if loader == "oracle":
            sc = O.load_obj(os.path.join(H.ROOT, "scenes", scene_name))
            normals, vertices, indices = sc.face_normals, sc.vertices, sc.indices
        else:
            normals, vertices, indices = workloads.scene_arrays(scene_name)
        tri = vertices.reshape(-1, 3)[indices].reshape(-1, 9)
        assert np.array_equal(tri, ref_tri), loader        # same triangles, same order, same corners
        assert normals.tobytes() == z[scene_name + ":normals"].tobytes(), loader


@pytest.mark.parametrize("name", STEP_CASES)
def test_oracle_reproduces_reference_steps_bit_for_bit(name):
    z, p, terms, scene = load_case(name)
    cur = z["initial"]
    for k in range(z["states"].shape[0]):
        cur = O.step(cur, p, terms, scene, taps=False).particles
        for f in EXACT_FIELDS:
            assert np.array_equal(cur[f], z["states"][k][f]), "%s step %d field %s" % (name, k, f)
        assert not cur["acceleration"].any()
    assert H.struct_bytes(p) == z["params_after"].tobytes()


def test_oracle_force_kernel_matches_reference_bit_for_bit():
    z = np.load(os.path.join(GOLDEN, "kernel_forces_water_n2048.npz"))
    p = from_bytes(abi.SimulationParameters, z["params"])
    terms = from_bytes(abi.PrecomputedKernelValues, z["terms"])
    scene = O.Scene(np.zeros(0, np.float32), np.zeros(0, np.uint32), np.zeros(0, np.float32))
    r = O.step(z["initial"], p, terms, scene)
    assert np.array_equal(r.permutation, z["permutation"])
    assert np.array_equal(r.cell_table, z["cell_table"])
    assert np.array_equal(r.density, z["density"])
    assert np.array_equal(r.pressure, z["pressure"])
    assert np.array_equal(r.acceleration, z["acceleration"])


# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", STEP_CASES)
def test_cuda_matches_reference_steps(name):
    from libclsph_b200 import capi
    z, p, terms, scene = load_case(name)
    ctx = capi.Context(z["initial"].size)
    ctx.set_scene(scene.face_normals, scene.vertices, scene.indices)
    ctx.set_parameters(p, terms)
    # single step from each golden state (the 1e-4 bar is a single-step bar)
    prev = z["initial"]
    for k in range(z["states"].shape[0]):
        ctx.upload(prev)
        ctx.step(1)
        got = ctx.download()
        want = z["states"][k]
        assert np.array_equal(got["grid_index"], want["grid_index"]), "%s step %d keys/order" % (name, k)
        H.assert_close_fields(got, want, tol=1e-4, what="%s step %d" % (name, k))
        prev = want
    assert H.struct_bytes(ctx.parameters()) == z["params_after"].tobytes()
    ctx.close()


@pytest.mark.gpu
def test_cuda_force_kernel_matches_reference():
    from libclsph_b200 import capi
    z = np.load(os.path.join(GOLDEN, "kernel_forces_water_n2048.npz"))
    p = from_bytes(abi.SimulationParameters, z["params"])
    terms = from_bytes(abi.PrecomputedKernelValues, z["terms"])
    ctx = capi.Context(z["initial"].size)
    ctx.set_scene(np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros(0, np.uint32))
    ctx.set_parameters(p, terms)
    ctx.set_debug(True)
    ctx.upload(z["initial"])
    ctx.step(1)
    ctx.synchronize()
    assert np.array_equal(ctx.fetch(capi.TAP_PERMUTATION), z["permutation"])
    assert np.array_equal(ctx.fetch(capi.TAP_CELL_TABLE), z["cell_table"])
    assert H.rel_err(ctx.fetch(capi.TAP_DENSITY), z["density"]) <= 1e-4
    assert H.rel_err(ctx.fetch(capi.TAP_PRESSURE), z["pressure"]) <= 1e-4
    assert H.rel_err(ctx.fetch(capi.TAP_ACCELERATION), z["acceleration"]) <= 1e-4
    ctx.close()
