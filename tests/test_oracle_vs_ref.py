"""Live comparison of the oracle with oracle/_ref (the reference's own kernels and host code
compiled by oracle/build_ref.sh). Skipped where _ref is absent and cannot be built; the
committed fixtures (tests/test_golden.py) carry the same evidence everywhere else."""
import numpy as np
import pytest

from libclsph_b200 import workloads
from oracle import oracle as O, ref as R
from tests import helpers as H

pytestmark = pytest.mark.skipif(not (R.available() or R.can_build()), reason="oracle/_ref not built and no reference tree")

EXACT_FIELDS = ("position", "velocity", "intermediate_velocity", "density", "pressure", "grid_index")


@pytest.fixture(scope="module", autouse=True)
def _built():
    assert R.build()


@pytest.mark.parametrize("fluid,n,scene_name,steps", [("water", 4096, "box.obj", 4), ("mucus", 3000, "cone.obj", 2),
                                                      ("water", 640, "cube.obj", 3)])
def test_simulate_matches_oracle_bit_for_bit(fluid, n, scene_name, steps):
    p, terms, vol, _ = workloads.make_config(fluid=fluid, particles_count=n, particle_mass=0.05)
    nrm, v, i = R.scene_load(H.ROOT, scene_name)
    scene = O.Scene(v, i, nrm)
    s = workloads.jittered_state(p, vol, seed=n)
    states, p_after, _ = R.simulate(p, terms, vol, scene, initial=s, substeps=steps, record_all=True)
    po = p.copy()
    cur = s
    for k in range(steps):
        cur = O.step(cur, po, terms, scene, taps=False).particles
        for f in EXACT_FIELDS:
            assert np.array_equal(cur[f], states[k][f]), (k, f)
    assert H.struct_bytes(po) == H.struct_bytes(p_after)


@pytest.mark.parametrize("kind", H.EDGE_KINDS)
def test_edge_states_match_bit_for_bit(kind):
    """States that sit ON the decisions of the path (coincident particles and erratum E3, pairs at
    distance h, positions on cell boundaries, one crowded cell, isolated particles): the oracle must
    follow the reference's own kernels there too."""
    p, terms, vol, _ = workloads.make_config(fluid="water", particles_count=1024, particle_mass=0.05)
    nrm, v, i = R.scene_load(H.ROOT, "box.obj")
    scene = O.Scene(v, i, nrm)
    s = H.edge_state(kind, p, vol)
    states, p_after, _ = R.simulate(p, terms, vol, scene, initial=s, substeps=2, record_all=True)
    po = p.copy()
    cur = s
    for k in range(2):
        cur = O.step(cur, po, terms, scene, taps=False).particles
        for f in EXACT_FIELDS:
            assert np.array_equal(cur[f], states[k][f], equal_nan=True), (kind, k, f)
    assert H.struct_bytes(po) == H.struct_bytes(p_after)


def test_default_lattice_of_init_particles_matches():
    """Without last_frame.bin the reference places its own lattice (sph_simulation.cpp:71-92)."""
    p, terms, vol, _ = workloads.make_config(fluid="water", particles_count=1000)
    nrm, v, i = R.scene_load(H.ROOT, "box.obj")
    scene = O.Scene(v, i, nrm)
    states, _, _ = R.simulate(p, terms, vol, scene, initial=None, substeps=1)
    for lattice in (O.init_particles(p, vol), workloads.lattice_state(p, vol)):
        got = O.step(lattice, p.copy(), terms, scene, taps=False).particles
        for f in EXACT_FIELDS:
            assert np.array_equal(got[f], states[f]), f


def test_single_kernels_match():
    p, terms, vol, _ = workloads.make_config(fluid="mucus", particles_count=2500)
    s = workloads.jittered_state(p, vol, seed=5)
    O.bounds_and_grid(s, p)
    a = O.locate_in_grid(s, p)
    assert np.array_equal(a["grid_index"], R.kernel_locate_in_grid(s, p)["grid_index"])
    srt, _ = O.sort_particles(a)
    table = O.cell_table(srt, p.grid_cell_count)
    d_o, _, _ = O.density_pressure(srt, p, terms, table)
    d_r = R.kernel_density_pressure(srt, p, terms, table)
    assert d_o.tobytes() == d_r.tobytes()
    f_o = O.forces(d_o, p, terms, table)
    f_r = R.kernel_forces(d_o, p, terms, table)
    assert f_o.tobytes() == f_r.tobytes()
    nrm, v, i = R.scene_load(H.ROOT, "shower.obj")
    scene = O.Scene(v, i, nrm)
    a_o, _ = O.advection_collision(f_o, p, scene, max_iters=0)
    a_r = R.kernel_advection_collision(f_o, p, terms, scene)
    for f in EXACT_FIELDS + ("acceleration",):
        assert np.array_equal(a_o[f], a_r[f]), f
