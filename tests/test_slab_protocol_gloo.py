"""world_size-2 (and 3) CPU runs of the slab protocol model over gloo (tests/slab_protocol_worker.py)
plus unit tests of the host-side slab arithmetic."""
import os
import subprocess
import sys

import numpy as np
import pytest

from libclsph_b200 import slabs, workloads
from tests import helpers as H


@pytest.mark.parametrize("world", [2, 3])
def test_protocol_over_gloo(world):
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(29700 + world), os.path.join(H.ROOT, "tests", "slab_protocol_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=H.ROOT, env=env)
    sys.stdout.write(r.stdout[-2000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0 and "PROTOCOL_OK" in r.stdout


def test_equal_count_planes_and_ownership():
    rng = np.random.default_rng(0)
    x = rng.normal(size=10001).astype(np.float32)
    for world in (1, 2, 4, 8):
        planes = slabs.equal_count_planes(x, world)
        assert planes.size == world + 1 and np.isneginf(planes[0]) and np.isposinf(planes[-1])
        owner = slabs.slab_of(x, planes)
        counts = np.bincount(owner, minlength=world)
        assert counts.sum() == x.size and counts.max() - counts.min() <= 1


def test_cell_planes_snap_to_nearest_boundary():
    planes = np.array([-np.inf, 0.26, 1.0, np.inf], dtype=np.float32)
    assert slabs.cell_planes(planes, 0.0, 0.1, 20) == [0, 3, 10, 0x7FFFFFFF]
    assert slabs.cell_planes(planes, 0.0, 0.1, 8) == [0, 3, 8, 0x7FFFFFFF]  # clamped to the grid


def test_slab_state_generation_matches_the_full_state():
    p, terms, vol, _ = workloads.make_config(fluid="water", particles_count=20000)
    full = workloads.jittered_state(p, vol)
    seen = 0
    for world in (3,):
        for r in range(world):
            idx, planes = workloads.slab_indices(p, vol, r, world)
            sub = workloads.jittered_state(p, vol, index=idx)
            assert sub.tobytes() == full[idx].tobytes()
            x = sub["position"][:, 0]
            assert (x >= planes[r]).all() and (x < planes[r + 1]).all()
            seen += idx.size
    assert seen == 20000
