"""Multi-GPU slab decomposition on real GPUs (needs >= 2; skipped otherwise): launches
tests/dist_worker.py under torchrun and checks its verdict."""
import os
import subprocess
import sys

import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def gpu_count():
    from libclsph_b200 import capi
    return capi.load_library().clsph_device_count()


@pytest.mark.parametrize("world,n", [(2, 60000), (4, 120000)])
def test_slab_decomposition_matches_single_gpu(world, n):
    if gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(H.ROOT, "tests", "dist_worker.py"),
           str(n), "4", "--no-sub"]  # the established organisation (cells of side 2h); the default one: test_zz_gpu_new_paths.py
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=H.ROOT)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0 and "DIST_OK" in r.stdout
