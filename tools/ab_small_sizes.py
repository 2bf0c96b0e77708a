"""A/B at small particle counts, one process, no torch: ms per sub-step (wall clock around a run of resident sub-steps) of the
default options against alternatives, for the water dam break at several sizes. Usage: python tools/ab_small_sizes.py [n ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libclsph_b200 import capi, workloads  # noqa: E402

SIZES = [int(a) for a in sys.argv[1:]] or [102400, 262144]
VARIANTS = [("default", {}), ("pair_density=0", {"pair_density": 0})]
STEPS = 400

for n in SIZES:
    mass = 0.05 * 102400 / n  # the fluid volume of config 1 at every size
    p, terms, vol, scene_file = workloads.make_config(fluid="water", particles_count=n, particle_mass=mass)
    state = workloads.jittered_state(p, vol)
    normals, vertices, indices = workloads.scene_arrays(scene_file)
    for name, options in VARIANTS + VARIANTS[:1]:
        ctx = capi.Context(n)
        for k, v in options.items():
            ctx.set_option(k, v)
        ctx.set_scene(normals, vertices, indices)
        ctx.set_parameters(p, terms)
        ctx.upload(state)
        ctx.step(20)
        ctx.synchronize()
        t0 = time.perf_counter()
        ctx.step(STEPS)
        ctx.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / STEPS
        print("n=%d %-16s %.4f ms per sub-step, sort passes %d" % (n, name, ms, ctx.sort_passes()), flush=True)
        ctx.close()
