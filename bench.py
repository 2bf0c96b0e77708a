#!/usr/bin/env python
"""bench.py -- particle-steps/s of the SPH sub-step (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME]

One "step" = one sub-step = one simulate_single_frame equivalent (libclsph/sph_simulation.cpp:
173-344 in the reference). Default workload at N=1: BASELINE config 2 (water dam break in
box.obj, 1 Mi particles, jittered-lattice state S1). Prints ONE JSON line (rank 0).

  value      device-resident throughput: state in HBM, K sub-steps timed with CUDA events on the
             library's stream, max over ranks
  e2e        same metric through clsph_simulate_single_frame with pinned HOST buffers: every
             step uploads the 80-byte AoS array, steps, and downloads it (the reference's
             call shape when a callback is installed)
  roofline   dominant kernel: algorithmic bytes per launch / its CUDA-event duration, against the
             measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own kernels (oracle/_ref, built from the reference sources) or the
             oracle port, on the host cores, on a bounded sample of the same workload

--impl reference times only the CPU side (reference arm of the driver).

The library's default kernel organisation is what gets timed (sub-cell order, merged rows, face grid);
--option name=value overrides single options for tuning runs. With N > 1 the run starts with a bitwise parity
check of the slab decomposition against a single-GPU run on the job's own ranks (libclsph_b200.distcheck,
`multi_gpu_parity` in the JSON line). `repeats` holds four more timed regions of K sub-steps and the median.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# configurations whose particle count is the TOTAL of the job (BASELINE config 4: 16 Mi over 8 GPUs); every other
# configuration is per GPU when N > 1 (weak scaling: N x the configured count as one block)
FIXED_TOTAL = {"config4_river_16m"}
METRIC = "particle_steps_per_sec"
UNIT = "particle-steps/s"
# Algorithmic bytes per particle-step (SURVEY 8d / DESIGN.md): 256 + 16 P, P = radix passes.
def step_bytes(passes):
    return 256 + 16 * passes
# Algorithmic bytes per particle of each kernel as built (DESIGN.md "Kernels"):
KERNEL_BYTES = {
    "keys": 20.0,        # read pos 16, write key 4
    "sort": None,        # 4 + 16 P, filled at run time
    "reorder": 104.0,    # perm 4 + gather 48 + write 48 + sorted key 4
    "density": 24.0,     # read pos 16, write rho,p 8
    "forces": 56.0,      # read pos 16 + vel 16 + rho,p 8, write acceleration 16
    "integrate": 96.0,   # read pos, ivel, acceleration 48, write pos, vel, ivel 48
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="config2_dambreak_1m")
    ap.add_argument("--particles", type=int, default=0, help="override the particle count of the config")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=102400,
                    help="particles of the cpu_baseline sample; given explicitly it also bounds the --impl reference run")
    ap.add_argument("--option", action="append", default=[], help="name=value passed to clsph_set_option (tuning)")
    ap.add_argument("--repeats", type=int, default=4, help="further timed regions of K sub-steps after the one `value` is taken from")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the bitwise multi-GPU parity check before timing")
    ap.add_argument("--no-large-point", action="store_true",
                    help="N = 1, default configuration: skip the extra 16 Mi-particle point (`large_point`), whose state does not fit L2")
    args = ap.parse_args()
    args.cpu_sample_given = any(a == "--cpu-sample" or a.startswith("--cpu-sample=") for a in sys.argv[1:])
    return args


def measured_traffic(config, world, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture, when this run matches the capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as fh:
            t = json.load(fh)
        if world != 1 or t["config"] != config or kernel not in t["kernels"]:
            return None, None
        k = t["kernels"][kernel]
        return (k["read_mb"] + k["write_mb"]) * 1e6, "ncu --set full, %s (profiles/r02_traffic.json)" % k["capture"]
    except Exception:
        return None, None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                mask = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def sample_workload(args, n_override=None):
    """(params, terms, volume, scene_file, state) for the configured workload."""
    from libclsph_b200 import workloads
    fluid, n, mass, scene = workloads.CONFIGS[args.config]
    if args.particles:
        n = args.particles
    if n_override:
        n = n_override
    p, terms, vol, _ = workloads.make_config(fluid=fluid, particles_count=n, particle_mass=mass)
    state = workloads.jittered_state(p, vol)
    return p, terms, vol, scene, state


# ------------------------------------------------------------------------------------------------
# CPU side: the reference's own kernels (oracle/_ref) or the oracle port
# ------------------------------------------------------------------------------------------------
def cpu_run(args, n, substeps, warmup):
    """Times `substeps` sub-steps of n particles on the host cores. Returns (rate, kind, cores)."""
    from oracle import oracle as O, ref as R
    p, terms, vol, scene_file, state = sample_workload(args, n_override=n)
    scene = O.load_obj(os.path.join(ROOT, "scenes", scene_file))
    # every host core this process may run on: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    # make the CPU side 1 / cores of what the box can do
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if R.available():
        R.set_num_threads(cores)
        _, _, secs = R.simulate(p, terms, vol, scene, initial=state, substeps=warmup + substeps, record_all=False)
        elapsed = float(secs[warmup:].sum())
        return n * substeps / elapsed, "reference", R.num_threads(), elapsed
    O.set_num_threads(cores)
    cur = state
    po = p.copy()
    for _ in range(warmup):
        cur = O.step(cur, po, terms, scene, taps=False).particles
    t0 = time.perf_counter()
    for _ in range(substeps):
        cur = O.step(cur, po, terms, scene, taps=False).particles
    elapsed = time.perf_counter() - t0
    return n * substeps / elapsed, "port", O.num_threads(), elapsed


def cpu_baseline(args, budget_s=20.0):
    """Bounded CPU sample: calibrate on 2 sub-steps, then run as many as fit the budget."""
    n = min(args.cpu_sample, sample_workload(args)[0].particles_count)
    rate, kind, cores, el = cpu_run(args, n, 2, 1)
    steps = int(max(2, min(20, budget_s * rate / n)))
    rate, kind, cores, el = cpu_run(args, n, steps, 1)
    what = ("reference kernels + host code compiled from the reference sources behind an in-process OpenCL shim "
            "(oracle/_ref, OpenMP over work-groups; PoCL unavailable)" if kind == "reference"
            else "CPU restatement of the reference kernels (oracle port, OpenMP; PoCL unavailable)")
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%s at %d particles (state S1), %d sub-steps after 1 warm-up, %.1f s; %s" % (args.config, n, steps, el, what)}


def run_reference(args, rank):
    if rank != 0:
        return
    fluid, n_full, mass, scene = __import__("libclsph_b200.workloads", fromlist=["CONFIGS"]).CONFIGS[args.config]
    n_full = args.particles or n_full
    if not (args.config in FIXED_TOTAL and not args.particles):
        n_full *= max(1, args.gpus)  # the GPU arm runs gpus x the configured count as one block
    # The configured particle count itself when K+W sub-steps of it fit about four minutes on this box's cores
    # (config 2 on 16 cores: ~2 s per sub-step); otherwise a bounded sample of the same fluid, sized to that budget
    # (same_config false). --cpu-sample N forces a sample.
    n_cal = min(65536, n_full)
    rate, kind, cores, _ = cpu_run(args, n_cal, 2, 1)
    total = max(1, args.steps + args.warmup)
    budget_n = int(rate * 240.0 / total)
    n = n_full if (budget_n >= n_full and not args.cpu_sample_given) else int(min(n_full, args.cpu_sample if args.cpu_sample_given else n_full,
                                                                                  max(4096, budget_n)))
    if n != n_full:
        n -= n % 4096 if n >= 4096 else 0
    t_wall = time.perf_counter()
    rate, kind, cores, elapsed = cpu_run(args, n, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.config, "particles_per_step_sample": n, "particles_full": n_full, "same_config": n == n_full,
                   "state": "S1 jittered lattice, seed 20261017",
                   "note": ("each step is one sub-step of the configured workload at its full particle count" if n == n_full else
                            "each step is a bounded sample of the workload (same fluid, same spacing)")},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d particles x %d sub-steps, %.1f s wall" % (n, args.steps, time.perf_counter() - t_wall)},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
_SAVED_STDOUT = None  # the real stdout while run_ours has fd 1 pointed at stderr


def ctx_capacity(ctx, n):
    """Room for a rank's download: its own particles plus what may have migrated in."""
    return int(n * 1.5) + 65536


def multi_gpu_workload(config, n_per_gpu, rank, world, sub_cell_order=True):
    """This rank's share of the weak-scaling workload: `world` x the configured count as ONE fluid block
    (same particle mass, hence same h and spacing), cut into x slabs. Returns a dict with the parameters,
    the rank's particles and ids, its slab planes and the capacities to create the context with."""
    from libclsph_b200 import workloads
    fluid, _, mass, _ = workloads.CONFIGS[config]
    p, terms, vol, _ = workloads.make_config(fluid=fluid, particles_count=n_per_gpu * world, particle_mass=mass)
    index, planes = workloads.slab_indices(p, vol, rank, world)
    state = workloads.jittered_state(p, vol, index=index)
    n = state.size
    # one grid-cell layer of the block's cross-section: the unit of ghost traffic (two layers per side)
    # and of bursty migration (a snapped slab boundary jumps by one cell now and then)
    per_side, side, spacing = workloads.lattice_geometry(p, vol)
    layer = int(p.particles_count / per_side * (2.0 * p.h / float(spacing))) + 1
    # Messages have a fixed size (no host-visible counts), so the capacities are what travels every sub-step:
    # emigrants = one layer in a burst (established kernels: slab boundaries snapped to cells), ghosts = two
    # cell layers (established) or the 2h next to the plane (sub-cell order), with 50 % slack for compression.
    # Sub-cell order: ownership follows the planes themselves, so only the particles that cross one migrate (at
    # most vmax dt = 0.08 h of a layer per sub-step) and there are no bursts.
    emigrant_cap = int((0.25 if sub_cell_order else 1.5) * layer) + 8192
    ghost_cap = int((1.5 if sub_cell_order else 3.0) * layer) + 8192
    return dict(params=p, terms=terms, volume=vol, state=state, ids=index, planes=planes, emigrant_capacity=emigrant_cap,
                ghost_capacity=ghost_cap, capacity=int(1.2 * n) + 2 * layer + 2 * ghost_cap + 65536)


def bind_to_gpu_numa(local_rank):
    """Keeps this rank's threads (and with them its pinned host buffers, first touch) on the CPUs next to its GPU:
    with eight ranks on one box the end-to-end copies otherwise all cross to one NUMA node."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        near = {i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1}
        allowed = set(os.sched_getaffinity(0)) & near
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return None


def large_point(capi, workloads, torch, local_rank, peak, config="sweep_16m", steps=10, warmup=3):
    """Device-resident particle-steps/s of `config` (state in HBM, CUDA events on the library's stream)."""
    fluid, n_cfg, mass, scene_file = workloads.CONFIGS[config]
    p, terms, vol, _ = workloads.make_config(fluid=fluid, particles_count=n_cfg, particle_mass=mass)
    state = workloads.jittered_state(p, vol)
    ctx = capi.Context(state.size, device=local_rank)
    try:
        ctx.set_scene(*workloads.scene_arrays(scene_file))
        ctx.set_parameters(p, terms)
        stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))
        ctx.upload(state)
        ctx.step(warmup)
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.step(steps)
        e1.record(stream)
        ctx.synchronize()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        grid = ctx.parameters()
    finally:
        ctx.close()
    passes = max(1, (int(grid.grid_cell_count * 8 - 1).bit_length() + 7) // 8)
    value = state.size * steps / (ms * 1e-3)
    return {"workload": config, "particles": int(state.size), "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "value": value,
            "unit": UNIT, "algorithmic_gb_s": value * step_bytes(passes) / 1e9, "frac_of_measured_hbm": value * step_bytes(passes) / 1e9 / peak,
            "note": "per-step working set ~3 GB vs 126 MB L2"}


def run_ours(args, rank, world, local_rank):
    # stdout must carry exactly one JSON line: native libraries (NCCL's version banner) write to fd 1 too,
    # so everything goes to stderr until the line is ready
    sys.stdout.flush()
    global _SAVED_STDOUT
    saved_stdout = _SAVED_STDOUT = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch
    from libclsph_b200 import capi, workloads

    dist = None
    near_cpus = bind_to_gpu_numa(local_rank) if world > 1 else None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    fluid, n_cfg, mass, scene_file = workloads.CONFIGS[args.config]
    n_cfg = args.particles or n_cfg
    fixed_total = args.config in FIXED_TOTAL and not args.particles
    if fixed_total and world > 1:
        n_cfg //= world
    options = list(args.option)
    sub = "sub_cell_order=0" not in options
    normals, vertices, indices = workloads.scene_arrays(scene_file)
    if world == 1:
        p, terms, vol, _, state = sample_workload(args)
        ids = None
        n = state.size
        ctx = capi.Context(n, device=local_rank)
    else:
        # bitwise parity of the slab decomposition on this job's own ranks, before anything is timed
        parity = None
        if not args.no_parity and sub:
            from libclsph_b200 import distcheck
            parity = distcheck.bitwise_parity(dist, rank, world, local_rank, n_total=60000 * world, steps=4, options=options)
        # weak scaling: each rank generates only its own slab of the common block
        w = multi_gpu_workload(args.config, n_cfg, rank, world, sub_cell_order=sub)
        p, terms, vol, state, ids, planes = w["params"], w["terms"], w["volume"], w["state"], w["ids"], w["planes"]
        emigrant_cap, ghost_cap = w["emigrant_capacity"], w["ghost_capacity"]
        n = state.size
        ctx = capi.Context(w["capacity"], device=local_rank)
    for opt in options:
        k, v = opt.split("=")
        ctx.set_option(k, int(v))
    ctx.set_scene(normals, vertices, indices)
    ctx.set_parameters(p, terms)
    if world > 1:
        if rank == 0:
            uid = torch.tensor(list(capi.comm_unique_id()), dtype=torch.uint8, device=device)
        else:
            uid = torch.zeros(128, dtype=torch.uint8, device=device)
        dist.broadcast(uid, 0)
        ctx.dist_init(rank, world, bytes(uid.cpu().numpy().tolist()), float(planes[rank]), float(planes[rank + 1]),
                      emigrant_capacity=emigrant_cap, ghost_capacity=ghost_cap)
    transport = ctx.dist_transport() if world > 1 else "none"
    stream = torch.cuda.ExternalStream(ctx.stream(), device=device)

    def upload_state():
        if world == 1:
            ctx.upload(state)
        else:
            ctx.dist_upload(state, ids)

    # ---- device-resident throughput ("value")
    upload_state()
    ctx.step(args.warmup)
    ctx.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.profile_enable(False)  # resets the launch counter
    barrier()
    torch.cuda.synchronize()
    def timed_region():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.step(args.steps)
        e1.record(stream)
        ctx.synchronize()
        torch.cuda.synchronize()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    ms_total = timed_region()
    launches = ctx.profile_read()["kernel_launches"]
    # BASELINE.md section 3 asks for the median of 5: four more regions of K sub-steps on the evolving state
    regions = [ms_total] + [timed_region() for _ in range(max(0, args.repeats))]
    clocks = sampler.finish()
    n_total = sum_over_ranks(float(n))
    value = n_total * args.steps / (ms_total * 1e-3)
    grid = ctx.parameters()
    passes = max(1, (int(grid.grid_cell_count - 1).bit_length() + 7) // 8)

    # ---- per-kernel durations over a second timed region (stage events on the same stream)
    ctx.profile_enable(True)
    ctx.step(args.steps)
    stage = ctx.profile_read()
    ctx.profile_enable(False)
    ctx.synchronize()  # surfaces device-side errors (grid or buffer overflow) of the profiled steps here
    per = {k[3:]: stage[k] / max(1, stage["substeps"]) for k in stage if k.startswith("ms_")}
    kb = dict(KERNEL_BYTES)
    kb["sort"] = 4.0 + 16.0 * passes
    dominant = max((k for k in kb), key=lambda k: per.get(k, 0.0))
    peak, peak_src = measured_peak()
    achieved = n * kb[dominant] / (per[dominant] * 1e-3) / 1e9
    pairs = sub and "pair_density=0" not in options and ("pair_density=1" in options or n >= 160000)  # the library's size rule
    factored = sub and "factored_forces=0" not in options and "fast_pairs=0" not in options
    direct = sub and "fast_pairs=0" not in options and os.environ.get("CLSPH_FORCES_DIRECT", "1") != "0"
    kernel_name = {"density": ("k_density_pairs" if pairs else "k_density_sub") if sub else "k_density_lists",
                   "forces": "k_forces_lists_direct" if direct else ("k_forces_lists_factored" if factored else "k_forces_lists"),
                   "reorder": "k_reorder_sub" if sub else "k_reorder", "sort": "k_onesweep", "keys": "k_keys_hist",
                   "integrate": "k_integrate"}[dominant]
    # the committed ncu capture is of the default options
    traffic, traffic_src = (None, None) if options else measured_traffic(args.config if not args.particles else None, world, kernel_name)
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch", "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": n * kb[dominant], "peak_source": peak_src,
                "algorithmic_bytes_per_particle": kb[dominant], "ms_per_launch": per[dominant],
                "whole_step": {"bytes_per_particle_step": step_bytes(passes), "radix_passes": passes,
                               "achieved": value / world * step_bytes(passes) / 1e9,
                               "frac": value / world * step_bytes(passes) / 1e9 / peak,
                               "frac_of_nominal_8TBs": value / world * step_bytes(passes) / 1e9 / 8000.0},
                "stage_ms": per,
                # what the library's sort did in the last sub-step: 0 = counting sort on the sub-cell table, else 8-bit radix passes
                # (the algorithmic bytes above keep SURVEY 8d's figure, 16 bytes per radix pass the keys need)
                "sort_passes_run": ctx.sort_passes(),
                "note": "density/force passes are instruction-issue / L1-gather / latency bound at the reference's 2h cell geometry, not HBM bound (SURVEY 8d, DESIGN 7)"}

    # ---- end to end through the host-buffer calls, pinned host memory: every step uploads the 80-byte AoS
    # array, runs one sub-step and downloads the result (the reference's call shape when a callback is installed)
    import ctypes
    lib = capi.load_library()
    if args.e2e_steps > 0:
        host_in = torch.empty(n * 80, dtype=torch.uint8).pin_memory()
        host_out = torch.empty(ctx_capacity(ctx, n) * 80, dtype=torch.uint8).pin_memory()
        host_in.numpy()[:] = state.view(np.uint8).reshape(-1)
        p_io = p.copy()
        if world > 1:
            ids_in = torch.from_numpy(ids.astype(np.uint32)).pin_memory()
            ids_out = torch.empty(ctx_capacity(ctx, n), dtype=torch.int32).pin_memory()
            got = ctypes.c_uint32()

        def e2e_step():
            if world == 1:
                rc = lib.clsph_simulate_single_frame(ctx._h, ctypes.c_void_p(host_in.data_ptr()),
                                                     ctypes.c_void_p(host_out.data_ptr()), ctypes.byref(p_io), ctypes.byref(terms))
            else:
                rc = lib.clsph_dist_upload(ctx._h, ctypes.c_void_p(host_in.data_ptr()), ctypes.c_void_p(ids_in.data_ptr()), n)
                rc = rc or lib.clsph_step(ctx._h, 1)
                rc = rc or lib.clsph_dist_download(ctx._h, ctypes.c_void_p(host_out.data_ptr()), ctypes.c_void_p(ids_out.data_ptr()),
                                                   ctx_capacity(ctx, n), ctypes.byref(got))
            if rc:
                raise RuntimeError(lib.clsph_last_error(ctx._h).decode())

        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        dt = max_over_ranks(time.perf_counter() - t0)
        zero_copy = world == 1 and os.environ.get("CLSPH_ZERO_COPY_UPLOAD", "0") not in ("0", "")
        e2e = {"value": n_total * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": n * ((48 if zero_copy else 80) if world == 1 else 84),
               "d2h_bytes_per_step": n * (80 if world == 1 else 84), "steps": args.e2e_steps, "ms_per_step": 1e3 * dt / args.e2e_steps,
               "path": ("clsph_simulate_single_frame(host AoS in, host AoS out)" if world == 1 else
                        "clsph_dist_upload + clsph_step + clsph_dist_download per rank") + ", pinned buffers",
               "cpus_near_gpu": near_cpus,
               "h2d": ("the upload kernel reads position, velocity and half-step velocity (48 of the 80 bytes of a record; ~64 with "
                       "32-byte sectors) straight from the pinned host array, inside the timed call" if zero_copy else
                       "cudaMemcpyAsync of the 80-byte records into a staging area, then a conversion kernel")}
    else:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "steps": 0,
               "path": "skipped (--e2e-steps 0)"}
    ctx.close()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong" if fixed_total else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.config, "particles_per_gpu": n, "particles_total": int(n_total), "fluid": fluid,
                   "scene": scene_file, "state": "S1 jittered lattice, seed 20261017",
                   "parallelism": "single GPU" if world == 1 else
                   "one fluid block of %d x the configured count, slab-decomposed along x over %d GPUs; per sub-step: global AABB, "
                   "migration + ghosts within 2h of each plane; transport: %s" % (world, world, transport),
                   "l2": "per-step working set ~%d MB vs 126 MB L2, no flush: sub-steps form a dependent chain" % (n * 200 // (1 << 20)),
                   "grid": [grid.grid_size_x, grid.grid_size_y, grid.grid_size_z], "grid_cell_count": grid.grid_cell_count,
                   "options": options},
        "repeats": {"ms_per_step": [m / args.steps for m in regions],
                    "median_value": n_total * args.steps / (sorted(regions)[len(regions) // 2] * 1e-3)},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
    }
    if world > 1:
        line["multi_gpu_parity"] = parity if parity is not None else {"skipped": True}
    # The configured 1 Mi particles keep ~50 MB of state, which stays in the 126 MB L2 between kernels; a second, shorter
    # measurement of the same fluid at 16 Mi particles (uniform block over plane.obj, config 5) shows the rate when
    # every pass really streams from HBM. Reported next to the headline, never instead of it.
    if world == 1 and args.config == "config2_dambreak_1m" and not args.particles and not options and not args.no_large_point:
        try:
            line["large_point"] = large_point(capi, workloads, torch, local_rank, peak)
        except Exception as exc:
            line["large_point"] = {"value": None, "error": repr(exc)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline(args)
        except Exception as exc:  # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (exc,)}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    # (NCCL_DEBUG is left as the caller set it: run_ours points fd 1 at stderr while native libraries may print,
    # so NCCL's communicator log cannot reach the JSON line on stdout)
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it, one rank per GPU
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    try:
        run_ours(args, rank, world, local_rank)
    except BaseException:
        # A failed rank must not linger: its peers would wait in NCCL forever. Report and leave at once,
        # without destructors that synchronise streams with unmatched receives on them.
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)


if __name__ == "__main__":
    main()
