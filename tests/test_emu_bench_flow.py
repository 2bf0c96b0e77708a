"""bench.py's GPU arm, single GPU, run end to end on the CPU: the library is the emulator build (tests/emu/),
torch is replaced by a stub with the handful of calls bench.py makes (events, the external stream, pinned
host buffers). Nothing here measures anything; the point is that every line of the benchmark's control flow --
option plumbing, the timed regions and their repeats, stage profiling, the end-to-end loop, the multi-GPU parity
check, the JSON line with its roofline -- executes before a real run on a B200."""
import io
import json
import sys
import time
import types
from contextlib import redirect_stdout

import numpy as np
import pytest

from libclsph_b200 import capi
from tests import helpers as H
from tests.emu import build_emu


@pytest.fixture(scope="module", autouse=True)
def emulator_library():
    saved = capi._lib
    capi._lib = capi.load_library(build_emu.build())
    yield capi._lib
    capi._lib = saved


class _Tensor:
    def __init__(self, arr):
        self.arr = arr

    def pin_memory(self):
        return self

    def numpy(self):
        return self.arr

    def data_ptr(self):
        return self.arr.ctypes.data


class _Event:
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return 1e3 * (other.t - self.t)


def fake_torch():
    t = types.ModuleType("torch")
    t.uint8, t.int32, t.float64, t.int64 = np.uint8, np.int32, np.float64, np.int64
    t.empty = lambda n, dtype=np.uint8: _Tensor(np.empty(n, dtype=dtype))
    t.from_numpy = lambda a: _Tensor(a)
    t.device = lambda *a: "cuda:0"
    cuda = types.ModuleType("torch.cuda")
    cuda.set_device = lambda d: None
    cuda.synchronize = lambda: None
    cuda.Event = _Event
    cuda.ExternalStream = lambda ptr, device=None: object()
    t.cuda = cuda
    return t, cuda


@pytest.mark.parametrize("options", [[], ["sub_cell_order=0"]])
def test_run_ours_single_gpu_prints_the_contract_line(options, monkeypatch, capfd):
    sys.path.insert(0, H.ROOT)
    import bench
    torch, cuda = fake_torch()
    monkeypatch.setitem(sys.modules, "torch", torch)
    monkeypatch.setitem(sys.modules, "torch.cuda", cuda)
    args = types.SimpleNamespace(gpus=1, steps=3, warmup=3, impl="ours", config="config3_mucus_labyrinth_4m", particles=1500, e2e_steps=2,
                                 no_cpu_baseline=True, cpu_sample=4096, option=list(options), repeats=2, no_parity=False, no_large_point=True)
    bench.run_ours(args, 0, 1, 0)
    out = capfd.readouterr().out
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out[-2000:]
    d = json.loads(lines[0])
    assert d["metric"] == "particle_steps_per_sec" and d["unit"] == "particle-steps/s" and d["n_gpus"] == 1
    assert d["steps"] == 3 and d["warmup"] == 3 and d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"]["workload"] == "config3_mucus_labyrinth_4m" and d["config"]["particles_per_gpu"] == 1500
    assert d["config"]["options"] == list(options)
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] in (1500 * 80, 1500 * 48) and d["e2e"]["d2h_bytes_per_step"] == 1500 * 80
    assert d["gpu_launches"] > 0
    assert len(d["repeats"]["ms_per_step"]) == 3 and d["repeats"]["median_value"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["peak"] > 0 and r["achieved"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert set(r["stage_ms"]) >= {"keys", "sort", "reorder", "density", "forces", "integrate"}
    assert r["kernel"] in ("k_density_pairs", "k_forces_lists_factored", "k_forces_lists_direct", "k_density_sub", "k_density_lists", "k_forces_lists", "k_integrate", "k_onesweep", "k_reorder_sub", "k_reorder",
                           "k_keys_hist")


# ---- N > 1: threads as ranks, a thread-based stand-in for torch.distributed ---------------------------------

import threading


class _DTensor(_Tensor):
    def item(self):
        return self.arr.reshape(-1)[0].item()

    def cpu(self):
        return self


class _FakeDist:
    """broadcast / all_reduce / barrier across the threads that play the ranks."""

    class ReduceOp:
        MAX, SUM = "max", "sum"

    def __init__(self, world):
        self.world = world
        self.local = threading.local()
        self.sync = threading.Barrier(world)
        self.slots = [None] * world

    def init_process_group(self, backend, device_id=None):
        pass

    def destroy_process_group(self):
        pass

    def barrier(self):
        self.sync.wait(timeout=900)

    def _exchange(self, arr):
        self.slots[self.local.rank] = arr.copy()
        self.sync.wait(timeout=900)
        got = [s.copy() for s in self.slots]
        self.sync.wait(timeout=900)
        return got

    def broadcast(self, t, src):
        t.arr[...] = self._exchange(t.arr)[src]

    def _exchange_objects(self, obj):
        self.slots[self.local.rank] = obj
        self.sync.wait(timeout=900)
        got = list(self.slots)
        self.sync.wait(timeout=900)
        return got

    def all_gather_object(self, out, obj):
        out[:] = self._exchange_objects(obj)

    def broadcast_object_list(self, box, src):
        box[:] = self._exchange_objects(list(box))[src]

    def all_reduce(self, t, op="sum"):
        got = self._exchange(t.arr)
        with np.errstate(over="ignore"):
            t.arr[...] = np.max(got, axis=0) if op == "max" else np.sum(got, axis=0, dtype=t.arr.dtype)


def test_run_ours_two_ranks_prints_the_contract_line(monkeypatch, capsys):
    """bench.py under torchrun with two ranks, played by two threads: the bitwise parity check of the slab
    decomposition (libclsph_b200.distcheck) on the job's own ranks, slab workload and capacities, resident stepping,
    stage profiling, the end-to-end loop through clsph_dist_upload / step / download, one JSON line on rank 0."""
    sys.path.insert(0, H.ROOT)
    import bench
    from libclsph_b200 import distcheck
    world = 2
    torch, cuda = fake_torch()
    fdist = _FakeDist(world)
    torch.tensor = lambda data, dtype=np.float64, device=None: _DTensor(np.array(data, dtype=dtype))
    torch.zeros = lambda n, dtype=np.float64, device=None: _DTensor(np.zeros(n, dtype=dtype))
    torch.from_numpy = lambda a: _DTensor(a)
    torch.empty = lambda n, dtype=np.uint8: _DTensor(np.empty(n, dtype=dtype))
    torch.distributed = fdist
    monkeypatch.setitem(sys.modules, "torch", torch)
    monkeypatch.setitem(sys.modules, "torch.cuda", cuda)
    monkeypatch.setitem(sys.modules, "torch.distributed", fdist)
    # the check itself at a size the emulator finishes in seconds
    real_parity = distcheck.bitwise_parity
    monkeypatch.setattr(distcheck, "bitwise_parity",
                        lambda dist, rank, world, local_rank, n_total, steps=4, options=(), verbose=False:
                        real_parity(dist, rank, world, local_rank, n_total=6000, steps=2, options=options))
    # the real script points fd 1 at stderr while native libraries are chatty; with two ranks in one process that
    # juggling would race, and the line is captured from sys.stdout here anyway
    shim = types.SimpleNamespace(**{k: getattr(bench.os, k) for k in ("path", "environ", "_exit")})
    shim.dup = lambda fd: fd
    shim.dup2 = lambda a, b: None
    monkeypatch.setattr(bench, "os", shim)

    errors = []

    def run_rank(rank):
        fdist.local.rank = rank
        args = types.SimpleNamespace(gpus=world, steps=3, warmup=3, impl="ours", config="config2_dambreak_1m", particles=12000, e2e_steps=2,
                                     no_cpu_baseline=True, cpu_sample=4096, option=[], repeats=1, no_parity=False, no_large_point=True)
        try:
            bench.run_ours(args, rank, world, 0)
        except BaseException as exc:  # noqa: BLE001
            import traceback
            errors.append((rank, traceback.format_exc()))
            fdist.sync.abort()

    threads = [threading.Thread(target=run_rank, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=1500)
    assert not errors, errors
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and d["value"] > 0 and d["config"]["particles_total"] == 24000
    par = d["multi_gpu_parity"]
    assert par["bitwise"] and par["order"] and par["ids_exact"] and par["world"] == 2, par
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
    assert d["roofline"]["stage_ms"]["exchange"] > 0
