"""bench.py contract pieces that run without a GPU: the reference arm's JSON line, the helper
tables, and that the GPU arm fails fast (non-zero, no hang, no stdout noise) when there is no GPU."""
import json
import os
import subprocess
import sys

from tests import helpers as H

BENCH = os.path.join(H.ROOT, "bench.py")


def run(args, env=None, timeout=600):
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout, cwd=H.ROOT,
                          env=dict(os.environ, **(env or {})))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--cpu-sample", "4096"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle_steps_per_sec" and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "config2_dambreak_1m"


def test_reference_arm_other_ranks_exit_silently():
    r = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_fails_fast_without_a_gpu():
    from libclsph_b200 import capi
    if capi.load_library().clsph_device_count() > 0:
        import pytest
        pytest.skip("a GPU is present")
    r = run(["--steps", "1", "--warmup", "0", "--no-cpu-baseline", "--e2e-steps", "0", "--particles", "4096"], timeout=300)
    assert r.returncode != 0
    assert r.stdout.strip() == ""          # nothing that could be mistaken for a result
    assert "NVIDIA" in r.stderr or "CUDA" in r.stderr


def test_kernel_byte_table_matches_the_survey_total():
    sys.path.insert(0, H.ROOT)
    import bench
    for passes in (2, 3, 4):
        kb = dict(bench.KERNEL_BYTES, sort=4.0 + 16.0 * passes)
        # keys 20 + sort (4 + 16 P) + reorder 104 (100 + the sorted key) + density 24 + forces 56 + integrate 96
        # = 304 + 16 P; SURVEY 8(d) fuses force+integrate (104) and counts the cell table read (4): 256 + 16 P
        assert sum(kb.values()) == 304 + 16 * passes
        assert bench.step_bytes(passes) == 256 + 16 * passes
