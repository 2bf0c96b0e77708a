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


def test_organisation_choice_adopts_the_candidate_only_after_a_clean_selfcheck(monkeypatch):
    sys.path.insert(0, H.ROOT)
    import subprocess as sp
    import types
    import bench

    def fake(stdout, returncode=0, raises=None):
        def run_(cmd, **kw):
            assert "libclsph_b200.selfcheck" in cmd and "--set" in cmd
            if raises:
                raise raises
            return types.SimpleNamespace(stdout=stdout, stderr="boom", returncode=returncode)
        return run_

    args = types.SimpleNamespace(option=[], organisation="auto", config="config2_dambreak_1m")
    a = dict(kv.split("=") for kv in bench.CANDIDATE_SETS[0])
    a = {k: int(v) for k, v in a.items()}
    b = dict(a, deferred_lists=1)
    ok = json.dumps({"agree": True, "ms_per_step_default": 0.9, "sets": [
        {"options": a, "agree": True, "max_rel_diff": 1e-7, "ms_per_step": 0.5},
        {"options": b, "agree": True, "max_rel_diff": 1e-7, "ms_per_step": 0.45}]})
    monkeypatch.setattr(sp, "run", fake("NCCL noise\n" + ok + "\n"))
    opts, rep = bench.choose_organisation(args, 0, 1 << 20)
    assert sorted(opts) == sorted("%s=%d" % kv for kv in b.items()) and rep["adopted"] and rep["agree"]   # the faster of the two
    one_bad = json.dumps({"agree": True, "ms_per_step_default": 0.9, "sets": [
        {"options": a, "agree": True, "max_rel_diff": 1e-7, "ms_per_step": 0.5},
        {"options": b, "agree": False, "error": "AssertionError: support_count differs"}]})
    monkeypatch.setattr(sp, "run", fake(one_bad))
    assert sorted(bench.choose_organisation(args, 0, 1 << 20)[0]) == sorted(bench.CANDIDATE_OPTIONS)
    slower = json.dumps({"agree": True, "ms_per_step_default": 0.5, "sets": [
        {"options": a, "agree": True, "max_rel_diff": 1e-7, "ms_per_step": 0.9}]})
    monkeypatch.setattr(sp, "run", fake(slower))
    assert bench.choose_organisation(args, 0, 1 << 20)[0] == []
    bad = json.dumps({"agree": False, "ms_per_step_default": 0.5, "sets": [{"options": a, "agree": False, "error": "AssertionError: permutation differs"}]})
    monkeypatch.setattr(sp, "run", fake(bad, returncode=1))
    opts, rep = bench.choose_organisation(args, 0, 1 << 20)
    assert opts == [] and not rep["adopted"] and "permutation" in rep["sets"][0]["error"]
    monkeypatch.setattr(sp, "run", fake("", returncode=-11))   # the subprocess crashed
    opts, rep = bench.choose_organisation(args, 0, 1 << 20)
    assert opts == [] and not rep["adopted"] and "exit code -11" in rep["error"]
    monkeypatch.setattr(sp, "run", fake("", raises=sp.TimeoutExpired("selfcheck", 600)))   # ... or hung
    assert bench.choose_organisation(args, 0, 1 << 20)[0] == []
    # explicit choices bypass the check
    args.organisation = "default"
    assert bench.choose_organisation(args, 0, 1 << 20)[0] == []
    args.organisation, args.option = "auto", ["list_rows=48"]
    assert bench.choose_organisation(args, 0, 1 << 20)[0] == ["list_rows=48"]
