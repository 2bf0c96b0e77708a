"""The C++ host side (sph_simulation, clsphparticles) end to end on the CPU: libclsph_host.so and the CLI
are the shipped binaries, linked against libclsph_cuda.so; here the emulator build of the same sources
(tests/emu/, same C ABI) is put in front of it with LD_PRELOAD for the duration of a subprocess. Test
infrastructure only -- see tests/emu/README.md. The same tests run against the real library on a GPU
(tests/test_host_api.py, -m gpu)."""
import os
import shutil
import subprocess
import sys

from libclsph_b200 import hostapi
from tests import helpers as H
from tests.emu import build_emu


def _env():
    return dict(os.environ, LD_PRELOAD=build_emu.build())


def test_sph_simulation_class_matches_the_oracle_under_the_emulator():
    hostapi.build()
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(H.ROOT, "tests", "test_host_api.py"), "-m", "gpu", "-q", "-x",
                        "-p", "no:cacheprovider", "-k", "simulate_matches_oracle_steps and 1-True"],
                       cwd=H.ROOT, env=_env(), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "1 passed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_cli_with_device_options_writes_frames_under_the_emulator(tmp_path):
    """clsphparticles ... --option name=value: the option pairs reach clsph_set_option (the candidate
    organisation here), frames are written by the asynchronous saver and flushed at exit; a bad option ends
    the run with the reference's print-and-exit(-1) convention."""
    hostapi.build()
    wd = str(tmp_path)
    for d in ("fluid_properties", "simulation_properties", "scenes"):
        shutil.copytree(os.path.join(H.ROOT, d), os.path.join(wd, d))
    os.makedirs(os.path.join(wd, "frames"))
    sim = open(os.path.join(wd, "simulation_properties", "default.json")).read()
    open(os.path.join(wd, "simulation_properties", "small.json"), "w").write(sim.replace('"particles_count" : 32000', '"particles_count" : 1024'))
    base = [hostapi.CLI_PATH, "water", "small", "box.obj", "", "--yes", "--frames", "2"]
    r = subprocess.run(base + ["--option", "sub_cell_order=1", "--option", "face_grid=1", "--option", "fast_pairs=1"], cwd=wd, env=_env(),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    assert sorted(os.listdir(os.path.join(wd, "frames"))) == ["frame0000001.geo", "frame0000002.geo"]
    lines = open(os.path.join(wd, "frames", "frame0000002.geo")).read().splitlines()
    assert lines[0] == "PGEOMETRY V5" and lines[1] == "NPoints 1024 NPrims 1" and lines[-1] == "endExtra"
    bad = subprocess.run(base + ["--option", "no_such_option=1"], cwd=wd, env=_env(), capture_output=True, text=True, timeout=600)
    assert bad.returncode != 0 and "unknown option" in bad.stderr


def _cli_frames(wd_root, name, extra, frames=3, particles=2048):
    wd = os.path.join(str(wd_root), name)
    os.makedirs(wd)
    for d in ("fluid_properties", "simulation_properties", "scenes"):
        shutil.copytree(os.path.join(H.ROOT, d), os.path.join(wd, d))
    os.makedirs(os.path.join(wd, "frames"))
    sim = open(os.path.join(wd, "simulation_properties", "default.json")).read()
    open(os.path.join(wd, "simulation_properties", "small.json"), "w").write(
        sim.replace('"particles_count" : 32000', '"particles_count" : %d' % particles))
    r = subprocess.run([hostapi.CLI_PATH, "water", "small", "box.obj", "", "--yes", "--frames", str(frames)] + extra, cwd=wd, env=_env(),
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-1500:]
    names = sorted(os.listdir(os.path.join(wd, "frames")))
    return {n: open(os.path.join(wd, "frames", n), "rb").read() for n in names}


def test_frames_packed_on_the_device_are_byte_identical_under_the_emulator(tmp_path):
    """--frame-export device (default with --sync full: clsph_frame_begin packs 28 bytes per particle on the GPU, the
    copy overlaps the next sub-steps, nothing else is downloaded) against --frame-export host (the reference's way:
    the pre_frame callback writes the downloaded 80-byte array): same file names, same bytes."""
    hostapi.build()
    dev = _cli_frames(tmp_path, "device", [])
    host = _cli_frames(tmp_path, "host", ["--frame-export", "host"])
    assert sorted(dev) == sorted(host) == ["frame0000001.geo", "frame0000002.geo", "frame0000003.geo"]
    for name in dev:
        assert dev[name] == host[name], name
    assert dev["frame0000001.geo"] != dev["frame0000003.geo"]  # the fluid moved


def test_cli_bgeo_frames_under_the_emulator(tmp_path):
    """clsphparticles --format bgeo (the reference's USE_PARTIO build): the device-packed and the callback path write
    the same binary files, and the points are those of the .geo run (positions to the 6 digits the text keeps)."""
    from tests.test_bgeo import read_bgeo
    hostapi.build()
    dev = _cli_frames(tmp_path, "device", ["--format", "bgeo"], frames=2)
    host = _cli_frames(tmp_path, "host", ["--format", "bgeo", "--frame-export", "host"], frames=2)
    text = _cli_frames(tmp_path, "text", [], frames=2)
    assert sorted(dev) == sorted(host) == ["frame0000001.bgeo", "frame0000002.bgeo"]
    for name in dev:
        assert dev[name] == host[name], name
    attributes, _ = read_bgeo(dev["frame0000002.bgeo"])
    lines = text["frame0000002.geo"].decode().splitlines()
    first = lines.index("mass 1 float 1") + 1
    for i in (0, 1, 1000, 2047):
        want = [float(v) for v in lines[first + i].split(" ")[:3]]
        got = attributes["position"][i]
        assert all(abs(g - w) <= 1e-5 * max(1.0, abs(w)) for g, w in zip(got, want)), (i, got, want)
    assert len(attributes["position"]) == 2048
    bad = subprocess.run([hostapi.CLI_PATH, "water", "small", "box.obj", "", "--yes", "--format", "obj"], cwd=os.path.join(str(tmp_path), "text"),
                         env=_env(), capture_output=True, text=True, timeout=600)
    assert bad.returncode != 0 and "--format expects" in bad.stderr
