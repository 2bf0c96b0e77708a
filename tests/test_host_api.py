"""The C++ drop-in host classes (libclsph_b200/host): settings, scene, frame writer on the CPU;
sph_simulation::simulate on the GPU."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from libclsph_b200 import abi, hostapi, workloads
from oracle import oracle as O
from tests import helpers as H

GOLDEN = os.path.join(H.ROOT, "tests", "golden")


@pytest.fixture(scope="module", autouse=True)
def _built():
    hostapi.build()


@pytest.mark.parametrize("fluid", ["water", "mucus"])
def test_load_settings_matches_the_reference(fluid):
    z = np.load(os.path.join(GOLDEN, "load_settings.npz"))
    p, t, vol, flags = hostapi.load_settings(os.path.join(H.ROOT, "fluid_properties", fluid + ".json"),
                                             os.path.join(H.ROOT, "simulation_properties", "default.json"))
    assert H.struct_bytes(p) == z[fluid + "_params"].tobytes()
    assert H.struct_bytes(t) == z[fluid + "_terms"].tobytes()
    assert np.float32(vol) == z[fluid + "_volume"]
    assert flags == dict(write_all_frames=False, serialize=False)


def test_load_settings_errors(tmp_path):
    sim = os.path.join(H.ROOT, "simulation_properties", "default.json")
    bad = tmp_path / "bad.json"
    bad.write_text('{"fluid_density": 1000, "dynamic_viscosity": 1, "restitution": 1.5, "k": 1,'
                   ' "surface_tension_threshold": 1, "surface_tension": 1, "particles_inside_influence_radius": 20}')
    with pytest.raises(RuntimeError, match="Restitution has an invalid value"):
        hostapi.load_settings(str(bad), sim)
    missing = tmp_path / "missing.json"
    missing.write_text('{"fluid_density": 1000};')
    with pytest.raises(RuntimeError, match="missing key"):
        hostapi.load_settings(str(missing), sim)
    with pytest.raises(RuntimeError, match="Cannot open"):
        hostapi.load_settings(str(tmp_path / "nope.json"), sim)


@pytest.mark.parametrize("scene_name", ["box.obj", "cone.obj", "cube.obj", "labyrinth.obj", "monkey.obj", "plane.obj",
                                        "river.obj", "shower.obj"])
def test_scene_load_matches_the_reference(scene_name):
    """Same three arrays as the reference's scene::load (tinyobj re-indexing included)."""
    z = np.load(os.path.join(GOLDEN, "scenes.npz"))
    normals, vertices, indices = hostapi.scene_load(scene_name, cwd=H.ROOT)
    assert np.array_equal(indices, z[scene_name + ":indices"])
    assert vertices.tobytes() == z[scene_name + ":vertices"].tobytes()
    assert normals.tobytes() == z[scene_name + ":normals"].tobytes()


def test_scene_load_failure(tmp_path):
    with pytest.raises(RuntimeError):
        hostapi.scene_load("does_not_exist.obj", cwd=H.ROOT)


def test_geo_frame_is_byte_identical_to_the_reference_writer(tmp_path):
    z = np.load(os.path.join(GOLDEN, "frame_water_n256_input.npz"))
    p = abi.SimulationParameters()
    ctypes.memmove(ctypes.addressof(p), z["params"].tobytes(), ctypes.sizeof(p))
    os.makedirs(tmp_path / "frames")
    hostapi.write_frames(str(tmp_path) + "/", z["particles"].copy(), p, frames=2)
    assert sorted(os.listdir(tmp_path / "frames")) == ["frame0000001.geo", "frame0000002.geo"]
    got = open(tmp_path / "frames" / "frame0000002.geo", "rb").read()
    want = open(os.path.join(GOLDEN, "frame_water_n256.geo"), "rb").read()
    assert got == want


# ------------------------------------------------------------------------------------------------
def _workdir(tmp_path):
    for d in ("scenes", "fluid_properties", "simulation_properties"):
        shutil.copytree(os.path.join(H.ROOT, d), tmp_path / d)
    os.makedirs(tmp_path / "frames")
    return str(tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("policy,callbacks", [(0, True), (1, True), (0, False)])
def test_simulate_matches_oracle_steps(tmp_path, policy, callbacks):
    """sph_simulation::simulate for one frame (10 sub-steps) from the lattice: same result whatever
    the host-sync policy, and within tolerance of ten oracle steps."""
    p, terms, vol, _ = workloads.make_config(fluid="water", particles_count=4096)
    normals, vertices, indices = hostapi.scene_load("box.obj", cwd=H.ROOT)
    got, p_after, calls = hostapi.simulate(p, terms, vol, normals, vertices, indices, frames=1, policy=policy,
                                           callbacks=callbacks, cwd=_workdir(tmp_path))
    assert calls == (2 * (1 + 10) if callbacks else 0)  # pre+post around the frame and around each sub-step
    scene = O.Scene(vertices, indices, normals)
    cur = O.init_particles(p, vol)
    po = p.copy()
    for _ in range(10):
        cur = O.step(cur, po, terms, scene, taps=False).particles
    assert np.array_equal(got["grid_index"], cur["grid_index"])
    H.assert_close_fields(got, cur, tol=1e-3, what="10 sub-steps")
    assert H.struct_bytes(p_after) == H.struct_bytes(po)


@pytest.mark.gpu
def test_truncated_checkpoint_is_refused_by_the_library(tmp_path):
    """A last_frame.bin shorter than particles_count records must raise (the reference's cereal loadBinary
    throws, libclsph/sph_simulation.cpp:59-67), not leave the tail particles zeroed at the origin."""
    p, terms, vol, _ = workloads.make_config(fluid="water", particles_count=4096)
    normals, vertices, indices = hostapi.scene_load("box.obj", cwd=H.ROOT)
    wd = _workdir(tmp_path)
    H.state_s0(p, vol)[:1000].tofile(os.path.join(wd, "last_frame.bin"))
    with pytest.raises(RuntimeError):
        hostapi.simulate(p, terms, vol, normals, vertices, indices, frames=1, policy=2, callbacks=False, cwd=wd)


REFERENCE_EXAMPLE = "/root/reference/example/particles.cpp"


@pytest.mark.skipif(not os.path.exists(REFERENCE_EXAMPLE), reason="the reference tree is not present on this machine")
def test_reference_example_compiles_and_links_unmodified(tmp_path):
    """Drop-in claim of INTEGRATION.md: the reference's own example/particles.cpp, unmodified, compiles against
    include/clsph (plus the reference's vendored cereal headers, which it includes itself) and links against
    libclsph_host.so + libclsph_cuda.so."""
    obj, exe = str(tmp_path / "particles.o"), str(tmp_path / "particles")
    r = subprocess.run(["g++", "-std=c++14", "-c", "-I", os.path.join(H.ROOT, "include", "clsph"), "-I", "/root/reference",
                        REFERENCE_EXAMPLE, "-o", obj], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    libdir = os.path.join(H.ROOT, "libclsph_b200")
    r = subprocess.run(["g++", "-o", exe, obj, "-L" + libdir, "-lclsph_host", "-lclsph_cuda", "-Wl,-rpath," + libdir],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert os.path.getsize(exe) > 0


@pytest.mark.gpu
def test_clsphparticles_cli_writes_frames_and_checkpoint(tmp_path):
    """The command-line driver end to end: JSON in, .geo frames and last_frame.bin out, resume."""
    hostapi.build()  # makes libclsph_host.so and the clsphparticles binary if they are missing or stale
    wd = _workdir(tmp_path)
    sim = open(os.path.join(wd, "simulation_properties", "default.json")).read()
    sim = sim.replace('"particles_count" : 32000', '"particles_count" : 2048').replace('"serialize" : false', '"serialize" : true')
    open(os.path.join(wd, "simulation_properties", "small.json"), "w").write(sim)
    run = lambda: subprocess.run([hostapi.CLI_PATH, "water", "small", "box.obj", "", "--yes", "--frames", "2"], cwd=wd,
                                 capture_output=True, text=True, timeout=300)
    r = run()
    assert r.returncode == 0, r.stderr
    assert "Kernel support radius (h):" in r.stdout
    frames = sorted(os.listdir(os.path.join(wd, "frames")))
    assert frames == ["frame0000001.geo", "frame0000002.geo"]
    head = open(os.path.join(wd, "frames", frames[0])).read().splitlines()
    assert head[0] == "PGEOMETRY V5" and head[1] == "NPoints 2048 NPrims 1"
    ckpt = os.path.join(wd, "last_frame.bin")
    assert os.path.getsize(ckpt) == 2048 * 80
    first = np.fromfile(ckpt, dtype=abi.PARTICLE)
    r = run()  # resumes from the checkpoint
    assert "Serialized frame found" in r.stdout
    second = np.fromfile(ckpt, dtype=abi.PARTICLE)
    assert first["position"][:, 1].mean() > second["position"][:, 1].mean()  # kept falling
    open(ckpt, "ab").write(b"x")  # wrong size -> refuses to run
    r = run()
    assert "incorrect size" in r.stdout


@pytest.mark.gpu
def test_frames_packed_on_the_device_are_byte_identical(tmp_path):
    """clsphparticles --frame-export device (default: 28 bytes per particle packed on the GPU, copied while the next
    sub-steps run) against --frame-export host (the downloaded 80-byte array written by the callback): same bytes."""
    hostapi.build()
    out = {}
    for mode, extra in (("device", []), ("host", ["--frame-export", "host"])):
        wd = _workdir(tmp_path / mode)
        sim = open(os.path.join(wd, "simulation_properties", "default.json")).read()
        open(os.path.join(wd, "simulation_properties", "small.json"), "w").write(sim.replace('"particles_count" : 32000', '"particles_count" : 20000'))
        r = subprocess.run([hostapi.CLI_PATH, "water", "small", "box.obj", "", "--yes", "--frames", "3"] + extra, cwd=wd,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        names = sorted(os.listdir(os.path.join(wd, "frames")))
        out[mode] = {n: open(os.path.join(wd, "frames", n), "rb").read() for n in names}
    assert sorted(out["device"]) == sorted(out["host"]) == ["frame0000001.geo", "frame0000002.geo", "frame0000003.geo"]
    for name in out["device"]:
        assert out["device"][name] == out["host"][name], name


@pytest.mark.gpu
def test_bgeo_frames_from_the_device_hold_the_simulated_state(tmp_path):
    """clsphparticles --format bgeo on the GPU (100 000 particles, so 32-bit vertex numbers): device-packed and callback
    paths write the same bytes; the file parses; densities in the colour ramp's range."""
    from tests.test_bgeo import read_bgeo
    hostapi.build()
    out = {}
    for mode, extra in (("device", []), ("host", ["--frame-export", "host"])):
        wd = _workdir(tmp_path / mode)
        sim = open(os.path.join(wd, "simulation_properties", "default.json")).read()
        open(os.path.join(wd, "simulation_properties", "small.json"), "w").write(sim.replace('"particles_count" : 32000', '"particles_count" : 100000'))
        r = subprocess.run([hostapi.CLI_PATH, "water", "small", "box.obj", "", "--yes", "--frames", "2", "--format", "bgeo"] + extra, cwd=wd,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        names = sorted(os.listdir(os.path.join(wd, "frames")))
        out[mode] = {n: open(os.path.join(wd, "frames", n), "rb").read() for n in names}
    assert sorted(out["device"]) == sorted(out["host"]) == ["frame0000001.bgeo", "frame0000002.bgeo"]
    for name in out["device"]:
        assert out["device"][name] == out["host"][name], name
    attributes, _ = read_bgeo(out["device"]["frame0000002.bgeo"])
    assert len(attributes["position"]) == 100000 and np.all(np.isfinite(attributes["position"]))
    assert np.all(attributes["color"] >= 0) and np.all(attributes["color"] <= 1) and attributes["color"].max() > 0
    assert np.any(attributes["velocity"][:, 1] != 0)  # all three velocity components are written


def test_geo_numbers_match_printf_g_on_adversarial_values(tmp_path):
    """The frame writer formats with std::to_chars on worker threads; every number must still read
    exactly like the reference's iostream output (printf %g): checked against Python's %g over
    magnitudes from denormal to huge, signed zeros and non-finite values, on enough particles to
    use several formatting threads."""
    rng = np.random.default_rng(5)
    n = 40000
    s = np.zeros(n, dtype=abi.PARTICLE)
    mag = (10.0 ** rng.uniform(-44, 38, size=(n, 6))) * rng.choice([-1.0, 1.0], size=(n, 6))
    with np.errstate(over="ignore"):
        vals = mag.astype(np.float32)
    vals[0] = [0.0, -0.0, np.inf, -np.inf, 1e-45, 3.4028235e38]
    vals[1] = [100000.0, 1000000.0, 999999.5, 0.0001, 0.00001, 123456.5]
    vals[2, 0] = np.nan
    s["position"][:, :3] = vals[:, :3]
    s["velocity"][:, :3] = vals[:, 3:]
    s["density"] = rng.uniform(-100, 2500, n).astype(np.float32)
    p, terms, vol = H.config("water", n)
    os.makedirs(tmp_path / "frames")
    hostapi.write_frames(str(tmp_path) + "/", s, p, frames=1)
    lines = open(tmp_path / "frames" / "frame0000001.geo").read().split("\n")
    first = lines.index("mass 1 float 1") + 1
    g = lambda x: "%g" % float(x)
    for i in list(range(64)) + list(rng.integers(0, n, 3000)):
        rho = np.float32(s["density"][i])
        one, k, half = np.float32(1), np.float32(1000), np.float32(500)
        r = (rho - k) / k if k < rho <= 2 * k else np.float32(0)
        gr = one - rho / k if 0 <= rho < k else np.float32(0)
        b = (rho - half) / half if half <= rho <= k else (one - (rho - k) / half if k <= rho <= k + half else np.float32(0))
        want = "%s %s %s 0 (%s %s %s\t%s %s %s\t%s)" % (g(vals[i, 0]), g(vals[i, 1]), g(vals[i, 2]), g(vals[i, 3]), g(vals[i, 4]),
                                                      g(vals[i, 5]), g(r), g(gr), g(b), g(np.float32(p.particle_mass)))
        assert lines[first + i] == want, (i, lines[first + i], want)
    assert lines[first + n + 3].startswith("Part %d 0 1 2 3" % n) and lines[first + n + 3].endswith(" %d [0\t0]" % (n - 1))


def test_number_formatting_equals_printf_g_over_a_strided_sweep_of_all_floats():
    """The writer formats with integer arithmetic (exact digits of m * 2^q); an exhaustive comparison with
    snprintf("%g") over all 2^32 bit patterns was run when it was written (0 mismatches); this is a strided
    sample of the same sweep plus the boundary cases, against Python's %g (correctly rounded, like glibc's)."""
    lib = hostapi.lib()
    lib.clsph_host_format_g.argtypes = [ctypes.c_float, ctypes.c_char_p]
    lib.clsph_host_format_g.restype = ctypes.c_int
    a = ctypes.create_string_buffer(64)
    bits = np.concatenate([np.arange(0, 1 << 32, 40009, dtype=np.uint64).astype(np.uint32),
                           np.array([0, 1, 0x80000000, 0x00800000, 0x007fffff, 0x7f7fffff, 0x3f800000, 0x7f800000, 0xff800000],
                                    dtype=np.uint32),
                           np.array([999999.5, 999999.44, 99999.95, 0.0001, 0.000099999994, 1e-24, 9.9999994e-25, 1e6, 123456.5, 0.5,
                                     1e-5, 100000.0], dtype=np.float32).view(np.uint32)])
    vals = bits.view(np.float32)
    vals = vals[~np.isnan(vals)]
    for v in vals.tolist():
        lib.clsph_host_format_g(v, a)
        want = "%g" % v   # Python formats doubles with correctly rounded digits, like glibc
        assert a.value.decode() == want, (v, a.value, want)
