"""The binary frame format (the Partio half): houdini_file_saver with format = bgeo writes the
"Bgeo V5" file the reference produces through libpartio when compiled with USE_PARTIO
(libclsph/file_save_delegates/houdini_file_saver.cpp:78-88, util/partio/PartioFunctions.h:5-65).

Parity is UNPINNED for this format: libpartio is not part of the reference tree (headers only), so there is no
reference-written .bgeo to compare bytes with. What is checked: the file parses with the independent reader below,
written from the layout of Partio's BGEO reader (magic, 'V', 5, nine counts, attribute definitions, points of
x y z w + attribute words, primitive attribute table, one 0x8000 particle primitive, 0x00 0xff), and every value is
the one PartioFunctions.h would store (position, velocity -- all three components, see SURVEY E11 --, the density
colour ramp, id = index, mass, pscale = h)."""
import ctypes
import os
import struct
import subprocess

import numpy as np
import pytest

from libclsph_b200 import abi, hostapi
from tests import helpers as H

GOLDEN = os.path.join(H.ROOT, "tests", "golden")
TYPE_NAMES = {0: "float", 1: "int", 5: "vector"}


def read_bgeo(data):
    """Independent parser of Houdini's binary "Bgeo V5" particle files as Partio writes them. Returns
    (attributes {name: (n, size) array}, definitions [(name, size, type)]). Raises on anything unexpected,
    trailing bytes included."""
    pos = 0

    def take(fmt):
        nonlocal pos
        vals = struct.unpack_from(">" + fmt, data, pos)
        pos += struct.calcsize(">" + fmt)
        return vals if len(vals) > 1 else vals[0]

    def take_str():
        nonlocal pos
        n = take("H")
        s = data[pos:pos + n].decode("ascii")
        pos += n
        return s

    assert data[:4] == b"Bgeo"
    pos = 4
    assert take("c") == b"V"
    version, n_points, n_prims, n_point_groups = take("iiii")
    n_prim_groups, n_point_attrib, n_vertex_attrib, n_prim_attrib, n_attrib = take("iiiii")
    assert (version, n_prims, n_point_groups, n_prim_groups, n_vertex_attrib, n_prim_attrib, n_attrib) == (5, 1, 0, 0, 0, 1, 0)
    definitions = []
    words = 4
    for _ in range(n_point_attrib):
        name = take_str()
        size = take("H")
        kind = take("i")
        assert kind in TYPE_NAMES
        defaults = [take("i") for _ in range(size)]
        assert defaults == [0] * size
        definitions.append((name, size, kind))
        words += size
    raw = np.frombuffer(data, dtype=">u4", count=n_points * words, offset=pos).reshape(n_points, words)
    pos += n_points * words * 4
    attributes = {"position": raw[:, :3].astype("<u4").view("<f4"), "w": raw[:, 3].astype("<u4").view("<f4")}
    col = 4
    for name, size, kind in definitions:
        block = raw[:, col:col + size].astype("<u4")
        attributes[name] = block.view("<i4") if kind == 1 else block.view("<f4")
        col += size
    # primitive attribute table: "generator", one index value, a table of one string
    assert take_str() == "generator"
    assert take("H") == 1 and take("i") == 4 and take("i") == 1
    assert take_str() == "papi"
    assert take("I") == 0x8000            # particle system
    assert take("i") == n_points
    wide = n_points > (1 << 16)
    idx = np.frombuffer(data, dtype=">u4" if wide else ">u2", count=n_points, offset=pos)
    pos += n_points * (4 if wide else 2)
    assert np.array_equal(idx.astype(np.int64), np.arange(n_points))
    assert take("i") == 0                 # the primitive's generator: string 0
    assert data[pos:] == b"\x00\xff"
    return attributes, definitions


def colour_ramp(rho):
    """houdini_file_saver.cpp:47-60 / PartioFunctions.h:41-56 in fp32."""
    rho = rho.astype(np.float32)
    f = np.float32
    r = np.where((rho > 1000) & (rho <= 2000), (rho - f(1000)) / f(1000), f(0))
    g = np.where((rho >= 0) & (rho < 1000), f(1) - rho / f(1000), f(0))
    b = np.where((rho >= 500) & (rho <= 1000), (rho - f(500)) / f(500),
                 np.where((rho >= 1000) & (rho <= 1500), f(1) - (rho - f(1000)) / f(500), f(0)))
    return np.stack([r, g, b], axis=1).astype(np.float32)


@pytest.fixture(scope="module", autouse=True)
def _built():
    hostapi.build()


def _golden_state():
    z = np.load(os.path.join(GOLDEN, "frame_water_n256_input.npz"))
    p = abi.SimulationParameters()
    ctypes.memmove(ctypes.addressof(p), z["params"].tobytes(), ctypes.sizeof(p))
    return z["particles"].copy(), p


def _check(attributes, definitions, particles, p):
    assert [(n, s, TYPE_NAMES[k]) for n, s, k in definitions] == [
        ("velocity", 3, "vector"), ("color", 3, "vector"), ("id", 1, "int"), ("mass", 1, "float"), ("pscale", 1, "float")]
    n = len(particles)
    assert np.array_equal(attributes["position"].view(np.uint32), np.ascontiguousarray(particles["position"][:, :3]).view(np.uint32))
    assert np.all(attributes["w"] == 1.0)
    assert np.array_equal(attributes["velocity"].view(np.uint32), np.ascontiguousarray(particles["velocity"][:, :3]).view(np.uint32))
    assert np.array_equal(attributes["color"].view(np.uint32), colour_ramp(particles["density"]).view(np.uint32))
    assert np.array_equal(attributes["id"][:, 0], np.arange(n))
    assert np.all(attributes["mass"] == np.float32(p.particle_mass))
    assert np.all(attributes["pscale"] == np.float32(p.h))


@pytest.mark.parametrize("packed", [False, True])
def test_bgeo_frame_parses_and_holds_the_reference_attributes(tmp_path, packed):
    particles, p = _golden_state()
    os.makedirs(tmp_path / "frames")
    hostapi.write_frames(str(tmp_path) + "/", particles, p, frames=2, fmt="bgeo", packed=packed)
    assert sorted(os.listdir(tmp_path / "frames")) == ["frame0000001.bgeo", "frame0000002.bgeo"]
    data = open(tmp_path / "frames" / "frame0000002.bgeo", "rb").read()
    attributes, definitions = read_bgeo(data)
    _check(attributes, definitions, particles, p)
    # size: header 41, definitions (2+len+2+4+4*size each), 13 words per point, 35 of primitive table + head, indices, 6
    names = sum(2 + len(n) + 2 + 4 + 4 * s for n, s, _ in definitions)
    assert len(data) == 41 + names + 52 * len(particles) + 35 + 2 * len(particles) + 6


def test_bgeo_matches_the_committed_fixture(tmp_path):
    """tests/golden/frame_water_n256.bgeo was written by this writer (tools/make_bgeo_fixture.py) and inspected by
    hand against the layout; it pins the format against accidental change, not against libpartio."""
    particles, p = _golden_state()
    os.makedirs(tmp_path / "frames")
    hostapi.write_frames(str(tmp_path) + "/", particles, p, frames=1, fmt="bgeo")
    got = open(tmp_path / "frames" / "frame0000001.bgeo", "rb").read()
    assert got == open(os.path.join(GOLDEN, "frame_water_n256.bgeo"), "rb").read()


def test_bgeo_wide_vertex_numbers_above_65536_points(tmp_path):
    particles, p = _golden_state()
    n = (1 << 16) + 37
    big = np.resize(particles, n).copy()
    big["position"][:, 0] = np.arange(n, dtype=np.float32)
    big["density"] = np.linspace(-10.0, 2500.0, n).astype(np.float32)   # the whole colour ramp, both ends included
    p.particles_count = n
    os.makedirs(tmp_path / "frames")
    hostapi.write_frames(str(tmp_path) + "/", big, p, frames=1, fmt="bgeo", packed=True)
    attributes, definitions = read_bgeo(open(tmp_path / "frames" / "frame0000001.bgeo", "rb").read())
    _check(attributes, definitions, big, p)
    # exactly 65536 points still use 16-bit vertex numbers (Partio: nPoints > 1 << 16)
    p.particles_count = 1 << 16
    hostapi.write_frames(str(tmp_path) + "/", big[:1 << 16].copy(), p, frames=1, fmt="bgeo")
    data = open(tmp_path / "frames" / "frame0000001.bgeo", "rb").read()
    attributes, definitions = read_bgeo(data)
    _check(attributes, definitions, big[:1 << 16], p)


def test_geo_format_is_unchanged_by_the_switch(tmp_path):
    particles, p = _golden_state()
    os.makedirs(tmp_path / "frames")
    hostapi.write_frames(str(tmp_path) + "/", particles, p, frames=2, fmt="geo", packed=True)
    got = open(tmp_path / "frames" / "frame0000002.geo", "rb").read()
    assert got == open(os.path.join(GOLDEN, "frame_water_n256.geo"), "rb").read()
