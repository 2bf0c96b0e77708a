"""Writes tests/golden/frame_water_n256.bgeo from tests/golden/frame_water_n256_input.npz with this repository's own
houdini_file_saver (format = bgeo). The fixture pins the format against accidental change; it is NOT a libpartio output
(libpartio is not part of the reference tree), see tests/test_bgeo.py."""
import ctypes
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libclsph_b200 import abi, hostapi  # noqa: E402

hostapi.build()
golden = os.path.join(ROOT, "tests", "golden")
z = np.load(os.path.join(golden, "frame_water_n256_input.npz"))
p = abi.SimulationParameters()
ctypes.memmove(ctypes.addressof(p), z["params"].tobytes(), ctypes.sizeof(p))
with tempfile.TemporaryDirectory() as tmp:
    os.makedirs(os.path.join(tmp, "frames"))
    hostapi.write_frames(tmp + "/", z["particles"].copy(), p, frames=1, fmt="bgeo")
    shutil.copy(os.path.join(tmp, "frames", "frame0000001.bgeo"), os.path.join(golden, "frame_water_n256.bgeo"))
print("wrote", os.path.join(golden, "frame_water_n256.bgeo"))
