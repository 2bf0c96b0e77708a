"""Binary records of the libclsph API as numpy / ctypes types.

Mirrors include/clsph/clsph_types.h (which follows the reference's
libclsph/common/structures.h:16-54): `particle` is 80 bytes, `simulation_parameters` 128,
`precomputed_kernel_values` 20. Used by the ctypes binding of the CUDA library and by the
test/bench harness; there is no arithmetic in this module.
"""
import ctypes

import numpy as np

PARTICLE = np.dtype(
    [
        ("position", "<f4", (4,)),
        ("velocity", "<f4", (4,)),
        ("intermediate_velocity", "<f4", (4,)),
        ("acceleration", "<f4", (4,)),
        ("density", "<f4"),
        ("pressure", "<f4"),
        ("grid_index", "<u4"),
        ("_pad", "<u4"),
    ]
)
assert PARTICLE.itemsize == 80


class Float3(ctypes.Structure):
    _fields_ = [("s", ctypes.c_float * 4)]

    def xyz(self):
        return (self.s[0], self.s[1], self.s[2])


class SimulationParameters(ctypes.Structure):
    _fields_ = [
        ("particles_count", ctypes.c_uint32),
        ("max_velocity", ctypes.c_float),
        ("fluid_density", ctypes.c_float),
        ("total_mass", ctypes.c_float),
        ("particle_mass", ctypes.c_float),
        ("dynamic_viscosity", ctypes.c_float),
        ("simulation_time", ctypes.c_float),
        ("target_fps", ctypes.c_float),
        ("h", ctypes.c_float),
        ("simulation_scale", ctypes.c_float),
        ("time_delta", ctypes.c_float),
        ("surface_tension_threshold", ctypes.c_float),
        ("surface_tension", ctypes.c_float),
        ("restitution", ctypes.c_float),
        ("K", ctypes.c_float),
        ("_pad0", ctypes.c_uint32),
        ("constant_acceleration", Float3),
        ("grid_size_x", ctypes.c_int32),
        ("grid_size_y", ctypes.c_int32),
        ("grid_size_z", ctypes.c_int32),
        ("grid_cell_count", ctypes.c_uint32),
        ("min_point", Float3),
        ("max_point", Float3),
    ]

    def copy(self):
        other = SimulationParameters()
        ctypes.memmove(ctypes.byref(other), ctypes.byref(self), ctypes.sizeof(self))
        return other


class PrecomputedKernelValues(ctypes.Structure):
    _fields_ = [
        ("poly_6", ctypes.c_float),
        ("poly_6_gradient", ctypes.c_float),
        ("poly_6_laplacian", ctypes.c_float),
        ("spiky", ctypes.c_float),
        ("viscosity", ctypes.c_float),
    ]


assert ctypes.sizeof(SimulationParameters) == 128
assert SimulationParameters.constant_acceleration.offset == 64
assert SimulationParameters.grid_size_x.offset == 80
assert SimulationParameters.grid_cell_count.offset == 92
assert SimulationParameters.min_point.offset == 96
assert SimulationParameters.max_point.offset == 112
assert ctypes.sizeof(PrecomputedKernelValues) == 20


def particle_ptr(arr):
    """ctypes pointer to a C-contiguous PARTICLE array (no copy)."""
    assert arr.dtype == PARTICLE and arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data_as(ctypes.c_void_p)


# Fluid / simulation settings as shipped in fluid_properties/*.json and
# simulation_properties/default.json (reference files of the same names).
FLUIDS = {
    "water": dict(fluid_density=998.29, dynamic_viscosity=3.5, restitution=0.0, k=100.0,
                  surface_tension_threshold=7.065, surface_tension=0.0728,
                  particles_inside_influence_radius=20),
    "mucus": dict(fluid_density=1000.0, dynamic_viscosity=36.0, restitution=0.5, k=5.0,
                  surface_tension_threshold=5.0, surface_tension=6.0,
                  particles_inside_influence_radius=40),
}
DEFAULT_SIM = dict(particles_count=32000, particle_mass=0.05, simulation_time=10.0,
                   target_fps=60.0, simulation_scale=0.1, constant_acceleration=(0.0, -9.8, 0.0))


def raw_parameters(fluid="water", particles_count=None, particle_mass=None, **overrides):
    """SimulationParameters holding only the JSON-level inputs (no derived constants).

    Returns (params, particles_inside_influence_radius). Derived fields (h, max_velocity,
    smoothing constants) come from the host library or, in tests, from the oracle.
    """
    f = dict(FLUIDS[fluid])
    s = dict(DEFAULT_SIM)
    if particles_count is not None:
        s["particles_count"] = particles_count
    if particle_mass is not None:
        s["particle_mass"] = particle_mass
    s.update({k: v for k, v in overrides.items() if k in s})
    f.update({k: v for k, v in overrides.items() if k in f})
    p = SimulationParameters()
    p.particles_count = int(s["particles_count"])
    p.particle_mass = s["particle_mass"]
    p.simulation_time = s["simulation_time"]
    p.target_fps = s["target_fps"]
    p.simulation_scale = s["simulation_scale"]
    for k in range(3):
        p.constant_acceleration.s[k] = s["constant_acceleration"][k]
    p.fluid_density = f["fluid_density"]
    p.dynamic_viscosity = f["dynamic_viscosity"]
    p.restitution = f["restitution"]
    p.K = f["k"]
    p.surface_tension_threshold = f["surface_tension_threshold"]
    p.surface_tension = f["surface_tension"]
    return p, int(f["particles_inside_influence_radius"])
