"""ctypes binding of the C ABI in include/clsph_cuda.h (libclsph_b200/libclsph_cuda.so).

This is the only way Python reaches the CUDA kernels; there is no other compute path. If the
shared library is missing or no CUDA device is usable, construction fails loudly.
"""
import ctypes
import os

import numpy as np

from . import build as _build
from .abi import PARTICLE, PrecomputedKernelValues, SimulationParameters, particle_ptr

# every symbol include/clsph_cuda.h declares
SYMBOLS = [
    "clsph_device_count", "clsph_create", "clsph_destroy", "clsph_last_error", "clsph_set_scene",
    "clsph_set_parameters", "clsph_set_option", "clsph_upload_particles", "clsph_step", "clsph_synchronize",
    "clsph_get_parameters", "clsph_download_particles", "clsph_simulate_single_frame", "clsph_set_debug",
    "clsph_debug_fetch", "clsph_kernel_advection_collision", "clsph_comm_unique_id", "clsph_dist_init",
    "clsph_dist_upload", "clsph_dist_download", "clsph_dist_transport", "clsph_frame_begin", "clsph_frame_end", "clsph_host_alloc", "clsph_host_free", "clsph_profile_enable", "clsph_profile_read", "clsph_particle_count", "clsph_sort_passes", "clsph_stream",
]

TAP_SORTED_KEYS, TAP_PERMUTATION, TAP_CELL_TABLE, TAP_KEYS_INPUT, TAP_CANDIDATE_COUNT = 0, 1, 2, 3, 4
TAP_SUPPORT_COUNT, TAP_DENSITY, TAP_PRESSURE, TAP_ACCELERATION, TAP_COLLISION_ITERS = 5, 6, 7, 8, 9

E_OK, E_INVAL, E_CUDA, E_GRID, E_STATE, E_COMM, E_NOMEM = range(7)


class StageTimes(ctypes.Structure):
    _fields_ = [(k, ctypes.c_double) for k in ("ms_bounds_grid", "ms_keys", "ms_sort", "ms_reorder", "ms_density",
                                               "ms_forces", "ms_integrate", "ms_exchange")] + [
        ("substeps", ctypes.c_uint64), ("kernel_launches", ctypes.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class ClsphError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("clsph error %d: %s" % (code, message))
        self.code = code


_lib = None


def load_library(path=None):
    """dlopen libclsph_cuda.so (building it first if the sources are newer) and type its entry points."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or _build.LIB_PATH
    if not os.path.exists(path):
        _build.build()
    L = ctypes.CDLL(path)
    vp, u32, sz = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_size_t
    L.clsph_device_count.restype = ctypes.c_int
    L.clsph_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int, u32, u32]
    L.clsph_destroy.argtypes = [vp]
    L.clsph_destroy.restype = None
    L.clsph_last_error.argtypes = [vp]
    L.clsph_last_error.restype = ctypes.c_char_p
    L.clsph_set_scene.argtypes = [vp, vp, vp, sz, vp, u32]
    L.clsph_set_parameters.argtypes = [vp, vp, vp]
    L.clsph_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_longlong]
    L.clsph_upload_particles.argtypes = [vp, vp, u32]
    L.clsph_step.argtypes = [vp, u32]
    L.clsph_synchronize.argtypes = [vp]
    L.clsph_get_parameters.argtypes = [vp, vp]
    L.clsph_download_particles.argtypes = [vp, vp]
    L.clsph_simulate_single_frame.argtypes = [vp, vp, vp, vp, vp]
    L.clsph_kernel_advection_collision.argtypes = [vp, vp, vp, u32]
    L.clsph_comm_unique_id.argtypes = [vp, sz]
    L.clsph_dist_init.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_float, ctypes.c_float, u32, u32]
    L.clsph_dist_upload.argtypes = [vp, vp, vp, u32]
    L.clsph_dist_download.argtypes = [vp, vp, vp, u32, ctypes.POINTER(u32)]
    L.clsph_set_debug.argtypes = [vp, ctypes.c_int]
    L.clsph_debug_fetch.argtypes = [vp, ctypes.c_int, vp, sz]
    L.clsph_profile_enable.argtypes = [vp, ctypes.c_int]
    L.clsph_profile_read.argtypes = [vp, vp]
    L.clsph_particle_count.argtypes = [vp, ctypes.POINTER(u32)]
    L.clsph_sort_passes.argtypes = [vp, ctypes.POINTER(u32)]
    L.clsph_stream.argtypes = [vp]
    L.clsph_stream.restype = vp
    L.clsph_frame_begin.argtypes = [vp, vp, u32]
    L.clsph_frame_end.argtypes = [vp]
    L.clsph_host_alloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), sz]
    L.clsph_host_free.argtypes = [vp]
    L.clsph_host_free.restype = None
    L.clsph_dist_transport.argtypes = [vp]
    L.clsph_dist_transport.restype = ctypes.c_char_p
    for name in SYMBOLS:
        if name not in ("clsph_destroy", "clsph_last_error", "clsph_stream", "clsph_dist_transport", "clsph_host_free"):
            getattr(L, name).restype = ctypes.c_int
    if path == _build.LIB_PATH:
        _lib = L
    return L


def _vp(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def comm_unique_id():
    """128-byte NCCL unique id (bytes) for clsph_dist_init; create on one rank, broadcast to all."""
    lib = load_library()
    buf = ctypes.create_string_buffer(128)
    rc = lib.clsph_comm_unique_id(buf, 128)
    if rc:
        raise ClsphError(rc, (lib.clsph_last_error(None) or b"").decode())
    return buf.raw


class Context:
    """Thin object wrapper over a clsph_context* (one GPU, one caller thread)."""

    def __init__(self, max_particles, device=0, cell_table_capacity=0):
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        rc = self._lib.clsph_create(ctypes.byref(self._h), device, max_particles, cell_table_capacity)
        if rc:
            raise ClsphError(rc, (self._lib.clsph_last_error(None) or b"").decode())

    def _check(self, rc):
        if rc:
            raise ClsphError(rc, (self._lib.clsph_last_error(self._h) or b"").decode())

    def close(self):
        if self._h:
            self._lib.clsph_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_scene(self, face_normals, vertices, indices):
        face_normals = np.ascontiguousarray(face_normals, dtype=np.float32).reshape(-1)
        vertices = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1)
        indices = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        self._check(self._lib.clsph_set_scene(self._h, _vp(face_normals), _vp(vertices), vertices.size, _vp(indices),
                                              indices.size // 3))

    def set_parameters(self, params, terms):
        self._check(self._lib.clsph_set_parameters(self._h, ctypes.byref(params), ctypes.byref(terms)))

    def set_option(self, name, value):
        self._check(self._lib.clsph_set_option(self._h, name.encode(), int(value)))

    def upload(self, particles):
        self._check(self._lib.clsph_upload_particles(self._h, particle_ptr(particles), particles.size))

    def upload_ptr(self, host_ptr, n):
        """Upload from a raw host pointer (e.g. a pinned torch tensor's data_ptr())."""
        self._check(self._lib.clsph_upload_particles(self._h, ctypes.c_void_p(host_ptr), n))

    def step(self, n_substeps=1):
        self._check(self._lib.clsph_step(self._h, n_substeps))

    def synchronize(self):
        self._check(self._lib.clsph_synchronize(self._h))

    def parameters(self):
        p = SimulationParameters()
        self._check(self._lib.clsph_get_parameters(self._h, ctypes.byref(p)))
        return p

    def particle_count(self):
        n = ctypes.c_uint32()
        self._check(self._lib.clsph_particle_count(self._h, ctypes.byref(n)))
        return n.value

    def sort_passes(self):
        """Radix passes of the last sub-step; 0 = counting sort on the dense sub-cell table."""
        n = ctypes.c_uint32()
        self._check(self._lib.clsph_sort_passes(self._h, ctypes.byref(n)))
        return n.value

    def download(self, out=None):
        if out is None:
            out = np.empty(self.particle_count(), dtype=PARTICLE)
        self._check(self._lib.clsph_download_particles(self._h, particle_ptr(out)))
        return out

    def download_ptr(self, host_ptr):
        self._check(self._lib.clsph_download_particles(self._h, ctypes.c_void_p(host_ptr)))

    def simulate_single_frame(self, particles_in, params, terms=None, out=None):
        """Host in, host out: the shape of the reference's simulate_single_frame(in, out)."""
        if out is None:
            out = np.empty_like(particles_in)
        self._check(self._lib.clsph_simulate_single_frame(self._h, particle_ptr(particles_in), particle_ptr(out),
                                                          ctypes.byref(params),
                                                          ctypes.byref(terms) if terms is not None else None))
        return out

    def kernel_advection_collision(self, particles_in, out=None):
        """The reference's advection_collision kernel alone (input carries the acceleration)."""
        if out is None:
            out = np.empty_like(particles_in)
        self._check(self._lib.clsph_kernel_advection_collision(self._h, particle_ptr(particles_in), particle_ptr(out),
                                                               particles_in.size))
        return out

    def dist_init(self, rank, world, unique_id, plane_lo, plane_hi, emigrant_capacity=0, ghost_capacity=0):
        """Join the slab decomposition: this rank owns plane_lo <= x < plane_hi (snapped to cells)."""
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        self._check(self._lib.clsph_dist_init(self._h, rank, world, buf, plane_lo, plane_hi, emigrant_capacity,
                                              ghost_capacity))

    def frame_points(self):
        """(n, 7) float32: position, velocity, density of every particle in the reference's output order, packed on
        the device (clsph_frame_begin / clsph_frame_end)."""
        import numpy as np
        out = np.empty((self.particle_count(), 7), dtype=np.float32)
        self._check(self._lib.clsph_frame_begin(self._h, _vp(out), out.shape[0]))
        self._check(self._lib.clsph_frame_end(self._h))
        return out

    def dist_transport(self):
        """How the ranks exchange particles: "peer stores ..." (NVLink, no collective per sub-step) or "nccl ..."."""
        return self._lib.clsph_dist_transport(self._h).decode()

    def dist_upload(self, particles, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        self._check(self._lib.clsph_dist_upload(self._h, particle_ptr(particles), _vp(ids), particles.size))

    def dist_count(self):
        n = ctypes.c_uint32()
        self._check(self._lib.clsph_dist_download(self._h, None, None, 0, ctypes.byref(n)))
        return n.value

    def dist_download(self):
        """(owned particles in local sorted order, their ids)."""
        n = self.dist_count()
        out = np.empty(n, dtype=PARTICLE)
        ids = np.empty(n, dtype=np.uint32)
        got = ctypes.c_uint32()
        self._check(self._lib.clsph_dist_download(self._h, particle_ptr(out), _vp(ids), n, ctypes.byref(got)))
        return out[: got.value], ids[: got.value]

    def set_debug(self, enable=True):
        self._check(self._lib.clsph_set_debug(self._h, 1 if enable else 0))

    def fetch(self, what):
        n = self.particle_count()
        if what == TAP_CELL_TABLE:
            arr = np.empty(self.parameters().grid_cell_count, dtype=np.uint32)
        elif what in (TAP_DENSITY, TAP_PRESSURE):
            arr = np.empty(n, dtype=np.float32)
        elif what == TAP_ACCELERATION:
            arr = np.empty((n, 3), dtype=np.float32)
        else:
            arr = np.empty(n, dtype=np.uint32)
        self._check(self._lib.clsph_debug_fetch(self._h, what, _vp(arr), arr.nbytes))
        return arr

    def profile_enable(self, enable=True):
        self._check(self._lib.clsph_profile_enable(self._h, 1 if enable else 0))

    def profile_read(self):
        t = StageTimes()
        self._check(self._lib.clsph_profile_read(self._h, ctypes.byref(t)))
        return t.as_dict()

    def stream(self):
        return self._lib.clsph_stream(self._h)
