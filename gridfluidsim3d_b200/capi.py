"""ctypes binding of include/gfs_b200.h (gridfluidsim3d_b200/libgfs_b200.so).

This is the reference-side binding a maintainer would write (cf. the reference's own
src/pyfluid/pybindings.py:11-36 over its C bindings).  There is no fallback of any kind: if the shared
library is missing or no CUDA device is usable, construction raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgfs_b200.so")

TRILINEAR, TRICUBIC = 0, 1
FAST, EXACT = 0, 1
FIELD_NEW, FIELD_SAVED, FIELD_P2G = 0, 1, 2
AIR, FLUID, SOLID = 0, 1, 2

#: every symbol include/gfs_b200.h declares (tests check the .so exports each one)
SYMBOLS = [
    "gfs_get_error_message", "gfs_create", "gfs_destroy", "gfs_device_info", "gfs_sync", "gfs_get_stats",
    "gfs_profile_enable", "gfs_profile_read",
    "gfs_sample", "gfs_advect", "gfs_add_point_values", "gfs_add_points",
    "gfs_domain_init", "gfs_set_material", "gfs_get_material", "gfs_get_fluid_cells", "gfs_set_sources",
    "gfs_emit_from_sources", "gfs_remove_in_sources", "gfs_set_particles", "gfs_num_particles", "gfs_get_particles", "gfs_get_particle_order",
    "gfs_set_field", "gfs_get_field", "gfs_set_field_layers", "gfs_get_field_layers", "gfs_get_material_layers", "gfs_sort", "gfs_sort_unstable", "gfs_set_option", "gfs_p2g", "gfs_g2p_advect", "gfs_substep", "gfs_advect_substep",
    "gfs_set_owned_layers", "gfs_p2g_begin", "gfs_p2g_end", "gfs_layer_bytes", "gfs_pack_layers", "gfs_unpack_layers",
    "gfs_copy_layers_batch", "gfs_extract_particles", "gfs_extract_particles_async", "gfs_extract_commit", "gfs_append_particles_device",
    "gfs_comm_alloc", "gfs_comm_export", "gfs_comm_connect", "gfs_comm_connect_local", "gfs_comm_push_layers",
    "gfs_comm_pull_layers", "gfs_comm_migrate_begin", "gfs_comm_migrate_finish", "gfs_comm_g2p_advect",
    "gfs_comm_world_alloc", "gfs_comm_world_export", "gfs_comm_world_connect", "gfs_comm_world_connect_local",
    "gfs_comm_allmax_scale", "gfs_comm_allmax_post", "gfs_sort_index", "gfs_comm_set_plan", "gfs_comm_substep", "gfs_extrapolate", "gfs_copy_field", "gfs_extrapolate_field",
    "gfs_apply_body_force", "gfs_pressure_solve", "gfs_apply_pressure", "gfs_get_pressure", "gfs_pressure_solve_field",
    "gfs_device_ptr", "gfs_resize_particles", "gfs_reserve", "gfs_state_hash",
    "gfs_mg_get_error_message", "gfs_mg_create", "gfs_mg_destroy", "gfs_mg_num_devices", "gfs_mg_context", "gfs_mg_get_slab", "gfs_mg_set_option",
    "gfs_mg_set_material", "gfs_mg_set_sources", "gfs_mg_set_field", "gfs_mg_get_field", "gfs_mg_get_material", "gfs_mg_scatter_particles",
    "gfs_mg_num_particles", "gfs_mg_gather_particles", "gfs_mg_substep", "gfs_mg_state_hash", "gfs_mg_sync", "gfs_slab_range", "gfs_slab_owner", "gfs_slab_halo_cells",
]


class GfsError(RuntimeError):
    pass


class Source(C.Structure):
    _fields_ = [("kind", C.c_int), ("p", C.c_float * 3), ("a", C.c_double), ("b", C.c_double),
                ("c", C.c_double), ("velocity", C.c_float * 3)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("num_particles", "out_of_grid", "in_solid", "solid_hits",
                                         "fluid_cells", "kernel_launches", "graph_replays", "removed_particles",
                                         "collision_overflow")]


_f32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8 = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_err = C.POINTER(C.c_int)
_lib = None


def load_library():
    """Load libgfs_b200.so and declare prototypes.  Raises if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GfsError("%s is missing: build it with `make -C gridfluidsim3d_b200/csrc` (or "
                       "__graft_entry__.build()).  There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    V, I, D, L64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
    L.gfs_get_error_message.restype = C.c_char_p
    L.gfs_create.restype = V
    L.gfs_create.argtypes = [I, V, _err]
    L.gfs_destroy.argtypes = [V, _err]
    L.gfs_device_info.argtypes = [V, C.c_char_p, I, _err]
    L.gfs_sync.argtypes = [V, _err]
    L.gfs_get_stats.argtypes = [V, C.POINTER(Stats), _err]
    L.gfs_profile_enable.argtypes = [V, I, _err]
    L.gfs_profile_read.argtypes = [V, C.c_char_p, C.POINTER(D), C.POINTER(L64), I, I, _err]
    L.gfs_profile_read.restype = I
    L.gfs_sample.argtypes = [V, _f32, L64, _f32, _f32, _f32, I, I, I, D, I, I, I, _f32, _err]
    L.gfs_advect.argtypes = [V, _f32, L64, _f32, _f32, _f32, I, I, I, D, D, I, I, I, _f32, _err]
    L.gfs_add_point_values.argtypes = [V, _f32, _f32, L64, D, _f32, D, I, I, I, _f32, V, I, I, _err]
    L.gfs_add_points.argtypes = [V, _f32, L64, D, _f32, D, I, I, I, _f32, I, I, C.c_float, I, _err]
    L.gfs_state_hash.argtypes = [V, C.POINTER(C.c_uint64), _err]
    L.gfs_domain_init.argtypes = [V, I, I, I, D, _err]
    L.gfs_set_material.argtypes = [V, _u8, _err]
    L.gfs_get_material.argtypes = [V, _u8, _err]
    L.gfs_set_sources.argtypes = [V, V, I, _err]
    L.gfs_emit_from_sources.argtypes = [V, D, C.c_uint64, C.POINTER(L64), _err]
    L.gfs_remove_in_sources.argtypes = [V, V, I, C.POINTER(L64), _err]
    L.gfs_set_particles.argtypes = [V, _f32, L64, _err]
    L.gfs_num_particles.argtypes = [V, _err]
    L.gfs_num_particles.restype = L64
    L.gfs_get_particles.argtypes = [V, _f32, _err]
    L.gfs_get_particle_order.argtypes = [V, _i32, _err]
    L.gfs_set_field.argtypes = [V, I, _f32, _f32, _f32, _err]
    L.gfs_get_field.argtypes = [V, I, _f32, _f32, _f32, _err]
    L.gfs_set_field_layers.argtypes = [V, I, _f32, _f32, _f32, I, I, _err]
    L.gfs_get_field_layers.argtypes = [V, I, _f32, _f32, _f32, I, I, _err]
    L.gfs_get_material_layers.argtypes = [V, _u8, I, I, _err]
    L.gfs_sort.argtypes = [V, _err]
    L.gfs_sort_unstable.argtypes = [V, _err]
    L.gfs_set_option.argtypes = [V, I, I, _err]
    L.gfs_p2g.argtypes = [V, I, _err]
    L.gfs_g2p_advect.argtypes = [V, D, D, I, I, I, _err]
    L.gfs_substep.argtypes = [V, D, D, I, I, I, _err]
    L.gfs_advect_substep.argtypes = [V, D, I, _err]
    L.gfs_set_owned_layers.argtypes = [V, I, I, _err]
    L.gfs_p2g_begin.argtypes = [V, I, _err]
    L.gfs_p2g_end.argtypes = [V, _err]
    L.gfs_layer_bytes.argtypes = [V, I, _err]
    L.gfs_layer_bytes.restype = L64
    L.gfs_pack_layers.argtypes = [V, I, I, I, V, _err]
    L.gfs_unpack_layers.argtypes = [V, I, I, I, V, I, _err]
    L.gfs_copy_layers_batch.argtypes = [V, I, I, C.POINTER(I), C.POINTER(I), C.POINTER(I), C.POINTER(L64), C.POINTER(I), V, _err]
    L.gfs_extract_particles.argtypes = [V, I, I, V, V, L64, C.POINTER(L64), C.POINTER(L64), _err]
    L.gfs_extract_particles_async.argtypes = [V, I, I, V, V, L64, V, _err]
    L.gfs_extract_commit.argtypes = [V, L64, _err]
    L.gfs_append_particles_device.argtypes = [V, V, L64, _err]
    L.gfs_comm_alloc.argtypes = [V, L64, L64, _err]
    L.gfs_comm_export.argtypes = [V, I, C.c_char_p, _err]
    L.gfs_comm_connect.argtypes = [V, I, C.c_char_p, _err]
    L.gfs_comm_connect_local.argtypes = [V, I, V, _err]
    L.gfs_comm_push_layers.argtypes = [V, I, I, C.POINTER(I), C.POINTER(I), C.POINTER(I), C.POINTER(L64), _err]
    L.gfs_comm_pull_layers.argtypes = [V, I, I, C.POINTER(I), C.POINTER(I), C.POINTER(I), C.POINTER(L64), C.POINTER(I), _err]
    L.gfs_comm_migrate_begin.argtypes = [V, I, I, _err]
    L.gfs_comm_migrate_finish.argtypes = [V, C.POINTER(L64), _err]
    L.gfs_comm_g2p_advect.argtypes = [V, C.c_double, C.c_double, I, I, I, I, I, _err]
    L.gfs_comm_world_alloc.argtypes = [V, I, I, _err]
    L.gfs_comm_world_export.argtypes = [V, C.c_char_p, _err]
    L.gfs_comm_world_connect.argtypes = [V, I, C.c_char_p, _err]
    L.gfs_comm_world_connect_local.argtypes = [V, I, V, _err]
    L.gfs_comm_allmax_scale.argtypes = [V, _err]
    L.gfs_comm_allmax_post.argtypes = [V, _err]
    L.gfs_sort_index.argtypes = [V, _err]
    L.gfs_extrapolate.argtypes = [V, I, I, _err]
    L.gfs_extrapolate_field.argtypes = [V, _f32, _f32, _f32, I, I, I, _u8, I, _err]
    L.gfs_copy_field.argtypes = [V, I, I, _err]
    L.gfs_get_fluid_cells.argtypes = [V, C.c_void_p, L64, C.POINTER(L64), _err]
    L.gfs_apply_body_force.argtypes = [V, I, C.c_float, C.c_float, C.c_float, C.c_double, _err]
    L.gfs_pressure_solve.argtypes = [V, I, C.c_double, C.c_double, C.c_double, I, C.POINTER(I), C.POINTER(C.c_double), _err]
    L.gfs_apply_pressure.argtypes = [V, I, I, C.c_double, C.c_double, _err]
    L.gfs_get_pressure.argtypes = [V, _f32, _err]
    L.gfs_pressure_solve_field.argtypes = [V, _f32, _f32, _f32, I, I, I, C.c_double, _u8, C.c_double, C.c_double, C.c_double, I,
                                           np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS"), C.POINTER(I), C.POINTER(C.c_double), _err]
    PI, PL = C.POINTER(I), C.POINTER(L64)
    L.gfs_comm_set_plan.argtypes = [V, I, I, PI, PI, PI, PL, I, PI, PI, PI, PL, PI, _err]
    L.gfs_comm_substep.argtypes = [V, C.c_double, C.c_double, I, I, I, I, I, PL, _err]
    L.gfs_device_ptr.argtypes = [V, I, _err]
    L.gfs_device_ptr.restype = V
    L.gfs_resize_particles.argtypes = [V, L64, _err]
    L.gfs_reserve.argtypes = [V, L64, _err]
    L.gfs_slab_range.argtypes = [I, I, I, C.POINTER(I), C.POINTER(I), _err]
    L.gfs_slab_owner.argtypes = [I, I, I, _err]
    L.gfs_slab_owner.restype = I
    L.gfs_slab_halo_cells.argtypes = [I, D, D, _err]
    L.gfs_slab_halo_cells.restype = I
    _lib = L
    return L


def _check(err):
    if err.value != 1:
        raise GfsError(load_library().gfs_get_error_message().decode("utf-8", "replace"))


def _c(a, dt=np.float32):
    return np.ascontiguousarray(a, dtype=dt)


def face_counts(dims):
    I, J, K = dims
    return (I + 1) * J * K, I * (J + 1) * K, I * J * (K + 1)


PICFLIP_RATIO = float(np.float32(0.05))      # double _ratioPICFLIP = 0.05f  (src/fluidsimulation.h:1161)


# ---- z-slab arithmetic (host only; usable without a GPU) ----------------------------------------
def slab_range(ksize, nranks, rank):
    L = load_library()
    k0, k1, err = C.c_int(), C.c_int(), C.c_int()
    L.gfs_slab_range(ksize, nranks, rank, C.byref(k0), C.byref(k1), C.byref(err))
    _check(err)
    return k0.value, k1.value


def slab_owner(ksize, nranks, k):
    L = load_library()
    err = C.c_int()
    r = L.gfs_slab_owner(ksize, nranks, int(k), C.byref(err))
    _check(err)
    return r


def slab_halo_cells(interp, max_displacement, dx):
    L = load_library()
    err = C.c_int()
    r = L.gfs_slab_halo_cells(interp, max_displacement, dx, C.byref(err))
    _check(err)
    return r


class Context:
    """One CUDA device + stream.  Mirrors the accelerator objects FluidSimulation owns
    (ParticleAdvector / CLScalarField, src/fluidsimulation.h:1170-1171)."""

    def __init__(self, device=0, stream=None):
        self.lib = load_library()
        err = C.c_int()
        self.h = self.lib.gfs_create(device, stream, C.byref(err))
        _check(err)
        self.dims = None
        self.dx = None

    def close(self):
        if getattr(self, "h", None):
            err = C.c_int()
            self.lib.gfs_destroy(self.h, C.byref(err))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, fn, *args):
        err = C.c_int()
        r = fn(self.h, *args, C.byref(err))
        _check(err)
        return r

    def device_info(self):
        buf = C.create_string_buffer(512)
        self._call(self.lib.gfs_device_info, buf, 512)
        return buf.value.decode()

    def sync(self):
        self._call(self.lib.gfs_sync)

    def stats(self):
        s = Stats()
        self._call(self.lib.gfs_get_stats, C.byref(s))
        return {n: getattr(s, n) for n, _ in Stats._fields_}

    def profile_enable(self, on=True):
        self._call(self.lib.gfs_profile_enable, int(on))

    def profile_read(self, reset=True):
        """{kernel name: (total milliseconds, launches)} since the last reset (CUDA events, device time)."""
        cap = 64
        names = C.create_string_buffer(64 * cap)
        ms, cnt = (C.c_double * cap)(), (C.c_int64 * cap)()
        n = self._call(self.lib.gfs_profile_read, names, ms, cnt, cap, int(reset))
        return {names.raw[64 * i:64 * (i + 1)].split(b"\0")[0].decode(): (ms[i], cnt[i]) for i in range(n)}

    # ---- host-pointer operators (ParticleAdvector / CLScalarField) --------------------------------
    def sample(self, pos, u, v, w, dims, dx, interp=TRICUBIC, arith=FAST, validate=True):
        pos = _c(pos)
        out = np.empty_like(pos)
        self._call(self.lib.gfs_sample, pos, len(pos), _c(u), _c(v), _c(w), *dims, dx, interp, arith, int(validate), out)
        return out

    def advect(self, pos, u, v, w, dims, dx, dt, order=4, interp=TRICUBIC, arith=FAST):
        pos = _c(pos)
        out = np.empty_like(pos)
        self._call(self.lib.gfs_advect, pos, len(pos), _c(u), _c(v), _c(w), *dims, dx, dt, order, interp, arith, out)
        return out

    def add_point_values(self, pos, values, radius, offset, dx, ndims, field=None, weight=None,
                         accumulate=False, arith=FAST, with_weight=True):
        pos, values = _c(pos), _c(values)
        cnt = ndims[0] * ndims[1] * ndims[2]
        field = np.zeros(cnt, np.float32) if field is None else field
        if with_weight and weight is None:
            weight = np.zeros(cnt, np.float32)
        wptr = weight.ctypes.data_as(C.c_void_p) if weight is not None else None
        self._call(self.lib.gfs_add_point_values, pos, values, len(pos), radius, _c(offset), dx, *ndims,
                   field, wptr, int(accumulate), arith)
        return field, weight

    def add_points(self, pos, radius, offset, dx, ndims, field=None, accumulate=True, threshold=None, arith=FAST):
        """CLScalarField::addPoints; threshold = the mesher's max-scalar-field-value threshold (None: not set)."""
        pos = _c(pos)
        field = np.zeros(ndims[0] * ndims[1] * ndims[2], np.float32) if field is None else field
        self._call(self.lib.gfs_add_points, pos, len(pos), radius, _c(offset), dx, *ndims, field, int(accumulate),
                   int(threshold is not None), float(threshold or 0.0), arith)
        return field

    def advect_substep(self, dt, order=4):
        """Index sort + RK advection of the resident positions through field NEW (trilinear brick kernel, positions only)."""
        self._call(self.lib.gfs_advect_substep, float(dt), int(order))

    def state_hash(self):
        """-> [material, p2g_u, p2g_v, p2g_w, particles] order/distribution-independent 64-bit hashes (ints)."""
        out = (C.c_uint64 * 5)()
        self._call(self.lib.gfs_state_hash, out)
        return [int(v) for v in out]

    # ---- device-resident domain ---------------------------------------------------------------------
    def domain_init(self, dims, dx):
        self.dims, self.dx = tuple(int(d) for d in dims), float(dx)
        self._call(self.lib.gfs_domain_init, *self.dims, self.dx)

    def set_material(self, material):
        self._call(self.lib.gfs_set_material, _c(material, np.uint8))

    def get_material(self, out=None):
        m = np.empty(self.dims[0] * self.dims[1] * self.dims[2], np.uint8) if out is None else out
        self._call(self.lib.gfs_get_material, m)
        return m

    def set_sources(self, sources):
        arr = (Source * max(1, len(sources)))()
        for s, d in zip(arr, sources):
            s.kind = d["kind"]
            s.p[:] = d["p"]
            s.a, s.b, s.c = d.get("a", 0.0), d.get("b", 0.0), d.get("c", 0.0)
            s.velocity[:] = d["velocity"]
        self._call(self.lib.gfs_set_sources, C.cast(arr, C.c_void_p), len(sources))

    @staticmethod
    def _source_array(sources):
        arr = (Source * max(1, len(sources)))()
        for s, d in zip(arr, sources):
            s.kind = d["kind"]
            s.p[:] = d["p"]
            s.a, s.b, s.c = d.get("a", 0.0), d.get("b", 0.0), d.get("c", 0.0)
            s.velocity[:] = d.get("velocity", (0.0, 0.0, 0.0))
        return arr

    def emit_from_sources(self, jitter, seed=1):
        """FluidSimulation::_updateFluidSources for the active inflow sources of set_sources -> particles added."""
        n = C.c_int64(0)
        self._call(self.lib.gfs_emit_from_sources, float(jitter), int(seed), C.byref(n))
        return int(n.value)

    def remove_in_sources(self, sources):
        """The outflow half: particles in the fluid cells the given sources overlap are removed -> how many."""
        n = C.c_int64(0)
        self._call(self.lib.gfs_remove_in_sources, C.cast(self._source_array(sources), C.c_void_p), len(sources), C.byref(n))
        return int(n.value)

    def set_particles(self, pos, vel):
        """pos, vel: (N,3) float32 -> MarkerParticle_t AoS (24 B) -> device SoA."""
        aos = np.ascontiguousarray(np.concatenate([_c(pos), _c(vel)], axis=1))
        self.set_particles_aos(aos)

    def set_particles_aos(self, aos):
        aos = _c(aos).reshape(-1)
        self._call(self.lib.gfs_set_particles, aos, aos.size // 6)

    @property
    def num_particles(self):
        return self._call(self.lib.gfs_num_particles)

    def get_particles_aos(self, out=None):
        n = self.num_particles
        out = np.empty(n * 6, np.float32) if out is None else out
        self._call(self.lib.gfs_get_particles, out)
        return out

    def get_particles(self):
        aos = self.get_particles_aos().reshape(-1, 6)
        return np.ascontiguousarray(aos[:, :3]), np.ascontiguousarray(aos[:, 3:])

    def get_particle_order(self):
        o = np.empty(self.num_particles, np.int32)
        self._call(self.lib.gfs_get_particle_order, o)
        return o

    def set_field(self, slot, u, v, w):
        self._call(self.lib.gfs_set_field, slot, _c(u), _c(v), _c(w))

    def get_field(self, slot, out=None):
        if out is None:
            nu, nv, nw = face_counts(self.dims)
            out = (np.empty(nu, np.float32), np.empty(nv, np.float32), np.empty(nw, np.float32))
        self._call(self.lib.gfs_get_field, slot, *out)
        return out

    def set_field_layers(self, slot, u, v, w, k_first, k_count):
        """Upload only cell layers [k_first, k_first + k_count) of the caller's WHOLE arrays (z-slab ranks)."""
        self._call(self.lib.gfs_set_field_layers, slot, _c(u), _c(v), _c(w), int(k_first), int(k_count))

    def get_field_layers(self, slot, out, k_first, k_count):
        self._call(self.lib.gfs_get_field_layers, slot, *out, int(k_first), int(k_count))
        return out

    def get_material_layers(self, out, k_first, k_count):
        self._call(self.lib.gfs_get_material_layers, out, int(k_first), int(k_count))
        return out

    def sort(self):
        self._call(self.lib.gfs_sort)

    def sort_unstable(self):
        self._call(self.lib.gfs_sort_unstable)

    def extrapolate_field(self, u, v, w, dims, material, num_layers):
        """Host-pointer operator: returns extrapolated copies of u, v, w."""
        u, v, w = [np.array(a, np.float32, copy=True).reshape(-1) for a in (u, v, w)]
        self._call(self.lib.gfs_extrapolate_field, u, v, w, *dims, np.ascontiguousarray(material, np.uint8).reshape(-1), int(num_layers))
        return u, v, w

    def extrapolate(self, slot, num_layers):
        self._call(self.lib.gfs_extrapolate, int(slot), int(num_layers))

    def get_fluid_cells(self):
        """FluidSimulation::_fluidCellIndices: (n, 3) int32 array of (i, j, k), in the reference's k, j, i scan order."""
        cnt = C.c_int64(0)
        self._call(self.lib.gfs_get_fluid_cells, None, 0, C.byref(cnt))
        out = np.empty((cnt.value, 3), np.int32)
        if cnt.value:
            self._call(self.lib.gfs_get_fluid_cells, out.ctypes.data, cnt.value, C.byref(cnt))
        return out

    def copy_field(self, dst_slot, src_slot):
        self._call(self.lib.gfs_copy_field, int(dst_slot), int(src_slot))

    # -- stages 6-8 on the resident grid (FluidSimulation::_applyConstantBodyForces, PressureSolver::solve,
    #    _applyPressureToVelocityField; reference defaults: density 20, tolerance 1e-6, 200 iterations)
    def apply_body_force(self, slot, force, dt):
        self._call(self.lib.gfs_apply_body_force, int(slot), float(force[0]), float(force[1]), float(force[2]), float(dt))

    def pressure_solve(self, slot, dt, density=20.0, tolerance=1e-6, max_iterations=200):
        """-> (iterations, last residual max-norm); iterations as the reference counts them (-1: nothing to solve)."""
        it, res = C.c_int(0), C.c_double(0.0)
        self._call(self.lib.gfs_pressure_solve, int(slot), float(dt), float(density), float(tolerance), int(max_iterations),
                   C.byref(it), C.byref(res))
        return it.value, res.value

    def apply_pressure(self, src_slot, dst_slot, dt, density=20.0):
        self._call(self.lib.gfs_apply_pressure, int(src_slot), int(dst_slot), float(dt), float(density))

    def get_pressure(self):
        out = np.empty(int(np.prod(self.dims)), np.float32)
        self._call(self.lib.gfs_get_pressure, out)
        return out

    def pressure_solve_field(self, u, v, w, dims, dx, material, dt, density=20.0, tolerance=1e-6, max_iterations=200):
        """Host-pointer operator (the body of PressureSolver::solve): -> (double pressure per cell, iterations, residual)."""
        out = np.zeros(int(np.prod(dims)), np.float64)
        it, res = C.c_int(0), C.c_double(0.0)
        self._call(self.lib.gfs_pressure_solve_field, _c(u), _c(v), _c(w), *dims, float(dx),
                   np.ascontiguousarray(material, np.uint8).reshape(-1), float(dt), float(density), float(tolerance),
                   int(max_iterations), out, C.byref(it), C.byref(res))
        return out, it.value, res.value

    def sort_index(self):
        self._call(self.lib.gfs_sort_index)

    def set_option(self, option, value):
        self._call(self.lib.gfs_set_option, option, value)

    def p2g(self, arith=FAST):
        self._call(self.lib.gfs_p2g, arith)

    def g2p_advect(self, dt, ratio=PICFLIP_RATIO, order=4, interp=TRICUBIC, arith=FAST):
        self._call(self.lib.gfs_g2p_advect, dt, ratio, order, interp, arith)

    def substep(self, dt, ratio=PICFLIP_RATIO, order=4, interp=TRICUBIC, arith=FAST):
        self._call(self.lib.gfs_substep, dt, ratio, order, interp, arith)

    # ---- z-slab sharding primitives (device buffers are raw pointers, e.g. torch tensor.data_ptr()) -----
    def set_owned_layers(self, k0, k1):
        self._call(self.lib.gfs_set_owned_layers, int(k0), int(k1))

    def p2g_begin(self, arith=FAST):
        self._call(self.lib.gfs_p2g_begin, arith)

    def p2g_end(self):
        self._call(self.lib.gfs_p2g_end)

    def layer_bytes(self, what):
        return self._call(self.lib.gfs_layer_bytes, what)

    def pack_layers(self, what, k_first, k_count, dst_ptr):
        self._call(self.lib.gfs_pack_layers, what, int(k_first), int(k_count), dst_ptr)

    def unpack_layers(self, what, k_first, k_count, src_ptr, add=False):
        self._call(self.lib.gfs_unpack_layers, what, int(k_first), int(k_count), src_ptr, int(add))

    def copy_layers_batch(self, direction, items, buffer_ptr):
        """items: [(what, k_first, k_count, byte offset, add)] -- one launch for all of them"""
        n = len(items)
        IA, LA = C.c_int * n, C.c_int64 * n
        self._call(self.lib.gfs_copy_layers_batch, int(direction), n, IA(*[int(i[0]) for i in items]), IA(*[int(i[1]) for i in items]),
                   IA(*[int(i[2]) for i in items]), LA(*[int(i[3]) for i in items]), IA(*[int(bool(i[4])) for i in items]), buffer_ptr)

    def extract_particles(self, k_lo, k_hi, down_ptr, up_ptr, cap):
        nd, nu = C.c_int64(), C.c_int64()
        self._call(self.lib.gfs_extract_particles, int(k_lo), int(k_hi), down_ptr, up_ptr, int(cap), C.byref(nd), C.byref(nu))
        return nd.value, nu.value

    def extract_particles_async(self, k_lo, k_hi, down_ptr, up_ptr, cap, counters_ptr):
        self._call(self.lib.gfs_extract_particles_async, int(k_lo), int(k_hi), down_ptr, up_ptr, int(cap), counters_ptr)

    def extract_commit(self, n_kept):
        self._call(self.lib.gfs_extract_commit, int(n_kept))

    def append_particles_device(self, aos_ptr, n):
        self._call(self.lib.gfs_append_particles_device, aos_ptr, int(n))

    # ---- peer-memory exchange (CUDA IPC) ---------------------------------------------------------------------
    def comm_alloc(self, layer_bytes, particle_cap):
        self._call(self.lib.gfs_comm_alloc, int(layer_bytes), int(particle_cap))

    def comm_export(self, side):
        buf = C.create_string_buffer(64)
        self._call(self.lib.gfs_comm_export, int(side), buf)
        return buf.raw

    def comm_connect(self, side, handle):
        self._call(self.lib.gfs_comm_connect, int(side), C.create_string_buffer(handle, 64))

    def comm_connect_local(self, side, other):
        self._call(self.lib.gfs_comm_connect_local, int(side), other.h)

    @staticmethod
    def _item_arrays(items):
        n = len(items)
        IA, LA = C.c_int * n, C.c_int64 * n
        return (n, IA(*[int(i[0]) for i in items]), IA(*[int(i[1]) for i in items]), IA(*[int(i[2]) for i in items]),
                LA(*[int(i[3]) for i in items]), IA(*[int(bool(i[4])) for i in items]))

    def comm_push_layers(self, side, items):
        n, w, f, k, o, _ = self._item_arrays(items)
        self._call(self.lib.gfs_comm_push_layers, int(side), n, w, f, k, o)

    def comm_pull_layers(self, side, items):
        n, w, f, k, o, a = self._item_arrays(items)
        self._call(self.lib.gfs_comm_pull_layers, int(side), n, w, f, k, o, a)

    def comm_migrate_begin(self, has_down, has_up):
        self._call(self.lib.gfs_comm_migrate_begin, int(bool(has_down)), int(bool(has_up)))

    def comm_g2p_advect(self, dt, has_down, has_up, ratio=0.05, order=4, interp=TRICUBIC, arith=FAST):
        self._call(self.lib.gfs_comm_g2p_advect, float(dt), float(ratio), int(order), int(interp), int(arith),
                   int(bool(has_down)), int(bool(has_up)))

    def comm_set_plan(self, side, push_items, pull_items):
        n, w, f, k, o, _ = self._item_arrays(push_items)
        m, w2, f2, k2, o2, a2 = self._item_arrays(pull_items)
        self._call(self.lib.gfs_comm_set_plan, int(side), n, w, f, k, o, m, w2, f2, k2, o2, a2)

    def comm_substep(self, dt, has_down, has_up, ratio=0.05, order=4, interp=TRICUBIC, arith=FAST):
        moved = (C.c_int64 * 2)()
        self._call(self.lib.gfs_comm_substep, float(dt), float(ratio), int(order), int(interp), int(arith),
                   int(bool(has_down)), int(bool(has_up)), moved)
        return moved[0], moved[1]

    def comm_world_alloc(self, rank, world):
        self._call(self.lib.gfs_comm_world_alloc, int(rank), int(world))

    def comm_world_export(self):
        buf = C.create_string_buffer(64)
        self._call(self.lib.gfs_comm_world_export, buf)
        return buf.raw

    def comm_world_connect(self, rank, handle):
        self._call(self.lib.gfs_comm_world_connect, int(rank), C.create_string_buffer(handle, 64))

    def comm_world_connect_local(self, rank, other):
        self._call(self.lib.gfs_comm_world_connect_local, int(rank), other.h)

    def comm_allmax_scale(self):
        self._call(self.lib.gfs_comm_allmax_scale)

    def comm_migrate_finish(self):
        moved = (C.c_int64 * 2)()
        self._call(self.lib.gfs_comm_migrate_finish, moved)
        return moved[0], moved[1]

    def resize_particles(self, n):
        self._call(self.lib.gfs_resize_particles, int(n))

    def device_ptr(self, which):
        return self._call(self.lib.gfs_device_ptr, which)
