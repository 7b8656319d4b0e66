"""ctypes front-ends for the two CPU checkers -- TEST INFRASTRUCTURE ONLY.

* ``Oracle``    -> oracle/liboracle.so   (plain-C restatement, oracle.c; exists everywhere)
* ``Reference`` -> oracle/_ref/libgfsref.so (the unmodified reference sources + ref_harness.cpp;
  built only where /root/reference exists, but the built .so travels to the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (gridfluidsim3d_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libgfsref.so")
DROPIN_SO = os.path.join(HERE, "_ref", "libgfsref_dropin.so")   # the reference simulator + the CUDA drop-in classes
RESIDENT_SO = os.path.join(HERE, "_ref", "libgfsref_resident.so")   # ... + stages 1/5/11/12 of _stepFluid device resident

_f32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8 = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


class Source(C.Structure):
    _fields_ = [("kind", C.c_int), ("p", C.c_float * 3), ("a", C.c_double), ("b", C.c_double),
                ("c", C.c_double), ("velocity", C.c_float * 3)]


def make_sources(sources):
    """sources: list of dicts {kind, p, a, b, c, velocity} -> (ctypes array, count)."""
    arr = (Source * max(1, len(sources)))()
    for s, d in zip(arr, sources):
        s.kind = d["kind"]
        s.p[:] = d["p"]
        s.a, s.b, s.c = d.get("a", 0.0), d.get("b", 0.0), d.get("c", 0.0)
        s.velocity[:] = d["velocity"]
    return arr, len(sources)


def build_oracle(force=False):
    if force or not os.path.exists(ORACLE_SO) or \
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return ORACLE_SO


def build_reference(force=False):
    """Build oracle/_ref/libgfsref.so if the reference tree is present; return path or None."""
    if os.path.isdir("/root/reference/src"):
        if force or not os.path.exists(REF_SO) or \
                os.path.getmtime(REF_SO) < os.path.getmtime(os.path.join(HERE, "ref_harness.cpp")):
            subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref"])
    return REF_SO if os.path.exists(REF_SO) else None


def face_dims(I, J, K):
    return (I + 1, J, K), (I, J + 1, K), (I, J, K + 1)


def face_counts(I, J, K):
    return tuple(a * b * c for a, b, c in face_dims(I, J, K))


def _c(a, dt=np.float32):
    return np.ascontiguousarray(a, dtype=dt)


class Oracle:
    def __init__(self):
        self.lib = L = C.CDLL(build_oracle())
        L.orc_max_threads.restype = C.c_int
        L.orc_cell_index.argtypes = [_f32, C.c_long, C.c_double, _i32]
        L.orc_sample.argtypes = [_f32, C.c_long, _f32, _f32, _f32, C.c_int, C.c_int, C.c_int, C.c_double,
                                 C.c_int, C.c_int, _f32]
        L.orc_advect.argtypes = [_f32, C.c_long, _f32, _f32, _f32, C.c_int, C.c_int, C.c_int, C.c_double,
                                 C.c_double, C.c_int, C.c_int, _f32]
        L.orc_picflip.argtypes = [_f32, _f32, C.c_long, _f32, _f32, _f32, _f32, _f32, _f32,
                                  C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, _f32]
        L.orc_splat.argtypes = [_f32, _f32, C.c_long, C.c_long, C.c_double, _f32, C.c_double,
                                C.c_int, C.c_int, C.c_int, _f32, _f32]
        L.orc_apply_weight.argtypes = [_f32, _f32, C.c_long]
        L.orc_border_solid.argtypes = [C.c_int, C.c_int, C.c_int, _u8]
        L.orc_classify.argtypes = [_f32, C.c_long, C.c_int, C.c_int, C.c_int, C.c_double, _u8]
        L.orc_classify.restype = C.c_long
        L.orc_p2g_component.argtypes = [_f32, _f32, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                        _u8, C.c_void_p, C.c_int, _f32]
        L.orc_finish_component.argtypes = [_f32, _f32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _u8, C.c_void_p, C.c_int, _f32]
        L.orc_p2g.argtypes = [_f32, _f32, C.c_long, C.c_int, C.c_int, C.c_int, C.c_double,
                              _u8, C.c_void_p, C.c_int, _f32, _f32, _f32]
        L.orc_solid_test.argtypes = [_f32, _f32, C.c_long, C.c_int, C.c_int, C.c_int, C.c_double, _u8, _u8]
        L.orc_solid_test.restype = C.c_long
        L.orc_g2p_advect.argtypes = [_f32, _f32, C.c_long, _f32, _f32, _f32, _f32, _f32, _f32,
                                     C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                     C.c_int, C.c_int, C.c_void_p, _f32, _f32, C.c_void_p]

    def cell_index(self, pos, dx):
        pos = _c(pos)
        out = np.empty((len(pos), 3), np.int32)
        self.lib.orc_cell_index(pos, len(pos), dx, out)
        return out

    def collide(self, p0, p1, dims, dx, material):
        """The post-advection pass with the collision resolve; returns (positions, flags)."""
        p0, p1 = _c(p0), np.array(p1, np.float32, copy=True)
        flags = np.zeros(len(p0), np.uint8)
        self.lib.orc_collide.argtypes = [_f32, _f32, C.c_long, C.c_int, C.c_int, C.c_int, C.c_double, _u8, _u8]
        self.lib.orc_collide.restype = C.c_long
        self.lib.orc_collide(p0, p1, len(p0), *dims, dx, np.ascontiguousarray(material, np.uint8).reshape(-1), flags)
        return p1, flags

    def extrapolate(self, u, v, w, dims, material, nlayers):
        """MACVelocityField::extrapolateVelocityField on copies of u, v, w."""
        u, v, w = [np.array(a, np.float32, copy=True).reshape(-1) for a in (u, v, w)]
        self.lib.orc_extrapolate.argtypes = [_f32, _f32, _f32, C.c_int, C.c_int, C.c_int, _u8, C.c_int]
        self.lib.orc_extrapolate(u, v, w, *dims, np.ascontiguousarray(material, np.uint8).reshape(-1), int(nlayers))
        return u, v, w

    # -- stages 6-8 (oracle_pressure.c)
    def body_force(self, u, v, w, dims, material, force, dt):
        """FluidSimulation::_applyConstantBodyForces on copies of u, v, w."""
        u, v, w = [np.array(a, np.float32, copy=True).reshape(-1) for a in (u, v, w)]
        f = np.asarray(force, np.float32)
        self.lib.orc_body_force.argtypes = [_f32, _f32, _f32, C.c_int, C.c_int, C.c_int, _u8, _f32, C.c_double]
        self.lib.orc_body_force(u, v, w, *dims, np.ascontiguousarray(material, np.uint8).reshape(-1), f, float(dt))
        return u, v, w

    def pressure_solve(self, u, v, w, dims, dx, material, dt, density=20.0, tolerance=1e-6, max_iterations=200):
        """PressureSolver::solve + the narrowing of FluidSimulation::_updatePressureGrid: (float pressure per cell,
        CG iterations (-1: right-hand side below tolerance), hit the iteration limit, last residual max-norm)."""
        out = np.zeros(int(np.prod(dims)), np.float32)
        info = np.zeros(2, np.int32)
        self.lib.orc_pressure_solve.restype = C.c_double
        self.lib.orc_pressure_solve.argtypes = [_f32, _f32, _f32, C.c_int, C.c_int, C.c_int, C.c_double, _u8, C.c_double,
                                                C.c_double, C.c_double, C.c_int, _f32, _i32]
        err = self.lib.orc_pressure_solve(_c(u), _c(v), _c(w), *dims, dx, np.ascontiguousarray(material, np.uint8).reshape(-1),
                                          float(dt), float(density), float(tolerance), int(max_iterations), out, info)
        return out, int(info[0]), bool(info[1]), float(err)

    def apply_pressure(self, u, v, w, dims, dx, material, pressure, dt, density=20.0):
        """FluidSimulation::_applyPressureToVelocityField on copies of u, v, w."""
        u, v, w = [np.array(a, np.float32, copy=True).reshape(-1) for a in (u, v, w)]
        self.lib.orc_apply_pressure.argtypes = [_f32, _f32, _f32, C.c_int, C.c_int, C.c_int, C.c_double, _u8, _f32, C.c_double, C.c_double]
        self.lib.orc_apply_pressure(u, v, w, *dims, dx, np.ascontiguousarray(material, np.uint8).reshape(-1),
                                    np.ascontiguousarray(pressure, np.float32).reshape(-1), float(dt), float(density))
        return u, v, w

    def sample(self, pos, u, v, w, dims, dx, mode, validate=True):
        pos = _c(pos)
        out = np.empty_like(pos)
        self.lib.orc_sample(pos, len(pos), _c(u), _c(v), _c(w), *dims, dx, mode, int(validate), out)
        return out

    def advect(self, pos, u, v, w, dims, dx, dt, order=4, mode=1):
        pos = _c(pos)
        out = np.empty_like(pos)
        self.lib.orc_advect(pos, len(pos), _c(u), _c(v), _c(w), *dims, dx, dt, order, mode, out)
        return out

    def picflip(self, pos, vel, new, saved, dims, dx, ratio=float(np.float32(0.05)), mode=1):
        pos, vel = _c(pos), _c(vel)
        out = np.empty_like(vel)
        self.lib.orc_picflip(pos, vel, len(pos), *[_c(a) for a in new], *[_c(a) for a in saved],
                             *dims, dx, ratio, mode, out)
        return out

    def splat(self, pos, values, radius, offset, dx, ndims, field=None, weight=None):
        pos, values = _c(pos), _c(values)
        cnt = ndims[0] * ndims[1] * ndims[2]
        field = np.zeros(cnt, np.float32) if field is None else field
        weight = np.zeros(cnt, np.float32) if weight is None else weight
        self.lib.orc_splat(pos, values, 1, len(pos), radius, _c(offset), dx, *ndims, field, weight)
        return field, weight

    def apply_weight(self, field, weight):
        self.lib.orc_apply_weight(field, weight, field.size)
        return field

    def border_material(self, dims):
        m = np.zeros(dims[0] * dims[1] * dims[2], np.uint8)
        self.lib.orc_border_solid(*dims, m)
        return m

    def classify(self, pos, dims, dx, material):
        pos = _c(pos)
        bad = self.lib.orc_classify(pos, len(pos), *dims, dx, material)
        return material, bad

    def p2g(self, pos, vel, dims, dx, material, sources=()):
        pos, vel = _c(pos), _c(vel)
        src, ns = make_sources(list(sources))
        nu, nv, nw = face_counts(*dims)
        u, v, w = np.empty(nu, np.float32), np.empty(nv, np.float32), np.empty(nw, np.float32)
        self.lib.orc_p2g(pos, vel, len(pos), *dims, dx, material, C.cast(src, C.c_void_p), ns, u, v, w)
        return u, v, w

    def splat_component(self, pos, vel, comp, dims, dx, field, weight):
        """accumulate velocity component `comp` of the particles into raw (un-normalised) node sums"""
        nd = face_dims(*dims)[comp]
        off = np.array([0.0 if comp == 0 else 0.5 * dx, 0.0 if comp == 1 else 0.5 * dx, 0.0 if comp == 2 else 0.5 * dx], np.float32)
        pos = _c(pos)
        vals = np.ascontiguousarray(_c(vel)[:, comp])
        self.lib.orc_splat(pos, vals, 1, len(pos), dx, off, dx, *nd, field, weight)

    def finish_component(self, field, weight, comp, dims, dx, material, sources=()):
        src, ns = make_sources(list(sources))
        out = np.empty_like(field)
        self.lib.orc_finish_component(field, weight, comp, *dims, dx, material, C.cast(src, C.c_void_p), ns, out)
        return out

    def g2p_advect(self, pos, vel, new, saved, dims, dx, dt, ratio=float(np.float32(0.05)), order=4, mode=1,
                   material=None, resolve=True):
        """resolve=True (the reference's behaviour): particles advected into a solid cell go through
        _resolveParticleSolidCellCollision (fluidsimulation.cpp:3145-3179); False: they just keep p0."""
        pos, vel = _c(pos), _c(vel)
        pos_out, vel_out = np.empty_like(pos), np.empty_like(vel)
        flags = np.zeros(len(pos), np.uint8)
        mptr = material.ctypes.data_as(C.c_void_p) if material is not None else None
        fn = self.lib.orc_g2p_advect_resolve if resolve else self.lib.orc_g2p_advect
        fn.argtypes = self.lib.orc_g2p_advect.argtypes
        fn(pos, vel, len(pos), *[_c(a) for a in new], *[_c(a) for a in saved],
                                *dims, dx, ratio, dt, order, mode, mptr, pos_out, vel_out,
                                flags.ctypes.data_as(C.c_void_p))
        return pos_out, vel_out, flags


class Reference:
    """The unmodified reference, driven through oracle/ref_harness.cpp."""

    def __init__(self, build=True, path=None):
        if path is None:
            path = build_reference() if build else (REF_SO if os.path.exists(REF_SO) else None)
        if path is None or not os.path.exists(path):
            raise FileNotFoundError("oracle/_ref/libgfsref.so is not built and /root/reference is absent")
        self.lib = L = C.CDLL(path)
        V = C.c_void_p
        L.ref_max_threads.restype = C.c_int
        L.ref_cell_index.argtypes = [_f32, C.c_long, C.c_double, _i32]
        L.ref_sample.argtypes = [_f32, C.c_long, _f32, _f32, _f32, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, _f32]
        L.ref_advect.argtypes = [_f32, C.c_long, _f32, _f32, _f32, C.c_int, C.c_int, C.c_int, C.c_double,
                                 C.c_double, C.c_int, _f32]
        L.ref_add_point_values.argtypes = [_f32, _f32, C.c_long, C.c_double, _f32, C.c_double,
                                           C.c_int, C.c_int, C.c_int, _f32, _f32]
        L.ref_sim_create.restype = V
        L.ref_sim_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double]
        for name in ("ref_sim_initialize", "ref_sim_destroy", "ref_sim_update_fluid_cells",
                     "ref_sim_advect_velocity_field", "ref_sim_update_particle_velocities"):
            getattr(L, name).argtypes = [V]
        L.ref_sim_set_accel.argtypes = [V, C.c_int, C.c_int]
        L.ref_sim_add_fluid_sphere.argtypes = [V, C.c_float, C.c_float, C.c_float, C.c_double]
        L.ref_sim_add_fluid_cuboid.argtypes = [V, C.c_float, C.c_float, C.c_float, C.c_double, C.c_double, C.c_double]
        L.ref_sim_add_body_force.argtypes = [V, C.c_float, C.c_float, C.c_float]
        L.ref_sim_add_solid_cells.argtypes = [V, _i32, C.c_long]
        L.ref_sim_add_inflow_source.argtypes = [V, C.c_int, C.c_float, C.c_float, C.c_float, C.c_double, C.c_double,
                                                C.c_double, C.c_float, C.c_float, C.c_float]
        L.ref_sim_num_particles.argtypes = [V]
        L.ref_sim_num_particles.restype = C.c_long
        L.ref_sim_set_particles.argtypes = [V, _f32, _f32, C.c_long]
        L.ref_sim_get_particles.argtypes = [V, _f32, _f32]
        L.ref_sim_get_material.argtypes = [V, _u8]
        L.ref_sim_set_fields.argtypes = [V, _f32, _f32, _f32, _f32, _f32, _f32]
        L.ref_sim_get_fields.argtypes = [V, _f32, _f32, _f32]
        L.ref_sim_advance_particles.argtypes = [V, C.c_double]
        L.ref_sim_update.argtypes = [V, C.c_double]
        L.ref_hotpath_create.restype = V
        L.ref_hotpath_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
        L.ref_hotpath_destroy.argtypes = [V]
        L.ref_hotpath_set_fields.argtypes = [V, _f32, _f32, _f32, _f32, _f32, _f32]
        L.ref_hotpath_step.argtypes = [V, _f32, _f32, C.c_long, C.c_double, C.c_int, _f64]
        L.ref_hotpath_step.restype = C.c_double

    # -- primitives
    def cell_index(self, pos, dx):
        pos = _c(pos)
        out = np.empty((len(pos), 3), np.int32)
        self.lib.ref_cell_index(pos, len(pos), dx, out)
        return out

    def sample(self, pos, u, v, w, dims, dx, mode):
        pos = _c(pos)
        out = np.empty_like(pos)
        self.lib.ref_sample(pos, len(pos), _c(u), _c(v), _c(w), *dims, dx, mode, out)
        return out

    def advect(self, pos, u, v, w, dims, dx, dt, order=4):
        pos = _c(pos)
        out = np.empty_like(pos)
        self.lib.ref_advect(pos, len(pos), _c(u), _c(v), _c(w), *dims, dx, dt, order, out)
        return out

    def extrapolate(self, u, v, w, dims, dx, material, nlayers):
        u, v, w = [np.array(a, np.float32, copy=True).reshape(-1) for a in (u, v, w)]
        self.lib.ref_extrapolate.argtypes = [_f32, _f32, _f32, C.c_int, C.c_int, C.c_int, C.c_double, _u8, C.c_int]
        self.lib.ref_extrapolate(u, v, w, *dims, dx, np.ascontiguousarray(material, np.uint8).reshape(-1), int(nlayers))
        return u, v, w

    def add_point_values(self, pos, values, radius, offset, dx, ndims):
        pos, values = _c(pos), _c(values)
        cnt = ndims[0] * ndims[1] * ndims[2]
        field, weight = np.zeros(cnt, np.float32), np.zeros(cnt, np.float32)
        self.lib.ref_add_point_values(pos, values, len(pos), radius, _c(offset), dx, *ndims, field, weight)
        return field, weight

    # -- save states (src/fluidsimulationsavestate.cpp)
    def state_read(self, path):
        """The reference's own reader on `path`: dict(dims, dx, frame, n_diffuse, pos, vel, solid_ijk) or None."""
        L = self.lib
        L.ref_state_read.argtypes = [C.c_char_p, _i32, C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_state_read.restype = C.c_int
        hdr, dx = np.zeros(7, np.int32), C.c_double()
        if not L.ref_state_read(path.encode(), hdr, C.byref(dx), None, None, None):
            return None
        pos, vel = np.empty((hdr[4], 3), np.float32), np.empty((hdr[4], 3), np.float32)
        solid = np.empty((hdr[6], 3), np.int32)
        L.ref_state_read(path.encode(), hdr, C.byref(dx), pos.ctypes.data_as(C.c_void_p), vel.ctypes.data_as(C.c_void_p),
                         solid.ctypes.data_as(C.c_void_p))
        return dict(dims=tuple(int(x) for x in hdr[:3]), dx=dx.value, frame=int(hdr[3]), n_diffuse=int(hdr[5]),
                    pos=pos, vel=vel, solid_ijk=solid)

    def sim_from_state(self, path):
        """FluidSimulation(FluidSimulationSaveState&) on `path`, wrapped as a RefSim (already initialised)."""
        self.lib.ref_sim_create_from_state.restype = C.c_void_p
        self.lib.ref_sim_create_from_state.argtypes = [C.c_char_p]
        h = self.lib.ref_sim_create_from_state(path.encode())
        if not h:
            return None
        st = self.state_read(path)
        sim = RefSim.__new__(RefSim)
        sim.lib, sim.dims, sim.dx, sim.h, sim._init = self.lib, st["dims"], st["dx"], h, True
        return sim

    # -- simulation-level
    def sim(self, dims, dx):
        return RefSim(self.lib, dims, dx)

    def hotpath(self, nthreads, dims, dx, new, saved):
        return RefHotpath(self.lib, nthreads, dims, dx, new, saved)


class RefHotpath:
    """nthreads private reference simulators timing the three hot-path stages (see ref_harness.cpp)."""

    def __init__(self, lib, nthreads, dims, dx, new, saved):
        self.lib, self.nthreads = lib, nthreads
        self.h = lib.ref_hotpath_create(nthreads, *dims, dx)
        lib.ref_hotpath_set_fields(self.h, *[_c(a) for a in new], *[_c(a) for a in saved])

    def step(self, pos, vel, dt, mode):
        """-> (seconds, [P2G, PIC/FLIP, RK4] seconds)"""
        stages = np.zeros(3)
        pos, vel = _c(pos), _c(vel)
        t = self.lib.ref_hotpath_step(self.h, pos, vel, len(pos), dt, mode, stages)
        return t, stages

    def close(self):
        if self.h:
            self.lib.ref_hotpath_destroy(self.h)
            self.h = None


class RefSim:
    def __init__(self, lib, dims, dx):
        self.lib, self.dims, self.dx = lib, tuple(dims), dx
        self.h = lib.ref_sim_create(*dims, dx)
        self._init = False

    def add_fluid_sphere(self, c, r):
        self.lib.ref_sim_add_fluid_sphere(self.h, *c, r)

    def add_fluid_cuboid(self, p, w, h, d):
        self.lib.ref_sim_add_fluid_cuboid(self.h, *p, w, h, d)

    def add_body_force(self, f):
        self.lib.ref_sim_add_body_force(self.h, *f)

    def add_swirl_force(self):
        """a variable body force field (host callback) about the vertical axis through (4, *, 4)"""
        self.lib.ref_sim_add_swirl_force.argtypes = [C.c_void_p]
        self.lib.ref_sim_add_swirl_force(self.h)

    def add_solid_cells(self, ijk):
        ijk = _c(ijk, np.int32)
        self.lib.ref_sim_add_solid_cells(self.h, ijk, len(ijk))

    def add_inflow_source(self, kind, p, a, b, c, velocity):
        self.lib.ref_sim_add_inflow_source(self.h, kind, *p, a, b, c, *velocity)

    def add_outflow_source(self, kind, p, a, b=0.0, c=0.0):
        fn = self.lib.ref_sim_add_outflow_source
        fn.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_double, C.c_double, C.c_double]
        fn(self.h, kind, *p, a, b, c)

    def update_fluid_sources(self):
        """FluidSimulation::_updateFluidSources on the current state (inflow emission, outflow removal)."""
        self.lib.ref_sim_update_fluid_sources.argtypes = [C.c_void_p]
        self.lib.ref_sim_update_fluid_sources(self.h)

    def initialize(self):
        self.lib.ref_sim_initialize(self.h)
        self._init = True

    def enable_mesh_output(self):
        """Stock output settings of the reference (surface mesh + isotropic reconstruction on); -> output directory."""
        self.lib.ref_sim_enable_mesh_output.restype = C.c_char_p
        self.lib.ref_sim_enable_mesh_output.argtypes = [C.c_void_p]
        return self.lib.ref_sim_enable_mesh_output(self.h).decode()

    def log_path(self):
        """Path of the simulator's own log (its directory is created); call before update()."""
        self.lib.ref_sim_log_path.restype = C.c_char_p
        self.lib.ref_sim_log_path.argtypes = [C.c_void_p]
        return self.lib.ref_sim_log_path(self.h).decode()

    @staticmethod
    def stage_times(path):
        """{stage name: seconds summed over the logged substeps} from the reference's own StopWatch lines."""
        out, steps = {}, 0
        names = ("Update Fluid Cells", "Reconstruct Fluid Surface", "Update Level set", "Reconstruct Output Surface",
                 "Advect Velocity Field", "Apply Body Forces", "Update Pressure Grid", "Apply Pressure",
                 "Extrapolate Fluid Velocities", "Update Diffuse Material", "Update PIC/FLIP Velocities",
                 "Advance Marker Particles", "Update time")
        try:
            with open(path) as f:
                lines = f.readlines()
        except OSError:
            return out, 0
        past_breakdown = False
        for ln in lines:
            if ln.startswith("---Percentage Breakdown---"):
                past_breakdown = True
            if ln.startswith("Frame:") or ln.startswith("Step time"):
                past_breakdown = False
            if past_breakdown and not ln.startswith("Update time"):
                continue
            for nm in names:
                if ln.startswith(nm + ":"):
                    try:
                        out[nm] = out.get(nm, 0.0) + float(ln.split(":", 1)[1].strip().split()[0])
                    except (ValueError, IndexError):
                        pass
                    if nm == "Update time":
                        steps += 1
        return out, steps

    def mesh_particles(self, use_accelerator, cap=4000000):
        """IsotropicParticleMesher::meshParticles on the current state -> (vertices (n,3), triangle count)."""
        fn = self.lib.ref_sim_mesh_particles
        fn.restype = C.c_long
        fn.argtypes = [C.c_void_p, C.c_int, np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS"), C.c_long, C.POINTER(C.c_long)]
        verts = np.empty((cap, 3), np.float32)
        ntris = C.c_long(0)
        nv = fn(self.h, int(use_accelerator), verts, cap, C.byref(ntris))
        return verts[:min(nv, cap)].copy(), int(ntris.value)

    def set_accel(self, particle_advection, scalar_field):
        self.lib.ref_sim_set_accel(self.h, int(particle_advection), int(scalar_field))

    @property
    def n(self):
        return self.lib.ref_sim_num_particles(self.h)

    def set_particles(self, pos, vel):
        pos, vel = _c(pos), _c(vel)
        self.lib.ref_sim_set_particles(self.h, pos, vel, len(pos))

    def get_particles(self):
        n = self.n
        pos, vel = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        self.lib.ref_sim_get_particles(self.h, pos, vel)
        return pos, vel

    def get_material(self):
        m = np.empty(self.dims[0] * self.dims[1] * self.dims[2], np.uint8)
        self.lib.ref_sim_get_material(self.h, m)
        return m

    def set_fields(self, new, saved):
        self.lib.ref_sim_set_fields(self.h, *[_c(a) for a in new], *[_c(a) for a in saved])

    def get_fields(self):
        nu, nv, nw = face_counts(*self.dims)
        u, v, w = np.empty(nu, np.float32), np.empty(nv, np.float32), np.empty(nw, np.float32)
        self.lib.ref_sim_get_fields(self.h, u, v, w)
        return u, v, w

    def update_fluid_cells(self):
        self.lib.ref_sim_update_fluid_cells(self.h)

    def advect_velocity_field(self):
        self.lib.ref_sim_advect_velocity_field(self.h)

    def update_particle_velocities(self):
        self.lib.ref_sim_update_particle_velocities(self.h)

    def advance_particles(self, dt):
        self.lib.ref_sim_advance_particles(self.h, dt)

    def apply_body_forces(self, dt):
        self.lib.ref_sim_apply_body_forces.argtypes = [C.c_void_p, C.c_double]
        self.lib.ref_sim_apply_body_forces(self.h, dt)

    def update_pressure_grid(self, dt):
        """FluidSimulation::_updatePressureGrid on the simulator's current _MACVelocity / material / fluid cell list."""
        out = np.zeros(int(np.prod(self.dims)), np.float32)
        self.lib.ref_sim_update_pressure_grid.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        self.lib.ref_sim_update_pressure_grid(self.h, dt, out.ctypes.data)
        return out

    def apply_pressure(self, dt, pressure):
        p = np.ascontiguousarray(pressure, np.float32)
        self.lib.ref_sim_apply_pressure.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        self.lib.ref_sim_apply_pressure(self.h, dt, p.ctypes.data)

    def extrapolate_velocities(self):
        self.lib.ref_sim_extrapolate_velocities.argtypes = [C.c_void_p]
        self.lib.ref_sim_extrapolate_velocities(self.h)

    def density(self):
        self.lib.ref_sim_density.restype = C.c_double
        self.lib.ref_sim_density.argtypes = [C.c_void_p]
        return self.lib.ref_sim_density(self.h)

    def save_state(self, path):
        self.lib.ref_sim_save_state.argtypes = [C.c_void_p, C.c_char_p]
        self.lib.ref_sim_save_state(self.h, path.encode())

    def update(self, dt):
        self.lib.ref_sim_update(self.h, dt)

    def close(self):
        if self.h:
            self.lib.ref_sim_destroy(self.h)
            self.h = None
