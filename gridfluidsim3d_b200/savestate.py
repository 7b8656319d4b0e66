"""GridFluidSim3D save-state files (`FluidSimulationSaveState`, src/fluidsimulationsavestate.cpp:31-144) as a wire
format for particle sets: what a user of the reference already has on disk can be loaded straight into the device-resident
domain (gfs_domain_init + gfs_set_material + gfs_set_particles), and a resident particle set can be written back in a form
the reference's `FluidSimulation(FluidSimulationSaveState&)` constructor accepts.  SURVEY 8(f) rank 4, second half.

Layout (native little-endian, no padding, as the reference writes it field by field):
    int32 isize, jsize, ksize; float64 dx; int32 next_frame; int32 n_marker; int32 n_diffuse; int32 n_solid; uint8 bricks
    float32 marker positions [n_marker*3];  float32 marker velocities [n_marker*3]
    float32 diffuse positions [n_diffuse*3]; float32 diffuse velocities [n_diffuse*3]
    float32 diffuse lifetimes [n_diffuse];   int8 diffuse types [n_diffuse]
    int32 solid cell indices [n_solid*3]  (i, j, k scan order of the writer: k outermost, i innermost)
    [FluidBrickGrid state when `bricks` is set: carried as opaque bytes]
"""
import struct

import numpy as np

HEADER = struct.Struct("<iiidiiii?")          # 37 bytes


def read_state(path):
    """Returns a dict: dims, dx, frame, pos, vel (float32 [n,3]), diffuse_pos, diffuse_vel, diffuse_lifetime,
    diffuse_type, solid_ijk (int32 [m,3]), brick_blob (bytes or None)."""
    with open(path, "rb") as f:
        raw = f.read()
    if len(raw) < HEADER.size:
        raise ValueError("%s: shorter than a save-state header" % path)
    I, J, K, dx, frame, nm, nd, ns, bricks = HEADER.unpack_from(raw, 0)
    if min(I, J, K) <= 0 or not dx > 0 or min(nm, nd, ns) < 0:
        raise ValueError("%s: not a GridFluidSim3D save state (header %r)" % (path, (I, J, K, dx, nm, nd, ns)))
    off = HEADER.size

    def take(dtype, count):
        nonlocal off
        nbytes = np.dtype(dtype).itemsize * count
        if off + nbytes > len(raw):
            raise ValueError("%s: truncated save state" % path)
        a = np.frombuffer(raw, dtype, count, off).copy()
        off += nbytes
        return a
    pos = take("<f4", 3 * nm).reshape(nm, 3)
    vel = take("<f4", 3 * nm).reshape(nm, 3)
    dpos = take("<f4", 3 * nd).reshape(nd, 3)
    dvel = take("<f4", 3 * nd).reshape(nd, 3)
    dlife = take("<f4", nd)
    dtype_ = take("i1", nd)
    solid = take("<i4", 3 * ns).reshape(ns, 3)
    return dict(dims=(I, J, K), dx=dx, frame=frame, pos=pos, vel=vel, diffuse_pos=dpos, diffuse_vel=dvel,
                diffuse_lifetime=dlife, diffuse_type=dtype_, solid_ijk=solid, brick_blob=raw[off:] if bricks else None)


def write_state(path, dims, dx, pos, vel, solid_ijk, frame=0, diffuse_pos=None, diffuse_vel=None,
                diffuse_lifetime=None, diffuse_type=None):
    """Writes a state without FluidBrickGrid data (the reference then starts with brick output disabled)."""
    pos = np.ascontiguousarray(pos, "<f4").reshape(-1, 3)
    vel = np.ascontiguousarray(vel, "<f4").reshape(-1, 3)
    if len(pos) != len(vel):
        raise ValueError("positions and velocities differ in length")
    nd = 0 if diffuse_pos is None else len(diffuse_pos)
    dpos = np.zeros((0, 3), "<f4") if diffuse_pos is None else np.ascontiguousarray(diffuse_pos, "<f4").reshape(nd, 3)
    dvel = np.zeros((nd, 3), "<f4") if diffuse_vel is None else np.ascontiguousarray(diffuse_vel, "<f4").reshape(nd, 3)
    dlife = np.zeros(nd, "<f4") if diffuse_lifetime is None else np.ascontiguousarray(diffuse_lifetime, "<f4").reshape(nd)
    dtyp = np.zeros(nd, "i1") if diffuse_type is None else np.ascontiguousarray(diffuse_type, "i1").reshape(nd)
    solid = np.ascontiguousarray(solid_ijk, "<i4").reshape(-1, 3)
    with open(path, "wb") as f:
        f.write(HEADER.pack(int(dims[0]), int(dims[1]), int(dims[2]), float(dx), int(frame), len(pos), nd, len(solid), False))
        for a in (pos, vel, dpos, dvel, dlife, dtyp, solid):
            f.write(a.tobytes())


def material_from_state(state):
    """uint8 material grid (0 air, 2 solid; fluid is classified from the particles) for gfs_set_material."""
    I, J, K = state["dims"]
    m = np.zeros((K, J, I), np.uint8)
    s = state["solid_ijk"]
    m[s[:, 2], s[:, 1], s[:, 0]] = 2
    return m.reshape(-1)


def solid_ijk_from_material(material, dims):
    """Solid cell indices in the reference writer's scan order (k outermost, i innermost)."""
    I, J, K = dims
    kk, jj, ii = np.nonzero(np.asarray(material, np.uint8).reshape(K, J, I) == 2)
    return np.stack([ii, jj, kk], 1).astype(np.int32)
