"""Seeded synthetic inputs for the particle<->grid transfer path.

Everything is plain numpy and deterministic in (config, seed); the same arrays feed the CUDA path,
the oracle and the reference, so results are comparable element by element.

Scene conventions follow the reference:
  * border cells are solid                         (src/fluidsimulation.cpp:1191-1213)
  * 8 particles per fluid cell on the 2x2x2 sub-cell centres, each jittered uniformly by
    +-0.25*0.1*dx                                  (src/fluidsimulation.cpp:1219-1247, fluidsimulation.h:1111)
  * faces sit at Grid3d::FaceIndexToPosition{U,V,W} (src/grid3d.h:113-135)
"""
import numpy as np

AIR, FLUID, SOLID = 0, 1, 2

#: BASELINE.json configs -> (dims, dx, fluid-shape name)
CONFIGS = {
    "hello64": ((64, 64, 64), 0.125, "sphere"),          # configs[0]: 64^3 sphere drop
    "dambreak128": ((128, 128, 128), 0.0625, "dam"),     # configs[1]: 128^3 cuboid dam break (~7.7M)
    "splash256": ((256, 256, 256), 0.03125, "splash"),   # configs[2]: 256^3 splash (~100M)
    "river512": ((512, 256, 256), 0.03125, "river"),     # configs[3]: 512x256x256 channel (~250M)
    "tiny16": ((16, 16, 16), 0.5, "sphere"),             # test-only
    "small32": ((32, 32, 32), 0.25, "sphere"),           # test-only
    "slab24": ((24, 20, 28), 0.25, "dam"),               # test-only, non-cubic
    "odd20": ((20, 18, 22), 0.3, "sphere"),              # test-only, dx not a power of two (fp64 index path)
}


def face_dims(dims):
    I, J, K = dims
    return (I + 1, J, K), (I, J + 1, K), (I, J, K + 1)


def border_material(dims):
    """uint8 material grid (flat, i fastest) with a one-cell solid border."""
    I, J, K = dims
    m = np.full((K, J, I), AIR, np.uint8)
    m[0], m[-1] = SOLID, SOLID
    m[:, 0], m[:, -1] = SOLID, SOLID
    m[:, :, 0], m[:, :, -1] = SOLID, SOLID
    return m.reshape(-1)


def fluid_cells(shape, dims, material=None):
    """Boolean (K,J,I) mask of cells seeded with particles; never a solid cell."""
    I, J, K = dims
    k, j, i = np.meshgrid(np.arange(K), np.arange(J), np.arange(I), indexing="ij", sparse=True)
    cx, cy, cz = (i + 0.5) / I, (j + 0.5) / J, (k + 0.5) / K
    if shape == "sphere":            # README Hello World: centre of the domain, effective radius 3/8 of the width
        mask = (cx - 0.5) ** 2 + (cy - 0.5) ** 2 + (cz - 0.5) ** 2 < 0.375 ** 2
    elif shape == "dam":             # cuboid x < 0.5 W, y < 0.95 H, all interior z
        mask = (cx < 0.5) & (cy < 0.95) & (cz > -1)
    elif shape == "splash":          # deep pool y < 0.72 H plus two inflow balls of radius 0.12 W above it
        ball1 = (cx - 0.3) ** 2 + (cy - 0.86) ** 2 + (cz - 0.3) ** 2 < 0.12 ** 2
        ball2 = (cx - 0.7) ** 2 + (cy - 0.85) ** 2 + (cz - 0.7) ** 2 < 0.12 ** 2
        mask = (cy < 0.72) | ball1 | ball2
        mask = mask & (cx > -1) & (cz > -1)
    elif shape == "river":           # channel y < 0.47 H
        mask = (cy < 0.47) & (cx > -1) & (cz > -1)
    elif shape == "full":
        mask = (cx > -1) & (cy > -1) & (cz > -1)
    else:
        raise ValueError(shape)
    mask = np.broadcast_to(mask, (K, J, I)).copy()
    solid = (border_material(dims) if material is None else material).reshape(K, J, I) == SOLID
    mask &= ~solid
    return mask


def make_particles(mask, dx, seed=12345, shuffle=True, jitter_factor=0.1):
    """(N,3) float32 positions: 8 jittered particles per True cell of mask[(K,J,I)]."""
    rng = np.random.Generator(np.random.Philox(seed))
    kk, jj, ii = np.nonzero(mask)
    ncell = len(ii)
    centre = np.stack([ii, jj, kk], 1).astype(np.float64) * dx + 0.5 * dx          # (ncell,3)
    q = 0.25 * dx
    sub = np.array([[-1, -1, -1], [1, -1, -1], [1, -1, 1], [-1, -1, 1],
                    [-1, 1, -1], [1, 1, -1], [1, 1, 1], [-1, 1, 1]], np.float64) * q
    pos = (centre[:, None, :] + sub[None, :, :]).reshape(ncell * 8, 3)
    jit = 0.25 * jitter_factor * dx
    pos += rng.uniform(-jit, jit, size=pos.shape)
    pos = pos.astype(np.float32)
    if shuffle:                                   # the reference reshuffles every step (fluidsimulation.cpp:3211-3219)
        pos = pos[rng.permutation(len(pos))]
    return np.ascontiguousarray(pos)


def vortex(x, y, z, extent):
    """Analytic smooth vortex, |v| <= 1: returns (u, v, w) at the given coordinates (float64)."""
    W, H, D = extent
    xh, yh, zh = x / W, y / H, z / D
    pi = np.pi
    u = -np.sin(pi * xh) ** 2 * np.sin(2 * pi * yh) * np.cos(pi * zh)
    v = np.sin(2 * pi * xh) * np.sin(pi * yh) ** 2 * np.cos(pi * zh)
    w = 0.5 * np.sin(2 * pi * xh) * np.sin(2 * pi * zh) * np.sin(pi * yh)
    return u, v, w


def particle_velocities(pos, dims, dx):
    ext = (dims[0] * dx, dims[1] * dx, dims[2] * dx)
    p = pos.astype(np.float64)
    u, v, w = vortex(p[:, 0], p[:, 1], p[:, 2], ext)
    return np.ascontiguousarray(np.stack([u, v, w], 1).astype(np.float32))


def make_fields(dims, dx, saved_scale=0.9):
    """(u,v,w) 'new' and 'saved' MAC fields (flat float32, i fastest) sampled at face centres."""
    I, J, K = dims
    ext = (I * dx, J * dx, K * dx)
    out = []
    for comp, (ni, nj, nk) in enumerate(face_dims(dims)):
        k, j, i = np.meshgrid(np.arange(nk, dtype=np.float64), np.arange(nj, dtype=np.float64),
                              np.arange(ni, dtype=np.float64), indexing="ij", sparse=True)
        x = (i + (0.0 if comp == 0 else 0.5)) * dx
        y = (j + (0.0 if comp == 1 else 0.5)) * dx
        z = (k + (0.0 if comp == 2 else 0.5)) * dx
        out.append(np.ascontiguousarray(
            np.broadcast_to(vortex(x, y, z, ext)[comp], (nk, nj, ni)).astype(np.float32).reshape(-1)))
    new = tuple(out)
    saved = tuple((a * np.float32(saved_scale)).astype(np.float32) for a in new)
    return new, saved


def cfl_dt(dx, cfl=0.5, vmax=1.0):
    """dt such that the fastest particle moves cfl*dx in one substep (the vortex has |v| <= ~1.1)."""
    return cfl * dx / vmax


def make_scene(name, seed=12345, shuffle=True, max_particles=None):
    """Everything one substep needs, as a dict of numpy arrays."""
    dims, dx, shape = CONFIGS[name]
    material = border_material(dims)
    mask = fluid_cells(shape, dims, material)
    pos = make_particles(mask, dx, seed, shuffle)
    if max_particles is not None and len(pos) > max_particles:
        pos = np.ascontiguousarray(pos[:max_particles])
    vel = particle_velocities(pos, dims, dx)
    new, saved = make_fields(dims, dx)
    return dict(name=name, dims=dims, dx=dx, material=material, pos=pos, vel=vel, new=new, saved=saved,
                dt=cfl_dt(dx))
