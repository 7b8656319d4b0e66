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
    "tall64": ((32, 32, 64), 0.25, "sphere"),            # test-only: enough layers for 8 slabs
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
    elif shape == "river":           # channel y < 0.93 H: ~245 M particles at 512x256x256 (BASELINE configs[3]: "~250M")
        mask = (cy < 0.93) & (cx > -1) & (cz > -1)
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


# ---------------------------------------------------------------------------------------------------------
# Counter-based scene generation with torch (CPU or CUDA): every particle's jitter is a pure function of
# (seed, global cell index, sub-cell slot, axis), so a slab-sharded run generates -- rank by rank, from its own
# layers only -- exactly the particle set the single-GPU run generates.  Used by bench.py and by the parity tests at
# the BASELINE sizes (numpy takes minutes at 10^8 particles).
# ---------------------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def _s64(x):
    """Python int -> the signed 64-bit value with the same bits (torch has no uint64 arithmetic)."""
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(x, s):
    """logical right shift of an int64 tensor"""
    return (x >> s) & ((1 << (64 - s)) - 1)


def splitmix64_t(x):
    """splitmix64 finaliser on an int64 tensor (wrap-around arithmetic); same bits as splitmix64_np."""
    z = x + _s64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _s64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _s64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def splitmix64_np(x):
    with np.errstate(over="ignore"):
        z = x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


SUB_CELL = np.array([[-1, -1, -1], [1, -1, -1], [1, -1, 1], [-1, -1, 1],
                     [-1, 1, -1], [1, 1, -1], [1, 1, 1], [-1, 1, 1]], np.float64)


def hashed_positions_np(cell_ijk, dims, dx, seed=12345, jitter_factor=0.1):
    """float32 positions of the 8 particles of each listed cell (numpy twin of the torch generator, for tests)."""
    I, J, K = dims
    c = np.asarray(cell_ijk, np.int64)
    lin = c[:, 0] + I * (c[:, 1] + J * c[:, 2])
    pid = (lin[:, None] * 8 + np.arange(8)[None, :]).reshape(-1)
    centre = (np.repeat(c, 8, 0).astype(np.float64) + 0.5) * dx + np.tile(SUB_CELL, (len(c), 1)) * (0.25 * dx)
    jit = 0.25 * jitter_factor * dx
    out = np.empty((len(pid), 3), np.float64)
    for a in range(3):
        h = splitmix64_np((pid * 3 + a).astype(np.uint64) ^ splitmix64_np(np.array([seed], np.uint64)))
        u = (h >> np.uint64(40)).astype(np.float64) * (1.0 / (1 << 24))
        out[:, a] = centre[:, a] + (2.0 * u - 1.0) * jit
    return out.astype(np.float32)


def vortex_t(x, y, z, ext):
    import torch
    pi = float(np.pi)
    xh, yh, zh = x / ext[0], y / ext[1], z / ext[2]
    u = -torch.sin(pi * xh) ** 2 * torch.sin(2 * pi * yh) * torch.cos(pi * zh)
    v = torch.sin(2 * pi * xh) * torch.sin(pi * yh) ** 2 * torch.cos(pi * zh)
    w = 0.5 * torch.sin(2 * pi * xh) * torch.sin(2 * pi * zh) * torch.sin(pi * yh)
    return u, v, w


def make_scene_torch(name, device, seed=12345, k_range=None, shuffle=True, fields=True):
    """Particles (AoS float32 [N,6]), new/saved fields and material of workload `name` as torch tensors on `device`.
    k_range=(k0,k1) keeps only the particles seeded in cell layers [k0,k1): the union over a partition of the layers is
    the particle SET of the unrestricted call, bit for bit (only the order differs: the reference shuffles its particles
    every step, src/fluidsimulation.cpp:3211-3219, so the arrays are handed over shuffled)."""
    import torch
    dims, dx, shape = CONFIGS[name]
    I, J, K = dims
    material = border_material(dims)
    mask = torch.from_numpy(fluid_cells(shape, dims, material)).to(device)
    if k_range is not None:
        keep = torch.zeros(K, dtype=torch.bool, device=device)
        keep[k_range[0]:k_range[1]] = True
        mask &= keep[:, None, None]
    kk, jj, ii = torch.nonzero(mask, as_tuple=True)
    del mask
    lin = ii + I * (jj + J * kk)                                              # int64 global cell index
    ncell = lin.shape[0]
    sub = torch.tensor(SUB_CELL, dtype=torch.float64, device=device) * (0.25 * dx)
    key = splitmix64_t(torch.tensor([_s64(seed)], dtype=torch.int64, device=device))
    pid = (lin[:, None] * 8 + torch.arange(8, device=device, dtype=torch.int64)[None, :]).reshape(-1)
    jit = 0.25 * 0.1 * dx
    pos = torch.empty((ncell * 8, 3), dtype=torch.float32, device=device)
    for a, idx in enumerate((ii, jj, kk)):
        centre = ((idx.to(torch.float64) + 0.5) * dx)[:, None] + sub[None, :, a]          # (ncell, 8)
        h = splitmix64_t((pid * 3 + a) ^ key)
        u = _lsr(h, 40).to(torch.float64) * (1.0 / (1 << 24))
        pos[:, a] = (centre.reshape(-1) + (2.0 * u - 1.0) * jit).to(torch.float32)
        del centre, h, u
    del pid, lin, ii, jj, kk
    if shuffle:
        gen = torch.Generator(device=device)
        gen.manual_seed(seed)
        perm = torch.randperm(pos.shape[0], generator=gen, device=device)
        pos = pos[perm].contiguous()
        del perm
    ext = (I * dx, J * dx, K * dx)
    u, v, w = vortex_t(pos[:, 0], pos[:, 1], pos[:, 2], ext)
    aos = torch.cat([pos, torch.stack([u, v, w], 1)], 1).contiguous()
    del pos, u, v, w
    out = dict(name=name, dims=dims, dx=dx, dt=cfl_dt(dx), material=material, aos=aos)
    if fields:
        new = []
        for comp, (ni, nj, nk) in enumerate(face_dims(dims)):
            i = torch.arange(ni, device=device, dtype=torch.float32)[None, None, :]
            j = torch.arange(nj, device=device, dtype=torch.float32)[None, :, None]
            k = torch.arange(nk, device=device, dtype=torch.float32)[:, None, None]
            x = (i + (0.0 if comp == 0 else 0.5)) * dx
            y = (j + (0.0 if comp == 1 else 0.5)) * dx
            z = (k + (0.0 if comp == 2 else 0.5)) * dx
            new.append(vortex_t(x, y, z, ext)[comp].expand(nk, nj, ni).contiguous().reshape(-1))
        out["new"] = new
        out["saved"] = [a * 0.9 for a in new]
    return out
