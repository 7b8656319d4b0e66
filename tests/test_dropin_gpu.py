"""The drop-in boundary, end to end: the UNMODIFIED reference simulator compiled against
gridfluidsim3d_b200/dropin/{particleadvector,clscalarfield}.{h,cpp} (oracle/Makefile target `dropin`) and linked to
libgfs_b200.so, against the same simulator with its own CPU accelerator paths.  GPU only; skipped where the two
prebuilt libraries (oracle/_ref/) are absent."""
import ctypes
import os

import numpy as np
import pytest

from gridfluidsim3d_b200 import synth

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def libs():
    from oracle import pyoracle
    if not (os.path.exists(pyoracle.REF_SO) and os.path.exists(pyoracle.DROPIN_SO)):
        pytest.skip("oracle/_ref reference builds are not present")
    return pyoracle.Reference(build=False), pyoracle.Reference(build=False, path=pyoracle.DROPIN_SO)


@pytest.fixture(scope="module")
def resident_lib():
    from oracle import pyoracle
    if not (os.path.exists(pyoracle.REF_SO) and os.path.exists(pyoracle.RESIDENT_SO)):
        pytest.skip("oracle/_ref reference builds are not present")
    return pyoracle.Reference(build=False), pyoracle.Reference(build=False, path=pyoracle.RESIDENT_SO)


def test_accelerator_classes_exact_mode_is_bit_exact(libs, oracle):
    """ParticleAdvector with OpenCL 'disabled' = the CUDA library's exact arithmetic: bit-identical to the reference's
    CPU loops for tricubicInterpolate (validated) and advectParticlesRK1..4."""
    ref, drop = libs
    dims, dx = (12, 10, 14), 0.25
    rng = np.random.default_rng(5)
    u, v, w = (rng.standard_normal(a * b * c).astype(np.float32) for a, b, c in synth.face_dims(dims))
    pos = rng.uniform(-0.5 * dx, (np.array(dims) + 0.5) * dx, size=(20000, 3)).astype(np.float32)
    assert np.array_equal(bits(drop.sample(pos, u, v, w, dims, dx, 1)), bits(ref.sample(pos, u, v, w, dims, dx, 1)))
    for order in (1, 2, 3, 4):
        a = drop.advect(pos, u, v, w, dims, dx, 0.21, order)
        b = ref.advect(pos, u, v, w, dims, dx, 0.21, order)
        assert np.array_equal(bits(a), bits(b))
    # CLScalarField::addPointValues: fixed-point accumulation agrees with the CPU splat to fp32 rounding
    inside = pos[np.all((pos > 0) & (pos < np.array(dims) * dx), 1)]
    vals = rng.standard_normal(len(inside)).astype(np.float32)
    off = np.array([0.0, 0.5 * dx, 0.5 * dx], np.float32)
    fa, wa = drop.add_point_values(inside, vals, dx, off, dx, synth.face_dims(dims)[0])
    fb, wb = ref.add_point_values(inside, vals, dx, off, dx, synth.face_dims(dims)[0])
    assert np.abs(wa - wb).max() <= 1e-5 * wb.max() and np.abs(fa - fb).max() <= 1e-5 * np.abs(fb).max()
    assert np.array_equal(wa > 0, wb > 0)


@pytest.mark.parametrize("fast", [False, True])
def test_whole_simulator_with_cuda_accelerators(libs, fast):
    """Hello World at 32^3 (README.md:113-117 scaled): the reference FluidSimulation::update() driven for two frames,
    once with its own CPU accelerator paths and once with the CUDA drop-in classes underneath.  Both runs see the same
    rand() sequence, so particles stay index-aligned."""
    ref, drop = libs
    libc = ctypes.CDLL(None)
    out = []
    for lib in (ref, drop):
        libc.srand(1)
        sim = lib.sim((32, 32, 32), 0.25)
        sim.add_fluid_sphere((4.0, 4.0, 4.0), 5.0)
        sim.add_body_force((0.0, -25.0, 0.0))
        if lib is drop:
            sim.set_accel(fast, fast)
        sim.initialize()
        for _ in range(2):
            sim.update(1.0 / 30.0)
        p, v = sim.get_particles()
        out.append((p, v, sim.get_material(), sim.get_fields()))
        sim.close()
    (p0, v0, m0, f0), (p1, v1, m1, f1) = out
    assert len(p0) == len(p1) > 20000
    # two frames through P2G -> pressure solve -> extrapolation -> G2P: fp32-level differences in the splat are
    # carried through the solver, so compare at 1e-4 of the field scale (positions: of the cell size)
    assert np.abs(p0 - p1).max() < 1e-4 * 0.25
    assert np.abs(v0 - v1).max() < 1e-4 * max(1.0, np.abs(v0).max())
    assert (m0 != m1).mean() < 1e-3
    for a, b in zip(f0, f1):
        assert np.abs(a - b).max() < 1e-4 * max(1.0, np.abs(a).max())


def test_whole_simulator_stock_settings_with_surface_mesher(libs):
    """The reference's stock scene (src/main.cpp:8-21 at 32^3) with surface-mesh output and isotropic reconstruction ON:
    IsotropicParticleMesher sets the max-scalar-field threshold and calls CLScalarField::addPoints every frame
    (src/isotropicparticlemesher.cpp:334-359).  The drop-in must run it (round 1 aborted here), leave the particle
    state where the CPU classes leave it, write its surface meshes, and produce the surface the mesher's own CPU loop
    produces.  (The plain reference build cannot run this setting with OpenCL disabled: CLScalarField::addPoints lacks a
    `return` after its NoCL branch, src/clscalarfield.cpp:75-77, and falls into the OpenCL path -- so the CPU side of
    the comparison is the run without mesh output plus IsotropicParticleMesher without an accelerator.)"""
    ref, drop = libs
    libc = ctypes.CDLL(None)
    out = []
    for lib in (ref, drop):
        libc.srand(1)
        sim = lib.sim((32, 32, 32), 0.25)
        sim.add_fluid_sphere((4.0, 4.0, 4.0), 5.0)
        sim.add_body_force((0.0, -25.0, 0.0))
        files = []
        if lib is drop:
            bakedir = sim.enable_mesh_output()
            sim.set_accel(True, True)
        sim.initialize()
        if lib is drop and os.path.isdir(bakedir):
            for f in os.listdir(bakedir):
                if f.endswith(".ply"):
                    os.remove(os.path.join(bakedir, f))
        for _ in range(2):
            sim.update(1.0 / 30.0)
        if lib is drop:
            files = sorted((f, os.path.getsize(os.path.join(bakedir, f))) for f in os.listdir(bakedir) if f.endswith(".ply"))
        p, v = sim.get_particles()
        mesh = sim.mesh_particles(use_accelerator=lib is drop)
        out.append((p, v, files, mesh))
        sim.close()
    (p0, v0, _, (mv0, mt0)), (p1, v1, files, (mv1, mt1)) = out
    assert len(p0) == len(p1) > 20000
    assert np.abs(p0 - p1).max() < 1e-4 * 0.25
    assert np.abs(v0 - v1).max() < 1e-4 * max(1.0, np.abs(v0).max())
    assert len(files) >= 2 and all(sz > 1000 for _, sz in files), files           # one surface mesh per frame was written
    # same iso-surface: the threshold (1.0) only limits how far ABOVE the iso level (0.5) a node may grow, and the rules
    # differ between the reference's own paths (gfs_add_points); topology and extent must agree, vertex positions to a
    # fraction of a cell
    assert mt0 > 1000 and abs(mt0 - mt1) <= 0.05 * mt0, (mt0, mt1)
    assert np.abs(mv0.min(0) - mv1.min(0)).max() < 0.25 * 0.5 and np.abs(mv0.max(0) - mv1.max(0)).max() < 0.25 * 0.5
    assert np.abs(mv0.mean(0) - mv1.mean(0)).max() < 0.25 * 0.1


@pytest.mark.parametrize("n,frames,fast", [(32, 3, True), (32, 2, False), (64, 1, True)])
def test_resident_fluidsimulation_step_parity(resident_lib, n, frames, fast):
    """The UNMODIFIED reference simulator with stages 1, 5-9, 11, 12 of _stepFluid (src/fluidsimulation.cpp:3262-3390) on the
    device-resident path (dropin/fluidsimulation_resident.cpp: gfs_set_particles once, gfs_p2g, gfs_extrapolate,
    gfs_apply_body_force, gfs_pressure_solve, gfs_apply_pressure, gfs_g2p_advect) against the same simulator on its CPU
    paths: N frames of FluidSimulation::update() at 32^3 and at BASELINE configs[0]'s 64^3.
    Particles are compared as sorted sets (the resident path keeps its own particle order; the reference shuffles its
    own every substep anyway), grids element by element.  The reference's own stage timers are printed side by side."""
    from oracle.pyoracle import RefSim
    ref, res = resident_lib
    libc = ctypes.CDLL(None)
    dx = 8.0 / n
    out, times = [], []
    for lib in (ref, res):
        libc.srand(1)
        sim = lib.sim((n, n, n), dx)
        sim.add_fluid_sphere((4.0, 4.0, 4.0), 5.0 if n == 32 else 6.0)
        sim.add_body_force((0.0, -25.0, 0.0))
        if lib is res:
            sim.set_accel(fast, fast)
        log = sim.log_path()
        if os.path.exists(log):
            os.remove(log)
        sim.initialize()
        for _ in range(frames):
            sim.update(1.0 / 30.0)
        p, v = sim.get_particles()
        out.append((p, v, sim.get_material(), sim.get_fields()))
        times.append(RefSim.stage_times(log))
        sim.close()
    (p0, v0, m0, f0), (p1, v1, m1, f1) = out
    (t0, s0), (t1, s1) = times
    print("\nstage seconds over %d/%d substeps (CPU reference | resident CUDA path), %d^3, %d particles:" % (s0, s1, n, len(p0)))
    for nm in t0:
        print("  %-30s %9.4f | %9.4f" % (nm, t0[nm], t1.get(nm, float("nan"))))
    assert len(p0) == len(p1) > 20000 and s0 == s1 > 0

    # match the two particle sets by nearest neighbour (the orders differ and cannot be aligned by sorting: thousands of
    # particles share a coordinate to within the tolerance)
    from scipy.spatial import cKDTree
    dist, idx = cKDTree(p1).query(p0)
    # several substeps through P2G -> pressure solve -> extrapolation -> G2P: fp32-level differences of the splat are
    # carried through the solver (same bar as test_whole_simulator_with_cuda_accelerators)
    tol_p = 1e-4 * dx if fast else 1e-5 * dx
    assert len(np.unique(idx)) > 0.999 * len(p0)                     # one-to-one up to coincident pairs
    # (isolated outliers are legitimate: a face whose splat weight sits at the 1e-9 `isValueSet` threshold, or a sample at
    # a cell face, can fall on the other side of a discontinuous rule after an fp32-level difference)
    assert np.median(dist) < tol_p and (dist < 50 * tol_p).mean() > 0.9999 and dist.max() < dx
    dv = np.abs(v0 - v1[idx]).max(1)
    assert np.median(dv) < 1e-4 * max(1.0, np.abs(v0).max()) and (dv < 1e-2 * max(1.0, np.abs(v0).max())).mean() > 0.999
    assert (m0 != m1).mean() < 1e-3
    for a, b in zip(f0, f1):
        assert np.abs(a - b).max() < 1e-3 * max(1.0, np.abs(a).max()) and np.median(np.abs(a - b)) < 1e-5


def test_resident_simulation_with_solids_inflow_and_variable_force(resident_lib):
    """The resident build on a scene that takes its less travelled paths: interior solid cells (solid corrections of the
    pressure right-hand side, collision resolve), an inflow source (the host edits the particle vector every substep, so
    the device set is re-uploaded) and a variable body-force field (a host callback: stage 6 makes its round trip through
    the reference's own loops).  Both builds call rand() at the same places, so the emitted particles coincide."""
    ref, res = resident_lib
    libc = ctypes.CDLL(None)
    n, dx = 32, 0.25
    kk, jj, ii = np.meshgrid(np.arange(2, 6), np.arange(1, 5), np.arange(12, 18), indexing="ij")
    solid = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], 1).astype(np.int32)
    out = []
    for lib in (ref, res):
        libc.srand(5)
        sim = lib.sim((n, n, n), dx)
        sim.add_solid_cells(solid)
        sim.add_fluid_cuboid((0.25, 0.25, 0.25), 7.5, 2.0, 7.5)
        sim.add_body_force((0.0, -25.0, 0.0))
        sim.add_swirl_force()
        sim.add_inflow_source(0, (4.0, 5.5, 4.0), 0.8, 0.0, 0.0, (0.5, -2.0, 0.25))
        if lib is res:
            sim.set_accel(True, True)
        sim.initialize()
        for _ in range(2):
            sim.update(1.0 / 30.0)
        p, v = sim.get_particles()
        out.append((p, v, sim.get_material(), sim.get_fields()))
        sim.close()
    (p0, v0, m0, f0), (p1, v1, m1, f1) = out
    assert len(p0) == len(p1) > 40000
    from scipy.spatial import cKDTree
    dist, idx = cKDTree(p1).query(p0)
    tol_p = 1e-4 * dx
    assert np.median(dist) < tol_p and (dist < 50 * tol_p).mean() > 0.999 and dist.max() < dx
    dv = np.abs(v0 - v1[idx]).max(1)
    assert np.median(dv) < 1e-4 * max(1.0, np.abs(v0).max()) and (dv < 1e-2 * max(1.0, np.abs(v0).max())).mean() > 0.999
    assert (m0 != m1).mean() < 1e-3
    for a, b in zip(f0, f1):
        assert np.abs(a - b).max() < 1e-2 * max(1.0, np.abs(a).max()) and np.median(np.abs(a - b)) < 1e-5
    assert np.abs(f0[0]).max() > 0.5 and np.abs(f0[2]).max() > 0.5          # the swirl did something


def test_sources_emission_and_outflow_match_reference(libs):
    """gfs_emit_from_sources / gfs_remove_in_sources against FluidSimulation::_updateFluidSources of the unmodified reference
    (src/fluidsimulation.cpp:1771-1879) on the same particles and material grid: a spherical and a cuboid inflow source
    (one partly over fluid, one over air) and a cuboid outflow source.  The reference jitters new particles with rand();
    the set of emitting half-dx sub-cells, the velocities and the surviving old particles must be identical, positions
    agree to within the jitter."""
    from gridfluidsim3d_b200 import capi
    ref, _ = libs
    n, dx = 32, 0.25
    ctypes.CDLL(None).srand(3)
    sim = ref.sim((n, n, n), dx)
    sim.add_fluid_cuboid((0.25, 0.25, 0.25), 7.5, 2.6, 7.5)
    sim.add_body_force((0.0, -25.0, 0.0))
    inflow = [dict(kind=0, p=(2.1, 2.9, 2.3), a=0.93, velocity=(1.5, -0.5, 0.25)),
              dict(kind=1, p=(4.6, 4.1, 5.2), a=1.3, b=0.8, c=1.1, velocity=(-0.75, 0.0, 0.5))]
    outflow = [dict(kind=1, p=(5.5, 0.3, 0.4), a=1.7, b=1.2, c=2.3)]
    sim.initialize()
    sim.update(1.0 / 30.0)                       # particles off their seeding lattice, material = last classification
    for s_ in inflow:                            # (added now, so that this very call of _updateFluidSources is their first)
        sim.add_inflow_source(s_["kind"], s_["p"], s_["a"], s_.get("b", 0.0), s_.get("c", 0.0), s_["velocity"])
    for s_ in outflow:
        sim.add_outflow_source(s_["kind"], s_["p"], s_["a"], s_.get("b", 0.0), s_.get("c", 0.0))
    p0, v0 = sim.get_particles()
    mat = sim.get_material()
    sim.update_fluid_sources()
    p1, v1 = sim.get_particles()
    sim.close()

    c = capi.Context(0)
    c.domain_init((n, n, n), dx); c.set_material(mat); c.set_sources(inflow)
    c.set_particles(p0, v0)
    jitter = 0.25 * 0.1 * dx
    emitted = c.emit_from_sources(jitter, seed=7)
    removed = c.remove_in_sources(outflow)
    pg, vg = c.get_particles()
    c.close()
    assert emitted > 500 and removed > 500
    assert len(pg) == len(p1) == len(p0) + emitted - removed

    def rows(p, v):
        a = np.ascontiguousarray(np.concatenate([p, v], 1))
        return np.sort(a.view([("f%d" % i, "f4") for i in range(6)]).reshape(-1), order=["f%d" % i for i in range(6)])
    old = rows(p0, v0)
    is_old_ref = np.isin(rows(p1, v1), old)
    is_old_gpu = np.isin(rows(pg, vg), old)
    assert is_old_ref.sum() == is_old_gpu.sum() == len(p0) - removed
    # survivors of the old set: identical
    assert np.array_equal(rows(p1, v1)[is_old_ref], rows(pg, vg)[is_old_gpu])

    # the new particles: same (half-dx sub-cell, velocity) multiset, positions within the jitter of each other
    def new_of(p, v):
        a = np.ascontiguousarray(np.concatenate([p, v], 1)).view([("f%d" % i, "f4") for i in range(6)]).reshape(-1)
        keep = ~np.isin(a, old)
        return p[keep], v[keep]
    (pn_ref, vn_ref), (pn_gpu, vn_gpu) = new_of(p1, v1), new_of(pg, vg)
    assert len(pn_ref) == len(pn_gpu) > 500

    def keyed(p, v):
        sub = np.floor(p.astype(np.float64) / (0.5 * dx)).astype(np.int64)
        key = sub[:, 0] + 4 * n * (sub[:, 1] + 4 * n * sub[:, 2])
        o = np.lexsort((v[:, 2], v[:, 1], v[:, 0], key))
        return key[o], p[o], v[o]
    (k_ref, pr, vr), (k_gpu, pq, vq) = keyed(pn_ref, vn_ref), keyed(pn_gpu, vn_gpu)
    assert np.array_equal(k_ref, k_gpu)
    assert np.array_equal(vr.view(np.uint32), vq.view(np.uint32))
    assert np.abs(pr - pq).max() <= 2.0 * jitter * (1 + 1e-5)
