"""Pin the C restatement (oracle/oracle.c) bit-for-bit against the UNMODIFIED reference sources
(oracle/_ref/libgfsref.so, built from /root/reference by oracle/Makefile).  CPU only.

Skipped where the reference library is absent (a checkout without /root/reference and without a
prebuilt oracle/_ref); the committed fixtures in tests/golden/ (test_oracle_golden.py) cover that case.
"""
import numpy as np
import pytest

from gridfluidsim3d_b200 import synth


def rough_fields(dims, seed):
    """Zero-mean random fields: every tap matters, nothing cancels by symmetry."""
    rng = np.random.default_rng(seed)
    return tuple(rng.standard_normal(a * b * c).astype(np.float32) for a, b, c in synth.face_dims(dims))


def probe_positions(dims, dx, n, seed):
    """Random positions covering the interior, the border cells, the exact domain faces and outside."""
    rng = np.random.default_rng(seed)
    ext = np.array(dims) * dx
    pos = rng.uniform(-0.6 * dx, ext + 0.6 * dx, size=(n, 3))
    pos[: n // 10] = np.round(pos[: n // 10] / dx) * dx                      # exactly on cell faces
    pos[n // 10: n // 5] = np.round(pos[n // 10: n // 5] / (0.5 * dx)) * 0.5 * dx   # on faces and centres
    return pos.astype(np.float32)


DIMS, DX = (12, 10, 14), 0.25


def test_cell_index_bit_exact(oracle, reference):
    for dx in (0.125, 0.0625, 0.1, 0.3, 1.0 / 3.0):
        pos = probe_positions(DIMS, dx, 20000, 1)
        assert np.array_equal(oracle.cell_index(pos, dx), reference.cell_index(pos, dx))


@pytest.mark.parametrize("mode", [0, 1])
def test_sample_bit_exact(oracle, reference, mode):
    u, v, w = rough_fields(DIMS, 2)
    pos = probe_positions(DIMS, DX, 20000, 3)
    a = oracle.sample(pos, u, v, w, DIMS, DX, mode, validate=(mode == 1))
    b = reference.sample(pos, u, v, w, DIMS, DX, mode)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_sample_validate_nan_inf(oracle, reference):
    u, v, w = rough_fields(DIMS, 4)
    u = u.copy(); u[::7] = np.inf; v = v.copy(); v[::11] = np.nan
    pos = probe_positions(DIMS, DX, 5000, 5)
    a = oracle.sample(pos, u, v, w, DIMS, DX, 1, validate=True)
    b = reference.sample(pos, u, v, w, DIMS, DX, 1)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.isfinite(a).all()


@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_advect_bit_exact(oracle, reference, order):
    u, v, w = rough_fields(DIMS, 6)
    pos = probe_positions(DIMS, DX, 8000, 7)
    for dt in (1.0 / 30.0, 0.1, 0.37):
        a = oracle.advect(pos, u, v, w, DIMS, DX, dt, order, 1)
        b = reference.advect(pos, u, v, w, DIMS, DX, dt, order)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("comp", [0, 1, 2])
def test_splat_bit_exact(oracle, reference, comp):
    dx = 0.25
    nd = synth.face_dims(DIMS)[comp]
    off = np.array([0.0 if comp == 0 else 0.5 * dx, 0.0 if comp == 1 else 0.5 * dx,
                    0.0 if comp == 2 else 0.5 * dx], np.float32)
    pos = probe_positions(DIMS, dx, 6000, 8 + comp)
    pos = pos[np.all((pos > 0) & (pos < np.array(DIMS) * dx), 1)]
    vals = np.random.default_rng(9).standard_normal(len(pos)).astype(np.float32)
    fa, wa = oracle.splat(pos, vals, dx, off, dx, nd)
    fb, wb = reference.add_point_values(pos, vals, dx, off, dx, nd)
    assert np.array_equal(fa.view(np.uint32), fb.view(np.uint32))
    assert np.array_equal(wa.view(np.uint32), wb.view(np.uint32))
    # a different radius exercises the general index-bound formula (src/grid3d.h:350-371)
    fa, wa = oracle.splat(pos, vals, 1.7 * dx, off, dx, nd)
    fb, wb = reference.add_point_values(pos, vals, 1.7 * dx, off, dx, nd)
    assert np.array_equal(fa.view(np.uint32), fb.view(np.uint32))
    assert np.array_equal(wa.view(np.uint32), wb.view(np.uint32))


def _scene(name, interior_solids=False, seed=12345):
    s = synth.make_scene(name, seed=seed)
    if interior_solids:
        I, J, K = s["dims"]
        m = s["material"].reshape(K, J, I)
        m[2:5, 1:4, 3:6] = synth.SOLID            # a solid block in a corner of the fluid
        mask = synth.fluid_cells(synth.CONFIGS[name][2], s["dims"], s["material"])
        s["pos"] = synth.make_particles(mask, s["dx"], seed)
        s["vel"] = synth.particle_velocities(s["pos"], s["dims"], s["dx"])
        kk, jj, ii = np.nonzero(m == synth.SOLID)
        s["solid_ijk"] = np.stack([ii, jj, kk], 1).astype(np.int32)
    return s


def _ref_sim(reference, s, sources=()):
    sim = reference.sim(s["dims"], s["dx"])
    if "solid_ijk" in s:
        sim.add_solid_cells(s["solid_ijk"])
    for src in sources:
        sim.add_inflow_source(src["kind"], src["p"], src.get("a", 0), src.get("b", 0), src.get("c", 0), src["velocity"])
    sim.initialize()
    sim.set_particles(s["pos"], s["vel"])
    return sim


@pytest.mark.parametrize("name,solids", [("tiny16", False), ("slab24", True)])
def test_classification_and_p2g_bit_exact(oracle, reference, name, solids):
    """Stage 1 + stage 5 of the reference's _stepFluid vs orc_p2g on the same particles."""
    s = _scene(name, solids)
    sources = [dict(kind=0, p=(1.5, 1.5, 1.5), a=0.9, velocity=(0.5, -1.0, 0.25)),
               dict(kind=1, p=(2.0, 0.5, 2.0), a=1.0, b=0.8, c=1.3, velocity=(-0.3, 0.2, 0.7))]
    sim = _ref_sim(reference, s)
    sim.update_fluid_cells()
    assert sim.n == len(s["pos"])                      # nothing was removed: no particle sat in a solid
    for src in sources:     # registered after stage 1, so that _updateFluidSources does not emit particles
        sim.add_inflow_source(src["kind"], src["p"], src.get("a", 0), src.get("b", 0), src.get("c", 0), src["velocity"])
    sim.advect_velocity_field()
    mat_ref = sim.get_material()
    u_ref, v_ref, w_ref = sim.get_fields()
    sim.close()

    mat = s["material"].copy()
    u, v, w = oracle.p2g(s["pos"], s["vel"], s["dims"], s["dx"], mat, sources)
    assert np.array_equal(mat, mat_ref)
    for a, b in ((u, u_ref), (v, v_ref), (w, w_ref)):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.count_nonzero(u) > 0 and np.count_nonzero(w) > 0


@pytest.mark.parametrize("name", ["tiny16", "slab24"])
def test_picflip_and_rk4_bit_exact(oracle, reference, name):
    """Stage 11 + stage 12 (no shuffle/cap) vs orc_g2p_advect; dt small enough that no particle
    reaches a solid cell, so the reference's collision resolve (outside this scope) never runs."""
    s = _scene(name)
    new, saved = rough_fields(s["dims"], 21), rough_fields(s["dims"], 22)
    new = tuple(0.3 * a for a in new)
    dt = 0.25 * s["dx"]
    sim = _ref_sim(reference, s)
    sim.update_fluid_cells()
    sim.set_fields(new, saved)
    sim.update_particle_velocities()
    sim.advance_particles(dt)
    pos_ref, vel_ref = sim.get_particles()
    mat = sim.get_material()
    sim.close()

    pos, vel, flags = oracle.g2p_advect(s["pos"], s["vel"], new, saved, s["dims"], s["dx"], dt, material=mat)
    assert flags.sum() == 0
    assert np.array_equal(vel.view(np.uint32), vel_ref.view(np.uint32))
    assert np.array_equal(pos.view(np.uint32), pos_ref.view(np.uint32))


@pytest.mark.parametrize("name,solids,dt", [("tiny16", False, 1.0 / 60), ("slab24", True, 1.0 / 30), ("slab24", False, 1.0 / 300)])
def test_body_force_pressure_solve_and_update_bit_exact(oracle, reference, name, solids, dt):
    """Stages 6, 7, 8 of _stepFluid (constant body forces, PressureSolver::solve behind _updatePressureGrid,
    _applyPressureToVelocityField) vs oracle_pressure.c on the reference's own stage-5 field: every float bit for bit --
    the restatement keeps the MICCG(0) operation order, so the CG trajectory and iteration count are the reference's."""
    s = _scene(name, solids)
    force = (0.3, -9.8, 0.05)
    sim = _ref_sim(reference, s)
    sim.add_body_force(force)
    sim.update_fluid_cells()
    sim.advect_velocity_field()
    mat = sim.get_material()
    f5 = sim.get_fields()
    sim.apply_body_forces(dt)
    f6 = sim.get_fields()
    p_ref = sim.update_pressure_grid(dt)
    sim.apply_pressure(dt, p_ref)
    f8 = sim.get_fields()
    density = sim.density()
    sim.close()

    g6 = oracle.body_force(*f5, s["dims"], mat, force, dt)
    for a, b in zip(g6, f6):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    p, iters, limit, err = oracle.pressure_solve(*g6, s["dims"], s["dx"], mat, dt, density)
    assert iters > 3 and not limit and err < 1e-6
    assert np.count_nonzero(p) > 0.9 * (mat == synth.FLUID).sum()
    assert np.array_equal(p.view(np.uint32), p_ref.view(np.uint32))
    g8 = oracle.apply_pressure(*g6, s["dims"], s["dx"], mat, p, dt, density)
    for a, b in zip(g8, f8):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # the projected field is divergence free in the fluid cells to the solver's tolerance
    I, J, K = s["dims"]
    u = g8[0].reshape(K, J, I + 1); v = g8[1].reshape(K, J + 1, I); w = g8[2].reshape(K + 1, J, I)
    div = (u[:, :, 1:] - u[:, :, :-1] + v[:, 1:, :] - v[:, :-1, :] + w[1:] - w[:-1]) / s["dx"]
    assert np.abs(div[mat.reshape(K, J, I) == synth.FLUID]).max() < 1e-3


@pytest.mark.parametrize("nlayers", [0, 1, 3, 7])
def test_extrapolate_bit_exact(oracle, reference, nlayers):
    """MACVelocityField::extrapolateVelocityField, with interior solids, a fluid blob against the
    solid border and isolated fluid cells; rough fields so that every averaged neighbour matters."""
    dims = (13, 9, 11)
    I, J, K = dims
    rng = np.random.default_rng(40 + nlayers)
    mat = synth.border_material(dims).reshape(K, J, I).copy()
    mat[3:6, 2:4, 5:8] = synth.SOLID
    fluid = (rng.random((K, J, I)) < 0.18) & (mat != synth.SOLID)
    fluid[4:8, 3:7, 2:6] |= mat[4:8, 3:7, 2:6] != synth.SOLID
    mat[fluid] = synth.FLUID
    u, v, w = rough_fields(dims, 41)
    a = oracle.extrapolate(u, v, w, dims, mat, nlayers)
    b = reference.extrapolate(u, v, w, dims, 0.25, mat, nlayers)
    for x, y in zip(a, b):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
    assert any((x != y.reshape(-1)).any() for x, y in zip(a, (u, v, w)))


@pytest.mark.parametrize("name,scale", [("slab24", 1.0), ("slab24", 4.0), ("tiny16", 2.5)])
def test_collision_resolve_bit_exact(oracle, reference, name, scale):
    """stage 12 with a step large enough that many particles are advected into solid cells (the
    border and an interior block), so FluidSimulation::_resolveParticleSolidCellCollision runs: voxel walk, ray/box
    intersection, back-off, and every "return p0" branch.  orc_g2p_advect_resolve must reproduce it bit for bit."""
    s = _scene(name, interior_solids=True)
    new, saved = rough_fields(s["dims"], 31), rough_fields(s["dims"], 32)
    dt = scale * s["dx"]
    sim = _ref_sim(reference, s)
    sim.update_fluid_cells()
    sim.set_fields(new, saved)
    sim.update_particle_velocities()
    sim.advance_particles(dt)
    pos_ref, vel_ref = sim.get_particles()
    mat = sim.get_material()
    sim.close()

    pos, vel, flags = oracle.g2p_advect(s["pos"], s["vel"], new, saved, s["dims"], s["dx"], dt, material=mat, resolve=True)
    assert flags.sum() > 20                                             # the resolve really ran
    assert np.array_equal(vel.view(np.uint32), vel_ref.view(np.uint32))
    assert np.array_equal(pos.view(np.uint32), pos_ref.view(np.uint32))
    hit = flags.astype(bool)
    moved = (pos[hit] != s["pos"][hit]).any(1)
    assert moved.any()                                                  # resolved positions, not just "keep p0"


def test_save_state_both_directions(reference, tmp_path):
    """SURVEY 8(f) rank 4 (wire format): a state written by the reference's FluidSimulationSaveState is read field for
    field by gridfluidsim3d_b200/savestate.py, and a state written by savestate.py is accepted by the reference's reader and
    by FluidSimulation(FluidSimulationSaveState&) with the same particles and solid cells."""
    from gridfluidsim3d_b200 import savestate
    s = _scene("slab24", interior_solids=True)
    sim = _ref_sim(reference, s)
    ref_path = str(tmp_path / "ref.state")
    sim.save_state(ref_path)
    p_ref, v_ref = sim.get_particles()
    mat_ref = sim.get_material()
    sim.close()

    st = savestate.read_state(ref_path)
    assert st["dims"] == tuple(s["dims"]) and st["dx"] == s["dx"] and st["frame"] == 0
    assert np.array_equal(st["pos"].view(np.uint32), p_ref.view(np.uint32))
    assert np.array_equal(st["vel"].view(np.uint32), v_ref.view(np.uint32))
    assert len(st["diffuse_pos"]) == 0 and st["brick_blob"] is None
    solid = savestate.material_from_state(st)
    assert np.array_equal(solid == synth.SOLID, mat_ref == synth.SOLID)
    assert np.array_equal(st["solid_ijk"], savestate.solid_ijk_from_material(mat_ref, s["dims"]))

    ours = str(tmp_path / "ours.state")
    savestate.write_state(ours, s["dims"], s["dx"], s["pos"], s["vel"], st["solid_ijk"], frame=3)
    back = reference.state_read(ours)
    assert back is not None and back["dims"] == tuple(s["dims"]) and back["dx"] == s["dx"] and back["frame"] == 3
    assert np.array_equal(back["pos"].view(np.uint32), s["pos"].view(np.uint32))
    assert np.array_equal(back["vel"].view(np.uint32), s["vel"].view(np.uint32))
    assert np.array_equal(back["solid_ijk"], st["solid_ijk"])
    assert open(ours, "rb").read()[37:] == open(ref_path, "rb").read()[37:]          # byte-identical payload
    sim2 = reference.sim_from_state(ours)
    assert sim2 is not None and sim2.n == len(s["pos"])
    p2, v2 = sim2.get_particles()
    assert np.array_equal(p2.view(np.uint32), s["pos"].view(np.uint32))
    assert np.array_equal(sim2.get_material() == synth.SOLID, mat_ref == synth.SOLID)
    sim2.close()
