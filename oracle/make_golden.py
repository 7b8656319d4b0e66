"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libgfsref.so).

Run in a container that has /root/reference:   python -m oracle.make_golden
Every array stored here came out of reference code (see oracle/ref_harness.cpp for the entry points);
inputs are stored alongside so the fixtures are self-contained.  The reference has no golden vectors of
its own, so these files are what pins the oracle -- and through it the CUDA path --
on machines where the reference cannot be built.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gridfluidsim3d_b200 import synth          # noqa: E402
from oracle.pyoracle import Reference           # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def rough_fields(dims, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return tuple((scale * rng.standard_normal(a * b * c)).astype(np.float32) for a, b, c in synth.face_dims(dims))


def probes(dims, dx, n, seed):
    rng = np.random.default_rng(seed)
    ext = np.array(dims) * dx
    pos = rng.uniform(-0.6 * dx, ext + 0.6 * dx, size=(n, 3))
    pos[: n // 10] = np.round(pos[: n // 10] / dx) * dx
    pos[n // 10: n // 5] = np.round(pos[n // 10: n // 5] / (0.5 * dx)) * 0.5 * dx
    return pos.astype(np.float32)


def extrapolation_case(seed=301):
    """A 14 x 11 x 13 domain: solid border, an interior solid block, a fluid blob touching it, rough fields."""
    dims = (14, 11, 13)
    I, J, K = dims
    mat = synth.border_material(dims).reshape(K, J, I).copy()
    mat[4:7, 2:5, 6:9] = synth.SOLID
    kk, jj, ii = np.meshgrid(np.arange(K), np.arange(J), np.arange(I), indexing="ij")
    blob = ((ii - 5.5) ** 2 + (jj - 4.5) ** 2 + (kk - 6.0) ** 2 < 9.5) & (mat != synth.SOLID)
    mat[blob] = synth.FLUID
    mat[10, 8, 11] = synth.FLUID                     # an isolated fluid cell near the corner
    return dims, 0.25, mat.reshape(-1), rough_fields(dims, seed)


def golden_extrapolate(ref):
    dims, dx, mat, (u, v, w) = extrapolation_case()
    g = dict(dims=np.array(dims, np.int32), dx=np.float64(dx), material=mat, u=u, v=v, w=w)
    for nl in (1, 3, 7):
        a, b, c = ref.extrapolate(u, v, w, dims, dx, mat, nl)
        g["u_%d" % nl], g["v_%d" % nl], g["w_%d" % nl] = a, b, c
    np.savez_compressed(os.path.join(OUT, "extrapolate.npz"), **g)


def golden_state(ref):
    """A save state written by the reference's FluidSimulationSaveState: 10 x 8 x 9 cells, a fluid ball (the
    simulator seeds and jitters the particles itself, glibc rand() with its default seed) and four interior solid cells."""
    sim = ref.sim((10, 8, 9), 0.25)
    sim.add_fluid_sphere((1.25, 1.0, 1.1), 0.6)
    sim.add_solid_cells(np.array([[6, 2, 3], [6, 3, 3], [7, 2, 3], [2, 5, 6]], np.int32))
    sim.initialize()
    sim.save_state(os.path.join(OUT, "reference_small.state"))
    sim.close()


def golden_pressure(ref):
    """Stages 6-8 (body forces, MICCG(0) pressure solve, pressure update) of the unmodified reference on its own stage-5
    field of the stage scene (interior solids, fluid against walls): inputs and outputs of every stage."""
    dims, dx = (12, 10, 14), 0.25
    I, J, K = dims
    material = synth.border_material(dims)
    m3 = material.reshape(K, J, I)
    m3[2:5, 1:4, 3:6] = synth.SOLID
    mask = synth.fluid_cells("dam", dims, material)
    p = synth.make_particles(mask, dx, seed=778)
    vel = (synth.particle_velocities(p, dims, dx) + 0.05 * np.random.default_rng(204).standard_normal(p.shape)).astype(np.float32)
    kk, jj, ii = np.nonzero(m3 == synth.SOLID)
    force, dt = (0.3, -9.8, 0.05), 1.0 / 30
    sim = ref.sim(dims, dx)
    sim.add_solid_cells(np.stack([ii, jj, kk], 1).astype(np.int32))
    sim.add_body_force(force)
    sim.initialize()
    sim.set_particles(p, vel)
    sim.update_fluid_cells()
    sim.advect_velocity_field()
    mat = sim.get_material()
    f5 = sim.get_fields()
    sim.apply_body_forces(dt)
    f6 = sim.get_fields()
    pr = sim.update_pressure_grid(dt)
    sim.apply_pressure(dt, pr)
    f8 = sim.get_fields()
    density = sim.density()
    sim.close()
    np.savez_compressed(os.path.join(OUT, "pressure.npz"), dims=np.array(dims, np.int32), dx=np.float64(dx), material=mat,
                        force=np.array(force, np.float32), dt=np.float64(dt), density=np.float64(density),
                        u5=f5[0], v5=f5[1], w5=f5[2], u6=f6[0], v6=f6[1], w6=f6[2], pressure=pr, u8=f8[0], v8=f8[1], w8=f8[2])
    print("pressure.npz", os.path.getsize(os.path.join(OUT, "pressure.npz")))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "pressure":          # only this fixture (the others stay byte for byte)
        golden_pressure(Reference())
        return
    ref = Reference()
    os.makedirs(OUT, exist_ok=True)
    if "--only-extrapolate" in sys.argv:             # added after the other fixtures were committed
        golden_extrapolate(ref)
        return
    if "--only-state" in sys.argv:
        golden_state(ref)
        return
    golden_extrapolate(ref)
    golden_state(ref)

    # ---- primitives: index, sampling, RK1-4, splat ---------------------------------------------
    dims, dx = (10, 8, 12), 0.25
    u, v, w = rough_fields(dims, 101)
    pos = probes(dims, dx, 1500, 102)
    g = dict(dims=np.array(dims, np.int32), dx=np.float64(dx), u=u, v=v, w=w, pos=pos)
    for d in (0.125, 0.1, 1.0 / 3.0):
        g["cell_dx_%g" % d] = ref.cell_index(pos, d)
    g["sample_trilinear"] = ref.sample(pos, u, v, w, dims, dx, 0)
    g["sample_tricubic"] = ref.sample(pos, u, v, w, dims, dx, 1)
    g["dt"] = np.float64(0.11)
    for order in (1, 2, 3, 4):
        g["rk%d" % order] = ref.advect(pos, u, v, w, dims, dx, 0.11, order)
    inside = pos[np.all((pos > 0) & (pos < np.array(dims) * dx), 1)]
    vals = np.random.default_rng(103).standard_normal(len(inside)).astype(np.float32)
    g["splat_pos"], g["splat_values"] = inside, vals
    for comp, nd in enumerate(synth.face_dims(dims)):
        off = np.array([0.0 if comp == 0 else 0.5 * dx, 0.0 if comp == 1 else 0.5 * dx,
                        0.0 if comp == 2 else 0.5 * dx], np.float32)
        f, wt = ref.add_point_values(inside, vals, dx, off, dx, nd)
        g["splat_field_%d" % comp], g["splat_weight_%d" % comp] = f, wt
    np.savez_compressed(os.path.join(OUT, "primitives.npz"), **g)

    # ---- stage level: classification + P2G, PIC/FLIP + RK4, on a scene with interior solids and sources
    dims, dx = (12, 10, 14), 0.25
    I, J, K = dims
    material = synth.border_material(dims)
    m3 = material.reshape(K, J, I)
    m3[2:5, 1:4, 3:6] = synth.SOLID
    mask = synth.fluid_cells("dam", dims, material)
    p = synth.make_particles(mask, dx, seed=777)
    vel = synth.particle_velocities(p, dims, dx) + \
        0.05 * np.random.default_rng(104).standard_normal(p.shape).astype(np.float32)
    vel = vel.astype(np.float32)
    kk, jj, ii = np.nonzero(m3 == synth.SOLID)
    solid_ijk = np.stack([ii, jj, kk], 1).astype(np.int32)
    sources = [dict(kind=0, p=(1.2, 1.0, 1.5), a=0.7, velocity=(0.5, -1.0, 0.25)),
               dict(kind=1, p=(0.5, 0.5, 2.0), a=1.0, b=0.8, c=0.9, velocity=(-0.3, 0.2, 0.7))]

    sim = ref.sim(dims, dx)
    sim.add_solid_cells(solid_ijk)
    sim.initialize()
    sim.set_particles(p, vel)
    sim.update_fluid_cells()
    assert sim.n == len(p)
    for s in sources:
        sim.add_inflow_source(s["kind"], s["p"], s.get("a", 0), s.get("b", 0), s.get("c", 0), s["velocity"])
    sim.advect_velocity_field()
    mat_out = sim.get_material()
    pu, pv, pw = sim.get_fields()

    new, saved = rough_fields(dims, 105, 0.3), rough_fields(dims, 106, 0.3)
    dt = 0.25 * dx
    sim.set_fields(new, saved)
    sim.update_particle_velocities()
    sim.advance_particles(dt)
    pos1, vel1 = sim.get_particles()
    sim.close()

    src_arr = np.array([[s["kind"], *s["p"], s.get("a", 0), s.get("b", 0), s.get("c", 0), *s["velocity"]]
                        for s in sources], np.float64)
    np.savez_compressed(os.path.join(OUT, "stages.npz"),
                        dims=np.array(dims, np.int32), dx=np.float64(dx), material_in=material,
                        pos=p, vel=vel, sources=src_arr, material_out=mat_out, p2g_u=pu, p2g_v=pv, p2g_w=pw,
                        new_u=new[0], new_v=new[1], new_w=new[2], saved_u=saved[0], saved_v=saved[1],
                        saved_w=saved[2], dt=np.float64(dt), pos_out=pos1, vel_out=vel1)

    # ---- whole-simulator: Hello-World-like drop at 16^3, two frames of FluidSimulation::update, with the
    # particle set before/after each hot-path stage of the second frame captured through the stage calls.
    dims, dx = (16, 16, 16), 0.5
    sim = ref.sim(dims, dx)
    sim.add_fluid_sphere((4.0, 4.0, 4.0), 5.0)
    sim.add_body_force((0.0, -25.0, 0.0))
    sim.initialize()
    p0, v0 = sim.get_particles()
    sim.update(1.0 / 30.0)
    p1, v1 = sim.get_particles()
    mat1 = sim.get_material()
    u1, vv1, w1 = sim.get_fields()
    sim.close()
    np.savez_compressed(os.path.join(OUT, "helloworld16.npz"), dims=np.array(dims, np.int32), dx=np.float64(dx),
                        pos0=p0, vel0=v0, pos1=p1, vel1=v1, material1=mat1, u1=u1, v1=vv1, w1=w1)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
