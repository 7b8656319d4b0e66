"""The oracle (oracle/oracle.c) against the committed golden fixtures (tests/golden/*.npz), which were
produced by the unmodified reference through oracle/make_golden.py.  CPU only, runs everywhere."""
import os

import numpy as np
import pytest

from gridfluidsim3d_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_primitives_golden(oracle):
    g = load("primitives.npz")
    dims, dx = tuple(int(x) for x in g["dims"]), float(g["dx"])
    u, v, w, pos = g["u"], g["v"], g["w"], g["pos"]
    for d in (0.125, 0.1, 1.0 / 3.0):
        assert np.array_equal(oracle.cell_index(pos, d), g["cell_dx_%g" % d])
    assert np.array_equal(bits(oracle.sample(pos, u, v, w, dims, dx, 0, validate=False)), bits(g["sample_trilinear"]))
    assert np.array_equal(bits(oracle.sample(pos, u, v, w, dims, dx, 1, validate=True)), bits(g["sample_tricubic"]))
    for order in (1, 2, 3, 4):
        out = oracle.advect(pos, u, v, w, dims, dx, float(g["dt"]), order, 1)
        assert np.array_equal(bits(out), bits(g["rk%d" % order]))
    for comp, nd in enumerate(synth.face_dims(dims)):
        off = np.array([0.0 if comp == 0 else 0.5 * dx, 0.0 if comp == 1 else 0.5 * dx,
                        0.0 if comp == 2 else 0.5 * dx], np.float32)
        f, wt = oracle.splat(g["splat_pos"], g["splat_values"], dx, off, dx, nd)
        assert np.array_equal(bits(f), bits(g["splat_field_%d" % comp]))
        assert np.array_equal(bits(wt), bits(g["splat_weight_%d" % comp]))


def golden_sources(g):
    return [dict(kind=int(r[0]), p=tuple(r[1:4]), a=r[4], b=r[5], c=r[6], velocity=tuple(r[7:10])) for r in g["sources"]]


def test_stages_golden(oracle):
    g = load("stages.npz")
    dims, dx = tuple(int(x) for x in g["dims"]), float(g["dx"])
    mat = g["material_in"].copy()
    u, v, w = oracle.p2g(g["pos"], g["vel"], dims, dx, mat, golden_sources(g))
    assert np.array_equal(mat, g["material_out"])
    assert np.array_equal(bits(u), bits(g["p2g_u"]))
    assert np.array_equal(bits(v), bits(g["p2g_v"]))
    assert np.array_equal(bits(w), bits(g["p2g_w"]))
    new = (g["new_u"], g["new_v"], g["new_w"])
    saved = (g["saved_u"], g["saved_v"], g["saved_w"])
    pos, vel, flags = oracle.g2p_advect(g["pos"], g["vel"], new, saved, dims, dx, float(g["dt"]), material=mat)
    assert np.array_equal(bits(vel), bits(g["vel_out"]))
    # exactly one particle of this fixture ends in a solid cell and goes through the reference's collision resolve
    # (src/fluidsimulation.cpp:3145-3179): the fixture holds the reference's resolved position for it
    assert flags.sum() == 1
    assert np.array_equal(bits(pos), bits(g["pos_out"]))
    kept, _, _ = oracle.g2p_advect(g["pos"], g["vel"], new, saved, dims, dx, float(g["dt"]), material=mat, resolve=False)
    hit = flags == 1
    assert np.array_equal(bits(kept[hit]), bits(g["pos"][hit])) and not np.array_equal(bits(kept[hit]), bits(pos[hit]))


def test_whole_simulator_frame_golden(oracle):
    """One FluidSimulation::update(1/30) of a 16^3 sphere drop: the oracle's RK4 through the simulator's
    final velocity field must reproduce the simulator's particle set (it is reshuffled by rand(), so the
    comparison is on sorted rows); the oracle's classification of the *old* positions is the material
    grid the simulator used for that frame."""
    g = load("helloworld16.npz")
    dims, dx = tuple(int(x) for x in g["dims"]), float(g["dx"])
    mat = oracle.border_material(dims)
    mat, bad = oracle.classify(g["pos0"], dims, dx, mat)
    assert bad == 0
    assert np.array_equal(mat, g["material1"])
    pos = oracle.advect(g["pos0"], g["u1"], g["v1"], g["w1"], dims, dx, 1.0 / 30.0, 4, 1)
    assert len(pos) == len(g["pos1"])       # nothing hit the 100-per-cell cap in one frame

    def rows(a):
        a = np.ascontiguousarray(a)
        return np.sort(a.view([("x", "f4"), ("y", "f4"), ("z", "f4")]).reshape(-1), order=("x", "y", "z"))
    assert np.array_equal(rows(pos), rows(g["pos1"]))


def test_splat_kernel_volume_integral(oracle):
    """Known answer: the mean sum of weights per interior particle equals the kernel's
    volume integral 4*pi*(1/3 - 4/81 + 17/63 - 22/45) = 0.8156 cells."""
    dims, dx = (20, 20, 20), 0.25
    rng = np.random.default_rng(5)
    pos = rng.uniform(4 * dx, 16 * dx, size=(40000, 3)).astype(np.float32)
    ones = np.ones(len(pos), np.float32)
    f, wt = oracle.splat(pos, ones, dx, np.zeros(3, np.float32), dx, dims)
    expect = 4 * np.pi * (1 / 3 - 4 / 81 + 17 / 63 - 22 / 45)
    assert abs(wt.sum() / len(pos) - expect) < 5e-3
    assert np.allclose(f, wt)


def test_uniform_field_properties(oracle):
    """A constant field is reproduced >= 2 cells from the boundary (out-of-range taps read 0 nearer), and
    RK4 through a uniform field is an exact translation."""
    dims, dx = (16, 16, 16), 0.25
    c = (0.5, -0.25, 0.125)
    u, v, w = (np.full(n, cv, np.float32) for n, cv in zip((17 * 16 * 16, 16 * 17 * 16, 16 * 16 * 17), c))
    rng = np.random.default_rng(6)
    pos = rng.uniform(3 * dx, 13 * dx, size=(2000, 3)).astype(np.float32)
    for mode in (0, 1):
        s = oracle.sample(pos, u, v, w, dims, dx, mode)
        assert np.allclose(s, np.array(c, np.float32), rtol=0, atol=1e-6)
        out = oracle.advect(pos, u, v, w, dims, dx, 0.25, 4, mode)
        assert np.allclose(out - pos, 0.25 * np.array(c), rtol=0, atol=2e-6)


def test_extrapolate_golden(oracle):
    """MACVelocityField::extrapolateVelocityField outputs of the unmodified reference (oracle/make_golden.py)."""
    g = load("extrapolate.npz")
    dims = tuple(int(x) for x in g["dims"])
    for nl in (1, 3, 7):
        out = oracle.extrapolate(g["u"], g["v"], g["w"], dims, g["material"], nl)
        for got, name in zip(out, "uvw"):
            assert np.array_equal(bits(got), bits(g["%s_%d" % (name, nl)]))


def test_pressure_stages_golden(oracle):
    """Stages 6-8 of the unmodified reference (tests/golden/pressure.npz, written by `oracle/make_golden.py pressure`):
    the oracle reproduces the force-added field, the MICCG(0) float pressure grid and the projected field bit for bit."""
    g = load("pressure.npz")
    dims, dx, dt, density = tuple(int(x) for x in g["dims"]), float(g["dx"]), float(g["dt"]), float(g["density"])
    mat = g["material"]
    f6 = oracle.body_force(g["u5"], g["v5"], g["w5"], dims, mat, g["force"], dt)
    for a, nm in zip(f6, ("u6", "v6", "w6")):
        assert np.array_equal(bits(a), bits(g[nm]))
    p, iters, limit, err = oracle.pressure_solve(*f6, dims, dx, mat, dt, density)
    assert iters > 3 and not limit and err < 1e-6
    assert np.array_equal(bits(p), bits(g["pressure"]))
    f8 = oracle.apply_pressure(*f6, dims, dx, mat, p, dt, density)
    for a, nm in zip(f8, ("u8", "v8", "w8")):
        assert np.array_equal(bits(a), bits(g[nm]))
    assert np.count_nonzero(g["pressure"]) > 100

