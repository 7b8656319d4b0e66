"""Stages 6-8 on the device: constant body forces, the reference's MICCG(0) pressure solve as tile
wavefronts, the pressure update -- against oracle/oracle_pressure.c, which tests/test_oracle_vs_ref.py pins bit for bit
against the unmodified reference.

Bar: body forces and the pressure update are bit-exact given equal inputs.  The solve keeps every operation of the
reference in its order except the summation order of the two dot products per iteration, so alpha / beta differ in the
last place: the iteration count must be the reference's and the FLOAT pressure grid must agree to 2e-6 relative to the
largest pressure (measured: almost all values are bit-identical)."""
import numpy as np
import pytest

from gridfluidsim3d_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def make_case(name, solids, seed=12345):
    s = synth.make_scene(name, seed=seed)
    if solids:
        I, J, K = s["dims"]
        m = s["material"].reshape(K, J, I)
        m[2:5, 1:4, 3:6] = synth.SOLID
        m[1:3, 1:3, I // 2:I // 2 + 4] = synth.SOLID
        mask = synth.fluid_cells(synth.CONFIGS[name][2], s["dims"], s["material"])
        s["pos"] = synth.make_particles(mask, s["dx"], seed)
        s["vel"] = synth.particle_velocities(s["pos"], s["dims"], s["dx"])
    return s


def stage5(ctx, s):
    ctx.domain_init(s["dims"], s["dx"])
    ctx.set_material(s["material"])
    ctx.set_particles(s["pos"], s["vel"])
    ctx.sort_index()
    ctx.p2g(capi.FAST)
    return ctx.get_material(), ctx.get_field(capi.FIELD_P2G)


@pytest.mark.parametrize("name,solids,dt", [("tiny16", False, 1.0 / 60), ("slab24", True, 1.0 / 30), ("small32", False, 1.0 / 30),
                                            ("odd20", False, 1.0 / 120), ("hello64", False, 1.0 / 30)])
def test_body_force_pressure_and_update_match_oracle(ctx, oracle, name, solids, dt):
    s = make_case(name, solids)
    force = (0.3, -9.8, 0.05)
    mat, f5 = stage5(ctx, s)

    ctx.apply_body_force(capi.FIELD_P2G, force, dt)
    f6 = ctx.get_field(capi.FIELD_P2G)
    g6 = oracle.body_force(*f5, s["dims"], mat, force, dt)
    for a, b in zip(f6, g6):
        assert np.array_equal(bits(a), bits(b))

    iters, resid = ctx.pressure_solve(capi.FIELD_P2G, dt)
    p = ctx.get_pressure()
    p_ref, it_ref, limit, err_ref = oracle.pressure_solve(*g6, s["dims"], s["dx"], mat, dt)
    assert not limit and it_ref > 3
    assert iters == it_ref                                       # same CG trajectory
    assert resid < 1e-6 and abs(resid - err_ref) <= 1e-3 * err_ref + 1e-12
    scale = np.abs(p_ref).max()
    assert scale > 0
    assert np.abs(p - p_ref).max() <= 2e-6 * scale
    same = (bits(p) == bits(p_ref)).mean()
    print("%s: %d iterations, residual %.3e (oracle %.3e), %.4f %% of the float pressures bit-identical" % (name, iters, resid, err_ref, 100 * same))
    assert same > 0.98
    assert np.array_equal(p == 0, p_ref == 0) or np.count_nonzero((p == 0) != (p_ref == 0)) < 4

    # the pressure update, fed the device's own pressure: bit-exact against the oracle's update of the same inputs
    ctx.apply_pressure(capi.FIELD_P2G, capi.FIELD_NEW, dt)
    f8 = ctx.get_field(capi.FIELD_NEW)
    g8 = oracle.apply_pressure(*g6, s["dims"], s["dx"], mat, p, dt)
    for a, b in zip(f8, g8):
        assert np.array_equal(bits(a), bits(b))
    assert all(np.array_equal(bits(a), bits(b)) for a, b in zip(ctx.get_field(capi.FIELD_P2G), f6))      # source untouched
    # in place gives the same field
    ctx.apply_pressure(capi.FIELD_P2G, capi.FIELD_P2G, dt)
    for a, b in zip(ctx.get_field(capi.FIELD_P2G), g8):
        assert np.array_equal(bits(a), bits(b))
    # and the projected field is divergence free in the fluid cells
    I, J, K = s["dims"]
    u = f8[0].reshape(K, J, I + 1); v = f8[1].reshape(K, J + 1, I); w = f8[2].reshape(K + 1, J, I)
    div = (u[:, :, 1:] - u[:, :, :-1] + v[:, 1:, :] - v[:, :-1, :] + w[1:] - w[:-1]) / s["dx"]
    assert np.abs(div[mat.reshape(K, J, I) == synth.FLUID]).max() < 1e-3


def test_pressure_solve_repeats_bit_for_bit_and_handles_limits(ctx, oracle):
    s = make_case("small32", False)
    dt = 1.0 / 30
    mat, f5 = stage5(ctx, s)
    ctx.apply_body_force(capi.FIELD_P2G, (0, -9.8, 0), dt)
    it1, r1 = ctx.pressure_solve(capi.FIELD_P2G, dt)
    p1 = ctx.get_pressure()
    it2, r2 = ctx.pressure_solve(capi.FIELD_P2G, dt)
    p2 = ctx.get_pressure()
    assert it1 == it2 and r1 == r2 and np.array_equal(bits(p1), bits(p2))
    # iteration limit: the estimate so far is kept, as the reference does
    f6 = ctx.get_field(capi.FIELD_P2G)
    it3, r3 = ctx.pressure_solve(capi.FIELD_P2G, dt, max_iterations=5)
    p3 = ctx.get_pressure()
    p_ref, it_ref, limit, err_ref = oracle.pressure_solve(*f6, s["dims"], s["dx"], mat, dt, max_iterations=5)
    assert limit and it3 == 5 and it_ref == 5
    assert np.abs(p3 - p_ref).max() <= 2e-6 * np.abs(p_ref).max()
    # nothing to solve: zero field -> zero pressure, the reference's early return
    ctx.set_field(capi.FIELD_NEW, *[np.zeros_like(a) for a in f6])
    it4, r4 = ctx.pressure_solve(capi.FIELD_NEW, dt)
    assert it4 == -1 and r4 == 0.0 and not ctx.get_pressure().any()


@pytest.mark.parametrize("name,solids", [("slab24", True), ("odd20", False), ("hello64", False)])
def test_pressure_sweep_variants_bit_identical(ctx, name, solids):
    """option 12: substitution kernels reading global memory / staged in shared memory behind tile flags / staged with
    data-flow (sentinel) synchronisation, waiting per tile or per value -- every float of the result identical, same iteration count and residual."""
    s = make_case(name, solids)
    dt = 1.0 / 30
    stage5(ctx, s)
    ctx.apply_body_force(capi.FIELD_P2G, (0.1, -9.8, 0.2), dt)
    out = []
    for variant in (0, 1, 2, 3, 3):
        ctx.set_option(12, variant)
        it, res = ctx.pressure_solve(capi.FIELD_P2G, dt)
        out.append((it, res, ctx.get_pressure()))
    ctx.set_option(12, 2)
    for other in out[1:]:
        assert out[0][0] == other[0] and out[0][1] == other[1]
        assert np.array_equal(bits(out[0][2]), bits(other[2]))
    assert out[0][0] > 3


def test_pressure_solve_field_host_pointer_operator(ctx, oracle):
    """gfs_pressure_solve_field (the body of the drop-in PressureSolver::solve): double pressures per cell from host arrays."""
    s = make_case("slab24", True)
    dt = 1.0 / 30
    mat, f5 = stage5(ctx, s)
    g6 = oracle.body_force(*f5, s["dims"], mat, (0.0, -9.8, 0.0), dt)
    p_ref, it_ref, limit, err_ref = oracle.pressure_solve(*g6, s["dims"], s["dx"], mat, dt)
    p, it, res = ctx.pressure_solve_field(*g6, s["dims"], s["dx"], mat, dt)
    assert it == it_ref and not limit
    assert np.array_equal(bits(p.astype(np.float32)), bits(p_ref)) or np.abs(p.astype(np.float32) - p_ref).max() <= 2e-6 * np.abs(p_ref).max()
    assert not p[mat != synth.FLUID].any()


def test_pressure_stages_golden_fixture_through_cuda(ctx):
    """The committed outputs of the unmodified reference (tests/golden/pressure.npz) straight against the CUDA path, no
    oracle in between: body forces and the pressure update bit-exact, the float pressure grid to 2e-6 of its maximum."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pressure.npz"))
    dims, dx, dt, density = tuple(int(x) for x in g["dims"]), float(g["dx"]), float(g["dt"]), float(g["density"])
    ctx.domain_init(dims, dx)
    ctx.set_material(g["material"])
    ctx.set_field(capi.FIELD_P2G, g["u5"], g["v5"], g["w5"])
    ctx.apply_body_force(capi.FIELD_P2G, g["force"], dt)
    for a, nm in zip(ctx.get_field(capi.FIELD_P2G), ("u6", "v6", "w6")):
        assert np.array_equal(bits(a), bits(g[nm]))
    iters, resid = ctx.pressure_solve(capi.FIELD_P2G, dt, density)
    p = ctx.get_pressure()
    assert iters > 3 and resid < 1e-6
    assert np.abs(p - g["pressure"]).max() <= 2e-6 * np.abs(g["pressure"]).max()
    assert (bits(p) == bits(g["pressure"])).mean() > 0.98
    ctx.apply_pressure(capi.FIELD_P2G, capi.FIELD_NEW, dt, density)
    for a, nm in zip(ctx.get_field(capi.FIELD_NEW), ("u8", "v8", "w8")):
        assert np.abs(a - g[nm]).max() <= 2e-6 * max(1.0, np.abs(g[nm]).max())


def test_pressure_requires_domain_and_solve(ctx):
    with pytest.raises(capi.GfsError):
        ctx.pressure_solve(capi.FIELD_P2G, 1.0 / 30)
    s = make_case("tiny16", False)
    stage5(ctx, s)
    with pytest.raises(capi.GfsError):
        ctx.apply_pressure(capi.FIELD_P2G, capi.FIELD_NEW, 1.0 / 30)
    with pytest.raises(capi.GfsError):
        ctx.pressure_solve(capi.FIELD_P2G, -1.0)
