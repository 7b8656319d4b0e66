"""Parity of the CUDA path against the oracle AT THE SIZES THE PERFORMANCE IS QUOTED ON (BASELINE.json configs[1..3]).

  * dambreak128 (128^3, 7.7 M particles): every array of one fast substep, trilinear and tricubic, against the oracle;
  * splash256 (256^3, 96 M particles) and river512 (512x256x256, 245 M particles): the sampled comparison SURVEY.md
    §8(d) prescribes -- three 32^3 sub-boxes of the P2G grids and the material (the oracle runs on just the particles
    that can touch them) and 1 M random particles through PIC/FLIP + RK4 -- trilinear and tricubic.

Scale-dependent machinery these exercise and the small cases do not: 32-bit products in the index arithmetic, the
brick count and key range, tensor-map extents and the 264-float row pitch, the tile/dense switch of the splat, counters
beyond 2^24.  Tolerance: the mixed fp32 bound of tests/test_gpu_parity.py; material and cell indices bit-exact.
"""
import numpy as np
import pytest

from gridfluidsim3d_b200 import capi, synth

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def assert_close(a, b, what, rtol=RTOL, scale=None):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = (np.abs(b).max() if b.size else 0.0) if scale is None else scale
    tol = rtol * np.maximum(np.abs(a), np.abs(b)) + rtol * scale
    bad = np.abs(a - b) > tol
    assert not bad.any(), "%s: %d of %d outside tolerance, worst |a-b|=%g at scale %g" % (
        what, bad.sum(), bad.size, np.abs(a - b).max(), scale)


def test_dambreak128_full_parity(oracle):
    """BASELINE.json configs[1]: one fast substep, every array in full, both interpolations."""
    s = synth.make_scene("dambreak128")
    mat = s["material"].copy()
    ref_uvw = oracle.p2g(s["pos"], s["vel"], s["dims"], s["dx"], mat)
    for interp in (capi.TRILINEAR, capi.TRICUBIC):
        p_ref, v_ref, _ = oracle.g2p_advect(s["pos"], s["vel"], s["new"], s["saved"], s["dims"], s["dx"], s["dt"],
                                            mode=interp, material=mat)
        c = capi.Context(0)
        c.domain_init(s["dims"], s["dx"]); c.set_material(s["material"]); c.set_sources([])
        c.set_particles(s["pos"], s["vel"])
        c.set_field(capi.FIELD_NEW, *s["new"]); c.set_field(capi.FIELD_SAVED, *s["saved"])
        c.substep(s["dt"], interp=interp, arith=capi.FAST)
        assert np.array_equal(c.get_material(), mat)
        for a, b, nm in zip(c.get_field(capi.FIELD_P2G), ref_uvw, "uvw"):
            assert_close(a, b, "p2g " + nm)
        o = c.get_particle_order()
        p, v = c.get_particles()
        st = c.stats()
        c.close()
        assert len(np.unique(o)) == len(s["pos"]) and st["collision_overflow"] == 0
        assert_close(v, v_ref[o], "velocity (interp %d)" % interp)
        assert_close(p, p_ref[o], "position (interp %d)" % interp)
        ca, cb = oracle.cell_index(p, s["dx"]), oracle.cell_index(p_ref[o], s["dx"])
        assert (ca != cb).any(1).mean() < 1e-4


def test_dambreak128_pressure_solve_parity(oracle):
    """Stages 6-8 at BASELINE.json configs[1] (128^3 dam break, ~0.9 M fluid cells): the device MICCG(0) must take the
    oracle's iteration count (or hit the reference's 200-iteration limit with it) and produce its float pressure grid."""
    s = synth.make_scene("dambreak128")
    dt = 1.0 / 30
    c = capi.Context(0)
    c.domain_init(s["dims"], s["dx"]); c.set_material(s["material"]); c.set_sources([])
    c.set_particles(s["pos"], s["vel"])
    c.sort_index(); c.p2g(capi.FAST)
    mat = c.get_material()
    f5 = c.get_field(capi.FIELD_P2G)
    force = (0.0, -9.8, 0.0)
    c.apply_body_force(capi.FIELD_P2G, force, dt)
    f6 = c.get_field(capi.FIELD_P2G)
    g6 = oracle.body_force(*f5, s["dims"], mat, force, dt)
    for a, b in zip(f6, g6):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    iters, resid = c.pressure_solve(capi.FIELD_P2G, dt)
    p = c.get_pressure()
    c.apply_pressure(capi.FIELD_P2G, capi.FIELD_NEW, dt)
    f8 = c.get_field(capi.FIELD_NEW)
    c.close()
    p_ref, it_ref, limit, err_ref = oracle.pressure_solve(*g6, s["dims"], s["dx"], mat, dt)
    print("dambreak128 pressure: %d iterations (oracle %d, limit %s), residual %.3e (oracle %.3e), %d fluid cells"
          % (iters, it_ref, limit, resid, err_ref, int((mat == synth.FLUID).sum())))
    scale = np.abs(p_ref).max()
    same = (p.view(np.uint32) == p_ref.view(np.uint32)).mean()
    print("  max |p - p_ref| / max |p_ref| = %.3e, %.2f %% of the float pressures bit-identical" % (np.abs(p - p_ref).max() / scale, 100 * same))
    assert iters == it_ref
    # 99 iterations on 10^6 unknowns amplify the last-place difference of the dot products' summation order: the final
    # residuals agree to a fraction of a percent, the pressures to fp32 rounding
    assert abs(resid - err_ref) <= 0.05 * err_ref
    assert np.abs(p - p_ref).max() <= RTOL * scale
    assert same > 0.5
    g8 = oracle.apply_pressure(*g6, s["dims"], s["dx"], mat, p, dt)
    for a, b in zip(f8, g8):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


# sub-boxes (lower corner, in cells) per workload: at the free surface, deep inside, and in a domain corner (border solids,
# boundary faces); river: also one straddling x = 256..288 where 32-bit linear indices pass 2^24 per plane
BOXES = {
    "splash256": [(100, 168, 60), (200, 40, 190), (0, 0, 0)],
    "river512": [(300, 220, 100), (250, 90, 120), (480, 224, 224)],
}


@pytest.mark.parametrize("name", ["splash256", "river512"])
def test_sampled_parity_at_scale(oracle, name):
    import torch
    dev = torch.device("cuda", 0)
    sc = synth.make_scene_torch(name, dev)
    dims, dx, dt = sc["dims"], sc["dx"], sc["dt"]
    I, J, K = dims
    aos = sc["aos"].cpu().numpy()
    new = [t.cpu().numpy() for t in sc["new"]]
    saved = [t.cpu().numpy() for t in sc["saved"]]
    mat0 = sc["material"]
    del sc
    torch.cuda.empty_cache()
    N = len(aos)
    pos, vel = np.ascontiguousarray(aos[:, :3]), np.ascontiguousarray(aos[:, 3:])
    rng = np.random.default_rng(2026)
    sample = np.sort(rng.choice(N, size=1_000_000, replace=False))
    fdims = synth.face_dims(dims)

    p2g_checked = False
    for interp in (capi.TRILINEAR, capi.TRICUBIC):
        c = capi.Context(0)
        c.domain_init(dims, dx); c.set_material(mat0); c.set_sources([])
        c.set_particles_aos(aos)
        c.set_field(capi.FIELD_NEW, *new); c.set_field(capi.FIELD_SAVED, *saved)
        c.substep(dt, interp=interp, arith=capi.FAST)
        st = c.stats()
        assert st["num_particles"] == N and st["out_of_grid"] == 0 and st["in_solid"] == 0 and st["collision_overflow"] == 0
        order = c.get_particle_order()
        p_gpu, v_gpu = c.get_particles()
        if not p2g_checked:
            mat_gpu = c.get_material()
            uvw_gpu = c.get_field(capi.FIELD_P2G)
        c.close()

        # ---- G2P: 1 M random particles through the oracle, matched through the order tags --------------------------
        inv = np.empty(N, np.int32)
        inv[order] = np.arange(N, dtype=np.int32)
        assert np.array_equal(order[inv[sample]], sample)                  # the tags are a permutation
        p_ref, v_ref, _ = oracle.g2p_advect(pos[sample], vel[sample], new, saved, dims, dx, dt, mode=interp, material=mat0)
        slots = inv[sample]
        assert_close(v_gpu[slots], v_ref, "%s velocity (interp %d)" % (name, interp))
        assert_close(p_gpu[slots], p_ref, "%s position (interp %d)" % (name, interp))
        ca, cb = oracle.cell_index(p_gpu[slots], dx), oracle.cell_index(p_ref, dx)
        assert (ca != cb).any(1).mean() < 1e-4
        del inv, order, p_gpu, v_gpu

        # ---- P2G + classification: three 32^3 sub-boxes (the splat does not depend on the interpolation: once) ----------
        if p2g_checked:
            continue
        p2g_checked = True
        assert st["fluid_cells"] == int((mat_gpu == synth.FLUID).sum())
        for (bi, bj, bk) in BOXES[name]:
            lo = np.array([bi, bj, bk]); hi = lo + 32
            # particles that can influence a face of the box: splat radius 1 cell + the 26-neighbour fill of unset faces
            sel = np.all((pos >= ((lo - 3) * dx).astype(np.float32)) & (pos < ((hi + 3) * dx).astype(np.float32)), axis=1)
            assert sel.sum() > 1000
            m = mat0.copy()
            u, v, w = oracle.p2g(pos[sel], vel[sel], dims, dx, m)
            sl = np.s_[max(bk, 0):min(bk + 32, K), max(bj, 0):min(bj + 32, J), max(bi, 0):min(bi + 32, I)]
            assert np.array_equal(mat_gpu.reshape(K, J, I)[sl], m.reshape(K, J, I)[sl]), "material in box %s" % ((bi, bj, bk),)
            for comp, (got, ref) in enumerate(zip(uvw_gpu, (u, v, w))):
                ni, nj, nk = fdims[comp]
                e = [1 if comp == a else 0 for a in range(3)]
                fs = np.s_[bk:min(bk + 32 + e[2], nk), bj:min(bj + 32 + e[1], nj), bi:min(bi + 32 + e[0], ni)]
                g, r = got.reshape(nk, nj, ni)[fs], ref.reshape(nk, nj, ni)[fs]
                assert np.abs(r).max() > 0
                assert_close(g, r, "%s p2g comp %d in box %s" % (name, comp, (bi, bj, bk)), scale=max(np.abs(r).max(), 1.0))
                assert np.array_equal(g != 0, r != 0)
