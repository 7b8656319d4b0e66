"""The z-slab exchange logic of gridfluidsim3d_b200.slabs on CPU: world_size-2 (and 3) process groups over gloo,
compute by the oracle through tests/slab_numpy_backend.py, checked against the unsharded oracle."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gridfluidsim3d_b200 import capi, slabs, synth


def test_slab_arithmetic_matches_the_c_abi():
    for K, world in ((256, 8), (28, 3), (7, 7), (100, 6)):
        ranges = slabs.slab_ranges(K, world)
        assert ranges[0][0] == 0 and ranges[-1][1] == K
        for r, (k0, k1) in enumerate(ranges):
            assert capi.slab_range(K, world, r) == (k0, k1)
            for k in range(k0, k1):
                assert capi.slab_owner(K, world, k) == r
    assert capi.slab_halo_cells(capi.TRILINEAR, 0.5 * 0.125, 0.125) == 2
    assert capi.slab_halo_cells(capi.TRICUBIC, 1.3 * 0.125, 0.125) == 4
    with pytest.raises(capi.GfsError):
        capi.slab_range(10, 0, 0)


def drift_scene(name):
    """the synthetic scene with a uniform +z drift added to the NEW field, so particles cross every slab cut"""
    s = synth.make_scene(name)
    s["new"] = (s["new"][0], s["new"][1], (s["new"][2] + np.float32(0.6)).astype(np.float32))
    return s


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, steps, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.pyoracle import Oracle
    from tests.slab_numpy_backend import NumpySlabBackend
    orc = Oracle()
    s = drift_scene(name)
    owned = slabs.slab_ranges(s["dims"][2], world)[rank]
    b = NumpySlabBackend(orc, s, owned, mode=0)
    drv = slabs.SlabDriver(b, rank, world, halo=2)
    tr = slabs.DistTransport()
    moved = 0
    for _ in range(steps):
        sent, got = slabs.substep(drv, tr, s["dt"] * 1.5, pressure_solve_between=(_ == 1))
        moved += sent
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), pos=b.pos, vel=b.vel, material=b.material,
             u=b.p2g[0], v=b.p2g[1], w=b.p2g[2], owned=np.array(owned), moved=moved, bytes=tr.bytes_sent)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,name", [(2, "tiny16"), (3, "slab24")])
def test_sharded_substeps_equal_unsharded(oracle, world, name):
    steps = 3
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), name, steps, d), nprocs=world, join=True)
        parts = [np.load(os.path.join(d, "rank%d.npz" % r)) for r in range(world)]
    s = drift_scene(name)
    I, J, K = s["dims"]
    pos, vel, mat = s["pos"].copy(), s["vel"].copy(), s["material"].copy()
    for _ in range(steps):
        u, v, w = oracle.p2g(pos, vel, s["dims"], s["dx"], mat)
        pos, vel, _ = oracle.g2p_advect(pos, vel, s["new"], s["saved"], s["dims"], s["dx"], s["dt"] * 1.5, mode=0, material=mat)

    # particles: every rank holds exactly the particles of its slab, and the union is the unsharded set, bit for bit
    def rows(p, v):
        a = np.ascontiguousarray(np.concatenate([p, v], 1))
        return np.sort(a.view([("f%d" % i, "f4") for i in range(6)]).reshape(-1), order=["f%d" % i for i in range(6)])
    allp = np.concatenate([p["pos"] for p in parts]); allv = np.concatenate([p["vel"] for p in parts])
    assert np.array_equal(rows(allp, allv), rows(pos, vel))
    assert sum(int(p["moved"]) for p in parts) > 0                     # migration was exercised
    for p in parts:
        k = oracle.cell_index(p["pos"], s["dx"])[:, 2]
        assert ((k >= p["owned"][0]) & (k < p["owned"][1])).all()
    # grids: each rank's owned layers of the last P2G (material exact, u/v/w to fp32 tolerance: partial sums are
    # merged in a different order than a single pass would add them)
    m3 = mat.reshape(K, J, I)
    for p in parts:
        k0, k1 = p["owned"]
        assert np.array_equal(p["material"].reshape(K, J, I)[k0:k1], m3[k0:k1])
        for got, ref, (ni, nj, nk) in zip((p["u"], p["v"], p["w"]), (u, v, w), synth.face_dims(s["dims"])):
            a, b = got.reshape(nk, nj, ni)[k0:k1], ref.reshape(nk, nj, ni)[k0:k1]
            assert np.abs(a - b).max() <= 2e-5 * np.abs(ref).max()
            assert np.array_equal(a != 0, b != 0)


def test_weighted_slab_ranges_balance_particles():
    from gridfluidsim3d_b200 import slabs
    counts = np.array([0, 10, 10, 10, 10, 40, 40, 40, 40, 10, 10, 10, 10, 10, 10, 0] * 2)
    for world in (2, 3, 4):
        r = slabs.slab_ranges_weighted(counts, world, min_layers=2)
        assert r[0][0] == 0 and r[-1][1] == len(counts) and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(b - a >= 2 for a, b in r)
        load = [counts[a:b].sum() for a, b in r]
        uni = [counts[a:b].sum() for a, b in slabs.slab_ranges(len(counts), world)]
        assert max(load) <= max(uni)
