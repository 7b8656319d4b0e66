"""Run under torchrun with >= 2 GPUs: the peer-memory transport (gfs_comm_*) against the torch.distributed one.

Every rank steps the same sharded scene twice, once per transport, and the owned P2G layers, the material and the
sorted particle rows must agree bit for bit after every substep.  Prints "PEER_CHECK_OK" on rank 0.
Used by tests/test_gpu_parity.py::test_peer_transport_two_gpus, by bench.py at every --gpus N > 1 (check(), after its
timed regions: the result is the line's verify.peer_vs_nccl) and by hand:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/peer_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gridfluidsim3d_b200 import capi, slabs, synth  # noqa: E402


def check(rank, world, local, dev, name=None):
    """-> (ok on every rank, particles of the last peer run); needs an initialised process group."""
    name = name or os.environ.get("PEER_CHECK_SCENE", "small32")
    s = synth.make_scene(name, seed=7)
    s["new"] = (s["new"][0], s["new"][1], (s["new"][2] + np.float32(0.6)).astype(np.float32))
    I, J, K = s["dims"]
    k0, k1 = slabs.slab_ranges(K, world)[rank]
    kcell = np.floor(s["pos"][:, 2].astype(np.float64) / s["dx"]).astype(np.int64)
    mine = (kcell >= k0) & (kcell < k1)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    results = {}
    for kind in ("nccl", "peer"):
        c = capi.Context(local, stream=stream.cuda_stream)
        c.domain_init(s["dims"], s["dx"]); c.set_material(s["material"]); c.set_sources([])
        c.set_particles(s["pos"][mine], s["vel"][mine])
        c.set_field(capi.FIELD_NEW, *s["new"]); c.set_field(capi.FIELD_SAVED, *s["saved"])
        if os.environ.get("PEER_CHECK_ALLMAX_EARLY"):
            c.set_option(9, int(os.environ["PEER_CHECK_ALLMAX_EARLY"]))
        for interp in (capi.TRILINEAR, capi.TRICUBIC):
            drv = slabs.SlabDriver(slabs.CudaSlabBackend(c, s["dims"], (k0, k1), interp, shared_stream=True), rank, world,
                                   halo=capi.slab_halo_cells(interp, 0.5 * s["dx"], s["dx"]))
            if interp == capi.TRILINEAR:
                big = slabs.PeerTransport.layer_bytes(slabs.SlabDriver(
                    slabs.CudaSlabBackend(c, s["dims"], (k0, k1), capi.TRICUBIC, shared_stream=True), rank, world,
                    halo=capi.slab_halo_cells(capi.TRICUBIC, 0.5 * s["dx"], s["dx"])))
                tr = slabs.DistTransport() if kind == "nccl" else slabs.PeerTransport(drv, particle_cap=len(s["pos"]), layer_bytes=big)
            out = []
            for step in range(3):
                if os.environ.get("PEER_CHECK_VERBOSE"):
                    print("rank %d %s interp %d step %d" % (rank, kind, interp, step), file=sys.stderr, flush=True)
                slabs.substep(drv, tr, 1.5 * s["dt"], pressure_solve_between=(step == 1))
                torch.cuda.synchronize()
                p, v = c.get_particles()
                rows = np.ascontiguousarray(np.concatenate([p, v], 1)).view([("f%d" % i, "f4") for i in range(6)]).reshape(-1)
                rows = np.sort(rows, order=["f%d" % i for i in range(6)])
                f = [a.view(np.uint32).copy() for a in c.get_field(capi.FIELD_P2G)]
                out.append((rows, f, c.get_material().copy()))
            results[(kind, interp)] = out
        dist.barrier()
        c.close()
    ok = True
    for interp in (capi.TRILINEAR, capi.TRICUBIC):
        for (ra, fa, ma), (rb, fb, mb) in zip(results[("nccl", interp)], results[("peer", interp)]):
            ok = ok and ra.shape == rb.shape and np.array_equal(ra, rb) and np.array_equal(ma[k0 * I * J:k1 * I * J], mb[k0 * I * J:k1 * I * J])
            for a, b, (ni, nj, nk) in zip(fa, fb, synth.face_dims(s["dims"])):
                ok = ok and np.array_equal(a.reshape(nk, nj, ni)[k0:k1], b.reshape(nk, nj, ni)[k0:k1])
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    moved = torch.tensor([len(results[("peer", capi.TRILINEAR)][-1][0])], device=dev)
    dist.all_reduce(moved)
    return int(t.item()) == 1, int(moved.item())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok, n = check(rank, world, local, dev)
    if rank == 0:
        print("PEER_CHECK_OK" if ok else "PEER_CHECK_FAILED", "particles", n, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
