#!/usr/bin/env python
"""Kernel A/B harness: the splash256 (or --workload) substep with each requested variant of a tuning option, per-kernel
CUDA-event times from the library's own profiling, and the state hashes (which must not depend on the variant).

    python profiles/kbench.py --option 0 --values 1,2,3 [--steps 6] [--workload splash256] [--interp trilinear]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import torch
    from gridfluidsim3d_b200 import capi, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--option", type=int, default=0)
    ap.add_argument("--values", default="1,2,3")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--workload", default="splash256")
    ap.add_argument("--interp", default="trilinear")
    ap.add_argument("--only", default="", help="substring: print only kernels containing it")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    sc = synth.make_scene_torch(args.workload, dev)
    aos = sc["aos"].cpu().numpy()
    new = [t.cpu().numpy() for t in sc["new"]]
    saved = [t.cpu().numpy() for t in sc["saved"]]
    interp = capi.TRILINEAR if args.interp == "trilinear" else capi.TRICUBIC
    out = {}
    for v in [int(x) for x in args.values.split(",")]:
        c = capi.Context(0)
        c.domain_init(sc["dims"], sc["dx"]); c.set_material(sc["material"]); c.set_sources([])
        c.set_option(args.option, v)
        c.set_particles_aos(aos)
        c.set_field(capi.FIELD_NEW, *new); c.set_field(capi.FIELD_SAVED, *saved)
        for _ in range(3):
            c.substep(sc["dt"], interp=interp, arith=capi.FAST)
        c.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # graph-replayed steps on the context's own stream: time with the host clock around a sync
        import time
        t0 = time.perf_counter()
        for _ in range(args.steps):
            c.substep(sc["dt"], interp=interp, arith=capi.FAST)
        c.sync()
        wall = (time.perf_counter() - t0) / args.steps * 1e3
        c.profile_enable(True); c.profile_read(reset=True)
        for _ in range(args.steps):
            c.substep(sc["dt"], interp=interp, arith=capi.FAST)
        prof = c.profile_read(reset=True)
        c.profile_enable(False)
        h = c.state_hash()
        kern = {k: round(t / max(1, n), 4) for k, (t, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]) if args.only in k}
        out[v] = dict(ms_per_step=round(wall, 4), kernels=kern, hashes=["%016x" % x for x in h])
        print("option %d = %d: %.3f ms/step  %s" % (args.option, v, wall, json.dumps(kern)), flush=True)
        c.close()
    hs = {tuple(o["hashes"]) for o in out.values()}
    print("hashes identical across variants:", len(hs) == 1, flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
