"""Per-tile timestamps of the pressure substitution sweeps (gfs_set_option 13): how long a tile stages, waits, steps, and how
far its completion lags the latest of its predecessors -- the length of the wavefront's critical path.

    python profiles/press_trace.py [splash256|hello64|...]
"""
import numpy as np, sys, ctypes, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from gridfluidsim3d_b200 import capi, synth
rt = ctypes.CDLL("libcudart.so")
def dev_read(ctx, which, n, dt=np.int64):
    err = ctypes.c_int()
    ctx.lib.gfs_device_ptr.restype = ctypes.c_void_p
    ptr = ctx.lib.gfs_device_ptr(ctx.h, which, ctypes.byref(err))
    out = np.empty(n, dt)
    rt.cudaDeviceSynchronize()
    assert rt.cudaMemcpy(ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(ptr), ctypes.c_size_t(out.nbytes), 2) == 0
    return out
name = sys.argv[1] if len(sys.argv) > 1 else "splash256"
dev = torch.device("cuda", 0)
sc = synth.make_scene_torch(name, dev)
aos = sc["aos"].cpu().numpy()
dims, dx = sc["dims"], sc["dx"]
ctx = capi.Context(0)
ctx.domain_init(dims, dx); ctx.set_material(sc["material"]); ctx.set_sources([])
ctx.set_particles_aos(aos)
ctx.sort_index(); ctx.p2g(capi.FAST)
dt = 1.0 / 30
ctx.apply_body_force(capi.FIELD_P2G, (0, -9.8, 0), dt)
ctx.set_option(13, 1)
it, r = ctx.pressure_solve(capi.FIELD_P2G, dt, max_iterations=12)
I, J, K = dims
ntx, nty, ntz = -(-I // 16), -(-J // 8), -(-K // 4)
nt = ntx * nty * ntz
tr = dev_read(ctx, 46, nt * 8).reshape(2, nt, 4)
for rev in (0, 1):
    t = tr[rev]
    done = t[:, 3] > 0
    t0 = t[done, 0].min()
    T = (t - t0) / 1e3          # us
    print("sweep", "backward" if rev else "forward", "tiles traced", done.sum(), "of", nt, "span %.1f us" % (T[done, 3].max()))
    print("  per tile: static %.2f us, wait %.2f us (median) / %.2f (mean), steps %.2f us median, %.2f p90" % (
        np.median(T[done, 1] - T[done, 0]), np.median(T[done, 2] - T[done, 1]), np.mean(T[done, 2] - T[done, 1]),
        np.median(T[done, 3] - T[done, 2]), np.percentile(T[done, 3] - T[done, 2], 90)))
    # critical path: completion time vs the latest predecessor's completion
    end = np.where(done, T[:, 3], np.nan).reshape(ntz, nty, ntx)
    lag = []
    d = 1 if rev else -1
    for tz in range(ntz):
        for ty in range(nty):
            for tx in range(ntx):
                e = end[tz, ty, tx]
                if np.isnan(e): continue
                preds = []
                for (a, b, c) in ((tz + d, ty, tx), (tz, ty + d, tx), (tz, ty, tx + d)):
                    if 0 <= a < ntz and 0 <= b < nty and 0 <= c < ntx and not np.isnan(end[a, b, c]): preds.append(end[a, b, c])
                if preds: lag.append(e - max(preds))
    lag = np.array(lag)
    print("  completion lag behind the latest predecessor: median %.2f us, mean %.2f, p10 %.2f, p90 %.2f" % (np.median(lag), lag.mean(), np.percentile(lag, 10), np.percentile(lag, 90)))
