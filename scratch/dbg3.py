import sys; sys.path.insert(0,'.')
import numpy as np
from gridfluidsim3d_b200 import capi, synth
from oracle.pyoracle import Oracle
o=Oracle(); ctx=capi.Context(0)
s=synth.make_scene("tiny16")
rng = np.random.default_rng(11)
blob = (np.array([5.25, 5.25, 5.25]) + rng.uniform(-0.2, 0.2, size=(700, 3))).astype(np.float32)
cases={
 'blob+randn': (np.concatenate([s["pos"], blob]), np.concatenate([s["vel"], rng.standard_normal((700, 3)).astype(np.float32)])),
 'blob+smooth': (np.concatenate([s["pos"], blob]), np.concatenate([s["vel"], synth.particle_velocities(blob,s["dims"],s["dx"])])),
 'noblob+scaled': (s["pos"], (s["vel"]*5).astype(np.float32)),
 'blob40+smooth': (np.concatenate([s["pos"], blob[:40]]), np.concatenate([s["vel"], synth.particle_velocities(blob[:40],s["dims"],s["dx"])])),
}
for name,(pos,vel) in cases.items():
    mat = s["material"].copy()
    ref = o.p2g(pos, vel, s["dims"], s["dx"], mat)
    ctx.domain_init(s["dims"],s["dx"]); ctx.set_material(s["material"]); ctx.set_sources([]); ctx.set_particles(pos,vel)
    ctx.sort_unstable(); ctx.p2g(0)
    out=ctx.get_field(2)
    print(name, [float(np.abs(a-b).max()) for a,b in zip(out,ref)], ctx.stats())
