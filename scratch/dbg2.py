import sys; sys.path.insert(0,'.')
import numpy as np
from gridfluidsim3d_b200 import capi, synth
from oracle.pyoracle import Oracle
o=Oracle(); ctx=capi.Context(0)
s=synth.make_scene("tiny16")
rng = np.random.default_rng(11)
blob = (np.array([5.25, 5.25, 5.25]) + rng.uniform(-0.2, 0.2, size=(700, 3))).astype(np.float32)
pos = np.concatenate([s["pos"], blob]); vel = np.concatenate([s["vel"], rng.standard_normal((700, 3)).astype(np.float32)])
mat = s["material"].copy()
ref = o.p2g(pos, vel, s["dims"], s["dx"], mat)
for variant in (1,0):
    ctx.domain_init(s["dims"],s["dx"]); ctx.set_material(s["material"]); ctx.set_sources([]); ctx.set_particles(pos,vel)
    ctx.set_option(0,variant)
    ctx.sort_unstable(); ctx.p2g(0)
    out=ctx.get_field(2)
    for a,b,nm,(ni,nj,nk) in zip(out,ref,"uvw",synth.face_dims(s["dims"])):
        d=np.abs(a-b); bad=np.nonzero(d>1e-4)[0]
        print(variant,nm,len(bad), 'max',d.max())
        if len(bad):
            i,j,k=bad%ni,(bad//ni)%nj,bad//(ni*nj)
            print('  i range',i.min(),i.max(),'j',j.min(),j.max(),'k',k.min(),k.max(), 'gpu vals',a[bad[:5]],'ref',b[bad[:5]])
