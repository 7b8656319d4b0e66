import sys; sys.path.insert(0,'.')
import numpy as np
from gridfluidsim3d_b200 import capi, synth
from oracle.pyoracle import Oracle
o=Oracle()
import os
if os.environ.get('GFS_LIB'): capi.LIB_PATH=os.environ['GFS_LIB']
for name in ("tiny16","small32","slab24","odd20"):
  for variant in (1,0):
    ctx=capi.Context(0)
    s=synth.make_scene(name)
    pos,vel=s["pos"],s["vel"]
    mat = s["material"].copy()
    ref = o.p2g(pos, vel, s["dims"], s["dx"], mat)
    ctx.domain_init(s["dims"],s["dx"]); ctx.set_material(s["material"]); ctx.set_sources([]); ctx.set_particles(pos,vel)
    ctx.set_option(0,variant)
    ctx.sort_unstable(); ctx.p2g(0)
    out=ctx.get_field(2)
    print(name, variant, [float(np.abs(a-b).max()) for a,b in zip(out,ref)], [int(np.count_nonzero(a)) for a in out])
    ctx.close()
