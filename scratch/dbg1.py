import sys; sys.path.insert(0,'.')
import numpy as np
from gridfluidsim3d_b200 import capi, synth
from oracle.pyoracle import Oracle
o=Oracle(); ctx=capi.Context(0)
s=synth.make_scene("small32")
ctx.domain_init(s["dims"],s["dx"]); ctx.set_material(s["material"]); ctx.set_sources([]); ctx.set_particles(s["pos"],s["vel"])
ctx.set_field(0,*s["new"]); ctx.set_field(1,*s["saved"])
mat=s["material"].copy(); pos,vel=s["pos"].copy(),s["vel"].copy()
I,J,K=s["dims"]; dx=s["dx"]
for step in range(3):
    u,v,w=o.p2g(pos,vel,s["dims"],dx,mat)
    ctx.substep(s["dt"],interp=0,arith=0)
    gu,gv,gw=ctx.get_field(2)
    d=np.abs(gv.astype(np.float64)-v)
    idx=int(d.argmax()); print(step,'max diff',d.max(),'at',idx)
    if d.max()>1e-5:
        ni,nj=I,J+1
        i,j,k=idx%ni,(idx//ni)%nj,idx//(ni*nj)
        print('node',i,j,k,'oracle',v[idx],'gpu',gv[idx])
        off=np.array([0.5*dx,0,0.5*dx],np.float32)
        f,wt=o.splat(pos,vel[:,1].copy(),dx,off,dx,(I,J+1,K))
        print('oracle wt',wt[idx],'num',f[idx], 'ratio', f[idx]/wt[idx] if wt[idx]>0 else None)
        # contributors
        q=pos-off; nodep=np.array([i,j,k])*dx
        d2=((q-nodep)**2).sum(1); m=d2<dx*dx*1.001
        print('contributors',m.sum(), d2[m]/(dx*dx), vel[m,1])
        m3=mat.reshape(K,J,I); print('mat cells', m3[k,j-1,i] if j>0 else None, m3[k,j,i] if j<J else None)
    oo=ctx.get_particle_order(); p,vv=ctx.get_particles(); pos[oo],vel[oo]=p,vv
