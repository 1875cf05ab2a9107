"""Companion of boundary_probe_prototype.py: per-chunk accuracy of the probe-solved starts."""
import sys, numpy as np
from boundary_probe_prototype import *
S=setup(N=90000)
N=S['N']; L=512; W=96
al=exact_forward(S)
starts=np.arange(0,N,L); nc=len(starts); ends=np.minimum(starts+L,N)
true_start=[None]+[al[starts[c]-1] for c in range(1,nc)]
start=[None]*nc; end=[None]*nc
def run_chunk(c,a0):
    a=a0
    for t in range(starts[c],ends[c]): a=step(a,t,S) if t>0 else (S['pi']*S['B'][0])/(S['pi']*S['B'][0]).sum()
    return a
for c in range(nc):
    if c>0:
        tb=max(starts[c]-W,0); a=S['B'][tb].copy(); a/=a.sum()
        for t in range(tb+1,starts[c]): a=step(a,t,S)
        start[c]=a
    end[c]=run_chunk(c,start[c])
def probe_E(c, basis):
    Q=np.stack(basis)
    co=S['co']; het=S['het']
    for t in range(starts[c],ends[c]):
        irs=np.where(het, co['irh'][t], co['irn'][t])
        w=Q*irs; Wt=w.sum((-1,-2),keepdims=True); Wz=w.sum(-1,keepdims=True); Wg=w.sum(-2,keepdims=True)
        o=np.where(het, co['cB'][t]*Wt+co['k4'][t]*Wg+0*w, co['k0'][t]*Wt+co['k1'][t]*Wz+co['k2'][t]*Wg+co['k3'][t]*w)
        Q=o*S['B'][t]; Q/=Q.sum()
    return Q
for r in range(1,4):
    new=[None]+[end[c-1] for c in range(1,nc)]
    for c in range(1,nc):
        if relerr(new[c],start[c])>1e-8:
            start[c]=new[c]; end[c]=run_chunk(c,start[c])
bad=[c for c in range(1,nc) if relerr(start[c],true_start[c])>1e-8]
print('after 3 rounds bad',bad)
# solve variants
for mode in ('mass','ls'):
    a=end[0]; errs={}
    for c in range(1,nc):
        a=a/a.sum()
        errs[c]=relerr(a,true_start[c])
        if c==nc-1: break
        basis=[]
        for z in range(S['Z']):
            q=np.zeros_like(start[c]); q[z]=start[c][z]; basis.append(q)
        if mode=='mass':
            x=a.sum(-1)/start[c].sum(-1)
        else:
            # relative LS: minimise sum ((a - sum x_z q_z)/a)^2
            Amat=np.stack([ (q/np.maximum(a,1e-300)).ravel() for q in basis],1)
            x=np.linalg.lstsq(Amat, np.ones(Amat.shape[0]), rcond=None)[0]
        E=probe_E(c,basis)
        a=(x[:,None,None]*E).sum(0)
    worst=sorted(errs.items(), key=lambda kv:-kv[1])[:6]
    print(mode,'worst solved-start errors',[(c,'%.2e'%e) for c,e in worst])
    c=worst[0][0]
    # detail
    a=end[0]
    for cc in range(1,c+1):
        a=a/a.sum()
        if cc==c: break
        basis=[]
        for z in range(S['Z']):
            q=np.zeros_like(start[cc]); q[z]=start[cc][z]; basis.append(q)
        x=a.sum(-1)/start[cc].sum(-1)
        E=probe_E(cc,basis); a=(x[:,None,None]*E).sum(0)
    b=true_start[c]; rel=np.abs(a-b)/np.maximum(a,b)
    for i in np.argsort(-rel.ravel())[:6]:
        z,g=np.unravel_index(i,b.shape); print('   chunk',c,'(z=%d,g=%d) true %.3e solved %.3e cur-est %.3e'%(z,g,b[z,g],a[z,g],start[c][z,g]))
