"""numpy prototype: boundary solve with different probe sets (cluster masses; top-R single states + cluster remainders).
Usage: python boundary_probe_sets_prototype.py <chromosome 1..23> <plain|cluster|topR> [first solve round]
Needs proto.py / rounds_per_chr.py helpers of the same directory tree (scripts/research/boundary_probe_prototype.py
holds the shared step / setup functions); kept as a record of the experiment described in DESIGN.md section 8."""
import sys, numpy as np
sys.path.insert(0,'/root/repo')
from proto import step, relerr, exact_forward
import rounds_per_chr as RP
from tests import synth
from oracle import titan_oracle as O
c=int(sys.argv[1]); mode=sys.argv[2]; first=int(sys.argv[3]) if len(sys.argv)>3 else 2
full,_=synth.make_data(1000000, Z=3, maxCN=8, n_chr=23, seed=20261020)
p=O.load_default_parameters(8,3,data=full)
S=RP.setup_chr(c, s=tuple(p['cellPrevParams']['s_0']), n=0.45, phi=2.04, var=np.asarray(p['genotypeParams']['var_0']))
N=S['N']; L=1024; W=192; tol=1e-8; Z=S['Z']; U=S['U']
al=exact_forward(S)
starts=np.arange(0,N,L); nc=len(starts); ends=np.minimum(starts+L,N)
true_start=[None]+[al[starts[k]-1] for k in range(1,nc)]
start=[None]*nc; end=[None]*nc; E=[None]*nc; M=[None]*nc
def run_chunk(k,a0):
    a=a0
    for t in range(starts[k],ends[k]): a=step(a,t,S) if t>0 else (S['pi']*S['B'][0])/(S['pi']*S['B'][0]).sum()
    return a
def masks_of(v):
    ms=[]
    if mode=='cluster':
        for z in range(Z):
            m=np.zeros(v.shape,bool); m[z]=True; ms.append(m)
    elif mode.startswith('top'):
        R=int(mode[3:]); flat=np.argsort(-v.ravel())[:R]
        taken=np.zeros(v.shape,bool)
        for i in flat:
            m=np.zeros(v.shape,bool); m.ravel()[i]=True; ms.append(m); taken|=m
        for z in range(Z):
            m=np.zeros(v.shape,bool); m[z]=True; m&=~taken; ms.append(m)
    return ms
def probes(k):
    ms=masks_of(start[k]); basis=np.stack([np.where(m,start[k],0.0) for m in ms])
    Q=basis.copy(); co=S['co']; het=S['het']
    for t in range(starts[k],ends[k]):
        irs=np.where(het, co['irh'][t], co['irn'][t])
        w=Q*irs; Wt=w.sum((-1,-2),keepdims=True); Wz=w.sum(-1,keepdims=True); Wg=w.sum(-2,keepdims=True)
        o=np.where(het, co['cB'][t]*Wt+co['k4'][t]*Wg+0*w, co['k0'][t]*Wt+co['k1'][t]*Wz+co['k2'][t]*Wg+co['k3'][t]*w)
        Q=o*S['B'][t]; Q/=Q.sum()
    return Q, ms
for k in range(nc):
    if k>0:
        tb=max(starts[k]-W,0); a=S['B'][tb].copy(); a/=a.sum()
        for t in range(tb+1,starts[k]): a=step(a,t,S)
        start[k]=a
    end[k]=run_chunk(k,start[k])
    if k>0: E[k],M[k]=probes(k)
for r in range(1,25):
    mism=[k for k in range(1,nc) if relerr(start[k],end[k-1])>tol]
    print('round',r,'mismatch',len(mism), mism[:12])
    if not mism: break
    if r>=first and mode!='plain':
        a=end[0]; new=[None]*nc
        for k in range(1,nc):
            a=a/a.sum(); new[k]=a
            if relerr(a,start[k])<=tol:
                a=end[k]; continue
            x=np.array([(a[m].sum()/start[k][m].sum()) if start[k][m].sum()>0 else 0.0 for m in M[k]])
            a=(x[:,None,None]*E[k]).sum(0)
    else:
        new=[None]+[end[k-1] for k in range(1,nc)]
    for k in range(1,nc):
        if relerr(new[k],start[k])>tol:
            start[k]=new[k]; end[k]=run_chunk(k,start[k]); E[k],M[k]=probes(k)
