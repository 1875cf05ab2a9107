"""Reduction look-ahead for the forward recursion: numerical check of the algebra (numpy, CPU).

The cluster sums of step t+1 follow from data that exists before the combine of step t:
    Wz_{t+1}[z] = sc_t (sum_g v_t E_t + W_t sum_g v_t cW + Wz_t[z] sum_g v_t cZ),   v_t = B_t irs_{t+1},
so the butterfly that forms them need not gate alpha_t.  This script runs the plain structured step and the
look-ahead form side by side on synthetic emissions (agreement ~2e-15 relative over 3000 steps, identical
power-of-two rescale exponents).  The CUDA kernel that implements it (fwd_chunk_la_kernel, see
reduction_lookahead_kernel.patch) is bit-compatible with the test-suite but was NOT adopted: measured on
B200 the single recursion warp still issues its 173 instructions per step in order at ~3.6 cycles each
(583 cycles per step against 650 for the plain consumer), i.e. the butterfly is off the dependence chain but
not off the issue stream; it needs a hand-scheduled two-step software pipeline to pay off.
"""
import numpy as np
rng=np.random.default_rng(1)
U,Z,T=25,3,3000
het=np.zeros(U,bool); het[3]=True
h=int(het.sum()); K=U*Z
def coeffs(rG,rZ):
    oG=(1-rG)/(K-1); nz=1-rZ; q=rZ+(Z-1)*nz
    rs_non=rG*q+oG*(h*Z*rZ+(U-1-h)*q); rs_het=rG*Z*rZ+oG*((h-1)*Z*rZ+(U-h)*q)
    return dict(k0=oG*nz,k1=oG*(rZ-nz),k2=(rG-oG)*nz,k3=(rG-oG)*(rZ-nz),cB=oG*rZ,k4=(rG-oG)*rZ,irn=1/rs_non,irh=1/rs_het)
gaps=rng.integers(1,20000,size=T+2).astype(float)
C=[coeffs(1-0.5*(1-np.exp(-(g+1)/(2*1e15))),1-0.5*(1-np.exp(-(g+1)/(2*1e15)))) for g in gaps]
# emissions: peaked, shifting
B=np.exp(-rng.gamma(2.0,8.0,size=(T+2,U,Z))); B/=B.max(axis=(1,2),keepdims=True)
def sel(c):
    irs=np.where(het,c['irh'],c['irn'])[:,None]
    cW=np.where(het,c['cB'],c['k0'])[:,None]; cG=np.where(het,c['k4'],c['k2'])[:,None]
    cZ=np.where(het,0.0,c['k1'])[:,None]; cw=np.where(het,0.0,c['k3'])[:,None]
    return irs,cW,cG,cZ,cw
def pow2(W):
    m,e=np.frexp(W); return np.ldexp(1.0,-(e-1)), e-1
a0=rng.random((U,Z)); a0/=a0.sum()
# direct
a=a0.copy(); E1=0; A1=[]
for t in range(1,T+1):
    irs,cW,cG,cZ,cw=sel(C[t])
    w=a*irs; Wz=w.sum(0,keepdims=True); W=Wz.sum(); Wg=w.sum(1,keepdims=True)
    sc,e=pow2(W); E1+=e
    a=(cW*W+cG*Wg+cZ*Wz+cw*w)*(B[t]*sc); A1.append(a.copy())
# look-ahead
a=a0.copy(); E2=0; A2=[]
irs,cW,cG,cZ,cw=sel(C[1])
w=a*irs; Wz_cur=w.sum(0,keepdims=True); W_cur=Wz_cur.sum()   # bootstrap: direct reduction for the first step
for t in range(1,T+1):
    irs,cW,cG,cZ,cw=sel(C[t]); irs_n=sel(C[t+1])[0]
    w=a*irs; Wg=w.sum(1,keepdims=True)
    E=cG*Wg+cw*w
    v=B[t]*irs_n                                    # producer
    PW=(v*cW).sum(0,keepdims=True); PZ=(v*cZ).sum(0,keepdims=True)   # producer
    S1=(v*E).sum(0,keepdims=True)                   # the butterfly on the critical path
    sc,e=pow2(W_cur); E2+=e
    a=(cW*W_cur+cZ*Wz_cur+E)*(B[t]*sc); A2.append(a.copy())
    Wz_cur=sc*(S1+W_cur*PW+Wz_cur*PZ); W_cur=Wz_cur.sum()
rel=max(np.max(np.abs(x-y)/np.maximum(np.maximum(x,y),1e-300)) for x,y in zip(A1,A2))
print('exponent sums',E1,E2,'max rel diff',rel)
# difference allowing power-of-two mismatch: compare normalised
reln=max(np.max(np.abs(x/x.sum()-y/y.sum())/np.maximum(x/x.sum(),1e-300)) for x,y in zip(A1,A2))
print('max rel diff normalised',reln)
