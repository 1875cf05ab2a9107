"""Research prototype (numpy, CPU): how many verification rounds the chunked forward / backward scan needs
under different boundary-acceptance criteria, and what error each criterion leaves in the log-likelihood
and the posteriors.  Vectorised over chunks (all chunks of a round advance in lockstep).

Criteria compared at every chunk boundary (used start vector s vs the predecessor's actual end vector e,
both normalised to sum 1):
  rel   : max_j |s_j - e_j| / max(s_j, e_j) <= tol                       (round-1 product: 1e-8)
  post  : sum_j gamma_j |s_j / e_j - 1| <= tol, gamma ~ e (.) (other direction's vector at the boundary)
          -- the first-order change of the log-likelihood and (up to 2x) of every downstream posterior.

Not part of the product or the tests; uses oracle/ only to build emissions.
usage: python scripts/research/round_criteria_prototype.py [N] [L] [W]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import titan_oracle as O  # noqa: E402
from tests import synth  # noqa: E402


def setup(N=60000, Z=3, s=(0.001, 0.2, 0.3), n=0.5, phi=2.0, var=None, chrom=1, txn=1e15):
    full, _ = synth.make_data(1000000, Z=Z, maxCN=8, n_chr=23, seed=20261020)
    sel = full["chr"] == chrom
    data = {k: v[sel][:N] for k, v in full.items()}
    N = data["posn"].size
    p = O.load_default_parameters(8, Z, data=data)
    g = p["genotypeParams"]
    U = len(g["rt"])
    K = U * Z
    muR, muC = O.clonal_two_component_mixture(g["rt"], g["rn"], n, np.array(s), g["ct"], phi)
    v = g["var_0"] if var is None else np.full(U, var)
    py = O.compute_py_c(data["ref"], data["tumDepth"], data["logR"], muR, muC, v)
    py = np.ascontiguousarray(py.T)
    m = py.max(1, keepdims=True)
    B = np.exp(py - m).reshape(N, Z, U)
    d = np.diff(data["posn"]) + 1.0
    rG = np.ones(N)
    rZ = np.ones(N)
    rG[1:] = 1 - 0.5 * (1 - np.exp(-d / (2 * txn)))
    rZ[1:] = 1 - 0.5 * (1 - np.exp(-d / (2 * txn)))
    het = g["ZS"] == -1
    h = het.sum()
    oG = (1 - rG) / (K - 1)
    nz = 1 - rZ
    q = rZ + (Z - 1) * nz
    rs_non = rG * q + oG * (h * Z * rZ + (U - 1 - h) * q)
    rs_het = rG * Z * rZ + oG * ((h - 1) * Z * rZ + (U - h) * q)
    co = dict(k0=oG * nz, k1=oG * (rZ - nz), k2=(rG - oG) * nz, k3=(rG - oG) * (rZ - nz), cB=oG * rZ, k4=(rG - oG) * rZ,
              irn=1 / rs_non, irh=1 / rs_het)
    pi = np.outer(p["cellPrevParams"]["piZ_0"], g["piG_0"])
    return dict(N=N, Z=Z, U=U, K=K, B=B, m=m[:, 0], co=co, het=het, pi=pi, data=data)


def fstep(a, t, S):
    """a: [n, Z, U]; t: [n] ints -> unnormalised alpha_t (the caller normalises)"""
    co, het = S["co"], S["het"]
    irs = np.where(het[None, :], co["irh"][t][:, None], co["irn"][t][:, None])
    w = a * irs[:, None, :]
    W = w.sum((1, 2), keepdims=True)
    Wz = w.sum(2, keepdims=True)
    Wg = w.sum(1, keepdims=True)
    ex = lambda k: co[k][t][:, None, None]
    o_n = ex("k0") * W + ex("k1") * Wz + ex("k2") * Wg + ex("k3") * w
    o_h = ex("cB") * W + ex("k4") * Wg + 0 * w
    return np.where(het[None, None, :], o_h, o_n) * S["B"][t]


def bstep(b, t1, S):
    """b = beta~_{t1} [n, Z, U], t1: [n] -> unnormalised beta~_{t1-1} (transition t1, emission B_{t1})"""
    co, het = S["co"], S["het"]
    bb = b * S["B"][t1]
    bn = np.where(het[None, None, :], 0.0, bb)
    bh = np.where(het[None, None, :], bb, 0.0)
    Wn = bn.sum((1, 2), keepdims=True)
    Wz = bn.sum(2, keepdims=True)
    H = bh.sum((1, 2), keepdims=True)
    loc = bb.sum(1, keepdims=True)      # per genotype: Wg (non-het lanes) or Hg (het lanes)
    ex = lambda k: co[k][t1][:, None, None]
    cG = np.where(het[None, None, :], ex("k4"), ex("k2"))
    cw = np.where(het[None, None, :], 0.0, ex("k3"))
    o = ex("k0") * Wn + ex("cB") * H + cG * loc + ex("k1") * Wz + cw * bb
    irs = np.where(het[None, :], co["irh"][t1][:, None], co["irn"][t1][:, None])
    return o * irs[:, None, :]


def norm(a):
    s = a.sum((1, 2), keepdims=True)
    return a / s, s[:, 0, 0]


def exact(S):
    N = S["N"]
    al = np.empty((N, S["Z"], S["U"]))
    be = np.empty_like(al)
    ll = 0.0
    a = (S["pi"] * S["B"][0])[None]
    a, c = norm(a)
    ll += np.log(c[0]) + S["m"][0]
    al[0] = a[0]
    for t in range(1, N):
        a, c = norm(fstep(a, np.array([t]), S))
        ll += np.log(c[0]) + S["m"][t]
        al[t] = a[0]
    b = np.full((1, S["Z"], S["U"]), 1.0 / S["K"])
    be[N - 1] = b[0]
    for t in range(N - 1, 0, -1):
        b, _ = norm(bstep(b, np.array([t]), S))
        be[t - 1] = b[0]
    g = al * be
    g /= g.sum((1, 2), keepdims=True)
    return al, be, g, ll


def run_chunks_fwd(S, starts, ends, a0, idx, al_out, ll_out):
    """run chunks idx from start vectors a0 (alpha^_{t0-1}, sum 1); store alpha rows and ll pieces; return end vectors"""
    a = a0.copy()
    L = (ends[idx] - starts[idx]).max()
    ll = np.zeros(len(idx))
    endv = np.empty_like(a)
    for s in range(L):
        t = starts[idx] + s
        live = t < ends[idx]
        tt = np.minimum(t, ends[idx] - 1)
        an, c = norm(fstep(a, tt, S))
        a = np.where(live[:, None, None], an, a)
        ll += np.where(live, np.log(c) + S["m"][tt], 0.0)
        al_out[tt[live]] = a[live]
    ll_out[idx] = ll
    return a


def run_chunks_bwd(S, starts, ends, b0, idx, be_out):
    """b0 = beta~ at SNP ends[idx] (first SNP of the next chunk); fills beta~ for t in [t0, t1); returns beta~_{t0}"""
    b = b0.copy()
    L = (ends[idx] - starts[idx]).max()
    for s in range(L):
        t = ends[idx] - 1 - s          # produce beta~_t from beta~_{t+1}
        live = t >= starts[idx]
        tt = np.maximum(t, starts[idx])
        bn, _ = norm(bstep(b, tt + 1, S))
        b = np.where(live[:, None, None], bn, b)
        be_out[tt[live]] = b[live]
    return b


def simulate(S, ex, L=512, W=192, crit="rel", tol=1e-8, verbose=False):
    N, Z, U = S["N"], S["Z"], S["U"]
    al_x, be_x, g_x, ll_x = ex
    starts = np.arange(0, N, L)
    nc = len(starts)
    ends = np.minimum(starts + L, N)
    al = np.zeros((N, Z, U))
    be = np.zeros((N, Z, U))
    llp = np.zeros(nc)
    # ---- round 0: warm-up from flat ----
    sF = np.zeros((nc, Z, U))
    sB = np.zeros((nc, Z, U))
    idx = np.arange(1, nc)
    tb = np.maximum(starts[idx] - W, 0)
    a = S["B"][tb].copy()
    a, _ = norm(a)
    for s in range(1, W):
        t = tb + s
        live = t < starts[idx]
        an, _ = norm(fstep(a, np.minimum(t, starts[idx] - 1), S))
        a = np.where(live[:, None, None], an, a)
    sF[idx] = a
    # chunk 0: t=0 handled by making start such that the first step is pi*B0; emulate: run separately
    a0 = (S["pi"] * S["B"][0])[None]
    a0, c0 = norm(a0)
    al[0] = a0[0]
    ll0 = np.log(c0[0]) + S["m"][0]
    st0 = starts.copy()
    st0[0] = 1
    sF[0] = a0[0]
    allidx = np.arange(nc)
    eF = run_chunks_fwd(S, st0, ends, sF, allidx, al, llp)
    llp[0] += ll0
    # backward round 0
    idxb = np.arange(0, nc - 1)
    te = np.minimum(ends[idxb] + W, N - 1)     # start beta~ at SNP te flat
    b = np.full((len(idxb), Z, U), 1.0 / S["K"])
    for s in range(W):
        t = te - 1 - s
        live = t >= ends[idxb]
        bn, _ = norm(bstep(b, np.maximum(t, ends[idxb]) + 1, S))
        b = np.where(live[:, None, None], bn, b)
    sB[idxb] = b
    sB[nc - 1] = 1.0 / S["K"]
    # last chunk: beta_{N-1} = flat; its "start" is beta~ at SNP N (nonexistent): handle by running from N-1
    be[N - 1] = 1.0 / S["K"]
    en0 = ends.copy()
    en0[nc - 1] = N - 1
    eB = run_chunks_bwd(S, starts, en0, sB, allidx, be)
    rounds_f = rounds_b = 1
    runs_f = runs_b = nc
    while True:
        # boundary c (between chunk c-1 and c), c = 1..nc-1
        wantF = eF[:-1]                 # alpha^_{t0_c - 1}
        usedF = sF[1:]
        wantB = eB[1:]                  # beta~_{t0_c}
        usedB = sB[:-1]
        if crit == "rel":
            hi = np.maximum(wantF, usedF)
            dF = (np.abs(wantF - usedF) / np.maximum(hi, 1e-290)).max((1, 2))
            hi = np.maximum(wantB, usedB)
            dB = (np.abs(wantB - usedB) / np.maximum(hi, 1e-290)).max((1, 2))
        else:
            # gamma at t0-1 ~ alpha^_{t0-1} (.) beta~_{t0-1};  beta~_{t0-1} = one more step from wantB
            t0 = starts[1:]
            bext, _ = norm(bstep(wantB, t0, S))
            g = wantF * bext
            g /= g.sum((1, 2), keepdims=True)
            dF = (g * np.abs(usedF / np.maximum(wantF, 1e-300) - 1)).sum((1, 2))
            # gamma at t0 ~ alpha^_{t0} (.) beta~_{t0};  alpha^_{t0} = one more step from wantF
            aext, _ = norm(fstep(wantF, t0, S))
            g = aext * wantB
            g /= g.sum((1, 2), keepdims=True)
            dB = (g * np.abs(usedB / np.maximum(wantB, 1e-300) - 1)).sum((1, 2))
        badF = np.where(dF > tol)[0] + 1
        badB = np.where(dB > tol)[0]
        if verbose:
            print(f"  round {rounds_f}: fwd dirty {len(badF)} (max {dF.max():.2e})  bwd dirty {len(badB)} (max {dB.max():.2e})")
        if len(badF) == 0 and len(badB) == 0:
            break
        newF = eF.copy()
        newB = eB.copy()
        if len(badF):
            sF[badF] = eF[badF - 1]
            newF[badF] = run_chunks_fwd(S, st0, ends, sF[badF], badF, al, llp)
            rounds_f += 1
            runs_f += len(badF)
        if len(badB):
            sB[badB] = eB[badB + 1]
            newB[badB] = run_chunks_bwd(S, starts, en0, sB[badB], badB, be)
            rounds_b += 1
            runs_b += len(badB)
        eF, eB = newF, newB
        if rounds_f > 300:
            break
    g = al * be
    g /= g.sum((1, 2), keepdims=True)
    ll = llp.sum()
    return dict(rounds_f=rounds_f, rounds_b=rounds_b, runs_f=runs_f, runs_b=runs_b, nc=nc,
                ll_rel=abs(ll - ll_x) / abs(ll_x), ll_abs=abs(ll - ll_x), post_err=np.abs(g - g_x).max(),
                stat_rel=np.abs(g.sum(0) - g_x.sum(0)).max() / g_x.sum(0).max())


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    cfgs = {"init": dict(s=(0.001, 0.2, 0.3), n=0.5, phi=2.0),
            "iter2": dict(s=(0.003, 0.33, 0.58), n=0.31, phi=2.0, var=0.0225)}
    for name, kw in cfgs.items():
        S = setup(N=N, **kw)
        ex = exact(S)
        print(f"== params {name}: N={S['N']} loglik {ex[3]:.6f}")
        for L in (1024, 512, 256):
            for W in (192,):
                for crit, tol in (("rel", 1e-8), ("rel", 1e-6), ("post", 1e-8), ("post", 1e-7), ("post", 1e-6)):
                    r = simulate(S, ex, L=L, W=W, crit=crit, tol=tol)
                    print(f"L={L:5d} W={W:4d} {crit:5s} tol={tol:.0e}: rounds f/b {r['rounds_f']:3d}/{r['rounds_b']:3d} "
                          f"runs f/b {r['runs_f']:5d}/{r['runs_b']:5d} of {r['nc']:4d}  ll_rel {r['ll_rel']:.1e} (abs {r['ll_abs']:.1e}) "
                          f"post_err {r['post_err']:.1e} stat_rel {r['stat_rel']:.1e}", flush=True)
