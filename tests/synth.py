"""Seeded synthetic genome-wide allele-count / logR data.

Shared by tests/ and bench.py.  Pure numpy, no reference or oracle dependency.
"""
from __future__ import annotations

import numpy as np

HG19_LEN = [249250621, 243199373, 198022430, 191154276, 180915260, 171115067, 159138663, 146364022,
            141213431, 135534747, 135006516, 133851895, 115169878, 107349540, 102531392, 90354753,
            81195210, 78077248, 59128983, 63025520, 48129895, 51304566, 155270560]
CHR_NAMES = [str(i) for i in range(1, 23)] + ["X"]
S_TRUE = (0.0, 0.35, 0.6, 0.75, 0.85, 0.9, 0.93, 0.96)   # first five: SURVEY.md 8d

# symmetric genotype table for CN 0..8 (same content as loadDefaultParameters: ct, major fraction)
_CT, _RT = [0, 1], [0.5, 1.0]
for _c in range(2, 9):
    for _m in range(_c, (_c + 1) // 2 - 1, -1):
        _CT.append(_c)
        _RT.append(_m / _c)


def _mixture(rt, ct, n, s, phi, rn=0.5):
    rt, ct, s = np.asarray(rt)[:, None], np.asarray(ct, float)[:, None], np.asarray(s)[None, :]
    num = n * rn * 2 + (1 - n) * s * rn * 2 + (1 - n) * (1 - s) * rt * ct
    tot = n * 2 + (1 - n) * s * 2 + (1 - n) * (1 - s) * ct
    return num / tot, np.log(tot / (n * 2 + (1 - n) * phi))


def make_data(N, Z=3, maxCN=8, n_chr=23, seed=20261019, mean_seg=5000, n=0.3, phi=2.0, depth_mean=60,
              logr_sd=0.15, exome=False, haplotype_bins=False, bin_size=100000):
    """-> (data dict with chr/posn/ref/tumDepth/logR, truth dict).  chr is an int array 1..n_chr."""
    rng = np.random.default_rng(seed)
    lens = np.array(HG19_LEN[:n_chr], dtype=np.float64)
    per = np.maximum(2, np.floor(N * lens / lens.sum()).astype(np.int64))
    per[0] += N - per.sum()
    U = int(np.sum(np.array(_CT) <= maxCN))
    rt, ct = np.array(_RT[:U]), np.array(_CT[:U])
    het_idx = [i for i in range(U) if ct[i] >= 2 and abs(rt[i] - 0.5) < 1e-12]
    s = np.array(S_TRUE[:Z])
    muR, muC = _mixture(rt, ct, n, s, phi)
    chr_, posn, gu, gz = [], [], [], []
    for c in range(n_chr):
        T = int(per[c])
        if exome:
            n_ex = max(1, int(T * 0.4))
            starts = np.sort(rng.choice(int(lens[c]) - 400, size=n_ex, replace=False)) + 1
            which = np.sort(rng.integers(0, n_ex, size=T))
            p = starts[which] + rng.integers(0, 300, size=T)
            p = np.unique(p)
            while p.size < T:
                p = np.unique(np.concatenate([p, rng.integers(1, int(lens[c]), size=T - p.size)]))
            p = np.sort(p[:T])
        else:
            p = np.unique(rng.integers(1, int(lens[c]) + 1, size=int(T * 1.02) + 8))
            while p.size < T:
                p = np.unique(np.concatenate([p, rng.integers(1, int(lens[c]) + 1, size=T)]))
            p = np.sort(rng.choice(p, size=T, replace=False))
        # piecewise-constant truth
        u = np.empty(T, dtype=np.int64)
        z = np.empty(T, dtype=np.int64)
        t = 0
        while t < T:
            L = int(rng.geometric(1.0 / mean_seg))
            uu = het_idx[0] if rng.random() < 0.6 else int(rng.integers(0, U))
            zz = int(rng.integers(0, Z))
            u[t:t + L] = uu
            z[t:t + L] = zz
            t += L
        chr_.append(np.full(T, c + 1, dtype=np.int32)); posn.append(p); gu.append(u); gz.append(z)
    chr_, posn, gu, gz = (np.concatenate(a) for a in (chr_, posn, gu, gz))
    depth = np.clip(rng.poisson(depth_mean, size=posn.size), 10, 1000).astype(np.float64)
    ref = rng.binomial(depth.astype(np.int64), muR[gu, gz]).astype(np.float64)
    ref = np.maximum(ref, depth - ref)                      # symmetric mode (R/utils.R:186-187)
    logR = rng.normal(muC[gu, gz], logr_sd)
    if haplotype_bins:                                      # 10X haplotype-block mode (R/haplotype.R:118-174)
        key = chr_.astype(np.int64) * 100000 + posn // bin_size
        _, inv = np.unique(key, return_inverse=True)
        ref = np.bincount(inv, weights=ref)[inv]
        depth = np.bincount(inv, weights=depth)[inv]
    data = {"chr": chr_, "posn": posn.astype(np.float64), "ref": ref, "tumDepth": depth, "logR": logR}
    truth = {"u": gu, "z": gz, "n": n, "phi": phi, "s": s, "U": U}
    return data, truth
