// titan_common.cuh -- types and small device helpers shared by every translation unit of the library.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

namespace titan {

struct __align__(16) SnpRec {   // parameter-independent per-SNP columns (HBM, read twice/E-step)
  double x;    // reference count              (data$ref)
  double nx;   // tumDepth - ref
  double lr;   // logR (natural log)
  double lb;   // lgamma(N+1)-lgamma(x+1)-lgamma(N-x+1), host libm (R/hmmClonal.R:618)
};
// Gaussian allele model (alleleEmissionModel = "Gaussian"): the same four slots carry r = ref / tumDepth, r*r, logR and
// r*logR, so that the six posterior-weighted sums of the final pass (slots x, nx, lr, lb, 1, lr*lr) come out as
// a, f, c, h, e, g of R/paramEstimation.R:30-32,37-39 without another kernel variant.

// Structured transition between SNP t-1 and t (unused at a chromosome start), stored once per lane class so that a
// lane reads its own coefficients without selects: half 0 serves non-HET genotypes, half 1 the diploid-HET genotype.
//   forward  (destination class):  o = cW*W + cG*Wg + cZ*Wz[z] + cw*w[z],  w = alpha*irs(source class)
//   backward (source class):       o = (k0*Wn + cB*H + cG*loc + k1*Wz[z] + cw*bb[z]) * irs
// k0 = oG(1-rhoZ), k1 = oG(2rhoZ-1), k2 = (rhoG-oG)(1-rhoZ), k3 = (rhoG-oG)(2rhoZ-1), cB = oG*rhoZ, k4 = (rhoG-oG)rhoZ
struct __align__(16) TxnHalf {
  double irs;          // 1/rowsum of a source of this class
  double cW, cG;       // k0, k2  |  cB, k4
  double cZ, cw;       // k1, k3  |  0, 0
  double k0, cB, k1;   // the class-independent terms of the backward step
};
struct __align__(16) TxnRec {
  TxnHalf h[2];
};

struct Chunk {
  int t0, t1;      // SNP range [t0, t1) owned by this chunk
  int cs, ce;      // chromosome range [cs, ce)
  int chr;         // chromosome index
  int pad;
};

struct ModelConsts {            // device pointers, refreshed every E-step (K- and U-sized)
  const double *lmu;    // [K] log(muR_j)
  const double *l1mu;   // [K] log(1-muR_j)
  const double *muC;    // [K]
  const double *cN;     // [U] log(1/(sqrt(var_u)*sqrt(2*pi)))
  const double *i2v;    // [U] 1/(2*var_u)       (forward-backward)
  const double *tv;     // [U] 2*var_u           (Viterbi, exact division as in the reference)
  const double *pi;     // [K] piG_u*piZ_z
  const double *logpi;  // [K] host-libm log of the above
  // Gaussian allele model (bivariateNormalpdflog, R/hmmClonal.R:550-559); then lmu holds muR itself, cN = -c_u,
  // i2v = l0/var_u and tv = var_u, and:
  const double *gR;     // [U] l0/varR_u
  const double *gX;     // [U] l0*2*corRho/sqrt(varR_u*var_u)
  const double *vR;     // [U] varR_u                  (Viterbi, exact divisions as in the reference)
  const double *sq;     // [U] sqrt(varR_u*var_u)
  double l0, rho2;      // 1/(2*(1-corRho^2)), 2*corRho
  int gauss;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_maxf(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 2^-exponent(v) for positive v: an exact power of two taken from the exponent field with integer instructions (it
// sits on the recursion's critical path).  Zero gives 2^1023 (0 stays 0); a non-finite v gives garbage next to a
// vector that is already non-finite -- the host rejects non-finite log-likelihoods.
__device__ __forceinline__ double pow2_rescale(double v, int &e_out) {
  const int be = (__double2hiint(v) >> 20) & 0x7ff;
  e_out = be - 1023;
  return __hiloint2double((2046 - be) << 20, 0);   // biased exponent 1023 - e
}

}  // namespace titan
