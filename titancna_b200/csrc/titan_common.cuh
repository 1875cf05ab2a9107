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

struct __align__(16) TxnRec {   // structured transition between SNP t-1 and t (unused at chr start)
  double k0, k1, k2, k3, cB, k4, irn, irh;   // irn/irh = 1/rowsum of a non-HET / HET source
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

// 2^-exponent(v) for positive finite v (exact power of two; 1.0 for 0 / subnormal / non-finite input).
// Branch-free: it sits on the recursion's critical path.
__device__ __forceinline__ double pow2_rescale(double v, int &e_out) {
  const int be = (__double2hiint(v) >> 20) & 0x7ff;
  const bool ok = (unsigned)(be - 1) < 0x7feu;
  e_out = ok ? be - 1023 : 0;
  return __hiloint2double(ok ? ((2046 - be) << 20) : 0x3ff00000, 0);   // biased exponent 1023 - e
}

}  // namespace titan
