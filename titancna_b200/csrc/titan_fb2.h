// titan_fb2.h -- host-visible interface of the forward-backward engine (titan_fb2.cu).
//
// E-step pipeline (reference: R/hmmClonal.R:160-171,191-234; src/fwd_backC_clonalCN.c:111-192;
// R/paramEstimation.R:27-44):
//   1. emission kernel     B[t][j] = exp(py[j,t] - m_t), m_t      one pass, fully parallel, K x N fp64 in HBM
//   2. round kernel r = 0  every chunk: forward (alpha stored) and backward (beta only) from warm-started guesses
//   3. round kernel r >= 1 every chunk compares the boundary vector it used with the one its neighbour actually
//                          produced, weighted by the posterior at the boundary; mismatching chunks re-run
//   4. final kernel        every chunk: backward from its verified boundary, posteriors, the six weighted sums
// One warp owns one chunk; its per-SNP rows (B, transition coefficients, alpha, records) arrive through a
// private shared-memory ring filled by TMA bulk copies (cp.async.bulk + mbarrier), issued by lane 0.
#pragma once
#include <cuda_runtime.h>

#include "titan_common.cuh"

namespace titan {

constexpr int kFbPadRows = 16;   // slack rows before and after every per-SNP array: tiles are copied whole
#ifndef TITAN_PROBE_PIECES
#define TITAN_PROBE_PIECES 3
#endif
constexpr int kProbes = TITAN_PROBE_PIECES;   // pieces a dirty chunk's start vector is split into: kProbes - 1 single states + the remainder

struct Fb2Args {
  const double *Bm;            // [N][Kp] scaled emissions; column K holds the shift m_t
  const TxnRec *txn;           // [N]
  const SnpRec *snp;           // [N]
  const Chunk *chunks;         // [nChunks]
  int nChunks;
  int U, K, Kp;
  int warm;                    // warm-up length (SNPs) of the speculative round-0 starts
  int round;
  double tol;                  // posterior-weighted boundary tolerance
  const unsigned char *het;    // [U]
  const double *pi;            // [K]
  double *alpha;               // [N][Kp] scaled forward variables
  double *startF, *endF, *extF;   // [nChunks][K], [2][nChunks][K], [2][nChunks][K]
  double *startB, *endB, *extB;
  double *subB;                // [nChunks][3][K] beta~ entering the lower three quarters of a chunk (final pass)
  double *chunk_ll;            // [nChunks]
  unsigned int *rerun;         // [rounds] chunk runs (both directions) of each round
  // probe-accelerated rounds (see fb2_probe_kernel): per direction [nChunks] flags and [nChunks][3][K] probe results
  int use_solved;              // round kernel: chunks flagged in solved* start from sol* without a test
  int probe_window;            // also probe this many chunks downstream of a dirty one (a re-run shifts its successors' inputs)
  unsigned char *dirtyF, *dirtyB, *solvedF, *solvedB, *probedF, *probedB;
  int *probeJF, *probeJB;      // [nChunks][kProbes - 1] the states probed individually
  int *probeEF, *probeEB;      // [nChunks][kProbes] accumulated power-of-two exponents of the probe results
  double *probeF, *probeB;     // [nChunks][3][K]
  double *solF, *solB;         // [nChunks][K] boundary vectors predicted by the chain solve
  const int *chr_chunk0;       // [nChr+1] first chunk of every chromosome
  int nChr;
  // final pass
  // final_mode 0: every chunk.  1 (speculative, on a second stream while the verification cycles still run, once per
  // cycle): every chunk without a valid pass that the current cycle does not re-run; a pass records the chunk's run
  // counter it started from.  2 (closing): the chunks that ran again since their last pass (or never had one).
  int final_mode;
  unsigned int *ver;           // [nChunks] += 1 whenever the chunk runs (either direction) in a round >= 1
  unsigned int *fin_ver;       // [nChunks] value of ver[c] the last final pass of chunk c started from
  double *part;                // [nChunks][6][K]
  double *rhoG0, *rhoZ0;       // [nChr][U], [nChr][Z]
  double *rhoG, *rhoZ;         // optional [N][U], [N][Z]
  double *loggamma;            // optional [N][K] (legacy boundary)
};

struct EmArgs {
  const SnpRec *snp;           // [N]
  const double *py;            // legacy: [N][K] log emissions, else nullptr
  ModelConsts M;
  int U, K, Kp;
  long long N;
  double *Bm;                  // [N][Kp]
};

// all launchers return the CUDA launch status; Z in 1..8, U <= 64
cudaError_t fb2_launch_emission(const EmArgs &E, int Z, cudaStream_t st);
cudaError_t fb2_launch_round(const Fb2Args &A, int Z, cudaStream_t st);
cudaError_t fb2_launch_probe(const Fb2Args &A, int Z, cudaStream_t st);    // boundary test + probe runs of the dirty chunks
cudaError_t fb2_launch_solve(const Fb2Args &A, int Z, cudaStream_t st);    // per-chromosome chain solve from the probe results
cudaError_t fb2_launch_final(const Fb2Args &A, int Z, bool write_post, bool from_py, cudaStream_t st);
bool fb2_supported(int U, int Z);

}  // namespace titan
