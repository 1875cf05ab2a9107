/* titancna_b200.h -- C ABI of the B200-native TitanCNA E-step hot path.
 *
 * Plain C: raw pointers + sizes, int status codes, no torch / CUDA types.  All matrices are
 * column-major exactly as R hands them over.  Every function returns 0 on success; on failure
 * it returns a non-zero code and titan_last_error() describes it (thread-local string).  The
 * library NEVER computes on the CPU: without a usable CUDA device every compute entry fails
 * with TITAN_ERR_NO_DEVICE.
 *
 * Reference interfaces replaced (paths relative to gavinha/TitanCNA 1.23.1):
 *   titan_fwd_back_clonalCN   <- SEXP fwd_backC_clonalCN(...)   src/titan.h:10, src/fwd_backC_clonalCN.c:40-211
 *   titan_viterbi_clonalCN    <- SEXP viterbiC_clonalCN(...)    src/titan.h:11, src/viterbiC_clonalCN.c:35-134
 *   titan_get_position_overlap<- SEXP getPositionOverlapC(...)  src/titan.h:9,  src/getPositionOverlapC.c:19-50
 *   titan_solver_* / titan_estep / titan_viterbi / titan_em_run
 *                             <- the N-sized bodies of runEMclonalCN / viterbiClonalCN that today run
 *                                in R around the .Call (R/hmmClonal.R:141-171,191-234,261-291,449-484;
 *                                R/paramEstimation.R:27-44,75-283,533-547)
 * The R-side glue that binds these (same .Call names/arity as src/register.c:8-13) is
 * titancna_b200/csrc/r_glue.c (reference-compatible routines) + r_glue_fused.c (solver routines); see INTEGRATION.md.
 */
#ifndef TITANCNA_B200_H
#define TITANCNA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TITAN_OK              0
#define TITAN_ERR_ARG         1   /* bad argument (message mirrors the reference's error() text) */
#define TITAN_ERR_NO_DEVICE   2   /* no CUDA device / driver: there is no CPU fallback */
#define TITAN_ERR_CUDA        3   /* CUDA runtime failure */
#define TITAN_ERR_UNSUPPORTED 4   /* configuration outside what the kernels cover */
#define TITAN_ERR_NUMERIC     5   /* non-finite values met where the reference would stop() */

const char *titan_last_error(void);
const char *titan_version(void);
int titan_device_count(int *count);
/* select the CUDA device used by the calling thread's subsequent calls (one process per GPU) */
int titan_set_device(int device);
/* the calling thread's current CUDA device (worker threads of a host program start on device 0, not on their
 * creator's device: a driver that fans solutions out to threads reads it here and passes it on) */
int titan_get_device(int *device);

/* ------------------------------------------------------------------------------------------
 * 1. Legacy-compatible entry points: one chromosome per call, materialised log-emissions in,
 *    same meaning of every argument as the reference's .Call routines.
 *    log_piGiZi[K]      log initial distribution (K = U*Z, +1 with the outlier state)
 *    py[K x T]          log emissions, column-major (state fastest)
 *    copyNumKey[nCT]    ct (unused by the maths, kept for ABI fidelity; may be NULL)
 *    zygosityKey[nZS]   ZS per unit state; -1 marks diploid HET
 *    numClust           Z;  numUnitStates = K / Z by integer division (fwd_backC_clonalCN.c:71)
 *    positions[T]       base-pair positions (ascending)
 *    zStrength          txnZstrength * txnExpLen ;  txnLen = txnExpLen
 *    useOutlier         0/1 (fwd_backC_clonalCN.c:224-247 semantics)
 *    Layouts outside the structured kernels (outlier state, repeated zygosity keys, K % numClust != 0,
 *    > 32 unit states) are served by dense O(K^2) kernels with the same semantics.
 * ---------------------------------------------------------------------------------------- */
/* rho_out[K x T] receives LOG posteriors (log 0 = -inf where the reference holds a finite
 * ~-750: compare after exp), loglik_out[1] the chromosome log-likelihood. */
int titan_fwd_back_clonalCN(const double *log_piGiZi, int K, const double *py, int T,
                            const double *copyNumKey, int nCT, const double *zygosityKey, int nZS,
                            int numClust, const double *positions, double zStrength, double txnLen,
                            int useOutlier, double *rho_out, double *loglik_out);

/* path_out[T]: 1-based joint state index, bit-identical to the reference (first-index ties). */
int titan_viterbi_clonalCN(const double *log_piGiZi, int K, const double *py, int T,
                           const double *copyNumKey, int nCT, const double *zygosityKey, int nZS,
                           int numClust, const double *positions, double zStrength, double txnLen,
                           int useOutlier, int *path_out);

/* out[T]: 1-based index of the first [start,stop] interval containing posn[i], else 0.
 * Registered but never called by the reference's R code; host-side, kept for ABI completeness. */
int titan_get_position_overlap(const double *posn, int T, const double *start, const double *stop,
                               int Tcn, double *out);

/* ------------------------------------------------------------------------------------------
 * 2. Fused path: the data set lives on the device; only O(U*Z) numbers cross per EM iteration.
 * ---------------------------------------------------------------------------------------- */
typedef struct titan_solver titan_solver;

typedef struct titan_data_desc {
  int64_t N;                 /* SNPs, all chromosomes concatenated in order                          */
  int32_t nChr;              /* chromosomes (contiguous runs, R/hmmClonal.R:120-123)                 */
  const int64_t *chr_start;  /* [nChr+1] offsets into the columns, chr_start[nChr] == N              */
  const double *posn;        /* [N] positions                                                        */
  const double *ref;         /* [N] symmetric reference counts  x = max(ref, nonRef)                 */
  const double *tumDepth;    /* [N] depths                                                           */
  const double *logR;        /* [N] natural-log ratios                                               */
  double txnExpLen;          /* runEMclonalCN(txnExpLen=)                                            */
  double txnZstrength;       /* runEMclonalCN(txnZstrength=); zStrength = txnZstrength*txnExpLen     */
  int32_t U;                 /* genotype (unit) states                                               */
  int32_t Z;                 /* clonal clusters                                                      */
  const double *ZS;          /* [U] zygosity key, -1 = diploid HET; other entries must be distinct   */
  int32_t alleleModel;       /* genotypeParams$alleleEmissionModel: TITAN_ALLELE_BINOMIAL (0, default) or
                                TITAN_ALLELE_GAUSSIAN (1): bivariate normal on (ref / tumDepth, logR),
                                R/hmmClonal.R:149-158,525-559                                         */
} titan_data_desc;
#define TITAN_ALLELE_BINOMIAL 0
#define TITAN_ALLELE_GAUSSIAN 1

/* Uploads the columns once, computes the parameter-independent per-SNP terms on the HOST with the
 * host libm (rhoG/rhoZ: fwd_backC_clonalCN.c:342-348,124-125; lgamma bracket: R/hmmClonal.R:618),
 * partitions chromosomes into position chunks and allocates all device scratch. */
int titan_solver_create(const titan_data_desc *desc, titan_solver **out);
void titan_solver_destroy(titan_solver *s);

/* Shape of the data set a solver holds (any pointer may be NULL). */
int titan_solver_shape(const titan_solver *s, int64_t *N, int32_t *nChr, int32_t *U, int32_t *Z);

/* Tuning knobs (defaults in DESIGN.md).  chunk_len <= 0: one chunk per chromosome (plain
 * sequential recursion).  Otherwise chromosomes are cut into chunks of chunk_len SNPs that run
 * in parallel from warm-started boundary vectors and are re-run until every boundary agrees
 * with its predecessor to rel_tol (exact chunked scan by fixed-point verification). */
int titan_solver_configure(titan_solver *s, int chunk_len, int warmup_len, double rel_tol, int max_rounds);
/* Launch on a caller-owned stream (a cudaStream_t passed as void*; NULL = the solver's own). */
int titan_solver_set_stream(titan_solver *s, void *cuda_stream);

typedef struct titan_params {
  /* model parameters of one EM iteration: clonalTwoComponentMixtureCN inputs + variances + pi */
  double n;                  /* normal contamination                                                 */
  double phi;                /* ploidy                                                               */
  const double *s;           /* [Z] cellular prevalence parameters                                   */
  const double *var;         /* [U] logR variances                                                   */
  const double *piG;         /* [U] initial genotype distribution                                    */
  const double *piZ;         /* [Z] initial cluster distribution                                     */
  const double *rt;          /* [U] tumour reference allelic ratios                                  */
  const double *ct;          /* [U] tumour copy numbers                                              */
  double rn;                 /* normal reference allelic ratio                                       */
  /* Gaussian allele model only (ignored by a binomial solver) */
  const double *varR;        /* [U] allelic-ratio variances (convergeParams$varR column)             */
  double corRho;             /* correlation of the bivariate normal: genotypeParams$corRho_0 in the EM; the
                                Viterbi wrapper passes the data correlation (R/hmmClonal.R:447,536-538) */
} titan_params;

typedef struct titan_estep_out {
  /* all host pointers, caller-allocated; any of the optional ones may be NULL */
  double *loglik_chr;        /* [nChr] forward-backward log-likelihood per chromosome               */
  double *stats;             /* [6][U*Z]: a,b,c,d,e,g of R/paramEstimation.R:34-40, joint index u+z*U;
                                Gaussian allele model: a,f,c,h,e,g of :30-32,37-39 in the same six slots */
  double *rhoG0;             /* [U x nChr] rhoG at the first SNP of each chromosome (chrPos_0)       */
  double *rhoZ0;             /* [Z x nChr] rhoZ at the first SNP of each chromosome                  */
  double *rhoG;              /* optional [U x N]  (R/hmmClonal.R:234)                                */
  double *rhoZ;              /* optional [Z x N]  (R/hmmClonal.R:233)                                */
  double *muR;               /* optional [U x Z] mixture means used (R/hmmClonal.R:595)              */
  double *muC;               /* optional [U x Z]                    (R/hmmClonal.R:596)              */
} titan_estep_out;

/* One E-step: emissions -> scaled forward-backward -> posterior marginals -> the six weighted
 * sums.  Replaces R/hmmClonal.R:191-234 + R/paramEstimation.R:27-44 for all chromosomes at once. */
int titan_estep(titan_solver *s, const titan_params *p, titan_estep_out *out);

/* Viterbi decode of all chromosomes: path_out[N], 1-based joint state index u + (z-1)*U.
 * Emission parameters from `p`; initial distribution from p->piG / p->piZ (the caller passes the
 * i-1 column, R/hmmClonal.R:466-467). */
int titan_viterbi(titan_solver *s, const titan_params *p, int32_t *path_out);

typedef struct titan_hyper {
  /* loadDefaultParameters() content needed by the M-step and the priors (R/utils.R:101-146) */
  const double *rt, *ct, *ZS;         /* [U]                                                         */
  double rn;
  const double *var_0;                /* [U]                                                         */
  const double *alphaKHyper, *betaKHyper, *kappaGHyper, *piG_0;   /* [U]                             */
  const double *s_0, *alphaSHyper, *betaSHyper, *kappaZHyper, *piZ_0; /* [Z]                         */
  double n_0, alphaNHyper, betaNHyper;
  double phi_0, alphaPHyper, betaPHyper;
  /* Gaussian allele model only (R/utils.R:108-119); may be NULL / 0 for the binomial model */
  const double *varR_0, *alphaRHyper, *betaRHyper;                /* [U]                             */
  double corRho_0;
} titan_hyper;

typedef struct titan_em_opts {
  int32_t maxiter;            /* runEMclonalCN(maxiter=)                                             */
  int32_t maxiterUpdate;      /* runEMclonalCN(maxiterUpdate=)                                       */
  int32_t normal_map;         /* normalEstimateMethod == "map"                                       */
  int32_t estimateS;
  int32_t estimatePloidy;
  double likChangeThreshold;
  int32_t want_posteriors;    /* fill rhoG / rhoZ of the last E-step                                 */
  int32_t verbose;
} titan_em_opts;

typedef struct titan_em_out {
  /* caller-allocated with an iteration axis of maxiter+1 columns (column 0 = initial values) */
  int32_t iters;              /* OUT: number of valid columns (the R code's final i)                 */
  double *n, *phi, *loglik;   /* [maxiter+1]                                                         */
  double *s, *piZ;            /* [Z x (maxiter+1)]                                                   */
  double *var, *piG;          /* [U x (maxiter+1)]                                                   */
  double *muR, *muC;          /* [U x Z x (maxiter+1)]                                               */
  double *rhoG, *rhoZ;        /* optional [U x N], [Z x N]                                           */
  int32_t *cd_sweeps;         /* optional [maxiter+1]: coordinate-descent sweeps of each M-step      */
  double estep_ms;            /* OUT: device time spent in E-step kernels (CUDA events)              */
  double mstep_ms;            /* OUT: host time spent in the M-step                                  */
  double reduce_ms;           /* OUT: part of estep_ms spent in the final reduction and, for a sharded solution, in
                                 the all-reduce (which includes waiting for the slowest rank)         */
  double *varR;               /* optional [U x (maxiter+1)]: allelic variances (constant columns for the binomial
                                 model, R/paramEstimation.R:212)                                      */
} titan_em_out;

/* The whole runEMclonalCN loop (the solver's allele model, no outlier state): E-steps on the device,
 * the O(U*Z) coordinate-descent M-step (R/paramEstimation.R:75-283) and priors
 * (R/hmmClonal.R:715-783) on the host. */
int titan_em_run(titan_solver *s, const titan_hyper *h, const titan_em_opts *o, titan_em_out *out);

/* Host M-step alone (what titan_em_run runs between E-steps), exposed for chromosome-sharded
 * multi-GPU runs where the statistics are all-reduced by the caller first.
 * stats as in titan_estep_out; prev_* are the i-1 column; out_* receive column i. */
int titan_mstep(const titan_hyper *h, int U, int Z, int nChr, int maxiterUpdate, int normal_map,
                int estimateS, int estimatePloidy, const double *stats, const double *rhoG0,
                const double *rhoZ0, double prev_n, const double *prev_s, const double *prev_var,
                double prev_phi, const double *prev_piG, const double *prev_piZ, double *out_n,
                double *out_s, double *out_var, double *out_phi, double *out_piG, double *out_piZ,
                int *out_sweeps, double *out_prior);

/* The same for the Gaussian allele model (R/paramEstimation.R:165-200,313-329,376-383,415-422,463-497,510-515):
 * stats = a,f,c,h,e,g as titan_estep returns them from a Gaussian solver; h->varR_0 / alphaRHyper / betaRHyper /
 * corRho_0 are required. */
int titan_mstep_gaussian(const titan_hyper *h, int U, int Z, int nChr, int maxiterUpdate, int normal_map,
                         int estimateS, int estimatePloidy, const double *stats, const double *rhoG0,
                         const double *rhoZ0, double prev_n, const double *prev_s, const double *prev_var,
                         const double *prev_varR, double prev_phi, const double *prev_piG, const double *prev_piZ,
                         double *out_n, double *out_s, double *out_var, double *out_varR, double *out_phi,
                         double *out_piG, double *out_piZ, int *out_sweeps, double *out_prior);

/* Introspection for benchmarks and tests */
typedef struct titan_timing {
  double last_estep_ms;       /* CUDA-event time of the last titan_estep (all its kernels)           */
  double last_fwd_ms, last_bwd_ms, last_reduce_ms;   /* engine 2: emission + all rounds / final pass / reductions */
  double last_fwd_r0_ms, last_bwd_r0_ms;             /* the round-0 launch alone (every chunk runs) */
  double last_viterbi_ms;
  int32_t last_fwd_rounds, last_bwd_rounds;   /* verification rounds the chunked scan needed          */
  int64_t last_fwd_chunk_runs, last_bwd_chunk_runs;
  int64_t kernel_launches;    /* cumulative launches of this library's kernels                       */
  int32_t n_chunks;
  int32_t engine;             /* 2: emission pass + combined rounds + final pass (default); 1: round-1 kernels */
  double last_emission_ms;    /* engine 2: the emission pass; last_fwd_r0_ms = round 0 (both directions),     */
  double last_rounds_ms;      /*           verification rounds >= 1, last_bwd_r0_ms = the final posterior pass */
} titan_timing;
int titan_solver_timing(const titan_solver *s, titan_timing *t);

/* Device buffers of destroyed solvers stay cached in the CUDA memory pool of the current device (a sweep
 * creates one solver per solution; cudaMalloc/cudaFree of such blocks cost more than an E-step), and the
 * parameter-independent tables of the last few (positions, transition settings, state layout) combinations --
 * transition coefficients (128 B per SNP) and the decode's log-transition tables (32 K bytes per distinct gap) --
 * stay resident so that the solutions of a sweep share them.  This drops those tables (live solvers keep theirs)
 * and returns the cached memory to the driver.  No counterpart in the reference (R frees through its GC). */
int titan_release_cached_memory(void);

/* ------------------------------------------------------------------------------------------
 * 3. One solution over several GPUs: chromosomes sharded across ranks (one process per GPU), ONE
 *    all-reduce(sum, fp64) of [a,b,c,d,e,g | loglik per chromosome | rhoG, rhoZ at chrPos_0] per E-step
 *    on the device, M-step replicated on every rank.  Replaces the sum over the reference's
 *    foreach(c = 1:numChrs) %dopar% workers (R/hmmClonal.R:203-234) when the workers are GPUs.
 *    NCCL is bound at run time (libnccl.so.2); without it these calls fail with TITAN_ERR_UNSUPPORTED.
 * ---------------------------------------------------------------------------------------- */
#define TITAN_COMM_ID_BYTES 128
typedef struct titan_comm titan_comm;
/* rank 0 draws an id and ships the 128 bytes to the other ranks by any means the host has
 * (R: parallel / Rmpi / a file; Python: torch.distributed.broadcast_object_list) */
int titan_comm_unique_id(char *id_out /* [TITAN_COMM_ID_BYTES] */);
/* collective over all `world` ranks, on the calling thread's current device (titan_set_device) */
int titan_comm_create(const char *id, int rank, int world, titan_comm **out);
void titan_comm_destroy(titan_comm *c);
int titan_comm_info(const titan_comm *c, int *rank, int *world, int *nccl_version);
/* The solver holds chromosomes chr_global[0..nChr) (ascending) of a data set with nChr_total chromosomes.
 * From here on titan_estep / titan_em_run return whole-genome quantities on every rank: stats summed
 * over ranks, loglik_chr[nChr_total], rhoG0[U x nChr_total], rhoZ0[Z x nChr_total]; the N-sized outputs
 * (rhoG, rhoZ, Viterbi path) stay per shard.  comm == NULL detaches. */
int titan_solver_set_comm(titan_solver *s, titan_comm *comm, int nChr_total, const int32_t *chr_global);

/* FP64 roofline denominator, measured: DFMA throughput (TFLOP/s, 2 flop per DFMA) of the current device with every
 * SM busy, best of five launches; ms_out (optional) = duration of that launch.  SURVEY.md 8(d) asks for it because
 * MEASURED_PEAKS.json has no fp64 figure.  No counterpart in the reference. */
int titan_measure_fp64_peak(double *tflops_out, double *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* TITANCNA_B200_H */
