/*
 * ntn_b200.h -- C-ABI of the B200-native NTN cost+gradient hot path.
 *
 * This is the drop-in boundary for the path that Run_NTN's L-BFGS loop calls once per
 * function evaluation in c0der1337/Neural_Tensor_Network_For_KB_Completion:
 *
 *     IPair<Double,double[]> NTN.computeAt(double[] theta)        src/NTN.java:92-443
 *
 * plus the state that method reads implicitly from the DataFactory singleton
 * (src/DataFactory.java:21,67,89-98,393-437,583-595,614-684).  A JNI / Panama-FFM / ctypes
 * binding of the reference binds exactly these symbols.  Plain pointers
 * and sizes only; no C++/torch types.  Every function returns 0 on success and a negative
 * NTN_E* code otherwise; ntn_last_error() gives the text.  There is NO CPU fallback: without
 * a CUDA device ntn_create fails with NTN_ECUDA.
 *
 * Threading: one handle = one host thread at a time (the reference's NTN/DataFactory are a
 * stateful object and a non-thread-safe singleton).  One handle drives one GPU; under data
 * parallelism there is one process and one handle per GPU (DESIGN.md "Multi-GPU").
 */
#ifndef NTN_B200_H
#define NTN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NTN_OK        0
#define NTN_EINVAL   -1   /* bad argument / index out of range / call order          */
#define NTN_ECUDA    -2   /* CUDA runtime error (no device, OOM, launch failure)      */
#define NTN_ESTATE   -3   /* missing set_entity_words / set_batch / frozen vectors    */
#define NTN_EUNSUP   -4   /* shape outside what the kernels are built for            */

#define NTN_ACT_TANH     0   /* Run_NTN.java:78 */
#define NTN_ACT_SIGMOID  1

#define NTN_WV_FROZEN    0   /* entity vectors from the load-time word vectors: what the
                                reference does (DataFactory.java:393-437 reads worvectorWordVec,
                                never theta -- SURVEY.md App. B1) */
#define NTN_WV_THETA     1   /* entity vectors from theta's word block (the differentiable,
                                intended behaviour; createEntityVectorsByWordEmbeddings :515-581) */

#define NTN_GEMM_AUTO     0  /* tcgen05 when the shape allows it, else SIMT            */
#define NTN_GEMM_SIMT     1  /* FP32 CUDA-core contractions (A/B reference path)       */
#define NTN_GEMM_TCGEN05  2  /* tcgen05/TMEM + TMA, 3xTF32 split                       */

#define NTN_SCORE_REFERENCE_EVAL 0 /* NTN.calculateScoreOfaTripple, NTN.java:713-736: no activation,
                                      U scales only the bilinear term (App. B6) */
#define NTN_SCORE_TRAIN          1 /* u' f(e1'W e2 + V[e1;e2] + b), NTN.java:163-205 */

typedef struct ntn_handle ntn_handle;

/* Constructor arguments of NTN (NTN.java:42-66) + the DataFactory sizes it reads. */
typedef struct ntn_config {
    int32_t d;              /* embeddingSize            (Run_NTN.java:73)  1..256            */
    int32_t k;              /* sliceSize                (Run_NTN.java:76)  1..8              */
    int32_t R;              /* numberOfRelations                                             */
    int32_t Ne;             /* numberOfEntities                                              */
    int32_t Nw;             /* numOfWords                                                    */
    int32_t batch_divisor;  /* batchSize: divisor of cost and gradient (NTN.java:398-401,429-430) */
    float   lambda;         /* lamda (Run_NTN.java:79)                                       */
    int32_t activation;     /* NTN_ACT_*                                                     */
    int32_t wv_source;      /* NTN_WV_*                                                      */
    int32_t gemm_path;      /* NTN_GEMM_*                                                    */
    int32_t device;         /* CUDA device ordinal                                           */
    float   reg_scale;      /* multiplies the regulariser (cost and gradient).  1 for a single
                               process; 1/G on each of G data-parallel ranks so that the SUM
                               all-reduce of the per-rank results is the full-batch result   */
} ntn_config;

/* ---- lifetime ------------------------------------------------------------------------- */
int  ntn_create(const ntn_config* cfg, ntn_handle** out);
void ntn_destroy(ntn_handle* h);
/* Text of the last error on this handle (h == NULL: last ntn_create failure on this thread). */
const char* ntn_last_error(const ntn_handle* h);

/* NTN.getDimension / domainDimension (NTN.java:445-448,770-773):
 * P = R*k*d*d + R*2d*k + R*k + R*k + d*Nw. */
int64_t ntn_dimension(const ntn_handle* h);

/* ---- DataFactory state the path reads ------------------------------------------------- */
/* Entity -> word maps.  fwd_* is the CSR of words averaged into the entity vector
 * (DataFactory.java:405-434, with multiplicity, "unknown" fallback already applied);
 * bwd_* is getWordIndexes(i) (DataFactory.java:646-684) and len_bwd[i] = entityLength(i)
 * (:583-595), the divisor in NTN.java:420.  bwd_off/bwd_idx may be NULL (= forward lists).
 * Word ids in [0,Nw); every entity needs >= 1 forward word; len_bwd[i] < 0 is rejected.  len_bwd[i] == 0 (the
 * Wordnet name "__2", DataFactory.java:650: the reference divides by zero there, App. B8) is accepted and that
 * entity contributes nothing to the word gradient.  Host pointers; copied. */
int ntn_set_entity_words(ntn_handle* h,
                         const int32_t* fwd_off /* Ne+1 */, const int32_t* fwd_idx,
                         const int32_t* bwd_off /* Ne+1 or NULL */, const int32_t* bwd_idx,
                         const int32_t* len_bwd /* Ne */);

/* Load-time word vectors, d x Nw row-major (= DataFactory.wordVectorMaxtrixLoaded,
 * DataFactory.java:383-387).  Required when wv_source == NTN_WV_FROZEN. */
int ntn_set_frozen_wv(ntn_handle* h, const float* wv /* d*Nw */);

/* The batch job (DataFactory.generateNewTrainingBatchJob, DataFactory.java:116-136): n columns
 * (e1, rel, e2, e3_corrupt) in batch order.  Indices are validated.  Columns sharing
 * (e1, rel, e2) are grouped internally; results depend on column order only through FP32
 * summation order. */
int ntn_set_batch(ntn_handle* h, int64_t n,
                  const int32_t* e1, const int32_t* rel, const int32_t* e2, const int32_t* e3);
/* The same with the columns already in device memory (e.g. corrupt entities sampled on the GPU).  The per-relation
 * grouping the reference redoes on every evaluation (DataFactory.getBatchJobTripplesOfRelation, DataFactory.java:89-98;
 * index vectors :614-645) is done here once per batch job, on the device; the arrays may be released on return. */
int ntn_set_batch_dev(ntn_handle* h, int64_t n,
                      const int32_t* d_e1, const int32_t* d_rel, const int32_t* d_e2, const int32_t* d_e3);

/* ---- batch-job generation on the device (sampler variants of DataFactory.generateNewTrainingBatchJob) ------------
 * The reference draws e3 = rand.nextInt(numOfentities) on the host, h-major, always from the first batch_size training
 * triples (DataFactory.java:116-136; the shuffle is commented out at :118).  ntn_set_batch with a host-generated job
 * (seeded java.util.Random, csrc/host/ntn_host.hpp) is the parity mode.  The calls below are the fast mode: the
 * training triples live in HBM, column j = h*batch + i copies training triple first+i, and its corrupt entity is the
 * counter-based draw  e3 = (mix(seed, job, j, attempt) >> 32) * Ne >> 32  (mix: a splitmix64-style finaliser, restated
 * in oracle/data_ref.py) -- reproducible anywhere, no host round trip.  filter_true != 0 redraws (attempt = 1..15) while
 * (e1, rel, e3) or (e3, rel, e2) is a known training triple.  The job is indexed like ntn_set_batch_dev. */
int ntn_set_training_triples(ntn_handle* h, int64_t T, const int32_t* e1, const int32_t* rel, const int32_t* e2);
int ntn_generate_batch_dev(ntn_handle* h, int64_t first, int32_t batch, int32_t corrupt, uint64_t seed, uint64_t job,
                           int32_t filter_true);
/* Download the columns of the current batch job when the library holds them (after ntn_set_batch with host pointers
 * or ntn_generate_batch_dev); n must be the job's column count. */
int ntn_get_batch_columns(ntn_handle* h, int64_t n, int32_t* e1, int32_t* rel, int32_t* e2, int32_t* e3);

/* ---- the hot path --------------------------------------------------------------------- */
/* computeAt (NTN.java:92-443).  theta and grad are host arrays of P doubles in the reference's
 * flattening order W | V | b | U | wordvectors (NTN.java:450-526); values are rounded to FP32
 * on entry exactly as Util.convertDoubleArrayToFlattenedINDArray does (Util.java:49-55).
 * corrupt_side[r] != 0  <=>  the reference's Math.random() > 0.5 for relation r (NTN.java:146):
 * negatives are (e1, e3); 0: (e3, e2).  n_active (may be NULL) = this call's increment of the
 * reference's `update` counter (NTN.java:222). */
int ntn_eval(ntn_handle* h, const double* theta, const uint8_t* corrupt_side /* R */,
             double* cost, double* grad, int64_t* n_active);

/* Same with FP32 host arrays (halves the bytes that cross PCIe; same arithmetic). */
int ntn_eval_f32(ntn_handle* h, const float* theta, const uint8_t* corrupt_side,
                 double* cost, float* grad, int64_t* n_active);

/* Device-resident fast path: theta never leaves HBM.  d_theta / d_grad are DEVICE pointers to
 * ntn_internal_dimension() floats in the library's internal layout (same small blocks, word
 * vectors word-major and padded; convert with the two functions below).  Runs on the handle's
 * stream; cost / n_active are host pointers, written after a stream synchronise (pass NULL for
 * both to skip the synchronise and read them later with ntn_last_cost). */
int ntn_eval_dev(ntn_handle* h, const float* d_theta, const uint8_t* corrupt_side,
                 float* d_grad, double* cost, int64_t* n_active);
int ntn_last_cost(ntn_handle* h, double* cost, int64_t* n_active);   /* synchronises */
/* DEVICE pointer to the last evaluation's 4 doubles {cost, n_active, data term, sum theta^2}: lets a
 * data-parallel host all-reduce the scalars on the device together with the gradient. */
double* ntn_result_dev(ntn_handle* h);
/* Data-parallel hosts: with on != 0 every ntn_eval_dev also writes 4 floats after the gradient, at
 * d_grad[ntn_internal_dimension() .. +4) = {cost_hi, cost_lo, n_active >> 12, n_active & 4095} (float pairs: exact on one
 * rank; after the FP32 sum over G ranks the count is still exact and the cost is good to ~1e-7 relative),
 * so ONE sum all-reduce of Pint+4 floats carries the gradient and the scalars.  d_grad must then have Pint+4 floats. */
int ntn_set_grad_tail(ntn_handle* h, int on);
/* Data-parallel hosts, gradient exchange INSIDE the evaluation (csrc/exchange.cu; NTN.java:411-430 is linear, so the
 * per-rank gradients of the column shards add up).  The gradient buffer is a symmetric allocation: the same numel floats
 * on every GPU of the box, every rank's copy mapped into this process (buffer_ptrs[world], this rank's own copy at
 * [rank]), a signal pad of >= 8 KiB of zero-initialised 32-bit flags per rank (signal_pad_ptrs[world]) and, optionally,
 * the NVSwitch multicast alias of the buffer (0: none).  After the bind -- and ntn_set_grad_tail(h, 1) -- every
 * ntn_eval_dev whose d_grad is buffer_ptrs[rank] ends with the SUM over the ranks of gradient and {cost, n_active} tail
 * in every rank's buffer: the word pull is cut into chunks and the library's own peer-memory kernels all-reduce chunk c
 * while chunk c+1 is pulled.  scratch_ptrs (optional, NULL: none) is a second symmetric allocation of
 * >= ntn_exchange_scratch_numel() floats per rank; with it the exchange runs in PUSH mode: the word-pull kernel stores
 * every 16-byte group another rank owns straight into that rank's scratch while it computes it, and the owner only
 * reads local memory (no peer loads).  All ranks must evaluate in lock-step (same sequence of calls).
 * ntn_exchange_check returns 1 when a peer did not show up within 2 s (the evaluation's results are then invalid). */
int ntn_exchange_bind(ntn_handle* h, int32_t rank, int32_t world, const uint64_t* buffer_ptrs,
                      const uint64_t* signal_pad_ptrs, uint64_t multicast_ptr, int64_t numel,
                      const uint64_t* scratch_ptrs, int64_t scratch_numel);
int64_t ntn_exchange_scratch_numel(const ntn_handle* h);
/* Entity-gradient exchange (after ntn_exchange_bind): bind a symmetric allocation of >= Ne * ceil4(d) floats as THE entity-
 * gradient buffer of the handle (+ a scratch of >= ntn_exchange_ge_scratch_numel() floats in push mode).  The sum over
 * the ranks is then taken on the entity gradient (Util.java:74-98's product, NTN.java:372,388) while the entity pull is
 * still running; the word pull that follows (NTN.java:411-430, linear in the entity gradient) yields the full-batch word
 * gradient on every rank and only the W/V/b/U blocks and the tail are exchanged after it.  Same volume over NVLink as
 * exchanging the word gradient, started one kernel earlier. */
int ntn_exchange_bind_ge(ntn_handle* h, const uint64_t* buffer_ptrs, const uint64_t* signal_pad_ptrs, uint64_t multicast_ptr,
                         int64_t numel, const uint64_t* scratch_ptrs, int64_t scratch_numel);
int64_t ntn_exchange_ge_scratch_numel(const ntn_handle* h);
int ntn_exchange_unbind(ntn_handle* h);
int ntn_exchange_check(ntn_handle* h);
/* measurement hook: iters back-to-back all-reduce kernels over floats [off, off+n) of a bound buffer (same call on every rank) */
int ntn_exchange_probe(ntn_handle* h, float* d_grad, int64_t off, int64_t n, int32_t last, int32_t iters);
const char* ntn_exchange_mode(const ntn_handle* h);   /* "none", "push", "p2p" (peer loads/stores) or "multimem" (NVLS) */

int64_t ntn_internal_dimension(const ntn_handle* h);
/* reference-order FP32 device vector (P) <-> internal-order device vector */
int ntn_to_internal_dev(ntn_handle* h, const float* d_ref, float* d_int);
int ntn_from_internal_dev(ntn_handle* h, const float* d_int, float* d_ref);
/* host double[P] (reference order) <-> internal-order device vector */
int ntn_upload_theta(ntn_handle* h, const double* theta, float* d_int);
int ntn_download_theta(ntn_handle* h, const float* d_int, double* theta);

/* Run on the caller's CUDA stream (a cudaStream_t cast to void*; as in CUDA, NULL is the legacy
 * default stream).  ntn_reset_stream goes back to the handle's own non-blocking stream. */
int ntn_set_stream(ntn_handle* h, void* cuda_stream);
int ntn_reset_stream(ntn_handle* h);
void* ntn_get_stream(ntn_handle* h);

/* ---- optimiser: device-resident L-BFGS (Run_NTN.java:119-124) ----------------------------- */
/* The reference's minimiser (umass-nlp LBFGSMinimizer) is an un-vendored jar; the algorithm used here is
 * defined in csrc/lbfgs.cu / DESIGN.md (parity unpinned).  theta stays in HBM (internal layout); every trial
 * point is one ntn_eval_dev on the current batch job; sides_fn is called once per evaluation to draw that
 * evaluation's corrupt sides (the reference draws Math.random() inside computeAt, NTN.java:146). */
typedef void (*ntn_sides_fn)(void* ctx, uint8_t* corrupt_side /* R, to fill */);
/* Data parallel: called after every evaluation with the DEVICE gradient of this rank's shard followed by the 4-float
 * {cost, n_active} tail (see ntn_set_grad_tail); must SUM all-reduce those n floats in place across the ranks (on the
 * handle's stream or synchronised with it).  Every rank then continues with identical numbers. */
typedef void (*ntn_reduce_fn)(void* ctx, float* d_grad_and_tail, int64_t n);
typedef struct ntn_lbfgs_opts {
    int32_t max_iters;      /* LBFGSMinimizer.Opts.maxIters; Run_NTN sets 5 (batch_iterations)   */
    int32_t history;        /* curvature pairs kept (0 = default 6)                              */
    double  tol;            /* relative function-change tolerance (0 = default 1e-4)             */
    ntn_reduce_fn reduce_fn;   /* NULL for a single process                                      */
    void*   reduce_ctx;
} ntn_lbfgs_opts;
typedef struct ntn_lbfgs_result {
    double  min_obj_val;    /* IOptimizer.Result.minObjVal */
    double  first_cost;     /* cost at the starting point   */
    int32_t iters, evals, did_converge;
    int64_t n_active_last;
} ntn_lbfgs_result;
int ntn_lbfgs_minimize(ntn_handle* h, float* d_theta /* device, internal layout, in/out */,
                       const ntn_lbfgs_opts* opts, ntn_sides_fn sides_fn, void* ctx, ntn_lbfgs_result* result);

/* ---- eval-side scoring ("next" row N1) -------------------------------------------------- */
/* Scores of n triples (NTN.computeBestThresholds / getPrediction, NTN.java:587-712). */
int ntn_score(ntn_handle* h, const double* theta, int64_t n,
              const int32_t* e1, const int32_t* rel, const int32_t* e2,
              int32_t mode /* NTN_SCORE_* */, double* scores);

/* ---- introspection --------------------------------------------------------------------- */
/* Kernel launches issued by the last ntn_eval* call, and which contraction path it took. */
int ntn_last_launch_count(const ntn_handle* h);
int ntn_active_gemm_path(const ntn_handle* h);
/* Mean per-phase device times (ms, CUDA events on the handle's stream) over the evaluations issued
 * since ntn_set_profiling(h, 1), which also resets the accumulators.
 * names: entity_build, w_prep, gemm_fwd, score, gemm_dw, gemm_ge, entity_pull, word_pull, finalize */
int ntn_set_profiling(ntn_handle* h, int on);
int ntn_last_phase_ms(ntn_handle* h, float* ms /* NTN_NUM_PHASES */);
#define NTN_NUM_PHASES 9
const char* ntn_phase_name(int i);

/* Test hook: copy an FP32 intermediate of the last evaluation ("T", "M", "GA", "dWpart", "E", "gE")
 * to the host; returns its element count (call with out == NULL to size the buffer), -1 if unknown. */
int64_t ntn_debug_fetch(ntn_handle* h, const char* name, float* out, int64_t cap);

/* ---- C++ host layer test hooks (csrc/host/ntn_host.hpp; no GPU needed) ----------------------- */
/* java.util.Random: n draws of nextInt(bound) (bound == 0: nextInt()), then m draws of nextDouble(). */
int ntn_host_java_random(int64_t seed, int32_t bound, int64_t n, int32_t* ints, int64_t m, double* doubles);
/* n consecutive java.util.Random.nextInt(bound) draws continuing from the 48-bit state *state (updated). */
int ntn_host_random_ints(uint64_t* state, int32_t bound, int64_t n, int32_t* out);
/* A native ntn_sides_fn (for ntn_lbfgs_minimize): ctx = uint64_t[2 + R] {java.util.Random state, R, presence flags};
 * one nextDouble() > 0.5 per present relation (NTN.java:146); the state is advanced in place. */
void ntn_host_sides_draw(void* ctx, uint8_t* corrupt_side);
/* DataFactory name parsing: '\n'-separated entity names and vocabulary -> forward / backward CSRs. */
int ntn_host_entity_word_maps(const char* names, const char* words, int32_t* fwd_off, int32_t* fwd_idx, int32_t* bwd_off,
                              int32_t* bwd_idx, int32_t* len_bwd, int64_t cap, int64_t* nf, int64_t* nb);
/* DataFactory.generateNewTrainingBatchJob (DataFactory.java:116-136), columns of the `calls`-th batch job. */
int ntn_host_generate_batch(const int32_t* train, int64_t ntrain, int32_t batch, int32_t corrupt, int32_t num_entities,
                            int64_t seed, int32_t calls, int32_t* e1, int32_t* rel, int32_t* e2, int32_t* e3);

/* java.util.Collections.shuffle(list, new Random(seed)) applied to [0, n) (the shuffle commented out at
 * DataFactory.java:118). */
int ntn_host_collections_shuffle(int64_t seed, int32_t n, int32_t* perm);
/* The invariant division of the flat thread index used by the gather kernels (q[i] = n[i] / d for every 32-bit n). */
int ntn_host_fastdiv(uint32_t d, int64_t count, const uint32_t* n, uint32_t* q);
/* chunk layout of the word pull (data-parallel exchange): out[4c..4c+3] = {first word, words, first partial slot, blocks};
 * returns the total number of regulariser-partial slots (> 0), or a negative NTN_E* code */
int ntn_host_word_chunks(int32_t Nw, int32_t dq, int32_t C, int32_t* out);
/* The staging conversions of the host-buffer entry points (double[] theta -> float, float gradient -> double[];
 * Util.java:49-55 rounding), exposed so that the CPU tests can check them on unaligned and ragged ranges. */
int ntn_host_convert_f64_f32(const double* src, float* dst, int64_t n);
int ntn_host_convert_f32_f64(const float* src, double* dst, int64_t n);

/* Theta checkpoints of the C++ host (Run_NTN.java:133-139; resume :108): write n values to `path` as the reference's
 * comma-separated text (binary == 0) or as the binary fast path ("NTNTHETA" | u32 1 | u32 32 | i64 P | P x f32), then read
 * them back -- every FP32 value must round-trip bit for bit. */
int ntn_host_theta_checkpoint_roundtrip(const char* path, int32_t binary, const double* theta, int64_t n, double* back);

const char* ntn_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NTN_B200_H */
