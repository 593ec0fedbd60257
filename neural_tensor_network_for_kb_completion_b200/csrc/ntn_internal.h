// Internal definitions shared by the C-ABI translation unit and the kernel translation units.
// Not part of the public boundary (that is include/ntn_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/ntn_b200.h"

// ------------------------------------------------------------------------------------------
// Shapes.  The CUDA path is built on the de-duplicated, augmented formulation:
//   * entity rows carry a constant 1 at index d, so the V and b terms of NTN.java:182-197 ride
//     along in the same contraction as the bilinear term of NTN.java:163-179;
//   * columns that share (e1, rel, e2) share one row of T' (the C corrupt copies of a training
//     triple, DataFactory.java:123-129).
//   dq  = d rounded up to 4       pitch of one slice inside a T'/M row, and of a word row
//   Kp  = d+1 rounded up to 8     pitch of an entity row (index d holds 1.0f)
//   Np  = k*dq+k rounded up to 16 pitch of a T'/M row: k slices of dq, then k "beta"/coef-sum slots
// ------------------------------------------------------------------------------------------
// Division of the flat thread index by the run-time row width (threads per row), Granlund-Montgomery style: the
// generic 64-bit division costs ~25 instructions in every thread of the gather kernels, which are issue-bound.
// Exact for every 32-bit n; the host checks that the grid fits 32 bits.
struct FastDiv {
    unsigned d, m, s;                                              // divisor, magic, shift (= ceil(log2 d) - 1, or 0)
};
static inline FastDiv make_fastdiv(unsigned d) {
    FastDiv f{d, 0u, 0u};
    if (d <= 1) return f;                                          // handled in fd_div
    unsigned l = 0; while ((1ull << l) < d) ++l;
    f.m = (unsigned)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    f.s = l - 1;
    return f;
}
__host__ __device__ __forceinline__ unsigned fd_div(unsigned n, const FastDiv& f) {
    if (f.d <= 1) return n;
#ifdef __CUDA_ARCH__
    const unsigned t = __umulhi(f.m, n);
#else
    const unsigned t = (unsigned)(((unsigned long long)f.m * n) >> 32);
#endif
    return (t + ((n - t) >> 1)) >> f.s;
}
// x = hi + lo with hi, lo exactly representable in TF32 (10 explicit mantissa bits): the operand split of the 3xTF32
// contractions (gemm_tcgen05.cu).  Rounding with full-rate integer ops (add half an ulp of the 13 dropped bits, mask):
// cvt.rna.tf32.f32 goes through the slow conversion pipe.  Shared with the kernels that PRODUCE a contraction operand
// already split (k_score2 -> M_hi / M_lo, k_tc_anchor_rows), so that the operand can be fed by TMA untouched.
#ifdef __CUDACC__
__device__ __forceinline__ float tf32_round(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void split_tf32(const float4& x, float4& hi, float4& lo) {
    hi.x = tf32_round(x.x); hi.y = tf32_round(x.y); hi.z = tf32_round(x.z); hi.w = tf32_round(x.w);
    lo.x = tf32_round(x.x - hi.x); lo.y = tf32_round(x.y - hi.y); lo.z = tf32_round(x.z - hi.z); lo.w = tf32_round(x.w - hi.w);
}
#endif

struct NtnDims {
    int d, k, R, Ne, Nw;
    int dq, Kp, Np;
    int64_t P;                   // reference dimension
    int64_t oW, oV, ob, oU;      // block offsets (same in both layouts)
    int64_t oWVref;              // reference layout: word block d x Nw row-major
    int64_t oWVint;              // internal layout: word block Nw x dq row-major (16-byte aligned)
    int64_t Pint;
};

#define NTN_TILE_ROWS   128      // unique triples per contraction tile (one relation per tile)
#define NTN_CHUNK_ROWS  256      // unique triples per dW split-K chunk (one relation per chunk)
#define NTN_SCORE_ROWS  32       // default unique triples per score tile (8 warps x 4); finish_batch picks per batch job
#define NTN_HUB_SEG     32       // incident-list segment length for hub entities
#define NTN_WORD_HUB_SEG 16      // ... and for hub words (their branch is a short, latency-bound tail: more, shorter segments)
#define NTN_ORDER_CLAMP 1023    // incident counts are clamped here when the pull's visiting order is sorted
#define NTN_RING        8        // in-flight evaluations the host may enqueue without synchronising
#define NTN_MAX_PEERS   8        // GPUs of one box (gradient exchange over peer memory, exchange.cu)
#define NTN_MAX_WORD_CHUNKS 8    // the word pull is cut into chunks so that the exchange of chunk c overlaps the pull of c+1
#define NTN_STAMP_SLOTS (NTN_NUM_PHASES + NTN_MAX_WORD_CHUNKS + 1)   // timeline stamp pairs: kernel groups, then exchange kernels
#define NTN_XCHG_MAX_BLOCKS 64   // blocks of one exchange kernel (each pairs with the same block index on every peer)
#define NTN_XCHG_SIG_WORD0 1024  // first 32-bit word of the signal pad used by the exchange kernels (torch's own
                                 // symmetric-memory kernels use the words below)

struct DeviceBatch {
    int64_t N = 0;               // columns
    int Nu = 0;                  // unique (rel, e1, e2)
    int ntiles = 0, nchunks = 0, nstiles = 0;
    int n_inc = 0;               // incidents = 2*Nu + N
    int n_ent_hub_segs = 0;
    bool use_gv = false;         // large batch: k_score materialises one contribution row per incident (GV)
    // unique triples, sorted by (rel, e1, e2)
    int *u_e1 = nullptr, *u_e2 = nullptr, *u_rel = nullptr;
    int *c_off = nullptr;        // Nu+1: corrupt columns of each unique triple
    int *c_e3 = nullptr;         // N
    int *t_rel = nullptr, *t_u0 = nullptr, *t_n = nullptr;        // contraction tiles
    int *g_rel = nullptr, *g_u0 = nullptr, *g_n = nullptr;        // tiles of the swapped-role entity-gradient contraction (<= 160 rows)
    int ngtiles = 0;
    int *st_rel = nullptr, *st_u0 = nullptr, *st_n = nullptr;     // score tiles
    int *rels_off = nullptr;     // R+1: score-tile range per relation
    int *ch_rel = nullptr, *ch_u0 = nullptr, *ch_n = nullptr;     // dW chunks
    int *relch_off = nullptr;    // R+1: chunk range per relation
    int *reltile_off = nullptr;  // R+1: tile range per relation
    // entity -> incidents (pull lists), sorted by (entity, code, u)
    int *inc_off = nullptr;      // Ne+1
    int4 *inc = nullptr;         // {unique triple, coefficient row, role (0 e1, 1 e2, 2 e3), relation}
    // hub entities (incident list longer than NTN_HUB_SEG): segments + partial rows
    int *eh_ent = nullptr, *eh_begin = nullptr, *eh_end = nullptr;   // per segment
    int *eh_first = nullptr;     // Ne: first segment id of a hub entity, -1 otherwise
    int *eh_count = nullptr;     // Ne: number of segments
    // per-evaluation intermediates
    float *T = nullptr;          // Nu x Np   T' = [E_anchor | 1] * Waug
    float *M = nullptr;          // Nu x Np   backward coefficients folded with the gathered rows (m_split: their TF32 hi parts)
    float *M_lo = nullptr;       // Nu x Np   m_split: the TF32 lo parts (M = M + M_lo), written by the score kernels
    bool m_split = false;        // the tcgen05 contractions take M pre-split (TMA-fed A operand of the entity-gradient contraction)
    float *GA = nullptr;         // Nu x Kp   anchor-entity gradient rows
    float *GV = nullptr;         // (Nu + N) x dq   per-incident contribution rows (non-anchor roles), written by k_score
    float *crec = nullptr;       // (Nu + N) x RS packed coefficient records (k_score2 -> k_entity_pull2), RS = ceil4(k + 1)
    int RS = 4;
    int4 *pull_rng = nullptr;    // entity ranges of the warp-per-range entity pull: {first entity, end entity, first incident, end incident}
    int *ent_order = nullptr;    // Ne: entities in descending order of their incident count (stable): the pull's visiting order
    float *acoef = nullptr;      // (Nu + N) x k : a_{u,s} rows, then c_{col,s} rows (one allocation; GV variant)
    float *ccoef = nullptr;      // = acoef + Nu*k
    float *dWpart = nullptr;     // nchunks x Kp x Np
    double *tp_cost = nullptr;   // ntiles
    int *tp_nact = nullptr;      // ntiles
    float *tp_dU = nullptr;      // ntiles x k
    float *eh_part = nullptr;    // n_ent_hub_segs x dq
};

struct DeviceWords {
    bool set = false;
    int nnz_f = 0, nnz_b = 0;
    int *fwd_off = nullptr, *fwd_idx = nullptr;    // entity -> forward words
    int4 *fwd_rec = nullptr;                       // per entity {first index, word count, word 0, word 1}
    int *w_off = nullptr, *w_ent = nullptr;        // word -> entities of the backward lists (inverse CSR)
    int4 *w_rec = nullptr;                         // per word {j0, j1, first entity, hub segment count}
    float *len_bwd = nullptr;                      // Ne: 1 / entityLength(n) as float (the kernels multiply)
    int n_hub_segs = 0, n_hub_words = 0;
    int *hub_words = nullptr;                      // ids of the hub words
    int *wh_word = nullptr, *wh_begin = nullptr, *wh_end = nullptr;
    int *wh_first = nullptr, *wh_count = nullptr;  // Nw
    float *wh_part = nullptr;                      // n_hub_segs x dq
};

// One symmetric gradient buffer bound for the in-evaluation exchange (exchange.cu): the same allocation exists on every
// rank of the box; peer[r] is rank r's copy mapped into this process, mc the NVSwitch multicast alias (or null).
struct ExchangeBinding {
    float* local = nullptr;
    float* peer[NTN_MAX_PEERS] = {};
    float* mc = nullptr;
    uint32_t* sig[NTN_MAX_PEERS] = {};   // signal pads (32-bit flags), same indexing as peer[]
    float* scratch[NTN_MAX_PEERS] = {};  // second symmetric allocation (push mode): rank r's copy receives the other ranks'
    int64_t numel = 0;                   // contributions to the slices r owns
    int64_t scratch_numel = 0;
    uint32_t* epoch = nullptr;           // two words of this rank's signal pad: {barrier epochs completed, blocks that have left the running kernel}
};

// Push mode: where the producer of a gradient range (the word pull) also stores the 16-byte groups that another rank
// owns -- straight into that rank's scratch, over NVLink.  A range [off, off + 4 n4) is cut into `world` slices of per4
// groups; rank r owns slice r.  dst[r] already points at this rank's row of r's scratch for this range.
struct XPushRange {
    float* dst[NTN_MAX_PEERS];
    int64_t off;                         // first float of the range in the gradient vector
    unsigned per4, per_m, per_s;         // slice length in float4 groups and its invariant-division magic
    int rank, world;                     // world == 0: nothing to push
};
#define NTN_XCHG_RANGE_PAD 64            // floats between the scratch regions of consecutive ranges (slices are rounded up)

struct ntn_handle {
    ntn_config cfg;
    NtnDims dm;
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t stream2 = nullptr;          // forked branch: hub words, joined inside the word pull
    cudaStream_t stream3 = nullptr;          // forked branch: w_prep, then gemm_dw -> finalize_small
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_f0 = nullptr, ev_j0 = nullptr, ev_f1 = nullptr, ev_j1 = nullptr;
    std::string err;
    std::unordered_map<void*, size_t> capacity;   // bytes behind the device buffers that dev_alloc may reuse
    DeviceWords w;
    DeviceBatch b;
    float *wv_frozen = nullptr;   // Nw x dq (internal word-major), or null
    bool E_frozen_valid = false;  // NTN_WV_FROZEN: E was built from wv_frozen and is reused (invariant across evaluations)
    float *E = nullptr;           // Ne x Kp
    float *gE = nullptr;          // Ne x dq
    float *Waug = nullptr;        // R x Kp x Np   (row p, column n)
    uint8_t *side = nullptr;      // R (device)
    uint8_t *h_side = nullptr;    // NTN_RING x R (pinned ring, so evaluations can be enqueued back to back)
    cudaEvent_t side_ev[NTN_RING] = {};
    int ring_pos = 0;
    double *reg_part = nullptr;   // partial sums of theta^2
    int n_reg_part = 0;
    // results
    double *d_out = nullptr;      // [0]=cost, [1]=n_active (as double)
    double *h_out = nullptr;      // pinned mirror
    // host-path staging
    float *d_ref_theta = nullptr, *d_ref_grad = nullptr;   // P (reference order)
    float *d_int_theta = nullptr, *d_int_grad = nullptr;   // Pint
    float *h_stage = nullptr;     // pinned, P floats
    std::vector<cudaEvent_t> xfer_ev;   // per-chunk D2H completion events
    int launches = 0;
    int gemm_path_active = NTN_GEMM_SIMT;
    bool profiling = false;
    cudaEvent_t ev[NTN_RING][NTN_NUM_PHASES][2] = {};   // begin/end of each phase, on the stream it runs on
    bool ev_pending[NTN_RING] = {};
    double phase_acc[NTN_NUM_PHASES] = {};
    int phase_n = 0;
    unsigned long long *d_stamps = nullptr;   // NTN_GRAPH_TIMELINE: %globaltimer at begin/end of every kernel group, then of
                                              // every exchange kernel (NTN_STAMP_SLOTS pairs in all)
    void *tc = nullptr;           // tcgen05 path state (gemm_tcgen05.cu)
    void *bi_ws = nullptr;        // batch-index workspace (batch_index.cu), grow-only
    size_t bi_ws_bytes = 0;
    int *h_meta = nullptr;        // pinned: counters read back by the batch index
    int *h_cols = nullptr;        // pinned staging of the batch columns (ntn_set_batch with host pointers)
    int *d_cols = nullptr;        // their device copy, 4 x n
    size_t cols_cap = 0;
    int64_t cols_n = -1;          // columns currently staged in d_cols (-1: none)
    int *d_train = nullptr;       // training triples [e1 | rel | e2] (ntn_set_training_triples), for the device batch generator
    int64_t n_train = 0;
    uint64_t *d_true_keys = nullptr;   // their packed keys, sorted (filtered corrupt sampling)
    void *lbfgs_ws = nullptr;     // L-BFGS workspace (lbfgs.cu), grow-only
    size_t lbfgs_ws_bytes = 0;
    double* lbfgs_scal = nullptr; // pinned: the optimiser's scalars (g.dir, first step, history state)
    // CUDA graphs of the evaluation, one per (theta, grad, stream) triple
    struct GraphEntry { const float* theta; float* grad; cudaStream_t stream; cudaGraphExec_t exec; int launches; };
    std::vector<GraphEntry> graphs;
    bool use_graph = true;
    bool grad_tail = false;       // the caller's gradient buffer has 4 extra floats for {cost, n_active}
    // in-evaluation gradient exchange over peer memory (exchange.cu)
    int x_rank = 0, x_world = 1;
    unsigned long long x_timeout_ns = 2000000000ull;   // bound of every flag wait (NTN_EXCHANGE_TIMEOUT_MS)
    int x_mcst = 0;                               // push mode: broadcast through the multicast alias
    int x_stamp = 0;                              // exchange kernels launched in the evaluation being enqueued (timeline slots)
    int x_chunks = 2, x_blocks = 64, x_mode = 0;   // word-pull chunks, blocks per exchange kernel, 0: peer loads/stores, 1: multimem, 2: push
    std::vector<ExchangeBinding> xb;
    // entity-gradient exchange (ntn_exchange_bind_ge): gE itself is a symmetric allocation and is summed over the ranks
    // while the entity pull is still running; the word pull then works on the full-batch gE on every rank and only the
    // small W/V/b/U blocks and the tail are left to exchange after it
    ExchangeBinding xge;
    bool xge_set = false;
    float* gE_own = nullptr;                      // the handle's own gE while a symmetric one is bound
    int xge_chunks = 2;
    bool ge_mode = false;                         // the evaluation being enqueued exchanges gE (the word block then takes the full lambda)
    cudaStream_t stream4 = nullptr;               // exchange of the finished chunks, beside the word pull (highest priority)
    cudaEvent_t ev_x[NTN_MAX_WORD_CHUNKS] = {}, ev_xdone = nullptr;
    unsigned* d_xerr = nullptr;                   // set when a bounded flag wait expired (results invalid)
    int word_chunks = 1;                          // chunks of the word pull in the evaluation being enqueued
};

// word-pull chunk c of C: words [w0, w0 + wn) and the first of its blocks' regulariser-partial slots
void ntn_word_chunk(const NtnDims& dm, int c, int C, int* w0, int* wn, int* block_off, int* blocks);
int ntn_word_blocks_total(const NtnDims& dm, int C);
// exchange.cu: all-reduce (SUM) of up to two float ranges of the bound buffer across the ranks, in place
struct XRange { int64_t off, n; int id; bool pushed; };   // floats; off and n are multiples of 4; id: ordinal of the range in
                                                          // the evaluation (its scratch region); pushed: the producer already
                                                          // stored the other ranks' slices into their scratch
void ntn_launch_exchange(ntn_handle* h, const ExchangeBinding* xb, XRange r0, XRange r1, XRange r2, bool last, cudaStream_t st);
XPushRange ntn_push_range(const ntn_handle* h, const ExchangeBinding* xb, int64_t off, int64_t n, int id);
const ExchangeBinding* ntn_find_binding(const ntn_handle* h, const float* d_grad);

// ---- batch indexing on the device (batch_index.cu)
int ntn_index_batch_dev(ntn_handle* h, int64_t n, const int* d_e1, const int* d_rel, const int* d_e2, const int* d_e3,
                        int* counts /* Nu, hub segments */, int* rel_count /* R */);

int ntn_build_true_keys_dev(ntn_handle* h, int64_t T, const int* d_train, uint64_t** d_keys);
int ntn_generate_columns_dev(ntn_handle* h, const int* d_train, int64_t T, int64_t first, int batch, int corrupt, uint64_t seed,
                             uint64_t job, const uint64_t* d_true_keys, int64_t n_true, int* d_cols);

// ---- kernel launchers (ntn_kernels.cu) ---------------------------------------------------
void ntn_launch_entity_build(ntn_handle* h, const float* wvt);
void ntn_launch_w_prep(ntn_handle* h, const float* theta, cudaStream_t st);
void ntn_launch_gemm_fwd_simt(ntn_handle* h);
void ntn_launch_score(ntn_handle* h, const float* theta);
int ntn_score2_rows(int k);   // tile height of the default score kernel
void ntn_launch_gemm_dw_simt(ntn_handle* h, cudaStream_t st);
void ntn_launch_gemm_ge_simt(ntn_handle* h);
void ntn_stamp(ntn_handle* h, int slot, cudaStream_t st);   // NTN_GRAPH_TIMELINE: store %globaltimer into stamp slot (no-op otherwise)
void ntn_launch_entity_pull(ntn_handle* h, const ExchangeBinding* xge);
int ntn_pull_ranges(const ntn_handle* h);
void ntn_launch_pull_ranges(ntn_handle* h);
void ntn_launch_word_pull(ntn_handle* h, const float* theta, float* grad, const ExchangeBinding* xb);
void ntn_launch_finalize_small(ntn_handle* h, const float* theta, float* grad, cudaStream_t st);
void ntn_launch_final_reduce(ntn_handle* h, float* tail);
void ntn_launch_to_internal(ntn_handle* h, const float* d_ref, float* d_int);
void ntn_launch_from_internal(ntn_handle* h, const float* d_int, float* d_ref);
void ntn_launch_transpose_wv_host_layout(ntn_handle* h, const float* d_wv_ref /* d x Nw */, float* d_wvt);
void ntn_launch_score_eval(ntn_handle* h, const float* theta, const int* e1, const int* rel, const int* e2,
                           int64_t n, int mode, double* out);
