// C-ABI entry points (include/ntn_b200.h): handle lifetime, DataFactory state upload and batch
// indexing, and the evaluation driver that strings the kernels together.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <numeric>
#include <thread>

#include "ntn_internal.h"

int ntn_small_blocks(const NtnDims& dm);                       // ntn_kernels.cu
int ntn_word_blocks(const NtnDims& dm);
int ntn_tc_supported(const ntn_handle* h);                     // gemm_tcgen05.cu
int ntn_tc_prepare_batch(ntn_handle* h);
void ntn_tc_release(ntn_handle* h);
void ntn_drop_graphs(ntn_handle* h);
int ntn_tc_w_prep(ntn_handle* h, const float* theta, cudaStream_t st);
int ntn_tc_gemm_fwd(ntn_handle* h);
int ntn_tc_gemm_dw(ntn_handle* h, cudaStream_t st);
int ntn_tc_gemm_ge(ntn_handle* h);
int ntn_tc_check_error(ntn_handle* h);
int64_t ntn_tc_fetch_timeline(ntn_handle* h, float* out, int64_t cap);

static thread_local std::string g_create_err;

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return NTN_ECUDA;                                                                      \
        }                                                                                          \
    } while (0)

#define FAIL(code, msg) do { h->err = (msg); return (code); } while (0)

// (re)allocate a device buffer; an existing allocation is kept when it is large enough -- ntn_set_batch runs once per
// batch job and its ~35 cudaMalloc/cudaFree pairs were a visible part of its cost
template <typename T> static int dev_alloc(ntn_handle* h, T** p, size_t n) {
    if (n == 0) n = 1;
    const size_t bytes = n * sizeof(T);
    if (*p) {
        auto it = h->capacity.find((void*)*p);
        if (it != h->capacity.end() && it->second >= bytes) return NTN_OK;
        if (it != h->capacity.end()) h->capacity.erase(it);
        cudaFree(*p); *p = nullptr;
    }
    CK(cudaMalloc((void**)p, bytes));
    h->capacity[(void*)*p] = bytes;
    return NTN_OK;
}
template <typename T> static int dev_upload(ntn_handle* h, T** p, const std::vector<T>& v) {
    int rc = dev_alloc(h, p, v.size());
    if (rc) return rc;
    if (!v.empty()) CK(cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return NTN_OK;
}
template <typename T> static void dev_free(T*& p) { if (p) { cudaFree(p); p = nullptr; } }
template <typename T> static void dev_free(ntn_handle* h, T*& p) { if (p) { h->capacity.erase((void*)p); cudaFree(p); p = nullptr; } }

static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

static NtnDims make_dims(const ntn_config& c) {
    NtnDims m{};
    m.d = c.d; m.k = c.k; m.R = c.R; m.Ne = c.Ne; m.Nw = c.Nw;
    m.dq = (int)round_up(c.d, 4);
    m.Kp = (int)round_up(c.d + 1, 8);
    m.Np = (int)round_up((int64_t)c.k * m.dq + c.k, 16);
    int64_t d = c.d, k = c.k, R = c.R;
    m.oW = 0;
    m.oV = m.oW + R * k * d * d;
    m.ob = m.oV + R * 2 * d * k;
    m.oU = m.ob + R * k;
    m.oWVref = m.oU + R * k;
    m.P = m.oWVref + d * (int64_t)c.Nw;
    m.oWVint = round_up(m.oWVref, 4);
    m.Pint = m.oWVint + (int64_t)c.Nw * m.dq;
    return m;
}

extern "C" const char* ntn_version(void) { return "ntn_b200 0.1 (sm_100a)"; }

extern "C" const char* ntn_last_error(const ntn_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

static const char* kPhaseNames[NTN_NUM_PHASES] = {"entity_build", "w_prep", "gemm_fwd", "score", "gemm_dw",
                                                  "gemm_ge", "entity_pull", "word_pull", "finalize"};
extern "C" const char* ntn_phase_name(int i) { return (i >= 0 && i < NTN_NUM_PHASES) ? kPhaseNames[i] : ""; }

extern "C" int ntn_create(const ntn_config* cfg, ntn_handle** out) {
    if (!cfg || !out) { g_create_err = "null argument"; return NTN_EINVAL; }
    *out = nullptr;
    if (cfg->d < 1 || cfg->d > 256 || cfg->k < 1 || cfg->k > 8) {
        g_create_err = "unsupported shape: need 1 <= d <= 256 and 1 <= k <= 8";
        return NTN_EUNSUP;
    }
    if (cfg->R < 1 || cfg->Ne < 1 || cfg->Nw < 1 || cfg->batch_divisor < 1) {
        g_create_err = "R, Ne, Nw and batch_divisor must be positive";
        return NTN_EINVAL;
    }
    if (cfg->activation != NTN_ACT_TANH && cfg->activation != NTN_ACT_SIGMOID) {
        g_create_err = "activation must be NTN_ACT_TANH or NTN_ACT_SIGMOID";
        return NTN_EINVAL;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e);
        return NTN_ECUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_err = "device ordinal out of range"; return NTN_EINVAL; }
    ntn_handle* h = new ntn_handle();
    h->cfg = *cfg;
    if (h->cfg.reg_scale == 0.f) h->cfg.reg_scale = 1.f;
    h->dm = make_dims(*cfg);
    h->device = cfg->device;
    h->use_graph = getenv("NTN_NO_GRAPH") == nullptr;
    if (getenv("NTN_GRAPH_TIMELINE")) {
        if (cudaMalloc((void**)&h->d_stamps, 2 * NTN_STAMP_SLOTS * sizeof(unsigned long long)) != cudaSuccess) h->d_stamps = nullptr;
        else cudaMemset(h->d_stamps, 0, 2 * NTN_STAMP_SLOTS * sizeof(unsigned long long));
    }
    auto bail = [&](int rc) { g_create_err = h->err; ntn_destroy(h); return rc; };
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return bail(NTN_ECUDA); } } while (0)
    CKC(cudaSetDevice(h->device));
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, h->device));
    h->sm_count = prop.multiProcessorCount;
    if (prop.major != 10) {
        h->err = "this library is built for sm_100a (B200) only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
        return bail(NTN_ECUDA);
    }
    // the main chain (forward -> score -> entity gradient -> pulls) is the critical path of an evaluation; the forked
    // branches (weight prep / dW contraction / finalize, hub words) only have to finish before the final reduce.  The
    // chain's stream gets the highest priority and the branches the lowest, so that free SM slots go to the chain first.
    int prio_lo = 0, prio_hi = 0;
    CKC(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));      // numerically: hi <= lo
    static const bool use_prio = [] { const char* e = getenv("NTN_STREAM_PRIORITY"); return !(e && e[0] == '0'); }();
    CKC(cudaStreamCreateWithPriority(&h->own_stream, cudaStreamNonBlocking, use_prio ? prio_hi : prio_lo));
    // (the hub-word branch: NTN_HUB_PRIORITY=1 gives it the chain's priority)
    static const bool hub_hi = [] { const char* e = getenv("NTN_HUB_PRIORITY"); return e && e[0] == '1'; }();
    CKC(cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, hub_hi ? prio_hi : prio_lo));
    CKC(cudaStreamCreateWithPriority(&h->stream3, cudaStreamNonBlocking, prio_lo));
    CKC(cudaEventCreateWithFlags(&h->ev_f0, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->ev_j0, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->ev_f1, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->ev_j1, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    h->stream = h->own_stream;
    const NtnDims& dm = h->dm;
    CKC(cudaMalloc((void**)&h->E, (size_t)dm.Ne * dm.Kp * sizeof(float)));
    CKC(cudaMalloc((void**)&h->gE, (size_t)dm.Ne * dm.dq * sizeof(float)));
    CKC(cudaMalloc((void**)&h->Waug, (size_t)dm.R * dm.Kp * dm.Np * sizeof(float)));
    CKC(cudaMalloc((void**)&h->side, dm.R));
    CKC(cudaMallocHost((void**)&h->h_side, (size_t)dm.R * NTN_RING));
    for (int i = 0; i < NTN_RING; ++i) CKC(cudaEventCreateWithFlags(&h->side_ev[i], cudaEventDisableTiming));
    h->n_reg_part = ntn_word_blocks(dm) + dm.Nw + ntn_small_blocks(dm) + dm.R * dm.k;   // word blocks | hub words (<= Nw) | small blocks | U
    CKC(cudaMalloc((void**)&h->reg_part, (size_t)h->n_reg_part * sizeof(double)));
    CKC(cudaMalloc((void**)&h->d_out, 4 * sizeof(double)));
    CKC(cudaMallocHost((void**)&h->h_out, 4 * sizeof(double)));
    for (int s = 0; s < NTN_RING; ++s)
        for (int i = 0; i < NTN_NUM_PHASES; ++i) { CKC(cudaEventCreate(&h->ev[s][i][0])); CKC(cudaEventCreate(&h->ev[s][i][1])); }
#undef CKC
    *out = h;
    return NTN_OK;
}

static void free_batch(DeviceBatch& b) {
    dev_free(b.u_e1); dev_free(b.u_e2); dev_free(b.u_rel); dev_free(b.c_off); dev_free(b.c_e3);
    dev_free(b.st_rel); dev_free(b.st_u0); dev_free(b.st_n); dev_free(b.rels_off);
    dev_free(b.t_rel); dev_free(b.t_u0); dev_free(b.t_n); dev_free(b.ch_rel); dev_free(b.ch_u0); dev_free(b.ch_n);
    dev_free(b.relch_off); dev_free(b.reltile_off); dev_free(b.inc_off); dev_free(b.inc);
    dev_free(b.eh_ent); dev_free(b.eh_begin); dev_free(b.eh_end); dev_free(b.eh_first); dev_free(b.eh_count);
    dev_free(b.T); dev_free(b.M); dev_free(b.M_lo); dev_free(b.GA); dev_free(b.g_rel); dev_free(b.g_u0); dev_free(b.g_n); dev_free(b.GV); dev_free(b.acoef); b.ccoef = nullptr; dev_free(b.crec); dev_free(b.ent_order); dev_free(b.pull_rng); dev_free(b.dWpart);
    dev_free(b.tp_cost); dev_free(b.tp_nact); dev_free(b.tp_dU); dev_free(b.eh_part);
    b = DeviceBatch();
}

extern "C" void ntn_destroy(ntn_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    ntn_drop_graphs(h);
    ntn_tc_release(h);
    free_batch(h->b);
    DeviceWords& w = h->w;
    dev_free(w.fwd_off); dev_free(w.fwd_idx); dev_free(w.w_off); dev_free(w.w_ent); dev_free(w.len_bwd);
    dev_free(w.wh_word); dev_free(w.wh_begin); dev_free(w.wh_end); dev_free(w.wh_first); dev_free(w.wh_count);
    dev_free(w.wh_part); dev_free(w.hub_words); dev_free(w.w_rec); dev_free(w.fwd_rec);
    if (h->gE_own) { h->gE = h->gE_own; h->gE_own = nullptr; }      // (a bound symmetric gE belongs to the caller)
    dev_free(h->wv_frozen); dev_free(h->E); dev_free(h->gE); dev_free(h->Waug); dev_free(h->side);
    dev_free(h->reg_part); dev_free(h->d_out); dev_free(h->d_ref_theta); dev_free(h->d_ref_grad);
    dev_free(h->d_int_theta); dev_free(h->d_int_grad);
    if (h->lbfgs_ws) cudaFree(h->lbfgs_ws);
    if (h->lbfgs_scal) cudaFreeHost(h->lbfgs_scal);
    if (h->d_stamps) cudaFree(h->d_stamps);
    if (h->bi_ws) cudaFree(h->bi_ws);
    if (h->h_meta) cudaFreeHost(h->h_meta);
    if (h->h_cols) cudaFreeHost(h->h_cols);
    if (h->d_cols) cudaFree(h->d_cols);
    if (h->d_train) cudaFree(h->d_train);
    if (h->d_true_keys) cudaFree(h->d_true_keys);
    if (h->h_side) cudaFreeHost(h->h_side);
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    for (cudaEvent_t e : h->xfer_ev) cudaEventDestroy(e);
    for (int s = 0; s < NTN_RING; ++s) {
        if (h->side_ev[s]) cudaEventDestroy(h->side_ev[s]);
        for (int i = 0; i < NTN_NUM_PHASES; ++i) for (int q = 0; q < 2; ++q) if (h->ev[s][i][q]) cudaEventDestroy(h->ev[s][i][q]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    for (cudaEvent_t e : {h->ev_f0, h->ev_j0, h->ev_f1, h->ev_j1}) if (e) cudaEventDestroy(e);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->stream3) cudaStreamDestroy(h->stream3);
    if (h->stream4) cudaStreamDestroy(h->stream4);
    for (int i = 0; i < NTN_MAX_WORD_CHUNKS; ++i) if (h->ev_x[i]) cudaEventDestroy(h->ev_x[i]);
    if (h->ev_xdone) cudaEventDestroy(h->ev_xdone);
    if (h->d_xerr) cudaFree(h->d_xerr);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

extern "C" int64_t ntn_dimension(const ntn_handle* h) { return h ? h->dm.P : -1; }
extern "C" int64_t ntn_internal_dimension(const ntn_handle* h) { return h ? h->dm.Pint : -1; }
extern "C" int ntn_last_launch_count(const ntn_handle* h) { return h ? h->launches : -1; }
extern "C" int ntn_active_gemm_path(const ntn_handle* h) { return h ? h->gemm_path_active : -1; }

extern "C" int ntn_set_stream(ntn_handle* h, void* s) {
    if (!h) return NTN_EINVAL;
    cudaStreamSynchronize(h->stream);
    h->stream = (cudaStream_t)s;
    return NTN_OK;
}
extern "C" int ntn_reset_stream(ntn_handle* h) {
    if (!h) return NTN_EINVAL;
    cudaStreamSynchronize(h->stream);
    h->stream = h->own_stream;
    return NTN_OK;
}
extern "C" void* ntn_get_stream(ntn_handle* h) { return h ? (void*)h->stream : nullptr; }
// fold the event pairs of ring slot s into the per-phase accumulators (waits for that evaluation)
static void drain_profile_slot(ntn_handle* h, int s) {
    if (!h->ev_pending[s]) return;
    for (int i = 0; i < NTN_NUM_PHASES; ++i) {
        float ms = 0.f;
        cudaEventSynchronize(h->ev[s][i][1]);
        if (cudaEventElapsedTime(&ms, h->ev[s][i][0], h->ev[s][i][1]) == cudaSuccess) h->phase_acc[i] += ms;
    }
    h->phase_n++;
    h->ev_pending[s] = false;
}
extern "C" int ntn_set_profiling(ntn_handle* h, int on) {
    if (!h) return NTN_EINVAL;
    for (int s = 0; s < NTN_RING; ++s) drain_profile_slot(h, s);
    h->profiling = on != 0;
    for (int i = 0; i < NTN_NUM_PHASES; ++i) h->phase_acc[i] = 0.0;
    h->phase_n = 0;
    return NTN_OK;
}

// Split the lists longer than NTN_HUB_SEG into segments (hub words like "unknown", hub entities).
static void build_hub_segments(const std::vector<int>& off, int nrows, std::vector<int>& seg_row,
                               std::vector<int>& seg_begin, std::vector<int>& seg_end, std::vector<int>& first,
                               std::vector<int>& count, int seg_len = NTN_HUB_SEG, int hub_threshold = NTN_HUB_SEG) {
    first.assign(nrows, -1);
    count.assign(nrows, 0);
    for (int r = 0; r < nrows; ++r) {
        int n = off[r + 1] - off[r];
        if (n <= hub_threshold) continue;
        first[r] = (int)seg_row.size();
        for (int b = off[r]; b < off[r + 1]; b += seg_len) {
            seg_row.push_back(r);
            seg_begin.push_back(b);
            seg_end.push_back(std::min(b + seg_len, off[r + 1]));
            count[r]++;
        }
    }
}

extern "C" int ntn_set_entity_words(ntn_handle* h, const int32_t* fwd_off, const int32_t* fwd_idx,
                                    const int32_t* bwd_off, const int32_t* bwd_idx, const int32_t* len_bwd) {
    if (!h) return NTN_EINVAL;
    if (!fwd_off || !fwd_idx || !len_bwd) FAIL(NTN_EINVAL, "null argument");
    CK(cudaSetDevice(h->device));
    const NtnDims& dm = h->dm;
    if (!bwd_off || !bwd_idx) { bwd_off = fwd_off; bwd_idx = fwd_idx; }
    if (fwd_off[0] != 0 || bwd_off[0] != 0) FAIL(NTN_EINVAL, "CSR offsets must start at 0");
    if ((int64_t)dm.Ne * (dm.Kp / 4) >= ((int64_t)1 << 32) || (int64_t)dm.Nw * (dm.dq / 4) >= ((int64_t)1 << 32))
        FAIL(NTN_EUNSUP, "entity or word table too large for the flat 32-bit thread index of the gather kernels");
    for (int n = 0; n < dm.Ne; ++n) {
        if (fwd_off[n + 1] <= fwd_off[n]) FAIL(NTN_EINVAL, "entity " + std::to_string(n) + " has no forward word");
        if (bwd_off[n + 1] < bwd_off[n]) FAIL(NTN_EINVAL, "backward CSR offsets not monotone");
        if (len_bwd[n] < 0)
            FAIL(NTN_EINVAL, "entityLength(" + std::to_string(n) + ") < 0");
    }
    int nf = fwd_off[dm.Ne], nb = bwd_off[dm.Ne];
    for (int j = 0; j < nf; ++j) if (fwd_idx[j] < 0 || fwd_idx[j] >= dm.Nw) FAIL(NTN_EINVAL, "forward word index out of range");
    for (int j = 0; j < nb; ++j) if (bwd_idx[j] < 0 || bwd_idx[j] >= dm.Nw) FAIL(NTN_EINVAL, "backward word index out of range");
    ntn_drop_graphs(h);
    DeviceWords& w = h->w;
    w.nnz_f = nf; w.nnz_b = nb;
    std::vector<int> vfo(fwd_off, fwd_off + dm.Ne + 1), vfi(fwd_idx, fwd_idx + nf);
    // inverse CSR word -> entities (stable counting sort: entities ascending, multiplicity kept),
    // the order NTN.java:414-426 visits them in.
    std::vector<int> woff(dm.Nw + 1, 0), went(nb);
    for (int j = 0; j < nb; ++j) woff[bwd_idx[j] + 1]++;
    for (int i = 0; i < dm.Nw; ++i) woff[i + 1] += woff[i];
    {
        std::vector<int> cur(woff.begin(), woff.end() - 1);
        for (int n = 0; n < dm.Ne; ++n)
            for (int j = bwd_off[n]; j < bwd_off[n + 1]; ++j) went[cur[bwd_idx[j]]++] = n;
    }
    std::vector<float> lb(dm.Ne);
    // the kernels multiply by 1/len (NTN.java:420 divides).  entityLength == 0 happens on the real Wordnet list (the
    // name "__2", DataFactory.java:650): the reference divides that entity's gradient column by zero; here the entity
    // simply contributes nothing to the word gradient (weight 0), which is what the guard at NTN.java:417 does for
    // every entity whose column is zero.
    for (int n = 0; n < dm.Ne; ++n) lb[n] = len_bwd[n] > 0 ? (float)(1.0 / (double)len_bwd[n]) : 0.f;
    std::vector<int> sr, sb, se, first, count;
    build_hub_segments(woff, dm.Nw, sr, sb, se, first, count, NTN_WORD_HUB_SEG, NTN_HUB_SEG);
    w.n_hub_segs = (int)sr.size();
    std::vector<int> hub_words;
    for (int i = 0; i < dm.Nw; ++i) if (count[i] > 0) hub_words.push_back(i);
    w.n_hub_words = (int)hub_words.size();
    std::vector<int4> frec(dm.Ne);
    for (int n = 0; n < dm.Ne; ++n) {
        int j0 = fwd_off[n], cnt = fwd_off[n + 1] - fwd_off[n];
        frec[n] = make_int4(j0, cnt, fwd_idx[j0], cnt > 1 ? fwd_idx[j0 + 1] : 0);
    }
    std::vector<int4> wrec(dm.Nw);
    for (int i = 0; i < dm.Nw; ++i)
        wrec[i] = make_int4(woff[i], woff[i + 1], woff[i + 1] > woff[i] ? went[woff[i]] : -1, count[i]);
    int rc;
    if ((rc = dev_upload(h, &w.fwd_off, vfo)) || (rc = dev_upload(h, &w.fwd_idx, vfi)) ||
        (rc = dev_upload(h, &w.w_off, woff)) || (rc = dev_upload(h, &w.w_ent, went)) ||
        (rc = dev_upload(h, &w.len_bwd, lb)) || (rc = dev_upload(h, &w.wh_word, sr)) ||
        (rc = dev_upload(h, &w.wh_begin, sb)) || (rc = dev_upload(h, &w.wh_end, se)) ||
        (rc = dev_upload(h, &w.wh_first, first)) || (rc = dev_upload(h, &w.wh_count, count)) ||
        (rc = dev_upload(h, &w.hub_words, hub_words)) || (rc = dev_upload(h, &w.w_rec, wrec)) || (rc = dev_upload(h, &w.fwd_rec, frec)) ||
        (rc = dev_alloc(h, &w.wh_part, (size_t)w.n_hub_segs * dm.dq)))
        return rc;
    CK(cudaStreamSynchronize(h->stream));     // host vectors go out of scope
    w.set = true;
    h->E_frozen_valid = false;
    return NTN_OK;
}

extern "C" int ntn_set_frozen_wv(ntn_handle* h, const float* wv) {
    if (!h) return NTN_EINVAL;
    if (!wv) FAIL(NTN_EINVAL, "null argument");
    CK(cudaSetDevice(h->device));
    const NtnDims& dm = h->dm;
    ntn_drop_graphs(h);
    float* tmp = nullptr;
    size_t n = (size_t)dm.d * dm.Nw;
    CK(cudaMalloc((void**)&tmp, n * sizeof(float)));
    h->E_frozen_valid = false;
    int rc = dev_alloc(h, &h->wv_frozen, (size_t)dm.Nw * dm.dq);
    if (rc) { cudaFree(tmp); return rc; }
    cudaError_t e = cudaMemcpyAsync(tmp, wv, n * sizeof(float), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) { ntn_launch_transpose_wv_host_layout(h, tmp, h->wv_frozen); e = cudaStreamSynchronize(h->stream); }
    cudaFree(tmp);
    if (e != cudaSuccess) FAIL(NTN_ECUDA, std::string("set_frozen_wv: ") + cudaGetErrorString(e));
    return NTN_OK;
}

// Tile / chunk tables from the per-relation unique-triple counts, per-evaluation intermediates, contraction path.
// The unique-triple tables, the incident CSR and the hub segments are already in h->b (device- or host-built).
static int finish_batch(ntn_handle* h, int64_t n, int Nu, const std::vector<int>& rel_count, int n_hub_segs) {
    const NtnDims& dm = h->dm;
    std::vector<int> rel_off(dm.R + 1, 0);
    for (int r = 0; r < dm.R; ++r) rel_off[r + 1] = rel_off[r] + rel_count[r];
    std::vector<int> t_rel, t_u0, t_n, ch_rel, ch_u0, ch_n, relch_off(dm.R + 1, 0), reltile_off(dm.R + 1, 0);
    std::vector<int> st_rel, st_u0, st_n, rels_off(dm.R + 1, 0);
    // split-K chunk of the dW contraction: about two chunks per SM (every chunk costs a Kp x Np partial that the
    // finalize kernel sums serially), never below NTN_CHUNK_ROWS, a multiple of the 32-row k-block
    const int chunk_rows = std::max(NTN_CHUNK_ROWS, (int)round_up((Nu + 2 * h->sm_count - 1) / (2 * h->sm_count), 32));
    // entity-gradient pull variant: while T' (Nu x Np floats) stays L2-resident every incident re-reads its k T' rows
    // (k_score2 + k_entity_pull2, the default); beyond that k_score materialises one contribution row per incident
    // (measured in round 1: C3 0.256 vs 0.278 ms with rows, C4 2.89 vs 2.42 ms).  NTN_GV=0/1 overrides.
    bool use_gv = (size_t)Nu * dm.Np * sizeof(float) > ((size_t)48 << 20);
    if (const char* e = getenv("NTN_GV")) use_gv = e[0] == '1';
    int score_rows = NTN_SCORE_ROWS;
    if (!use_gv) {
        score_rows = ntn_score2_rows(dm.k);          // one block of k_score2: 4 warps x (32 / k) unique triples
    } else {
        // k_score (GV variant) tile height: a block is 8 warps and runs 2 per SM; every warp takes rows/8 triples in turn.
        // Short grids pick the height that minimises (waves x triples per warp) with half a triple of fixed cost per
        // block; long grids keep 32 (64 measured 6 % slower at 316k triples).  NTN_SCORE_ROWS overrides.
        const int slots = 2 * h->sm_count;
        auto tiles_at = [&](int rows) { int64_t t = 0; for (int r = 0; r < dm.R; ++r) t += (rel_count[r] + rows - 1) / rows; return t; };
        if (tiles_at(NTN_SCORE_ROWS) <= 4 * (int64_t)slots) {
            double best = 1e300;
            for (int cand : {16, 24, 32, 40}) {
                const double cost = (double)((tiles_at(cand) + slots - 1) / slots) * (cand / 8 + 0.5);
                if (cost < best - 1e-9) { best = cost; score_rows = cand; }
            }
        }
        if (const char* e = getenv("NTN_SCORE_ROWS")) score_rows = std::max(1, atoi(e));
    }
    // Contraction tile height (rows of unique triples, <= 128 = the UMMA M).  The persistent tcgen05 grid deals
    // (tile x N-chunk) items round-robin to one CTA per SM, so a kernel lasts ceil(items / SMs) item times: 163 tiles of
    // 128 rows on 148 SMs are TWO rounds for the entity-gradient contraction (the second 10 % full) and three for the
    // forward one (two N-chunks).  Measured at 20,000 x 10 (NTN_TILE_ROWS = 128 / 96 / 72 / 64): forward 24.2 / 23.3 /
    // 26.4 / 30.3 us, entity-gradient 28.8 / 28.4 / 28.3 / 36.9 us -- an item costs nearly the same whatever its height
    // (the MMA runs M = 128 regardless, every producer thread walks its 8 rows, the weight tile is the same), so what
    // counts is the number of rounds and only then the height.  Cost model: rounds x (rows + a fixed per-item cost worth
    // 256 rows), summed over the two kernels.
    int tile_rows = NTN_TILE_ROWS;
    {
        const int ny_fwd = (dm.Np + 159) / 160, ny_ge = (dm.Kp + 159) / 160;     // N-chunks of <= 160 accumulator columns
        auto tiles_at = [&](int rows) { int64_t t = 0; for (int r = 0; r < dm.R; ++r) t += (rel_count[r] + rows - 1) / rows; return t; };
        double best = 1e300;
        for (int rows = NTN_TILE_ROWS; rows >= 32; rows -= 8) {
            const int64_t t = tiles_at(rows);
            const double cost = (double)((t * ny_fwd + h->sm_count - 1) / h->sm_count + (t * ny_ge + h->sm_count - 1) / h->sm_count) * (rows + 256);
            if (cost < best - 1e-9) { best = cost; tile_rows = rows; }
        }
        if (const char* e = getenv("NTN_TILE_ROWS")) tile_rows = std::max(8, std::min(NTN_TILE_ROWS, atoi(e) / 8 * 8));
    }
    // tiles of the swapped-role entity-gradient contraction (gemm_tcgen05.cu, TC_GE2): the tile height is its N dimension,
    // up to 160 rows; the tallest height that needs the fewest rounds of the persistent grid
    std::vector<int> g_rel, g_u0, g_n;
    {
        const int nzg = (dm.Kp + 127) / 128;
        auto tiles_at = [&](int rows) { int64_t t = 0; for (int r = 0; r < dm.R; ++r) t += (rel_count[r] + rows - 1) / rows; return t; };
        int ge_rows = 160;
        int64_t best = (tiles_at(160) * nzg + h->sm_count - 1) / h->sm_count;
        for (int rows = 144; rows >= 64; rows -= 16) {
            const int64_t rounds = (tiles_at(rows) * nzg + h->sm_count - 1) / h->sm_count;
            if (rounds < best) { best = rounds; ge_rows = rows; }
        }
        if (const char* e = getenv("NTN_GE_ROWS")) ge_rows = std::max(16, std::min(160, atoi(e) / 16 * 16));
        for (int r = 0; r < dm.R; ++r)
            for (int u0 = rel_off[r]; u0 < rel_off[r + 1]; u0 += ge_rows) {
                g_rel.push_back(r); g_u0.push_back(u0); g_n.push_back(std::min(ge_rows, rel_off[r + 1] - u0));
            }
    }
    for (int r = 0; r < dm.R; ++r) {
        for (int u0 = rel_off[r]; u0 < rel_off[r + 1]; u0 += tile_rows) {
            t_rel.push_back(r); t_u0.push_back(u0); t_n.push_back(std::min(tile_rows, rel_off[r + 1] - u0));
        }
        for (int u0 = rel_off[r]; u0 < rel_off[r + 1]; u0 += chunk_rows) {
            ch_rel.push_back(r); ch_u0.push_back(u0); ch_n.push_back(std::min(chunk_rows, rel_off[r + 1] - u0));
        }
        for (int u0 = rel_off[r]; u0 < rel_off[r + 1]; u0 += score_rows) {
            st_rel.push_back(r); st_u0.push_back(u0); st_n.push_back(std::min(score_rows, rel_off[r + 1] - u0));
        }
        rels_off[r + 1] = (int)st_rel.size();
        reltile_off[r + 1] = (int)t_rel.size();
        relch_off[r + 1] = (int)ch_rel.size();
    }
    DeviceBatch& b = h->b;
    b.N = n; b.Nu = Nu; b.ntiles = (int)t_rel.size(); b.nchunks = (int)ch_rel.size(); b.n_inc = 2 * Nu + (int)n;
    b.nstiles = (int)st_rel.size();
    b.ngtiles = (int)g_rel.size();
    b.use_gv = use_gv;
    b.n_ent_hub_segs = n_hub_segs;
    int rc;
    if ((rc = dev_upload(h, &b.t_rel, t_rel)) || (rc = dev_upload(h, &b.t_u0, t_u0)) || (rc = dev_upload(h, &b.t_n, t_n)) ||
        (rc = dev_upload(h, &b.st_rel, st_rel)) || (rc = dev_upload(h, &b.st_u0, st_u0)) || (rc = dev_upload(h, &b.st_n, st_n)) ||
        (rc = dev_upload(h, &b.rels_off, rels_off)) ||
        (rc = dev_upload(h, &b.g_rel, g_rel)) || (rc = dev_upload(h, &b.g_u0, g_u0)) || (rc = dev_upload(h, &b.g_n, g_n)) ||
        (rc = dev_upload(h, &b.ch_rel, ch_rel)) || (rc = dev_upload(h, &b.ch_u0, ch_u0)) || (rc = dev_upload(h, &b.ch_n, ch_n)) ||
        (rc = dev_upload(h, &b.relch_off, relch_off)) || (rc = dev_upload(h, &b.reltile_off, reltile_off)) ||
        (rc = dev_alloc(h, &b.T, (size_t)Nu * dm.Np)) ||
        // (M, M_lo: at least one 128-row TMA box, so that the tensor map of a tiny batch job covers allocated memory)
        (rc = dev_alloc(h, &b.M, (size_t)std::max(Nu, 128) * dm.Np)) ||
        (rc = dev_alloc(h, &b.M_lo, h->cfg.gemm_path != NTN_GEMM_SIMT ? (size_t)std::max(Nu, 128) * dm.Np : 1)) ||
        (rc = dev_alloc(h, &b.GA, (size_t)Nu * dm.Kp)) || (rc = dev_alloc(h, &b.GV, use_gv ? ((size_t)Nu + (size_t)n) * dm.dq : 1)) || (rc = dev_alloc(h, &b.acoef, use_gv ? ((size_t)Nu + (size_t)n) * dm.k : 1)) ||
        (rc = dev_alloc(h, &b.crec, use_gv ? 1 : ((size_t)Nu + (size_t)n) * (size_t)round_up(dm.k + 1, 4))) ||
        (rc = dev_alloc(h, &b.dWpart, (size_t)b.nchunks * dm.Kp * dm.Np)) ||
        (rc = dev_alloc(h, &b.tp_cost, (size_t)b.nstiles)) || (rc = dev_alloc(h, &b.tp_nact, (size_t)b.nstiles)) ||
        (rc = dev_alloc(h, &b.tp_dU, (size_t)b.nstiles * dm.k)) ||
        (rc = dev_alloc(h, &b.eh_part, (size_t)b.n_ent_hub_segs * dm.dq)) ||
        (rc = dev_alloc(h, &b.pull_rng, (size_t)ntn_pull_ranges(h))))
        return rc;
    ntn_launch_pull_ranges(h);
    b.ccoef = b.acoef + (use_gv ? (size_t)Nu * dm.k : 0);
    b.RS = (int)round_up(dm.k + 1, 4);
    CK(cudaStreamSynchronize(h->stream));            // the small tables above come from host vectors that die here
    // contraction path for this batch
    h->gemm_path_active = NTN_GEMM_SIMT;
    b.m_split = false;
    if (h->cfg.gemm_path != NTN_GEMM_SIMT) {
        if (ntn_tc_supported(h)) {
            rc = ntn_tc_prepare_batch(h);
            if (rc) { ntn_tc_release(h); b.m_split = false; return rc; }     // (a half-built state must not be reused by the next job)
            h->gemm_path_active = NTN_GEMM_TCGEN05;
        } else if (h->cfg.gemm_path == NTN_GEMM_TCGEN05) {
            FAIL(NTN_EUNSUP, "tcgen05 contraction path does not support this shape");
        }
    }
    return NTN_OK;
}

// Host builder of the batch tables: the same tables batch_index.cu builds on the device, element for element.  Kept as
// the cross-check of the device builder (NTN_SET_BATCH_HOST=1, tests/test_gpu_parity.py); not the default.
static int set_batch_host_tables(ntn_handle* h, int64_t n, const int32_t* e1, const int32_t* rel, const int32_t* e2,
                                 const int32_t* e3) {
    const NtnDims& dm = h->dm;
    for (int64_t j = 0; j < n; ++j) {
        if (e1[j] < 0 || e1[j] >= dm.Ne || e2[j] < 0 || e2[j] >= dm.Ne || e3[j] < 0 || e3[j] >= dm.Ne || rel[j] < 0 || rel[j] >= dm.R)
            FAIL(NTN_EINVAL, "entity or relation index out of range in column " + std::to_string(j));
    }
    // stable LSD radix sort of the column ids by the packed key (rel, e1, e2): stability keeps a triple's corrupt
    // copies in batch order
    std::vector<int> order((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    {
        int eb = 1; while ((1 << eb) < dm.Ne) ++eb;
        int rb = 1; while ((1 << rb) < dm.R) ++rb;
        if (2 * eb + rb > 63) FAIL(NTN_EUNSUP, "key (rel, e1, e2) does not fit 63 bits");
        std::vector<uint64_t> key((size_t)n), key2((size_t)n);
        std::vector<int> order2((size_t)n);
        for (int64_t j = 0; j < n; ++j) key[j] = ((uint64_t)rel[j] << (2 * eb)) | ((uint64_t)e1[j] << eb) | (uint64_t)e2[j];
        std::vector<size_t> cnt(2049);
        for (int shift = 0; shift < 2 * eb + rb; shift += 11) {
            std::fill(cnt.begin(), cnt.end(), 0);
            for (int64_t i = 0; i < n; ++i) cnt[((key[i] >> shift) & 0x7ff) + 1]++;
            for (int b2 = 0; b2 < 2048; ++b2) cnt[b2 + 1] += cnt[b2];
            for (int64_t i = 0; i < n; ++i) { size_t p2 = cnt[(key[i] >> shift) & 0x7ff]++; key2[p2] = key[i]; order2[p2] = order[i]; }
            key.swap(key2); order.swap(order2);
        }
    }
    std::vector<int> u_e1, u_e2, u_rel, c_off, c_e3((size_t)n), col_u((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        int j = order[i];
        if (i == 0 || rel[j] != u_rel.back() || e1[j] != u_e1.back() || e2[j] != u_e2.back()) {
            u_e1.push_back(e1[j]); u_e2.push_back(e2[j]); u_rel.push_back(rel[j]);
            c_off.push_back((int)i);
        }
        c_e3[i] = e3[j];
        col_u[i] = (int)u_e1.size() - 1;
    }
    c_off.push_back((int)n);
    int Nu = (int)u_e1.size();
    std::vector<int> rel_count(dm.R, 0);
    for (int u = 0; u < Nu; ++u) rel_count[u_rel[u]]++;
    // entity -> incidents: (e1 role, e2 role, corrupt columns), stable counting sort by entity
    int n_inc = 2 * Nu + (int)n;
    std::vector<int> inc_off(dm.Ne + 1, 0);
    std::vector<int4> inc(n_inc);
    for (int u = 0; u < Nu; ++u) { inc_off[u_e1[u] + 1]++; inc_off[u_e2[u] + 1]++; }
    for (int64_t c = 0; c < n; ++c) inc_off[c_e3[c] + 1]++;
    for (int e = 0; e < dm.Ne; ++e) inc_off[e + 1] += inc_off[e];
    {
        std::vector<int> cur(inc_off.begin(), inc_off.end() - 1);
        for (int u = 0; u < Nu; ++u) inc[cur[u_e1[u]]++] = make_int4(u, u, 0, u_rel[u]);
        for (int u = 0; u < Nu; ++u) inc[cur[u_e2[u]]++] = make_int4(u, u, 1, u_rel[u]);
        for (int64_t c = 0; c < n; ++c) inc[cur[c_e3[c]]++] = make_int4(col_u[c], Nu + (int)c, 2, u_rel[col_u[c]]);
    }
    std::vector<int> sr, sb, se, first, count;
    build_hub_segments(inc_off, dm.Ne, sr, sb, se, first, count);
    // visiting order of the entity pull: descending incident count (clamped like the device builder), stable
    std::vector<int> ent_order(dm.Ne);
    std::iota(ent_order.begin(), ent_order.end(), 0);
    std::stable_sort(ent_order.begin(), ent_order.end(), [&](int x, int y) {
        return std::min(inc_off[x + 1] - inc_off[x], NTN_ORDER_CLAMP) > std::min(inc_off[y + 1] - inc_off[y], NTN_ORDER_CLAMP);
    });
    DeviceBatch& b = h->b;
    int rc;
    if ((rc = dev_upload(h, &b.ent_order, ent_order)) || (rc = dev_upload(h, &b.u_e1, u_e1)) || (rc = dev_upload(h, &b.u_e2, u_e2)) || (rc = dev_upload(h, &b.u_rel, u_rel)) ||
        (rc = dev_upload(h, &b.c_off, c_off)) || (rc = dev_upload(h, &b.c_e3, c_e3)) ||
        (rc = dev_upload(h, &b.inc_off, inc_off)) || (rc = dev_upload(h, &b.inc, inc)) ||
        (rc = dev_upload(h, &b.eh_ent, sr)) || (rc = dev_upload(h, &b.eh_begin, sb)) || (rc = dev_upload(h, &b.eh_end, se)) ||
        (rc = dev_upload(h, &b.eh_first, first)) || (rc = dev_upload(h, &b.eh_count, count)))
        return rc;
    CK(cudaStreamSynchronize(h->stream));
    return finish_batch(h, n, Nu, rel_count, (int)sr.size());
}

static int set_batch_dev_columns(ntn_handle* h, int64_t n, const int* d_e1, const int* d_rel, const int* d_e2, const int* d_e3) {
    std::vector<int> rel_count(h->dm.R, 0);
    int counts[2] = {0, 0};
    int rc = ntn_index_batch_dev(h, n, d_e1, d_rel, d_e2, d_e3, counts, rel_count.data());
    if (rc) return rc;
    return finish_batch(h, n, counts[0], rel_count, counts[1]);
}

static int ensure_cols(ntn_handle* h, int64_t n) {
    if (h->cols_cap >= (size_t)n && h->h_cols) return NTN_OK;
    if (h->h_cols) cudaFreeHost(h->h_cols);
    if (h->d_cols) cudaFree(h->d_cols);
    h->h_cols = nullptr; h->d_cols = nullptr; h->cols_cap = 0;
    const size_t cap = std::max<size_t>((size_t)n, 1);
    CK(cudaMallocHost((void**)&h->h_cols, 4 * cap * sizeof(int)));
    CK(cudaMalloc((void**)&h->d_cols, 4 * cap * sizeof(int)));
    h->cols_cap = cap;
    return NTN_OK;
}

extern "C" int ntn_set_batch(ntn_handle* h, int64_t n, const int32_t* e1, const int32_t* rel, const int32_t* e2,
                             const int32_t* e3) {
    if (!h) return NTN_EINVAL;
    if (n < 0 || (n > 0 && (!e1 || !rel || !e2 || !e3))) FAIL(NTN_EINVAL, "null argument");
    if (n > ((int64_t)1 << 30) / 3) FAIL(NTN_EUNSUP, "batch too large");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    ntn_drop_graphs(h);                     // graphs bake buffer pointers and grid sizes; device buffers are reused when large enough
    if (getenv("NTN_SET_BATCH_HOST")) return set_batch_host_tables(h, n, e1, rel, e2, e3);
    // columns -> pinned staging -> device; everything else happens there (batch_index.cu)
    { int rc0 = ensure_cols(h, n); if (rc0) return rc0; }
    h->cols_n = n;
    const int32_t* src[4] = {e1, rel, e2, e3};
    for (int c = 0; c < 4 && n > 0; ++c) memcpy(h->h_cols + (size_t)c * n, src[c], (size_t)n * sizeof(int));
    if (n > 0) CK(cudaMemcpyAsync(h->d_cols, h->h_cols, 4 * (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    return set_batch_dev_columns(h, n, h->d_cols, h->d_cols + n, h->d_cols + 2 * n, h->d_cols + 3 * n);
}

extern "C" int ntn_set_training_triples(ntn_handle* h, int64_t T, const int32_t* e1, const int32_t* rel, const int32_t* e2) {
    if (!h) return NTN_EINVAL;
    if (T < 0 || (T > 0 && (!e1 || !rel || !e2))) FAIL(NTN_EINVAL, "null argument");
    if (T > ((int64_t)1 << 31) - 1) FAIL(NTN_EUNSUP, "too many training triples");
    const NtnDims& dm = h->dm;
    for (int64_t j = 0; j < T; ++j)
        if (e1[j] < 0 || e1[j] >= dm.Ne || e2[j] < 0 || e2[j] >= dm.Ne || rel[j] < 0 || rel[j] >= dm.R)
            FAIL(NTN_EINVAL, "index out of range in training triple " + std::to_string(j));
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    if (h->d_train) { cudaFree(h->d_train); h->d_train = nullptr; }
    h->n_train = 0;
    const size_t nn = (size_t)std::max<int64_t>(T, 1);
    CK(cudaMalloc((void**)&h->d_train, 3 * nn * sizeof(int)));
    if (T > 0) {
        CK(cudaMemcpy(h->d_train, e1, (size_t)T * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->d_train + T, rel, (size_t)T * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->d_train + 2 * T, e2, (size_t)T * sizeof(int), cudaMemcpyHostToDevice));
    }
    int rc = ntn_build_true_keys_dev(h, T, h->d_train, &h->d_true_keys);
    if (rc) return rc;
    h->n_train = T;
    return NTN_OK;
}

extern "C" int ntn_generate_batch_dev(ntn_handle* h, int64_t first, int32_t batch, int32_t corrupt, uint64_t seed, uint64_t job,
                                      int32_t filter_true) {
    if (!h) return NTN_EINVAL;
    if (!h->d_train) FAIL(NTN_ESTATE, "ntn_set_training_triples has not been called");
    if (batch < 0 || corrupt < 0 || first < 0 || first + batch > h->n_train) FAIL(NTN_EINVAL, "batch range outside the training triples");
    const int64_t n = (int64_t)batch * corrupt;
    if (n > ((int64_t)1 << 30) / 3) FAIL(NTN_EUNSUP, "batch too large");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    ntn_drop_graphs(h);
    int rc = ensure_cols(h, n);
    if (rc) return rc;
    rc = ntn_generate_columns_dev(h, h->d_train, h->n_train, first, batch, corrupt, seed, job,
                                  filter_true ? h->d_true_keys : nullptr, h->n_train, h->d_cols);
    if (rc) return rc;
    h->cols_n = n;
    return set_batch_dev_columns(h, n, h->d_cols, h->d_cols + n, h->d_cols + 2 * n, h->d_cols + 3 * n);
}

extern "C" int ntn_get_batch_columns(ntn_handle* h, int64_t n, int32_t* e1, int32_t* rel, int32_t* e2, int32_t* e3) {
    if (!h) return NTN_EINVAL;
    if (n != h->cols_n) FAIL(NTN_EINVAL, "no staged batch job of this size (columns are kept for ntn_set_batch with host pointers and ntn_generate_batch_dev)");
    if (n > 0 && (!e1 || !rel || !e2 || !e3)) FAIL(NTN_EINVAL, "null argument");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    int32_t* dst[4] = {e1, rel, e2, e3};
    for (int c = 0; c < 4 && n > 0; ++c) CK(cudaMemcpy(dst[c], h->d_cols + (size_t)c * n, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    return NTN_OK;
}

extern "C" int ntn_set_batch_dev(ntn_handle* h, int64_t n, const int32_t* d_e1, const int32_t* d_rel, const int32_t* d_e2,
                                 const int32_t* d_e3) {
    if (!h) return NTN_EINVAL;
    if (n < 0 || (n > 0 && (!d_e1 || !d_rel || !d_e2 || !d_e3))) FAIL(NTN_EINVAL, "null argument");
    if (n > ((int64_t)1 << 30) / 3) FAIL(NTN_EUNSUP, "batch too large");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    ntn_drop_graphs(h);
    h->cols_n = -1;
    return set_batch_dev_columns(h, n, d_e1, d_rel, d_e2, d_e3);
}

// Enqueue the kernels of one evaluation on h->stream with its forked branches (also what gets captured
// into the CUDA graph).  prof: bracket every phase with its own event pair.
// NTN_GRAPH_TIMELINE=1: one-thread kernels that store %globaltimer before and after every kernel group, also inside the
// captured graph -- the only way to see how the forked branches really interleave under graph replay (the event
// brackets of the profiling mode launch the kernels outside the graph).  Fetched with ntn_debug_fetch("graph_timeline").
__global__ void k_stamp(unsigned long long* dst) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    *dst = t;
}

void ntn_stamp(ntn_handle* h, int slot, cudaStream_t st) {
    if (h->d_stamps && slot >= 0 && slot < 2 * NTN_STAMP_SLOTS) k_stamp<<<1, 1, 0, st>>>(h->d_stamps + slot);
}

static int enqueue_eval_kernels(ntn_handle* h, const float* d_theta, float* d_grad, const float* wvt, bool prof, int slot) {
    auto begin = [&](int p, cudaStream_t st) {
        if (prof) cudaEventRecord(h->ev[slot][p][0], st);
        if (h->d_stamps) k_stamp<<<1, 1, 0, st>>>(h->d_stamps + 2 * p);
    };
    auto end = [&](int p, cudaStream_t st) {
        if (prof) cudaEventRecord(h->ev[slot][p][1], st);
        if (h->d_stamps) k_stamp<<<1, 1, 0, st>>>(h->d_stamps + 2 * p + 1);
    };
    enum { P_EB, P_WP, P_FWD, P_SC, P_DW, P_GE, P_EP, P_WPULL, P_FIN };
    cudaStream_t s0 = h->stream, s1 = h->stream3;
    const bool tc = h->gemm_path_active == NTN_GEMM_TCGEN05 && h->b.ntiles > 0;
    int rc;
    // data-parallel: when the gradient buffer is a bound symmetric allocation (and carries the {cost, n_active} tail) the
    // all-reduce over the ranks runs inside the evaluation, chunk by chunk beside the word pull (exchange.cu)
    const ExchangeBinding* xb = h->grad_tail ? ntn_find_binding(h, d_grad) : nullptr;
    // ... and when gE itself is bound (ntn_exchange_bind_ge) the sum over the ranks is taken one stage earlier, on the entity
    // gradient, while the entity pull is running: the word pull then produces the full-batch word gradient on every rank
    // and only the small blocks and the tail are exchanged after it
    const ExchangeBinding* xge = (xb && h->xge_set) ? &h->xge : nullptr;
    h->ge_mode = xge != nullptr;
    h->word_chunks = (xb && !xge) ? h->x_chunks : 1;
    h->x_stamp = 0;
    if (h->d_stamps) cudaMemsetAsync(h->d_stamps, 0, 2 * NTN_STAMP_SLOTS * sizeof(unsigned long long), h->stream);
    // ---- fork: weight preparation runs beside the entity build
    CK(cudaEventRecord(h->ev_f0, s0));
    CK(cudaStreamWaitEvent(s1, h->ev_f0, 0));
    begin(P_WP, s1);
    if (tc) { if ((rc = ntn_tc_w_prep(h, d_theta, s1))) return rc; } else ntn_launch_w_prep(h, d_theta, s1);
    end(P_WP, s1);
    CK(cudaEventRecord(h->ev_j0, s1));
    // NTN.java:103 rebuilds the entity vectors on every call; from frozen word vectors the result never changes, so
    // it is built once (eval_internal) and only the theta-sourced variant pays for it per evaluation
    begin(P_EB, s0); if (h->cfg.wv_source != NTN_WV_FROZEN) ntn_launch_entity_build(h, wvt); end(P_EB, s0);
    CK(cudaStreamWaitEvent(s0, h->ev_j0, 0));
    begin(P_FWD, s0);
    if (tc) { if ((rc = ntn_tc_gemm_fwd(h))) return rc; } else ntn_launch_gemm_fwd_simt(h);   // :163-197
    end(P_FWD, s0);
    begin(P_SC, s0); ntn_launch_score(h, d_theta); end(P_SC, s0);                            // :200-289
    // ---- fork: dW contraction + small-block finalize beside the entity/word gradient chain
    CK(cudaEventRecord(h->ev_f1, s0));
    CK(cudaStreamWaitEvent(s1, h->ev_f1, 0));
    begin(P_DW, s1);
    if (tc) { if ((rc = ntn_tc_gemm_dw(h, s1))) return rc; } else ntn_launch_gemm_dw_simt(h, s1);   // :301-349
    end(P_DW, s1);
    begin(P_FIN, s1); ntn_launch_finalize_small(h, d_theta, d_grad, s1); end(P_FIN, s1);      // :398-401,433-438
    CK(cudaEventRecord(h->ev_j1, s1));
    begin(P_GE, s0);
    if (tc) { if ((rc = ntn_tc_gemm_ge(h))) return rc; } else ntn_launch_gemm_ge_simt(h);     // :352-388
    end(P_GE, s0);
    begin(P_EP, s0); ntn_launch_entity_pull(h, xge); end(P_EP, s0);                           // Util.java:74-98
    if (xge) CK(cudaStreamWaitEvent(s0, h->ev_xdone, 0));                                     // gE summed over the ranks, everywhere
    begin(P_WPULL, s0); ntn_launch_word_pull(h, d_theta, d_grad, xge ? nullptr : xb); end(P_WPULL, s0);   // :411-430
    CK(cudaStreamWaitEvent(s0, h->ev_j1, 0));
    ntn_launch_final_reduce(h, h->grad_tail ? d_grad + h->dm.Pint : nullptr);                 // :430,437
    if (xb) {
        // the last word chunk (everything when the pull is not chunked) and the tail; the earlier chunks went out on
        // the side stream while the pull was running
        const NtnDims& dm = h->dm;
        int w0, wn, off, nb;
        ntn_word_chunk(dm, h->word_chunks - 1, h->word_chunks, &w0, &wn, &off, &nb);
        const bool pushed = h->x_mode == 2;
        XRange r0{dm.oWVint + (int64_t)w0 * dm.dq, (int64_t)wn * dm.dq, h->word_chunks, pushed};
        XRange r1{0, dm.oWVint, 0, false}, r2{dm.Pint, 4, h->word_chunks + 1, false};
        if (xge) r0 = XRange{0, 0, 0, false};                    // the word block is already the full-batch one on every rank
        else if (h->word_chunks > 1) CK(cudaStreamWaitEvent(s0, h->ev_xdone, 0));
        ntn_launch_exchange(h, xb, r0, r1, r2, true, s0);
    }
    CK(cudaMemcpyAsync(h->h_out, h->d_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, s0));
    return NTN_OK;
}

void ntn_drop_graphs(ntn_handle* h) {
    for (auto& g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    h->graphs.clear();
}

// ------------------------------------------------------------------------------------------
// the evaluation driver                                          NTN.computeAt, NTN.java:92-443
// ------------------------------------------------------------------------------------------
static int eval_internal(ntn_handle* h, const float* d_theta, const uint8_t* corrupt_side, float* d_grad) {
    const NtnDims& dm = h->dm;
    if (!h->w.set) FAIL(NTN_ESTATE, "ntn_set_entity_words has not been called");
    if (!corrupt_side) FAIL(NTN_EINVAL, "corrupt_side is null");
    const float* wvt;
    if (h->cfg.wv_source == NTN_WV_FROZEN) {
        if (!h->wv_frozen) FAIL(NTN_ESTATE, "wv_source is NTN_WV_FROZEN but ntn_set_frozen_wv has not been called");
        wvt = h->wv_frozen;
    } else {
        wvt = d_theta + dm.oWVint;
    }
    const bool prof = h->profiling;
    const int slot = h->ring_pos;
    h->ring_pos = (slot + 1) % NTN_RING;
    if (prof) drain_profile_slot(h, slot);
    // pinned ring slot: wait only for the evaluation that used this slot NTN_RING calls ago
    CK(cudaEventSynchronize(h->side_ev[slot]));
    uint8_t* hs = h->h_side + (size_t)slot * dm.R;
    for (int r = 0; r < dm.R; ++r) hs[r] = corrupt_side[r] ? 1 : 0;
    CK(cudaMemcpyAsync(h->side, hs, dm.R, cudaMemcpyHostToDevice, h->stream));
    CK(cudaEventRecord(h->side_ev[slot], h->stream));
    int rc;
    if (h->cfg.wv_source == NTN_WV_FROZEN && !h->E_frozen_valid) {
        ntn_launch_entity_build(h, wvt);
        h->E_frozen_valid = true;
    }
    if (prof || !h->use_graph || h->stream == nullptr /* legacy default stream cannot be captured */) {
        h->launches = 0;
        if ((rc = enqueue_eval_kernels(h, d_theta, d_grad, wvt, prof, slot))) return rc;
        if (prof) h->ev_pending[slot] = true;
    } else {
        // the kernel sequence (with its forked branches) is captured once per (theta, grad) buffer pair and
        // replayed: the evaluation is a dozen short kernels, so launch latency and inter-kernel gaps matter
        ntn_handle::GraphEntry* ge = nullptr;
        for (auto& g : h->graphs) if (g.theta == d_theta && g.grad == d_grad && g.stream == h->stream) ge = &g;
        if (!ge) {
            if (h->graphs.size() >= 8) { cudaGraphExecDestroy(h->graphs.front().exec); h->graphs.erase(h->graphs.begin()); }
            cudaGraph_t graph = nullptr;
            h->launches = 0;
            CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            rc = enqueue_eval_kernels(h, d_theta, d_grad, wvt, false, slot);
            cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) FAIL(NTN_ECUDA, std::string("graph capture: ") + cudaGetErrorString(e));
            ntn_handle::GraphEntry ne{d_theta, d_grad, h->stream, nullptr, h->launches};
            e = cudaGraphInstantiate(&ne.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) FAIL(NTN_ECUDA, std::string("graph instantiate: ") + cudaGetErrorString(e));
            h->graphs.push_back(ne);
            ge = &h->graphs.back();
        }
        h->launches = ge->launches;
        CK(cudaGraphLaunch(ge->exec, h->stream));
    }
    CK(cudaGetLastError());
    return NTN_OK;
}

extern "C" int ntn_last_cost(ntn_handle* h, double* cost, int64_t* n_active) {
    if (!h) return NTN_EINVAL;
    CK(cudaStreamSynchronize(h->stream));
    if (cost) *cost = h->h_out[0];
    if (n_active) *n_active = (int64_t)llround(h->h_out[1]);
    if (h->gemm_path_active == NTN_GEMM_TCGEN05 && ntn_tc_check_error(h))
        FAIL(NTN_ECUDA, "tcgen05 pipeline: a bounded mbarrier wait expired (results invalid)");
    return NTN_OK;
}

extern "C" double* ntn_result_dev(ntn_handle* h) { return h ? h->d_out : nullptr; }
extern "C" int ntn_set_grad_tail(ntn_handle* h, int on) {
    if (!h) return NTN_EINVAL;
    ntn_drop_graphs(h);
    h->grad_tail = on != 0;
    return NTN_OK;
}

extern "C" int ntn_last_phase_ms(ntn_handle* h, float* ms) {
    if (!h || !ms) return NTN_EINVAL;
    for (int s = 0; s < NTN_RING; ++s) drain_profile_slot(h, s);
    for (int i = 0; i < NTN_NUM_PHASES; ++i) ms[i] = h->phase_n ? (float)(h->phase_acc[i] / h->phase_n) : 0.f;
    return NTN_OK;
}

extern "C" int ntn_eval_dev(ntn_handle* h, const float* d_theta, const uint8_t* corrupt_side, float* d_grad,
                            double* cost, int64_t* n_active) {
    if (!h) return NTN_EINVAL;
    if (!d_theta || !d_grad) FAIL(NTN_EINVAL, "null device pointer");
    CK(cudaSetDevice(h->device));
    int rc = eval_internal(h, d_theta, corrupt_side, d_grad);
    if (rc) return rc;
    if (cost || n_active) return ntn_last_cost(h, cost, n_active);
    return NTN_OK;
}

extern "C" int ntn_to_internal_dev(ntn_handle* h, const float* d_ref, float* d_int) {
    if (!h) return NTN_EINVAL;
    if (!d_ref || !d_int) FAIL(NTN_EINVAL, "null device pointer");
    CK(cudaSetDevice(h->device));
    ntn_launch_to_internal(h, d_ref, d_int);
    CK(cudaGetLastError());
    return NTN_OK;
}
extern "C" int ntn_from_internal_dev(ntn_handle* h, const float* d_int, float* d_ref) {
    if (!h) return NTN_EINVAL;
    if (!d_ref || !d_int) FAIL(NTN_EINVAL, "null device pointer");
    CK(cudaSetDevice(h->device));
    ntn_launch_from_internal(h, d_int, d_ref);
    CK(cudaGetLastError());
    return NTN_OK;
}

// host <-> pinned staging with type conversion.  The arrays are ~57 MB; conversion runs on an OpenMP pool and is
// pipelined against the PCIe copies in NTN_XFER_CHUNK-element chunks (convert chunk i+1 while chunk i is in flight).
static int64_t xfer_chunk() {
    static int64_t c = [] {
        const char* e = getenv("NTN_XFER_CHUNK_LOG2");
        int l = e ? atoi(e) : 20;       // 4 MB of floats per chunk: measured best of 2^18..2^22 on the B200 host
        return (int64_t)1 << std::min(24, std::max(16, l));
    }();
    return c;
}
#define NTN_XFER_CHUNK xfer_chunk()
#include "host/convert.hpp"
template <typename S, typename D> static void convert_parallel(const S* src, D* dst, int64_t n) { convert_range(src, dst, n); }
extern "C" int ntn_host_convert_f64_f32(const double* src, float* dst, int64_t n) {
    if ((!src || !dst) && n > 0) return NTN_EINVAL;
    convert_range(src, dst, n);
    return NTN_OK;
}
extern "C" int ntn_host_convert_f32_f64(const float* src, double* dst, int64_t n) {
    if ((!src || !dst) && n > 0) return NTN_EINVAL;
    convert_range(src, dst, n);
    return NTN_OK;
}

// chunk boundaries of a pipelined transfer: the first chunks are small (the pipeline fills quickly: what is exposed is
// one conversion before the first copy and one copy before the first conversion), then they double up to NTN_XFER_CHUNK
static std::vector<int64_t> xfer_chunks(int64_t n) {
    std::vector<int64_t> off{0};
    int64_t c = std::min<int64_t>(NTN_XFER_CHUNK, (int64_t)1 << 18);
    while (off.back() < n) {
        off.push_back(std::min(n, off.back() + c));
        c = std::min<int64_t>(NTN_XFER_CHUNK, c * 2);
    }
    return off;
}

// theta (host, S) -> pinned float staging -> device, chunk-pipelined
template <typename S> static cudaError_t upload_pipelined(ntn_handle* h, const S* src, float* d_dst, int64_t n) {
    const std::vector<int64_t> off = xfer_chunks(n);
    for (size_t i = 0; i + 1 < off.size(); ++i) {
        const int64_t o = off[i], c = off[i + 1] - o;
        convert_range(src + o, h->h_stage + o, c);
        cudaError_t e = cudaMemcpyAsync(d_dst + o, h->h_stage + o, (size_t)c * sizeof(float), cudaMemcpyHostToDevice, h->stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
// device -> pinned float staging -> host (D), chunk-pipelined on the handle's stream
template <typename D> static cudaError_t download_pipelined(ntn_handle* h, const float* d_src, D* dst, int64_t n) {
    const std::vector<int64_t> off = xfer_chunks(n);
    const int nchunks = (int)off.size() - 1;
    if ((int)h->xfer_ev.size() < nchunks) {
        size_t old = h->xfer_ev.size();
        h->xfer_ev.resize(nchunks);
        for (size_t i = old; i < h->xfer_ev.size(); ++i) cudaEventCreateWithFlags(&h->xfer_ev[i], cudaEventDisableTiming);
    }
    for (int i = 0; i < nchunks; ++i) {
        const int64_t o = off[i], c = off[i + 1] - o;
        cudaError_t e = cudaMemcpyAsync(h->h_stage + o, d_src + o, (size_t)c * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
        if (e != cudaSuccess) return e;
        cudaEventRecord(h->xfer_ev[i], h->stream);
    }
    static const bool timing = getenv("NTN_TIMING") != nullptr;
    double t_wait = 0, t_conv = 0;
    for (int i = 0; i < nchunks; ++i) {
        const int64_t o = off[i], c = off[i + 1] - o;
        const auto a0 = std::chrono::steady_clock::now();
        // poll instead of a blocking wait: a chunk is ~70 us of DMA, the wake-up of a blocked thread is of that order
        cudaError_t e;
        while ((e = cudaEventQuery(h->xfer_ev[i])) == cudaErrorNotReady) {}
        if (e != cudaSuccess) return e;
        const auto a1 = std::chrono::steady_clock::now();
        convert_range(h->h_stage + o, dst + o, c);
        if (timing) {
            const auto a2 = std::chrono::steady_clock::now();
            t_wait += std::chrono::duration<double, std::milli>(a1 - a0).count();
            t_conv += std::chrono::duration<double, std::milli>(a2 - a1).count();
        }
    }
    if (timing) fprintf(stderr, "[download] %d chunks: waiting for DMA %.3f ms, converting %.3f ms\n", nchunks, t_wait, t_conv);
    return cudaSuccess;
}

static int ensure_host_path(ntn_handle* h) {
    const NtnDims& dm = h->dm;
    int rc;
    if (!h->d_ref_theta && (rc = dev_alloc(h, &h->d_ref_theta, (size_t)dm.P))) return rc;
    if (!h->d_ref_grad && (rc = dev_alloc(h, &h->d_ref_grad, (size_t)dm.P))) return rc;
    if (!h->d_int_theta && (rc = dev_alloc(h, &h->d_int_theta, (size_t)dm.Pint))) return rc;
    if (!h->d_int_grad && (rc = dev_alloc(h, &h->d_int_grad, (size_t)dm.Pint))) return rc;
    if (!h->h_stage) CK(cudaMallocHost((void**)&h->h_stage, (size_t)dm.P * sizeof(float)));
    return NTN_OK;
}

template <typename TH, typename TG>
static int eval_host(ntn_handle* h, const TH* theta, const uint8_t* corrupt_side, double* cost, TG* grad, int64_t* n_active) {
    if (!h) return NTN_EINVAL;
    if (!theta || !grad || !cost) FAIL(NTN_EINVAL, "null argument");
    CK(cudaSetDevice(h->device));
    int rc = ensure_host_path(h);
    if (rc) return rc;
    const NtnDims& dm = h->dm;
    static const bool timing = getenv("NTN_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    CK(cudaStreamSynchronize(h->stream));
    const auto t0 = now();
    CK(upload_pipelined(h, theta, h->d_ref_theta, dm.P));
    const auto t1 = now();
    ntn_launch_to_internal(h, h->d_ref_theta, h->d_int_theta);
    rc = eval_internal(h, h->d_int_theta, corrupt_side, h->d_int_grad);
    if (rc) return rc;
    int launches = h->launches + 2;
    ntn_launch_from_internal(h, h->d_int_grad, h->d_ref_grad);
    h->launches = launches + 2;
    if (timing) cudaStreamSynchronize(h->stream);
    const auto t2 = now();
    CK(download_pipelined(h, h->d_ref_grad, grad, dm.P));
    const auto t3 = now();
    if (timing)
        fprintf(stderr, "[ntn_eval] convert+enqueue upload %.3f ms, H2D tail + evaluation %.3f ms, D2H + convert %.3f ms\n",
                ms(t0, t1), ms(t1, t2), ms(t2, t3));
    return ntn_last_cost(h, cost, n_active);
}

extern "C" int ntn_eval(ntn_handle* h, const double* theta, const uint8_t* corrupt_side, double* cost, double* grad,
                        int64_t* n_active) {
    return eval_host<double, double>(h, theta, corrupt_side, cost, grad, n_active);
}
extern "C" int ntn_eval_f32(ntn_handle* h, const float* theta, const uint8_t* corrupt_side, double* cost, float* grad,
                            int64_t* n_active) {
    return eval_host<float, float>(h, theta, corrupt_side, cost, grad, n_active);
}

extern "C" int ntn_upload_theta(ntn_handle* h, const double* theta, float* d_int) {
    if (!h) return NTN_EINVAL;
    if (!theta || !d_int) FAIL(NTN_EINVAL, "null argument");
    CK(cudaSetDevice(h->device));
    int rc = ensure_host_path(h);
    if (rc) return rc;
    const NtnDims& dm = h->dm;
    CK(cudaStreamSynchronize(h->stream));
    CK(upload_pipelined(h, theta, h->d_ref_theta, dm.P));
    ntn_launch_to_internal(h, h->d_ref_theta, d_int);
    CK(cudaStreamSynchronize(h->stream));
    return NTN_OK;
}
extern "C" int ntn_download_theta(ntn_handle* h, const float* d_int, double* theta) {
    if (!h) return NTN_EINVAL;
    if (!theta || !d_int) FAIL(NTN_EINVAL, "null argument");
    CK(cudaSetDevice(h->device));
    int rc = ensure_host_path(h);
    if (rc) return rc;
    const NtnDims& dm = h->dm;
    ntn_launch_from_internal(h, d_int, h->d_ref_grad);
    CK(download_pipelined(h, h->d_ref_grad, theta, dm.P));
    return NTN_OK;
}

// test hook: copy an intermediate of the last evaluation to the host (T, M, GA: FP32; returns count)
extern "C" int64_t ntn_debug_fetch(ntn_handle* h, const char* name, float* out, int64_t cap) {
    if (!h || !name) return -1;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    const NtnDims& dm = h->dm;
    const DeviceBatch& b = h->b;
    const float* src = nullptr; int64_t n = 0;
    std::string s(name);
    if (s == "T") { src = b.T; n = (int64_t)b.Nu * dm.Np; }
    else if (s == "M" && b.m_split) {
        // the contractions take M pre-split: hand back hi + lo
        n = (int64_t)b.Nu * dm.Np;
        if (out && cap >= n) {
            std::vector<float> lo((size_t)n);
            cudaMemcpy(out, b.M, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost);
            cudaMemcpy(lo.data(), b.M_lo, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost);
            for (int64_t i = 0; i < n; ++i) out[i] += lo[(size_t)i];
        }
        return n;
    }
    else if (s == "M") { src = b.M; n = (int64_t)b.Nu * dm.Np; }
    else if (s == "GA") { src = b.GA; n = (int64_t)b.Nu * dm.Kp; }
    else if (s == "dWpart") { src = b.dWpart; n = (int64_t)b.nchunks * dm.Kp * dm.Np; }
    else if (s == "E") { src = h->E; n = (int64_t)dm.Ne * dm.Kp; }
    else if (s == "gE") { src = h->gE; n = (int64_t)dm.Ne * dm.dq; }
    else if (s == "tc_timeline") return ntn_tc_fetch_timeline(h, out, cap);
    else if (s == "graph_timeline") {
        // begin/end of every kernel group of the LAST evaluation, microseconds relative to the earliest stamp
        if (!h->d_stamps) return 0;
        const int m = 2 * NTN_STAMP_SLOTS;
        if (out && cap >= m) {
            unsigned long long t[2 * NTN_STAMP_SLOTS], t0 = ~0ull;
            cudaMemcpy(t, h->d_stamps, sizeof(t), cudaMemcpyDeviceToHost);
            for (int i = 0; i < m; ++i) if (t[i] && t[i] < t0) t0 = t[i];
            for (int i = 0; i < m; ++i) out[i] = t[i] ? (float)((double)(t[i] - t0) * 1e-3) : -1.f;
        }
        return m;
    }
    else if (s.rfind("i:", 0) == 0) {
        // integer batch tables, copied bit for bit (the caller views the buffer as int32)
        const int* isrc = nullptr;
        const std::string t = s.substr(2);
        if (t == "u_e1") { isrc = b.u_e1; n = b.Nu; }
        else if (t == "u_e2") { isrc = b.u_e2; n = b.Nu; }
        else if (t == "u_rel") { isrc = b.u_rel; n = b.Nu; }
        else if (t == "c_off") { isrc = b.c_off; n = b.Nu + 1; }
        else if (t == "c_e3") { isrc = b.c_e3; n = b.N; }
        else if (t == "inc_off") { isrc = b.inc_off; n = dm.Ne + 1; }
        else if (t == "inc") { isrc = (const int*)b.inc; n = 4 * (int64_t)b.n_inc; }
        else if (t == "eh_first") { isrc = b.eh_first; n = dm.Ne; }
        else if (t == "eh_count") { isrc = b.eh_count; n = dm.Ne; }
        else if (t == "eh_begin") { isrc = b.eh_begin; n = b.n_ent_hub_segs; }
        else if (t == "eh_end") { isrc = b.eh_end; n = b.n_ent_hub_segs; }
        else if (t == "ent_order") { isrc = b.ent_order; n = dm.Ne; }
        else return -1;
        if (out && cap >= n && n > 0) cudaMemcpy(out, isrc, n * sizeof(int), cudaMemcpyDeviceToHost);
        return n;
    }
    else return -1;
    if (out && cap >= n && n > 0) cudaMemcpy(out, src, n * sizeof(float), cudaMemcpyDeviceToHost);
    return n;
}

extern "C" int ntn_score(ntn_handle* h, const double* theta, int64_t n, const int32_t* e1, const int32_t* rel,
                         const int32_t* e2, int32_t mode, double* scores) {
    if (!h) return NTN_EINVAL;
    if (!theta || (n > 0 && (!e1 || !rel || !e2 || !scores))) FAIL(NTN_EINVAL, "null argument");
    if (mode != NTN_SCORE_REFERENCE_EVAL && mode != NTN_SCORE_TRAIN) FAIL(NTN_EINVAL, "bad score mode");
    if (!h->w.set) FAIL(NTN_ESTATE, "ntn_set_entity_words has not been called");
    CK(cudaSetDevice(h->device));
    const NtnDims& dm = h->dm;
    for (int64_t j = 0; j < n; ++j)
        if (e1[j] < 0 || e1[j] >= dm.Ne || e2[j] < 0 || e2[j] >= dm.Ne || rel[j] < 0 || rel[j] >= dm.R)
            FAIL(NTN_EINVAL, "index out of range in triple " + std::to_string(j));
    int rc = ensure_host_path(h);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    convert_parallel(theta, h->h_stage, dm.P);
    CK(cudaMemcpyAsync(h->d_ref_theta, h->h_stage, (size_t)dm.P * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    ntn_launch_to_internal(h, h->d_ref_theta, h->d_int_theta);
    const float* wvt;
    if (h->cfg.wv_source == NTN_WV_FROZEN) {                   // NTN.java:592,677: same stale vectors at eval time
        if (!h->wv_frozen) FAIL(NTN_ESTATE, "wv_source is NTN_WV_FROZEN but ntn_set_frozen_wv has not been called");
        wvt = h->wv_frozen;
    } else {
        wvt = h->d_int_theta + dm.oWVint;
    }
    ntn_launch_entity_build(h, wvt);
    int *d_e1 = nullptr, *d_rel = nullptr, *d_e2 = nullptr; double* d_s = nullptr;
    size_t nn = (size_t)std::max<int64_t>(n, 1);
    cudaError_t e = cudaSuccess;
    if ((e = cudaMalloc((void**)&d_e1, nn * 4)) == cudaSuccess && (e = cudaMalloc((void**)&d_rel, nn * 4)) == cudaSuccess &&
        (e = cudaMalloc((void**)&d_e2, nn * 4)) == cudaSuccess && (e = cudaMalloc((void**)&d_s, nn * 8)) == cudaSuccess && n > 0) {
        cudaMemcpyAsync(d_e1, e1, n * 4, cudaMemcpyHostToDevice, h->stream);
        cudaMemcpyAsync(d_rel, rel, n * 4, cudaMemcpyHostToDevice, h->stream);
        cudaMemcpyAsync(d_e2, e2, n * 4, cudaMemcpyHostToDevice, h->stream);
        ntn_launch_score_eval(h, h->d_int_theta, d_e1, d_rel, d_e2, n, mode, d_s);
        cudaMemcpyAsync(scores, d_s, n * 8, cudaMemcpyDeviceToHost, h->stream);
        e = cudaStreamSynchronize(h->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
    }
    cudaFree(d_e1); cudaFree(d_rel); cudaFree(d_e2); cudaFree(d_s);
    if (e != cudaSuccess) FAIL(NTN_ECUDA, std::string("ntn_score: ") + cudaGetErrorString(e));
    return NTN_OK;
}
