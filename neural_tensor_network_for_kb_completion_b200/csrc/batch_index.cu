// Batch indexing on the device: the tables ntn_set_batch builds from the (e1, rel, e2, e3) columns of a batch job.
//
// The reference filters the batch list by relation on every evaluation (DataFactory.java:89-98, :614-645)
// and re-gathers entity columns per relation (NTN.java:129-143).  Here the grouping is done once per batch job:
//   1. columns are sorted (stable) by the packed key (rel, e1, e2)  -> the C corrupt copies of a training triple
//      (DataFactory.java:123-129) become one unique triple u with the list c_e3[c_off[u] .. c_off[u+1]);
//   2. every entity gets its pull list of incidents -- e1 roles (u ascending), e2 roles (u ascending), corrupt columns
//      (sorted-column order) -- by a stable sort of the concatenated role list on the entity id.
// Both sorts are stable LSD radix sorts (CUB), so the tables -- and with them every summation order of the evaluation --
// are a pure function of the batch columns: identical to what the host builder (ntn_capi.cu, NTN_SET_BATCH_HOST=1)
// produces, run to run and across devices.  The host only sees R+2 counters (for the tile tables) and the hub-segment
// count.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <string>

#include "ntn_internal.h"

namespace {

#define BCK(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string("batch index: ") + #call + ": " + cudaGetErrorString(e_);         \
            return NTN_ECUDA;                                                                      \
        }                                                                                          \
    } while (0)

inline int cdiv64(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// meta[0] = smallest offending column (INT_MAX if none), meta[1] = Nu, meta[2] = number of hub segments,
// meta[4 .. 4+R) = unique triples per relation
__global__ void k_bi_init(int* meta, int n_meta) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_meta) meta[i] = (i == 0) ? 0x7fffffff : 0;
}

__global__ void k_bi_keys(const int* __restrict__ e1, const int* __restrict__ rel, const int* __restrict__ e2,
                          const int* __restrict__ e3, int64_t n, int Ne, int R, int eb, uint64_t* __restrict__ key,
                          int* __restrict__ val, int* meta) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int a = e1[i], r = rel[i], b = e2[i], c = e3[i];
    const bool bad = a < 0 || a >= Ne || b < 0 || b >= Ne || c < 0 || c >= Ne || r < 0 || r >= R;
    if (bad) atomicMin(meta, (int)i);
    key[i] = bad ? ~0ull : (((uint64_t)r << (2 * eb)) | ((uint64_t)a << eb) | (uint64_t)b);
    val[i] = (int)i;
}

__global__ void k_bi_heads(const uint64_t* __restrict__ key, int64_t n, int* __restrict__ head) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}

__global__ void k_bi_unique(const uint64_t* __restrict__ key, const int* __restrict__ order, const int* __restrict__ scan,
                            const int* __restrict__ e3, int64_t n, int eb, int* __restrict__ u_e1, int* __restrict__ u_e2,
                            int* __restrict__ u_rel, int* __restrict__ c_off, int* __restrict__ c_e3,
                            int* __restrict__ col_u, int* meta) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int u = scan[i] - 1;
    col_u[i] = u;
    c_e3[i] = e3[order[i]];
    const uint64_t kx = key[i];
    if ((i == 0 || kx != key[i - 1]) && kx != ~0ull) {
        const uint64_t em = (1ull << eb) - 1;
        const int r = (int)(kx >> (2 * eb));
        u_e1[u] = (int)((kx >> eb) & em);
        u_e2[u] = (int)(kx & em);
        u_rel[u] = r;
        c_off[u] = (int)i;
        atomicAdd(meta + 4 + r, 1);                 // integer count: order-independent
    }
    if (i == n - 1) { c_off[u + 1] = (int)n; meta[1] = u + 1; }
}

// role list in pull order: [e1 roles | e2 roles | corrupt columns]
__global__ void k_bi_inc_keys(const int* __restrict__ u_e1, const int* __restrict__ u_e2, const int* __restrict__ c_e3,
                              int Nu, int64_t n, uint32_t* __restrict__ key, int* __restrict__ val) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * (int64_t)Nu + n) return;
    key[t] = (uint32_t)(t < Nu ? u_e1[t] : t < 2 * (int64_t)Nu ? u_e2[t - Nu] : c_e3[t - 2 * (int64_t)Nu]);
    val[t] = (int)t;
}

// sorted role list -> incident records + CSR offsets (an entity without incidents gets an empty range)
__global__ void k_bi_inc_records(const uint32_t* __restrict__ key, const int* __restrict__ val, const int* __restrict__ u_rel,
                                 const int* __restrict__ col_u, int Nu, int64_t n_inc, int Ne, int4* __restrict__ inc,
                                 int* __restrict__ inc_off) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_inc) return;
    const int prev = (i == 0) ? -1 : (int)key[i - 1];
    const int cur = (i == n_inc) ? Ne : (int)key[i];
    for (int e = prev + 1; e <= cur; ++e) inc_off[e] = (int)i;
    if (i == n_inc) return;
    const int t = val[i];
    int4 rec;
    if (t < Nu) rec = make_int4(t, t, 0, u_rel[t]);
    else if (t < 2 * Nu) rec = make_int4(t - Nu, t - Nu, 1, u_rel[t - Nu]);
    else { const int c = t - 2 * Nu, u = col_u[c]; rec = make_int4(u, Nu + c, 2, u_rel[u]); }
    inc[i] = rec;
}

// hub entities: pull lists longer than `thr` are cut in segments of `seg` incidents
__global__ void k_bi_hub_count(const int* __restrict__ inc_off, int Ne, int thr, int seg, int* __restrict__ count) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Ne) return;
    const int deg = inc_off[e + 1] - inc_off[e];
    count[e] = deg > thr ? (deg + seg - 1) / seg : 0;
}
__global__ void k_bi_hub_fill(const int* __restrict__ inc_off, const int* __restrict__ count, int* __restrict__ first, int Ne,
                              int seg, int* __restrict__ s_ent, int* __restrict__ s_begin, int* __restrict__ s_end, int* meta) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Ne) return;
    const int c = count[e], f = first[e];                // first = exclusive scan of count
    if (e == Ne - 1) meta[2] = f + c;
    if (c == 0) { first[e] = -1; return; }
    const int b0 = inc_off[e], b1 = inc_off[e + 1];
    for (int s = 0; s < c; ++s) {
        s_ent[f + s] = e;
        s_begin[f + s] = b0 + s * seg;
        s_end[f + s] = min(b0 + (s + 1) * seg, b1);
    }
}

// visiting order of the entity pull: entities by descending incident count (clamped), stable -> key = clamp - min(deg, clamp)
__global__ void k_bi_order_keys(const int* __restrict__ inc_off, int Ne, uint32_t* __restrict__ key, int* __restrict__ val) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Ne) return;
    key[e] = (uint32_t)(NTN_ORDER_CLAMP - min(inc_off[e + 1] - inc_off[e], NTN_ORDER_CLAMP));
    val[e] = e;
}

// ---- on-device batch-job generator ---------------------------------------------------------------
// Counter-based corrupt sampler: e3 of column j of batch job `job` is a pure function of (seed, job, j, attempt), so a
// batch job can be regenerated anywhere (tests restate it in NumPy, oracle/data_ref.py: device_corrupt_sample).
__host__ __device__ inline uint64_t bs_mix(uint64_t seed, uint64_t job, uint64_t j, uint64_t attempt) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (1ull + job);
    z ^= j * 0xBF58476D1CE4E5B9ull;
    z += attempt * 0x94D049BB133111EBull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
__device__ inline bool bs_contains(const uint64_t* __restrict__ keys, int64_t n, uint64_t k) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (keys[mid] < k) lo = mid + 1; else hi = mid; }
    return lo < n && keys[lo] == k;
}
// columns h-major like DataFactory.java:123-129: column j = h*batch + i copies training triple first+i
__global__ void k_bs_generate(const int* __restrict__ tr_e1, const int* __restrict__ tr_rel, const int* __restrict__ tr_e2,
                              int64_t first, int batch, int64_t n, int Ne, int eb, uint64_t seed, uint64_t job,
                              const uint64_t* __restrict__ true_keys, int64_t n_true, int max_attempts,
                              int* __restrict__ e1, int* __restrict__ rel, int* __restrict__ e2, int* __restrict__ e3) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int64_t t = first + (j % batch);
    const int a = tr_e1[t], r = tr_rel[t], b = tr_e2[t];
    int c = 0;
    for (int att = 0; att < max_attempts; ++att) {
        c = (int)(((bs_mix(seed, job, (uint64_t)j, (uint64_t)att) >> 32) * (uint64_t)Ne) >> 32);
        if (!true_keys) break;
        // filtered mode: neither (e1, rel, c) nor (c, rel, e2) may be a known true triple (the corrupt side is only
        // drawn at evaluation time, NTN.java:146, so both readings are excluded)
        const uint64_t kt = ((uint64_t)r << (2 * eb)) | ((uint64_t)a << eb) | (uint64_t)c;
        const uint64_t kh = ((uint64_t)r << (2 * eb)) | ((uint64_t)c << eb) | (uint64_t)b;
        if (!bs_contains(true_keys, n_true, kt) && !bs_contains(true_keys, n_true, kh)) break;
    }
    e1[j] = a; rel[j] = r; e2[j] = b; e3[j] = c;
}

__global__ void k_bs_true_keys(const int* __restrict__ e1, const int* __restrict__ rel, const int* __restrict__ e2, int64_t n,
                               int eb, uint64_t* __restrict__ key) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) key[i] = ((uint64_t)rel[i] << (2 * eb)) | ((uint64_t)e1[i] << eb) | (uint64_t)e2[i];
}

template <typename T> int grow(ntn_handle* h, T** p, size_t n) {
    if (n == 0) n = 1;
    const size_t bytes = n * sizeof(T);
    if (*p) {
        auto it = h->capacity.find((void*)*p);
        if (it != h->capacity.end() && it->second >= bytes) return NTN_OK;
        if (it != h->capacity.end()) h->capacity.erase(it);
        cudaFree(*p); *p = nullptr;
    }
    BCK(cudaMalloc((void**)p, bytes));
    h->capacity[(void*)*p] = bytes;
    return NTN_OK;
}

}  // namespace

// Device columns -> unique-triple tables, incident CSR and hub segments of h->b.  On return (stream synchronised)
// counts[0] = Nu, counts[1] = hub segments, rel_count[r] = unique triples of relation r.
int ntn_index_batch_dev(ntn_handle* h, int64_t n, const int* d_e1, const int* d_rel, const int* d_e2, const int* d_e3,
                        int* counts, int* rel_count) {
    const NtnDims& dm = h->dm;
    DeviceBatch& b = h->b;
    cudaStream_t st = h->stream;
    int eb = 1; while ((1 << eb) < dm.Ne) ++eb;
    int rb = 1; while ((1 << rb) < dm.R) ++rb;
    if (2 * eb + rb > 63) { h->err = "batch index: key (rel, e1, e2) does not fit 63 bits"; return NTN_EUNSUP; }
    const int n_meta = 4 + dm.R;
    const size_t n3 = (size_t)3 * (size_t)n;

    // workspace: sort double buffers, scan, temp storage (grow-only, kept in the handle)
    size_t tmp_sort64 = 0, tmp_sort32 = 0, tmp_scan = 0, tmp_scan_e = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort64, (uint64_t*)nullptr, (uint64_t*)nullptr, (int*)nullptr, (int*)nullptr,
                                    (int)n, 0, 2 * eb + rb, st);
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort32, (uint32_t*)nullptr, (uint32_t*)nullptr, (int*)nullptr, (int*)nullptr,
                                    (int)n3, 0, eb, st);
    cub::DeviceScan::InclusiveSum(nullptr, tmp_scan, (int*)nullptr, (int*)nullptr, (int)n, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan_e, (int*)nullptr, (int*)nullptr, dm.Ne, st);
    size_t tmp_sort_o = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort_o, (uint32_t*)nullptr, (uint32_t*)nullptr, (int*)nullptr, (int*)nullptr,
                                    dm.Ne, 0, 10, st);
    const size_t tmp_bytes = std::max(std::max(std::max(tmp_sort64, tmp_sort32), std::max(tmp_scan, tmp_scan_e)), tmp_sort_o);
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_key = 0, o_key2 = o_key + al((size_t)n * 8), o_val = o_key2 + al((size_t)n * 8),
                 o_val2 = o_val + al((size_t)n * 4), o_scan = o_val2 + al((size_t)n * 4), o_colu = o_scan + al((size_t)n * 4),
                 o_ik = o_colu + al((size_t)n * 4), o_ik2 = o_ik + al(n3 * 4), o_iv = o_ik2 + al(n3 * 4), o_iv2 = o_iv + al(n3 * 4),
                 o_meta = o_iv2 + al(n3 * 4), o_ok = o_meta + al((size_t)n_meta * 4), o_ok2 = o_ok + al((size_t)dm.Ne * 4),
                 o_ov = o_ok2 + al((size_t)dm.Ne * 4), o_tmp = o_ov + al((size_t)dm.Ne * 4), total = o_tmp + al(tmp_bytes);
    if (h->bi_ws_bytes < total) {
        if (h->bi_ws) cudaFree(h->bi_ws);
        h->bi_ws = nullptr; h->bi_ws_bytes = 0;
        BCK(cudaMalloc(&h->bi_ws, total));
        h->bi_ws_bytes = total;
    }
    if (!h->h_meta) BCK(cudaMallocHost((void**)&h->h_meta, (size_t)(4 + 4096) * sizeof(int)));
    if (n_meta > 4 + 4096) { h->err = "batch index: more than 4096 relations"; return NTN_EUNSUP; }
    char* ws = (char*)h->bi_ws;
    uint64_t *key = (uint64_t*)(ws + o_key), *key2 = (uint64_t*)(ws + o_key2);
    int *val = (int*)(ws + o_val), *val2 = (int*)(ws + o_val2), *scan = (int*)(ws + o_scan), *col_u = (int*)(ws + o_colu);
    uint32_t *ik = (uint32_t*)(ws + o_ik), *ik2 = (uint32_t*)(ws + o_ik2);
    int *iv = (int*)(ws + o_iv), *iv2 = (int*)(ws + o_iv2), *meta = (int*)(ws + o_meta);
    void* tmp = ws + o_tmp;

    int rc;
    if ((rc = grow(h, &b.u_e1, (size_t)n)) || (rc = grow(h, &b.u_e2, (size_t)n)) || (rc = grow(h, &b.u_rel, (size_t)n)) ||
        (rc = grow(h, &b.c_off, (size_t)n + 1)) || (rc = grow(h, &b.c_e3, (size_t)n)) || (rc = grow(h, &b.inc_off, (size_t)dm.Ne + 1)) ||
        (rc = grow(h, &b.inc, n3)) || (rc = grow(h, &b.eh_first, (size_t)dm.Ne)) || (rc = grow(h, &b.eh_count, (size_t)dm.Ne)) ||
        (rc = grow(h, &b.ent_order, (size_t)dm.Ne)))
        return rc;

    const int T = 256;
    k_bi_init<<<cdiv64(n_meta, T), T, 0, st>>>(meta, n_meta);
    size_t tb;
    if (n > 0) {
        k_bi_keys<<<cdiv64(n, T), T, 0, st>>>(d_e1, d_rel, d_e2, d_e3, n, dm.Ne, dm.R, eb, key, val, meta);
        tb = tmp_bytes;
        BCK(cub::DeviceRadixSort::SortPairs(tmp, tb, key, key2, val, val2, (int)n, 0, 2 * eb + rb, st));
        k_bi_heads<<<cdiv64(n, T), T, 0, st>>>(key2, n, val);                       // val is free again: head flags
        tb = tmp_bytes;
        BCK(cub::DeviceScan::InclusiveSum(tmp, tb, val, scan, (int)n, st));
        k_bi_unique<<<cdiv64(n, T), T, 0, st>>>(key2, val2, scan, d_e3, n, eb, b.u_e1, b.u_e2, b.u_rel, b.c_off, b.c_e3, col_u, meta);
    } else {
        BCK(cudaMemsetAsync(b.c_off, 0, sizeof(int), st));
    }
    BCK(cudaMemcpyAsync(h->h_meta, meta, (size_t)n_meta * sizeof(int), cudaMemcpyDeviceToHost, st));
    BCK(cudaStreamSynchronize(st));
    if (h->h_meta[0] != 0x7fffffff) {
        h->err = "entity or relation index out of range in column " + std::to_string(h->h_meta[0]);
        return NTN_EINVAL;
    }
    const int Nu = h->h_meta[1];
    for (int r = 0; r < dm.R; ++r) rel_count[r] = h->h_meta[4 + r];
    const int64_t n_inc = 2 * (int64_t)Nu + n;
    if (n_inc > 0) {
        k_bi_inc_keys<<<cdiv64(n_inc, T), T, 0, st>>>(b.u_e1, b.u_e2, b.c_e3, Nu, n, ik, iv);
        tb = tmp_bytes;
        BCK(cub::DeviceRadixSort::SortPairs(tmp, tb, ik, ik2, iv, iv2, (int)n_inc, 0, eb, st));
    }
    k_bi_inc_records<<<cdiv64(n_inc + 1, T), T, 0, st>>>(ik2, iv2, b.u_rel, col_u, Nu, n_inc, dm.Ne, b.inc, b.inc_off);
    {
        uint32_t *ok = (uint32_t*)(ws + o_ok), *ok2 = (uint32_t*)(ws + o_ok2);
        int* ov = (int*)(ws + o_ov);
        k_bi_order_keys<<<cdiv64(dm.Ne, T), T, 0, st>>>(b.inc_off, dm.Ne, ok, ov);
        tb = tmp_bytes;
        BCK(cub::DeviceRadixSort::SortPairs(tmp, tb, ok, ok2, ov, b.ent_order, dm.Ne, 0, 10, st));
    }
    k_bi_hub_count<<<cdiv64(dm.Ne, T), T, 0, st>>>(b.inc_off, dm.Ne, NTN_HUB_SEG, NTN_HUB_SEG, b.eh_count);
    tb = tmp_bytes;
    BCK(cub::DeviceScan::ExclusiveSum(tmp, tb, b.eh_count, b.eh_first, dm.Ne, st));
    // segment tables: a hub list has deg > NTN_HUB_SEG incidents and ceil(deg / NTN_HUB_SEG) <= 2 deg / NTN_HUB_SEG segments
    const size_t seg_cap = (size_t)(2 * n_inc / NTN_HUB_SEG) + 1;
    if ((rc = grow(h, &b.eh_ent, seg_cap)) || (rc = grow(h, &b.eh_begin, seg_cap)) || (rc = grow(h, &b.eh_end, seg_cap))) return rc;
    k_bi_hub_fill<<<cdiv64(dm.Ne, T), T, 0, st>>>(b.inc_off, b.eh_count, b.eh_first, dm.Ne, NTN_HUB_SEG, b.eh_ent, b.eh_begin, b.eh_end, meta);
    BCK(cudaMemcpyAsync(h->h_meta + 2, meta + 2, sizeof(int), cudaMemcpyDeviceToHost, st));
    BCK(cudaStreamSynchronize(st));
    BCK(cudaGetLastError());
    counts[0] = Nu;
    counts[1] = h->h_meta[2];
    return NTN_OK;
}

// Sorted key set of the training triples (for the filtered sampler).  d_train = [e1 | rel | e2], T each, validated by the caller.
int ntn_build_true_keys_dev(ntn_handle* h, int64_t T, const int* d_train, uint64_t** d_keys) {
    const NtnDims& dm = h->dm;
    cudaStream_t st = h->stream;
    int eb = 1; while ((1 << eb) < dm.Ne) ++eb;
    int rb = 1; while ((1 << rb) < dm.R) ++rb;
    if (2 * eb + rb > 63) { h->err = "batch index: key (rel, e1, e2) does not fit 63 bits"; return NTN_EUNSUP; }
    uint64_t* tmp_keys = nullptr; void* tmp = nullptr; size_t tb = 0;
    const size_t nn = (size_t)std::max<int64_t>(T, 1);
    if (*d_keys) { cudaFree(*d_keys); *d_keys = nullptr; }
    BCK(cudaMalloc((void**)d_keys, nn * 8));
    BCK(cudaMalloc((void**)&tmp_keys, nn * 8));
    cub::DeviceRadixSort::SortKeys(nullptr, tb, tmp_keys, *d_keys, (int)T, 0, 2 * eb + rb, st);
    cudaError_t e = cudaMalloc(&tmp, std::max<size_t>(tb, 1));
    if (e == cudaSuccess && T > 0) {
        k_bs_true_keys<<<cdiv64(T, 256), 256, 0, st>>>(d_train, d_train + T, d_train + 2 * T, T, eb, tmp_keys);
        e = cub::DeviceRadixSort::SortKeys(tmp, tb, tmp_keys, *d_keys, (int)T, 0, 2 * eb + rb, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(tmp_keys); cudaFree(tmp);
    if (e != cudaSuccess) { h->err = std::string("batch index: true-triple keys: ") + cudaGetErrorString(e); return NTN_ECUDA; }
    return NTN_OK;
}

// Fill d_cols = [e1 | rel | e2 | e3] (n = batch*corrupt each) from training triples [first, first+batch).
int ntn_generate_columns_dev(ntn_handle* h, const int* d_train, int64_t T, int64_t first, int batch, int corrupt, uint64_t seed,
                             uint64_t job, const uint64_t* d_true_keys, int64_t n_true, int* d_cols) {
    const NtnDims& dm = h->dm;
    int eb = 1; while ((1 << eb) < dm.Ne) ++eb;
    const int64_t n = (int64_t)batch * corrupt;
    if (n > 0)
        k_bs_generate<<<cdiv64(n, 256), 256, 0, h->stream>>>(d_train, d_train + T, d_train + 2 * T, first, batch, n, dm.Ne, eb, seed,
                                                            job, d_true_keys, n_true, d_true_keys ? 16 : 1,
                                                            d_cols, d_cols + n, d_cols + 2 * n, d_cols + 3 * n);
    BCK(cudaGetLastError());
    return NTN_OK;
}
