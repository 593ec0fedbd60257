// Hand-written sm_100a kernels of the NTN cost+gradient path: everything except the tcgen05
// contractions (gemm_tcgen05.cu).  Each kernel cites the reference statements it replaces
// (/root/reference/src/NTN.java unless another file is named) and DESIGN.md gives its roofline.
//
// Formulation (SURVEY.md App. A "dedupe identity"): for a unique training triple u
// of relation r with anchor entity A (e1 if the relation corrupts the tail, NTN.java:146-150,
// else e2) and other entity O:
//   T'_u[s,:] = [E_A | 1] * Waug_r  = W_r[s]-contracted anchor row + the V half that multiplies
//               the non-anchor side; beta_u[s] = anchor-side V term + b      (NTN.java:163-197)
//   pre+ = T'_u[s,:].E_O + beta,  pre-_c = T'_u[s,:].E_{e3_c} + beta          (per corrupt column c)
//   M_u[s,:] = a_{u,s} E_O + sum_c c_{c,s} E_{e3_c},  coefsum = a + sum_c c   (NTN.java:280-388)
//   dWaug_r = [E_A | 1]^T [M | coefsum]  -> dW, dV, db                        (NTN.java:301-349)
//   GA_u = [M | coefsum] Waug_r^T        -> gradient of the anchor entity     (NTN.java:352-388)
//   gE[O] += sum_s a T',  gE[e3_c] += sum_s c_c T'                            (Util.java:74-98)
#include <assert.h>
#include <algorithm>
#include <math.h>
#include <stdio.h>

#include "ntn_internal.h"

#define FULL 0xffffffffu

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ float dot4_acc(const float4& a, const float4& b, float acc) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
}
__device__ __forceinline__ void fma4(float4& acc, float s, const float4& v) {
    acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y);
    acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}
// store of an M element group: plain, or split into TF32 hi / lo arrays when the contractions take M pre-split
__device__ __forceinline__ void store_m4(float* Mhi, float* Mlo, size_t off, const float4& v) {
    if (Mlo) {
        float4 hi, lo;
        split_tf32(v, hi, lo);
        *reinterpret_cast<float4*>(Mhi + off) = hi;
        *reinterpret_cast<float4*>(Mlo + off) = lo;
    } else {
        *reinterpret_cast<float4*>(Mhi + off) = v;
    }
}
__device__ __forceinline__ void store_m1(float* Mhi, float* Mlo, size_t off, float v) {
    if (Mlo) { const float hi = tf32_round(v); Mhi[off] = hi; Mlo[off] = tf32_round(v - hi); }
    else Mhi[off] = v;
}
// store of one float4 group of an entity-gradient row; PUSH (data-parallel entity-gradient exchange, exchange.cu): a group
// that another rank owns also goes straight into that rank's scratch -- a posted store over NVLink from the kernel that
// produces it
template <bool PUSH>
__device__ __forceinline__ void ge_store(float* gE, const XPushRange& xp, size_t o, const float4& v) {
    *reinterpret_cast<float4*>(gE + o) = v;
    if (PUSH) {
        const unsigned rel4 = (unsigned)((o - (size_t)xp.off) >> 2);
        const unsigned owner = fd_div(rel4, FastDiv{xp.per4, xp.per_m, xp.per_s});
        if ((int)owner != xp.rank) {
            float* dst = xp.dst[0];
#pragma unroll
            for (int s = 1; s < NTN_MAX_PEERS; ++s) dst = (owner == (unsigned)s) ? xp.dst[s] : dst;
            *reinterpret_cast<float4*>(dst + 4 * (size_t)(rel4 - owner * xp.per4)) = v;
        }
    }
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// test hook (CPU): q[i] = n[i] / d through the same magic numbers the kernels use
extern "C" int ntn_host_fastdiv(uint32_t d, int64_t count, const uint32_t* n, uint32_t* q) {
    if (d == 0 || count < 0 || (count > 0 && (!n || !q))) return NTN_EINVAL;
    const FastDiv f = make_fastdiv(d);
    for (int64_t i = 0; i < count; ++i) q[i] = fd_div(n[i], f);
    return NTN_OK;
}
__device__ __forceinline__ unsigned flat_index() { return blockIdx.x * blockDim.x + threadIdx.x; }

template <int ACT> __device__ __forceinline__ float act_fn(float x) {
    if (ACT == NTN_ACT_TANH) return tanhf(x);               // NTN.java:200-201
    return 1.0f / (1.0f + expf(-x));
}
template <int ACT> __device__ __forceinline__ float act_diff(float z) {   // NTN.java:744-759
    if (ACT == NTN_ACT_TANH) return 1.0f - z * z;
    return z * (1.0f - z);
}

// ==========================================================================================
// K1  entity build: segmented mean of word rows               DataFactory.java:393-437
//     E[n, 0:d] = (1/cnt) sum_j WVt[word_j(n), :],  E[n, d] = 1,  E[n, d+1:Kp] = 0
//     One thread per (entity, float4 chunk).  HBM-bound: reads nnz*dq, writes Ne*Kp.
// ==========================================================================================
__device__ __forceinline__ float4 entity_mean_chunk(const float* __restrict__ wvt, const int4& r, const int* __restrict__ idx,
                                                    int c, int d, int dq) {
    const int cnt = r.y;
    float4 acc = zero4();
    if (c * 4 < dq) {
        if (cnt <= 2) {
            acc = ldg4(wvt + (size_t)r.z * dq + c * 4);
            if (cnt == 2) { float4 v = ldg4(wvt + (size_t)r.w * dq + c * 4); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        } else {
            for (int jb = r.x; jb < r.x + cnt; jb += 4) {            // word order, like :410-418
                float4 v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    v[i] = (jb + i < r.x + cnt) ? ldg4(wvt + (size_t)__ldg(idx + jb + i) * dq + c * 4) : zero4();
#pragma unroll
                for (int i = 0; i < 4; ++i) { acc.x += v[i].x; acc.y += v[i].y; acc.z += v[i].z; acc.w += v[i].w; }
            }
        }
        if (cnt > 1) {                                               // :420 (single word: copy, :429)
            float fc = (float)cnt;
            acc.x /= fc; acc.y /= fc; acc.z /= fc; acc.w /= fc;
        }
    }
    float* av = reinterpret_cast<float*>(&acc);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int p = c * 4 + i;
        if (p == d) av[i] = 1.0f; else if (p > d) av[i] = 0.0f;
    }
    return acc;
}

__global__ void __launch_bounds__(256) k_entity_build(const float* __restrict__ wvt, const int4* __restrict__ rec,
                                                      const int* __restrict__ idx, float* __restrict__ E,
                                                      int Ne, int half, int d, int dq, int Kp, FastDiv fd) {
    // flat mapping: one thread per (entity pair, float4 chunk of the Kp-wide row): entities n and n + half, so that two
    // independent record -> word-row chains are in flight per thread (the kernel is a few waves of two dependent
    // round trips).  rec = {first index into idx, word count, word 0, word 1}: entities of one or two words (92 %)
    // need no second dependent index load.
    const unsigned gid = flat_index();
    const int n = (int)fd_div(gid, fd), c = (int)(gid - (unsigned)n * fd.d);       // fd.d = Kp / 4 chunks per row
    if (n >= half) return;
    const int n2 = n + half;
    const int4 r1 = __ldg(rec + n);
    const int4 r2 = (n2 < Ne) ? __ldg(rec + n2) : r1;
    const float4 a1 = entity_mean_chunk(wvt, r1, idx, c, d, dq);
    const float4 a2 = entity_mean_chunk(wvt, r2, idx, c, d, dq);
    *reinterpret_cast<float4*>(E + (size_t)n * Kp + c * 4) = a1;
    if (n2 < Ne) *reinterpret_cast<float4*>(E + (size_t)n2 * Kp + c * 4) = a2;
}

// ==========================================================================================
// K2  Waug_r [Kp x Np] from theta and this evaluation's corrupt sides.
//     tail (side=1):  Waug[p][s*dq+q] = W_r[s][p][q]   Waug[d][s*dq+q] = V_r[d+q][s]
//                     Waug[p][k*dq+s] = V_r[p][s]      Waug[d][k*dq+s] = b_r[s]
//     head (side=0):  Waug[p][s*dq+q] = W_r[s][q][p]   Waug[d][s*dq+q] = V_r[q][s]
//                     Waug[p][k*dq+s] = V_r[d+p][s]    Waug[d][k*dq+s] = b_r[s]
//     (replaces the element-wise slice copies of NTN.java:166-169,376-383)
// ==========================================================================================
__global__ void __launch_bounds__(256) k_w_prep(const float* __restrict__ theta, const uint8_t* __restrict__ side,
                                                float* __restrict__ Waug, NtnDims dm) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t per = (int64_t)dm.Kp * dm.Np;
    if (i >= per * dm.R) return;
    int r = (int)(i / per);
    int rem = (int)(i - (int64_t)r * per);
    int p = rem / dm.Np, n = rem - p * dm.Np;
    int d = dm.d, k = dm.k, dq = dm.dq;
    bool tail = side[r] != 0;
    const float* W = theta + dm.oW + (int64_t)r * k * d * d;
    const float* V = theta + dm.oV + (int64_t)r * 2 * d * k;
    float v = 0.f;
    if (n < k * dq) {
        int s = n / dq, q = n - s * dq;
        if (q < d) {
            if (p < d) v = tail ? W[((int64_t)s * d + p) * d + q] : W[((int64_t)s * d + q) * d + p];
            else if (p == d) v = tail ? V[(d + q) * k + s] : V[q * k + s];
        }
    } else if (n < k * dq + k) {
        int s = n - k * dq;
        if (p < d) v = tail ? V[p * k + s] : V[(d + p) * k + s];
        else if (p == d) v = theta[dm.ob + (int64_t)r * k + s];
    }
    Waug[i] = v;
}

// ==========================================================================================
// SIMT FP32 contractions (the A/B reference path for the tcgen05 kernels; NTN_GEMM_SIMT).
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile.
// ==========================================================================================
#define GBM 64
#define GBN 64
#define GBK 16
#define GPAD 68

struct GemmTables {
    const int *t_rel, *t_u0, *t_n;      // tile / chunk table
    const int *u_e1, *u_e2;
    const uint8_t* side;
};

__device__ __forceinline__ int anchor_of(const GemmTables& tb, int u, bool tail) {
    return tail ? __ldg(tb.u_e1 + u) : __ldg(tb.u_e2 + u);
}

__device__ __forceinline__ void gemm_compute_tile(const float (*As)[GPAD], const float (*Bs)[GPAD], float acc[4][4],
                                                  int ty, int tx) {
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
        float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}

// forward: T'[u0+m][n] = sum_p E[anchor(u0+m)][p] * Waug_r[p][n]     (NTN.java:163-197)
// grid.x = tiles*2 (64-row halves of the 128-row score tile), grid.y = ceil(Np/64)
__global__ void __launch_bounds__(256) k_gemm_fwd_simt(GemmTables tb, const float* __restrict__ E,
                                                       const float* __restrict__ Waug, float* __restrict__ T,
                                                       int Kp, int Np) {
    __shared__ __align__(16) float As[GBK][GPAD];
    __shared__ __align__(16) float Bs[GBK][GPAD];
    int tile = blockIdx.x >> 1, half = blockIdx.x & 1;
    int r = tb.t_rel[tile], u0 = tb.t_u0[tile] + half * GBM, rows = tb.t_n[tile] - half * GBM;
    if (rows <= 0) return;
    if (rows > GBM) rows = GBM;
    bool tail = tb.side[r] != 0;
    int n0 = blockIdx.y * GBN;
    const float* B = Waug + (size_t)r * Kp * Np;
    int t = threadIdx.x, ty = t >> 4, tx = t & 15;
    int am = t >> 2, akq = (t & 3) * 4;            // A: k-contiguous rows, gathered
    int bk = t >> 4, bnq = (t & 15) * 4;           // B: n-contiguous
    const float* arow = nullptr;
    if (am < rows) arow = E + (size_t)anchor_of(tb, u0 + am, tail) * Kp;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < Kp; k0 += GBK) {
        float4 av = zero4(), bv = zero4();
        if (arow && k0 + akq < Kp) av = ldg4(arow + k0 + akq);
        if (k0 + bk < Kp && n0 + bnq < Np) bv = ldg4(B + (size_t)(k0 + bk) * Np + n0 + bnq);
        As[akq + 0][am] = av.x; As[akq + 1][am] = av.y; As[akq + 2][am] = av.z; As[akq + 3][am] = av.w;
        *reinterpret_cast<float4*>(&Bs[bk][bnq]) = bv;
        __syncthreads();
        gemm_compute_tile(As, Bs, acc, ty, tx);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = ty * 4 + i;
        if (m < rows && n0 + tx * 4 < Np)
            *reinterpret_cast<float4*>(T + (size_t)(u0 + m) * Np + n0 + tx * 4) =
                make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
}

// anchor-entity gradient rows: GA[u0+m][p] = sum_n M[u0+m][n] * Waug_r[p][n]   (NTN.java:352-388)
__global__ void __launch_bounds__(256) k_gemm_ge_simt(GemmTables tb, const float* __restrict__ M,
                                                      const float* __restrict__ Waug, float* __restrict__ GA,
                                                      int Kp, int Np) {
    __shared__ __align__(16) float As[GBK][GPAD];
    __shared__ __align__(16) float Bs[GBK][GPAD];
    int tile = blockIdx.x >> 1, half = blockIdx.x & 1;
    int r = tb.t_rel[tile], u0 = tb.t_u0[tile] + half * GBM, rows = tb.t_n[tile] - half * GBM;
    if (rows <= 0) return;
    if (rows > GBM) rows = GBM;
    int n0 = blockIdx.y * GBN;                     // over p in [0,Kp)
    const float* B = Waug + (size_t)r * Kp * Np;
    int t = threadIdx.x, ty = t >> 4, tx = t & 15;
    int am = t >> 2, akq = (t & 3) * 4;            // A = M rows, k(=n')-contiguous
    int bn = t >> 2, bkq = (t & 3) * 4;            // B(k=n', n=p) = Waug[p][n'], k-contiguous
    const float* arow = (am < rows) ? M + (size_t)(u0 + am) * Np : nullptr;
    const float* brow = (n0 + bn < Kp) ? B + (size_t)(n0 + bn) * Np : nullptr;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < Np; k0 += GBK) {
        float4 av = zero4(), bv = zero4();
        if (arow && k0 + akq < Np) av = ldg4(arow + k0 + akq);
        if (brow && k0 + bkq < Np) bv = ldg4(brow + k0 + bkq);
        As[akq + 0][am] = av.x; As[akq + 1][am] = av.y; As[akq + 2][am] = av.z; As[akq + 3][am] = av.w;
        Bs[bkq + 0][bn] = bv.x; Bs[bkq + 1][bn] = bv.y; Bs[bkq + 2][bn] = bv.z; Bs[bkq + 3][bn] = bv.w;
        __syncthreads();
        gemm_compute_tile(As, Bs, acc, ty, tx);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = ty * 4 + i;
        if (m < rows && n0 + tx * 4 < Kp)
            *reinterpret_cast<float4*>(GA + (size_t)(u0 + m) * Kp + n0 + tx * 4) =
                make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
}

// dWaug split-K partial: part[chunk][n][p] = sum_{u in chunk} E[anchor(u)][p] * M[u][n]   (NTN.java:301-349)
// grid.x = chunks, grid.y = ceil(Kp/64) * ceil(Np/64)
__global__ void __launch_bounds__(256) k_gemm_dw_simt(GemmTables tb, const float* __restrict__ E,
                                                      const float* __restrict__ M, float* __restrict__ part,
                                                      int Kp, int Np) {
    __shared__ __align__(16) float As[GBK][GPAD];
    __shared__ __align__(16) float Bs[GBK][GPAD];
    __shared__ int s_anchor[GBK];
    int ch = blockIdx.x;
    int r = tb.t_rel[ch], u0 = tb.t_u0[ch], rows = tb.t_n[ch];
    bool tail = tb.side[r] != 0;
    int ntn = (Np + GBN - 1) / GBN;
    int m0 = (blockIdx.y / ntn) * GBM, n0 = (blockIdx.y % ntn) * GBN;
    int t = threadIdx.x, ty = t >> 4, tx = t & 15;
    int ak = t >> 4, amq = (t & 15) * 4;           // A(m=p, k=u): m-contiguous, rows gathered over k
    int bk = t >> 4, bnq = (t & 15) * 4;           // B(k=u, n): n-contiguous
    float acc[4][4] = {};
    for (int k0 = 0; k0 < rows; k0 += GBK) {
        if (t < GBK) s_anchor[t] = (k0 + t < rows) ? anchor_of(tb, u0 + k0 + t, tail) : -1;
        __syncthreads();
        float4 av = zero4(), bv = zero4();
        int an = s_anchor[ak];
        if (an >= 0 && m0 + amq < Kp) av = ldg4(E + (size_t)an * Kp + m0 + amq);
        if (k0 + bk < rows && n0 + bnq < Np) bv = ldg4(M + (size_t)(u0 + k0 + bk) * Np + n0 + bnq);
        *reinterpret_cast<float4*>(&As[ak][amq]) = av;
        *reinterpret_cast<float4*>(&Bs[bk][bnq]) = bv;
        __syncthreads();
        gemm_compute_tile(As, Bs, acc, ty, tx);
        __syncthreads();
    }
    // stored transposed, [chunk][n][p] (p contiguous): the layout the tcgen05 epilogue writes coalesced
    float* C = part + (size_t)ch * Kp * Np;
    const int m = m0 + ty * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int n = n0 + tx * 4 + j;
        if (n < Np && m < Kp)
            *reinterpret_cast<float4*>(C + (size_t)n * Kp + m) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
    }
}

// ==========================================================================================
// K4  score / hinge / backward-coefficient kernel (the fused epilogue of the forward contraction)
//     NTN.java:200-222 (activation, u-dot, hinge, cost, update), :272-289 (dU, temp, db),
//     and the column-wise halves of :301-388 folded into M.
//     One block per NTN_SCORE_ROWS-row tile of one relation; one warp per unique triple at a time,
//     two corrupt columns in flight; lanes over float4 chunks of the d-vector.  Gathers of E rows
//     are L2 hits (E is 30 MB vs 126 MB L2).
// ==========================================================================================
struct ScoreArgs {
    const float *theta, *E, *T;
    float *M, *M_lo, *acoef, *ccoef;
    float* GV;                 // (Nu + N) x dq: per-incident contribution rows for the entity pull
    int Nu;
    const int *t_rel, *t_u0, *t_n, *u_e1, *u_e2, *c_off, *c_e3;
    const uint8_t* side;
    double* tp_cost; int* tp_nact; float* tp_dU;
    int d, dq, Kp, Np;
    int64_t oU;
};

// Transpose-reduce: every lane holds 32 partial sums v[0..31]; afterwards v[0] on lane l is the
// warp-wide total of partial sum number l.  31 shuffles instead of 32 x 5.
__device__ __forceinline__ void warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const float send = up ? v[i] : v[i + o];
            const float keep = up ? v[i + o] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL, send, o);
        }
    }
}

template <int KS, int CH, int ACT, bool GVW>
__global__ void __launch_bounds__(256, (KS * CH > 8) ? 1 : 2) k_score(ScoreArgs a) {   // wide rows (k = 8, d = 256): one block per SM, 255 registers, no spill
    // corrupt columns per batch: their KS pre-activations each fill the 32 lanes (KS=3 -> 10 columns,
    // the reference's corrupt_size); all row gathers of a batch are issued before any is used.
    constexpr int COLB = (CH == 1) ? ((32 / KS) < 10 ? (32 / KS) : 10) : ((16 / KS) < 1 ? 1 : ((16 / KS) < 5 ? (16 / KS) : 5));
    constexpr int V = COLB * KS;
    __shared__ double s_cost[8];
    __shared__ int s_nact[8];
    __shared__ float s_dU[8][32];
    const int tile = blockIdx.x;
    const int r = a.t_rel[tile], u0 = a.t_u0[tile], rows = a.t_n[tile];
    const bool tail = a.side[r] != 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dq = a.dq, d = a.d;
    float U[KS];
#pragma unroll
    for (int s = 0; s < KS; ++s) U[s] = __ldg(a.theta + a.oU + (int64_t)r * KS + s);
    const int my_s = lane % KS, my_j = lane / KS;       // lane l owns (column my_j, slice my_s) of a batch
    const float my_U = U[0] * (my_s == 0) + [&] { float t = 0.f;
#pragma unroll
        for (int s = 1; s < KS; ++s) t += U[s] * (my_s == s); return t; }();
    // entity rows carry the augmentation 1.0 at index d (and padding after it): when d is not a multiple of 4 the last
    // float4 chunk of a row reaches into it and is masked; otherwise (d = 100) no chunk does
    const bool masked = dq != d;
    bool cv[CH]; float4 cm[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        int p = (lane + 32 * c) * 4;
        cv[c] = p < dq;
        cm[c] = make_float4(p < d ? 1.f : 0.f, p + 1 < d ? 1.f : 0.f, p + 2 < d ? 1.f : 0.f, p + 3 < d ? 1.f : 0.f);
    }
    // lane-local accumulators, reduced once at the end of the tile
    float cost_l = 0.f, dU_l = 0.f; int nact_w = 0;

    // the index data of a warp's NEXT triple (other entity, corrupt range, the first 32 corrupt ids -- one per lane) is
    // fetched while the current one is computed: the dependent chain per triple is then one round trip (rows of E and
    // T'), not c_off -> c_e3 -> rows
    const int* oth = tail ? a.u_e2 : a.u_e1;
    int nx_other = 0, nx_c0 = 0, nx_c1 = 0, nx_e3 = 0;
    if (warp < rows) {
        nx_other = __ldg(oth + u0 + warp); nx_c0 = __ldg(a.c_off + u0 + warp); nx_c1 = __ldg(a.c_off + u0 + warp + 1);
        nx_e3 = (nx_c0 + lane < nx_c1) ? __ldg(a.c_e3 + nx_c0 + lane) : 0;
    }
    for (int i = warp; i < rows; i += 8) {
        const int u = u0 + i;
        const int other = nx_other, c0 = nx_c0, c1 = nx_c1, my_e3 = nx_e3;
        if (i + 8 < rows) {
            nx_other = __ldg(oth + u + 8); nx_c0 = __ldg(a.c_off + u + 8); nx_c1 = __ldg(a.c_off + u + 9);
        }
        const float* Trow = a.T + (size_t)u * a.Np;
        float4 t[KS][CH]; float beta[KS];
#pragma unroll
        for (int s = 0; s < KS; ++s) {
#pragma unroll
            for (int c = 0; c < CH; ++c) t[s][c] = cv[c] ? ldg4(Trow + s * dq + (lane + 32 * c) * 4) : zero4();
            beta[s] = __ldg(Trow + KS * dq + s);
        }
        float my_beta = 0.f;
#pragma unroll
        for (int s = 0; s < KS; ++s) my_beta += beta[s] * (my_s == s);
        float4 eo[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            eo[c] = cv[c] ? ldg4(a.E + (size_t)other * a.Kp + (lane + 32 * c) * 4) : zero4();
            if (masked) { eo[c].x *= cm[c].x; eo[c].y *= cm[c].y; eo[c].z *= cm[c].z; eo[c].w *= cm[c].w; }
        }
        // positive column (shared by all corrupt copies of this triple)
        float zp[KS], tp[KS], sp = 0.f;
        {
            float acc[KS];
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                acc[s] = 0.f;
#pragma unroll
                for (int c = 0; c < CH; ++c) acc[s] += dot4(t[s][c], eo[c]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int s = 0; s < KS; ++s) acc[s] += __shfl_xor_sync(FULL, acc[s], o);
            // lane s evaluates the activation of slice s, then everyone reads all KS of them
            float mine = 0.f;
#pragma unroll
            for (int s = 0; s < KS; ++s) mine += (acc[s] + beta[s]) * (lane == s);
            const float zmine = act_fn<ACT>(mine);
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                zp[s] = __shfl_sync(FULL, zmine, s);
                sp = fmaf(U[s], zp[s], sp);                           // :204
                tp[s] = act_diff<ACT>(zp[s]) * U[s];                  // :280
            }
        }
        // the next triple's corrupt ids: its range (requested at the top of this iteration) has arrived by now, and the
        // ids have the whole corrupt-column loop below to land
        if (i + 8 < rows) nx_e3 = (nx_c0 + lane < nx_c1) ? __ldg(a.c_e3 + nx_c0 + lane) : 0;
        float my_zp = 0.f;
#pragma unroll
        for (int s = 0; s < KS; ++s) my_zp += zp[s] * (my_s == s);
        float4 Macc[KS][CH]; float csum[KS];
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            csum[s] = 0.f;
#pragma unroll
            for (int c = 0; c < CH; ++c) Macc[s][c] = zero4();
        }
        int cnt = 0;
        for (int col = c0; col < c1; col += COLB) {
            const int nb = min(COLB, c1 - col);
            float4 ev[COLB][CH];
#pragma unroll
            for (int j = 0; j < COLB; ++j) {
                const bool ok = j < nb;
                const int ci = col + j - c0;                              // warp-uniform
                const int e3s = __shfl_sync(FULL, my_e3, ci & 31);
                const int e3 = ok ? (ci < 32 ? e3s : __ldg(a.c_e3 + col + j)) : 0;
#pragma unroll
                for (int c = 0; c < CH; ++c)
                    ev[j][c] = (ok && cv[c]) ? ldg4(a.E + (size_t)e3 * a.Kp + (lane + 32 * c) * 4) : zero4();
            }
            float v[32];
#pragma unroll
            for (int q = V; q < 32; ++q) v[q] = 0.f;                      // slots beyond the batch's COLB x KS values
#pragma unroll
            for (int j = 0; j < COLB; ++j)
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    if (masked) { ev[j][c].x *= cm[c].x; ev[j][c].y *= cm[c].y; ev[j][c].z *= cm[c].z; ev[j][c].w *= cm[c].w; }
#pragma unroll
                    for (int s = 0; s < KS; ++s) {                        // first chunk assigns: no zero-fill + add
                        if (c == 0) v[j * KS + s] = dot4(t[s][c], ev[j][c]);
                        else v[j * KS + s] += dot4(t[s][c], ev[j][c]);
                    }
                }
            warp_transpose_reduce32(v, lane);                          // lane l: pre-activation (my_j, my_s)
            const bool mine_valid = lane < V && my_j < nb;
            const float zn = act_fn<ACT>(v[0] + my_beta);              // one activation per lane
            // negative score of column j = sum over its KS lanes (NTN.java:205)
            float sn = my_U * zn;
#pragma unroll
            for (int s = 1; s < KS; ++s) sn += __shfl_down_sync(FULL, my_U * zn, s);
            const bool active_here = mine_valid && my_s == 0 && ((sp + 1.0f) > sn);   // :213 (strict, FP32)
            const unsigned amask = __ballot_sync(FULL, active_here);   // bit j*KS = column j active
            const bool my_active = mine_valid && ((amask >> (my_j * KS)) & 1u);
            if (active_here) cost_l += (sp - sn) + 1.0f;               // :221
            const int na = __popc(amask);
            nact_w += na; cnt += na;                                   // :222
            const float my_cc = my_active ? -act_diff<ACT>(zn) * my_U : 0.f;          // :281
            if (my_active) dU_l += my_zp - zn;                         // :273
            if (mine_valid) a.ccoef[(size_t)col * KS + lane] = my_cc;  // coalesced: (col+j)*KS + s = col*KS + lane
#pragma unroll
            for (int j = 0; j < COLB; ++j) {
                float4 gv[CH];
#pragma unroll
                for (int c = 0; c < CH; ++c) gv[c] = zero4();
#pragma unroll
                for (int s = 0; s < KS; ++s) {
                    const float cc = __shfl_sync(FULL, my_cc, j * KS + s);
                    csum[s] += cc;
#pragma unroll
                    for (int c = 0; c < CH; ++c) { fma4(Macc[s][c], cc, ev[j][c]); if (GVW) fma4(gv[c], cc, t[s][c]); }
                }
                // gradient row of the corrupt entity of column col+j: sum_s c_{c,s} T'_u[s,:]  (zero if inactive);
                // materialised here, where T' is in registers, so the entity pull reads ONE row per incident
                if (GVW && j < nb) {
#pragma unroll
                    for (int c = 0; c < CH; ++c)
                        if (cv[c]) *reinterpret_cast<float4*>(a.GV + (size_t)(a.Nu + col + j) * dq + (lane + 32 * c) * 4) = gv[c];
                }
            }
        }
        const size_t Mrow = (size_t)u * a.Np;
        float4 gvo[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) gvo[c] = zero4();
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            float as = (float)cnt * tp[s];                            // positive side: one per active copy
            csum[s] += as;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                fma4(Macc[s][c], as, eo[c]);
                if (GVW) fma4(gvo[c], as, t[s][c]);
                if (cv[c]) store_m4(a.M, a.M_lo, Mrow + s * dq + (lane + 32 * c) * 4, Macc[s][c]);
            }
            if (lane == 0) a.acoef[(size_t)u * KS + s] = as;
        }
        // gradient row of the non-anchor positive entity: sum_s a_{u,s} T'_u[s,:]
        if (GVW) {
#pragma unroll
            for (int c = 0; c < CH; ++c)
                if (cv[c]) *reinterpret_cast<float4*>(a.GV + (size_t)u * dq + (lane + 32 * c) * 4) = gvo[c];
        }
        // coefficient sums ride in the k slots after the slices; zero the rest of the row
        for (int n = KS * dq + lane; n < a.Np; n += 32) {
            float vv = 0.f;
#pragma unroll
            for (int s = 0; s < KS; ++s) if (n == KS * dq + s) vv = csum[s];
            store_m1(a.M, a.M_lo, Mrow + n, vv);
        }
    }
    // tile partials, fixed order: lanes, then warps
    double cost_w = (double)cost_l;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cost_w += __shfl_xor_sync(FULL, cost_w, o);
    s_dU[warp][lane] = dU_l;
    if (lane == 0) { s_cost[warp] = cost_w; s_nact[warp] = nact_w; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0.0; int n = 0;
        for (int w = 0; w < 8; ++w) { c += s_cost[w]; n += s_nact[w]; }
        a.tp_cost[tile] = c; a.tp_nact[tile] = n;
    }
    if (threadIdx.x < KS) {
        float vv = 0.f;
        for (int w = 0; w < 8; ++w)
            for (int l = threadIdx.x; l < V; l += KS) vv += s_dU[w][l];   // lanes that own this slice
        a.tp_dU[(size_t)tile * KS + threadIdx.x] = vv;
    }
}

// ==========================================================================================
// K4' score / hinge / backward-coefficient kernel (default; k_score above serves the large-batch GV variant)
//     NTN.java:200-222 (activation, u-dot, hinge, cost, update), :272-289 (dU, temp, db), column halves of :301-388.
//     Eight lanes share a unique triple: lane l takes the float4 chunks p = l, l+8, ... of every row, so a warp-wide
//     load covers whole 128-byte lines (4 triples x 128 B) -- a lane-per-(triple, slice) variant measured 68 us because
//     its 16-byte gathers cost one L1 wavefront each -- and every lane accumulates ALL dot products of its chunks
//     (slices x {10 corrupt columns + the positive entity}), so the d-reduction is a 3-level butterfly over 8 lanes
//     instead of a 32-lane transpose, and nothing is predicated per column.  Slot j of a group is then owned by lane
//     j % 8: activation, score, hinge and coefficients are evaluated once per column, not once per lane, and the
//     coefficients are broadcast back for pass 2, which rebuilds the M row  M_u[s,:] = a_{u,s} E_O + sum_c c_{c,s} E_c
//     from the same entity rows (L1/L2 hits; columns with zero coefficients -- hinge inactive, NTN.java:229-239 -- are
//     redirected to the positive row and cost no traffic).  k > 3: the slices are split over SG lane groups.
//     Coefficients go out as one packed record per incident for the entity pull:
//       crec[u]      = {a_{u,0..KS-1}, anchor role (0: e1, 1: e2), pad}        unique triple u
//       crec[Nu + c] = {c_{c,0..KS-1}, -1, pad}                                corrupt column c
// ==========================================================================================
#define SC2_NC 10                  // corrupt columns per group (the reference's corrupt_size, Run_NTN.java:77)
#define SC2_WARPS 2

struct Score2Args {
    const float *theta, *E, *T;
    float *M, *M_lo, *crec;
    int Nu, RS;                    // RS = floats per coefficient record (KS + 1 rounded up to 4)
    const int *t_rel, *t_u0, *t_n, *u_e1, *u_e2, *c_off, *c_e3;
    const uint8_t* side;
    double* tp_cost; int* tp_nact; float* tp_dU;
    int d, dq, Kp, Np;
    int64_t oU;
};

template <int KS> struct Score2Shape {
    static constexpr int SG = KS <= 3 ? 1 : (KS <= 6 ? 2 : 4);     // lane groups over the slices
    static constexpr int SPL = (KS + SG - 1) / SG;                 // slices per lane
    static constexpr int L = 8;                                    // lanes over the float4 chunks of a row (16 measured equal: +40 % instructions)
    static constexpr int G = 32 / (L * SG);                        // unique triples per warp: 4, 2, 1
};

// 16-byte asynchronous copy global -> shared (L2 only: a row is used once per triple), and its completion
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <int KS, int ACT, int DQT>                  // DQT: the row width dq when it is known at compile time (the reference's
__global__ void __launch_bounds__(SC2_WARPS * 32) k_score2(Score2Args a) {   // d = 100 and the scaled d = 256), else 0
    constexpr int L = Score2Shape<KS>::L, SG = Score2Shape<KS>::SG, SPL = Score2Shape<KS>::SPL, G = Score2Shape<KS>::G;
    constexpr int LT = L * SG;                       // lanes of one unique triple
    constexpr int NC = SC2_NC, NS = NC + 1;          // slots of a group: NC corrupt columns + the positive entity (slot NC)
    constexpr int NR = (NS + L - 1) / L;             // owner rounds: slot j belongs to lane j % L in round j / L
    constexpr int PR = NC / L, PO = NC % L;          // the positive slot NC is owned by lane PO in round PR
    constexpr unsigned LMASK = (L == 32) ? 0xffffffffu : ((1u << L) - 1u);
    // Row staging: the KS T' slices and the NS entity rows of every triple of the warp are copied to shared memory with
    // cp.async -- all of a triple's ~5.6 KB in flight at once without holding a register, where the register-staged
    // version of this kernel serialised its loads under the 128-register cap (72 % long-scoreboard stalls) -- and both
    // passes read them from there, so the entity rows cross the L2 once.
    extern __shared__ __align__(16) float sc2_smem[];
    __shared__ double s_cost[SC2_WARPS];
    __shared__ int s_nact[SC2_WARPS];
    __shared__ float s_dU[SC2_WARPS][32][SPL];
    const int tile = blockIdx.x;
    const int r = a.t_rel[tile], u0 = a.t_u0[tile], rows = a.t_n[tile];
    const bool tail = a.side[r] != 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / LT, sg = (lane / L) % SG, l = lane % L, lt = lane % LT;
    const int lbase = lane - l;                      // first lane of this (triple, slice group)
    const bool valid = warp * G + g < rows;
    const int u = u0 + (valid ? warp * G + g : 0);   // idle lane groups shadow the tile's first triple and never store
    const int dq = DQT ? DQT : a.dq, nch = dq >> 2, RS = a.RS;   // compile-time dq: every shared-memory offset is an immediate
    const int nit = (nch + L - 1) / L;
    const float* __restrict__ E = a.E;
    const float* Tr = a.T + (size_t)u * a.Np;
    float* S = sc2_smem + (size_t)(warp * G + g) * (KS + NS) * dq;       // this triple's rows: KS slices, then NS slots
    bool sv[SPL]; float Us[SPL], beta[SPL]; int so[SPL];
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
        const int s = sg * SPL + i;
        sv[i] = s < KS;
        const int sc = sv[i] ? s : 0;
        so[i] = sc * dq;
        Us[i] = sv[i] ? __ldg(a.theta + a.oU + (int64_t)r * KS + sc) : 0.f;
        beta[i] = sv[i] ? __ldg(Tr + KS * dq + sc) : 0.f;
    }
    const int c0 = __ldg(a.c_off + u), c1 = __ldg(a.c_off + u + 1);
    const int eo_other = __ldg((tail ? a.u_e2 : a.u_e1) + u) * a.Kp;      // element offset of the positive row
    const int ncol = valid ? c1 - c0 : 0;
    const int maxcol = __reduce_max_sync(FULL, ncol);                      // warp-uniform trip count of the group loop
    // the T' slices: staged once (KS * dq floats are contiguous in a T' row)
#pragma unroll (DQT ? 4 : 1)
    for (int p = lt; p < KS * nch; p += LT) cp_async16(S + 4 * p, Tr + 4 * p);

    float zp[SPL], cost_l = 0.f, dU_l[SPL], csum[SPL], sp = 0.f;
    int cnt = 0, nact_l = 0;
    float cc[SPL][NC];                               // group 0 stays in registers (and its rows in shared memory) for pass 2
#pragma unroll
    for (int i = 0; i < SPL; ++i) { zp[i] = 0.f; dU_l[i] = 0.f; csum[i] = 0.f; }

    // ---- pass 1: pre-activations, scores, hinge, coefficients; group gb covers columns [c0+gb, c0+gb+NC)
    for (int gb = max(maxcol, 1) - 1 - (max(maxcol, 1) - 1) % NC; gb >= 0; gb -= NC) {   // last group first: group 0's rows stay staged
        if (gb + NC < max(maxcol, 1)) __syncwarp();                      // every lane is done reading the previous group
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            int eg = eo_other;
            if (j < NC && gb + j < ncol) eg = __ldg(a.c_e3 + c0 + gb + j) * a.Kp;
            float* dst = S + (KS + j) * dq;
#pragma unroll (DQT ? 8 : 1)
            for (int p = lt; p < nch; p += LT) cp_async16(dst + 4 * p, E + eg + 4 * p);
        }
        cp_async_wait_all();
        __syncwarp();
        float acc[SPL][NS];
#pragma unroll
        for (int i = 0; i < SPL; ++i)
#pragma unroll
            for (int j = 0; j < NS; ++j) acc[i][j] = 0.f;
#pragma unroll (DQT ? 8 : 1)
        for (int it = 0; it < nit; ++it) {
            const int p = l + L * it;
            if (p < nch) {
                float4 t[SPL];
#pragma unroll
                for (int i = 0; i < SPL; ++i) t[i] = *reinterpret_cast<const float4*>(S + so[i] + 4 * p);
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    const float4 v = *reinterpret_cast<const float4*>(S + (KS + j) * dq + 4 * p);
#pragma unroll
                    for (int i = 0; i < SPL; ++i) acc[i][j] = dot4_acc(t[i], v, acc[i][j]);
                }
            }
        }
        // the T' columns beyond d are zero (Waug has no such rows), so the 1.0 at E[n][d] never enters a dot product.
        // butterfly over the L lanes of the triple: every lane ends with the same totals (x + y == y + x)
#pragma unroll
        for (int o = 1; o < L; o <<= 1)
#pragma unroll
            for (int i = 0; i < SPL; ++i)
#pragma unroll
                for (int j = 0; j < NS; ++j) acc[i][j] += __shfl_xor_sync(FULL, acc[i][j], o);
        // owner lanes: slot j = rd*L + l
        float z[NR][SPL], tot[NR];
#pragma unroll
        for (int rd = 0; rd < NR; ++rd) {
            float ps = 0.f;
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                float pre = 0.f;
#pragma unroll
                for (int jj = 0; jj < L; ++jj) if (rd * L + jj < NS) pre = (l == jj) ? acc[i][rd * L + jj] : pre;
                z[rd][i] = act_fn<ACT>(pre + beta[i]);                   // :200-201
                ps = fmaf(Us[i], z[rd][i], ps);                          // :204-205 (Us = 0 for a padding slice)
            }
#pragma unroll
            for (int o = L; o < L * SG; o <<= 1) ps += __shfl_xor_sync(FULL, ps, o);   // the other slice groups
            tot[rd] = ps;
        }
        sp = __shfl_sync(FULL, tot[PR], lbase + PO);                     // (restaged with every group: same value)
#pragma unroll
        for (int i = 0; i < SPL; ++i) zp[i] = __shfl_sync(FULL, z[PR][i], lbase + PO);
        float cmine[NR][SPL];
#pragma unroll
        for (int rd = 0; rd < NR; ++rd) {
            const int j = rd * L + l;
            const bool ok = j < NC && gb + j < ncol;
            const bool act = ok && ((sp + 1.0f) > tot[rd]);              // :213 (strict, FP32)
            const unsigned amask = __ballot_sync(FULL, act);
            if (act && sg == 0) { cost_l += (sp - tot[rd]) + 1.0f; nact_l++; }    // :221-222
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                cmine[rd][i] = act ? -act_diff<ACT>(z[rd][i]) * Us[i] : 0.f;       // :281
                if (act) dU_l[i] += zp[i] - z[rd][i];                    // :273
            }
            if (ok) {
                float* cr = a.crec + (size_t)(a.Nu + c0 + gb + j) * RS;
                if (SG == 1 && SPL == 3) {
                    *reinterpret_cast<float4*>(cr) = make_float4(cmine[rd][0], cmine[rd][1], cmine[rd][2], -1.f);
                } else {
#pragma unroll
                    for (int i = 0; i < SPL; ++i) if (sv[i]) cr[sg * SPL + i] = cmine[rd][i];
                    if (sg == 0) cr[KS] = -1.f;
                }
            }
            cnt += __popc((amask >> lbase) & LMASK);
        }
        // coefficients of every slot back to all L lanes (pass 2 needs them for its own chunks)
#pragma unroll
        for (int j = 0; j < NC; ++j)
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                const float c = __shfl_sync(FULL, cmine[j / L][i], lbase + (j % L));
                csum[i] += c;
                cc[i][j] = c;                                            // the last iteration is group 0
            }
    }
    float as[SPL];
#pragma unroll
    for (int i = 0; i < SPL; ++i) { as[i] = (float)cnt * (act_diff<ACT>(zp[i]) * Us[i]); csum[i] += as[i]; }   // :280

    // ---- pass 2: the M row, from the rows of group 0 still in shared memory
    if (maxcol > NC) __syncwarp();                   // the rare path below re-reads records stored by other lanes of this warp
    const bool masked = dq != a.d;
    float4 msk = make_float4(1.f, 1.f, 1.f, 1.f);
    if (masked) { const int p = dq - 4; msk = make_float4(p < a.d ? 1.f : 0.f, p + 1 < a.d ? 1.f : 0.f, p + 2 < a.d ? 1.f : 0.f, 0.f); }
    const size_t Mr = (size_t)u * a.Np;
#pragma unroll (DQT ? 8 : 1)
    for (int it = 0; it < nit; ++it) {
        const int p = l + L * it;
        if (p >= nch) continue;
        const int p4 = 4 * p;
        const float4 vo = *reinterpret_cast<const float4*>(S + (KS + NC) * dq + p4);
        float4 m[SPL];
#pragma unroll
        for (int i = 0; i < SPL; ++i) m[i] = make_float4(as[i] * vo.x, as[i] * vo.y, as[i] * vo.z, as[i] * vo.w);
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const float4 v = *reinterpret_cast<const float4*>(S + (KS + j) * dq + p4);
#pragma unroll
            for (int i = 0; i < SPL; ++i) fma4(m[i], cc[i][j], v);
        }
        if (maxcol > NC) {                                               // rare: duplicate training triples merged into one unique triple
            for (int c = c0 + NC; c < c0 + ncol; ++c) {
                const float4 ve = ldg4(E + (size_t)__ldg(a.c_e3 + c) * a.Kp + p4);
#pragma unroll
                for (int i = 0; i < SPL; ++i) {
                    const float cx = sv[i] ? a.crec[(size_t)(a.Nu + c) * RS + sg * SPL + i] : 0.f;   // written in pass 1 by this warp
                    fma4(m[i], cx, ve);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
            if (masked && p == nch - 1) { m[i].x *= msk.x; m[i].y *= msk.y; m[i].z *= msk.z; m[i].w *= msk.w; }
            if (valid && sv[i]) store_m4(a.M, a.M_lo, Mr + so[i] + p4, m[i]);
        }
    }
    if (valid && l == 0) {
#pragma unroll
        for (int i = 0; i < SPL; ++i)
            if (sv[i]) {
                store_m1(a.M, a.M_lo, Mr + KS * dq + sg * SPL + i, csum[i]);   // coefficient sums ride after the slices
                a.crec[(size_t)u * RS + sg * SPL + i] = as[i];
            }
        if (sg == 0) {
            a.crec[(size_t)u * RS + KS] = tail ? 0.f : 1.f;              // anchor role: e1 when the tail is corrupted
            for (int n = KS * dq + KS; n < a.Np; ++n) store_m1(a.M, a.M_lo, Mr + n, 0.f);
        }
    }
    // ---- tile partials, fixed order: lanes, then warps
    double cost_w = valid ? (double)cost_l : 0.0;
    int nact_w = valid ? nact_l : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { cost_w += __shfl_xor_sync(FULL, cost_w, o); nact_w += __shfl_xor_sync(FULL, nact_w, o); }
#pragma unroll
    for (int i = 0; i < SPL; ++i) s_dU[warp][lane][i] = (valid && sv[i]) ? dU_l[i] : 0.f;
    if (lane == 0) { s_cost[warp] = cost_w; s_nact[warp] = nact_w; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0.0; int n = 0;
        for (int w = 0; w < SC2_WARPS; ++w) { c += s_cost[w]; n += s_nact[w]; }
        a.tp_cost[tile] = c; a.tp_nact[tile] = n;
    }
    if (threadIdx.x < KS) {
        const int s = threadIdx.x, ssg = s / SPL, si = s - ssg * SPL;
        float vv = 0.f;
        for (int w = 0; w < SC2_WARPS; ++w)
            for (int ln = 0; ln < 32; ++ln)
                if ((ln / L) % SG == ssg) vv += s_dU[w][ln][si];
        a.tp_dU[(size_t)tile * KS + s] = vv;
    }
}

// ==========================================================================================
// K7  entity-gradient pull (replaces the 8*k*R dense scatter products of Util.java:74-98 and the
//     dense adds of NTN.java:372,388).  One thread per (entity, float4 chunk) walks the entity's
//     incident list in a fixed order -- no float atomics, bit-reproducible.  An incident is a packed
//     16-byte record {unique triple, contribution row, role, relation} and costs ONE row read: the
//     anchor role reads its GA row (entity-gradient contraction), every other role the row k_score
//     materialised in GV.  Hub entities (lists longer than NTN_HUB_SEG) are pre-reduced per segment
//     by k_entity_pull_hub.
// ==========================================================================================
struct PullArgs {
    const float *T, *coef;               // k-row variant: T' rows + coefficient rows [acoef (Nu x k) | ccoef (N x k)]
    const float *GV, *GA;                // GV: (Nu + N) x dq contribution rows written by k_score; GA: anchor rows
    const int4* inc;                     // {u, coefficient / GV row, role (0: e1, 1: e2, 2: e3), rel}
    const int* inc_off;
    const uint8_t* side;
    float* gE;
    const int *eh_first, *eh_count, *eh_begin, *eh_end;
    float* eh_part;
    int Ne, dq, Kp, Np, n_segs;
    FastDiv fd;                          // threads per row = dq / 4
};

// One batch of N incidents, all valid: records first, then every row / coefficient load of the batch (latency overlaps),
// then the FMAs in incident order (deterministic).
template <int KS, bool GVR, int N>
__device__ __forceinline__ void pull_batch(const PullArgs& a, int jb, int c, float4& acc) {
    if (GVR) {
        float4 rows[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int4 rec = __ldg(a.inc + jb + i);
            // a corrupt incident (role 2, >90 % of them) is never the anchor: the side flag is only looked up for
            // the e1 / e2 roles, so the common chain is record -> rows, not record -> flag -> rows
            const bool anchor = rec.z != 2 && ((a.side[rec.w] != 0) == (rec.z == 0));
            rows[i] = ldg4(anchor ? a.GA + (size_t)rec.x * a.Kp + c * 4 : a.GV + (size_t)rec.y * a.dq + c * 4);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) { acc.x += rows[i].x; acc.y += rows[i].y; acc.z += rows[i].z; acc.w += rows[i].w; }
    } else {
        float co[N][KS]; float4 rows[N][KS];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int4 rec = __ldg(a.inc + jb + i);
            const bool anchor = rec.z != 2 && ((a.side[rec.w] != 0) == (rec.z == 0));
            // coefficient and row loads are issued together: gating the rows on "any coefficient non-zero"
            // (inactive column, NTN.java:229-239) would add one more dependent round trip to every incident
            const float* Trow = a.T + (size_t)rec.x * a.Np + c * 4;
            const float* crow = a.coef + (size_t)rec.y * KS;
#pragma unroll
            for (int s = 0; s < KS; ++s) co[i][s] = anchor ? 0.f : __ldg(crow + s);
            rows[i][0] = ldg4(anchor ? a.GA + (size_t)rec.x * a.Kp + c * 4 : Trow);
#pragma unroll
            for (int s = 1; s < KS; ++s) rows[i][s] = anchor ? zero4() : ldg4(Trow + s * a.dq);
            if (anchor) co[i][0] = 1.f;
        }
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int s = 0; s < KS; ++s) fma4(acc, co[i][s], rows[i][s]);
    }
}

// GVR = true : every non-anchor incident reads ONE materialised row (large batches: T' does not fit in L2)
// GVR = false: it reads its coefficients and the k T' rows (small batches: T' is L2-resident, k_score stays lean)
// Full batches run without per-incident predicates (the kernel is issue-bound: 62 % issue-active, a fifth of it address
// arithmetic and zero-filling of the predicated tail); the remainder goes one incident at a time.
template <int KS, bool GVR>
__device__ __forceinline__ float4 pull_range(const PullArgs& a, int j0, int j1, int c) {
    constexpr int NB = GVR ? 4 : ((KS <= 4) ? 2 : 1);               // incidents in flight (ILP)
    float4 acc = zero4();
    int jb = j0;
    for (; jb + NB <= j1; jb += NB) pull_batch<KS, GVR, NB>(a, jb, c, acc);
    for (; jb < j1; ++jb) pull_batch<KS, GVR, 1>(a, jb, c, acc);
    return acc;
}

template <int KS, bool GVR>
__global__ void __launch_bounds__(256) k_entity_pull_hub(PullArgs a) {
    const unsigned gid = flat_index();
    const int seg = (int)fd_div(gid, a.fd), c = (int)(gid - (unsigned)seg * a.fd.d);
    if (seg >= a.n_segs) return;
    float4 acc = pull_range<KS, GVR>(a, a.eh_begin[seg], a.eh_end[seg], c);
    *reinterpret_cast<float4*>(a.eh_part + (size_t)seg * a.dq + c * 4) = acc;
}

template <int KS, bool GVR, bool PUSH>
__global__ void __launch_bounds__(256, (KS <= 4) ? 6 : 2) k_entity_pull(PullArgs a, XPushRange xp, int row0, int nrows) {
    const unsigned gid = flat_index();
    const int nl = (int)fd_div(gid, a.fd), c = (int)(gid - (unsigned)nl * a.fd.d);
    if (nl >= nrows) return;
    const int n = row0 + nl;                                         // entities [row0, row0 + nrows): all, or one exchange chunk
    float4 acc;
    int nseg = a.eh_count ? a.eh_count[n] : 0;
    if (nseg == 0) {
        acc = pull_range<KS, GVR>(a, __ldg(a.inc_off + n), __ldg(a.inc_off + n + 1), c);
    } else {
        acc = zero4();
        int s0 = a.eh_first[n];
        for (int sg = 0; sg < nseg; sg += 4) {
            float4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                v[i] = (sg + i < nseg) ? ldg4(a.eh_part + (size_t)(s0 + sg + i) * a.dq + c * 4) : zero4();
#pragma unroll
            for (int i = 0; i < 4; ++i) { acc.x += v[i].x; acc.y += v[i].y; acc.z += v[i].z; acc.w += v[i].w; }
        }
    }
    ge_store<PUSH>(a.gE, xp, (size_t)n * a.dq + c * 4, acc);
}

// ==========================================================================================
// K7' entity-gradient pull over packed coefficient records (the default; the kernels above serve the GV variant).
//     Same sorted-segment pull (Util.java:74-98, NTN.java:372,388; no float atomics), trimmed for issue slots:
//       * one 16-byte coefficient record per incident (k_score2) instead of k scalar loads, and the record of a unique
//         triple names its anchor role, so no corrupt-side lookup;
//       * entities are visited in descending order of their incident count (ent_order, built per batch job): the
//         threads of a warp -- which span two entities -- run the same number of iterations;
//       * the rows of an incident are requested right after its record: a positive-role incident requests its GA row
//         AND its k T' rows and picks when the coefficient record arrives (17 % of the incidents, +5 % row traffic),
//         so no incident waits for record -> coefficients -> rows.
// ==========================================================================================
struct Pull2Args {
    const float *T, *crec, *GA;
    const int4* inc;                     // {u, coefficient record, role (0: e1, 1: e2, 2: e3), rel}
    const int *inc_off, *order;
    float* gE;
    const int *eh_first, *eh_count, *eh_begin, *eh_end;
    float* eh_part;
    int Ne, dq, Kp, Np, RS, n_segs;
    FastDiv fd;                          // threads per row = dq / 4
};

template <int KS, int N>
__device__ __forceinline__ void pull2_batch(const Pull2Args& a, int jb, int c4, float4& acc) {
    constexpr int RQ = (KS + 4) / 4;                                  // float4s per record: KS coefficients + the role
    float4 co[N][RQ], rows[N][KS], ga[N];
    int role[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const int4 rec = __ldg(a.inc + jb + i);
        role[i] = rec.z;
        const float* cr = a.crec + (size_t)rec.y * a.RS;
#pragma unroll
        for (int q = 0; q < RQ; ++q) co[i][q] = ldg4(cr + 4 * q);
        const float* Trow = a.T + (size_t)rec.x * a.Np + c4;
#pragma unroll
        for (int s = 0; s < KS; ++s) rows[i][s] = ldg4(Trow + s * a.dq);
        ga[i] = (rec.z != 2) ? ldg4(a.GA + (size_t)rec.x * a.Kp + c4) : zero4();
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const float* cf = reinterpret_cast<const float*>(&co[i][0]);
        const bool anchor = role[i] != 2 && cf[KS] == (float)role[i];
        const float wg = anchor ? 1.f : 0.f;
#pragma unroll
        for (int s = 0; s < KS; ++s) fma4(acc, anchor ? 0.f : cf[s], rows[i][s]);
        fma4(acc, wg, ga[i]);
    }
}

template <int KS>
__device__ __forceinline__ float4 pull2_range(const Pull2Args& a, int j0, int j1, int c4) {
    constexpr int NB = (KS <= 4) ? 2 : 1;                             // incidents in flight
    float4 acc = zero4();
    int jb = j0;
    for (; jb + NB <= j1; jb += NB) pull2_batch<KS, NB>(a, jb, c4, acc);
    for (; jb < j1; ++jb) pull2_batch<KS, 1>(a, jb, c4, acc);
    return acc;
}

template <int KS>
__global__ void __launch_bounds__(256) k_entity_pull2_hub(Pull2Args a) {
    const unsigned gid = flat_index();
    const int seg = (int)fd_div(gid, a.fd), c = (int)(gid - (unsigned)seg * a.fd.d);
    if (seg >= a.n_segs) return;
    float4 acc = pull2_range<KS>(a, a.eh_begin[seg], a.eh_end[seg], c * 4);
    *reinterpret_cast<float4*>(a.eh_part + (size_t)seg * a.dq + c * 4) = acc;
}

template <int KS, bool PUSH>
__global__ void __launch_bounds__(256, (KS <= 4) ? 4 : 2) k_entity_pull2(Pull2Args a, XPushRange xp, int row0, int nrows) {
    const unsigned gid = flat_index();
    const int slot = (int)fd_div(gid, a.fd), c = (int)(gid - (unsigned)slot * a.fd.d);
    if (slot >= nrows) return;
    // single GPU: slots in descending order of degree (a.order); exchange chunks: entities [row0, row0 + nrows) as they come
    const int n = a.order ? __ldg(a.order + slot) : row0 + slot;
    float4 acc;
    const int nseg = a.eh_count ? a.eh_count[n] : 0;
    if (nseg == 0) {
        acc = pull2_range<KS>(a, __ldg(a.inc_off + n), __ldg(a.inc_off + n + 1), c * 4);
    } else {
        acc = zero4();
        const int s0 = a.eh_first[n];
        for (int sg = 0; sg < nseg; sg += 4) {
            float4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                v[i] = (sg + i < nseg) ? ldg4(a.eh_part + (size_t)(s0 + sg + i) * a.dq + c * 4) : zero4();
#pragma unroll
            for (int i = 0; i < 4; ++i) { acc.x += v[i].x; acc.y += v[i].y; acc.z += v[i].z; acc.w += v[i].w; }
        }
    }
    ge_store<PUSH>(a.gE, xp, (size_t)n * a.dq + c * 4, acc);
}

// ==========================================================================================
// K7'' entity-gradient pull, one WARP per range of entities (the default).  Same sorted-segment pull (Util.java:74-98,
//     NTN.java:372,388; fixed summation order, no float atomics), re-cut for issue slots and latency:
//       * work is dealt by INCIDENTS, not by entities: range r holds the consecutive entities whose key
//         inc_off[n] + n falls in [32 r, 32 r + 32) -- at most 32 entities and (hubs aside) at most 63 incidents, so
//         every warp has the same amount of work whatever the degree distribution, and empty entities (which only need
//         their zero row) cost one unit each;
//       * the lanes of the warp first fetch the range's incident records and coefficient records in ONE coalesced
//         load each (lane l takes incident l) -- the record -> coefficient chain is paid once per 32 incidents, where a
//         thread per (entity, chunk) paid it per incident and 25 threads decoded the same record -- and then walk the
//         incidents with the lanes over the float4 chunks of a row, the decoded fields broadcast by shuffle, the rows of
//         NB incidents requested before the first is consumed;
//       * an anchor incident reads its GA row only, any other its k T' rows only (the role is known before the rows are
//         requested).
//     A hub entity (more than NTN_HUB_SEG incidents, pre-reduced per segment by k_entity_pull2_hub) is always the last
//     entity of its range: the next key is more than 32 away.
// ==========================================================================================
struct Pull3Args {
    const float *T, *crec, *GA;
    const int4* inc;
    const int* inc_off;
    const int4* rng;                     // {first entity, end entity, first incident, end incident (a trailing hub's excluded)}
    float* gE;
    const int *eh_first, *eh_count;
    const float* eh_part;
    int nranges, dq, Kp, Np, RS;
};

__global__ void __launch_bounds__(256) k_pull_ranges(const int* __restrict__ inc_off, const int* __restrict__ eh_count, int Ne,
                                                     int4* __restrict__ rng) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= Ne) return;
    const int o = __ldg(inc_off + n), o1 = __ldg(inc_off + n + 1);
    const int r = (o + n) >> 5;
    const int rp = n ? ((__ldg(inc_off + n - 1) + n - 1) >> 5) : -1;
    const int rn = (n + 1 < Ne) ? ((o1 + n + 1) >> 5) : -2;
    if (r != rp) { rng[r].x = n; rng[r].z = o; }
    if (r != rn) { rng[r].y = n + 1; rng[r].w = (eh_count && eh_count[n]) ? o : o1; }
}

template <int KS, int CPL>
__global__ void __launch_bounds__(128) k_entity_pull3(Pull3Args a) {
    constexpr int RQ = (KS + 4) / 4;                                  // float4s per coefficient record (KS coefficients + role)
    constexpr int NB = (12 / (KS * CPL)) < 1 ? 1 : (12 / (KS * CPL)); // incidents whose rows are in flight together
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= a.nranges) return;
    const int4 rg = __ldg(a.rng + warp);
    if (rg.x < 0) return;                                             // no entity starts in this key interval
    const int n_ent = rg.y - rg.x;
    const bool last_hub = a.eh_count != nullptr && __ldg(a.eh_count + rg.y - 1) != 0;
    const int n_flat = n_ent - (last_hub ? 1 : 0);
    const int my_end = (lane < n_flat) ? __ldg(a.inc_off + rg.x + lane + 1) : 0x7fffffff;
    int c4[CPL]; bool cv[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) { c4[c] = (lane + 32 * c) * 4; cv[c] = c4[c] < a.dq; }
    float4 acc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[c] = zero4();
    int e = 0, e_end = __shfl_sync(FULL, my_end, 0);
    // entity e is complete once the next incident to consume is e_end: store its row, move on (empty entities: zeros)
    auto flush_until = [&](int jj) {
        while (e < n_flat && jj == e_end) {
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                if (cv[c]) *reinterpret_cast<float4*>(a.gE + (size_t)(rg.x + e) * a.dq + c4[c]) = acc[c];
                acc[c] = zero4();
            }
            ++e;
            e_end = __shfl_sync(FULL, my_end, e & 31);
        }
    };
    for (int jb = rg.z; jb < rg.w; jb += 32) {
        const int bend = min(jb + 32, rg.w);
        int4 rec = make_int4(0, 0, 2, 0);
        float cr[RQ * 4];
#pragma unroll
        for (int q = 0; q < RQ * 4; ++q) cr[q] = 0.f;
        if (jb + lane < bend) {
            rec = __ldg(a.inc + jb + lane);
            const float* cp = a.crec + (size_t)rec.y * a.RS;
#pragma unroll
            for (int q = 0; q < RQ; ++q) {
                const float4 t = ldg4(cp + 4 * q);
                cr[4 * q] = t.x; cr[4 * q + 1] = t.y; cr[4 * q + 2] = t.z; cr[4 * q + 3] = t.w;
            }
        }
        for (int j = jb; j < bend; j += NB) {
            float4 rows[NB][KS][CPL];
            float cf[NB][KS];
            bool anchor[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const int src = (j + i - jb) & 31;
                const int u = __shfl_sync(FULL, rec.x, src), role = __shfl_sync(FULL, rec.z, src);
                const float crole = __shfl_sync(FULL, cr[KS], src);
#pragma unroll
                for (int s = 0; s < KS; ++s) cf[i][s] = __shfl_sync(FULL, cr[s], src);
                anchor[i] = role != 2 && crole == (float)role;
                if (j + i < bend) {
                    if (anchor[i]) {
#pragma unroll
                        for (int c = 0; c < CPL; ++c) if (cv[c]) rows[i][0][c] = ldg4(a.GA + (size_t)u * a.Kp + c4[c]);
                    } else {
                        const float* Trow = a.T + (size_t)u * a.Np;
#pragma unroll
                        for (int s = 0; s < KS; ++s)
#pragma unroll
                            for (int c = 0; c < CPL; ++c) if (cv[c]) rows[i][s][c] = ldg4(Trow + s * a.dq + c4[c]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                if (j + i < bend) {
                    flush_until(j + i);
                    if (anchor[i]) {
#pragma unroll
                        for (int c = 0; c < CPL; ++c) if (cv[c]) { acc[c].x += rows[i][0][c].x; acc[c].y += rows[i][0][c].y; acc[c].z += rows[i][0][c].z; acc[c].w += rows[i][0][c].w; }
                    } else {
#pragma unroll
                        for (int s = 0; s < KS; ++s)
#pragma unroll
                            for (int c = 0; c < CPL; ++c) if (cv[c]) fma4(acc[c], cf[i][s], rows[i][s][c]);
                    }
                }
            }
        }
    }
    flush_until(rg.w);
    // (rg.w is the end offset of the last flat entity: the loop above has stored it; anything left is an empty tail)
    while (e < n_flat) {
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            if (cv[c]) *reinterpret_cast<float4*>(a.gE + (size_t)(rg.x + e) * a.dq + c4[c]) = acc[c];
            acc[c] = zero4();
        }
        ++e;
    }
    if (last_hub) {                                                   // sum of the pre-reduced segment rows, in segment order
        const int n = rg.y - 1, s0 = __ldg(a.eh_first + n), nseg = __ldg(a.eh_count + n);
        float4 hacc[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) hacc[c] = zero4();
        for (int sg = 0; sg < nseg; sg += 4) {
            float4 v[4][CPL];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < CPL; ++c)
                    v[i][c] = (sg + i < nseg && cv[c]) ? ldg4(a.eh_part + (size_t)(s0 + sg + i) * a.dq + c4[c]) : zero4();
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < CPL; ++c) { hacc[c].x += v[i][c].x; hacc[c].y += v[i][c].y; hacc[c].z += v[i][c].z; hacc[c].w += v[i][c].w; }
        }
#pragma unroll
        for (int c = 0; c < CPL; ++c) if (cv[c]) *reinterpret_cast<float4*>(a.gE + (size_t)n * a.dq + c4[c]) = hacc[c];
    }
}

// ==========================================================================================
// K8  word-gradient pull + regulariser for the word block          NTN.java:411-430,437-438
//     gWV[w] = (1/B) sum_{n : w in words(n)} gE[n]/len_n  + lambda*theta_w ;  reg += theta_w^2
//     Sorted-segment pull over the inverse (word -> entity) CSR: deterministic, no float atomics.
//     Lists are cut in NTN_HUB_SEG-entry segments; a segment is one warp, its row loads issued 8 at a
//     time (latency-bound otherwise).  Hub words ("unknown", frequent words) have their segments
//     pre-reduced by k_word_pull_hub and are finished by one block each in k_word_finish_hub.
// ==========================================================================================
struct WordArgs {
    const float *gE, *inv_len, *theta;   // inv_len[n] = 1 / entityLength(n), rounded once on the host (NTN.java:420)
    float* grad;
    const int *w_off, *w_ent;
    const int4* w_rec;                   // {j0, j1, first entity (or -1), hub segment count}: one load instead of a chain
    const int *wh_first, *wh_count, *wh_begin, *wh_end, *hub_words;
    float* wh_part;
    double* reg_part;      // [word blocks | hub words | small-block finalize]
    int Nw, d, dq, n_segs, n_hub_words, n_word_blocks;
    int64_t oWV;
    float inv_B, lam;
    FastDiv fd;                          // threads per row = dq / 4
};

// one batch of N entities of a word's list, all valid: loads first, then the adds in entity order (NTN.java:414-426)
template <int N>
__device__ __forceinline__ void word_batch(const WordArgs& a, int jb, int c, float4& acc) {
    float4 v[N]; float inv[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const int n = __ldg(a.w_ent + jb + i);
        inv[i] = __ldg(a.inv_len + n);                               // :420
        v[i] = ldg4(a.gE + (size_t)n * a.dq + c * 4);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        acc.x = fmaf(v[i].x, inv[i], acc.x); acc.y = fmaf(v[i].y, inv[i], acc.y);
        acc.z = fmaf(v[i].z, inv[i], acc.z); acc.w = fmaf(v[i].w, inv[i], acc.w);
    }
}
template <int NB>                                                    // entities in flight (ILP)
__device__ __forceinline__ float4 word_range(const WordArgs& a, int j0, int j1, int c) {
    float4 acc = zero4();
    int jb = j0;
    for (; jb + NB <= j1; jb += NB) word_batch<NB>(a, jb, c, acc);   // full batches: no per-entity predicates
    for (; jb < j1; ++jb) word_batch<1>(a, jb, c, acc);
    return acc;
}

// c4 = first column of this float4 chunk.  The internal word rows are dq = ceil4(d) wide: the padding columns are not
// parameters (the host layout has none), so their gradient is written as 0 and they stay out of the regulariser --
// gE's padding columns carry the bias product of the anchor rows (GA[u][d]) and must not leak into the optimiser.
template <bool PUSH>
__device__ __forceinline__ float word_store(const WordArgs& a, const XPushRange& xp, size_t o, float4 th, float4 acc, int c4) {
    if (c4 + 4 > a.d) {
        if (c4 + 1 > a.d) { th.x = 0.f; acc.x = 0.f; }
        if (c4 + 2 > a.d) { th.y = 0.f; acc.y = 0.f; }
        if (c4 + 3 > a.d) { th.z = 0.f; acc.z = 0.f; }
        th.w = 0.f; acc.w = 0.f;
    }
    float4 g = make_float4(fmaf(a.lam, th.x, acc.x * a.inv_B), fmaf(a.lam, th.y, acc.y * a.inv_B),
                           fmaf(a.lam, th.z, acc.z * a.inv_B), fmaf(a.lam, th.w, acc.w * a.inv_B));
    *reinterpret_cast<float4*>(a.grad + o) = g;                      // :429,438
    if (PUSH) {
        // data parallel, push exchange (exchange.cu): a group that another rank owns also goes straight into that rank's
        // scratch -- a posted store over NVLink, issued where the value is produced
        const unsigned rel4 = (unsigned)((o - (size_t)xp.off) >> 2);
        const unsigned owner = fd_div(rel4, FastDiv{xp.per4, xp.per_m, xp.per_s});
        if ((int)owner != xp.rank) {
            float* dst = xp.dst[0];                                  // (selects: a dynamic index would put the table on the stack)
#pragma unroll
            for (int s = 1; s < NTN_MAX_PEERS; ++s) dst = (owner == (unsigned)s) ? xp.dst[s] : dst;
            *reinterpret_cast<float4*>(dst + 4 * (size_t)(rel4 - owner * xp.per4)) = g;
        }
    }
    return dot4(th, th);                                             // :437
}

struct XPushAll {                        // every word chunk's push descriptor (the hub words are spread over all of them)
    XPushRange r[NTN_MAX_WORD_CHUNKS];
    int w_end[NTN_MAX_WORD_CHUNKS];      // chunk c holds words [w_end[c-1], w_end[c])
    int n;                               // 0: no push
};

// one thread per (hub segment, float4 chunk)
__global__ void __launch_bounds__(256) k_word_pull_hub(WordArgs a) {
    const unsigned gid = flat_index();
    const int seg = (int)fd_div(gid, a.fd), c = (int)(gid - (unsigned)seg * a.fd.d);
    if (seg >= a.n_segs) return;
    float4 acc = word_range<8>(a, a.wh_begin[seg], a.wh_end[seg], c);   // short latency-bound branch: 8 rows in flight
    *reinterpret_cast<float4*>(a.wh_part + (size_t)seg * a.dq + c * 4) = acc;
}

// one block (32 warps) per hub word: every warp sums an interleaved subset of the partial rows into
// 8 independent accumulators (8 loads in flight), then a fixed-order combine.  Lanes over float4 chunks.
template <int CH, bool PUSH>
__global__ void __launch_bounds__(1024) k_word_finish_hub(WordArgs a, XPushAll xa) {
    __shared__ float4 s_acc[32][CH][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = a.hub_words[blockIdx.x];
    const int s0 = a.wh_first[w], nseg = a.wh_count[w];
    bool cv[CH]; float4 acc[8][CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        cv[c] = (lane + 32 * c) * 4 < a.dq;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][c] = zero4();
    }
    for (int sb = warp; sb < nseg; sb += 256)
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < CH; ++c)
                if (sb + 32 * i < nseg && cv[c]) {
                    float4 v = ldg4(a.wh_part + (size_t)(s0 + sb + 32 * i) * a.dq + (lane + 32 * c) * 4);
                    acc[i][c].x += v.x; acc[i][c].y += v.y; acc[i][c].z += v.z; acc[i][c].w += v.w;
                }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        float4 t = acc[0][c];
#pragma unroll
        for (int i = 1; i < 8; ++i) { t.x += acc[i][c].x; t.y += acc[i][c].y; t.z += acc[i][c].z; t.w += acc[i][c].w; }
        s_acc[warp][c][lane] = t;
    }
    __syncthreads();
    if (warp == 0) {
        float reg = 0.f;
        XPushRange xp = xa.r[0];
        if (PUSH) {
#pragma unroll
            for (int i = 1; i < NTN_MAX_WORD_CHUNKS; ++i) if (i < xa.n && w >= xa.w_end[i - 1]) xp = xa.r[i];
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            float4 t = s_acc[0][c][lane];
            for (int ww = 1; ww < 32; ++ww) { float4 v = s_acc[ww][c][lane]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
            if (cv[c]) {
                size_t o = (size_t)a.oWV + (size_t)w * a.dq + (lane + 32 * c) * 4;
                reg += word_store<PUSH>(a, xp, o, ldg4(a.theta + o), t, (lane + 32 * c) * 4);
            }
        }
        reg = warp_sum(reg);
        if (lane == 0) a.reg_part[a.n_word_blocks + blockIdx.x] = (double)reg;
    }
}

// one thread per (word pair, float4 chunk): words w and w + half, two independent record -> entity-row chains in flight
// per thread (the kernel is a few waves of dependent DRAM round trips); hub words are finished by k_word_finish_hub
template <bool PUSH>
__device__ __forceinline__ float word_one(const WordArgs& a, const XPushRange& xp, int w, int c, const int4& rec, const float4& th, size_t o) {
    if (rec.w != 0) return 0.f;                                      // hub words are finished elsewhere
    float4 acc = zero4();
    if (rec.y - rec.x == 1) {                                        // the common case: exactly one entity
        const float inv = __ldg(a.inv_len + rec.z);
        float4 v = ldg4(a.gE + (size_t)rec.z * a.dq + c * 4);
        acc = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
    } else if (rec.y > rec.x) {
        acc = word_range<4>(a, rec.x, rec.y, c);
    }
    return word_store<PUSH>(a, xp, o, th, acc, c * 4);
}
template <bool PUSH>
__global__ void __launch_bounds__(256, 5) k_word_pull(WordArgs a, XPushRange xp, int w0, int wn, int half, int block_off) {
    // words [w0, w0 + wn): the whole block (single GPU) or one chunk of it (the exchange of a finished chunk overlaps the
    // pull of the next, exchange.cu); block_off = this chunk's first regulariser-partial slot
    __shared__ double s_reg[8];
    const unsigned gid = flat_index();
    const int wl = (int)fd_div(gid, a.fd), c = (int)(gid - (unsigned)wl * a.fd.d);
    float reg = 0.f;
    if (wl < half) {
        const int w = w0 + wl, w2 = w + half;
        const bool two = wl + half < wn;
        const size_t o1 = (size_t)a.oWV + (size_t)w * a.dq + c * 4, o2 = (size_t)a.oWV + (size_t)(two ? w2 : w) * a.dq + c * 4;
        const int4 r1 = __ldg(a.w_rec + w);
        const int4 r2 = __ldg(a.w_rec + (two ? w2 : w));
        const float4 t1 = ldg4(a.theta + o1), t2 = ldg4(a.theta + o2);
        reg = word_one<PUSH>(a, xp, w, c, r1, t1, o1);
        if (two) reg += word_one<PUSH>(a, xp, w2, c, r2, t2, o2);
    }
    reg = warp_sum(reg);
    if ((threadIdx.x & 31) == 0) s_reg[threadIdx.x >> 5] = (double)reg;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += s_reg[i];
        a.reg_part[block_off + blockIdx.x] = t;
    }
}

// ==========================================================================================
// K9  finalize the W, V, b, U blocks: fixed-order reduction of the dWaug split-K partials and of the
//     tile partials, /batchSize, + lambda*theta, regulariser partial    NTN.java:398-401,433-438
// ==========================================================================================
struct FinArgs {
    const float *theta, *dWpart, *tp_dU;
    float* grad;
    const int *relch_off, *rels_off;
    const uint8_t* side;
    double* reg_part;      // [gridDim.x] slots after the word-pull slots
    NtnDims dm;
    float inv_B, lam;
};

__global__ void __launch_bounds__(256) k_finalize_small(FinArgs a) {
    __shared__ double s_reg[8];
    const NtnDims& dm = a.dm;
    const int d = dm.d, k = dm.k, dq = dm.dq;
    const int64_t per = (int64_t)dm.Kp * dm.Np;
    const int64_t nW = per * dm.R, nU = (int64_t)dm.R * k;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float th2 = 0.f;
    if (i < nW) {
        int r = (int)(i / per);
        int rem = (int)(i - (int64_t)r * per);
        int n = rem / dm.Kp, p = rem - n * dm.Kp;                      // partials are [n][p], p contiguous
        bool tail = a.side[r] != 0;
        int64_t idx = -1;
        if (n < k * dq) {
            int s = n / dq, q = n - s * dq;
            if (q < d) {
                if (p < d) idx = dm.oW + (((int64_t)r * k + s) * d + (tail ? p : q)) * d + (tail ? q : p);
                else if (p == d) idx = dm.oV + ((int64_t)r * 2 * d + (tail ? d + q : q)) * k + s;
            }
        } else if (n < k * dq + k) {
            int s = n - k * dq;
            if (p < d) idx = dm.oV + ((int64_t)r * 2 * d + (tail ? p : d + p)) * k + s;
            else if (p == d) idx = dm.ob + (int64_t)r * k + s;
        }
        if (idx >= 0) {
            float sum = 0.f;
            for (int c = a.relch_off[r]; c < a.relch_off[r + 1]; ++c) sum += __ldg(a.dWpart + (size_t)c * per + rem);
            float th = a.theta[idx];
            a.grad[idx] = fmaf(a.lam, th, sum * a.inv_B);
            th2 = th * th;
        }
    } else if (i >= nW + nU && i < nW + nU + (dm.oWVint - (dm.oU + nU))) {
        a.grad[dm.oU + nU + (i - nW - nU)] = 0.f;                     // alignment padding before the word block
    }
    float s = warp_sum(th2);
    if ((threadIdx.x & 31) == 0) s_reg[threadIdx.x >> 5] = (double)s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_reg[w];
        a.reg_part[blockIdx.x] = t;
    }
}

// dU_r[s] = (1/B) sum over the relation's score tiles of the per-tile partial + lambda*theta   (NTN.java:273,401,438)
// One block per (relation, slice): threads take interleaved tiles, then a fixed-order tree (deterministic).
__global__ void __launch_bounds__(256) k_finalize_u(FinArgs a, double* reg_slot) {
    __shared__ float sm[256];
    const int k = a.dm.k, r = blockIdx.x / k, s = blockIdx.x - r * k;
    float v = 0.f;
    for (int t = a.rels_off[r] + threadIdx.x; t < a.rels_off[r + 1]; t += 256) v += a.tp_dU[(size_t)t * k + s];
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) {
        const int64_t idx = a.dm.oU + blockIdx.x;
        const float th = a.theta[idx];
        a.grad[idx] = fmaf(a.lam, th, sm[0] * a.inv_B);
        reg_slot[blockIdx.x] = (double)th * (double)th;
    }
}

// single block: cost = sum(tile costs)/B + 0.5*lambda*sum(theta^2)      NTN.java:430,437
__global__ void __launch_bounds__(1024) k_final_reduce(const double* __restrict__ tp_cost, const int* __restrict__ tp_nact,
                                                       int ntiles, const double* __restrict__ reg_part, int nreg,
                                                       double inv_B, double half_lam, double* __restrict__ out,
                                                       float* __restrict__ tail) {
    __shared__ double s_a[1024], s_b[1024], s_c[1024];
    double c = 0.0, n = 0.0, r = 0.0;
    for (int i = threadIdx.x; i < ntiles; i += 1024) { c += tp_cost[i]; n += (double)tp_nact[i]; }
    for (int i = threadIdx.x; i < nreg; i += 1024) r += reg_part[i];
    s_a[threadIdx.x] = c; s_b[threadIdx.x] = n; s_c[threadIdx.x] = r;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s_a[threadIdx.x] += s_a[threadIdx.x + o];
            s_b[threadIdx.x] += s_b[threadIdx.x + o];
            s_c[threadIdx.x] += s_c[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = s_a[0] * inv_B + half_lam * s_c[0];
        out[1] = s_b[0];
        out[2] = s_a[0] * inv_B;       // data term alone
        out[3] = s_c[0];               // sum theta^2
        if (tail) {                    // {cost, n_active} as float pairs after the gradient (exact per rank; the count stays exact under the FP32
                                       // sum over ranks, the cost is then good to ~1e-7 relative): one collective covers all
            const float chi = (float)out[0];
            tail[0] = chi; tail[1] = (float)(out[0] - (double)chi);
            const long long na = (long long)s_b[0];
            tail[2] = (float)(na >> 12); tail[3] = (float)(na & 4095);
        }
    }
}

// ==========================================================================================
// layout conversion between the reference flattening (NTN.java:450-526; word block d x Nw) and the
// internal one (word block Nw x dq).  32x32 shared-memory transposes, coalesced on both sides.
// ==========================================================================================
__global__ void __launch_bounds__(256) k_copy_small(const float* __restrict__ src, float* __restrict__ dst, int64_t n,
                                                    int64_t pad_to) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
    else if (i < pad_to) dst[i] = 0.f;
}
// ref [d][Nw] -> int [Nw][dq]
__global__ void __launch_bounds__(256) k_wv_to_internal(const float* __restrict__ src, float* __restrict__ dst, int d, int Nw, int dq) {
    __shared__ float tile[32][33];
    int w0 = blockIdx.x * 32, p0 = blockIdx.y * 32;
    int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        int p = p0 + j, w = w0 + tx;
        tile[j][tx] = (p < d && w < Nw) ? src[(size_t)p * Nw + w] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        int w = w0 + j, p = p0 + tx;
        if (w < Nw && p < dq) dst[(size_t)w * dq + p] = tile[tx][j];
    }
}
// int [Nw][dq] -> ref [d][Nw]
__global__ void __launch_bounds__(256) k_wv_from_internal(const float* __restrict__ src, float* __restrict__ dst, int d, int Nw, int dq) {
    __shared__ float tile[32][33];
    int w0 = blockIdx.x * 32, p0 = blockIdx.y * 32;
    int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        int w = w0 + j, p = p0 + tx;
        tile[j][tx] = (w < Nw && p < dq) ? src[(size_t)w * dq + p] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        int p = p0 + j, w = w0 + tx;
        if (p < d && w < Nw) dst[(size_t)p * Nw + w] = tile[tx][j];
    }
}

// ==========================================================================================
// eval-side scoring ("next" row N1)             NTN.calculateScoreOfaTripple, NTN.java:713-736
// one warp per triple; FP64 accumulation is not needed: the reference is FP32.
// ==========================================================================================
template <int ACT>
__global__ void __launch_bounds__(256) k_score_eval(const float* __restrict__ theta, const float* __restrict__ E,
                                                    const int* __restrict__ e1, const int* __restrict__ rel,
                                                    const int* __restrict__ e2, int64_t n, int mode,
                                                    double* __restrict__ out, NtnDims dm) {
    int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (j >= n) return;
    const int d = dm.d, k = dm.k;
    const int r = rel[j];
    const float* a = E + (size_t)e1[j] * dm.Kp;
    const float* c = E + (size_t)e2[j] * dm.Kp;
    float total = 0.f;
    for (int s = 0; s < k; ++s) {
        const float* W = theta + dm.oW + ((int64_t)r * k + s) * d * d;
        const float* V = theta + dm.oV + (int64_t)r * 2 * d * k;
        float bil = 0.f, lin = 0.f;
        for (int p = lane; p < d; p += 32) {
            float wc = 0.f;
            for (int q = 0; q < d; ++q) wc = fmaf(W[(size_t)p * d + q], c[q], wc);    // :727
            bil = fmaf(a[p], wc, bil);                                                // :728
            lin = fmaf(V[p * k + s], a[p], lin);                                      // :729
            lin = fmaf(V[(d + p) * k + s], c[p], lin);
        }
        bil = warp_sum(bil); lin = warp_sum(lin);
        float u = theta[dm.oU + (int64_t)r * k + s], b = theta[dm.ob + (int64_t)r * k + s];
        if (mode == NTN_SCORE_REFERENCE_EVAL) total += u * bil + lin + b;             // :730 (no activation)
        else total += u * act_fn<ACT>(bil + lin + b);
    }
    if (lane == 0) out[j] = (double)total;
}

// ==========================================================================================
// launchers
// ==========================================================================================
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

void ntn_launch_entity_build(ntn_handle* h, const float* wvt) {
    const NtnDims& dm = h->dm;
    const int half = (dm.Ne + 1) / 2;
    k_entity_build<<<cdiv((int64_t)half * (dm.Kp / 4), 256), 256, 0, h->stream>>>(wvt, h->w.fwd_rec, h->w.fwd_idx, h->E,
                                                                                 dm.Ne, half, dm.d, dm.dq, dm.Kp, make_fastdiv((unsigned)(dm.Kp / 4)));
    h->launches++;
}

void ntn_launch_w_prep(ntn_handle* h, const float* theta, cudaStream_t st) {
    const NtnDims& dm = h->dm;
    int64_t n = (int64_t)dm.R * dm.Kp * dm.Np;
    k_w_prep<<<cdiv(n, 256), 256, 0, st>>>(theta, h->side, h->Waug, dm);
    h->launches++;
}

static GemmTables tile_tables(ntn_handle* h) {
    return GemmTables{h->b.t_rel, h->b.t_u0, h->b.t_n, h->b.u_e1, h->b.u_e2, h->side};
}

void ntn_launch_gemm_fwd_simt(ntn_handle* h) {
    const NtnDims& dm = h->dm;
    if (h->b.ntiles == 0) return;
    dim3 g(h->b.ntiles * 2, cdiv(dm.Np, GBN));
    k_gemm_fwd_simt<<<g, 256, 0, h->stream>>>(tile_tables(h), h->E, h->Waug, h->b.T, dm.Kp, dm.Np);
    h->launches++;
}

void ntn_launch_gemm_ge_simt(ntn_handle* h) {
    const NtnDims& dm = h->dm;
    if (h->b.ntiles == 0) return;
    dim3 g(h->b.ntiles * 2, cdiv(dm.Kp, GBN));
    k_gemm_ge_simt<<<g, 256, 0, h->stream>>>(tile_tables(h), h->b.M, h->Waug, h->b.GA, dm.Kp, dm.Np);
    h->launches++;
}

void ntn_launch_gemm_dw_simt(ntn_handle* h, cudaStream_t st) {
    const NtnDims& dm = h->dm;
    if (h->b.nchunks == 0) return;
    GemmTables tb{h->b.ch_rel, h->b.ch_u0, h->b.ch_n, h->b.u_e1, h->b.u_e2, h->side};
    dim3 g(h->b.nchunks, cdiv(dm.Kp, GBM) * cdiv(dm.Np, GBN));
    k_gemm_dw_simt<<<g, 256, 0, st>>>(tb, h->E, h->b.M, h->b.dWpart, dm.Kp, dm.Np);
    h->launches++;
}

#define DISPATCH_KS(KSv, ...)                                              \
    switch (KSv) {                                                         \
        case 1: { constexpr int KS = 1; __VA_ARGS__; } break;              \
        case 2: { constexpr int KS = 2; __VA_ARGS__; } break;              \
        case 3: { constexpr int KS = 3; __VA_ARGS__; } break;              \
        case 4: { constexpr int KS = 4; __VA_ARGS__; } break;              \
        case 5: { constexpr int KS = 5; __VA_ARGS__; } break;              \
        case 6: { constexpr int KS = 6; __VA_ARGS__; } break;              \
        case 7: { constexpr int KS = 7; __VA_ARGS__; } break;              \
        default: { constexpr int KS = 8; __VA_ARGS__; } break;             \
    }
#define DISPATCH_CH(CHv, ...)                                              \
    if ((CHv) <= 1) { constexpr int CH = 1; __VA_ARGS__; } else { constexpr int CH = 2; __VA_ARGS__; }

void ntn_launch_score(ntn_handle* h, const float* theta) {
    const NtnDims& dm = h->dm;
    if (h->b.ntiles == 0) return;
    float* const m_lo = h->b.m_split ? h->b.M_lo : nullptr;
    ScoreArgs a{theta, h->E, h->b.T, h->b.M, m_lo, h->b.acoef, h->b.ccoef, h->b.GV, h->b.Nu, h->b.st_rel, h->b.st_u0, h->b.st_n,
                h->b.u_e1, h->b.u_e2, h->b.c_off, h->b.c_e3, h->side, h->b.tp_cost, h->b.tp_nact, h->b.tp_dU,
                dm.d, dm.dq, dm.Kp, dm.Np, dm.oU};
    int ch = cdiv(dm.dq / 4, 32);
    const bool tanh_act = h->cfg.activation == NTN_ACT_TANH;
    if (!h->b.use_gv) {
        // default: lane-per-(triple, slice) kernel writing packed coefficient records (tiles of ntn_score2_rows(k) triples)
        Score2Args s2{theta, h->E, h->b.T, h->b.M, m_lo, h->b.crec, h->b.Nu, h->b.RS, h->b.st_rel, h->b.st_u0, h->b.st_n,
                      h->b.u_e1, h->b.u_e2, h->b.c_off, h->b.c_e3, h->side, h->b.tp_cost, h->b.tp_nact, h->b.tp_dU,
                      dm.d, dm.dq, dm.Kp, dm.Np, dm.oU};
#define SC2_LAUNCH(KSv, ACTv, DQv)                                                                                       \
        do {                                                                                                             \
            const size_t smem = (size_t)SC2_WARPS * Score2Shape<KSv>::G * (KSv + SC2_NC + 1) * dm.dq * sizeof(float);    \
            if (smem > 48 * 1024) cudaFuncSetAttribute(k_score2<KSv, ACTv, DQv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            k_score2<KSv, ACTv, DQv><<<h->b.nstiles, SC2_WARPS * 32, smem, h->stream>>>(s2);                             \
        } while (0)
        if (dm.k == 3 && dm.dq == 100) {
            if (tanh_act) SC2_LAUNCH(3, NTN_ACT_TANH, 100); else SC2_LAUNCH(3, NTN_ACT_SIGMOID, 100);
        } else if (dm.k == 8 && dm.dq == 256) {
            if (tanh_act) SC2_LAUNCH(8, NTN_ACT_TANH, 256); else SC2_LAUNCH(8, NTN_ACT_SIGMOID, 256);
        } else {
            DISPATCH_KS(dm.k, { if (tanh_act) SC2_LAUNCH(KS, NTN_ACT_TANH, 0); else SC2_LAUNCH(KS, NTN_ACT_SIGMOID, 0); });
        }
#undef SC2_LAUNCH
        h->launches++;
        return;
    }
    DISPATCH_KS(dm.k, DISPATCH_CH(ch, {
        if (h->b.use_gv) {
            if (tanh_act) k_score<KS, CH, NTN_ACT_TANH, true><<<h->b.nstiles, 256, 0, h->stream>>>(a);
            else k_score<KS, CH, NTN_ACT_SIGMOID, true><<<h->b.nstiles, 256, 0, h->stream>>>(a);
        } else {
            if (tanh_act) k_score<KS, CH, NTN_ACT_TANH, false><<<h->b.nstiles, 256, 0, h->stream>>>(a);
            else k_score<KS, CH, NTN_ACT_SIGMOID, false><<<h->b.nstiles, 256, 0, h->stream>>>(a);
        }
    }));
    h->launches++;
}

int ntn_score2_rows(int k) {
    int g = 4;
    DISPATCH_KS(k, { g = Score2Shape<KS>::G; });
    return SC2_WARPS * g;
}

int ntn_pull_ranges(const ntn_handle* h) { return (int)(((int64_t)h->b.n_inc + h->dm.Ne) >> 5) + 1; }

// range table of k_entity_pull3, built once per batch job from the incident CSR (b.pull_rng holds ntn_pull_ranges(h) records)
void ntn_launch_pull_ranges(ntn_handle* h) {
    const DeviceBatch& b = h->b;
    cudaMemsetAsync(b.pull_rng, 0xff, (size_t)ntn_pull_ranges(h) * sizeof(int4), h->stream);
    k_pull_ranges<<<cdiv(h->dm.Ne, 256), 256, 0, h->stream>>>(b.inc_off, b.n_ent_hub_segs ? b.eh_count : nullptr, h->dm.Ne, b.pull_rng);
}

// Entity pull.  xge != null (data parallel, entity-gradient exchange): the entities are cut into contiguous row chunks;
// the kernel of a chunk also stores every group another rank owns into that rank's scratch, and the chunk's all-reduce
// runs on the exchange stream while the next chunk is pulled; the caller waits for ev_xdone before it reads gE.
void ntn_launch_entity_pull(ntn_handle* h, const ExchangeBinding* xge) {
    const NtnDims& dm = h->dm;
    const DeviceBatch& b = h->b;
    const int cpr = dm.dq / 4;
    const int C = xge ? std::max(1, std::min(NTN_MAX_WORD_CHUNKS, h->xge_chunks)) : 1;
    // (the warp-per-range kernel is the opt-in variant: at 20,000 x 10 it measured 48.7 us against 42.3 us of the
    //  thread-per-(entity, chunk) kernel -- 83 warp instructions per incident and 26 % occupancy, ncu profiles/r2_*)
    static const int pull_variant = [] { const char* e = getenv("NTN_PULL"); return e ? atoi(e) : 2; }();
    if (!xge && !b.use_gv && pull_variant == 3 && dm.dq <= 256) {
        Pull2Args a2{b.T, b.crec, b.GA, b.inc, b.inc_off, b.ent_order, h->gE,
                     b.n_ent_hub_segs ? b.eh_first : nullptr, b.n_ent_hub_segs ? b.eh_count : nullptr,
                     b.eh_begin, b.eh_end, b.eh_part, dm.Ne, dm.dq, dm.Kp, dm.Np, b.RS, b.n_ent_hub_segs,
                     make_fastdiv((unsigned)(dm.dq / 4))};
        Pull3Args a{b.T, b.crec, b.GA, b.inc, b.inc_off, b.pull_rng, h->gE, b.n_ent_hub_segs ? b.eh_first : nullptr,
                    b.n_ent_hub_segs ? b.eh_count : nullptr, b.eh_part, ntn_pull_ranges(h), dm.dq, dm.Kp, dm.Np, b.RS};
        const int blocks = cdiv((int64_t)a.nranges * 32, 128);
        DISPATCH_KS(dm.k, {
            if (b.n_ent_hub_segs) {
                k_entity_pull2_hub<KS><<<cdiv((int64_t)b.n_ent_hub_segs * cpr, 256), 256, 0, h->stream>>>(a2);
                h->launches++;
            }
            if (cpr <= 32) k_entity_pull3<KS, 1><<<blocks, 128, 0, h->stream>>>(a);
            else k_entity_pull3<KS, 2><<<blocks, 128, 0, h->stream>>>(a);
        });
        h->launches++;
        return;
    }
    Pull2Args a2{b.T, b.crec, b.GA, b.inc, b.inc_off, xge ? nullptr : b.ent_order, h->gE,
                 b.n_ent_hub_segs ? b.eh_first : nullptr, b.n_ent_hub_segs ? b.eh_count : nullptr,
                 b.eh_begin, b.eh_end, b.eh_part, dm.Ne, dm.dq, dm.Kp, dm.Np, b.RS, b.n_ent_hub_segs,
                 make_fastdiv((unsigned)cpr)};
    PullArgs a1{b.T, b.acoef, b.GV, b.GA, b.inc, b.inc_off, h->side, h->gE,
                b.n_ent_hub_segs ? b.eh_first : nullptr, b.n_ent_hub_segs ? b.eh_count : nullptr,
                b.eh_begin, b.eh_end, b.eh_part, dm.Ne, dm.dq, dm.Kp, dm.Np, b.n_ent_hub_segs, make_fastdiv((unsigned)cpr)};
    // hub entities: their segments are pre-reduced first (one kernel for all chunks)
    if (b.n_ent_hub_segs) {
        const int hb = cdiv((int64_t)b.n_ent_hub_segs * cpr, 256);
        DISPATCH_KS(dm.k, {
            if (b.use_gv) k_entity_pull_hub<KS, true><<<hb, 256, 0, h->stream>>>(a1);
            else k_entity_pull2_hub<KS><<<hb, 256, 0, h->stream>>>(a2);
        });
        h->launches++;
    }
    for (int c = 0; c < C; ++c) {
        const int row0 = (int)((int64_t)dm.Ne * c / C), nrows = (int)((int64_t)dm.Ne * (c + 1) / C) - row0;
        const XPushRange xp = ntn_push_range(h, xge, (int64_t)row0 * dm.dq, (int64_t)nrows * dm.dq, c);
        const bool push = xp.world != 0;
        const int nb = cdiv((int64_t)nrows * cpr, 256);
        if (nb > 0) {
            DISPATCH_KS(dm.k, {
                if (b.use_gv) {
                    if (push) k_entity_pull<KS, true, true><<<nb, 256, 0, h->stream>>>(a1, xp, row0, nrows);
                    else k_entity_pull<KS, true, false><<<nb, 256, 0, h->stream>>>(a1, xp, row0, nrows);
                } else {
                    if (push) k_entity_pull2<KS, true><<<nb, 256, 0, h->stream>>>(a2, xp, row0, nrows);
                    else k_entity_pull2<KS, false><<<nb, 256, 0, h->stream>>>(a2, xp, row0, nrows);
                }
            });
            h->launches++;
        }
        if (xge) {
            cudaEventRecord(h->ev_x[c], h->stream);
            cudaStreamWaitEvent(h->stream4, h->ev_x[c], 0);
            const XRange none{0, 0, 0, false};
            ntn_launch_exchange(h, xge, XRange{(int64_t)row0 * dm.dq, (int64_t)nrows * dm.dq, c, push}, none, none, c + 1 == C, h->stream4);
        }
    }
    if (xge) cudaEventRecord(h->ev_xdone, h->stream4);
}

// The word pull covers the Nw words in C chunks of whole words (C = 1 unless a gradient exchange is bound): chunk c is
// words [w0, w0 + wn), one thread per (word pair, float4 chunk), and owns `blocks` regulariser-partial slots from block_off.
// (Equal chunks: an exchange kernel costs ~10 us of launch + flag barrier before it moves a byte, so the schedule that
// ends first is the one with the fewest kernels that still overlaps half of the bytes -- two chunks, measured.)
int ntn_word_blocks_total(const NtnDims& dm, int C);
void ntn_word_chunk(const NtnDims& dm, int c, int C, int* w0, int* wn, int* block_off, int* blocks) {
    const int cpr = dm.dq / 4;
    int64_t bound[NTN_MAX_WORD_CHUNKS + 1];
    bound[0] = 0;
    for (int i = 1; i <= C; ++i) bound[i] = (i == C) ? dm.Nw : ((int64_t)dm.Nw * i + C - 1) / C;     // equal chunks
    int off = 0;
    for (int i = 0; i < c; ++i) off += cdiv(((bound[i + 1] - bound[i] + 1) / 2) * cpr, 256);
    *w0 = (int)bound[c]; *wn = (int)(bound[c + 1] - bound[c]);
    *block_off = off; *blocks = cdiv(((int64_t)(*wn + 1) / 2) * cpr, 256);
}
// test hook (CPU): the chunk layout for Nw words of dq floats in C chunks -> out[4 c .. 4 c + 3] = {first word, words,
// first regulariser-partial slot, blocks}; returns the total number of slots
extern "C" int ntn_host_word_chunks(int32_t Nw, int32_t dq, int32_t C, int32_t* out) {
    if (Nw < 1 || dq < 4 || (dq & 3) || C < 1 || C > NTN_MAX_WORD_CHUNKS || !out) return NTN_EINVAL;
    NtnDims dm{};
    dm.Nw = Nw; dm.dq = dq;
    for (int c = 0; c < C; ++c) ntn_word_chunk(dm, c, C, out + 4 * c, out + 4 * c + 1, out + 4 * c + 2, out + 4 * c + 3);
    return ntn_word_blocks_total(dm, C);
}

int ntn_word_blocks_total(const NtnDims& dm, int C) {
    int w0, wn, off, nb;
    ntn_word_chunk(dm, C - 1, C, &w0, &wn, &off, &nb);
    return off + nb;
}

void ntn_launch_word_pull(ntn_handle* h, const float* theta, float* grad, const ExchangeBinding* xb) {
    const NtnDims& dm = h->dm;
    const DeviceWords& w = h->w;
    const int C = h->word_chunks;
    WordArgs a{h->gE, w.len_bwd /* reciprocals */, theta, grad, w.w_off, w.w_ent, w.w_rec, w.wh_first, w.wh_count, w.wh_begin, w.wh_end,
               w.hub_words, w.wh_part, h->reg_part, dm.Nw, dm.d, dm.dq, w.n_hub_segs, w.n_hub_words, ntn_word_blocks_total(dm, C),
               // (entity-gradient exchange: gE already holds the sum over the ranks, so every rank computes the FULL word
               //  gradient and nothing adds the word block up afterwards -- it takes lambda undivided)
               dm.oWVint, 1.0f / (float)h->cfg.batch_divisor, h->cfg.lambda * (h->ge_mode ? 1.0f : h->cfg.reg_scale), make_fastdiv((unsigned)(dm.dq / 4))};
    const int cpr = dm.dq / 4;
    int ch = cdiv(cpr, 32);
    // range ids of an evaluation (their scratch regions, exchange.cu): 0 = the W/V/b/U blocks, 1 + c = word chunk c,
    // C + 1 = the {cost, n_active} tail
    XPushAll xa{};
    for (int c = 0; c < C; ++c) {
        int w0, wn, off, nb;
        ntn_word_chunk(dm, c, C, &w0, &wn, &off, &nb);
        xa.r[c] = ntn_push_range(h, xb, dm.oWVint + (int64_t)w0 * dm.dq, (int64_t)wn * dm.dq, 1 + c);
        xa.w_end[c] = w0 + wn;
    }
    xa.n = (xb && xa.r[0].world) ? C : 0;
    const bool pushed = xa.n != 0;
    // hub words: a forked branch that only needs gE and writes rows the main kernel skips.  With a chunked exchange the
    // branch runs on the exchange's own (high-priority) stream AHEAD of the first chunk's exchange: hub rows lie in any
    // chunk, and on the low-priority branch stream they would only finish when the main pull has drained.
    cudaStream_t hub_st = (xb && C > 1) ? h->stream4 : h->stream2;
    if (w.n_hub_segs) {
        cudaEventRecord(h->ev_fork, h->stream);
        cudaStreamWaitEvent(hub_st, h->ev_fork, 0);
        k_word_pull_hub<<<cdiv((int64_t)w.n_hub_segs * cpr, 256), 256, 0, hub_st>>>(a);
        DISPATCH_CH(ch, {
            if (pushed) k_word_finish_hub<CH, true><<<w.n_hub_words, 1024, 0, hub_st>>>(a, xa);
            else k_word_finish_hub<CH, false><<<w.n_hub_words, 1024, 0, hub_st>>>(a, xa);
        });
        cudaEventRecord(h->ev_join, hub_st);
        h->launches += 2;
    }
    for (int c = 0; c < C; ++c) {
        int w0, wn, off, nb;
        ntn_word_chunk(dm, c, C, &w0, &wn, &off, &nb);
        if (nb > 0) {
            if (pushed) k_word_pull<true><<<nb, 256, 0, h->stream>>>(a, xa.r[c], w0, wn, (wn + 1) / 2, off);
            else k_word_pull<false><<<nb, 256, 0, h->stream>>>(a, xa.r[c], w0, wn, (wn + 1) / 2, off);
            h->launches++;
        }
        if (xb && c + 1 < C) {
            // chunk c is final in this rank's buffer (the hub rows precede it in stream4's order): exchange it on the side
            // stream while the pull goes on with chunk c + 1.  The W/V/b/U blocks -- finished by a low-priority branch that
            // ends with the evaluation -- and the tail go out with the last chunk.
            cudaEventRecord(h->ev_x[c], h->stream);
            cudaStreamWaitEvent(h->stream4, h->ev_x[c], 0);
            const XRange none{0, 0, 0, false};
            ntn_launch_exchange(h, xb, XRange{dm.oWVint + (int64_t)w0 * dm.dq, (int64_t)wn * dm.dq, 1 + c, pushed}, none, none, false, h->stream4);
        }
    }
    if (xb && C > 1) cudaEventRecord(h->ev_xdone, h->stream4);
    if (w.n_hub_segs) cudaStreamWaitEvent(h->stream, h->ev_join, 0);
}

int ntn_small_blocks(const NtnDims& dm) {
    int64_t nsmall = (int64_t)dm.R * dm.Kp * dm.Np + (int64_t)dm.R * dm.k + 4;
    return cdiv(nsmall, 256);
}
int ntn_word_blocks(const NtnDims& dm) { return ntn_word_blocks_total(dm, 1) + NTN_MAX_WORD_CHUNKS; }   // capacity (any chunking)

void ntn_launch_finalize_small(ntn_handle* h, const float* theta, float* grad, cudaStream_t st) {
    const NtnDims& dm = h->dm;
    const DeviceBatch& b = h->b;
    int nb = ntn_small_blocks(dm);
    int wb = ntn_word_blocks_total(dm, h->word_chunks) + h->w.n_hub_words;
    FinArgs a{theta, b.dWpart, b.tp_dU, grad, b.relch_off, b.rels_off, h->side, h->reg_part + wb, dm,
              1.0f / (float)h->cfg.batch_divisor, h->cfg.lambda * h->cfg.reg_scale};
    k_finalize_small<<<nb, 256, 0, st>>>(a);
    k_finalize_u<<<dm.R * dm.k, 256, 0, st>>>(a, h->reg_part + wb + nb);
    h->launches += 2;
}

void ntn_launch_final_reduce(ntn_handle* h, float* tail) {
    const NtnDims& dm = h->dm;
    const DeviceBatch& b = h->b;
    int nb = ntn_small_blocks(dm);
    int wb = ntn_word_blocks_total(dm, h->word_chunks) + h->w.n_hub_words;
    double lam = (double)h->cfg.lambda * (double)h->cfg.reg_scale;
    k_final_reduce<<<1, 1024, 0, h->stream>>>(b.tp_cost, b.tp_nact, b.nstiles, h->reg_part, wb + nb + dm.R * dm.k,
                                             1.0 / (double)h->cfg.batch_divisor, 0.5 * lam, h->d_out, tail);
    h->launches++;
}

void ntn_launch_to_internal(ntn_handle* h, const float* d_ref, float* d_int) {
    const NtnDims& dm = h->dm;
    k_copy_small<<<cdiv(dm.oWVint, 256), 256, 0, h->stream>>>(d_ref, d_int, dm.oWVref, dm.oWVint);
    dim3 g(cdiv(dm.Nw, 32), cdiv(dm.dq, 32));
    k_wv_to_internal<<<g, 256, 0, h->stream>>>(d_ref + dm.oWVref, d_int + dm.oWVint, dm.d, dm.Nw, dm.dq);
    h->launches += 2;
}

void ntn_launch_from_internal(ntn_handle* h, const float* d_int, float* d_ref) {
    const NtnDims& dm = h->dm;
    k_copy_small<<<cdiv(dm.oWVref, 256), 256, 0, h->stream>>>(d_int, d_ref, dm.oWVref, dm.oWVref);
    dim3 g(cdiv(dm.Nw, 32), cdiv(dm.dq, 32));
    k_wv_from_internal<<<g, 256, 0, h->stream>>>(d_int + dm.oWVint, d_ref + dm.oWVref, dm.d, dm.Nw, dm.dq);
    h->launches += 2;
}

void ntn_launch_transpose_wv_host_layout(ntn_handle* h, const float* d_wv_ref, float* d_wvt) {
    const NtnDims& dm = h->dm;
    dim3 g(cdiv(dm.Nw, 32), cdiv(dm.dq, 32));
    k_wv_to_internal<<<g, 256, 0, h->stream>>>(d_wv_ref, d_wvt, dm.d, dm.Nw, dm.dq);
}

void ntn_launch_score_eval(ntn_handle* h, const float* theta, const int* e1, const int* rel, const int* e2, int64_t n,
                           int mode, double* out) {
    if (n == 0) return;
    if (h->cfg.activation == NTN_ACT_TANH)
        k_score_eval<NTN_ACT_TANH><<<cdiv(n * 32, 256), 256, 0, h->stream>>>(theta, h->E, e1, rel, e2, n, mode, out, h->dm);
    else
        k_score_eval<NTN_ACT_SIGMOID><<<cdiv(n * 32, 256), 256, 0, h->stream>>>(theta, h->E, e1, rel, e2, n, mode, out, h->dm);
}
