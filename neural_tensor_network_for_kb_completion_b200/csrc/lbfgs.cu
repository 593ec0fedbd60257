// Device-resident L-BFGS over the flattened parameter vector (north-star item 3): theta, the gradient,
// the curvature pairs and the search direction all stay in HBM; the host only sees scalars.
//
// The reference calls `new LBFGSMinimizer().minimize(t, theta, opts)` with opts.maxIters = 5 and a fresh
// minimiser (empty history) per batch job (/root/reference/src/Run_NTN.java:119-124).  The minimiser itself
// lives in umass-nlp-1.0-SNAPSHOT.jar, which is not in the reference tree, so its line-search constants and
// history size are not visible: PARITY UNPINNED.  This file defines the algorithm used here:
//   two-loop recursion, history m (default 6), H0 = (s.y / y.y) I, first iteration d = -g / ||g||;
//   backtracking line search: t0 = 1, Armijo constant 1e-4, shrink 0.5 (0.1 on the first iteration),
//   at most 20 trials; stop after max_iters iterations or when |f_new - f_old| <= tol * (|f_new|+|f_old|+eps)/2.
// Every trial point costs one full cost+gradient evaluation (ntn_eval_dev) with its own corrupt-side draw,
// exactly like the reference (NTN.java:146 draws inside computeAt).
//
// Vector work per iteration -- three kernels, no host round trip of their own:
//   k_pair_dots   ONE pass forms the new curvature pair s = trial - theta, y = g_new - g into the free slot of the
//                 history ring AND every inner product the recursion will need: s, y and g_new against all stored
//                 S_j, Y_j, and among themselves (FP64 accumulation, fixed two-stage order -> deterministic);
//   k_two_loop    one block: reduces the partials, updates the Gram matrix G = <b_i, b_j> of the basis
//                 b = [S_0..S_m, Y_0..Y_m, g], decides whether the pair is kept (y.s > 0), and runs the two-loop
//                 recursion ON THE GRAM MATRIX: the direction is a linear combination dir = sum_j delta_j b_j, so the
//                 recursion costs O(m^2) scalar work instead of 4m passes over theta-sized vectors; also g.dir and the
//                 first step length;
//   k_lincomb     dir = sum_j delta_j b_j and the first trial point trial = theta - t0 * dir in the same pass.
// That is (2m' + 4) + (2m' + 2) vector reads and 4 writes per iteration for m' stored pairs, against (4m' + 6) reads
// and 2m' + ... writes of the textbook vector recursion, and 3 + (trials - 1) launches.  theta/trial and g/g_new swap by
// pointer when a step is accepted.  The host synchronises once per trial point (it needs the cost for the Armijo
// test); the scalars of k_two_loop (g.dir, t0, history state) ride to pinned memory before that evaluation and are
// read at the same synchronisation.
#include <math.h>
#include <string.h>

#include <algorithm>

#include "ntn_internal.h"

int ntn_tc_check_error(ntn_handle* h);                          // gemm_tcgen05.cu

#define LB_BLOCKS 296            // 2 blocks per SM
#define LB_THREADS 256
#define LB_MAX_HIST 8
#define LB_MAX_SLOTS (LB_MAX_HIST + 1)                  // one free slot: a rejected pair never overwrites a kept one
#define LB_MAX_BASIS (2 * LB_MAX_SLOTS + 1)
#define LB_MAX_DOTS (6 * LB_MAX_SLOTS + 6)

__device__ __forceinline__ double lb_block_sum(double v, double* sm) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) for (int w = 0; w < LB_THREADS / 32; ++w) t += sm[w];
    return t;
}

__device__ __forceinline__ double dot4d(const float4& a, const float4& b) {
    // products of two floats are exact in double; the four of an element group are added pairwise
    return ((double)a.x * (double)b.x + (double)a.y * (double)b.y) + ((double)a.z * (double)b.z + (double)a.w * (double)b.w);
}

// device-resident optimiser state
struct LbState {
    int hist, head, m, accepted, converged, pad;        // ring of m + 1 slots: newest pair at head - 1, slot `head` is free
    double gd, t0, gg, sy;                              // g.dir (of -dir: negative), first step, g.g, last y.s
    double G[LB_MAX_BASIS][LB_MAX_BASIS];
    double delta[LB_MAX_BASIS];
};

struct PairArgs {
    const float *theta, *trial, *g, *gnew;
    float *S, *Y;                 // slot j at S + j * stride
    int64_t n4, stride;
    int slot_new;                 // free slot that receives (s, y); -1: no pair (initial call: only g_new . g_new)
    double* part;                 // [dot index][LB_BLOCKS]
};

// dot layout: for stored slot j < NS: [6j + 0..5] = s.S_j, s.Y_j, y.S_j, y.Y_j, gn.S_j, gn.Y_j (slot_new's entries unused);
// then [6 NS + 0..5] = s.s, s.y, y.y, s.gn, y.gn, gn.gn
template <int NS>
__global__ void __launch_bounds__(LB_THREADS) k_pair_dots(PairArgs a) {
    __shared__ double sm[LB_THREADS / 32];
    double acc[6 * NS + 6];
#pragma unroll
    for (int i = 0; i < 6 * NS + 6; ++i) acc[i] = 0.0;
    const float4* th4 = reinterpret_cast<const float4*>(a.theta);
    const float4* tr4 = reinterpret_cast<const float4*>(a.trial);
    const float4* g4 = reinterpret_cast<const float4*>(a.g);
    const float4* gn4 = reinterpret_cast<const float4*>(a.gnew);
    const bool pair = a.slot_new >= 0;
    for (int64_t i = (int64_t)blockIdx.x * LB_THREADS + threadIdx.x; i < a.n4; i += (int64_t)gridDim.x * LB_THREADS) {
        const float4 gn = __ldg(gn4 + i);
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f), y = s;
        if (pair) {
            const float4 th = __ldg(th4 + i), tr = __ldg(tr4 + i), g = __ldg(g4 + i);
            s = make_float4(tr.x - th.x, tr.y - th.y, tr.z - th.z, tr.w - th.w);
            y = make_float4(gn.x - g.x, gn.y - g.y, gn.z - g.z, gn.w - g.w);
        }
        float4 sv[NS > 0 ? NS : 1], yv[NS > 0 ? NS : 1];
#pragma unroll
        for (int j = 0; j < NS; ++j) {               // the old content of slot_new is read too (and ignored): no branch
            sv[j] = reinterpret_cast<const float4*>(a.S + (size_t)j * a.stride)[i];
            yv[j] = reinterpret_cast<const float4*>(a.Y + (size_t)j * a.stride)[i];
        }
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            acc[6 * j + 0] += dot4d(s, sv[j]); acc[6 * j + 1] += dot4d(s, yv[j]);
            acc[6 * j + 2] += dot4d(y, sv[j]); acc[6 * j + 3] += dot4d(y, yv[j]);
            acc[6 * j + 4] += dot4d(gn, sv[j]); acc[6 * j + 5] += dot4d(gn, yv[j]);
        }
        acc[6 * NS + 0] += dot4d(s, s); acc[6 * NS + 1] += dot4d(s, y); acc[6 * NS + 2] += dot4d(y, y);
        acc[6 * NS + 3] += dot4d(s, gn); acc[6 * NS + 4] += dot4d(y, gn); acc[6 * NS + 5] += dot4d(gn, gn);
        if (pair) {
            reinterpret_cast<float4*>(a.S + (size_t)a.slot_new * a.stride)[i] = s;
            reinterpret_cast<float4*>(a.Y + (size_t)a.slot_new * a.stride)[i] = y;
        }
    }
#pragma unroll
    for (int q = 0; q < 6 * NS + 6; ++q) {
        const double t = lb_block_sum(acc[q], sm);
        if (threadIdx.x == 0) a.part[(size_t)q * LB_BLOCKS + blockIdx.x] = t;
    }
}

// One block.  Reduces the partial dots (one warp per dot, fixed order), updates the Gram matrix and the history state,
// then thread 0 runs the two-loop recursion on the Gram matrix.
__global__ void __launch_bounds__(256) k_two_loop(LbState* st, const double* __restrict__ part, int NS, int slot_new, int nblocks,
                                                  double* __restrict__ scal /* pinned mirror source: 8 doubles */) {
    __shared__ double dots[LB_MAX_DOTS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nd = 6 * NS + 6;
    for (int q = warp; q < nd; q += 8) {
        double v = 0.0;
        for (int b = lane; b < nblocks; b += 32) v += part[(size_t)q * LB_BLOCKS + b];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) dots[q] = v;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int m = st->m, nslot = m + 1;
    const int IS = 0, IY = LB_MAX_SLOTS, IG = 2 * LB_MAX_SLOTS;      // basis indices: S_j, Y_j, g
    double (*G)[LB_MAX_BASIS] = st->G;
    const double* tail = dots + 6 * NS;
    // ---- Gram update: the current gradient is now g_new
    for (int j = 0; j < NS; ++j) {
        if (j == slot_new) continue;
        G[IG][IS + j] = G[IS + j][IG] = dots[6 * j + 4];
        G[IG][IY + j] = G[IY + j][IG] = dots[6 * j + 5];
    }
    G[IG][IG] = tail[5];
    st->gg = tail[5];
    st->accepted = 0;
    if (slot_new >= 0) {
        const int a = slot_new;
        for (int j = 0; j < NS; ++j) {
            if (j == a) continue;
            G[IS + a][IS + j] = G[IS + j][IS + a] = dots[6 * j + 0];
            G[IS + a][IY + j] = G[IY + j][IS + a] = dots[6 * j + 1];
            G[IY + a][IS + j] = G[IS + j][IY + a] = dots[6 * j + 2];
            G[IY + a][IY + j] = G[IY + j][IY + a] = dots[6 * j + 3];
        }
        G[IS + a][IS + a] = tail[0];
        G[IS + a][IY + a] = G[IY + a][IS + a] = tail[1];
        G[IY + a][IY + a] = tail[2];
        G[IG][IS + a] = G[IS + a][IG] = tail[3];
        G[IG][IY + a] = G[IY + a][IG] = tail[4];
        st->sy = tail[1];
        const double rho = 1.0 / tail[1];
        if (tail[1] > 0.0 && isfinite(rho)) {                        // keep the pair: the free slot moves on
            st->accepted = 1;
            st->head = (st->head + 1) % nslot;
            if (st->hist < m) st->hist++;
        }
    }
    // ---- two-loop recursion on coefficient vectors over the basis
    double* c = st->delta;
    double alpha[LB_MAX_HIST];
    for (int restart = 0; restart < 2; ++restart) {
        for (int j = 0; j < LB_MAX_BASIS; ++j) c[j] = 0.0;
        c[IG] = 1.0;                                                 // q = g
        const int hist = st->hist, head = st->head;
        for (int i = 0; i < hist; ++i) {                             // newest to oldest
            const int a = (head - 1 - i + 2 * nslot) % nslot;
            double sq = 0.0;
            for (int j = 0; j < LB_MAX_BASIS; ++j) sq += c[j] * G[IS + a][j];
            alpha[i] = sq / G[IS + a][IY + a];                       // rho_a * s_a.q
            c[IY + a] -= alpha[i];
        }
        if (hist > 0) {
            const int a = (head - 1 + nslot) % nslot;
            const double gamma = G[IS + a][IY + a] / G[IY + a][IY + a];   // H0 = (s.y / y.y) I
            for (int j = 0; j < LB_MAX_BASIS; ++j) c[j] *= gamma;
        }
        for (int i = hist - 1; i >= 0; --i) {                        // oldest to newest
            const int a = (head - 1 - i + 2 * nslot) % nslot;
            double yr = 0.0;
            for (int j = 0; j < LB_MAX_BASIS; ++j) yr += c[j] * G[IY + a][j];
            c[IS + a] += alpha[i] - yr / G[IS + a][IY + a];
        }
        double gdir = 0.0;
        for (int j = 0; j < LB_MAX_BASIS; ++j) gdir += c[j] * G[IG][j];
        st->gd = -gdir;
        st->t0 = 1.0;
        if (hist == 0) st->t0 = st->gg > 0.0 ? 1.0 / sqrt(st->gg) : 1.0;       // unit-norm steepest descent
        if (st->gd < 0.0) break;
        if (hist == 0) { st->converged = 1; break; }                 // g.g <= 0: nowhere to go
        st->hist = 0;                                                // not a descent direction: restart from steepest descent
    }
    scal[0] = st->gd; scal[1] = st->t0; scal[2] = (double)st->hist; scal[3] = (double)st->head;
    scal[4] = (double)st->accepted; scal[5] = (double)st->converged; scal[6] = st->gg; scal[7] = st->sy;
}

struct CombArgs {
    const float *theta, *g, *S, *Y;
    float *dir, *trial;
    int64_t n4, stride;
    const LbState* st;
};

template <int NS>
__global__ void __launch_bounds__(LB_THREADS) k_lincomb(CombArgs a) {
    float cs[NS > 0 ? NS : 1], cy[NS > 0 ? NS : 1];
#pragma unroll
    for (int j = 0; j < NS; ++j) { cs[j] = (float)a.st->delta[j]; cy[j] = (float)a.st->delta[LB_MAX_SLOTS + j]; }
    const float cg = (float)a.st->delta[2 * LB_MAX_SLOTS];
    const float t0 = (float)a.st->t0;
    const float4* th4 = reinterpret_cast<const float4*>(a.theta);
    const float4* g4 = reinterpret_cast<const float4*>(a.g);
    for (int64_t i = (int64_t)blockIdx.x * LB_THREADS + threadIdx.x; i < a.n4; i += (int64_t)gridDim.x * LB_THREADS) {
        const float4 g = __ldg(g4 + i), th = __ldg(th4 + i);
        float4 sv[NS > 0 ? NS : 1], yv[NS > 0 ? NS : 1];
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            sv[j] = __ldg(reinterpret_cast<const float4*>(a.S + (size_t)j * a.stride) + i);
            yv[j] = __ldg(reinterpret_cast<const float4*>(a.Y + (size_t)j * a.stride) + i);
        }
        float4 d = make_float4(cg * g.x, cg * g.y, cg * g.z, cg * g.w);
#pragma unroll
        for (int j = 0; j < NS; ++j) {               // fixed order; a free or dropped slot has coefficient 0
            d.x = fmaf(cs[j], sv[j].x, d.x); d.y = fmaf(cs[j], sv[j].y, d.y); d.z = fmaf(cs[j], sv[j].z, d.z); d.w = fmaf(cs[j], sv[j].w, d.w);
            d.x = fmaf(cy[j], yv[j].x, d.x); d.y = fmaf(cy[j], yv[j].y, d.y); d.z = fmaf(cy[j], yv[j].z, d.z); d.w = fmaf(cy[j], yv[j].w, d.w);
        }
        reinterpret_cast<float4*>(a.dir)[i] = d;
        reinterpret_cast<float4*>(a.trial)[i] = make_float4(fmaf(-t0, d.x, th.x), fmaf(-t0, d.y, th.y), fmaf(-t0, d.z, th.z), fmaf(-t0, d.w, th.w));
    }
}

// out = x + t*d  (host scalar t): the backtracking trials after the first
__global__ void __launch_bounds__(LB_THREADS) k_step(float* __restrict__ out, const float* __restrict__ x,
                                                     const float* __restrict__ d, int64_t n4, float t) {
    float4* o4 = reinterpret_cast<float4*>(out);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* d4 = reinterpret_cast<const float4*>(d);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 xv = __ldg(x4 + i), dv = __ldg(d4 + i);
        o4[i] = make_float4(fmaf(t, dv.x, xv.x), fmaf(t, dv.y, xv.y), fmaf(t, dv.z, xv.z), fmaf(t, dv.w, xv.w));
    }
}

__global__ void k_lb_init(LbState* st, int m) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        st->hist = 0; st->head = 0; st->m = m; st->accepted = 0; st->converged = 0; st->pad = 0;
        st->gd = 0.0; st->t0 = 1.0; st->gg = 0.0; st->sy = 0.0;
        for (int i = 0; i < LB_MAX_BASIS; ++i) { st->delta[i] = 0.0; for (int j = 0; j < LB_MAX_BASIS; ++j) st->G[i][j] = 0.0; }
    }
}

#define LB_DISPATCH_NS(NSv, KERNEL, ARGS, STREAM)                                                           \
    switch (NSv) {                                                                                          \
        case 0: KERNEL<0><<<LB_BLOCKS, LB_THREADS, 0, STREAM>>>(ARGS); break;                               \
        case 1: KERNEL<1><<<LB_BLOCKS, LB_THREADS, 0, STREAM>>>(ARGS); break;                               \
        case 2: KERNEL<2><<<LB_BLOCKS, LB_THREADS, 0, STREAM>>>(ARGS); break;                               \
        case 3: KERNEL<3><<<LB_BLOCKS, LB_THREADS, 0, STREAM>>>(ARGS); break;                               \
        case 4: KERNEL<4><<<LB_BLOCKS, LB_THREADS, 0, STREAM>>>(ARGS); break;                               \
        case 5: KERNEL<5><<<LB_BLOCKS, LB_THREADS, 0, STREAM>>>(ARGS); break;                               \
        case 6: KERNEL<6><<<LB_BLOCKS, LB_THREADS, 0, STREAM>>>(ARGS); break;                               \
        case 7: KERNEL<7><<<LB_BLOCKS, LB_THREADS, 0, STREAM>>>(ARGS); break;                               \
        case 8: KERNEL<8><<<LB_BLOCKS, LB_THREADS, 0, STREAM>>>(ARGS); break;                               \
        default: KERNEL<9><<<LB_BLOCKS, LB_THREADS, 0, STREAM>>>(ARGS); break;                              \
    }

extern "C" int ntn_lbfgs_minimize(ntn_handle* h, float* d_theta, const ntn_lbfgs_opts* o, ntn_sides_fn sides_fn, void* ctx,
                                  ntn_lbfgs_result* res) {
    if (!h) return NTN_EINVAL;
    if (!d_theta || !o || !sides_fn || !res) { h->err = "null argument"; return NTN_EINVAL; }
    cudaSetDevice(h->device);
    const NtnDims& dm = h->dm;
    const int m = o->history > 0 ? (o->history < LB_MAX_HIST ? o->history : LB_MAX_HIST) : 6;
    const int nslot = m + 1;
    const int max_iters = o->max_iters > 0 ? o->max_iters : 5;
    const double tol = o->tol > 0 ? o->tol : 1e-4;
    const double c1 = 1e-4;
    const int64_t n = dm.Pint;                     // multiple of 4 by construction (oWVint % 4 == 0, dq % 4 == 0)
    if (n % 4) { h->err = "internal dimension not a multiple of 4"; return NTN_EUNSUP; }
    const int64_t n4 = n / 4;
    cudaStream_t st = h->stream;
    // workspace: S[m+1] Y[m+1] g gnew dir trial | block partials | state | pinned-bound scalars; kept in the handle across
    // calls (stable pointers also let the evaluation graphs of (theta, g) / (trial, gnew) be reused)
    const size_t nvec = (size_t)2 * nslot + 4;
    const int64_t ns = n + 4;                      // vector stride: room for the {cost, n_active} tail after a gradient
    const size_t vec_bytes = nvec * ns * sizeof(float);
    const size_t part_bytes = (size_t)LB_MAX_DOTS * LB_BLOCKS * sizeof(double);
    const size_t need = vec_bytes + part_bytes + sizeof(LbState) + 64 * sizeof(double);
    if (h->lbfgs_ws_bytes < need) {
        cudaFree(h->lbfgs_ws); h->lbfgs_ws = nullptr; h->lbfgs_ws_bytes = 0;
        cudaError_t e = cudaMalloc(&h->lbfgs_ws, need);
        if (e != cudaSuccess) { h->err = std::string("lbfgs alloc: ") + cudaGetErrorString(e); return NTN_ECUDA; }
        h->lbfgs_ws_bytes = need;
    }
    if (!h->lbfgs_scal && cudaMallocHost((void**)&h->lbfgs_scal, 8 * sizeof(double)) != cudaSuccess) {
        h->err = "lbfgs: pinned alloc"; return NTN_ECUDA;
    }
    float* buf = (float*)h->lbfgs_ws;
    double* part = (double*)((char*)h->lbfgs_ws + vec_bytes);
    LbState* state = (LbState*)((char*)part + part_bytes);
    double* d_scal = (double*)((char*)state + sizeof(LbState));
    volatile double* hs = h->lbfgs_scal;
    float* S = buf; float* Y = buf + (size_t)nslot * ns;
    float* g = buf + (size_t)2 * nslot * ns; float* gnew = g + ns; float* dir = gnew + ns; float* spare = dir + ns;
    float* theta = d_theta; float* trial = spare;          // swapped by pointer when a step is accepted
    std::vector<uint8_t> side(dm.R);
    int rc = NTN_OK, evals = 0, iters = 0;
    int attempted = 0, head = 0, hist = 0;                 // host mirror of the history ring
    bool converged = false;
    double f = 0.0, fnew = 0.0;
    int64_t nact = 0;
    // single process: plain evaluation.  Data parallel: the evaluation writes the {cost, n_active} tail after the
    // gradient, the caller's reduce_fn all-reduces gradient + tail, and cost / n_active are read back from the tail.
    const bool saved_tail = h->grad_tail;
    if (h->grad_tail != (o->reduce_fn != nullptr)) ntn_set_grad_tail(h, o->reduce_fn != nullptr);
    auto evaluate = [&](const float* th, float* gr, double* cost) {
        sides_fn(ctx, side.data());
        ++evals;
        if (!o->reduce_fn) return ntn_eval_dev(h, th, side.data(), gr, cost, &nact);
        int rc2 = ntn_eval_dev(h, th, side.data(), gr, nullptr, nullptr);
        if (rc2) return rc2;
        o->reduce_fn(o->reduce_ctx, gr, n + 4);
        float tail[4];
        cudaMemcpyAsync(tail, gr + n, sizeof(tail), cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { h->err = "lbfgs: stream error after the gradient exchange"; return (int)NTN_ECUDA; }
        if (h->gemm_path_active == NTN_GEMM_TCGEN05 && ntn_tc_check_error(h)) {
            h->err = "tcgen05 pipeline: a bounded mbarrier wait expired (results invalid)"; return (int)NTN_ECUDA;
        }
        *cost = (double)tail[0] + (double)tail[1];
        nact = (int64_t)llround((double)tail[2]) * 4096 + (int64_t)llround((double)tail[3]);
        return (int)NTN_OK;
    };
    // pass 1 of an iteration: (pair,) dots, Gram update + two-loop; pass 2: direction + first trial point
    auto plan_direction = [&](bool pair) {
        const int NS = std::min(nslot, attempted);          // slots that have ever been written
        PairArgs pa{theta, trial, g, gnew, S, Y, n4, ns, pair ? head : -1, part};
        if (!pair) pa.gnew = g;                             // initial call: only g.g
        const int NSd = pair ? std::min(nslot, attempted + 1) : NS;
        if (pair) ++attempted;
        LB_DISPATCH_NS(NSd, k_pair_dots, pa, st);
        k_two_loop<<<1, 256, 0, st>>>(state, part, NSd, pa.slot_new, LB_BLOCKS, d_scal);
        cudaMemcpyAsync((void*)hs, d_scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, st);
        if (pair) { std::swap(theta, trial); std::swap(g, gnew); }      // the accepted point becomes the current one
        CombArgs ca{theta, g, S, Y, dir, trial, n4, ns, state};
        LB_DISPATCH_NS(NSd, k_lincomb, ca, st);
    };
    k_lb_init<<<1, 1, 0, st>>>(state, m);
    if ((rc = evaluate(theta, g, &f))) goto done;
    res->first_cost = f;
    plan_direction(false);
    for (iters = 0; iters < max_iters; ++iters) {
        // ---- backtracking line search along -dir; the first trial point is already in `trial`
        bool ok = false;
        double t = 1.0, gd = 0.0, shrink = 0.5;
        for (int ls = 0; ls < 20; ++ls) {
            if (ls > 0) k_step<<<LB_BLOCKS, LB_THREADS, 0, st>>>(trial, theta, dir, n4, (float)(-t));
            if ((rc = evaluate(trial, gnew, &fnew))) goto done;      // synchronises: the scalars of k_two_loop have landed too
            if (ls == 0) {
                gd = hs[0]; t = hs[1]; hist = (int)hs[2]; head = (int)hs[3];
                if (hs[5] != 0.0) { converged = true; break; }
                if (hist == 0) shrink = 0.1;
            }
            if (fnew <= f + c1 * t * gd) { ok = true; break; }
            t *= shrink;
        }
        if (converged) { --evals; break; }                 // g.g == 0 exactly: the probe evaluation is not counted
        if (!ok) break;                                    // line search failed: keep the current point
        const double fold = f;
        f = fnew;
        const bool stop = fabs(fold - f) <= tol * (fabs(fold) + fabs(f) + 1e-10) / 2.0;
        if (stop || iters + 1 == max_iters) {              // no further direction needed: just adopt the point
            std::swap(theta, trial); std::swap(g, gnew);
            if (stop) converged = true;
            ++iters;
            break;
        }
        plan_direction(true);
    }
done:
    if (theta != d_theta) cudaMemcpyAsync(d_theta, theta, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
    cudaStreamSynchronize(st);
    if (h->grad_tail != saved_tail) ntn_set_grad_tail(h, saved_tail);
    res->min_obj_val = f; res->iters = iters; res->evals = evals; res->did_converge = converged ? 1 : 0;
    res->n_active_last = nact;
    return rc;
}
