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
// Vector kernels: dots accumulate in FP64 with a fixed two-stage order (deterministic); scalars that feed the
// next vector op (alpha_i, beta) stay on the device so the two-loop recursion runs without host round trips.
#include <math.h>
#include <string.h>

#include "ntn_internal.h"

#define LB_BLOCKS 592            // 4 blocks per SM
#define LB_THREADS 256

__device__ __forceinline__ double block_sum(double v, double* sm) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) for (int w = 0; w < LB_THREADS / 32; ++w) t += sm[w];
    return t;
}

// part[b] = sum over the block's slice of x*y
__global__ void __launch_bounds__(LB_THREADS) k_dot_partial(const float* __restrict__ x, const float* __restrict__ y,
                                                            int64_t n4, double* __restrict__ part) {
    __shared__ double sm[LB_THREADS / 32];
    double acc = 0.0;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 a = __ldg(x4 + i), b = __ldg(y4 + i);
        acc += (double)(a.x * b.x + a.y * b.y) + (double)(a.z * b.z + a.w * b.w);
    }
    double t = block_sum(acc, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = t;
}
// out[0] = sum(part); optional post-ops so that dependent scalars never visit the host:
//   mode 1: out[0] = rho * dot            (alpha_i = rho_i * s_i.q)
//   mode 2: out[0] = aux - rho * dot      (alpha_i - beta, beta = rho_i * y_i.r)
//   mode 3: out[0] = 1/dot                (rho = 1 / y.s)
__global__ void k_dot_final(const double* __restrict__ part, int nb, double* out, int mode, const double* rho, const double* aux) {
    __shared__ double sm[256];
    double v = 0.0;
    for (int i = threadIdx.x; i < nb; i += 256) v += part[i];
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) {
        double d = sm[0];
        if (mode == 1) d = rho[0] * d;
        else if (mode == 2) d = aux[0] - rho[0] * d;
        else if (mode == 3) d = 1.0 / d;
        out[0] = d;
    }
}
// y = a*x + b*y with device scalars (coef = sign * (*sc)); x may be null (pure scale)
__global__ void __launch_bounds__(LB_THREADS) k_axpby(float* __restrict__ y, const float* __restrict__ x, int64_t n4,
                                                      const double* sc, double sign, double b) {
    const float a = sc ? (float)(sign * sc[0]) : (float)sign;
    const float bf = (float)b;
    float4* y4 = reinterpret_cast<float4*>(y);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 yv = y4[i];
        float4 xv = x ? __ldg(x4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        yv.x = fmaf(a, xv.x, bf * yv.x); yv.y = fmaf(a, xv.y, bf * yv.y);
        yv.z = fmaf(a, xv.z, bf * yv.z); yv.w = fmaf(a, xv.w, bf * yv.w);
        y4[i] = yv;
    }
}
// out = x + t*d  (host scalar t)
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
// s = a - b
__global__ void __launch_bounds__(LB_THREADS) k_diff(float* __restrict__ s, const float* __restrict__ a,
                                                     const float* __restrict__ b, int64_t n4) {
    float4* s4 = reinterpret_cast<float4*>(s);
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 x = __ldg(a4 + i), y = __ldg(b4 + i);
        s4[i] = make_float4(x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w);
    }
}

struct Lb {
    ntn_handle* h;
    int64_t n4;
    double *part, *sc;   // device: block partials; scalars [alpha 0..m) [rho 0..m) [tmp m*2 ..]
    cudaStream_t st;
    void dot(const float* x, const float* y, double* out, int mode = 0, const double* rho = nullptr, const double* aux = nullptr) {
        k_dot_partial<<<LB_BLOCKS, LB_THREADS, 0, st>>>(x, y, n4, part);
        k_dot_final<<<1, 256, 0, st>>>(part, LB_BLOCKS, out, mode, rho, aux);
    }
    double fetch(const double* d) {
        double v;
        cudaMemcpyAsync(&v, d, sizeof(double), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        return v;
    }
};

extern "C" int ntn_lbfgs_minimize(ntn_handle* h, float* d_theta, const ntn_lbfgs_opts* o, ntn_sides_fn sides_fn, void* ctx,
                                  ntn_lbfgs_result* res) {
    if (!h) return NTN_EINVAL;
    if (!d_theta || !o || !sides_fn || !res) { h->err = "null argument"; return NTN_EINVAL; }
    cudaSetDevice(h->device);
    const NtnDims& dm = h->dm;
    const int m = o->history > 0 ? o->history : 6;
    const int max_iters = o->max_iters > 0 ? o->max_iters : 5;
    const double tol = o->tol > 0 ? o->tol : 1e-4;
    const double c1 = 1e-4;
    const int64_t n = dm.Pint;                     // multiple of 4 by construction (oWVint % 4 == 0, dq % 4 == 0)
    if (n % 4) { h->err = "internal dimension not a multiple of 4"; return NTN_EUNSUP; }
    Lb lb{h, n / 4, nullptr, nullptr, h->stream};
    // workspace S[m] Y[m] g gnew dir trial + scalars, kept in the handle across calls (stable pointers also let the
    // evaluation graphs of (theta, g) and (trial, gnew) be reused while the batch job stays the same)
    const size_t nvec = (size_t)2 * m + 4;
    const int64_t ns = n + 4;                      // vector stride: room for the {cost, n_active} tail after a gradient
    const size_t need = nvec * ns * sizeof(float) + LB_BLOCKS * sizeof(double) + (2 * m + 8) * sizeof(double);
    if (h->lbfgs_ws_bytes < need) {
        cudaFree(h->lbfgs_ws); h->lbfgs_ws = nullptr; h->lbfgs_ws_bytes = 0;
        cudaError_t e = cudaMalloc(&h->lbfgs_ws, need);
        if (e != cudaSuccess) { h->err = std::string("lbfgs alloc: ") + cudaGetErrorString(e); return NTN_ECUDA; }
        h->lbfgs_ws_bytes = need;
    }
    float* buf = (float*)h->lbfgs_ws;
    lb.part = (double*)((char*)h->lbfgs_ws + nvec * ns * sizeof(float));
    lb.sc = lb.part + LB_BLOCKS;
    float* S = buf; float* Y = buf + (size_t)m * ns;
    float* g = buf + (size_t)2 * m * ns; float* gnew = g + ns; float* dir = gnew + ns; float* trial = dir + ns;
    double* alpha = lb.sc; double* rho = lb.sc + m; double* tmp = lb.sc + 2 * m;
    std::vector<uint8_t> side(dm.R);
    int rc = NTN_OK, evals = 0, iters = 0, hist = 0, head = 0;     // ring: newest pair at (head-1) mod m
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
        cudaMemcpyAsync(tail, gr + n, sizeof(tail), cudaMemcpyDeviceToHost, lb.st);
        if (cudaStreamSynchronize(lb.st) != cudaSuccess) { h->err = "lbfgs: stream error after the gradient exchange"; return (int)NTN_ECUDA; }
        *cost = (double)tail[0] + (double)tail[1];
        nact = (int64_t)llround((double)tail[2]) * 4096 + (int64_t)llround((double)tail[3]);
        return (int)NTN_OK;
    };
    if ((rc = evaluate(d_theta, g, &f))) goto done;
    res->first_cost = f;
    for (iters = 0; iters < max_iters; ++iters) {
        // ---- direction: two-loop recursion on the device
        cudaMemcpyAsync(dir, g, n * sizeof(float), cudaMemcpyDeviceToDevice, lb.st);
        for (int i = 0; i < hist; ++i) {                           // newest to oldest
            int j = (head - 1 - i + 2 * m) % m;
            lb.dot(S + (size_t)j * ns, dir, alpha + j, 1, rho + j);                                      // alpha_j = rho_j s_j.q
            k_axpby<<<LB_BLOCKS, LB_THREADS, 0, lb.st>>>(dir, Y + (size_t)j * ns, lb.n4, alpha + j, -1.0, 1.0);   // q -= alpha_j y_j
        }
        if (hist > 0) {
            int j = (head - 1 + m) % m;                            // H0 = (s.y / y.y) I = 1 / (rho_j * y_j.y_j)
            lb.dot(Y + (size_t)j * ns, Y + (size_t)j * ns, tmp, 1, rho + j);
            k_dot_final<<<1, 256, 0, lb.st>>>(tmp, 1, tmp + 1, 3, nullptr, nullptr);                    // tmp[1] = 1/(rho y.y)
            // r = gamma * q
            double gamma = lb.fetch(tmp + 1);
            k_axpby<<<LB_BLOCKS, LB_THREADS, 0, lb.st>>>(dir, nullptr, lb.n4, nullptr, 0.0, gamma);
        }
        for (int i = hist - 1; i >= 0; --i) {                      // oldest to newest
            int j = (head - 1 - i + 2 * m) % m;
            lb.dot(Y + (size_t)j * ns, dir, tmp, 2, rho + j, alpha + j);                                 // tmp = alpha_j - rho_j y_j.r
            k_axpby<<<LB_BLOCKS, LB_THREADS, 0, lb.st>>>(dir, S + (size_t)j * ns, lb.n4, tmp, 1.0, 1.0); // r += (alpha_j - beta) s_j
        }
        double gd;                                                 // directional derivative of -dir
        lb.dot(g, dir, tmp);
        gd = -lb.fetch(tmp);
        double t = 1.0, shrink = 0.5;
        if (hist == 0) {                                           // first iteration: unit-norm steepest descent
            double gn = sqrt(-gd);
            t = gn > 0 ? 1.0 / gn : 1.0;
            shrink = 0.1;
        }
        if (!(gd < 0)) {                                           // not a descent direction: restart from steepest descent
            cudaMemcpyAsync(dir, g, n * sizeof(float), cudaMemcpyDeviceToDevice, lb.st);
            lb.dot(g, g, tmp);
            gd = -lb.fetch(tmp);
            hist = 0;
            if (!(gd < 0)) { converged = true; break; }
            t = 1.0 / sqrt(-gd); shrink = 0.1;
        }
        // ---- backtracking line search along -dir
        bool ok = false;
        for (int ls = 0; ls < 20; ++ls) {
            k_step<<<LB_BLOCKS, LB_THREADS, 0, lb.st>>>(trial, d_theta, dir, lb.n4, (float)(-t));
            if ((rc = evaluate(trial, gnew, &fnew))) goto done;
            if (fnew <= f + c1 * t * gd) { ok = true; break; }
            t *= shrink;
        }
        if (!ok) break;                                            // line search failed: keep the current point
        // ---- curvature pair s = trial - theta, y = gnew - g
        {
            int j = head;
            k_diff<<<LB_BLOCKS, LB_THREADS, 0, lb.st>>>(S + (size_t)j * ns, trial, d_theta, lb.n4);
            k_diff<<<LB_BLOCKS, LB_THREADS, 0, lb.st>>>(Y + (size_t)j * ns, gnew, g, lb.n4);
            lb.dot(S + (size_t)j * ns, Y + (size_t)j * ns, rho + j, 3);                                   // rho_j = 1 / y.s
            double rj = lb.fetch(rho + j);
            if (rj > 0 && isfinite(rj)) { head = (head + 1) % m; if (hist < m) ++hist; }              // skip pairs with y.s <= 0
        }
        cudaMemcpyAsync(d_theta, trial, n * sizeof(float), cudaMemcpyDeviceToDevice, lb.st);
        cudaMemcpyAsync(g, gnew, n * sizeof(float), cudaMemcpyDeviceToDevice, lb.st);
        double fold = f;
        f = fnew;
        if (fabs(fold - f) <= tol * (fabs(fold) + fabs(f) + 1e-10) / 2.0) { converged = true; ++iters; break; }
    }
done:
    cudaStreamSynchronize(lb.st);
    if (h->grad_tail != saved_tail) ntn_set_grad_tail(h, saved_tail);
    res->min_obj_val = f; res->iters = iters; res->evals = evals; res->did_converge = converged ? 1 : 0;
    res->n_active_last = nact;
    return rc;
}
