// Host <-> staging type conversion for the host-buffer entry points (ntn_eval, ntn_upload_theta, ...): theta and the
// gradient are double[] on the reference side (NTN.java:92, Util.java:49-55) and float in HBM.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <thread>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

// torchrun exports OMP_NUM_THREADS=1 to every rank; the staging loops size their own team: the host cores divided
// among the ranks of this box (LOCAL_WORLD_SIZE), at most 16.
static int convert_threads() {
    static int nt = [] {
        unsigned hw = std::thread::hardware_concurrency();
        if (const char* e = getenv("NTN_CONVERT_THREADS")) return std::max(1, atoi(e));
        const char* lws = getenv("LOCAL_WORLD_SIZE");
        int ranks = lws ? std::max(1, atoi(lws)) : 1;
        return (int)std::max(1u, std::min(16u, (hw ? hw : 1u) / (unsigned)ranks));
    }();
    return nt;
}
// one thread's slice.  The arrays are far larger than the caches and each element is touched once, so the SSE2 variants
// store with non-temporal writes (no read-for-ownership of the destination lines); cvtpd2ps rounds to nearest-even like
// the (float) cast, i.e. like Util.java:49-55.  NTN_NT_STORES=0 selects the plain loops.
template <typename S, typename D> static void convert_slice(const S* src, D* dst, int64_t n) {
    for (int64_t i = 0; i < n; ++i) dst[i] = (D)src[i];
}
#if defined(__SSE2__)
static bool nt_stores() {
    static bool on = [] { const char* e = getenv("NTN_NT_STORES"); return !(e && e[0] == '0'); }();
    return on;
}
template <> inline void convert_slice<double, float>(const double* src, float* dst, int64_t n) {
    int64_t i = 0;
    if (nt_stores()) {
        for (; i < n && ((uintptr_t)(dst + i) & 15); ++i) dst[i] = (float)src[i];
        for (; i + 4 <= n; i += 4) {
            __m128 lo = _mm_cvtpd_ps(_mm_loadu_pd(src + i)), hi = _mm_cvtpd_ps(_mm_loadu_pd(src + i + 2));
            _mm_stream_ps(dst + i, _mm_movelh_ps(lo, hi));
        }
        _mm_sfence();
    }
    for (; i < n; ++i) dst[i] = (float)src[i];
}
template <> inline void convert_slice<float, double>(const float* src, double* dst, int64_t n) {
    int64_t i = 0;
    if (nt_stores() && ((uintptr_t)dst & 7) == 0) {
        for (; i < n && ((uintptr_t)(dst + i) & 15); ++i) dst[i] = (double)src[i];
        for (; i + 4 <= n; i += 4) {
            __m128 v = _mm_loadu_ps(src + i);
            _mm_stream_pd(dst + i, _mm_cvtps_pd(v));
            _mm_stream_pd(dst + i + 2, _mm_cvtps_pd(_mm_movehl_ps(v, v)));
        }
        _mm_sfence();
    }
    for (; i < n; ++i) dst[i] = (double)src[i];
}
#endif
template <typename S, typename D> static void convert_range(const S* src, D* dst, int64_t n) {
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(convert_threads(), n / 4096));
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int t = 0; t < nt; ++t) {
        int64_t a = n * t / nt, b = n * (t + 1) / nt;
        convert_slice(src + a, dst + a, b - a);
    }
}
