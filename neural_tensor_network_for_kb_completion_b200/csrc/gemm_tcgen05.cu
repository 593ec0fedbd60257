// tcgen05 / TMEM contractions of the NTN path with a 3xTF32 split (FP32-level accuracy on the
// 5th-generation tensor cores).  Three grouped GEMMs per evaluation:
//
//   fwd  T'[u, n]      = sum_p  [E_anchor(u) | 1][p] * Waug_r[p, n]        NTN.java:163-197
//   dw   dWaug_r[p, n] = sum_u  [E_anchor(u) | 1][p] * M[u, n]  (split-K)  NTN.java:301-349
//   ge   GA[u, p]      = sum_n  M[u, n] * Waug_r[p, n]                     NTN.java:352-388
//
// One CTA = 128-row accumulator tiles in TMEM (tcgen05.alloc, two buffers), N <= 160 columns, 416 threads.
//   warps 0-3, 5-8  two producer groups taking alternate k-blocks: gather / load FP32 operand rows, split x = hi + lo
//                   (both exactly representable in TF32), store both into swizzled UMMA shared-memory tiles,
//                   fence.proxy.async, arrive on the stage's "full" mbarrier.
//   warps 9-12      epilogue (tcgen05.ld -> registers -> staged in shared memory -> coalesced global stores).
//   warp 4          TMEM allocation + the single-thread MMA issuer: per k-step of 8 it issues tcgen05.mma kind::tf32
//                   lo*hi, hi*lo, hi*hi into the same FP32 accumulator, then tcgen05.commit releases the stage
//                   ("empty" mbarrier) / signals the epilogue.
// fwd and ge use K-major SWIZZLE_128B tiles (row = M/N index, 128 bytes of K per row).  dw contracts over the ROW
// index of both of its operands, so it uses MN-major tiles (SWIZZLE_128B_BASE32B, the only MN-major layout for 32-bit
// types); a K-major variant whose producers transpose while they store is kept as NTN_TC_DW=k.
// The weight operand Waug is written once per evaluation by k_tc_w_prep in both K-major orientations; in fwd and ge
// it is fed by TMA (cp.async.bulk.tensor 2-D boxes from pre-split hi/lo arrays, SWIZZLE_128B tensor maps, completion
// bytes on the stage's full mbarrier); NTN_TC_TMA=0 selects the software-loaded fallback.
#include <cuda.h>
#include <algorithm>
#include <stdio.h>
#include <stdlib.h>

#include "ntn_internal.h"

#define TC_STAGES 3
#define TC_BK 32                  // floats of K per stage (= one 128-byte swizzle row)
#define TC_NCH 160                // accumulator columns per CTA
#define TC_A_BYTES (128 * 128)    // one split of the A tile per stage
#define TC_B_BYTES (TC_NCH * 128) // one split of the B tile per stage
#define TC_STAGE_BYTES (2 * TC_A_BYTES + 2 * TC_B_BYTES)
#define TC_EPI_BYTES (4 * 32 * 16 * 4)   // epilogue staging: one 32 x 16 float slab per epilogue warp
#define TC_SMEM_BYTES (TC_STAGES * TC_STAGE_BYTES + TC_EPI_BYTES + 256 /*barriers*/)
#define TC_ACC_COLS 256            // TMEM columns per accumulator buffer (N <= 160)
#define TC_TMEM_COLS 512           // two accumulators: the epilogue of item i overlaps the main loop of item i+1
#define TC_SPIN_LIMIT (1u << 22)

struct TcState {
    float *WT = nullptr;                        // [R][Np][KpP]   Waug^T, K(=p)-major rows      (fwd B)
    float *W = nullptr;                         // [R][KpN][NpP]  Waug,   K(=n)-major rows      (ge  B)
    // TMA path: the same two arrays pre-split into TF32 hi/lo, and their tiled tensor maps (SWIZZLE_128B boxes)
    bool use_tma = false;
    bool dw_mn = false;                         // dw with MN-major (SWIZZLE_128B_BASE32B) operand tiles
    float *WT_hi = nullptr, *WT_lo = nullptr, *W_hi = nullptr, *W_lo = nullptr;
    CUtensorMap mWT_hi, mWT_lo, mW_hi, mW_lo;
    CUtensorMap mW2_hi, mW2_lo;                 // W again with boxes of <= 128 rows: the A operand of the swapped-role contraction
    int box_fwd = 0, box_ge = 0, box_ge2 = 0;   // rows per TMA box
    bool ge2 = true;                            // entity-gradient contraction with swapped operand roles (NTN_TC_GE2=0: off)
    int KpP = 0, KpN = 0, NpP = 0;
    int* t_anchor = nullptr;                    // [Nu] anchor entity per unique triple, rewritten per evaluation (grow-only)
    size_t anchor_cap = 0;
    // TMA-fed A operands (fwd: the anchor rows of E gathered into unique-triple order and pre-split by k_tc_anchor_rows;
    // ge: M as the score kernel writes it, pre-split): no producer warp touches the data
    bool tma_a = false;
    float *EA_hi = nullptr, *EA_lo = nullptr;   // [Nu][Kp]
    size_t ea_cap = 0;
    CUtensorMap mEA_hi, mEA_lo, mM_hi, mM_lo;
    int* err = nullptr;                         // device flag: a bounded wait expired
    long long* dbg = nullptr;                   // 3 x 256 cycle stamps of CTA (0,0,0) per mode (debug timeline)
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug must not hang the GPU
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* err) {
    for (uint32_t i = 0; i < TC_SPIN_LIMIT; ++i)
        if (mbar_try_wait(bar, parity)) return true;
    atomicExch(err, 1);
    return false;
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA: one 2-D tile (inner coordinate x in elements, row y) global -> shared, completion on an mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], TF32 inputs, FP32 accumulate, single-CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): SWIZZLE_128B, version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): F32 accumulate, TF32 x TF32, M=128
__host__ __device__ constexpr uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

// (tf32_round / split_tf32: ntn_internal.h)
__device__ __forceinline__ float4 ldg4g(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void sts4(uint8_t* base, uint32_t off, const float4& v) { *reinterpret_cast<float4*>(base + off) = v; }

// ------------------------------------------------------------------------------------------
struct TcArgs {
    const int *t_rel, *t_u0, *t_n;     // tile table (fwd/ge: <=128 rows) or split-K chunk table (dw)
    const int *u_e1, *u_e2;
    const int* t_anchor;               // anchor entity of every unique triple for this evaluation's corrupt sides (k_tc_w_prep)
    const uint8_t* side;
    const float *E, *M;
    const float* M_lo;                 // non-null: M holds the TF32 hi parts and M_lo the lo parts (written split by the score kernel)
    const float *WT, *W;
    float *T, *GA, *dWpart;
    int Kp, Np, KpP, KpN, NpP;
    int* err;
    long long* dbg;                    // cycle-stamp timeline of CTA (0,0,0), or null
    int b_rows_per_rel;                // TMA path: rows of the weight tensor per relation (Np resp. KpN)
    int b_box_rows;                    // TMA path: rows per box (expect_tx bytes = 2 * rows * 128)
    int nitems, ny, nz;                // work items = nitems x ny (N chunks) x nz (M chunks, dw)
};

enum { TC_FWD = 0, TC_DW = 1, TC_GE = 2, TC_GE2 = 3 };   // GE2: the entity-gradient contraction with swapped operand roles

#define TC_THREADS 416            // warps 0-3, 5-8: producer groups, warp 4: MMA issuer, warps 9-12: epilogue

__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, float v[16]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// DW_MN (dw only): both operands as MN-major tiles.  dw contracts over the ROW index u of E_anchor and M, i.e. its K
// index is the slow dimension of both arrays; an MN-major UMMA operand takes the rows as they are (coalesced loads,
// 16-byte stores).  For 32-bit types the only MN-major layout is SWIZZLE_128B_BASE32B: atoms of 4 K-rows x 128 bytes,
// 32-byte chunks XOR-ed with (row & 3); tile = [32-float MN atom][32 rows][128 B] -> LBO = 4096, SBO = 512, a k-step
// (8 rows) advances the start address by 1024.
template <int MODE, bool TMA_B, bool DW_MN = false, bool TMA_A = false>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_gemm(TcArgs a, const __grid_constant__ CUtensorMap mapB_hi,
                                                           const __grid_constant__ CUtensorMap mapB_lo,
                                                           const __grid_constant__ CUtensorMap mapA_hi,
                                                           const __grid_constant__ CUtensorMap mapA_lo) {
    // SWIZZLE_128B tiles need 1024-byte alignment.  The array is used directly (no pointer arithmetic through
    // integers) so that the stores below stay in the shared state space (STS): with generic stores the compiler
    // cannot hoist the global loads above them and every load round trip is exposed.
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* epi_stage = smem + TC_STAGES * TC_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES + TC_EPI_BYTES);
    // bars[0..S) full, bars[S..2S) empty, then accumulator full[2], accumulator empty[2]; then the TMEM base address
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 4);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (TC_STAGES + s); };
    auto acc_full = [&](int b) { return bar0 + 8u * (2 * TC_STAGES + b); };
    auto acc_empty = [&](int b) { return bar0 + 8u * (2 * TC_STAGES + 2 + b); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // debug timeline: thread 0 (producer group 0) stamps slots [0,128), thread 128 (MMA issuer) slots [128,256)
    long long* dbg = (a.dbg && blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 128))
                         ? a.dbg + (MODE == TC_GE2 ? TC_GE : MODE) * 256 + (threadIdx.x == 128 ? 128 : 0) : nullptr;
    int dbg_n = 0;
    auto stamp = [&]() { if (dbg && dbg_n < 128) dbg[dbg_n++] = clock64(); };
    stamp();

    if (threadIdx.x == 0) {
        // full: 128 producer arrivals (+ the TMA issuer's arrive.expect_tx when the weight tile comes by TMA)
        // (TMA_A: both operands come by TMA, the only arrival is the issuing thread's arrive.expect_tx)
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full_bar(s), TMA_A ? 1 : ((TMA_B || MODE == TC_GE2) ? 129 : 128)); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 128); }
        fence_barrier_init();
        if (TMA_B || MODE == TC_GE2) { tma_prefetch_desc(&mapB_hi); tma_prefetch_desc(&mapB_lo); }
        if (TMA_A) { tma_prefetch_desc(&mapA_hi); tma_prefetch_desc(&mapA_lo); }
    }
    if (warp == 4) tmem_alloc(smem_u32(tmem_slot), TC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    stamp();

    // Work items: (tile or chunk) x N-chunk x M-chunk, taken round-robin by the CTAs of a persistent grid.  Every role
    // walks the same sequence and derives the same geometry; the pipeline (stage ring, the two TMEM accumulators)
    // runs across item boundaries: while the epilogue warps drain accumulator b, the producers and the MMA issuer
    // are already on the next item in accumulator b^1.
    const int nyz = a.ny * a.nz, nwork = a.nitems * nyz;
#define TC_ITEM_GEOMETRY(w)                                                                              \
    const int item = (w) / nyz, yz_ = (w) - item * nyz, by_ = yz_ % a.ny, bz_ = yz_ / a.ny;              \
    const int r = a.t_rel[item], u0 = a.t_u0[item], rows = a.t_n[item];                                  \
    int n0, nlen, m0 = 0, kext;                                                                          \
    if (MODE == TC_GE) { n0 = by_ * TC_NCH; nlen = min(TC_NCH, a.KpN - n0); kext = a.Np; }               \
    else { n0 = by_ * TC_NCH; nlen = min(TC_NCH, a.Np - n0); kext = (MODE == TC_FWD) ? a.Kp : rows; }    \
    if (MODE == TC_DW) m0 = bz_ * 128;                                                                   \
    const int nkb = (kext + TC_BK - 1) / TC_BK;                                                          \
    (void)bz_; (void)r; (void)m0; (void)u0; (void)n0; (void)nlen;


    if (MODE == TC_GE2) {
        // ================================ entity-gradient contraction, operand roles swapped ================================
        //   GA^T[p][u] = sum_n Waug_r[p][n] * M[u][n]        (NTN.java:352-388)
        // A (the 128 TMEM lanes) = a 128-row chunk of Waug_r -- the pre-split K-major copy k_tc_w_tiles writes, by TMA --
        // and B (the N dimension) = the M rows of up to TC_NCH unique triples, loaded and split by the producers.  The
        // issue cost of a kind::tf32 MMA does not depend on N (~125-139 cycles for N = 104 ... 256, measured), so the
        // orientation with Kp = 104 as N runs the tensor pipe at 104/256 of its rate; here N is the tile height
        // (160: 130 tiles at 20,000 unique triples -- ONE round of the persistent grid instead of two).
        // Work items: (tile of the ge table) x (128-row chunk of p); accumulator = [128 lanes p] x [N columns u].
        const int nz2 = a.nz, nwork2 = a.nitems * nz2;
        const int nkb2 = (a.Np + TC_BK - 1) / TC_BK;
        if (warp < 9 && warp != 4) {
            const int group = warp > 4 ? 1 : 0;
            const int t = threadIdx.x - (group ? 160 : 0);
            const int j = t & 7, rsub = t >> 3;
            bool alive = true;
            int g = 0;
            for (int w = blockIdx.x; w < nwork2 && alive; w += gridDim.x) {
                const int item = w / nz2, bz = w - item * nz2;
                const int r = __ldg(a.t_rel + item), u0 = __ldg(a.t_u0 + item), rows = __ldg(a.t_n + item);
                const int nlen = (rows + 15) & ~15;
                for (int kb = 0; kb < nkb2 && alive; ++kb, ++g) {
                    if ((g & 1) != group) continue;
                    const int s = g % TC_STAGES;
                    const uint32_t ph = (g / TC_STAGES) & 1;
                    alive = mbar_wait(empty_bar(s), ph ^ 1, a.err);
                    stamp();
                    uint8_t* sA_hi = smem + s * TC_STAGE_BYTES;
                    uint8_t* sA_lo = sA_hi + TC_A_BYTES;
                    uint8_t* sB_hi = sA_lo + TC_A_BYTES;
                    uint8_t* sB_lo = sB_hi + TC_B_BYTES;
                    const int k0 = kb * TC_BK;
                    if (t == 0) {
                        // weight chunk: two boxes (hi, lo) of 32 floats x b_box_rows rows of Waug_r (rows past the chunk's
                        // valid p belong to the next relation or arrive as zeros: lanes nobody reads)
                        mbar_arrive_expect_tx(full_bar(s), 2u * (uint32_t)a.b_box_rows * 128u);
                        const int y = r * a.b_rows_per_rel + bz * 128;
                        tma_load_2d(smem_u32(sA_hi), &mapB_hi, k0, y, full_bar(s));
                        tma_load_2d(smem_u32(sA_lo), &mapB_lo, k0, y, full_bar(s));
                    }
                    constexpr int BR = TC_NCH / 16;                  // M rows per thread
                    float4 xb[BR];
                    const int kk = k0 + j * 4;
#pragma unroll
                    for (int i = 0; i < BR; ++i) {
                        const int row = rsub + 16 * i;
                        xb[i] = (row < rows && kk < a.Np) ? ldg4g(a.M + (size_t)(u0 + row) * a.Np + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int i = 0; i < BR; ++i) {
                        const int row = rsub + 16 * i;
                        if (row < nlen) {
                            float4 hi, lo;
                            split_tf32(xb[i], hi, lo);
                            const uint32_t off = row * 128 + ((j ^ (row & 7)) << 4);
                            sts4(sB_hi, off, hi); sts4(sB_lo, off, lo);
                        }
                    }
                    fence_proxy_async();
                    mbar_arrive(full_bar(s));
                    stamp();
                }
            }
        } else if (warp == 4) {
            bool alive = true;
            int g = 0, tl = 0;
            for (int w = blockIdx.x; w < nwork2 && alive; w += gridDim.x, ++tl) {
                const int item = w / nz2;
                const int rows = a.t_n[item];
                const int nlen = (rows + 15) & ~15;
                const uint32_t idesc = make_idesc(nlen, 0, 0);
                const int b = tl & 1;
                const uint32_t tacc = tmem_base + (uint32_t)(b * TC_ACC_COLS);
                alive = mbar_wait(acc_empty(b), (uint32_t)((tl >> 1) & 1) ^ 1u, a.err);
                tc_fence_after();
                for (int kb = 0; kb < nkb2 && alive; ++kb, ++g) {
                    const int s = g % TC_STAGES;
                    const uint32_t ph = (g / TC_STAGES) & 1;
                    alive = mbar_wait(full_bar(s), ph, a.err);
                    tc_fence_after();
                    stamp();
                    if (lane == 0 && alive) {
                        const uint32_t sA_hi = smem_u32(smem + s * TC_STAGE_BYTES);
                        const uint32_t sA_lo = sA_hi + TC_A_BYTES, sB_hi = sA_lo + TC_A_BYTES, sB_lo = sB_hi + TC_B_BYTES;
                        const int kleft = a.Np - kb * TC_BK;
                        const int nks = (kleft >= TC_BK) ? 4 : (kleft + 7) / 8;
                        for (int ks = 0; ks < nks; ++ks) {
                            const uint64_t dAh = make_desc(sA_hi + ks * 32, 16, 1024), dAl = make_desc(sA_lo + ks * 32, 16, 1024);
                            const uint64_t dBh = make_desc(sB_hi + ks * 32, 16, 1024), dBl = make_desc(sB_lo + ks * 32, 16, 1024);
                            const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
                            umma_tf32(tacc, dAl, dBh, idesc, first);
                            umma_tf32(tacc, dAh, dBl, idesc, 1u);
                            umma_tf32(tacc, dAh, dBh, idesc, 1u);
                        }
                        umma_commit(empty_bar(s));
                        if (kb == nkb2 - 1) umma_commit(acc_full(b));
                    }
                    stamp();
                    __syncwarp();
                }
            }
            tc_fence_before();
        } else {
            // epilogue: TMEM lane = row p of the chunk, columns = the tile's unique triples; GA is [u][p], so for a fixed
            // column the 32 lanes of a warp store 32 consecutive floats
            const int q = warp & 3;
            const int row = q * 32 + lane;
            bool alive = true;
            int tl = 0;
            for (int w = blockIdx.x; w < nwork2 && alive; w += gridDim.x, ++tl) {
                const int item = w / nz2, bz = w - item * nz2;
                const int u0 = a.t_u0[item], rows = a.t_n[item];
                const int nlen = (rows + 15) & ~15;
                const int b = tl & 1;
                alive = mbar_wait(acc_full(b), (uint32_t)((tl >> 1) & 1), a.err);
                tc_fence_after();
                const uint32_t trow = tmem_base + (uint32_t)(b * TC_ACC_COLS) + ((uint32_t)(q * 32) << 16);
                const int p = bz * 128 + row;
                for (int c0 = 0; c0 < nlen && alive; c0 += 32) {
                    float v[32];
                    tmem_ld16_issue(trow + c0, v);
                    if (c0 + 16 < nlen) tmem_ld16_issue(trow + c0 + 16, v + 16);
                    tmem_ld_wait();
                    if (p < a.Kp) {
                        float* dst = a.GA + (size_t)(u0 + c0) * a.Kp + p;
#pragma unroll
                        for (int qq = 0; qq < 32; ++qq)
                            if (c0 + qq < rows) dst[(size_t)qq * a.Kp] = v[qq];
                    }
                }
                tc_fence_before();
                mbar_arrive(acc_empty(b));
                stamp();
            }
        }
    } else
    if (TMA_A && warp < 9 && warp != 4) {
        // ================================ producers, all-TMA ================================
        // One thread per group (the two groups still take alternate k-blocks, so two stages are requested while the
        // issuer works on a third): wait for the stage, announce its bytes, request the four boxes -- A hi/lo (128
        // rows of the pre-split operand, rows past the array end arrive as zeros, rows of the next tile are computed
        // and never stored) and the weight tile hi/lo.  Nothing is loaded, split or stored by a thread.
        const int group = warp > 4 ? 1 : 0;
        if (threadIdx.x == (group ? 160 : 0)) {
            bool alive = true;
            int g = 0;
            for (int w = blockIdx.x; w < nwork && alive; w += gridDim.x) {
                TC_ITEM_GEOMETRY(w)
                (void)rows; (void)kext;
                for (int kb = 0; kb < nkb && alive; ++kb, ++g) {
                    if ((g & 1) != group) continue;
                    const int s = g % TC_STAGES;
                    const uint32_t ph = (g / TC_STAGES) & 1;
                    alive = mbar_wait(empty_bar(s), ph ^ 1, a.err);
                    if (!alive) break;
                    stamp();
                    const uint32_t sA_hi = smem_u32(smem + s * TC_STAGE_BYTES);
                    const uint32_t sA_lo = sA_hi + TC_A_BYTES, sB_hi = sA_lo + TC_A_BYTES, sB_lo = sB_hi + TC_B_BYTES;
                    const int k0 = kb * TC_BK;
                    mbar_arrive_expect_tx(full_bar(s), 2u * 128u * 128u + 2u * (uint32_t)a.b_box_rows * 128u);
                    tma_load_2d(sA_hi, &mapA_hi, k0, u0, full_bar(s));
                    tma_load_2d(sA_lo, &mapA_lo, k0, u0, full_bar(s));
                    const int y = r * a.b_rows_per_rel + n0;
                    tma_load_2d(sB_hi, &mapB_hi, k0, y, full_bar(s));
                    tma_load_2d(sB_lo, &mapB_lo, k0, y, full_bar(s));
                    stamp();
                }
            }
        }
    } else if (warp < 9 && warp != 4) {
        // ================================ producers ================================
        // two groups of 128 threads take alternate k-blocks (of the CTA's whole item sequence), so two k-blocks'
        // global loads are in flight
        const int group = warp > 4 ? 1 : 0;
        const int t = threadIdx.x - (group ? 160 : 0);   // 0..127 inside the group
        const int gw = t >> 5;                           // warp inside the group
        const int j = t & 7;                             // 16-byte chunk inside a 128-byte swizzle row
        const int rsub = t >> 3;                         // 0..15
        (void)gw;
        bool alive = true;
        int g = 0;                                       // k-blocks issued so far by this CTA (all items)
        // The tables of the NEXT item (and, fwd, the anchor entities of this thread's 8 rows) are fetched while the
        // current item is produced, so that an item starts with a single round trip (its operand rows) instead of
        // the chain  tile table -> anchor index -> rows.
        int w = blockIdx.x;
        int nx_r = 0, nx_u0 = 0, nx_rows = 0;
        int nx_anchor[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) nx_anchor[i] = -1;
        auto fetch_tables = [&](int ww) {
            if (ww < nwork) { const int it = ww / nyz; nx_r = __ldg(a.t_rel + it); nx_u0 = __ldg(a.t_u0 + it); nx_rows = __ldg(a.t_n + it); }
        };
        auto fetch_anchors = [&](int ww) {
            if (MODE == TC_FWD && ww < nwork) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = rsub + 16 * i;
                    nx_anchor[i] = (row < nx_rows) ? __ldg(a.t_anchor + nx_u0 + row) : -1;
                }
            }
        };
        fetch_tables(w);
        fetch_anchors(w);
        for (; w < nwork && alive; w += gridDim.x) {
            const int item = w / nyz, yz_ = w - item * nyz, by_ = yz_ % a.ny, bz_ = yz_ / a.ny;
            const int r = nx_r, u0 = nx_u0, rows = nx_rows;
            int anchor[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) anchor[i] = nx_anchor[i];
            fetch_tables(w + (int)gridDim.x);            // lands during this item
            bool anchors_fetched = false;
            int n0, nlen, m0 = 0, kext;
            if (MODE == TC_GE) { n0 = by_ * TC_NCH; nlen = min(TC_NCH, a.KpN - n0); kext = a.Np; }
            else { n0 = by_ * TC_NCH; nlen = min(TC_NCH, a.Np - n0); kext = (MODE == TC_FWD) ? a.Kp : rows; }
            if (MODE == TC_DW) m0 = bz_ * 128;
            const int nkb = (kext + TC_BK - 1) / TC_BK;
            (void)bz_; (void)m0; (void)item; (void)anchor;
            const float* Bw = (MODE == TC_FWD) ? a.WT + (size_t)r * a.Np * a.KpP
                             : (MODE == TC_GE) ? a.W + (size_t)r * a.KpN * a.NpP : nullptr;
            const int ldb = (MODE == TC_FWD) ? a.KpP : a.NpP;
            (void)Bw; (void)ldb;
            for (int kb = 0; kb < nkb && alive; ++kb, ++g) {
                if ((g & 1) != group) continue;
                const int s = g % TC_STAGES;
                const uint32_t ph = (g / TC_STAGES) & 1;
                alive = mbar_wait(empty_bar(s), ph ^ 1, a.err);
                stamp();
                uint8_t* sA_hi = smem + s * TC_STAGE_BYTES;
                uint8_t* sA_lo = sA_hi + TC_A_BYTES;
                uint8_t* sB_hi = sA_lo + TC_A_BYTES;
                uint8_t* sB_lo = sB_hi + TC_B_BYTES;
                const int k0 = kb * TC_BK;
                constexpr int BROWS = TC_NCH / 16;                                       // B rows per thread (K-major loads)
                if (TMA_B && t == 0) {
                    // weight tile by TMA: two boxes (hi, lo) of 32 floats x b_box_rows rows land directly in the swizzled
                    // UMMA layout; their bytes complete on the stage's full barrier
                    mbar_arrive_expect_tx(full_bar(s), 2u * (uint32_t)a.b_box_rows * 128u);
                    const int y = r * a.b_rows_per_rel + n0;
                    tma_load_2d(smem_u32(sB_hi), &mapB_hi, k0, y, full_bar(s));
                    tma_load_2d(smem_u32(sB_lo), &mapB_lo, k0, y, full_bar(s));
                }
                if (MODE == TC_FWD || MODE == TC_GE) {
                    // all global loads of this k-block are issued first (latency overlaps), then split + store
                    float4 xa[8], xb[BROWS];
                    const int kk = k0 + j * 4;
    #pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = rsub + 16 * i;
                        xa[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (MODE == TC_FWD) { if (anchor[i] >= 0 && kk < a.Kp) xa[i] = ldg4g(a.E + (size_t)anchor[i] * a.Kp + kk); }
                        else { if (row < rows && kk < a.Np) xa[i] = ldg4g(a.M + (size_t)(u0 + row) * a.Np + kk); }
                    }
                    if (!TMA_B) {
    #pragma unroll
                        for (int i = 0; i < BROWS; ++i) {
                            const int row = rsub + 16 * i;
                            xb[i] = (row < nlen) ? ldg4g(Bw + (size_t)(n0 + row) * ldb + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    // ---- A: K-major, 128 rows x 128 bytes, chunk j of row `row` at row*128 + ((j ^ (row&7))<<4)
    #pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = rsub + 16 * i;
                        float4 hi, lo;
                        split_tf32(xa[i], hi, lo);
                        const uint32_t off = row * 128 + ((j ^ (row & 7)) << 4);
                        sts4(sA_hi, off, hi); sts4(sA_lo, off, lo);
                    }
                    // ---- B: K-major, nlen rows (n resp. p) x 128 bytes of Waug, split here as well: reading one FP32
                    //      copy instead of pre-split hi/lo arrays halves the L2 traffic of the operand every CTA re-reads
                    if (!TMA_B) {
    #pragma unroll
                        for (int i = 0; i < BROWS; ++i) {
                            const int row = rsub + 16 * i;
                            if (row < nlen) {
                                float4 hi, lo;
                                split_tf32(xb[i], hi, lo);
                                const uint32_t off = row * 128 + ((j ^ (row & 7)) << 4);
                                sts4(sB_hi, off, hi); sts4(sB_lo, off, lo);
                            }
                        }
                    }
                } else if (DW_MN) {
                    // ---- dw, MN-major tiles: rows as they are.  Thread (rsub, j): rows ul = rsub, rsub+16 of the k-block,
                    //      16-byte chunk j of every 32-float atom.
                    float4 xa[2][4], xb[2][TC_NCH / 32], xl[2][TC_NCH / 32];
                    const bool msplit = a.M_lo != nullptr;
    #pragma unroll
                    for (int i2 = 0; i2 < 2; ++i2) {
                        const int u = k0 + rsub + 16 * i2;
                        const int an = (u < rows) ? __ldg(a.t_anchor + u0 + u) : -1;
    #pragma unroll
                        for (int mi = 0; mi < 4; ++mi) {
                            const int p = m0 + mi * 32 + j * 4;
                            xa[i2][mi] = (an >= 0 && p < a.Kp) ? ldg4g(a.E + (size_t)an * a.Kp + p) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
    #pragma unroll
                        for (int ni = 0; ni < TC_NCH / 32; ++ni) {
                            const int n = ni * 32 + j * 4;
                            xb[i2][ni] = (an >= 0 && n < nlen) ? ldg4g(a.M + (size_t)(u0 + u) * a.Np + n0 + n) : make_float4(0.f, 0.f, 0.f, 0.f);
                            xl[i2][ni] = (msplit && an >= 0 && n < nlen) ? ldg4g(a.M_lo + (size_t)(u0 + u) * a.Np + n0 + n) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
    #pragma unroll
                    for (int i2 = 0; i2 < 2; ++i2) {
                        const int ul = rsub + 16 * i2;
                        const uint32_t rowoff = ul * 128 + ((((uint32_t)j >> 1) ^ (ul & 3)) << 5) + (j & 1) * 16;
    #pragma unroll
                        for (int mi = 0; mi < 4; ++mi) {
                            float4 hi, lo;
                            split_tf32(xa[i2][mi], hi, lo);
                            sts4(sA_hi, mi * 4096 + rowoff, hi); sts4(sA_lo, mi * 4096 + rowoff, lo);
                        }
    #pragma unroll
                        for (int ni = 0; ni < TC_NCH / 32; ++ni) {
                            if (ni * 32 < nlen) {
                                float4 hi = xb[i2][ni], lo = xl[i2][ni];          // already split by the score kernel ...
                                if (!msplit) split_tf32(xb[i2][ni], hi, lo);      // ... or split here
                                sts4(sB_hi, ni * 4096 + rowoff, hi); sts4(sB_lo, ni * 4096 + rowoff, lo);
                            }
                        }
                    }
                } else {
                    // ---- dw (K-major fallback): the contraction runs over the ROW index u of both operands, so the producers
                    //      transpose while storing: lane = u row of this k-block (K index), warps stride over the
                    //      16-byte column chunks.  For a fixed chunk and component the 32 lanes hit 32 different
                    //      banks (8 swizzled k-chunks x 4 words), so the 4-byte stores are conflict-free.
                    constexpr int BCH = TC_NCH / 16;                                      // B column chunks per thread
                    const int ul = lane, u = k0 + ul;
                    const int an = (u < rows) ? __ldg(a.t_anchor + u0 + u) : -1;
                    const uint32_t kc = ul >> 2, ke = (ul & 3) * 4;
                    float4 xa[8], xb[BCH];
    #pragma unroll
                    for (int i = 0; i < 8; ++i) {                                         // A rows: p = m0 + 4c + e
                        const int p = m0 + 4 * (gw + 4 * i);
                        xa[i] = (an >= 0 && p < a.Kp) ? ldg4g(a.E + (size_t)an * a.Kp + p) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
    #pragma unroll
                    for (int i = 0; i < BCH; ++i) {                                       // B rows: n = n0 + 4c + e
                        const int c = gw + 4 * i;
                        xb[i] = (an >= 0 && 4 * c < nlen) ? ldg4g(a.M + (size_t)(u0 + u) * a.Np + n0 + 4 * c)
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
    #pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float4 hi, lo;
                        split_tf32(xa[i], hi, lo);
                        const float hv[4] = {hi.x, hi.y, hi.z, hi.w}, lv[4] = {lo.x, lo.y, lo.z, lo.w};
    #pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const uint32_t row = 4 * (gw + 4 * i) + e, off = row * 128 + ((kc ^ (row & 7)) << 4) + ke;
                            *reinterpret_cast<float*>(sA_hi + off) = hv[e];
                            *reinterpret_cast<float*>(sA_lo + off) = lv[e];
                        }
                    }
    #pragma unroll
                    for (int i = 0; i < BCH; ++i) {
                        const int c = gw + 4 * i;
                        if (4 * c < nlen) {
                            float4 hi, lo;
                            split_tf32(xb[i], hi, lo);
                            const float hv[4] = {hi.x, hi.y, hi.z, hi.w}, lv[4] = {lo.x, lo.y, lo.z, lo.w};
    #pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const uint32_t row = 4 * c + e, off = row * 128 + ((kc ^ (row & 7)) << 4) + ke;
                                *reinterpret_cast<float*>(sB_hi + off) = hv[e];
                                *reinterpret_cast<float*>(sB_lo + off) = lv[e];
                            }
                        }
                    }
                }
                fence_proxy_async();                       // generic-proxy stores -> visible to the tensor core
                mbar_arrive(full_bar(s));
                stamp();
                if (!anchors_fetched) { fetch_anchors(w + (int)gridDim.x); anchors_fetched = true; }
            }
            if (!anchors_fetched) fetch_anchors(w + (int)gridDim.x);
        }
    } else if (warp == 4) {
        // ================================ MMA issuer ================================
        bool alive = true;
        int g = 0, tl = 0;
        for (int w = blockIdx.x; w < nwork && alive; w += gridDim.x, ++tl) {
            TC_ITEM_GEOMETRY(w)
            const uint32_t idesc = make_idesc(nlen, DW_MN ? 1 : 0, DW_MN ? 1 : 0);
            const int b = tl & 1;
            const uint32_t tacc = tmem_base + (uint32_t)(b * TC_ACC_COLS);
            alive = mbar_wait(acc_empty(b), (uint32_t)((tl >> 1) & 1) ^ 1u, a.err);   // epilogue has drained this accumulator
            tc_fence_after();
            for (int kb = 0; kb < nkb && alive; ++kb, ++g) {
                const int s = g % TC_STAGES;
                const uint32_t ph = (g / TC_STAGES) & 1;
                alive = mbar_wait(full_bar(s), ph, a.err);
                tc_fence_after();
                stamp();
                if (lane == 0 && alive) {
                    const uint32_t sA_hi = smem_u32(smem + s * TC_STAGE_BYTES);
                    const uint32_t sA_lo = sA_hi + TC_A_BYTES, sB_hi = sA_lo + TC_A_BYTES, sB_lo = sB_hi + TC_B_BYTES;
                    const int kleft = kext - kb * TC_BK;
                    const int nks = (kleft >= TC_BK) ? 4 : (kleft + 7) / 8;
                    for (int ks = 0; ks < nks; ++ks) {
                        uint64_t dAh, dAl, dBh, dBl;
                        if (DW_MN) {     // MN-major SWIZZLE_128B_BASE32B: k-step = 8 rows = 1024 B; MN atoms 4096 B apart; 4-row K groups 512 B apart
                            dAh = make_desc(sA_hi + ks * 1024, 4096, 512, 1); dAl = make_desc(sA_lo + ks * 1024, 4096, 512, 1);
                            dBh = make_desc(sB_hi + ks * 1024, 4096, 512, 1); dBl = make_desc(sB_lo + ks * 1024, 4096, 512, 1);
                        } else {         // K-major: a k-step is 32 bytes inside the 128-byte swizzle row; 8-row groups are 1024 B apart
                            dAh = make_desc(sA_hi + ks * 32, 16, 1024); dAl = make_desc(sA_lo + ks * 32, 16, 1024);
                            dBh = make_desc(sB_hi + ks * 32, 16, 1024); dBl = make_desc(sB_lo + ks * 32, 16, 1024);
                        }
                        const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
                        umma_tf32(tacc, dAl, dBh, idesc, first);           // small terms first
                        umma_tf32(tacc, dAh, dBl, idesc, 1u);
                        umma_tf32(tacc, dAh, dBh, idesc, 1u);
                    }
                    umma_commit(empty_bar(s));                              // stage free once these MMAs have read it
                    if (kb == nkb - 1) umma_commit(acc_full(b));            // accumulator complete
                }
                stamp();
                __syncwarp();
            }
        }
        tc_fence_before();
    } else {
        // ================================ epilogue (warps 9-12; warp w owns TMEM lanes 32*(w%4) .. +31) ================
        const int q = warp & 3;
        const int row = q * 32 + lane;                       // TMEM lane = accumulator row
        float* stg = reinterpret_cast<float*>(epi_stage + (warp - 9) * (TC_EPI_BYTES / 4));   // 32 rows x 16 floats
        bool alive = true;
        int tl = 0;
        for (int w = blockIdx.x; w < nwork && alive; w += gridDim.x, ++tl) {
            TC_ITEM_GEOMETRY(w)
            const int b = tl & 1;
            alive = mbar_wait(acc_full(b), (uint32_t)((tl >> 1) & 1), a.err);
            tc_fence_after();
            const uint32_t trow = tmem_base + (uint32_t)(b * TC_ACC_COLS) + ((uint32_t)(q * 32) << 16);
            if (MODE == TC_DW) {
                // partial stored transposed, [n][p]: lanes = consecutive p -> coalesced 128-byte stores
                const int p = m0 + row;
                for (int c0 = 0; c0 < nlen && alive; c0 += 32) {
                    float v[32];
                    tmem_ld16_issue(trow + c0, v);
                    if (c0 + 16 < nlen) tmem_ld16_issue(trow + c0 + 16, v + 16);
                    tmem_ld_wait();
                    if (p < a.Kp) {
                        float* dst = a.dWpart + (size_t)item * a.Kp * a.Np + (size_t)(n0 + c0) * a.Kp + p;
#pragma unroll
                        for (int qq = 0; qq < 32; ++qq)
                            if (c0 + qq < nlen) dst[(size_t)qq * a.Kp] = v[qq];
                    }
                }
            } else {
                // 16-column slabs through the warp's private staging tile (XOR-swizzled 16-byte chunks, conflict-free
                // per quarter warp), written out as 64-byte row segments: 8 rows x 2 full sectors per store instruction
                float* out = (MODE == TC_FWD) ? a.T : a.GA;
                const int ldo = (MODE == TC_FWD) ? a.Np : a.Kp;
                const int rr0 = lane >> 2, cc = lane & 3;
                for (int c0 = 0; c0 < nlen; c0 += 16) {
                    float v[16];
                    tmem_ld16_issue(trow + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        *reinterpret_cast<float4*>(stg + lane * 16 + 4 * (jj ^ ((lane >> 1) & 3))) =
                            make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr = rr0 + 8 * i;
                        const float4 x = *reinterpret_cast<const float4*>(stg + rr * 16 + 4 * (cc ^ ((rr >> 1) & 3)));
                        const int col = n0 + c0 + 4 * cc;
                        if (q * 32 + rr < rows && col < ldo)
                            *reinterpret_cast<float4*>(out + (size_t)(u0 + q * 32 + rr) * ldo + col) = x;
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            mbar_arrive(acc_empty(b));                       // 128 arrivals: the accumulator may be overwritten
            stamp();
        }
    }
#undef TC_ITEM_GEOMETRY
    tc_fence_before();
    __syncthreads();
    if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, TC_TMEM_COLS); }
}

// ------------------------------------------------------------------------------------------
// per-evaluation weight preparation: Waug_r (see k_w_prep in ntn_kernels.cu) in the two K-major
// layouts the kernels above read.  Padding rows/columns are zero.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float waug_at(const float* theta, const NtnDims& dm, int r, bool tail, int p, int n) {
    const int d = dm.d, k = dm.k, dq = dm.dq;
    if (p > d || n >= k * dq + k) return 0.f;
    const float* W = theta + dm.oW + (int64_t)r * k * d * d;
    const float* V = theta + dm.oV + (int64_t)r * 2 * d * k;
    if (n < k * dq) {
        int s = n / dq, q = n - s * dq;
        if (q >= d) return 0.f;
        if (p < d) return tail ? W[((int64_t)s * d + p) * d + q] : W[((int64_t)s * d + q) * d + p];
        return tail ? V[(d + q) * k + s] : V[q * k + s];
    }
    int s = n - k * dq;
    if (p < d) return tail ? V[p * k + s] : V[(d + p) * k + s];
    return theta[dm.ob + (int64_t)r * k + s];
}
__device__ __forceinline__ void split1(float v, float& h, float& l) {
    h = tf32_round(v);
    l = tf32_round(v - h);
}
__global__ void __launch_bounds__(256) k_tc_w_prep(const float* __restrict__ theta, const uint8_t* __restrict__ side,
                                                   float* WT, float* W, float* WT_hi, float* WT_lo, float* W_hi, float* W_lo,
                                                   NtnDims dm, int KpP, int KpN, int NpP, const int* __restrict__ u_rel,
                                                   const int* __restrict__ u_e1, const int* __restrict__ u_e2, int Nu,
                                                   int* __restrict__ t_anchor) {
    const int64_t nA = (int64_t)dm.R * dm.Np * KpP, nB = (int64_t)dm.R * KpN * NpP;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nA) {
        int r = (int)(i / ((int64_t)dm.Np * KpP));
        int rem = (int)(i - (int64_t)r * dm.Np * KpP);
        int n = rem / KpP, p = rem - n * KpP;
        const float v = waug_at(theta, dm, r, side[r] != 0, p, n);
        if (WT_hi) { float h, l; split1(v, h, l); WT_hi[i] = h; WT_lo[i] = l; }   // TMA path: only the pre-split copies are read
        else WT[i] = v;
    } else if (i < nA + nB) {
        i -= nA;
        int r = (int)(i / ((int64_t)KpN * NpP));
        int rem = (int)(i - (int64_t)r * KpN * NpP);
        int p = rem / NpP, n = rem - p * NpP;
        const float v = waug_at(theta, dm, r, side[r] != 0, p, n);
        if (W_hi) { float h, l; split1(v, h, l); W_hi[i] = h; W_lo[i] = l; }
        else W[i] = v;
    } else if (i < nA + nB + Nu) {
        // anchor entity of every unique triple under this evaluation's corrupt sides (NTN.java:146-156): the entity
        // that is NOT replaced -- e1 when the tail is corrupted, e2 when the head is
        const int u = (int)(i - nA - nB);
        t_anchor[u] = side[u_rel[u]] != 0 ? u_e1[u] : u_e2[u];
    }
}


// Weight preparation for the TMA path, tiled.  The generic kernel above walks the OUTPUT arrays element by element
// (two 64-bit divisions and a transposed, uncoalesced read of W per element: 11 us for 13 x 3 slices of 100 x 100, all of
// it beside the entity build on the critical path).  Here one block takes a 32 x 32 tile of a slice W_r[s], reads it
// once, coalesced, splits it, and writes it straight into one orientation and -- through a shared-memory transpose --
// into the other, both coalesced.  Which orientation is the straight one depends on the relation's corrupt side
// (NTN.java:146-156): tail: Waug[p][s dq + q] = W[s][p][q];  head: Waug[p][s dq + q] = W[s][q][p].
// The V / b border (row d and the k trailing columns) and the anchor table are done by k_tc_w_border.  Everything else
// in the four arrays is padding: zeroed once when they are allocated, never written again.
__global__ void __launch_bounds__(256) k_tc_w_tiles(const float* __restrict__ theta, const uint8_t* __restrict__ side,
                                                    float* __restrict__ WT_hi, float* __restrict__ WT_lo,
                                                    float* __restrict__ W_hi, float* __restrict__ W_lo,
                                                    NtnDims dm, int KpP, int KpN, int NpP, int tiles) {
    __shared__ float t_hi[32][33], t_lo[32][33];
    const int r = blockIdx.z, s = blockIdx.y, ti = blockIdx.x / tiles, tj = blockIdx.x - ti * tiles;
    const int d = dm.d, dq = dm.dq;
    const bool tail = side[r] != 0;
    const float* Wsl = theta + dm.oW + ((int64_t)r * dm.k + s) * d * d;
    float* WTh = WT_hi + (size_t)r * dm.Np * KpP; float* WTl = WT_lo + (size_t)r * dm.Np * KpP;     // [n][p]
    float* Wh = W_hi + (size_t)r * KpN * NpP;     float* Wl = W_lo + (size_t)r * KpN * NpP;         // [p][n]
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int a = ti * 32 + j, b = tj * 32 + tx;                  // element W[s][a][b], b contiguous
        float hi = 0.f, lo = 0.f;
        if (a < d && b < d) {
            const float v = __ldg(Wsl + (size_t)a * d + b);
            hi = tf32_round(v); lo = tf32_round(v - hi);
            // straight orientation: tail -> W[p = a][n = s dq + b];  head -> WT[n = s dq + a][p = b]
            if (tail) { Wh[(size_t)a * NpP + s * dq + b] = hi; Wl[(size_t)a * NpP + s * dq + b] = lo; }
            else { WTh[(size_t)(s * dq + a) * KpP + b] = hi; WTl[(size_t)(s * dq + a) * KpP + b] = lo; }
        }
        t_hi[j][tx] = hi; t_lo[j][tx] = lo;
    }
    __syncthreads();
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int a = ti * 32 + tx, b = tj * 32 + j;                  // element W[s][a][b], now a contiguous
        if (a < d && b < d) {
            const float hi = t_hi[tx][j], lo = t_lo[tx][j];
            // transposed orientation: tail -> WT[n = s dq + b][p = a];  head -> W[p = b][n = s dq + a]
            if (tail) { WTh[(size_t)(s * dq + b) * KpP + a] = hi; WTl[(size_t)(s * dq + b) * KpP + a] = lo; }
            else { Wh[(size_t)b * NpP + s * dq + a] = hi; Wl[(size_t)b * NpP + s * dq + a] = lo; }
        }
    }
}

// the V / b border of Waug (k_w_prep in ntn_kernels.cu spells the layout out) in both orientations, and the anchor
// entity of every unique triple under this evaluation's corrupt sides
__global__ void __launch_bounds__(256) k_tc_w_border(const float* __restrict__ theta, const uint8_t* __restrict__ side,
                                                     float* __restrict__ WT_hi, float* __restrict__ WT_lo,
                                                     float* __restrict__ W_hi, float* __restrict__ W_lo,
                                                     NtnDims dm, int KpP, int KpN, int NpP, const int* __restrict__ u_rel,
                                                     const int* __restrict__ u_e1, const int* __restrict__ u_e2, int Nu,
                                                     int* __restrict__ t_anchor) {
    const int d = dm.d, k = dm.k, dq = dm.dq;
    const int per = k * d + k * (d + 1);                              // row d under the slices, then the k trailing columns
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)dm.R * per) {
        const int r = (int)(i / per), e = (int)(i - (int64_t)r * per);
        int p, n;
        if (e < k * d) { const int s = e / d; p = d; n = s * dq + (e - s * d); }
        else { const int e2 = e - k * d, s = e2 / (d + 1); p = e2 - s * (d + 1); n = k * dq + s; }
        const float v = waug_at(theta, dm, r, side[r] != 0, p, n);
        const float hi = tf32_round(v), lo = tf32_round(v - hi);
        const size_t ia = (size_t)r * dm.Np * KpP + (size_t)n * KpP + p, ib = (size_t)r * KpN * NpP + (size_t)p * NpP + n;
        WT_hi[ia] = hi; WT_lo[ia] = lo; W_hi[ib] = hi; W_lo[ib] = lo;
    } else if (i < (int64_t)dm.R * per + Nu) {
        const int u = (int)(i - (int64_t)dm.R * per);
        t_anchor[u] = side[u_rel[u]] != 0 ? u_e1[u] : u_e2[u];
    }
}

// A operand of the forward contraction for the all-TMA path: the anchor entity's augmented row of every unique triple,
// gathered into unique-triple order and split into TF32 hi / lo (NTN.java:146-156 picks the anchor, :163-197 contracts
// it).  One thread per (unique triple, float4 chunk); E was written by the entity build just before and sits in L2.
__global__ void __launch_bounds__(256) k_tc_anchor_rows(const float* __restrict__ E, const int* __restrict__ t_anchor,
                                                        float* __restrict__ EA_hi, float* __restrict__ EA_lo, int Nu, int Kp) {
    const int cpr = Kp >> 2;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)Nu * cpr) return;
    const int u = (int)(i / cpr), c = (int)(i - (int64_t)u * cpr);
    const float4 v = ldg4g(E + (size_t)__ldg(t_anchor + u) * Kp + 4 * c);
    float4 hi, lo;
    split_tf32(v, hi, lo);
    *reinterpret_cast<float4*>(EA_hi + (size_t)u * Kp + 4 * c) = hi;
    *reinterpret_cast<float4*>(EA_lo + (size_t)u * Kp + 4 * c) = lo;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static inline int rup(int a, int b) { return (a + b - 1) / b * b; }
static inline int cdivi(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

int ntn_tc_supported(const ntn_handle* h) {
    (void)h;
    return 1;
}

// tiled tensor map over a row-major [rows][inner] FP32 array: boxes of 32 floats (one 128-byte swizzle row) x box_rows
// (the driver entry point is resolved at run time so that the library does not link libcuda and still loads on a
//  machine without a driver, where every entry point reports NTN_ECUDA)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool make_map(CUtensorMap* m, const float* base, uint64_t inner, uint64_t rows, uint32_t box_rows) {
    static EncodeTiledFn cuTensorMapEncodeTiled = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            fn = nullptr;
        return (EncodeTiledFn)fn;
    }();
    if (!cuTensorMapEncodeTiled) return false;
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {inner * sizeof(float)};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return cuTensorMapEncodeTiled(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static TcArgs tc_args(ntn_handle* h, bool chunks) {
    TcState* st = (TcState*)h->tc;
    const DeviceBatch& b = h->b;
    TcArgs a{};
    a.t_rel = chunks ? b.ch_rel : b.t_rel; a.t_u0 = chunks ? b.ch_u0 : b.t_u0; a.t_n = chunks ? b.ch_n : b.t_n;
    a.u_e1 = b.u_e1; a.u_e2 = b.u_e2; a.t_anchor = st->t_anchor; a.side = h->side; a.E = h->E; a.M = b.M; a.M_lo = b.m_split ? b.M_lo : nullptr;
    a.WT = st->WT; a.W = st->W;
    a.T = b.T; a.GA = b.GA; a.dWpart = b.dWpart;
    a.Kp = h->dm.Kp; a.Np = h->dm.Np; a.KpP = st->KpP; a.KpN = st->KpN; a.NpP = st->NpP;
    a.err = st->err;
    a.dbg = st->dbg;
    a.b_rows_per_rel = 0; a.b_box_rows = 0;
    return a;
}

int ntn_tc_prepare_batch(ntn_handle* h) {
    const NtnDims& dm = h->dm;
    if (!h->tc) {
        TcState* st = new TcState();
        h->tc = st;
        st->KpP = rup(dm.Kp, TC_BK); st->KpN = rup(dm.Kp, 16); st->NpP = rup(dm.Np, TC_BK);
        size_t nA = (size_t)dm.R * dm.Np * st->KpP, nB = (size_t)dm.R * st->KpN * st->NpP;
        cudaError_t e;
        if ((e = cudaMalloc((void**)&st->WT, nA * 4)) || (e = cudaMalloc((void**)&st->W, nB * 4)) ||
            (e = cudaMalloc((void**)&st->err, 4)) || (e = cudaMemset(st->err, 0, 4)) ||
            (getenv("NTN_TC_TIMELINE") && ((e = cudaMalloc((void**)&st->dbg, 3 * 256 * 8)) || (e = cudaMemset(st->dbg, 0, 3 * 256 * 8))))) {
            h->err = std::string("tcgen05 state: ") + cudaGetErrorString(e);
            return NTN_ECUDA;
        }
        const char* tma_env = getenv("NTN_TC_TMA");
        st->use_tma = !(tma_env && tma_env[0] == '0');
        if (st->use_tma) {
            if ((e = cudaMalloc((void**)&st->WT_hi, nA * 4)) || (e = cudaMalloc((void**)&st->WT_lo, nA * 4)) ||
                (e = cudaMalloc((void**)&st->W_hi, nB * 4)) || (e = cudaMalloc((void**)&st->W_lo, nB * 4))) {
                h->err = std::string("tcgen05 TMA state: ") + cudaGetErrorString(e);
                return NTN_ECUDA;
            }
            // padding rows / columns stay zero for the life of the handle (the per-evaluation kernels only write live entries)
            if ((e = cudaMemset(st->WT_hi, 0, nA * 4)) || (e = cudaMemset(st->WT_lo, 0, nA * 4)) ||
                (e = cudaMemset(st->W_hi, 0, nB * 4)) || (e = cudaMemset(st->W_lo, 0, nB * 4))) {
                h->err = std::string("tcgen05 TMA state: ") + cudaGetErrorString(e);
                return NTN_ECUDA;
            }
            st->box_fwd = std::min(TC_NCH, dm.Np);
            st->box_ge = std::min(TC_NCH, st->KpN);
            st->box_ge2 = std::min(128, st->KpN);
            if (!make_map(&st->mWT_hi, st->WT_hi, st->KpP, (uint64_t)dm.R * dm.Np, st->box_fwd) ||
                !make_map(&st->mWT_lo, st->WT_lo, st->KpP, (uint64_t)dm.R * dm.Np, st->box_fwd) ||
                !make_map(&st->mW_hi, st->W_hi, st->NpP, (uint64_t)dm.R * st->KpN, st->box_ge) ||
                !make_map(&st->mW_lo, st->W_lo, st->NpP, (uint64_t)dm.R * st->KpN, st->box_ge) ||
                !make_map(&st->mW2_hi, st->W_hi, st->NpP, (uint64_t)dm.R * st->KpN, st->box_ge2) ||
                !make_map(&st->mW2_lo, st->W_lo, st->NpP, (uint64_t)dm.R * st->KpN, st->box_ge2)) {
                h->err = "cuTensorMapEncodeTiled failed";
                return NTN_ECUDA;
            }
        }
        { const char* e2 = getenv("NTN_TC_GE2"); st->ge2 = st->use_tma && !(e2 && e2[0] == '0'); }
        const char* dw_env = getenv("NTN_TC_DW");
        st->dw_mn = dw_env ? (dw_env[0] == 'm') : true;      // NTN_TC_DW=k selects the K-major transposing producers
        // all-TMA forward / entity-gradient contractions: opt-in (NTN_TC_TMA_A=1); needs the TMA-fed weights and the
        // MN-major dw (the only dw producers that read M pre-split).  Measured at 20,000 x 10 on one B200 (same box,
        // alternating runs): entity-gradient contraction 28.5 -> 26.9 us, but forward 23.4 -> 27.1 us (the anchor-row gather
        // kernel), score 43 -> 47 us (M written twice), entity pull 56 -> 65 us (48 MB of M_hi / M_lo push T' out of the
        // L2), evaluation 0.2164 -> 0.2255 ms.  The cycle timeline (scripts/tc_timeline.py) shows why feeding the operand
        // faster buys so little: with every tile coming by TMA a k-block still takes ~1460 cycles -- twelve
        // 128 x 104 x 8 MMAs at ~122 cycles each, and the 128 x 152 x 8 MMAs of the forward contraction take ~126: the
        // issue cost of a kind::tf32 MMA is that of a full 128 x 256 x 8 one (139 cycles at the measured cuBLAS rate), so
        // these contractions run the tensor pipe at N/256 of its rate whatever feeds them.
        const char* ta_env = getenv("NTN_TC_TMA_A");
        st->tma_a = st->use_tma && st->dw_mn && ta_env && ta_env[0] == '1';
        if ((e = cudaFuncSetAttribute(k_tc_gemm<TC_DW, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES)) ||
            (e = cudaFuncSetAttribute(k_tc_gemm<TC_FWD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES)) ||
            (e = cudaFuncSetAttribute(k_tc_gemm<TC_FWD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES)) ||
            (e = cudaFuncSetAttribute(k_tc_gemm<TC_DW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES)) ||
            (e = cudaFuncSetAttribute(k_tc_gemm<TC_GE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES)) ||
            (e = cudaFuncSetAttribute(k_tc_gemm<TC_GE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES)) ||
            (e = cudaFuncSetAttribute(k_tc_gemm<TC_FWD, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES)) ||
            (e = cudaFuncSetAttribute(k_tc_gemm<TC_GE, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES)) ||
            (e = cudaFuncSetAttribute(k_tc_gemm<TC_GE2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES))) {
            h->err = std::string("tcgen05 smem attribute: ") + cudaGetErrorString(e);
            return NTN_ECUDA;
        }
    }
    {
        // per-evaluation anchor table, sized for this batch job
        TcState* st = (TcState*)h->tc;
        const size_t need = (size_t)std::max(h->b.Nu, 1);
        if (st->anchor_cap < need) {
            cudaFree(st->t_anchor); st->t_anchor = nullptr; st->anchor_cap = 0;
            cudaError_t e = cudaMalloc((void**)&st->t_anchor, need * sizeof(int));
            if (e != cudaSuccess) { h->err = std::string("tcgen05 anchor table: ") + cudaGetErrorString(e); return NTN_ECUDA; }
            st->anchor_cap = need;
        }
        // all-TMA path: the gathered anchor rows (grow-only) and the tensor maps of this batch job's A operands
        h->b.m_split = false;
        if (st->tma_a && h->b.Nu > 0) {
            const uint64_t rows_a = (uint64_t)std::max(h->b.Nu, 128);       // at least one 128-row box of allocated memory
            const size_t ea_need = (size_t)rows_a * (size_t)dm.Kp;
            if (st->ea_cap < ea_need) {
                cudaFree(st->EA_hi); cudaFree(st->EA_lo); st->EA_hi = st->EA_lo = nullptr; st->ea_cap = 0;
                cudaError_t e;
                if ((e = cudaMalloc((void**)&st->EA_hi, ea_need * 4)) || (e = cudaMalloc((void**)&st->EA_lo, ea_need * 4))) {
                    h->err = std::string("tcgen05 anchor rows: ") + cudaGetErrorString(e); return NTN_ECUDA;
                }
                st->ea_cap = ea_need;
            }
            if (!make_map(&st->mEA_hi, st->EA_hi, dm.Kp, rows_a, 128) || !make_map(&st->mEA_lo, st->EA_lo, dm.Kp, rows_a, 128) ||
                !make_map(&st->mM_hi, h->b.M, dm.Np, rows_a, 128) || !make_map(&st->mM_lo, h->b.M_lo, dm.Np, rows_a, 128)) {
                h->err = "cuTensorMapEncodeTiled failed (A operands)";
                return NTN_ECUDA;
            }
            h->b.m_split = true;
        }
    }
    return NTN_OK;
}

void ntn_tc_release(ntn_handle* h) {
    TcState* st = (TcState*)h->tc;
    if (!st) return;
    cudaFree(st->WT); cudaFree(st->W); cudaFree(st->err); cudaFree(st->dbg); cudaFree(st->t_anchor);
    cudaFree(st->WT_hi); cudaFree(st->WT_lo); cudaFree(st->W_hi); cudaFree(st->W_lo);
    cudaFree(st->EA_hi); cudaFree(st->EA_lo);
    delete st;
    h->tc = nullptr;
}

int ntn_tc_w_prep(ntn_handle* h, const float* theta, cudaStream_t stq) {
    TcState* st = (TcState*)h->tc;
    const NtnDims& dm = h->dm;
    static const bool tiled = [] { const char* e = getenv("NTN_TC_WPREP"); return !(e && e[0] == '0'); }();
    if (st->use_tma && tiled) {
        const int tiles = (dm.d + 31) / 32;
        k_tc_w_tiles<<<dim3(tiles * tiles, dm.k, dm.R), 256, 0, stq>>>(theta, h->side, st->WT_hi, st->WT_lo, st->W_hi, st->W_lo, dm,
                                                                         st->KpP, st->KpN, st->NpP, tiles);
        const int64_t nb = (int64_t)dm.R * (dm.k * dm.d + dm.k * (dm.d + 1)) + h->b.Nu;
        k_tc_w_border<<<cdivi(nb, 256), 256, 0, stq>>>(theta, h->side, st->WT_hi, st->WT_lo, st->W_hi, st->W_lo, dm, st->KpP, st->KpN,
                                                       st->NpP, h->b.u_rel, h->b.u_e1, h->b.u_e2, h->b.Nu, st->t_anchor);
        h->launches += 2;
        return NTN_OK;
    }
    int64_t n = (int64_t)dm.R * dm.Np * st->KpP + (int64_t)dm.R * st->KpN * st->NpP + h->b.Nu;
    k_tc_w_prep<<<cdivi(n, 256), 256, 0, stq>>>(theta, h->side, st->WT, st->W, st->WT_hi, st->WT_lo, st->W_hi, st->W_lo, dm,
                                                      st->KpP, st->KpN, st->NpP, h->b.u_rel, h->b.u_e1, h->b.u_e2, h->b.Nu, st->t_anchor);
    h->launches++;
    return NTN_OK;
}

// persistent grid: one CTA per SM walks the work items round-robin (NTN_TC_PERSIST=0: one CTA per item)
static dim3 tc_grid(ntn_handle* h, TcArgs& a, int nitems, int ny, int nz) {
    static const bool persist = [] { const char* e = getenv("NTN_TC_PERSIST"); return !(e && e[0] == '0'); }();
    a.nitems = nitems; a.ny = ny; a.nz = nz;
    const int64_t nwork = (int64_t)nitems * ny * nz;
    return dim3((unsigned)std::max<int64_t>(1, persist ? std::min<int64_t>(nwork, h->sm_count) : nwork));
}

int ntn_tc_gemm_fwd(ntn_handle* h) {
    TcArgs a = tc_args(h, false);
    TcState* st = (TcState*)h->tc;
    dim3 g = tc_grid(h, a, h->b.ntiles, cdivi(h->dm.Np, TC_NCH), 1);
    if (h->b.m_split) {
        // all-TMA: gather + split the anchor rows once, then every operand tile is a TMA box
        const int64_t nthr = (int64_t)h->b.Nu * (h->dm.Kp / 4);
        k_tc_anchor_rows<<<cdivi(nthr, 256), 256, 0, h->stream>>>(h->E, st->t_anchor, st->EA_hi, st->EA_lo, h->b.Nu, h->dm.Kp);
        h->launches++;
        a.b_rows_per_rel = h->dm.Np; a.b_box_rows = st->box_fwd;
        k_tc_gemm<TC_FWD, true, false, true><<<g, TC_THREADS, TC_SMEM_BYTES, h->stream>>>(a, st->mWT_hi, st->mWT_lo, st->mEA_hi, st->mEA_lo);
    } else if (st->use_tma) {
        a.b_rows_per_rel = h->dm.Np; a.b_box_rows = st->box_fwd;
        k_tc_gemm<TC_FWD, true><<<g, TC_THREADS, TC_SMEM_BYTES, h->stream>>>(a, st->mWT_hi, st->mWT_lo, st->mWT_hi, st->mWT_lo);
    } else {
        k_tc_gemm<TC_FWD, false><<<g, TC_THREADS, TC_SMEM_BYTES, h->stream>>>(a, st->mWT_hi, st->mWT_lo, st->mWT_hi, st->mWT_lo);
    }
    h->launches++;
    return NTN_OK;
}

int ntn_tc_gemm_dw(ntn_handle* h, cudaStream_t stq) {
    TcArgs a = tc_args(h, true);
    dim3 g = tc_grid(h, a, h->b.nchunks, cdivi(h->dm.Np, TC_NCH), cdivi(h->dm.Kp, 128));
    // the dW contraction runs beside the entity/word pulls (forked stream); a persistent CTA owns its SM (229 KB of
    // shared memory, 53 k registers), so its grid size decides how many SMs the pulls can use meanwhile
    // (measured at 20,000 x 10: 148 CTAs 0.2258 ms per evaluation, 111: 0.2191, 74: 0.2180, 48: 0.2168).  At the
    // reference's batch size the contraction has ~120 us of shadow to hide in, so it gets a third of the SMs; a large batch
    // job makes it the longest branch and it keeps the whole grid.  NTN_DW_CTAS overrides (0 = all SMs).
    static const int dw_env = [] { const char* e = getenv("NTN_DW_CTAS"); return e ? atoi(e) : -1; }();
    const int dw_ctas = dw_env >= 0 ? dw_env : (h->b.Nu <= 65536 ? (h->sm_count + 2) / 3 : 0);
    if (dw_ctas > 0 && (int)g.x > dw_ctas) g.x = dw_ctas;
    TcState* st = (TcState*)h->tc;
    if (st->dw_mn) k_tc_gemm<TC_DW, false, true><<<g, TC_THREADS, TC_SMEM_BYTES, stq>>>(a, st->mWT_hi, st->mWT_lo, st->mWT_hi, st->mWT_lo);
    else k_tc_gemm<TC_DW, false><<<g, TC_THREADS, TC_SMEM_BYTES, stq>>>(a, st->mWT_hi, st->mWT_lo, st->mWT_hi, st->mWT_lo);
    h->launches++;
    return NTN_OK;
}

int ntn_tc_gemm_ge(ntn_handle* h) {
    TcArgs a = tc_args(h, false);
    TcState* st = (TcState*)h->tc;
    if (st->ge2 && !h->b.m_split && h->b.ngtiles > 0) {
        // swapped operand roles: A = a 128-row chunk of Waug_r (TMA), N = the tile's unique triples
        a.t_rel = h->b.g_rel; a.t_u0 = h->b.g_u0; a.t_n = h->b.g_n;
        dim3 g2 = tc_grid(h, a, h->b.ngtiles, 1, cdivi(h->dm.Kp, 128));
        a.b_rows_per_rel = st->KpN; a.b_box_rows = st->box_ge2;
        k_tc_gemm<TC_GE2, false><<<g2, TC_THREADS, TC_SMEM_BYTES, h->stream>>>(a, st->mW2_hi, st->mW2_lo, st->mW2_hi, st->mW2_lo);
        h->launches++;
        return NTN_OK;
    }
    dim3 g = tc_grid(h, a, h->b.ntiles, cdivi(st->KpN, TC_NCH), 1);
    if (h->b.m_split) {
        a.b_rows_per_rel = st->KpN; a.b_box_rows = st->box_ge;
        k_tc_gemm<TC_GE, true, false, true><<<g, TC_THREADS, TC_SMEM_BYTES, h->stream>>>(a, st->mW_hi, st->mW_lo, st->mM_hi, st->mM_lo);
    } else if (st->use_tma) {
        a.b_rows_per_rel = st->KpN; a.b_box_rows = st->box_ge;
        k_tc_gemm<TC_GE, true><<<g, TC_THREADS, TC_SMEM_BYTES, h->stream>>>(a, st->mW_hi, st->mW_lo, st->mW_hi, st->mW_lo);
    } else {
        k_tc_gemm<TC_GE, false><<<g, TC_THREADS, TC_SMEM_BYTES, h->stream>>>(a, st->mW_hi, st->mW_lo, st->mW_hi, st->mW_lo);
    }
    h->launches++;
    return NTN_OK;
}

// debug: copy the 3 x 256 cycle stamps (as doubles relative to each mode's first stamp); returns count or 0
int64_t ntn_tc_fetch_timeline(ntn_handle* h, float* out, int64_t cap) {
    TcState* st = (TcState*)h->tc;
    if (!st || !st->dbg) return 0;
    if (out && cap >= 768) {
        long long tmp[768];
        cudaMemcpy(tmp, st->dbg, sizeof(tmp), cudaMemcpyDeviceToHost);
        for (int m = 0; m < 3; ++m) {
            long long t0 = tmp[m * 256];
            for (int i = 0; i < 256; ++i) out[m * 256 + i] = tmp[m * 256 + i] ? (float)(tmp[m * 256 + i] - t0) : -1.f;
        }
    }
    return 768;
}

int ntn_tc_check_error(ntn_handle* h) {
    TcState* st = (TcState*)h->tc;
    if (!st) return 0;
    int v = 0;
    cudaMemcpy(&v, st->err, 4, cudaMemcpyDeviceToHost);
    return v;
}
