// Gradient exchange inside the evaluation, over peer memory (NVLink 5 / NVSwitch), for the data-parallel split of
// NTN.computeAt: the batch columns are sharded over the GPUs of one box (the cost and the gradient are sums over
// columns, /root/reference/src/NTN.java:114-409, and the word-gradient accumulation of :411-430 is linear), so every
// rank's gradient buffer must end with the SUM over the ranks.
//
// The buffers are symmetric allocations (same size on every rank, every rank's copy mapped into every process; torch
// symmetric memory does the allocation and the handle exchange -- plumbing).  The kernels here are the data plane:
//   * the word pull is cut into chunks (ntn_word_chunk); as soon as chunk c is final in this rank's buffer, an exchange
//     kernel all-reduces it on a side stream while the pull goes on with chunk c + 1; the last chunk, the W/V/b/U blocks and
//     the {cost, n_active} tail follow the final reduce.  Only that last exchange is not overlapped.  (Variant,
//     ntn_exchange_bind_ge: the same done one stage earlier, on the entity gradient, beside the entity pull.);
//   * an all-reduce of a range is two-shot: rank r owns the r-th slice, SUMS that slice over all ranks and writes the
//     result into every rank's buffer.
//       mode 2 (push, the default when a scratch allocation is bound): NO peer loads at all.  The producer of a range --
//              the word-pull kernel itself, k_word_pull / k_word_finish_hub in ntn_kernels.cu -- stores every 16-byte group
//              that another rank owns straight into that rank's scratch while it computes it (posted writes over NVLink:
//              the reduce-scatter half rides inside the pull and costs it nothing but store slots); after ONE flag
//              barrier the owner adds its own group and the G-1 scratch copies in rank order (all local reads) and stores
//              the sum into every rank's buffer -- G stores, or one multimem.st on the multicast alias (the NVSwitch
//              replicates it) on a box of >= 4 ranks.  Loads from a peer cost a 2-3 us round trip each and a few dozen blocks
//              cannot keep enough of them in flight (measured on 2 GPUs: the pull-based exchange of mode 0 took
//              ~35 us per 7 MB chunk, longer than the chunk's pull); stores are fire-and-forget;
//       mode 0 (peer loads / stores): 16-byte loads from every peer in rank order 0..G-1 (a fixed order: every rank
//              ends with bit-identical sums and repeated runs agree), 16-byte stores to every peer;
//       mode 1 (multimem, NVLS): multimem.ld_reduce on the multicast alias -- the NVSwitch adds the G copies in the
//              fabric and returns one value -- then multimem.st broadcasts it; one load and one store per 16 bytes
//              whatever G is;
//   * ranks synchronise through 32-bit flags in the signal pad of the symmetric allocation: block b of a kernel pairs
//     with block b of the same kernel on every peer (x_barrier below: a signal is a store of the barrier's epoch number
//     with release semantics at system scope, a wait polls the own flag with acquire loads).  A start barrier says "every
//     rank has finished producing this range", the end barrier -- only in the kernel that closes an evaluation -- "every
//     rank's stores have landed".  Every wait is bounded: a lost peer sets an error flag (ntn_exchange_check) instead of
//     hanging the GPU.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

#include "ntn_internal.h"

struct XArgs {
    float* peer[NTN_MAX_PEERS];
    float* mc;
    uint32_t* sig[NTN_MAX_PEERS];
    float* scratch[NTN_MAX_PEERS];
    int rank, world, mode, last, mcst;
    XRange rg[3];
    unsigned* err;
    unsigned long long timeout_ns;
    uint32_t* epoch;                     // {barriers completed by this binding / 1, blocks that have left the running kernel}
};

__device__ __forceinline__ unsigned long long x_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Flag barrier between block b of this kernel and block b of the same kernel on every peer.  Every barrier of a
// binding's life has a number (its epoch: the ranks launch the same sequence of exchange kernels, so they count alike);
// a rank SIGNALS by storing the epoch into its slot of every peer's signal pad -- a posted store with release semantics
// at system scope, nothing comes back over NVLink -- and WAITS until every peer's slot in its own pad has reached the
// epoch (acquire loads of local memory).  Slots only grow, so a peer that is already one barrier ahead still satisfies
// the wait.  (A compare-and-swap handshake, which needs no counter, measured 9-10 us per barrier on 2 GPUs: every put
// waits for its atomic's round trip; the store-based one costs one one-way trip.)  The epoch counter is device memory of
// this rank: read by every block when the kernel starts, advanced by the last block to leave.
__device__ __forceinline__ void x_signal(uint32_t* flag, uint32_t epoch) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
}
__device__ __forceinline__ bool x_wait(const uint32_t* flag, uint32_t epoch, unsigned long long timeout_ns) {
    unsigned long long t0 = 0;
    for (unsigned spins = 0;; ++spins) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int32_t)(v - epoch) >= 0) return true;
        if ((spins & 255u) == 255u) {                          // the clock is only consulted now and then
            const unsigned long long t = x_now();
            if (t0 == 0) t0 = t;
            else if (t - t0 > timeout_ns) return false;
        }
    }
}

// block b of this rank <-> block b of every peer.  The block's earlier stores are ordered before the signal by the
// __syncthreads (every thread's stores happen before the signalling thread goes on) and the release at system scope of
// the signalling store (cumulative) -- no fence per thread (32 k of them measured ~5 us per barrier).
__device__ __forceinline__ void x_barrier(const XArgs& a, uint32_t epoch) {
    __syncthreads();
    const int t = threadIdx.x;
    if (t < a.world && t != a.rank) {
        const size_t slot = NTN_XCHG_SIG_WORD0 + (size_t)blockIdx.x * a.world;
        x_signal(a.sig[t] + slot + a.rank, epoch);
        if (!x_wait(a.sig[a.rank] + slot + t, epoch, a.timeout_ns)) atomicExch(a.err, 1u);
    }
    __syncthreads();
}
// epochs of this launch: base + 1 (start barrier), base + 2 (end barrier); the last block to leave advances the counter
__device__ __forceinline__ uint32_t x_epoch_begin(const XArgs& a) {
    __shared__ uint32_t s_base;
    if (threadIdx.x == 0) s_base = *reinterpret_cast<volatile uint32_t*>(a.epoch);
    __syncthreads();
    return s_base;
}
__device__ __forceinline__ void x_epoch_end(const XArgs& a, uint32_t base) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(a.epoch + 1, 1u) == gridDim.x - 1) { a.epoch[1] = 0u; __threadfence(); a.epoch[0] = base + 2u; }
    }
}

__device__ __forceinline__ float4 x_ld(const float* p) {      // peer data written by another GPU's earlier kernel: not through
    float4 v;                                                  // the non-coherent path
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void x_st(float* p, const float4& v) {
    asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 x_mc_ld_reduce(const float* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void x_mc_st(float* p, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

#define X_THREADS 512
#define X_UNROLL_MAX 4                              // 16-byte requests in flight per thread and peer (NVLink round trips are ~2-3 us)

template <int MODE, int G>                      // G = world size (compile time: the peer loop unrolls), 0: run time
__global__ void __launch_bounds__(X_THREADS) k_exchange(XArgs a) {
    constexpr int X_UNROLL = (MODE == 0 && G == 8) ? 2 : X_UNROLL_MAX;   // 8 peers x 4 float4 would not fit the registers
    const int world = G ? G : a.world;
    const uint32_t ep = x_epoch_begin(a);
    x_barrier(a, ep + 1);                       // every rank has finished producing the ranges
#pragma unroll 1
    for (int q = 0; q < 3; ++q) {
        const int64_t n4 = a.rg[q].n >> 2;
        if (n4 == 0) continue;
        const int64_t per = (n4 + world - 1) / world;
        const int64_t lo = min(n4, (int64_t)a.rank * per), hi = min(n4, lo + per);     // this rank's slice, in float4s
        const int64_t base = a.rg[q].off >> 2;
        const int64_t stride = (int64_t)gridDim.x * X_THREADS;
        for (int64_t i0 = lo + (int64_t)blockIdx.x * X_THREADS + threadIdx.x; i0 < hi; i0 += stride * X_UNROLL) {
            float4 acc[X_UNROLL];
            if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < X_UNROLL; ++j) {
                    const int64_t i = i0 + j * stride;
                    if (i < hi) acc[j] = x_mc_ld_reduce(a.mc + 4 * (base + i));
                }
#pragma unroll
                for (int j = 0; j < X_UNROLL; ++j) {
                    const int64_t i = i0 + j * stride;
                    if (i < hi) x_mc_st(a.mc + 4 * (base + i), acc[j]);
                }
            } else {
                float4 v[X_UNROLL][G ? G : 1];
                if (G) {
#pragma unroll
                    for (int j = 0; j < X_UNROLL; ++j) {
                        const int64_t i = i0 + j * stride;
#pragma unroll
                        for (int s = 0; s < (G ? G : 1); ++s)
                            if (i < hi) v[j][s] = x_ld(a.peer[s] + 4 * (base + i));
                    }
#pragma unroll
                    for (int j = 0; j < X_UNROLL; ++j) {
                        acc[j] = v[j][0];
#pragma unroll
                        for (int s = 1; s < (G ? G : 1); ++s) { acc[j].x += v[j][s].x; acc[j].y += v[j][s].y; acc[j].z += v[j][s].z; acc[j].w += v[j][s].w; }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < X_UNROLL; ++j) {
                        const int64_t i = i0 + j * stride;
                        acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (i < hi)
                            for (int s = 0; s < world; ++s) {                      // rank order: a fixed summation order
                                const float4 t = x_ld(a.peer[s] + 4 * (base + i));
                                if (s == 0) acc[j] = t; else { acc[j].x += t.x; acc[j].y += t.y; acc[j].z += t.z; acc[j].w += t.w; }
                            }
                    }
                }
#pragma unroll
                for (int j = 0; j < X_UNROLL; ++j) {
                    const int64_t i = i0 + j * stride;
                    if (i < hi)
#pragma unroll
                        for (int s = 0; s < (G ? G : world); ++s) x_st(a.peer[s] + 4 * (base + i), acc[j]);
                }
            }
        }
    }
    if (a.last) x_barrier(a, ep + 2);           // end of the evaluation: every rank's stores have landed everywhere
    x_epoch_end(a, ep);
}


// ---- mode 2: push.  Scratch region of a range: NTN_XCHG_RANGE_PAD * id floats after the range's own offset; inside it
// row s (per4 groups) holds rank s's contribution to the slice the scratch's owner owns.
__device__ __forceinline__ int64_t x_scr_off(const XRange& r) { return r.off + (int64_t)r.id * NTN_XCHG_RANGE_PAD; }

template <int G>
__global__ void __launch_bounds__(X_THREADS) k_exchange_push(XArgs a) {
    const int world = G ? G : a.world;
    const int64_t stride = (int64_t)gridDim.x * X_THREADS;
    const int64_t tid = (int64_t)blockIdx.x * X_THREADS + threadIdx.x;
    const uint32_t ep = x_epoch_begin(a);
    // phase 0: ranges whose producer did not push (the W/V/b/U blocks, the {cost, n_active} tail): store this rank's
    // groups of every other rank's slice into the owner's scratch
#pragma unroll 1
    for (int q = 0; q < 3; ++q) {
        const int64_t n4 = a.rg[q].n >> 2;
        if (n4 == 0 || a.rg[q].pushed) continue;
        const int64_t per = (n4 + world - 1) / world;
        const float4* src = reinterpret_cast<const float4*>(a.peer[a.rank] + a.rg[q].off);
        for (int s = 0; s < world; ++s) {
            if (s == a.rank) continue;
            const int64_t lo = min(n4, (int64_t)s * per), hi = min(n4, lo + per);
            float4* dst = reinterpret_cast<float4*>(a.scratch[s] + x_scr_off(a.rg[q])) + (int64_t)a.rank * per;
            for (int64_t i = lo + tid; i < hi; i += stride) dst[i - lo] = src[i];
        }
    }
    x_barrier(a, ep + 1);                       // every rank's contributions to my slices have landed in my scratch
#pragma unroll 1
    for (int q = 0; q < 3; ++q) {
        const int64_t n4 = a.rg[q].n >> 2;
        if (n4 == 0) continue;
        const int64_t per = (n4 + world - 1) / world;
        const int64_t lo = min(n4, (int64_t)a.rank * per), hi = min(n4, lo + per);
        const float* mine = a.peer[a.rank] + a.rg[q].off;
        const float* scr = a.scratch[a.rank] + x_scr_off(a.rg[q]);
        constexpr int U = (G == 0 || G == 8) ? 2 : 4;                // 16-byte groups in flight per thread and source
        for (int64_t i0 = lo + tid; i0 < hi; i0 += stride * U) {
            float4 acc[U];
            float4 v[U][G ? G : NTN_MAX_PEERS];
            // (plain L2-cached loads: everything read here is local memory, complete since the barrier above; L1 holds
            //  none of it -- the kernel has not touched these lines before)
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const int64_t i = i0 + j * stride;
#pragma unroll
                for (int s = 0; s < (G ? G : NTN_MAX_PEERS); ++s)
                    if (s < world && i < hi)
                        v[j][s] = __ldcg(reinterpret_cast<const float4*>((s == a.rank) ? mine + 4 * i : scr + 4 * ((int64_t)s * per + (i - lo))));
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                acc[j] = v[j][0];                                   // rank order: a fixed summation order
#pragma unroll
                for (int s = 1; s < (G ? G : NTN_MAX_PEERS); ++s)
                    if (s < world) { acc[j].x += v[j][s].x; acc[j].y += v[j][s].y; acc[j].z += v[j][s].z; acc[j].w += v[j][s].w; }
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const int64_t i = i0 + j * stride;
                if (i >= hi) continue;
                if (a.mcst) {
                    // one store to the multicast alias: the NVSwitch replicates it into every rank's buffer, so the
                    // all-gather half costs this rank one slice of NVLink bandwidth instead of G - 1
                    x_mc_st(a.mc + a.rg[q].off + 4 * i, acc[j]);
                } else {
#pragma unroll
                    for (int s = 0; s < (G ? G : NTN_MAX_PEERS); ++s)
                        if (s < world) *reinterpret_cast<float4*>(a.peer[s] + a.rg[q].off + 4 * i) = acc[j];
                }
            }
        }
    }
    if (a.last) x_barrier(a, ep + 2);           // end of the evaluation: every rank's sums have landed everywhere
    x_epoch_end(a, ep);
}

// the push descriptor of a range for its producer kernel (ntn_kernels.cu)
XPushRange ntn_push_range(const ntn_handle* h, const ExchangeBinding* xb, int64_t off, int64_t n, int id) {
    XPushRange r{};
    if (!xb || h->x_mode != 2) return r;        // world = 0: nothing to push
    const int64_t n4 = n >> 2, per = (n4 + h->x_world - 1) / h->x_world;
    const FastDiv f = make_fastdiv((unsigned)per);
    r.off = off; r.per4 = (unsigned)per; r.per_m = f.m; r.per_s = f.s; r.rank = h->x_rank; r.world = h->x_world;
    for (int s = 0; s < h->x_world; ++s)
        r.dst[s] = xb->scratch[s] + off + (int64_t)id * NTN_XCHG_RANGE_PAD + (int64_t)h->x_rank * per * 4;
    return r;
}

const ExchangeBinding* ntn_find_binding(const ntn_handle* h, const float* d_grad) {
    for (const auto& b : h->xb) if (b.local == d_grad) return &b;
    return nullptr;
}

void ntn_launch_exchange(ntn_handle* h, const ExchangeBinding* xb, XRange r0, XRange r1, XRange r2, bool last, cudaStream_t st) {
    XArgs a{};
    for (int r = 0; r < h->x_world; ++r) { a.peer[r] = xb->peer[r]; a.sig[r] = xb->sig[r]; a.scratch[r] = xb->scratch[r]; }
    a.mc = xb->mc;
    a.rank = h->x_rank; a.world = h->x_world;
    a.mode = h->x_mode;
    a.last = last ? 1 : 0;
    a.mcst = (h->x_mcst && xb->mc) ? 1 : 0;
    a.rg[0] = r0; a.rg[1] = r1; a.rg[2] = r2;
    a.err = h->d_xerr;
    a.epoch = xb->epoch;
    a.timeout_ns = h->x_timeout_ns;
    const int nb = h->x_blocks;
    h->launches++;
    const int sslot = 2 * (NTN_NUM_PHASES + std::min(h->x_stamp++, NTN_MAX_WORD_CHUNKS));
    ntn_stamp(h, sslot, st);
    if (a.mode == 2) {
        if (a.world == 2) k_exchange_push<2><<<nb, X_THREADS, 0, st>>>(a);
        else if (a.world == 4) k_exchange_push<4><<<nb, X_THREADS, 0, st>>>(a);
        else if (a.world == 8) k_exchange_push<8><<<nb, X_THREADS, 0, st>>>(a);
        else k_exchange_push<0><<<nb, X_THREADS, 0, st>>>(a);
        ntn_stamp(h, sslot + 1, st);
        return;
    }
    if (a.mode == 1) k_exchange<1, 0><<<nb, X_THREADS, 0, st>>>(a);
    else if (a.world == 2) k_exchange<0, 2><<<nb, X_THREADS, 0, st>>>(a);
    else if (a.world == 4) k_exchange<0, 4><<<nb, X_THREADS, 0, st>>>(a);
    else if (a.world == 8) k_exchange<0, 8><<<nb, X_THREADS, 0, st>>>(a);
    else k_exchange<0, 0><<<nb, X_THREADS, 0, st>>>(a);
    ntn_stamp(h, sslot + 1, st);
}

// ------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------
void ntn_drop_graphs(ntn_handle* h);

extern "C" int64_t ntn_exchange_scratch_numel(const ntn_handle* h) {
    return h ? h->dm.Pint + 4 + (int64_t)NTN_XCHG_RANGE_PAD * (NTN_MAX_WORD_CHUNKS + 3) : -1;
}

extern "C" int ntn_exchange_bind(ntn_handle* h, int32_t rank, int32_t world, const uint64_t* buffer_ptrs,
                                 const uint64_t* signal_pad_ptrs, uint64_t multicast_ptr, int64_t numel,
                                 const uint64_t* scratch_ptrs, int64_t scratch_numel) {
    if (!h) return NTN_EINVAL;
    if (world < 2 || world > NTN_MAX_PEERS || rank < 0 || rank >= world) { h->err = "exchange: need 2 <= world <= 8 and 0 <= rank < world"; return NTN_EINVAL; }
    if (!buffer_ptrs || !signal_pad_ptrs) { h->err = "exchange: null pointer table"; return NTN_EINVAL; }
    if (numel < h->dm.Pint + 4) { h->err = "exchange: buffer shorter than ntn_internal_dimension() + 4"; return NTN_EINVAL; }
    if (!h->xb.empty() && (h->x_rank != rank || h->x_world != world)) { h->err = "exchange: rank / world differ from the earlier binding"; return NTN_EINVAL; }
    cudaSetDevice(h->device);
    ExchangeBinding b;
    for (int r = 0; r < world; ++r) {
        if (!buffer_ptrs[r] || !signal_pad_ptrs[r] || (buffer_ptrs[r] & 15)) { h->err = "exchange: null or misaligned peer pointer"; return NTN_EINVAL; }
        b.peer[r] = reinterpret_cast<float*>(buffer_ptrs[r]);
        b.sig[r] = reinterpret_cast<uint32_t*>(signal_pad_ptrs[r]);
    }
    b.local = b.peer[rank];
    b.mc = reinterpret_cast<float*>(multicast_ptr);
    b.numel = numel;
    bool have_scratch = scratch_ptrs != nullptr && scratch_numel >= ntn_exchange_scratch_numel(h);
    for (int r = 0; r < world && have_scratch; ++r) {
        if (!scratch_ptrs[r] || (scratch_ptrs[r] & 15)) have_scratch = false;
        b.scratch[r] = reinterpret_cast<float*>(scratch_ptrs[r]);
    }
    if (scratch_ptrs && !have_scratch) { h->err = "exchange: scratch too short (ntn_exchange_scratch_numel), null or misaligned"; return NTN_EINVAL; }
    b.scratch_numel = have_scratch ? scratch_numel : 0;
    // barrier counter of this buffer: two words of this rank's own signal pad just below the flag slots (no peer writes
    // there), so that every handle bound to the buffer counts on.  The pads must be zero when the first handle binds (a
    // fresh symmetric allocation is) and every rank binds before any evaluates (the caller's barrier).
    b.epoch = b.sig[rank] + NTN_XCHG_SIG_WORD0 - 2;
    if (!h->stream4) {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        const char* pe = getenv("NTN_STREAM_PRIORITY");
        if (cudaStreamCreateWithPriority(&h->stream4, cudaStreamNonBlocking, (pe && pe[0] == '0') ? prio_lo : prio_hi) != cudaSuccess) { h->err = "exchange: stream"; return NTN_ECUDA; }
        for (int i = 0; i < NTN_MAX_WORD_CHUNKS; ++i) cudaEventCreateWithFlags(&h->ev_x[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&h->ev_xdone, cudaEventDisableTiming);
        if (cudaMalloc((void**)&h->d_xerr, sizeof(unsigned)) != cudaSuccess) { h->err = "exchange: alloc"; return NTN_ECUDA; }
        cudaMemset(h->d_xerr, 0, sizeof(unsigned));
    }
    h->x_rank = rank; h->x_world = world;
    // NTN_EXCHANGE=push|p2p|multimem, NTN_EXCHANGE_CHUNKS, NTN_EXCHANGE_BLOCKS override the defaults
    // defaults: push when a scratch allocation is bound (its broadcast half through the multicast alias on a large box);
    // otherwise NVLS loads on a large box, peer loads on a small one
    h->x_mode = have_scratch ? 2 : ((multicast_ptr && world >= 4) ? 1 : 0);
    if (const char* e = getenv("NTN_EXCHANGE")) {
        if (!strcmp(e, "p2p") || !strcmp(e, "pull")) h->x_mode = 0;
        else if (!strcmp(e, "multimem") && multicast_ptr) h->x_mode = 1;
        else if (!strcmp(e, "push") && have_scratch) h->x_mode = 2;
    }
    for (const auto& o : h->xb) if ((o.scratch_numel != 0) != have_scratch && h->x_mode == 2) { h->err = "exchange: all bound buffers need a scratch in push mode"; return NTN_EINVAL; }
    for (const auto& o : h->xb) if (!o.mc && h->x_mode == 1) { h->err = "exchange: all bound buffers need a multicast alias in multimem mode"; return NTN_EINVAL; }
    h->x_mcst = world >= 4 ? 1 : 0;          // push mode: broadcast the sums through the multicast alias (when there is one)
    if (const char* e = getenv("NTN_EXCHANGE_MCST")) h->x_mcst = atoi(e) != 0;
    if (const char* e = getenv("NTN_EXCHANGE_TIMEOUT_MS")) h->x_timeout_ns = (unsigned long long)std::max(1, atoi(e)) * 1000000ull;
    if (const char* e = getenv("NTN_EXCHANGE_CHUNKS")) h->x_chunks = std::max(1, std::min(NTN_MAX_WORD_CHUNKS, atoi(e)));
    if (const char* e = getenv("NTN_EXCHANGE_BLOCKS")) h->x_blocks = std::max(1, std::min(NTN_XCHG_MAX_BLOCKS, atoi(e)));
    for (auto& o : h->xb) if (o.local == b.local) { o = b; ntn_drop_graphs(h); return NTN_OK; }
    h->xb.push_back(b);
    ntn_drop_graphs(h);
    return NTN_OK;
}

// Entity-gradient exchange: bind a symmetric allocation of >= Ne * dq floats (and its scratch of
// >= ntn_exchange_ge_scratch_numel() floats) as THE entity-gradient buffer of this handle.  Evaluations into a bound
// gradient buffer then sum gE over the ranks while the entity pull is running (chunks of entity rows, the pull kernel
// pushes the groups another rank owns) and the word pull that follows produces the full-batch word gradient everywhere;
// only the W/V/b/U blocks and the tail are exchanged after it.  Needs a prior ntn_exchange_bind (rank, world, mode).
extern "C" int64_t ntn_exchange_ge_scratch_numel(const ntn_handle* h) {
    return h ? (int64_t)h->dm.Ne * h->dm.dq + (int64_t)NTN_XCHG_RANGE_PAD * (NTN_MAX_WORD_CHUNKS + 1) : -1;
}
extern "C" int ntn_exchange_bind_ge(ntn_handle* h, const uint64_t* buffer_ptrs, const uint64_t* signal_pad_ptrs,
                                    uint64_t multicast_ptr, int64_t numel, const uint64_t* scratch_ptrs, int64_t scratch_numel) {
    if (!h) return NTN_EINVAL;
    if (h->xb.empty()) { h->err = "exchange: bind a gradient buffer first (ntn_exchange_bind)"; return NTN_ESTATE; }
    if (!buffer_ptrs || !signal_pad_ptrs) { h->err = "exchange: null pointer table"; return NTN_EINVAL; }
    if (numel < (int64_t)h->dm.Ne * h->dm.dq) { h->err = "exchange: gE buffer shorter than Ne * dq"; return NTN_EINVAL; }
    const bool want_scratch = h->x_mode == 2;
    if (want_scratch && (!scratch_ptrs || scratch_numel < ntn_exchange_ge_scratch_numel(h))) {
        h->err = "exchange: gE scratch missing or too short (ntn_exchange_ge_scratch_numel)"; return NTN_EINVAL;
    }
    if (h->x_mode == 1 && !multicast_ptr) { h->err = "exchange: multimem mode needs the multicast alias of the gE buffer"; return NTN_EINVAL; }
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    ExchangeBinding b;
    for (int r = 0; r < h->x_world; ++r) {
        if (!buffer_ptrs[r] || !signal_pad_ptrs[r] || (buffer_ptrs[r] & 15)) { h->err = "exchange: null or misaligned peer pointer"; return NTN_EINVAL; }
        b.peer[r] = reinterpret_cast<float*>(buffer_ptrs[r]);
        b.sig[r] = reinterpret_cast<uint32_t*>(signal_pad_ptrs[r]);
        if (want_scratch) {
            if (!scratch_ptrs[r] || (scratch_ptrs[r] & 15)) { h->err = "exchange: null or misaligned scratch pointer"; return NTN_EINVAL; }
            b.scratch[r] = reinterpret_cast<float*>(scratch_ptrs[r]);
        }
    }
    b.local = b.peer[h->x_rank];
    b.mc = reinterpret_cast<float*>(multicast_ptr);
    b.numel = numel;
    b.scratch_numel = want_scratch ? scratch_numel : 0;
    b.epoch = b.sig[h->x_rank] + NTN_XCHG_SIG_WORD0 - 2;
    if (!h->gE_own) h->gE_own = h->gE;
    h->gE = b.local;
    h->xge = b;
    h->xge_set = true;
    if (const char* e = getenv("NTN_EXCHANGE_GE_CHUNKS")) h->xge_chunks = std::max(1, std::min(NTN_MAX_WORD_CHUNKS, atoi(e)));
    ntn_drop_graphs(h);
    return NTN_OK;
}

extern "C" int ntn_exchange_unbind(ntn_handle* h) {
    if (!h) return NTN_EINVAL;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->gE_own) { h->gE = h->gE_own; h->gE_own = nullptr; }
    h->xge_set = false;
    h->xb.clear();
    ntn_drop_graphs(h);
    return NTN_OK;
}

// 1 when a flag wait of an exchange kernel expired since the last call (results of that evaluation are invalid)
extern "C" int ntn_exchange_check(ntn_handle* h) {
    if (!h || !h->d_xerr) return 0;
    unsigned v = 0;
    cudaMemcpyAsync(&v, h->d_xerr, sizeof(v), cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    if (v) cudaMemsetAsync(h->d_xerr, 0, sizeof(unsigned), h->stream);
    return v ? 1 : 0;
}

// Measurement hook: `iters` back-to-back exchange kernels over floats [off, off + n) of a bound buffer (not pushed by a
// producer: the kernel's own phase 0 stores the other ranks' slices), each ending with the end barrier when last != 0.
// Every rank must make the same call.  The buffer's content is summed over the ranks `iters` times.
extern "C" int ntn_exchange_probe(ntn_handle* h, float* d_grad, int64_t off, int64_t n, int32_t last, int32_t iters) {
    if (!h) return NTN_EINVAL;
    const ExchangeBinding* xb = ntn_find_binding(h, d_grad);
    if (!xb || (off & 3) || (n & 3) || off < 0 || off + n > xb->numel) { h->err = "exchange probe: unbound buffer or bad range"; return NTN_EINVAL; }
    cudaSetDevice(h->device);
    for (int i = 0; i < iters; ++i)
        ntn_launch_exchange(h, xb, XRange{off, n, 0, false}, XRange{0, 0, 0, false}, XRange{0, 0, 0, false}, last != 0, h->stream);
    return cudaGetLastError() == cudaSuccess ? NTN_OK : NTN_ECUDA;
}

extern "C" const char* ntn_exchange_mode(const ntn_handle* h) {
    if (!h || h->xb.empty()) return "none";
    return h->x_mode == 2 ? "push" : (h->x_mode == 1 ? "multimem" : "p2p");
}
