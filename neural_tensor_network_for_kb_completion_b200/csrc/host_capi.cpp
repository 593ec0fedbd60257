// C-ABI test hooks over the C++ host layer (csrc/host/ntn_host.hpp): the CPU test-suite checks these
// bit-exactly against the oracle (java.util.Random stream, entity/word maps, batch job order).
#include <algorithm>
#include <cstring>

#include "host/ntn_host.hpp"

using namespace ntnhost;

extern "C" {

// n draws of nextInt(bound) (bound > 0) or nextInt() (bound == 0), then m draws of nextDouble()
int ntn_host_java_random(int64_t seed, int32_t bound, int64_t n, int32_t* ints, int64_t m, double* doubles) {
    try {
        JavaRandom r(seed);
        for (int64_t i = 0; i < n; ++i) ints[i] = bound > 0 ? r.nextInt(bound) : r.nextInt();
        for (int64_t i = 0; i < m; ++i) doubles[i] = r.nextDouble();
        return NTN_OK;
    } catch (...) { return NTN_EINVAL; }
}

// A native ntn_sides_fn for hosts that are not compiled code (the Python mirror): one Math.random() > 0.5 per non-empty
// relation, increasing r (NTN.java:120,146), on the 48-bit java.util.Random state the host hands over -- so a device
// L-BFGS run draws its corrupt sides without calling back into the interpreter.  state[0] is the LCG state (as JavaRandom
// holds it; read back after the run to keep the host's stream in step), state[1] = R, then R presence flags.
void ntn_host_sides_draw(void* ctx, uint8_t* side) {
    uint64_t* st = static_cast<uint64_t*>(ctx);
    const uint64_t kMult = 0x5DEECE66DULL, kAdd = 0xBULL, kMask = (1ULL << 48) - 1;
    uint64_t s = st[0];
    const int R = (int)st[1];
    for (int r = 0; r < R; ++r) {
        if (!st[2 + r]) { side[r] = 0; continue; }
        s = (s * kMult + kAdd) & kMask; const int64_t hi = (int64_t)(s >> 22);      // next(26)
        s = (s * kMult + kAdd) & kMask; const int64_t lo = (int64_t)(s >> 21);      // next(27)
        const double v = (double)((hi << 27) + lo) * (1.0 / (double)(1LL << 53));
        side[r] = v > 0.5 ? 1 : 0;
    }
    st[0] = s;
}

// n consecutive nextInt(bound) draws continuing from *state (the 48-bit LCG state, as JavaRandom holds it); the advanced
// state is written back.  Lets the Python DataFactory mirror draw 200k corrupt entities at native speed.
int ntn_host_random_ints(uint64_t* state, int32_t bound, int64_t n, int32_t* out) {
    if (!state || !out || bound <= 0) return NTN_EINVAL;
    const uint64_t kMult = 0x5DEECE66DULL, kAdd = 0xBULL, kMask = (1ULL << 48) - 1;
    uint64_t s = *state;
    const int32_t m = bound - 1;
    const bool pow2 = (bound & m) == 0;
    for (int64_t i = 0; i < n; ++i) {
        for (;;) {
            s = (s * kMult + kAdd) & kMask;
            const int32_t u = (int32_t)(s >> 17);
            if (pow2) { out[i] = (int32_t)(((int64_t)bound * (int64_t)u) >> 31); break; }
            const int32_t r = u % bound;
            if ((int32_t)((uint32_t)u - (uint32_t)r + (uint32_t)m) >= 0) { out[i] = r; break; }
        }
    }
    *state = s;
    return NTN_OK;
}

// names / words: '\n'-separated.  Outputs sized by the caller: fwd_off, bwd_off (Ne+1), len_bwd (Ne),
// fwd_idx / bwd_idx (cap entries each); *nf, *nb receive the used counts.
int ntn_host_entity_word_maps(const char* names, const char* words, int32_t* fwd_off, int32_t* fwd_idx, int32_t* bwd_off,
                              int32_t* bwd_idx, int32_t* len_bwd, int64_t cap, int64_t* nf, int64_t* nb) {
    try {
        auto split = [](const char* s) { std::vector<std::string> v; std::string cur;
            for (const char* p = s; *p; ++p) { if (*p == '\n') { v.push_back(cur); cur.clear(); } else cur.push_back(*p); }
            if (!cur.empty()) v.push_back(cur);
            return v; };
        DataFactory f(1, 1, 1);
        f.setEntities(split(names)); f.setVocabulary(split(words));
        std::vector<int32_t> fo, fi, bo, bi, lb;
        f.entityWordMaps(fo, fi, bo, bi, lb);
        if ((int64_t)fi.size() > cap || (int64_t)bi.size() > cap) return NTN_EINVAL;
        memcpy(fwd_off, fo.data(), fo.size() * 4); memcpy(bwd_off, bo.data(), bo.size() * 4); memcpy(len_bwd, lb.data(), lb.size() * 4);
        memcpy(fwd_idx, fi.data(), fi.size() * 4); memcpy(bwd_idx, bi.data(), bi.size() * 4);
        *nf = (int64_t)fi.size(); *nb = (int64_t)bi.size();
        return NTN_OK;
    } catch (...) { return NTN_EINVAL; }
}

// DataFactory.generateNewTrainingBatchJob over `ntrain` training triples (e1, rel, e2 interleaved), called `calls` times;
// the columns of the LAST call are written (batch*corrupt entries each).
int ntn_host_generate_batch(const int32_t* train, int64_t ntrain, int32_t batch, int32_t corrupt, int32_t num_entities,
                            int64_t seed, int32_t calls, int32_t* e1, int32_t* rel, int32_t* e2, int32_t* e3) {
    try {
        DataFactory f(batch, corrupt, 1, false, seed);
        std::vector<std::string> names((size_t)num_entities, "__x_1");
        f.entitiesNumWord = names;
        for (int64_t i = 0; i < ntrain; ++i) f.trainingTripples.push_back(Tripple{train[3 * i], train[3 * i + 1], train[3 * i + 2], 0});
        for (int c = 0; c < calls; ++c) f.generateNewTrainingBatchJob();
        size_t n = f.bj_e1.size();
        memcpy(e1, f.bj_e1.data(), n * 4); memcpy(rel, f.bj_rel.data(), n * 4); memcpy(e2, f.bj_e2.data(), n * 4); memcpy(e3, f.bj_e3.data(), n * 4);
        return NTN_OK;
    } catch (...) { return NTN_EINVAL; }
}


// java.util.Collections.shuffle(list, new Random(seed)) of the list [0, n): perm[p] = element now at position p
int ntn_host_collections_shuffle(int64_t seed, int32_t n, int32_t* perm) {
    if (n < 0 || (n > 0 && !perm)) return NTN_EINVAL;
    JavaRandom r(seed);
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int i = n; i > 1; --i) std::swap(perm[i - 1], perm[r.nextInt(i)]);
    return NTN_OK;
}

// theta checkpoints of the C++ host (csrc/host/ntn_host.hpp; Run_NTN.java:133-139 and the resume of :108):
// write n values as text (binary != 0: the binary fast path) to `path`, read them back into `back`
int ntn_host_theta_checkpoint_roundtrip(const char* path, int32_t binary, const double* theta, int64_t n, double* back) {
    if (!path || n < 0 || (n > 0 && (!theta || !back))) return NTN_EINVAL;
    try {
        std::vector<double> th(theta, theta + n), got;
        if (binary) writeThetaBin(th, path); else writeThetaTxt(th, path);
        if (!readTheta(path, got) || (int64_t)got.size() != n) return NTN_EINVAL;
        for (int64_t i = 0; i < n; ++i) back[i] = got[(size_t)i];
        return NTN_OK;
    } catch (...) { return NTN_EINVAL; }
}
}  // extern "C"
