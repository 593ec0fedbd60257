// C++ host layer above the C-ABI (include/ntn_b200.h): the reference's Java classes restated for the
// cost+gradient path -- java.util.Random, DataFactory (entity-name parsing, batch job, Socher text loaders)
// and the NTN function object -- so that a compiled host can drive libntn_b200.so the way Run_NTN.java does.
// Header-only; used by run_ntn.cpp (CLI) and host_capi.cpp (test hooks checked against the oracle).
#pragma once
#include <sys/stat.h>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/ntn_b200.h"

namespace ntnhost {

// java.util.Random, JDK specification (DataFactory.java:119,128; NTN.java:146 via Math.random)
class JavaRandom {
    uint64_t s_;
    static constexpr uint64_t kMult = 0x5DEECE66DULL, kAdd = 0xBULL, kMask = (1ULL << 48) - 1;
    int32_t next(int bits) {
        s_ = (s_ * kMult + kAdd) & kMask;
        return (int32_t)(s_ >> (48 - bits));
    }
public:
    explicit JavaRandom(int64_t seed) : s_(((uint64_t)seed ^ kMult) & kMask) {}
    int32_t nextInt() { return next(32); }
    int32_t nextInt(int32_t bound) {
        if (bound <= 0) throw std::invalid_argument("bound must be positive");
        int32_t r = next(31), m = bound - 1;
        if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
        for (int32_t u = r; (int32_t)((uint32_t)u - (uint32_t)(r = u % bound) + (uint32_t)m) < 0; u = next(31)) {}
        return r;
    }
    double nextDouble() { return (double)(((int64_t)next(26) << 27) + next(27)) * (1.0 / (double)(1LL << 53)); }
};

// String.split("_") with Java semantics: leading empties kept, trailing empties removed
inline std::vector<std::string> javaSplit(const std::string& s, char sep = '_') {
    std::vector<std::string> parts;
    size_t start = 0;
    for (size_t i = 0; i <= s.size(); ++i)
        if (i == s.size() || s[i] == sep) { parts.emplace_back(s.substr(start, i - start)); start = i + 1; }
    while (parts.size() > 1 && parts.back().empty()) parts.pop_back();
    if (parts.size() == 1 && parts[0].empty() && !s.empty()) parts.clear();
    return parts;
}

struct Tripple { int32_t e1, rel, e2, label; };      // Tripple.java:8-47 as plain indices

class DataFactory {
public:
    int batch_size, corrupt_size, embeddings_size;
    bool reduced_RelationSize;
    JavaRandom rand;
    JavaRandom shuffle_rand;                          // Collections.shuffle has its own Random in the JDK
    std::vector<std::string> entitiesNumWord, relationsNumWord, vocabNumWord;
    std::unordered_map<std::string, int> entitiesWordNum, relationsWordNum, vocabWordNum;
    std::vector<Tripple> trainingTripples, devTripples, testTripples;
    std::vector<float> wordVectorMaxtrixLoaded;       // d x Nw row-major
    std::vector<int32_t> bj_e1, bj_rel, bj_e2, bj_e3; // the batch job columns

    DataFactory(int batch, int corrupt, int d, bool reduced = false, int64_t seed = 42)
        : batch_size(batch), corrupt_size(corrupt), embeddings_size(d), reduced_RelationSize(reduced), rand(seed),
          shuffle_rand(seed + 1) {}

    void setEntities(const std::vector<std::string>& names) {
        entitiesNumWord = names; entitiesWordNum.clear();
        for (size_t i = 0; i < names.size(); ++i) entitiesWordNum[names[i]] = (int)i;
    }
    void setRelations(const std::vector<std::string>& names) {
        relationsNumWord = names; relationsWordNum.clear();
        for (size_t i = 0; i < names.size(); ++i) relationsWordNum[names[i]] = (int)i;
    }
    void setVocabulary(const std::vector<std::string>& words) {
        vocabNumWord = words; vocabWordNum.clear();
        for (size_t i = 0; i < words.size(); ++i) vocabWordNum[words[i]] = (int)i;
    }
    int getNumOfentities() const { return (int)entitiesNumWord.size(); }
    int getNumOfRelations() const { return (int)relationsNumWord.size(); }
    int getNumOfWords() const { return (int)vocabNumWord.size(); }

    static std::vector<std::string> readLines(const std::string& path) {
        std::ifstream f(path);
        if (!f) throw std::runtime_error("cannot open " + path);
        std::vector<std::string> out; std::string ln;
        while (std::getline(f, ln)) { if (!ln.empty() && ln.back() == '\r') ln.pop_back(); if (!ln.empty()) out.push_back(ln); }
        return out;
    }
    void loadEntitiesFromSocherFile(const std::string& p) { setEntities(readLines(p)); }                    // :153-187
    void loadRelationsFromSocherFile(const std::string& p) {                                               // :189-217
        auto v = readLines(p);
        if (reduced_RelationSize && v.size() > 2) v.resize(2);
        setRelations(v);
    }
    std::vector<Tripple> loadTriples(const std::string& p, bool labelled) const {                          // :219-325
        std::vector<Tripple> out;
        for (const auto& ln : readLines(p)) {
            std::istringstream ss(ln); std::string a, r, b; int lab = 0;
            ss >> a >> r >> b; if (labelled) ss >> lab;
            auto ir = relationsWordNum.find(r);
            if (ir == relationsWordNum.end()) { if (reduced_RelationSize) continue; throw std::runtime_error("unknown relation " + r); }
            out.push_back(Tripple{entitiesWordNum.at(a), ir->second, entitiesWordNum.at(b), lab});
        }
        return out;
    }
    void loadTrainingDataTripplesE1rE2(const std::string& p) { trainingTripples = loadTriples(p, false); }
    void loadDevDataTripplesE1rE2Label(const std::string& p) { devTripples = loadTriples(p, true); }
    void loadTestDataTripplesE1rE2Label(const std::string& p) { testTripples = loadTriples(p, true); }
    // word vectors as text: one line per word "word v0 v1 ... v{d-1}" (the .mat reader needs jmatio; the
    // Python host reads initEmbed.mat with scipy and can export this format)
    void loadWordVectorsFromTextFile(const std::string& p) {
        auto lines = readLines(p);
        std::vector<std::string> words; std::vector<std::vector<float>> vecs;
        for (const auto& ln : lines) {
            std::istringstream ss(ln); std::string w; ss >> w; std::vector<float> v(embeddings_size);
            for (auto& x : v) ss >> x;
            words.push_back(w); vecs.push_back(v);
        }
        setVocabulary(words);
        const int Nw = (int)words.size();
        wordVectorMaxtrixLoaded.assign((size_t)embeddings_size * Nw, 0.f);
        for (int w = 0; w < Nw; ++w) for (int p2 = 0; p2 < embeddings_size; ++p2) wordVectorMaxtrixLoaded[(size_t)p2 * Nw + w] = vecs[w][p2];
    }

    // ---- entity names -> words (DataFactory.java:397-434,583-595,646-684) ----
    static std::string clearName(const std::string& raw) {
        size_t last = raw.rfind('_');
        if (last != std::string::npos && last >= 2) return raw.substr(2, last - 2);
        if (raw.size() < 2) throw std::runtime_error("entity name too short: " + raw);
        return raw.substr(2);
    }
    int wordNum(const std::string& w) const {
        auto it = vocabWordNum.find(w);
        return it != vocabWordNum.end() ? it->second : vocabWordNum.at("unknown");
    }
    static std::vector<std::string> forwardWords(const std::string& raw) {
        std::string name = clearName(raw);
        if (name.find('_') != std::string::npos) return javaSplit(name);
        return {name};
    }
    int entityLength(int i) const { return (int)javaSplit(entitiesNumWord[i]).size() - 3; }
    std::vector<int32_t> getWordIndexes(int i) const {
        int n = entityLength(i);
        std::vector<int32_t> out((size_t)(n != 0 ? n : 1), 0);
        auto words = forwardWords(entitiesNumWord[i]);
        if (words.size() > out.size()) throw std::out_of_range("reference would throw ArrayIndexOutOfBounds for " + entitiesNumWord[i]);
        for (size_t j = 0; j < words.size(); ++j) out[j] = wordNum(words[j]);
        return out;
    }
    void entityWordMaps(std::vector<int32_t>& fo, std::vector<int32_t>& fi, std::vector<int32_t>& bo, std::vector<int32_t>& bi,
                        std::vector<int32_t>& lb) const {
        fo.assign(1, 0); bo.assign(1, 0); fi.clear(); bi.clear(); lb.clear();
        for (int i = 0; i < getNumOfentities(); ++i) {
            for (const auto& w : forwardWords(entitiesNumWord[i])) fi.push_back(wordNum(w));
            fo.push_back((int32_t)fi.size());
            for (int32_t w : getWordIndexes(i)) bi.push_back(w);
            bo.push_back((int32_t)bi.size());
            lb.push_back(entityLength(i));
        }
    }

    // Collections.shuffle(trainingTripples): the line commented out at DataFactory.java:118, JDK algorithm
    // (for i = size..2: swap(i-1, rnd.nextInt(i))), seeded stream
    void shuffleTrainingTripples() {
        for (int i = (int)trainingTripples.size(); i > 1; --i) std::swap(trainingTripples[i - 1], trainingTripples[shuffle_rand.nextInt(i)]);
    }

    // ---- batch job (DataFactory.java:116-136): h-major, first batch_size training triples ----
    void generateNewTrainingBatchJob(bool shuffle = false) {
        if (shuffle) shuffleTrainingTripples();
        const int n = std::min<int>(batch_size, (int)trainingTripples.size());
        bj_e1.clear(); bj_rel.clear(); bj_e2.clear(); bj_e3.clear();
        for (int h = 0; h < corrupt_size; ++h)
            for (int i = 0; i < n; ++i) {
                int32_t e3 = rand.nextInt(getNumOfentities());
                bj_e1.push_back(trainingTripples[i].e1); bj_rel.push_back(trainingTripples[i].rel);
                bj_e2.push_back(trainingTripples[i].e2); bj_e3.push_back(e3);
            }
    }
};

// The NTN function object (NTN.java:19-84, 92-448) over the C-ABI.
// ---- theta checkpoints (Run_NTN.java:133-139: Nd4j.writeTxt(theta, path, ","); resume: the commented :108) ----------
// Text: one comma-separated line of P floats, 9 significant digits (every FP32 value round-trips bit for bit).
// Binary fast path: "NTNTHETA" | uint32 version = 1 | uint32 dtype = 32 | int64 P | P little-endian float32.
inline void writeThetaTxt(const std::vector<double>& th, const std::string& path) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) throw std::runtime_error("cannot write " + path);
    for (size_t i = 0; i < th.size(); ++i) fprintf(f, i ? ",%.9g" : "%.9g", (double)(float)th[i]);
    fclose(f);
}
inline void writeThetaBin(const std::vector<double>& th, const std::string& path) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + path);
    const uint32_t hdr[2] = {1u, 32u};
    const int64_t P = (int64_t)th.size();
    std::vector<float> v(th.begin(), th.end());
    fwrite("NTNTHETA", 1, 8, f); fwrite(hdr, 4, 2, f); fwrite(&P, 8, 1, f); fwrite(v.data(), 4, v.size(), f);
    fclose(f);
}
// reads either format (the binary one is recognised by its magic); returns false if the file does not exist
inline bool readTheta(const std::string& path, std::vector<double>& th) {
    struct stat sb;
    if (stat(path.c_str(), &sb) != 0 || !S_ISREG(sb.st_mode)) return false;     // missing, or a directory
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char magic[8] = {0};
    size_t got = fread(magic, 1, 8, f);
    if (got == 8 && std::string(magic, 8) == "NTNTHETA") {
        uint32_t hdr[2]; int64_t P = 0;
        if (fread(hdr, 4, 2, f) != 2 || fread(&P, 8, 1, f) != 1 || hdr[0] != 1u || hdr[1] != 32u || P < 0) { fclose(f); throw std::runtime_error("bad theta header in " + path); }
        std::vector<float> v((size_t)P);
        if (fread(v.data(), 4, (size_t)P, f) != (size_t)P) { fclose(f); throw std::runtime_error("truncated theta file " + path); }
        th.assign(v.begin(), v.end());
    } else {
        fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
        std::string buf((size_t)n, '\0');
        if (fread(&buf[0], 1, (size_t)n, f) != (size_t)n) { fclose(f); throw std::runtime_error("cannot read " + path); }
        th.clear();
        const char* p = buf.c_str();
        while (*p) {
            char* e = nullptr;
            double v = strtod(p, &e);
            if (e == p) break;
            th.push_back((double)(float)v);
            p = e;
            while (*p == ',' || *p == ' ' || *p == '\n' || *p == '\r') ++p;
        }
    }
    fclose(f);
    return true;
}

class NTN {
    ntn_handle* h_ = nullptr;
    DataFactory* tbj_;
    JavaRandom side_rand_;
    int R_;
    bool train_uploaded_ = false, present_valid_ = false;
    size_t present_n_ = 0;
    std::vector<char> present_;
    void check(int rc) const { if (rc != NTN_OK) throw std::runtime_error(ntn_last_error(h_)); }
public:
    int64_t update = 0;
    NTN(int d, int Ne, int R, int Nw, int batchSize, int k, const std::string& activation, DataFactory* tbj, float lamda,
        int wv_source = NTN_WV_FROZEN, int64_t side_seed = 7)
        : tbj_(tbj), side_rand_(side_seed), R_(R) {
        ntn_config c{};
        c.d = d; c.k = k; c.R = R; c.Ne = Ne; c.Nw = Nw; c.batch_divisor = batchSize; c.lambda = lamda;
        c.activation = activation == "tanh" ? NTN_ACT_TANH : NTN_ACT_SIGMOID;
        c.wv_source = wv_source; c.gemm_path = NTN_GEMM_AUTO; c.device = 0; c.reg_scale = 1.f;
        int rc = ntn_create(&c, &h_);
        if (rc != NTN_OK) throw std::runtime_error(ntn_last_error(nullptr));
        std::vector<int32_t> fo, fi, bo, bi, lb;
        tbj->entityWordMaps(fo, fi, bo, bi, lb);
        check(ntn_set_entity_words(h_, fo.data(), fi.data(), bo.data(), bi.data(), lb.data()));
        if (wv_source == NTN_WV_FROZEN) check(ntn_set_frozen_wv(h_, tbj->wordVectorMaxtrixLoaded.data()));
    }
    ~NTN() { ntn_destroy(h_); }
    ntn_handle* handle() { return h_; }
    int getDimension() const { return (int)ntn_dimension(h_); }                  // NTN.java:445-448
    void syncBatch() { present_valid_ = false; check(ntn_set_batch(h_, (int64_t)tbj_->bj_e1.size(), tbj_->bj_e1.data(), tbj_->bj_rel.data(), tbj_->bj_e2.data(), tbj_->bj_e3.data())); }
    // sampler variant: batch job generated and indexed in HBM (ntn_generate_batch_dev); only the relation column of the
    // first batch_size training triples is kept on the host (for drawCorruptSides)
    void generateBatchJobOnDevice(uint64_t seed, uint64_t job, bool filter_true) {
        const auto& tr = tbj_->trainingTripples;
        if (!train_uploaded_) {
            std::vector<int32_t> e1, rel, e2;
            for (auto& x : tr) { e1.push_back(x.e1); rel.push_back(x.rel); e2.push_back(x.e2); }
            check(ntn_set_training_triples(h_, (int64_t)tr.size(), e1.data(), rel.data(), e2.data()));
            train_uploaded_ = true;
        }
        const int n = std::min<int>(tbj_->batch_size, (int)tr.size());
        check(ntn_generate_batch_dev(h_, 0, n, tbj_->corrupt_size, seed, job, filter_true ? 1 : 0));
        tbj_->bj_e1.clear(); tbj_->bj_e2.clear(); tbj_->bj_e3.clear(); tbj_->bj_rel.clear();
        for (int i = 0; i < n; ++i) tbj_->bj_rel.push_back(tr[i].rel);
        present_valid_ = false;
    }
    // one Math.random() > 0.5 per non-empty relation, increasing r (NTN.java:120,146)
    void drawCorruptSides(uint8_t* side) {
        if (!present_valid_ || present_n_ != tbj_->bj_rel.size()) {
            present_.assign(R_, 0);
            for (int32_t r : tbj_->bj_rel) present_[r] = 1;
            present_valid_ = true; present_n_ = tbj_->bj_rel.size();
        }
        const std::vector<char>& present = present_;
        for (int r = 0; r < R_; ++r) side[r] = present[r] ? (side_rand_.nextDouble() > 0.5 ? 1 : 0) : 0;
    }
    // IPair<Double,double[]> computeAt(double[] theta)  (NTN.java:92-443)
    double computeAt(const double* theta, double* grad) {
        std::vector<uint8_t> side(R_);
        drawCorruptSides(side.data());
        double cost; int64_t nact;
        check(ntn_eval(h_, theta, side.data(), &cost, grad, &nact));
        update += nact;
        return cost;
    }
    static void sidesThunk(void* ctx, uint8_t* out) { static_cast<NTN*>(ctx)->drawCorruptSides(out); }
};

}  // namespace ntnhost
