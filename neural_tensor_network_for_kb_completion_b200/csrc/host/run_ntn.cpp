// run_ntn: the reference's driver (/root/reference/src/Run_NTN.java:38-150) as a compiled host over the C-ABI.
//   run_ntn <data_path> <theta_load_path> <theta_save_path> [numIterations] [sampler]
// sampler: "reference" (default: host java.util.Random stream, first batch_size triples, DataFactory.java:116-136),
// "shuffle" (re-enables the Collections.shuffle commented out at :118), "device" / "device-filter" (batch job generated
// and indexed in HBM by the counter-based sampler, ntn_generate_batch_dev; with "-filter" known true triples are redrawn).
// theta_load_path: a checkpoint to resume from (text as written below, or its .bin twin); ignored if it does not exist.
// data_path holds entities.txt, relations.txt, train.txt, dev.txt, test.txt (Socher layout) and wordvectors.txt
// ("word v0 ... v{d-1}" per line; export of initEmbed.mat).  Training runs the device L-BFGS with maxIters = 5
// per batch job (Run_NTN.java:119-124), checkpoints theta as comma-separated text (:133-139), then thresholds and
// accuracy (:145-149) with the reference's eval formula and sweep (NTN.java:587-736).
#include <algorithm>
#include <cstdio>
#include <ctime>
#include <cuda_runtime.h>

#include "ntn_host.hpp"

using namespace ntnhost;

// text checkpoint in the reference's format + the binary fast path beside it (same name, .bin)
static void writeTheta(const std::vector<double>& th, const std::string& path) {
    writeThetaTxt(th, path);
    writeThetaBin(th, path.substr(0, path.size() - 4) + ".bin");
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: run_ntn <data_path> <theta_load_path> <theta_save_path> [numIterations]\n"); return 2; }
    const std::string data_path = argv[1], theta_load_path = argv[2], theta_save_path = argv[3];
    // Run_NTN.java:72-83
    const int batchSize = 20000, numWVdimensions = 100, batch_iterations = 5, sliceSize = 3, corrupt_size = 10;
    const int numIterations = argc > 4 ? atoi(argv[4]) : 500;
    const float lamda = 0.0001f;
    const std::string sampler = argc > 5 ? argv[5] : "reference";
    if (sampler != "reference" && sampler != "shuffle" && sampler != "device" && sampler != "device-filter") {
        fprintf(stderr, "run_ntn: unknown sampler '%s'\n", sampler.c_str());
        return 2;
    }
    try {
        DataFactory tbj(batchSize, corrupt_size, numWVdimensions);
        tbj.loadEntitiesFromSocherFile(data_path + "/entities.txt");
        tbj.loadRelationsFromSocherFile(data_path + "/relations.txt");
        tbj.loadTrainingDataTripplesE1rE2(data_path + "/train.txt");
        tbj.loadDevDataTripplesE1rE2Label(data_path + "/dev.txt");
        tbj.loadTestDataTripplesE1rE2Label(data_path + "/test.txt");
        tbj.loadWordVectorsFromTextFile(data_path + "/wordvectors.txt");
        const int R = tbj.getNumOfRelations(), Ne = tbj.getNumOfentities(), Nw = tbj.getNumOfWords(), d = numWVdimensions, k = sliceSize;
        NTN t(d, Ne, R, Nw, batchSize, k, "tanh", &tbj, lamda);
        const int64_t P = t.getDimension();
        // theta_inital (NTN.java:67-82): W = rand * 1/sqrt(2d), V = 0.4, b = 0.01, U = 0.8, word block = loaded vectors
        std::vector<double> theta((size_t)P);
        JavaRandom init(3);
        const double r = 1.0 / std::sqrt(2.0 * d);
        size_t o = 0;
        for (int64_t i = 0; i < (int64_t)R * k * d * d; ++i) theta[o++] = init.nextDouble() * r;
        for (int64_t i = 0; i < (int64_t)R * 2 * d * k; ++i) theta[o++] = 0.4;
        for (int i = 0; i < R * k; ++i) theta[o++] = 0.01;
        for (int i = 0; i < R * k; ++i) theta[o++] = 0.8;
        for (float v : tbj.wordVectorMaxtrixLoaded) theta[o++] = v;
        // resume (the reference parses theta_load_path but only has the load commented out, Run_NTN.java:64,108):
        // a text or binary checkpoint of the right dimension replaces the initial theta
        {
            std::vector<double> loaded;
            if (readTheta(theta_load_path, loaded)) {
                if ((int64_t)loaded.size() != P) throw std::runtime_error("theta_load_path holds " + std::to_string(loaded.size()) + " values, the model has " + std::to_string(P));
                theta = loaded;
                printf("theta loaded from %s\n", theta_load_path.c_str());
            }
        }
        float* d_theta = nullptr;
        if (cudaMalloc((void**)&d_theta, (size_t)ntn_internal_dimension(t.handle()) * sizeof(float)) != cudaSuccess) throw std::runtime_error("cudaMalloc");
        if (ntn_upload_theta(t.handle(), theta.data(), d_theta)) throw std::runtime_error(ntn_last_error(t.handle()));
        for (int i = 0; i < numIterations; ++i) {                                   // Run_NTN.java:113
            if (sampler == "device" || sampler == "device-filter") {
                t.generateBatchJobOnDevice(/*seed*/ 42, /*job*/ (uint64_t)i, sampler == "device-filter");
            } else {
                tbj.generateNewTrainingBatchJob(sampler == "shuffle");                  // :115
                t.syncBatch();
            }
            ntn_lbfgs_opts opts{batch_iterations, 0, 0.0};
            ntn_lbfgs_result res{};
            if (ntn_lbfgs_minimize(t.handle(), d_theta, &opts, &NTN::sidesThunk, &t, &res)) throw std::runtime_error(ntn_last_error(t.handle()));
            printf("res: %s| %.9g\nParamters for batchjob optimized, iteration: %d completed\n", res.did_converge ? "true" : "false", res.min_obj_val, i);
            if (i == 5 || i % 10 == 0) {                                            // :133-135
                ntn_download_theta(t.handle(), d_theta, theta.data());
                writeTheta(theta, theta_save_path + "/theta_opt_iteration_" + std::to_string(i) + ".txt");
            }
        }
        ntn_download_theta(t.handle(), d_theta, theta.data());
        time_t now = time(nullptr);
        writeTheta(theta, theta_save_path + "/theta_opt" + std::to_string(localtime(&now)->tm_mday) + ".txt");   // :139
        printf("Model saved!\n");
        // thresholds on dev (NTN.java:587-669, quirks of App. B7 kept) and accuracy on test (:672-712)
        auto scores = [&](const std::vector<Tripple>& ts) {
            std::vector<int32_t> e1, rel, e2; for (auto& x : ts) { e1.push_back(x.e1); rel.push_back(x.rel); e2.push_back(x.e2); }
            std::vector<double> s(ts.size());
            if (ntn_score(t.handle(), theta.data(), (int64_t)ts.size(), e1.data(), rel.data(), e2.data(), NTN_SCORE_REFERENCE_EVAL, s.data()))
                throw std::runtime_error(ntn_last_error(t.handle()));
            return s;
        };
        std::vector<double> sd = scores(tbj.devTripples), thr((size_t)R, -1.0), acc((size_t)R, 0.0);
        if (!sd.empty()) {
            double smin = *std::min_element(sd.begin(), sd.end()), smax = *std::max_element(sd.begin(), sd.end());
            for (double st = smin; st <= smax;)
                for (int i = 0; i < R; ++i) {
                    int n = 0, ok = 0;
                    for (size_t j = 0; j < sd.size(); ++j) if (tbj.devTripples[j].rel == i) { ++n; ok += (sd[j] <= st) ? (tbj.devTripples[j].label == 1) : (tbj.devTripples[j].label == -1); }
                    if (n && (double)ok / n > acc[i]) { acc[i] = (double)ok / n; thr[i] = st; }
                    st += 0.01;
                }
        }
        std::vector<double> stt = scores(tbj.testTripples);
        size_t good = 0;
        for (size_t j = 0; j < stt.size(); ++j) good += ((stt[j] <= thr[tbj.testTripples[j].rel]) ? 1 : -1) == tbj.testTripples[j].label;
        printf("Accuracy of predictions: %.6f\n", stt.empty() ? 0.0 : (double)good / stt.size());
        cudaFree(d_theta);
    } catch (const std::exception& e) { fprintf(stderr, "run_ntn: %s\n", e.what()); return 1; }
    return 0;
}
