// See W_final.hh.  Replaces src/W_final.cc:20-105 (constructor + ccj()); the reference's error behaviour
// (messages, exit codes, "Should not be here!" lines) is re-created from the per-sequence ccj_result.
#include "W_final.hh"

#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "ccj_render.hpp"

int noGU = 0;
static std::string g_param_file;

int ccj_params_load(const char *par_file) {
    FILE *f = fopen(par_file, "r");
    if (!f) return 0;
    fclose(f);
    g_param_file = par_file;
    return 1;
}

void ccj_params_load_DNA_Mathews2004() { g_param_file = "@dna_mathews2004"; }

namespace {
// one context per process and device (CCJ_DEVICE selects the GPU); the reference is equally process-global
ccj_ctx *shared_ctx() {
    static ccj_ctx *ctx = nullptr;
    if (!ctx) {
        const char *d = getenv("CCJ_DEVICE");
        const int rc = ccj_ctx_create(d ? atoi(d) : 0, &ctx);
        if (rc != 0) {
            std::cerr << "ccj_b200: no CUDA device available (this build has no CPU path)" << std::endl;
            exit(EXIT_FAILURE);
        }
    }
    return ctx;
}
}  // namespace

W_final::W_final(std::string seq, int dangle) : params_(new ccj_params_view{g_param_file, dangle}), P(nullptr), V(nullptr) {
    seq_ = seq;
    n = (cand_pos_t)seq.length();
    ctx_ = shared_ctx();
    const int mrc = !g_param_file.empty() && g_param_file[0] == '@'
                        ? ccj_model_load_embedded(ctx_, g_param_file.c_str() + 1, dangle, noGU)
                        : ccj_model_load(ctx_, g_param_file.c_str(), dangle, noGU);
    if (mrc != 0) {
        std::cerr << "Not a valid parameter file!" << std::endl;  // src/CCJ.cc:84,95
        exit(EXIT_FAILURE);
    }
    W.resize(n + 1, 0);
    structure = std::string(n + 1, '.');
    V = new s_energy_matrix(seq_, n, ctx_);
    P = new pseudo_loop(seq_, V, ctx_);
}

W_final::~W_final() {
    delete P;
    delete V;
    delete params_;
}

double W_final::ccj() {
    const int64_t offsets[2] = {0, (int64_t)n};
    ccj_result res;
    pairs.assign(n, -1);
    std::string dots(n, '.');
    int rc = ccj_batch_prepare(ctx_, seq_.data(), offsets, 1);
    if (!rc) rc = ccj_batch_fill(ctx_);
    if (!rc) rc = ccj_batch_traceback(ctx_);
    if (!rc) rc = ccj_batch_fetch(ctx_, &res, pairs.data(), &dots[0]);
    if (rc != 0) {
        std::cerr << "ccj_b200: " << ccj_last_error(ctx_) << std::endl;
        exit(EXIT_FAILURE);
    }
    // the reference prints these from inside the traceback, before main() prints the result
    for (int x = 0; x < res.n_should_not_be_here; ++x) printf("Should not be here!\n");
    if (res.status == CCJ_EXIT_FAILURE) {
        fflush(stdout);
        std::cerr << ccj::traceback_message(res.msg_id) << std::endl;
        exit(EXIT_FAILURE);
    }
    if (res.status == CCJ_EXIT_ZERO_NOT_GOOD) {
        fflush(stdout);
        fprintf(stderr, "NOT GOOD RESTR INTER, i=%d, j=%d, best_ip=%d, best_jp=%d\n", res.aux_i, res.aux_j, res.aux_j,
                res.aux_i);
        exit(0);
    }
    if (res.status != CCJ_OK) {
        std::cerr << "ccj_b200: internal traceback error " << res.status << std::endl;
        exit(70);
    }
    W[n] = res.energy_dcal;
    structure = dots;
    return res.energy_dcal / 100.0;
}
