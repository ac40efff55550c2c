// See W_final.hh.  Replaces src/W_final.cc:20-105 (constructor + ccj()); the reference's error behaviour
// (messages, exit codes, "Should not be here!" lines) is re-created from the per-sequence ccj_result.
#include "W_final.hh"

#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "ccj_render.hpp"
#include "ccj_shell.hpp"

int ccj_params_load(const char *par_file) { return vrna_params_load(par_file, VRNA_PARAMETER_FORMAT_DEFAULT); }
void ccj_params_load_DNA_Mathews2004() { vrna_params_load_DNA_Mathews2004(); }

// src/W_final.cc:20-30,45-56
W_final::W_final(std::string seq, int dangle) : params_(scale_parameters()), P(nullptr), V(nullptr) {
    if (!params_) {
        std::cerr << "Not a valid parameter file!" << std::endl;  // src/CCJ.cc:84,95
        exit(EXIT_FAILURE);
    }
    seq_ = seq;
    n = (cand_pos_t)seq.length();
    make_pair_matrix();
    params_->model_details.dangles = dangle;
    S_ = encode_sequence(seq.c_str(), 0);
    S1_ = encode_sequence(seq.c_str(), 1);
    W.resize(n + 1, 0);
    structure = std::string(n + 1, '.');
    V = new s_energy_matrix(seq_, n, S_, S1_, params_);
    P = new pseudo_loop(seq_, V, S_, S1_, params_);
}

W_final::~W_final() {
    delete P;
    delete V;
    free(params_);
    free(S_);
    free(S1_);
}

double W_final::ccj() {
    ccj_ctx *ctx = ccj::shell_ctx();
    ccj_result res;
    pairs.assign(n, -1);
    std::string dots(n, '.');
    int rc;
    const char *force = getenv("CCJ_FORCE_SHARD");   // testing aid: take the multi-GPU path for any length
    if (n > 0 && ccj_device_count() > 1 && (ccj_wave_capacity(ctx, n) < 1 || (force && force[0] == '1'))) {
        // the tables of this sequence exceed one GPU: deal the rows of the gap tables to every GPU of the box
        // (ccj_shard_fold; getters of P / V are not available for such a fold)
        const int ndev = ccj_device_count();
        std::vector<ccj_ctx *> ctxs;
        for (int d = 0; d < ndev; ++d) {
            ccj_ctx *c = nullptr;
            if (ccj_ctx_create(d, &c) != 0 || ccj_model_upload(c, &V->fold()->model, sizeof(ccj_model)) != 0) {
                std::cerr << "ccj_b200: cannot use GPU " << d << std::endl;
                exit(EXIT_FAILURE);
            }
            ctxs.push_back(c);
        }
        rc = ccj_shard_fold(ctxs.data(), ndev, seq_.data(), n, &res, pairs.data(), &dots[0], nullptr);
        for (ccj_ctx *c : ctxs) ccj_ctx_destroy(c);
        if (rc != 0) {
            std::cerr << "ccj_b200: the sequence does not fit the GPUs of this box" << std::endl;
            exit(EXIT_FAILURE);
        }
    } else {
        V->fold()->ensure_resident();   // the bulk fill (V, P, the 22 gap tables, W), src/W_final.cc:60-77
        rc = ccj_batch_traceback(ctx);
        if (!rc) rc = ccj_batch_fetch(ctx, &res, pairs.data(), &dots[0]);
        if (rc != 0) {
            std::cerr << "ccj_b200: " << ccj_last_error(ctx) << std::endl;
            exit(EXIT_FAILURE);
        }
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
