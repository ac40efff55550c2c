// s_energy_matrix / pseudo_loop method bodies (see the headers).  Every value comes from the bulk GPU fill through the
// host mirror of ccj_shell.hpp; the loop-energy helpers call the product's own host/device functions.
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "ccj_render.hpp"
#include "ccj_shell.hpp"
#include "pseudo_loop.hh"

// ---------------------------------------------------------------------------------------------- s_energy_matrix
s_energy_matrix::s_energy_matrix(std::string seq, cand_pos_t length, short *S, short *S1, vrna_param_t *params)
    : params_(params), S_(S), S1_(S1), seq_(seq), n(length) {
    fold_ = ccj::shell_fold(seq_, params);
}
s_energy_matrix::~s_energy_matrix() {}

energy_t s_energy_matrix::raw(int table, cand_pos_t i, cand_pos_t j) {
    if (i < 1 || j > n || i > j) return INF;
    return fold_->raw2(table, i, j);
}

free_energy_node *s_energy_matrix::get_node(cand_pos_t i, cand_pos_t j) {
    if (nodes_.empty()) nodes_.resize((size_t)(n + 1) * (n + 2) / 2 + 1);
    // same addressing as the reference's index[] (src/matrices.hh:69-75): row i starts after rows 1..i-1
    const size_t ij = (size_t)(i - 1) * (n + 1) - (size_t)(i - 1) * (i - 2) / 2 + (j - i);
    free_energy_node &nd = nodes_[ij];
    nd.energy = raw(T2_V, i, j);
    nd.type = (char)raw(T2_VTYPE, i, j);
    return &nd;
}

const ccj_model *s_energy_matrix::model_for(const paramT *params) {
    if (!params || params == params_) return &fold_->model;
    if (params != other_params_ || !other_model_) {
        other_model_.reset(new ccj_model());
        other_params_ = params;
    }
    ccj::model_from_vrna(*params, noGU, *other_model_);   // re-read every time: the caller may have edited it
    return other_model_.get();
}

energy_t s_energy_matrix::HairpinE(const std::string &seq, const short *, const short *, const paramT *params, cand_pos_t i,
                                   cand_pos_t j) {
    // the reference indexes the sequence it is handed; the encoding is the one of this object's sequence
    return ccj_HairpinE(model_for(params), fold_->S8.data(), seq.c_str(), i, j);
}

energy_t s_energy_matrix::compute_int(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l, const paramT *params) {
    return ccj_compute_int(model_for(params), fold_->S8.data(), i, j, k, l) + get_energy(k, l);
}

energy_t s_energy_matrix::compute_stack(cand_pos_t i, cand_pos_t j, const paramT *params) {
    return compute_int(i, j, i + 1, j - 1, params);
}

energy_t s_energy_matrix::compute_internal(cand_pos_t i, cand_pos_t j, const paramT *params) {
    energy_t v_iloop = INF;
    const cand_pos_t max_k = std::min(j - TURN - 2, i + MAXLOOP + 1);
    for (cand_pos_t k = i + 1; k <= max_k; ++k) {
        const cand_pos_t min_l = std::max(k + TURN + 1 + MAXLOOP + 2, k + j - i) - MAXLOOP - 2;
        for (cand_pos_t l = j - 1; l >= min_l; --l) v_iloop = std::min(v_iloop, compute_int(i, j, k, l, params));
    }
    return v_iloop;
}

energy_t s_energy_matrix::compute_energy_VM(cand_pos_t i, cand_pos_t j) {
    const ccj_cx c = fold_->cx();
    energy_t mn = INF;
    for (cand_pos_t k = i + 1; k <= j - 3; ++k) mn = std::min(mn, (energy_t)ccj_VM_term(c, i, j, k));
    return mn;
}

energy_t s_energy_matrix::E_MLStem(const energy_t &vij, const energy_t &vi1j, const energy_t &vij1, const energy_t &vi1j1,
                                   const short *, paramT *params, cand_pos_t i, cand_pos_t j, cand_pos_t) {
    ccj_cx c = fold_->cx();
    c.M = model_for(params);
    return ccj_E_MLStem(c, vij, vi1j, vij1, vi1j1, i, j);
}

energy_t s_energy_matrix::E_MbLoop(const energy_t WM2ij, const energy_t WM2ip1j, const energy_t WM2ijm1, const energy_t WM2ip1jm1,
                                   const short *, paramT *params, cand_pos_t i, cand_pos_t j) {
    ccj_cx c = fold_->cx();
    c.M = model_for(params);
    return ccj_E_MbLoop(c, WM2ij, WM2ip1j, WM2ijm1, WM2ip1jm1, i, j);
}

// -------------------------------------------------------------------------------------------------- pseudo_loop
pseudo_loop::pseudo_loop(std::string seq, s_energy_matrix *V, short *S, short *S1, vrna_param_t *params)
    : n((cand_pos_t)seq.length()), seq(seq), V(V), stack_interval(nullptr), f(nullptr), params_(params), S_(S), S1_(S1) {
    fold_ = ccj::shell_fold(seq, params);
    // TriangleMatrix::init(n+1, index) leaves return value INF for i>j (src/pseudo_loop.cc:32-34, src/matrices.hh:18-27)
    P.bind(fold_.get(), T2_P);
    WBP.bind(fold_.get(), T2_WBP);
    WPP.bind(fold_.get(), T2_WPP);
}
pseudo_loop::~pseudo_loop() {}

energy_t pseudo_loop::wbwp(int table, cand_pos_t i, cand_pos_t j) {
    if (i <= 0 || j <= 0 || i > n || j > n) return INF;
    if (i > j) return 0;
    return fold_->raw2(table, i, j);
}

energy_t pseudo_loop::get_gap(int table, cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) {
    return fold_->get4(table, i, j, k, l);
}

// get_PfromMdoubleprime (src/pseudo_loop.cc:663-679)
energy_t pseudo_loop::get_PfromMdoubleprime(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) {
    if (!(i <= j && j < k - 1 && k <= l)) return INF;
    if (i == j && k == l) return ccj_ptype(&fold_->model, fold_->S8.data(), i, l) == 0 ? INF : 0;
    const energy_t PB = fold_->model.PB_penalty;
    return std::min(get_gap(T_PL, i, j, k, l) + PB, get_gap(T_PR, i, j, k, l) + PB);
}

// the window getters read PL/PR/PM/PO of other cells: make those mirrors present, then run the product's cell function
#define CCJ_SHELL_VALID(i, j, k, l) if (!((i) <= (j) && (j) < (k)-1 && (k) <= (l))) return INF
energy_t pseudo_loop::get_PLiloop(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) {
    CCJ_SHELL_VALID(i, j, k, l);
    fold_->need4(T_PL);
    return ccj_PLiloop(fold_->cx(), i, j, k, l);
}
energy_t pseudo_loop::get_PRiloop(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) {
    CCJ_SHELL_VALID(i, j, k, l);
    fold_->need4(T_PR);
    return ccj_PRiloop(fold_->cx(), i, j, k, l);
}
energy_t pseudo_loop::get_PMiloop(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) {
    CCJ_SHELL_VALID(i, j, k, l);
    fold_->need4(T_PM);
    return ccj_PMiloop(fold_->cx(), i, j, k, l);
}
energy_t pseudo_loop::get_POiloop(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) {
    CCJ_SHELL_VALID(i, j, k, l);
    fold_->need4(T_PO);
    return ccj_POiloop(fold_->cx(), i, j, k, l);
}
energy_t pseudo_loop::get_PLmloop(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) {
    CCJ_SHELL_VALID(i, j, k, l);
    fold_->need4(T_PLmloop10);
    fold_->need4(T_PLmloop01);
    return ccj_PXmloop(fold_->cx(), T_PLmloop10, T_PLmloop01, i + 1, j - 1, k, l);
}
energy_t pseudo_loop::get_PRmloop(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) {
    CCJ_SHELL_VALID(i, j, k, l);
    fold_->need4(T_PRmloop10);
    fold_->need4(T_PRmloop01);
    return ccj_PXmloop(fold_->cx(), T_PRmloop10, T_PRmloop01, i, j, k + 1, l - 1);
}
energy_t pseudo_loop::get_PMmloop(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) {
    CCJ_SHELL_VALID(i, j, k, l);
    fold_->need4(T_PMmloop10);
    fold_->need4(T_PMmloop01);
    return ccj_PXmloop(fold_->cx(), T_PMmloop10, T_PMmloop01, i, j - 1, k + 1, l);
}
energy_t pseudo_loop::get_POmloop(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) {
    CCJ_SHELL_VALID(i, j, k, l);
    fold_->need4(T_POmloop10);
    fold_->need4(T_POmloop01);
    return ccj_PXmloop(fold_->cx(), T_POmloop10, T_POmloop01, i + 1, j, k, l - 1);
}
#undef CCJ_SHELL_VALID

void pseudo_loop::insert_node(int i, int j, int k, int l, char type) {
    seq_interval *tmp = new seq_interval;
    tmp->i = i;
    tmp->j = j;
    tmp->k = k;
    tmp->l = l;
    tmp->type = type;
    tmp->next = stack_interval;
    stack_interval = tmp;
}

void pseudo_loop::backtrack(minimum_fold *f, seq_interval *cur_interval) {
    this->f = f;
    fold_->ensure_resident();
    ccj_ctx *ctx = ccj::shell_ctx();
    const char ty = cur_interval->type;
    // 2-index nodes (insert_node(i,j,type), src/pseudo_loop.cc:2823-2833) leave k,l uninitialised in the reference
    const bool two = ty == P_P || ty == P_WB || ty == P_WBP || ty == P_WP || ty == P_WPP;
    const int32_t node[5] = {cur_interval->i, cur_interval->j, two ? 0 : cur_interval->k, two ? 0 : cur_interval->l, (int32_t)ty};
    std::vector<int32_t> pushed(5 * 64);
    int32_t np = 0, st[3] = {0, 0, 0};
    if (ccj_traceback_step(ctx, 0, node, pushed.data(), 64, &np, st) != 0) {
        std::cerr << "ccj_b200: " << ccj_last_error(ctx) << std::endl;
        exit(EXIT_FAILURE);
    }
    if (st[0] == CCJ_EXIT_FAILURE) {   // the reference prints its message and exits from inside backtrack
        std::cerr << ccj::traceback_message(st[1]) << std::endl;
        exit(EXIT_FAILURE);
    }
    for (int x = 0; x < np; ++x) {
        const int32_t *s = &pushed[5 * x];
        insert_node(s[0], s[1], s[2], s[3], (char)s[4]);
    }
    // pairs the node fixed (f[i].pair / f[i].type, e.g. src/pseudo_loop.cc:1046-1050)
    std::vector<int32_t> pr(n + 2, -1);
    std::vector<int8_t> tp(n + 2, 'N');
    if (ccj_fetch_fold_state(ctx, 0, pr.data(), tp.data()) != 0) {
        std::cerr << "ccj_b200: " << ccj_last_error(ctx) << std::endl;
        exit(EXIT_FAILURE);
    }
    for (int x = 1; x <= n; ++x)
        if (pr[x] > 0) {
            f[x].pair = pr[x];
            f[x].type = (char)tp[x];
        }
}
