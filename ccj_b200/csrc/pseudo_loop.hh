// pseudo_loop -- the reference's pseudoknot-table class surface (src/pseudo_loop.hh:13-56) on top of the
// device-resident tables: getters with Matrix4D::get / TriangleMatrix::get semantics; the fill and the
// traceback themselves run on the GPU (W_final::ccj()).
#ifndef CCJ_B200_PSEUDO_LOOP_HH
#define CCJ_B200_PSEUDO_LOOP_HH
#include <string>

#include "s_energy_matrix.hh"

class pseudo_loop {
public:
    pseudo_loop(std::string seq, s_energy_matrix *V, ccj_ctx *ctx) : n((cand_pos_t)seq.length()), V_(V), ctx_(ctx) {}

    // bulk-filled on the GPU; kept for source compatibility (src/pseudo_loop.cc:69-132)
    void compute_energies(cand_pos_t, cand_pos_t) {}

    // P(i,j): TriangleMatrix::get with return value INF (src/pseudo_loop.hh:32, src/matrices.hh:37-40)
    energy_t get_energy(cand_pos_t i, cand_pos_t j) { return (i > j) ? INF : raw2(T2_P, i, j); }
    // src/pseudo_loop.cc:647-661
    energy_t get_WB(cand_pos_t i, cand_pos_t j) { return wbwp(T2_WB, i, j); }
    energy_t get_WP(cand_pos_t i, cand_pos_t j) { return wbwp(T2_WP, i, j); }
    energy_t get_WBP(cand_pos_t i, cand_pos_t j) { return (i > j) ? INF : raw2(T2_WBP, i, j); }
    energy_t get_WPP(cand_pos_t i, cand_pos_t j) { return (i > j) ? INF : raw2(T2_WPP, i, j); }
    // any of the 22 gap tables (enum ccj_table4), Matrix4D::get semantics
    energy_t get_gap(int table, cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) {
        int32_t v = INF;
        ccj_table4_get(ctx_, 0, table, i, j, k, l, &v);
        return v;
    }
    energy_t get_PK(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) { return get_gap(T_PK, i, j, k, l); }
    energy_t get_PL(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) { return get_gap(T_PL, i, j, k, l); }
    energy_t get_PR(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) { return get_gap(T_PR, i, j, k, l); }
    energy_t get_PM(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) { return get_gap(T_PM, i, j, k, l); }
    energy_t get_PO(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) { return get_gap(T_PO, i, j, k, l); }
    // get_PfromMdoubleprime (src/pseudo_loop.cc:663-679); PB_penalty = 246 (src/h_globals.hh:11)
    energy_t get_PfromMdoubleprime(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l, bool il_can_pair) {
        if (!(i <= j && j < k - 1 && k <= l)) return INF;
        if (i == j && k == l) return il_can_pair ? 0 : INF;
        const energy_t b1 = get_PL(i, j, k, l) + 246, b2 = get_PR(i, j, k, l) + 246;
        return b1 < b2 ? b1 : b2;
    }

private:
    energy_t raw2(int table, cand_pos_t i, cand_pos_t j) {
        int32_t v = 0;
        if (i < 1 || j > n || ccj_table2_get(ctx_, 0, table, i, j, &v) != 0) return INF;
        return v;
    }
    energy_t wbwp(int table, cand_pos_t i, cand_pos_t j) {
        if (i <= 0 || j <= 0 || i > n || j > n) return INF;
        if (i > j) return 0;
        return raw2(table, i, j);
    }
    cand_pos_t n;
    s_energy_matrix *V_;
    ccj_ctx *ctx_;
};
#endif
