// s_energy_matrix -- the reference's nested-table class surface (src/s_energy_matrix.hh:16-68) on top of the
// device-resident tables.  The per-cell compute_* entry points of the reference are no-ops here: the whole
// fill is one bulk GPU sweep driven by W_final::ccj(); the getters read what that sweep left in HBM.
#ifndef CCJ_B200_S_ENERGY_MATRIX_HH
#define CCJ_B200_S_ENERGY_MATRIX_HH
#include <string>

#include "ccj_b200.h"
#include "ccj_types.h"

typedef int32_t energy_t;
typedef int32_t cand_pos_t;
#ifndef INF
#define INF CCJ_INF
#endif

struct free_energy_node {
    int energy;
    char type;
};

class W_final;

class s_energy_matrix {
public:
    s_energy_matrix(std::string seq, cand_pos_t length, ccj_ctx *ctx) : seq_(seq), n(length), ctx_(ctx) {}

    // bulk-filled on the GPU; kept for source compatibility (src/s_energy_matrix.cc:315-358,206-241)
    void compute_energy(cand_pos_t, cand_pos_t) {}
    void compute_WMv_WMp(cand_pos_t, cand_pos_t, energy_t) {}
    template <class T> void compute_energy_WM(cand_pos_t, cand_pos_t, T &) {}

    energy_t get_energy(cand_pos_t i, cand_pos_t j) { return (i >= j) ? INF : raw(T2_V, i, j); }
    energy_t get_energy_WM(cand_pos_t i, cand_pos_t j) { return (i >= j) ? INF : raw(T2_WM, i, j); }
    energy_t get_energy_WMv(cand_pos_t i, cand_pos_t j) { return (i >= j) ? INF : raw(T2_WMv, i, j); }
    energy_t get_energy_WMp(cand_pos_t i, cand_pos_t j) { return (i >= j) ? INF : raw(T2_WMp, i, j); }
    char get_type(cand_pos_t i, cand_pos_t j) { return (char)raw(T2_VTYPE, i, j); }
    free_energy_node get_node(cand_pos_t i, cand_pos_t j) { return free_energy_node{raw(T2_V, i, j), get_type(i, j)}; }

private:
    energy_t raw(int table, cand_pos_t i, cand_pos_t j) {
        int32_t v = 0;
        if (i < 1 || j > n || i > j || ccj_table2_get(ctx_, 0, table, i, j, &v) != 0) return INF;
        return v;
    }
    std::string seq_;
    cand_pos_t n;
    ccj_ctx *ctx_;
};
#endif
