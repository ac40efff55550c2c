// Internal glue of the C++ class shells (W_final / pseudo_loop / s_energy_matrix): one "ShellFold" per
// (sequence, energy model) holds the host mirror the getters read -- the model blob, the encoded sequence, the 2D
// tables and, table by table on first use, the gap tables of the bulk GPU fill.  Everything reaches the GPU through
// the C ABI of include/ccj_b200.h only.
#ifndef CCJ_B200_SHELL_HPP
#define CCJ_B200_SHELL_HPP
#include <memory>
#include <string>
#include <vector>

// the product's cell functions, compiled for the host here (CCJ_HD = inline); they use INF as a local name, which the
// reference-compatible macro of ccj_compat.hh would clobber
#pragma push_macro("INF")
#undef INF
#include "ccj_cells4.cuh"
#pragma pop_macro("INF")
#include "ccj_compat.hh"

namespace ccj {

ccj_ctx *shell_ctx();   // one context per process and device (CCJ_DEVICE selects the GPU), like the reference's globals

// vrna_param_t (possibly edited by the caller) -> the library's model blob, and back for scale_parameters()
void model_from_vrna(const vrna_param_t &p, int no_gu, ccj_model &m);
void vrna_from_model(const ccj_model &m, vrna_param_t &p);

}  // namespace ccj

struct ccj_shell_fold {
    std::string seq;
    int n = 0;
    ccj_model model;
    std::vector<int8_t> S8;       // S8[0..n+1], S8[0]=S8[n], S8[n+1]=S8[1]
    bool filled = false;
    std::vector<int32_t> t2;      // CCJ_NT2 x stride2, fetched right after the fill
    int16_t *t4 = nullptr;        // CCJ_NT4 x C(n+1,4), lazily committed; a table is copied on first use
    bool have4[CCJ_NT4] = {false};

    ~ccj_shell_fold();
    void ensure_filled();         // bulk fill on the GPU (once) and make this fold the context's resident wave
    void ensure_resident();       // tables of THIS fold on the device (re-fills if another fold took the context)
    void need4(int table);
    ccj_cx cx();                  // host view for the cell functions (reads t2 / t4)
    energy_t raw2(int table, int i, int j) { ensure_filled(); return t2[(size_t)table * ccj_stride2(n) + ccj_idx2(n, i, j)]; }
    energy_t get4(int table, int i, int j, int k, int l);   // Matrix4D::get semantics
};

namespace ccj {
using ShellFold = ::ccj_shell_fold;

// the fold shared by all shell objects constructed for the same sequence and parameters
std::shared_ptr<ShellFold> shell_fold(const std::string &seq, const vrna_param_t *params);

}  // namespace ccj
#endif
