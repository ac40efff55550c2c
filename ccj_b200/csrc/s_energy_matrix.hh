// s_energy_matrix -- the reference's nested-table class surface (src/s_energy_matrix.hh:16-68) on top of the
// device-resident tables.  Same constructor, same public members and methods.  The per-cell compute_* entry points
// of the reference are accepted and ignored: the whole fill is ONE bulk GPU sweep, started the first time a value is
// asked for (so code that drives the reference's own i/j loops and then reads the tables gets the same numbers);
// the getters read a host mirror of what that sweep left in HBM; the loop-energy helpers evaluate the product's own
// host/device energy functions (ccj_energy.cuh / ccj_cells.cuh) for the given parameters.
#ifndef CCJ_B200_S_ENERGY_MATRIX_HH
#define CCJ_B200_S_ENERGY_MATRIX_HH
#include <memory>
#include <string>
#include <vector>

#include "ccj_compat.hh"

struct ccj_shell_fold;   // the shared bulk fold behind the shell objects (ccj_shell.hpp)

class s_energy_matrix {
public:
    // src/s_energy_matrix.hh:20; S/S1 are kept for the callers that read them back, the encoding the GPU uses is
    // derived from `seq` (same alphabet: A=1 C=2 G=3 U/T=4)
    s_energy_matrix(std::string seq, cand_pos_t length, short *S, short *S1, vrna_param_t *params);
    ~s_energy_matrix();

    vrna_param_t *params_;
    short *S_;
    short *S1_;

    void compute_energy(cand_pos_t, cand_pos_t) {}                      // src/s_energy_matrix.cc:315-358, bulk on the GPU
    void compute_WMv_WMp(cand_pos_t, cand_pos_t, energy_t) {}           // :206-217
    void compute_energy_WM(cand_pos_t, cand_pos_t, TriangleMatrix &) {} // :219-241
    energy_t compute_energy_VM(cand_pos_t i, cand_pos_t j);             // :243-268 (evaluated on the host mirror)

    free_energy_node *get_node(cand_pos_t i, cand_pos_t j);
    energy_t get_energy(cand_pos_t i, cand_pos_t j) { return (i >= j) ? INF : raw(T2_V, i, j); }
    energy_t get_energy_WM(cand_pos_t i, cand_pos_t j) { return (i >= j) ? INF : raw(T2_WM, i, j); }
    energy_t get_energy_WMv(cand_pos_t i, cand_pos_t j) { return (i >= j) ? INF : raw(T2_WMv, i, j); }
    energy_t get_energy_WMp(cand_pos_t i, cand_pos_t j) { return (i >= j) ? INF : raw(T2_WMp, i, j); }
    char get_type(cand_pos_t i, cand_pos_t j) { return (char)raw(T2_VTYPE, i, j); }

    // src/s_energy_matrix.cc:275-313
    energy_t HairpinE(const std::string &seq, const short *S, const short *S1, const paramT *params, cand_pos_t i, cand_pos_t j);
    energy_t compute_stack(cand_pos_t i, cand_pos_t j, const paramT *params);
    energy_t compute_internal(cand_pos_t i, cand_pos_t j, const paramT *params);
    energy_t compute_int(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l, const paramT *params);
    // :54-112, :122-205
    energy_t E_MLStem(const energy_t &vij, const energy_t &vi1j, const energy_t &vij1, const energy_t &vi1j1, const short *S,
                      paramT *params, cand_pos_t i, cand_pos_t j, cand_pos_t n);
    energy_t E_MbLoop(const energy_t WM2ij, const energy_t WM2ip1j, const energy_t WM2ijm1, const energy_t WM2ip1jm1,
                      const short *S, paramT *params, cand_pos_t i, cand_pos_t j);

    ccj_shell_fold *fold() { return fold_.get(); }   // extension: the shared bulk fold behind this object

protected:
    energy_t raw(int table, cand_pos_t i, cand_pos_t j);
    const ccj_model *model_for(const paramT *params);   // the object's model, or a conversion of a different `params`
    std::string seq_;
    cand_pos_t n;
    std::shared_ptr<ccj_shell_fold> fold_;
    std::vector<free_energy_node> nodes_;   // get_node() hands out stable pointers like the reference
    std::unique_ptr<ccj_model> other_model_;
    const paramT *other_params_ = nullptr;
};
#endif
