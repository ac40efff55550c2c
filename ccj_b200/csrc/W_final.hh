// W_final -- the reference's fold orchestrator (src/W_final.hh:18-71) with the same public surface:
//     W_final(std::string seq, int dangle);   double ccj();   std::string structure;   params_
// Internally everything is one call sequence into the C ABI of include/ccj_b200.h; the tables stay in HBM
// for the life of the object and are reachable through P / V getters.
#ifndef CCJ_B200_W_FINAL_HH
#define CCJ_B200_W_FINAL_HH
#include <string>
#include <vector>

#include "pseudo_loop.hh"
#include "s_energy_matrix.hh"

// Process-global configuration, as in the reference: vrna_params_load() stores the parameter set the next
// W_final uses (src/CCJ.cc:80-99), `noGU` is the ViennaRNA global read by make_pair_matrix (src/CCJ.cc:77).
extern int noGU;
int ccj_params_load(const char *par_file);  // 1 on success (like vrna_params_load), 0 if unreadable
void ccj_params_load_DNA_Mathews2004();     // vrna_params_load_DNA_Mathews2004: the set linked into the library

struct ccj_params_view {   // what W_final::params_ exposes of the loaded model
    std::string param_file;
    int dangles;
};

class W_final {
public:
    W_final(std::string seq, int dangle);
    ~W_final();
    double ccj();  // fold: returns the MFE in kcal/mol, leaves the dot-bracket in `structure`

    ccj_params_view *params_;
    std::string structure;

    // tables of the finished fold
    pseudo_loop *P;
    s_energy_matrix *V;
    std::vector<energy_t> W;          // only W[n] is fetched (the exterior energy)
    std::vector<int32_t> pairs;       // minimum_fold::pair per nucleotide (1-based partner, -1 unpaired)

protected:
    cand_pos_t n;
    std::string seq_;
    ccj_ctx *ctx_;
};
#endif
