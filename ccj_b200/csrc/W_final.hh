// W_final -- the reference's fold orchestrator (src/W_final.hh:18-71) with the same public surface:
//     W_final(std::string seq, int dangle);   double ccj();   vrna_param_t *params_;   std::string structure;
// Internally everything is one call sequence into the C ABI of include/ccj_b200.h; the tables stay in HBM
// for the life of the object and are reachable through P / V (protected in the reference, public here).
#ifndef CCJ_B200_W_FINAL_HH
#define CCJ_B200_W_FINAL_HH
#include <string>
#include <vector>

#include "pseudo_loop.hh"
#include "s_energy_matrix.hh"

// kept from round 1 for callers of the C++ shell: 1 on success (like vrna_params_load), 0 if unreadable
int ccj_params_load(const char *par_file);
void ccj_params_load_DNA_Mathews2004();     // vrna_params_load_DNA_Mathews2004: the set linked into the library

class W_final {
public:
    W_final(std::string seq, int dangle);
    ~W_final();
    double ccj();  // fold: returns the MFE in kcal/mol, leaves the dot-bracket in `structure`

    vrna_param_t *params_;
    std::string structure;

    // tables of the finished fold
    pseudo_loop *P;
    s_energy_matrix *V;
    std::vector<energy_t> W;          // only W[n] is fetched (the exterior energy)
    std::vector<int32_t> pairs;       // minimum_fold::pair per nucleotide (1-based partner, -1 unpaired)

protected:
    cand_pos_t n;
    std::string seq_;
    short *S_;
    short *S1_;
};
#endif
