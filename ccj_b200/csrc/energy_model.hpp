// Host-side energy-model loader: "## RNAfold parameter file v2.0" -> ccj_model (device blob).
//
// Written from scratch; replaces the reference's use of the vendored ViennaRNA reader/scaler:
//   src/ViennaRNA/params/io.c:454-674   set_parameters_from_string (sections, order, sizes)
//   src/ViennaRNA/params/io.c:713-764   get_array1 (tokens INF/DEF/NST/x/*, one comment per line)
//   src/ViennaRNA/params/io.c:1011-1078 rd_Tetraloop37 / rd_Hexaloop37 / rd_Triloop37
//   src/ViennaRNA/params/params.c:400-555 get_scaled_params at 37 C (identity + clamps)
//   src/ViennaRNA/pair_mat.h:80-155     make_pair_matrix (noGU)
#pragma once
#include <string>
#include <vector>
#include "ccj_types.h"

namespace ccj {

struct RawParams {
    // 37 C free energies in the reference's index space (pair 0..7, base 0..4)
    int stack[8][8];
    int hairpin[31], bulge[31], internal_loop[31];
    int mismatchI[8][5][5], mismatchH[8][5][5], mismatchM[8][5][5];
    int mismatch1nI[8][5][5], mismatch23I[8][5][5], mismatchExt[8][5][5];
    int dangle5[8][5], dangle3[8][5];
    int int11[8][8][5][5];
    int int21[8][8][5][5][5];
    int int22[8][8][5][5][5][5];
    int ML_BASE, ML_closing, ML_intern;
    int ninio, MAX_NINIO;
    int DuplexInit, TerminalAU;
    double lxc;
    char Tetraloops[281], Triloops[241], Hexaloops[361];
    int Tetraloop_E[40], Triloop_E[40], Hexaloop_E[40];
    unsigned present;  // bit per section actually read
    // check_symmetry() of the reference's reader (io.c:1126-1178): one line per asymmetric entry, in its order
    std::vector<std::string> warnings;
    RawParams();
};

// Reads `path`. Returns false and sets `err` on I/O or syntax errors. Enthalpy sections are parsed
// (to stay in sync with the line stream) and, except for the symmetry check, discarded: at 37 C the rescaling
// is the identity.
bool load_par_file(const char *path, RawParams &rp, std::string &err);
// the same reader on a text held in memory (set_parameters_from_string, io.c:454-674)
bool load_par_text(const char *text, size_t len, RawParams &rp, std::string &err);
// Brings `rp` to the state the reference starts from: its compiled-in Turner-2004 defaults
// (src/ViennaRNA/params/default.c).  A file read afterwards only overrides the sections it holds.
bool load_defaults(RawParams &rp, std::string &err);
// Parameter sets linked into the library (embedded_params.cpp): "rna_turner2004" (the defaults) and
// "dna_mathews2004" (vrna_params_load_DNA_Mathews2004, src/ViennaRNA/params/io.c; byte-identical to the
// reference's static/misc/dna_mathews2004.hex).  Returns nullptr for an unknown name.
const char *embedded_par(const char *name, size_t *len);

// get_scaled_params at 37 C with the default model details (dangles=2 at scaling time, special_hp=1),
// then model_details.dangles := dangles (src/W_final.cc:20-25), pair matrix with noGU, PK penalties.
void build_model(const RawParams &rp, int dangles, int noGU, ccj_model &out);

// Sequence encoding (src/ViennaRNA/pair_mat.h:47-73,158-183): A=1 C=2 G=3 U/T=4, others 0.
int encode_base(char c);

}  // namespace ccj
