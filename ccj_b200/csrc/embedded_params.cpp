// Parameter sets linked into libccj_b200.so, taken at build time from the reference's own data files under params/
// (the build passes their absolute paths; nothing is copied into the source tree):
//   rna_turner2004  = params/rna_Turner04.par    -- the values of the reference's compiled-in defaults
//                                                    (src/ViennaRNA/params/default.c)
//   dna_mathews2004 = params/dna_Matthews04.par  -- byte-identical to src/ViennaRNA/static/misc/dna_mathews2004.hex,
//                                                    the set vrna_params_load_DNA_Mathews2004 loads (src/CCJ.cc:88-90)
#include <cstddef>
#include <cstring>

#include "energy_model.hpp"

#ifndef CCJ_PAR_TURNER04
#error "build with -DCCJ_PAR_TURNER04=\"<abs path>/params/rna_Turner04.par\" -DCCJ_PAR_DNA_MATHEWS04=..."
#endif

__asm__(".section .rodata\n"
        ".global ccj_par_turner04_begin\n.global ccj_par_turner04_end\n"
        "ccj_par_turner04_begin:\n.incbin \"" CCJ_PAR_TURNER04 "\"\nccj_par_turner04_end:\n.byte 0\n"
        ".global ccj_par_dna_mathews04_begin\n.global ccj_par_dna_mathews04_end\n"
        "ccj_par_dna_mathews04_begin:\n.incbin \"" CCJ_PAR_DNA_MATHEWS04 "\"\nccj_par_dna_mathews04_end:\n.byte 0\n"
        ".previous\n");

extern "C" {
extern const char ccj_par_turner04_begin[], ccj_par_turner04_end[];
extern const char ccj_par_dna_mathews04_begin[], ccj_par_dna_mathews04_end[];
}

namespace ccj {

const char *embedded_par(const char *name, size_t *len) {
    if (!name) return nullptr;
    if (!strcmp(name, "rna_turner2004")) {
        if (len) *len = (size_t)(ccj_par_turner04_end - ccj_par_turner04_begin);
        return ccj_par_turner04_begin;
    }
    if (!strcmp(name, "dna_mathews2004")) {
        if (len) *len = (size_t)(ccj_par_dna_mathews04_end - ccj_par_dna_mathews04_begin);
        return ccj_par_dna_mathews04_begin;
    }
    return nullptr;
}

}  // namespace ccj
