// TEST TOOL: prints size and member offsets of vrna_param_t / vrna_md_t / seq_interval / minimum_fold /
// free_energy_node.  Compiled once against the reference's headers (oracle/Makefile, -DPROBE_REFERENCE) and once
// against ccj_b200/csrc/ccj_compat.hh; the two outputs must be identical (tests/test_shell.py).
#include <cstddef>
#include <cstdio>
#ifdef PROBE_REFERENCE
#include "W_final.hh"
#else
#include "ccj_compat.hh"
#endif

#define OFF(T, m) printf(#T "." #m " %zu %zu\n", offsetof(T, m), sizeof(((T *)0)->m))
int main() {
    printf("sizeof vrna_param_t %zu\nsizeof vrna_md_t %zu\nsizeof seq_interval %zu\nsizeof minimum_fold %zu\nsizeof free_energy_node %zu\n",
           sizeof(vrna_param_t), sizeof(vrna_md_t), sizeof(seq_interval), sizeof(minimum_fold), sizeof(free_energy_node));
    OFF(vrna_param_t, id); OFF(vrna_param_t, stack); OFF(vrna_param_t, hairpin); OFF(vrna_param_t, bulge);
    OFF(vrna_param_t, internal_loop); OFF(vrna_param_t, mismatchExt); OFF(vrna_param_t, mismatchI);
    OFF(vrna_param_t, mismatch1nI); OFF(vrna_param_t, mismatch23I); OFF(vrna_param_t, mismatchH); OFF(vrna_param_t, mismatchM);
    OFF(vrna_param_t, dangle5); OFF(vrna_param_t, dangle3); OFF(vrna_param_t, int11); OFF(vrna_param_t, int21);
    OFF(vrna_param_t, int22); OFF(vrna_param_t, ninio); OFF(vrna_param_t, lxc); OFF(vrna_param_t, MLbase);
    OFF(vrna_param_t, MLintern); OFF(vrna_param_t, MLclosing); OFF(vrna_param_t, PS_penalty); OFF(vrna_param_t, PSM_penalty);
    OFF(vrna_param_t, PSP_penalty); OFF(vrna_param_t, PB_penalty); OFF(vrna_param_t, PUP_penalty); OFF(vrna_param_t, PPS_penalty);
    OFF(vrna_param_t, e_stP_penalty); OFF(vrna_param_t, e_intP_penalty); OFF(vrna_param_t, ap_penalty);
    OFF(vrna_param_t, bp_penalty); OFF(vrna_param_t, cp_penalty); OFF(vrna_param_t, a_penalty); OFF(vrna_param_t, b_penalty);
    OFF(vrna_param_t, c_penalty); OFF(vrna_param_t, TerminalAU); OFF(vrna_param_t, DuplexInit); OFF(vrna_param_t, Tetraloop_E);
    OFF(vrna_param_t, Tetraloops); OFF(vrna_param_t, Triloop_E); OFF(vrna_param_t, Triloops); OFF(vrna_param_t, Hexaloop_E);
    OFF(vrna_param_t, Hexaloops); OFF(vrna_param_t, TripleC); OFF(vrna_param_t, MultipleCA); OFF(vrna_param_t, MultipleCB);
    OFF(vrna_param_t, gquad); OFF(vrna_param_t, gquadLayerMismatch); OFF(vrna_param_t, gquadLayerMismatchMax);
    OFF(vrna_param_t, temperature); OFF(vrna_param_t, model_details); OFF(vrna_param_t, param_file);
    OFF(vrna_md_t, temperature); OFF(vrna_md_t, betaScale); OFF(vrna_md_t, pf_smooth); OFF(vrna_md_t, dangles);
    OFF(vrna_md_t, special_hp); OFF(vrna_md_t, noLP); OFF(vrna_md_t, noGU); OFF(vrna_md_t, noGUclosure); OFF(vrna_md_t, logML);
    OFF(vrna_md_t, circ); OFF(vrna_md_t, gquad); OFF(vrna_md_t, uniq_ML); OFF(vrna_md_t, energy_set); OFF(vrna_md_t, backtrack);
    OFF(vrna_md_t, backtrack_type); OFF(vrna_md_t, compute_bpp); OFF(vrna_md_t, nonstandards); OFF(vrna_md_t, max_bp_span);
    OFF(vrna_md_t, min_loop_size); OFF(vrna_md_t, window_size); OFF(vrna_md_t, oldAliEn); OFF(vrna_md_t, ribo);
    OFF(vrna_md_t, cv_fact); OFF(vrna_md_t, nc_fact); OFF(vrna_md_t, sfact); OFF(vrna_md_t, rtype); OFF(vrna_md_t, alias);
    OFF(vrna_md_t, pair);
    OFF(seq_interval, i); OFF(seq_interval, j); OFF(seq_interval, energy); OFF(seq_interval, type); OFF(seq_interval, next);
    OFF(seq_interval, k); OFF(seq_interval, l); OFF(seq_interval, asym);
    OFF(minimum_fold, pair); OFF(minimum_fold, type); OFF(free_energy_node, energy); OFF(free_energy_node, type);
    return 0;
}
