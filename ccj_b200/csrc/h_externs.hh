// The pseudoknot penalties of the reference are mutable process globals (src/h_externs.hh:5-22, defined once in
// src/h_globals.hh:7-25 by the translation unit that holds main()).  The shells keep that contract: the values in
// force when a W_final / s_energy_matrix / pseudo_loop is constructed are the ones the GPU uses.
#ifndef CCJ_B200_H_EXTERNS_HH
#define CCJ_B200_H_EXTERNS_HH
extern int PS_penalty, PSM_penalty, PSP_penalty, PB_penalty, PUP_penalty, PPS_penalty;
extern int a_penalty, b_penalty, c_penalty;
extern double e_stP_penalty, e_intP_penalty;
extern int ap_penalty, bp_penalty, cp_penalty;
#endif
