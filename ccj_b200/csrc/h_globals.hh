// Definitions of the penalty globals declared in h_externs.hh, with the reference's values (src/h_globals.hh:7-25:
// the HotKnots 2.0 / DP09 set).  Include from exactly one translation unit of a program -- the one with main(), as
// the reference's src/CCJ.cc:5 does.  All energies in dcal/mol.
#ifndef CCJ_B200_H_GLOBALS_HH
#define CCJ_B200_H_GLOBALS_HH
#include "h_externs.hh"
// pseudoloop initiation: exterior / inside a multiloop / inside a pseudoloop; band; unpaired base in a pseudoloop or band;
// nested closed region in a pseudoloop or band-spanning multiloop
int PS_penalty = -138, PSM_penalty = 1007, PSP_penalty = 1500, PB_penalty = 246, PUP_penalty = 6, PPS_penalty = 96;
// factors on the stacking / interior-loop energy of a pair that spans a band
double e_stP_penalty = 0.89, e_intP_penalty = 0.74;
// multiloop initiation, branch, unpaired base: ordinary (a, b, c) and spanning a band (ap, bp, cp)
int a_penalty = 339, b_penalty = 3, c_penalty = 2, ap_penalty = 341, bp_penalty = 56, cp_penalty = 12;
#endif
