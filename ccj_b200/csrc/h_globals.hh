// Definitions of the penalty globals declared in h_externs.hh, with the reference's values (src/h_globals.hh:7-25:
// the HotKnots 2.0 / DP09 set).  Include from exactly one translation unit of a program -- the one with main(), as
// the reference's src/CCJ.cc:5 does.  All energies in dcal/mol.
#ifndef CCJ_B200_H_GLOBALS_HH
#define CCJ_B200_H_GLOBALS_HH
#include "h_externs.hh"
int PS_penalty = -138;          // exterior pseudoloop initiation
int PSM_penalty = 1007;         // pseudoknot inside a multiloop
int PSP_penalty = 1500;         // pseudoknot inside a pseudoloop
int PB_penalty = 246;           // band
int PUP_penalty = 6;            // unpaired base in a pseudoloop or band
int PPS_penalty = 96;           // nested closed region in a pseudoloop / band-spanning multiloop
double e_stP_penalty = 0.89;    // stacked pair spanning a band: factor on the stacking energy
double e_intP_penalty = 0.74;   // interior loop spanning a band: factor on the loop energy
int a_penalty = 339, b_penalty = 3, c_penalty = 2;       // ordinary multiloop: initiation, branch, unpaired base
int ap_penalty = 341, bp_penalty = 56, cp_penalty = 12;  // multiloop that spans a band
#endif
