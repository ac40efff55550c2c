// Loop-energy functions on the ccj_model blob.  Integer dcal/mol, bit-exact restatement targets:
//   E_Hairpin        src/ViennaRNA/loops/hairpin.h:149-200
//   E_IntLoop        src/ViennaRNA/loops/internal.h:478-569
//   E_MLstem         src/ViennaRNA/loops/multibranch.h:226-246
//   vrna_E_ext_stem  src/ViennaRNA/loops/external.c:384-404  (== E_ExtLoop external.c:2191-2209)
//   get_e_stP/get_e_intP/compute_int  src/pseudo_loop.cc:822-840
#pragma once
#include "ccj_types.h"

#if defined(__CUDA_ARCH__)
#define CCJ_LD(x) __ldg(&(x))
#else
#define CCJ_LD(x) (x)
#endif

CCJ_HD int ccj_min(int a, int b) { return a < b ? a : b; }
CCJ_HD int ccj_max(int a, int b) { return a > b ? a : b; }

// lrint(p * e) under the default rounding mode (round-half-even): cvt.rni on device, lrint on host
CCJ_HD int ccj_scale_round(double p, int e) {
#if defined(__CUDA_ARCH__)
    return __double2int_rn(p * (double)e);
#else
    return (int)__builtin_lrint(p * (double)e);
#endif
}

CCJ_HD int ccj_ptype(const ccj_model *M, const int8_t *S, int i, int j) { return CCJ_LD(M->pair[S[i]][S[j]]); }

CCJ_HD int ccj_E_IntLoop(const ccj_model *P, int n1, int n2, int type, int type_2, int si1, int sj1, int sp1,
                         int sq1) {
    int nl, ns, u, energy;
    if (n1 > n2) { nl = n1; ns = n2; } else { nl = n2; ns = n1; }
    if (nl == 0) return CCJ_LD(P->stack[type][type_2]);
    if (ns == 0) {
        energy = CCJ_LD(P->bulge[nl]);
        if (nl == 1) {
            energy += CCJ_LD(P->stack[type][type_2]);
        } else {
            if (type > 2) energy += CCJ_LD(P->TerminalAU);
            if (type_2 > 2) energy += CCJ_LD(P->TerminalAU);
        }
        return energy;
    }
    if (ns == 1) {
        if (nl == 1) return CCJ_LD(P->int11[type][type_2][si1][sj1]);
        if (nl == 2) {
            if (n1 == 1) return CCJ_LD(P->int21[type][type_2][si1][sq1][sj1]);
            return CCJ_LD(P->int21[type_2][type][sq1][si1][sp1]);
        }
        energy = CCJ_LD(P->internal_loop[nl + 1]);
        energy += ccj_min(CCJ_LD(P->max_ninio), (nl - ns) * CCJ_LD(P->ninio2));
        energy += CCJ_LD(P->mismatch1nI[type][si1][sj1]) + CCJ_LD(P->mismatch1nI[type_2][sq1][sp1]);
        return energy;
    }
    if (ns == 2) {
        if (nl == 2) return CCJ_LD(P->int22[type][type_2][si1][sp1][sq1][sj1]);
        if (nl == 3) {
            energy = CCJ_LD(P->internal_loop[5]) + CCJ_LD(P->ninio2);
            energy += CCJ_LD(P->mismatch23I[type][si1][sj1]) + CCJ_LD(P->mismatch23I[type_2][sq1][sp1]);
            return energy;
        }
    }
    u = nl + ns;
    energy = CCJ_LD(P->internal_loop[u]);
    energy += ccj_min(CCJ_LD(P->max_ninio), (nl - ns) * CCJ_LD(P->ninio2));
    energy += CCJ_LD(P->mismatchI[type][si1][sj1]) + CCJ_LD(P->mismatchI[type_2][sq1][sp1]);
    return energy;
}

// compute_int(i,j,k,l): closing pair (i,j), inner pair (k,l)  (src/pseudo_loop.cc:822-826,
// src/s_energy_matrix.cc:309-313 without the V term)
CCJ_HD int ccj_compute_int(const ccj_model *M, const int8_t *S, int i, int j, int k, int l) {
    const int t1 = ccj_ptype(M, S, i, j);
    const int t2 = CCJ_LD(M->rtype[ccj_ptype(M, S, k, l)]);
    return ccj_E_IntLoop(M, k - i - 1, j - l - 1, t1, t2, S[i + 1], S[j - 1], S[k - 1], S[l + 1]);
}

CCJ_HD int ccj_e_stP(const ccj_model *M, const int8_t *S, int i, int j) {
    if (i + 1 == j - 1) return CCJ_INF;
    return ccj_scale_round(M->e_stP_penalty, ccj_compute_int(M, S, i, j, i + 1, j - 1));
}

CCJ_HD int ccj_e_intP(const ccj_model *M, const int8_t *S, int i, int ip, int jp, int j) {
    return ccj_scale_round(M->e_intP_penalty, ccj_compute_int(M, S, i, j, ip, jp));
}

CCJ_HD int ccj_E_MLstem(const ccj_model *P, int type, int si1, int sj1) {
    int energy = 0;
    if (si1 >= 0 && sj1 >= 0) energy += CCJ_LD(P->mismatchM[type][si1][sj1]);
    else if (si1 >= 0) energy += CCJ_LD(P->dangle5[type][si1]);
    else if (sj1 >= 0) energy += CCJ_LD(P->dangle3[type][sj1]);
    if (type > 2) energy += CCJ_LD(P->TerminalAU);
    energy += CCJ_LD(P->MLintern[type]);
    return energy;
}

CCJ_HD int ccj_E_ext_stem(const ccj_model *P, int type, int n5d, int n3d) {
    int energy = 0;
    if (n5d >= 0 && n3d >= 0) energy += CCJ_LD(P->mismatchExt[type][n5d][n3d]);
    else if (n5d >= 0) energy += CCJ_LD(P->dangle5[type][n5d]);
    else if (n3d >= 0) energy += CCJ_LD(P->dangle3[type][n3d]);
    if (type > 2) energy += CCJ_LD(P->TerminalAU);
    return energy;
}

CCJ_HD bool ccj_match(const char *entry, const char *s, int len) {
    for (int x = 0; x < len; ++x)
        if (entry[x] != s[x]) return false;
    return true;
}

// HairpinE(i,j) (src/s_energy_matrix.cc:275-282) -> E_Hairpin(size,type,si1,sj1,&seq[i-1])
CCJ_HD int ccj_HairpinE(const ccj_model *P, const int8_t *S, const char *seq, int i, int j) {
    const int type = ccj_ptype(P, S, i, j);
    if (type == 0) return CCJ_INF;
    const int size = j - i - 1;
    int energy = CCJ_LD(P->hairpin[size]);
    if (size < 3) return energy;
    const char *str = seq + (i - 1);
    if (P->special_hp) {
        if (size == 4) {
            for (int e = 0; e < P->n_tetra; ++e)
                if (ccj_match(P->tetra[e], str, 6)) return CCJ_LD(P->tetra_E[e]);
        } else if (size == 6) {
            for (int e = 0; e < P->n_hexa; ++e)
                if (ccj_match(P->hexa[e], str, 8)) return CCJ_LD(P->hexa_E[e]);
        } else if (size == 3) {
            for (int e = 0; e < P->n_tri; ++e)
                if (ccj_match(P->tri[e], str, 5)) return CCJ_LD(P->tri_E[e]);
            return energy + (type > 2 ? CCJ_LD(P->TerminalAU) : 0);
        }
    }
    energy += CCJ_LD(P->mismatchH[type][S[i + 1]][S[j - 1]]);
    return energy;
}
