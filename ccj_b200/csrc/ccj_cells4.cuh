// The 22 four-dimensional gap tables of one cell (i,j,k,l), evaluated in the reference's in-cell
// order (src/pseudo_loop.cc:85-127).  Generic-index version: every read goes through ccj_get4 /
// ccj_idx4.  The tuned kernels in ccj_fill4.cu walk the same candidates with strided pointers and are
// checked against this function table-for-table.
#pragma once
#include "ccj_cells.cuh"

// get_PLiloop (src/pseudo_loop.cc:682-703)
CCJ_HD int ccj_PLiloop(const ccj_cx &c, int i, int j, int k, int l) {
    if (!ccj_can_pair(c, i, j)) return CCJ_INF;
    int mn = CCJ_INF;
    if (i + CCJ_TURN + 2 < j) mn = ccj_get4(c, T_PL, i + 1, j - 1, k, l) + ccj_e_stP(c.M, c.q.S, i, j);
    const int max_d = ccj_min(j, i + CCJ_MAXLOOP);
    for (int d = i + 1; d < max_d; ++d) {
        const int min_dp = ccj_max(d + CCJ_TURN, j - CCJ_MAXLOOP);
        for (int dp = j - 1; dp > min_dp; --dp) {
            if (!ccj_can_pair(c, d, dp)) continue;
            mn = ccj_min(mn, ccj_e_intP(c.M, c.q.S, i, d, dp, j) + ccj_get4(c, T_PL, d, dp, k, l));
        }
    }
    return mn;
}
// get_PRiloop (src/pseudo_loop.cc:717-738)
CCJ_HD int ccj_PRiloop(const ccj_cx &c, int i, int j, int k, int l) {
    if (!ccj_can_pair(c, k, l)) return CCJ_INF;
    int mn = CCJ_INF;
    if (k + CCJ_TURN + 2 < l) mn = ccj_get4(c, T_PR, i, j, k + 1, l - 1) + ccj_e_stP(c.M, c.q.S, k, l);
    const int max_d = ccj_min(l, k + CCJ_MAXLOOP);
    for (int d = k + 1; d < max_d; ++d) {
        const int min_dp = ccj_max(d + CCJ_TURN, l - CCJ_MAXLOOP);
        for (int dp = l - 1; dp > min_dp; --dp) {
            if (!ccj_can_pair(c, d, dp)) continue;
            mn = ccj_min(mn, ccj_e_intP(c.M, c.q.S, k, d, dp, l) + ccj_get4(c, T_PR, i, j, d, dp));
        }
    }
    return mn;
}
// get_PMiloop (src/pseudo_loop.cc:752-773)
CCJ_HD int ccj_PMiloop(const ccj_cx &c, int i, int j, int k, int l) {
    if (!ccj_can_pair(c, j, k)) return CCJ_INF;
    int mn = CCJ_INF;
    if (i < j && k < l) mn = ccj_get4(c, T_PM, i, j - 1, k + 1, l) + ccj_e_stP(c.M, c.q.S, j - 1, k + 1);
    const int max_d = ccj_max(i, j - CCJ_MAXLOOP);
    for (int d = j - 1; d > max_d; --d) {
        const int min_dp = ccj_min(l, k + CCJ_MAXLOOP);
        for (int dp = k + 1; dp < min_dp; ++dp) {
            if (!ccj_can_pair(c, d, dp)) continue;
            mn = ccj_min(mn, ccj_e_intP(c.M, c.q.S, d, j, k, dp) + ccj_get4(c, T_PM, i, d, dp, l));
        }
    }
    return mn;
}
// get_POiloop (src/pseudo_loop.cc:787-808): the window reads PO.get(d,j,dp,k) with dp>k, always an
// invalid index -> INF, so only the stacking term can contribute.
CCJ_HD int ccj_POiloop(const ccj_cx &c, int i, int j, int k, int l) {
    if (!ccj_can_pair(c, i, l)) return CCJ_INF;
    int mn = CCJ_INF;
    if (i < j && k < l) mn = ccj_get4(c, T_PO, i + 1, j, k, l - 1) + ccj_e_stP(c.M, c.q.S, i, l);
    return mn;
}

// get_P?mloop (src/pseudo_loop.cc:705-715,740-750,775-785,810-820); the cell (i,j,k,l) itself is valid
CCJ_HD int ccj_PXmloop(const ccj_cx &c, int t10, int t01, int i, int j, int k, int l) {
    const int add = c.M->ap_penalty + c.M->bp_penalty;
    return ccj_min(ccj_get4(c, t10, i, j, k, l) + add, ccj_get4(c, t01, i, j, k, l) + add);
}

CCJ_HD void ccj_cell4d(const ccj_cx &c, int i, int j, int k, int l) {
    const ccj_model *M = c.M;
    const ccj_pos4 idx = ccj_pos_of(c.q, i, j, k, l);
    const int INF = CCJ_INF;
    const int bp = M->bp_penalty, cp = M->cp_penalty, PB = M->PB_penalty;
    int mn, tmp;

    // ---- PLmloop00 / 01 / 10 (src/pseudo_loop.cc:445-493) ----
    mn = CCJ_INTERN_INF + bp;  // PL(i,j,k,l) is still unset here
    for (int d = i; d <= j; ++d) {
        if (d > i) mn = ccj_min(mn, ccj_WB(c, i, d - 1) + ccj_get4u(c, T_PLmloop00, d, j, k, l));
        if (d < j) mn = ccj_min(mn, ccj_get4u(c, T_PLmloop00, i, d, k, l) + ccj_WB(c, d + 1, j));
    }
    ccj_put4(c, T_PLmloop00, idx, mn);
    mn = INF;
    for (int d = i; d < j; ++d)
        mn = ccj_min(mn, ccj_get4u(c, T_PLmloop00, i, d, k, l) + ccj_tri_get(c, T2_WBP, d + 1, j));
    ccj_put4(c, T_PLmloop01, idx, mn);
    mn = INF;
    for (int d = i + 1; d <= j; ++d) {
        mn = ccj_min(mn, ccj_tri_get(c, T2_WBP, i, d - 1) + ccj_get4u(c, T_PLmloop00, d, j, k, l));
        if (d < j) mn = ccj_min(mn, ccj_get4u(c, T_PLmloop10, i, d, k, l) + ccj_WB(c, d + 1, j));
    }
    ccj_put4(c, T_PLmloop10, idx, mn);

    // ---- PRmloop00 / 01 / 10 (src/pseudo_loop.cc:495-542) ----
    mn = CCJ_INTERN_INF + bp;
    for (int d = k; d <= l; ++d) {
        if (d > k) mn = ccj_min(mn, ccj_WB(c, k, d - 1) + ccj_get4u(c, T_PRmloop00, i, j, d, l));
        if (d < l) mn = ccj_min(mn, ccj_get4u(c, T_PRmloop00, i, j, k, d) + ccj_WB(c, d + 1, l));
    }
    ccj_put4(c, T_PRmloop00, idx, mn);
    mn = ccj_get4(c, T_PRmloop01, i, j, k, l - 1) + cp;
    for (int d = k; d < l; ++d)
        mn = ccj_min(mn, ccj_get4u(c, T_PRmloop00, i, j, k, d) + ccj_tri_get(c, T2_WBP, d + 1, l));
    ccj_put4(c, T_PRmloop01, idx, mn);
    mn = ccj_get4(c, T_PRmloop10, i, j, k + 1, l) + cp;
    for (int d = k + 1; d <= l; ++d)
        mn = ccj_min(mn, ccj_tri_get(c, T2_WBP, k, d - 1) + ccj_get4u(c, T_PRmloop00, i, j, d, l));
    ccj_put4(c, T_PRmloop10, idx, mn);

    // ---- PMmloop00 / 01 / 10 (src/pseudo_loop.cc:544-593) ----
    mn = CCJ_INTERN_INF + bp;
    for (int d = i; d < j; ++d) mn = ccj_min(mn, ccj_get4u(c, T_PMmloop00, i, d, k, l) + ccj_WB(c, d + 1, j));
    for (int d = k + 1; d <= l; ++d) mn = ccj_min(mn, ccj_get4u(c, T_PMmloop00, i, j, d, l) + ccj_WB(c, k, d - 1));
    ccj_put4(c, T_PMmloop00, idx, mn);
    mn = ccj_get4(c, T_PMmloop01, i, j, k + 1, l) + cp;
    for (int d = k; d < l; ++d)
        mn = ccj_min(mn, ccj_get4u(c, T_PMmloop00, i, j, k, d) + ccj_tri_get(c, T2_WBP, d + 1, l));
    ccj_put4(c, T_PMmloop01, idx, mn);
    mn = ccj_get4(c, T_PMmloop10, i, j - 1, k, l) + cp;
    for (int d = i + 1; d <= j; ++d)
        mn = ccj_min(mn, ccj_tri_get(c, T2_WBP, i, d - 1) + ccj_get4u(c, T_PMmloop00, d, j, k, l));
    for (int d = k + 1; d < l; ++d) mn = ccj_min(mn, ccj_get4u(c, T_PMmloop10, i, j, k, d) + ccj_WB(c, d + 1, l));
    ccj_put4(c, T_PMmloop10, idx, mn);

    // ---- POmloop00 / 01 / 10 (src/pseudo_loop.cc:595-644) ----
    mn = CCJ_INTERN_INF + bp;
    for (int d = i + 1; d <= j; ++d) mn = ccj_min(mn, ccj_WB(c, i, d - 1) + ccj_get4u(c, T_POmloop00, d, j, k, l));
    for (int d = k; d < l; ++d) mn = ccj_min(mn, ccj_get4u(c, T_POmloop00, i, j, k, d) + ccj_WB(c, d + 1, l));
    ccj_put4(c, T_POmloop00, idx, mn);
    mn = INF;
    for (int d = k; d < l; ++d)
        mn = ccj_min(mn, ccj_get4u(c, T_POmloop00, i, j, k, d) + ccj_tri_get(c, T2_WBP, d + 1, l));
    ccj_put4(c, T_POmloop01, idx, mn);
    mn = INF;
    for (int d = i + 1; d <= j; ++d)
        mn = ccj_min(mn, ccj_tri_get(c, T2_WBP, i, d - 1) + ccj_get4u(c, T_POmloop00, d, j, k, l));
    for (int d = k + 1; d < l; ++d) mn = ccj_min(mn, ccj_get4u(c, T_POmloop10, i, j, k, d) + ccj_WB(c, d + 1, l));
    ccj_put4(c, T_POmloop10, idx, mn);

    // ---- PL (src/pseudo_loop.cc:232-253) ----
    mn = INF;
    if (ccj_pt(c, i, j) > 0) {
        mn = ccj_PLiloop(c, i, j, k, l);
        mn = ccj_min(mn, ccj_PXmloop(c, T_PLmloop10, T_PLmloop01, i + 1, j - 1, k, l) + bp);
        if (j >= i + CCJ_TURN + 1) mn = ccj_min(mn, ccj_get4(c, T_PfromL, i + 1, j - 1, k, l));
    }
    const int vPL = ccj_put4(c, T_PL, idx, mn);
    // ---- PR (src/pseudo_loop.cc:255-275) ----
    mn = INF;
    if (ccj_pt(c, k, l) > 0) {
        mn = ccj_PRiloop(c, i, j, k, l);
        mn = ccj_min(mn, ccj_PXmloop(c, T_PRmloop10, T_PRmloop01, i, j, k + 1, l - 1) + bp);
        if (l >= k + CCJ_TURN + 1) mn = ccj_min(mn, ccj_get4(c, T_PfromR, i, j, k + 1, l - 1));
    }
    const int vPR = ccj_put4(c, T_PR, idx, mn);
    // ---- PM (src/pseudo_loop.cc:277-300) ----
    mn = INF;
    if (ccj_pt(c, j, k) > 0) {
        mn = ccj_PMiloop(c, i, j, k, l);
        mn = ccj_min(mn, ccj_PXmloop(c, T_PMmloop10, T_PMmloop01, i, j - 1, k + 1, l) + bp);
        if (k >= j + CCJ_TURN - 1) mn = ccj_min(mn, ccj_get4(c, T_PfromM, i, j - 1, k + 1, l));
        if (i == j && k == l) mn = ccj_min(mn, 0);
    }
    const int vPM = ccj_put4(c, T_PM, idx, mn);
    // ---- PO (src/pseudo_loop.cc:302-322) ----
    mn = INF;
    if (ccj_pt(c, i, l) > 0) {
        mn = ccj_POiloop(c, i, j, k, l);
        mn = ccj_min(mn, ccj_PXmloop(c, T_POmloop10, T_POmloop01, i + 1, j, k, l - 1) + bp);
        if (l >= i + CCJ_TURN + 1) mn = ccj_min(mn, ccj_get4(c, T_PfromO, i + 1, j, k, l - 1));
    }
    const int vPO = ccj_put4(c, T_PO, idx, mn);

    // ---- PfromL (src/pseudo_loop.cc:354-374) ----
    mn = INF;
    for (int d = i + 1; d < j; ++d) {
        mn = ccj_min(mn, ccj_get4u(c, T_PfromL, d, j, k, l) + ccj_WP(c, i, d - 1));
        mn = ccj_min(mn, ccj_get4u(c, T_PfromL, i, d, k, l) + ccj_WP(c, d + 1, j));
    }
    mn = ccj_min(mn, ccj_min(vPR + PB, ccj_min(vPM + PB, vPO + PB)));
    ccj_put4(c, T_PfromL, idx, mn);
    // ---- PfromR (src/pseudo_loop.cc:376-394) ----
    mn = INF;
    for (int d = k + 1; d < l; ++d) {
        mn = ccj_min(mn, ccj_get4u(c, T_PfromR, i, j, d, l) + ccj_WP(c, k, d - 1));
        mn = ccj_min(mn, ccj_get4u(c, T_PfromR, i, j, k, d) + ccj_WP(c, d + 1, l));
    }
    mn = ccj_min(mn, ccj_min(vPM + PB, vPO + PB));
    ccj_put4(c, T_PfromR, idx, mn);
    // ---- PfromM (src/pseudo_loop.cc:396-407) ----
    mn = INF;
    for (int d = i + 1; d < j; ++d) mn = ccj_min(mn, ccj_get4u(c, T_PfromMprime, i, d, k, l) + ccj_WP(c, d + 1, j));
    ccj_put4(c, T_PfromM, idx, mn);
    // ---- PfromMprime (src/pseudo_loop.cc:409-420) with get_PfromMdoubleprime (:663-679); inside the
    //      loop d<l so the i==j&&k==l base case of M'' cannot occur ----
    mn = INF;
    for (int d = k + 1; d < l; ++d) {
        tmp = ccj_min(ccj_get4u(c, T_PL, i, j, d, l) + PB, ccj_get4u(c, T_PR, i, j, d, l) + PB);
        mn = ccj_min(mn, tmp + ccj_WP(c, k, d - 1));
    }
    ccj_put4(c, T_PfromMprime, idx, mn);
    // ---- PfromO (src/pseudo_loop.cc:422-443) ----
    mn = INF;
    for (int d = i + 1; d < j; ++d) mn = ccj_min(mn, ccj_get4u(c, T_PfromO, d, j, k, l) + ccj_WP(c, i, d - 1));
    for (int d = k + 1; d < l; ++d) mn = ccj_min(mn, ccj_get4u(c, T_PfromO, i, j, k, d) + ccj_WP(c, d + 1, l));
    mn = ccj_min(mn, ccj_min(vPL + PB, vPR + PB));
    ccj_put4(c, T_PfromO, idx, mn);
    // ---- PK (src/pseudo_loop.cc:181-202) ----
    mn = INF;
    for (int d = i + 1; d < j; ++d) mn = ccj_min(mn, ccj_get4u(c, T_PK, i, d, k, l) + ccj_WP(c, d + 1, j));
    for (int d = k + 1; d < l; ++d) mn = ccj_min(mn, ccj_get4u(c, T_PK, i, j, d, l) + ccj_WP(c, k, d - 1));
    mn = ccj_min(mn, ccj_min(ccj_min(vPL + PB, vPM + PB), ccj_min(vPR + PB, vPO + PB)));
    ccj_put4(c, T_PK, idx, mn);
}

// one (j,d,k) candidate of compute_P (src/pseudo_loop.cc:166-179)
CCJ_HD int ccj_P_term(const ccj_cx &c, int i, int l, int j, int d, int k) {
    return ccj_get4(c, T_PK, i, j, d + 1, k) + ccj_get4(c, T_PK, j + 1, d, k + 1, l);
}

// W_final::E_ext_Stem (src/W_final.cc:118-173)
CCJ_HD int ccj_E_ext_Stem(const ccj_cx &c, int vij, int vi1j, int vij1, int vi1j1, int i, int j) {
    const ccj_model *P = c.M;
    const int8_t *S = c.q.S;
    const int n = c.q.n;
    int e = CCJ_INF, en;
    int tt = ccj_pt(c, i, j);
    en = vij;
    if (en != CCJ_INF) {
        if (P->dangles == 2) {
            int si1 = i > 1 ? S[i - 1] : -1;
            int sj1 = j < n ? S[j + 1] : -1;
            en += ccj_E_ext_stem(P, tt, si1, sj1);
        } else {
            en += ccj_E_ext_stem(P, tt, -1, -1);
        }
        e = ccj_min(e, en);
    }
    if (P->dangles == 1) {
        tt = ccj_pt(c, i + 1, j);
        en = (j - i - 1 > CCJ_TURN) ? vi1j : CCJ_INF;
        if (en != CCJ_INF) en += ccj_E_ext_stem(P, tt, S[i], -1);
        e = ccj_min(e, en);
        tt = ccj_pt(c, i, j - 1);
        en = (j - 1 - i > CCJ_TURN) ? vij1 : CCJ_INF;
        if (en != CCJ_INF) en += ccj_E_ext_stem(P, tt, -1, S[j]);
        e = ccj_min(e, en);
        tt = ccj_pt(c, i + 1, j - 1);
        en = (j - 1 - i - 1 > CCJ_TURN) ? vi1j1 : CCJ_INF;
        if (en != CCJ_INF) en += ccj_E_ext_stem(P, tt, S[i], S[j]);
        e = ccj_min(e, en);
    }
    return e;
}

// exterior W[j] for one j, candidates spread over lanes (src/W_final.cc:68-77). W[0..j-1] final.
template <class Par>
CCJ_HD int ccj_W_at(const ccj_cx &c, int j, const Par &par) {
    const int32_t *W = c.q.W;
    int m2 = CCJ_INF, m3 = CCJ_INF;
    for (int k = 1 + par.lane(); k <= j - CCJ_TURN - 1; k += Par::nlanes) {
        const int acc = (k > 1) ? W[k - 1] : 0;
        m2 = ccj_min(m2, acc + ccj_E_ext_Stem(c, ccj_V(c, T2_V, k, j), ccj_V(c, T2_V, k + 1, j),
                                                ccj_V(c, T2_V, k, j - 1), ccj_V(c, T2_V, k + 1, j - 1), k, j));
        const int p4 = ccj_min(ccj_min(ccj_tri_get(c, T2_P, k, j), ccj_tri_get(c, T2_P, k + 1, j)),
                               ccj_min(ccj_tri_get(c, T2_P, k, j - 1), ccj_tri_get(c, T2_P, k + 1, j - 1)));
        m3 = ccj_min(m3, acc + p4 + c.M->PS_penalty);
    }
    m2 = par.red(m2);
    m3 = par.red(m3);
    return ccj_min(ccj_min(W[j - 1], m2), m3);
}
