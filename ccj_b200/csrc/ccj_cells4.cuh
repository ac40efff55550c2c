// The 22 four-dimensional gap tables of one cell (i,j,k,l) (src/pseudo_loop.cc:85-127).  Generic-index version: every
// read goes through ccj_pos_of / ccj_addr4 (ordinary or sharded layout).  The tuned kernels in ccj_fill4.cu walk the same candidates with strided pointers and are
// checked against this function table-for-table.
#pragma once
#include "ccj_cells.cuh"

// The interior windows (get_P{L,R,M}iloop, src/pseudo_loop.cc:682-773) in two equivalent forms: the reference's scan of
// the 29 x 29 window with a can_pair test per slot, and -- where the fold has per-pair partner lists (ccj_seq::inlist /
// outlist, written once per sequence by k_prep with the already rounded energies) -- a walk over the pairable partners
// only.  Same candidates, same energies, same minimum; the scan touches ~7 slots per candidate it evaluates.
//   inlist  of (p,q): partners INSIDE  the pair, entry = int16 energy | x << 16 | y << 24 for (p+x, q-y)
//   outlist of (p,q): partners OUTSIDE the pair, entry.x the same for (p-x, q+y); candidates with a negative energy first
// status[5]: k_prep met an interior-loop energy outside int16 -- the lists cannot hold this model, scan instead
CCJ_HD bool ccj_lists_ok(const ccj_cx &c) { return c.q.use_lists != 0 && c.q.status[5] == 0; }
CCJ_HD int ccj_win_inside(const ccj_cx &c, int tbl, int p, int q, int i, int j, int k, int l, bool left) {
    const int slot = ccj_tri(p, q);
    const uint32_t *lst = c.q.inlist + (int64_t)slot * CCJ_WIN_IN;
    const int cnt = c.q.incnt[slot];
    int mn = CCJ_INF;
    for (int e = 0; e < cnt; ++e) {
        const uint32_t en = lst[e];
        const int x = (en >> 16) & 0xff, y = en >> 24, energy = (int)(int16_t)(en & 0xffff);
        // PL: PL(i+x, j-y, k, l) closed by (i,j);  PR: PR(i, j, k+x, l-y) closed by (k,l)
        mn = ccj_min(mn, energy + (left ? ccj_get4u(c, tbl, i + x, j - y, k, l) : ccj_get4u(c, tbl, i, j, k + x, l - y)));
    }
    return mn;
}

// get_PLiloop (src/pseudo_loop.cc:682-703)
CCJ_HD int ccj_PLiloop(const ccj_cx &c, int i, int j, int k, int l) {
    if (!ccj_can_pair(c, i, j)) return CCJ_INF;
    int mn = CCJ_INF;
    if (i + CCJ_TURN + 2 < j) mn = ccj_get4(c, T_PL, i + 1, j - 1, k, l) + ccj_e_stP(c.M, c.q.S, i, j);
    if (ccj_lists_ok(c)) return ccj_min(mn, ccj_win_inside(c, T_PL, i, j, i, j, k, l, true));
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
    if (ccj_lists_ok(c)) return ccj_min(mn, ccj_win_inside(c, T_PR, k, l, i, j, k, l, false));
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
    if (ccj_lists_ok(c)) {
        // partners (d,dp) = (j-x, k+y) of the inner pair (j,k); this cell admits d > max(i, j-30), dp < min(l, k+30),
        // i.e. x < j-i and y < l-k (x, y <= 29 by construction of the list)
        const int slot = ccj_tri(j, k), a = j - i, b = l - k;
        const uint32_t *lst = c.q.outlist + (int64_t)slot * CCJ_WIN_OUT * 2;   // 8-byte entries: (packed term, PMW row offset)
        const int cnt = c.q.outcnt[slot] & 0xffff;
        for (int e = 0; e < cnt; ++e) {
            const uint32_t en = lst[2 * e];
            const int x = (en >> 16) & 0xff, y = en >> 24;
            if (x < a && y < b) mn = ccj_min(mn, (int)(int16_t)(en & 0xffff) + ccj_get4u(c, T_PM, i, j - x, k + y, l));
        }
        return mn;
    }
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

// {WB, WP, WBP} of an interval 1 <= i <= j <= n: one 16-byte record where the fold keeps them packed (ccj_seq::w3, written
// with the 2D tables), else the three getters (get_WB / get_WP, src/pseudo_loop.cc:647-661; WBP.get)
struct ccj_w3v {
    int wb, wp, wbp;
};
CCJ_HD ccj_w3v ccj_w3_at(const ccj_cx &c, int i, int j) {
    ccj_w3v r;
    if (c.q.w3) {
        const int32_t *w = c.q.w3 + 4 * (int64_t)ccj_idx2(c.q.n, i, j);
#if defined(__CUDA_ARCH__)
        const int4 v = __ldg(reinterpret_cast<const int4 *>(w));
        r.wb = v.x; r.wp = v.y; r.wbp = v.z;
#else
        r.wb = w[0]; r.wp = w[1]; r.wbp = w[2];
#endif
    } else {
        r.wb = ccj_WB(c, i, j);
        r.wp = ccj_WP(c, i, j);
        r.wbp = ccj_tri_get(c, T2_WBP, i, j);
    }
    return r;
}

// split-point loops: iterations are independent (only the running minima carry over), unrolling lets the loads of
// several split points be in flight at once
#ifndef CCJ_CELL_UNROLL_N
#define CCJ_CELL_UNROLL_N 1
#endif
#define CCJ_PRAGMA(x) _Pragma(#x)
#define CCJ_UNROLL_BY(n) CCJ_PRAGMA(unroll n)
#define CCJ_CELL_UNROLL CCJ_UNROLL_BY(CCJ_CELL_UNROLL_N)

// All 22 tables of one cell.  Every split-point candidate of the 22 recurrences (src/pseudo_loop.cc:181-644) reads a
// cell of a LOWER level, so their order inside the cell is free: they are walked by the four access patterns
//     L1: X(i,d,k,l) with the 2D interval (d+1,j)     L2: X(d,j,k,l) with (i,d-1)
//     R3: X(i,j,d,l) with (k,d-1)                     R4: X(i,j,k,d) with (d+1,l)
// -- one position computation per split point serves every table read there (the dominant cost of this generic
// version, in the ordinary and even more in the sharded layout) --, and the same-cell terms are then applied in the
// reference's in-cell order (:85-127), which is what fixes the "PX(cell) is still unset = 32767" reads of P?mloop00.
CCJ_HD void ccj_cell4d(const ccj_cx &c, int i, int j, int k, int l) {
    const ccj_model *M = c.M;
    const ccj_seq &q = c.q;
    const ccj_pos4 idx = ccj_pos_of(q, i, j, k, l);
    const int INF = CCJ_INF;
    const int bp = M->bp_penalty, cp = M->cp_penalty, PB = M->PB_penalty;
#define RD(tbl, pos) ((int)*ccj_addr4(q, (tbl), (pos)))
    // partial minima per recurrence, by pattern
    int PLm00 = CCJ_INTERN_INF + bp, PLm01 = INF, PLm10 = INF;   // PL(i,j,k,l) is still unset where PLmloop00 reads it
    int PRm00 = CCJ_INTERN_INF + bp, PRm01, PRm10;
    int PMm00 = CCJ_INTERN_INF + bp, PMm01, PMm10;
    int POm00 = CCJ_INTERN_INF + bp, POm01 = INF, POm10 = INF;
    int PfL = INF, PfR = INF, PfM = INF, PfMp = INF, PfO = INF, PK = INF;
    PRm01 = ccj_get4(c, T_PRmloop01, i, j, k, l - 1) + cp;   // :517-519
    PRm10 = ccj_get4(c, T_PRmloop10, i, j, k + 1, l) + cp;   // :531-533
    PMm01 = ccj_get4(c, T_PMmloop01, i, j, k + 1, l) + cp;   // :564-566
    PMm10 = ccj_get4(c, T_PMmloop10, i, j - 1, k, l) + cp;   // :578-580

    // ---- L1: X(i,d,k,l), d = i .. j-1 ----
    CCJ_CELL_UNROLL
    for (int d = i; d < j; ++d) {
        const ccj_pos4 p = ccj_pos_of(q, i, d, k, l);
        const ccj_w3v w = ccj_w3_at(c, d + 1, j);
        const int wb = w.wb, wbp = w.wbp;
        const int x00 = RD(T_PLmloop00, p);
        PLm00 = ccj_min(PLm00, x00 + wb);                       // :455-458
        PLm01 = ccj_min(PLm01, x00 + wbp);                      // :468-471
        PMm00 = ccj_min(PMm00, RD(T_PMmloop00, p) + wb);        // :548-551
        if (d > i) {
            const int wp = w.wp;
            PLm10 = ccj_min(PLm10, RD(T_PLmloop10, p) + wb);    // :484-487
            PfL = ccj_min(PfL, RD(T_PfromL, p) + wp);           // :360-361
            PfM = ccj_min(PfM, RD(T_PfromMprime, p) + wp);      // :399-402
            PK = ccj_min(PK, RD(T_PK, p) + wp);                 // :184-187
        }
    }
    // ---- L2: X(d,j,k,l), d = i+1 .. j ----
    CCJ_CELL_UNROLL
    for (int d = i + 1; d <= j; ++d) {
        const ccj_pos4 p = ccj_pos_of(q, d, j, k, l);
        const ccj_w3v w = ccj_w3_at(c, i, d - 1);
        const int wb = w.wb, wbp = w.wbp;
        const int x00 = RD(T_PLmloop00, p), o00 = RD(T_POmloop00, p);
        PLm00 = ccj_min(PLm00, wb + x00);                       // :450-453
        PLm10 = ccj_min(PLm10, wbp + x00);                      // :481-483
        PMm10 = ccj_min(PMm10, wbp + RD(T_PMmloop00, p));       // :581-584
        POm00 = ccj_min(POm00, wb + o00);                       // :599-602
        POm10 = ccj_min(POm10, wbp + o00);                      // :632-635
        if (d < j) {
            const int wp = w.wp;
            PfL = ccj_min(PfL, RD(T_PfromL, p) + wp);           // :357-359
            PfO = ccj_min(PfO, RD(T_PfromO, p) + wp);           // :425-428
        }
    }
    // ---- R3: X(i,j,d,l), d = k+1 .. l ----
    CCJ_CELL_UNROLL
    for (int d = k + 1; d <= l; ++d) {
        const ccj_pos4 p = ccj_pos_of(q, i, j, d, l);
        const ccj_w3v w = ccj_w3_at(c, k, d - 1);
        const int wb = w.wb, wbp = w.wbp;
        const int r00 = RD(T_PRmloop00, p);
        PRm00 = ccj_min(PRm00, wb + r00);                       // :499-503
        PRm10 = ccj_min(PRm10, wbp + r00);                      // :534-537
        PMm00 = ccj_min(PMm00, RD(T_PMmloop00, p) + wb);        // :552-555
        if (d < l) {
            const int wp = w.wp;
            PfR = ccj_min(PfR, RD(T_PfromR, p) + wp);           // :379-381
            // get_PfromMdoubleprime (:663-679); d<l, so its i==j&&k==l base case cannot occur
            PfMp = ccj_min(PfMp, ccj_min(RD(T_PL, p) + PB, RD(T_PR, p) + PB) + wp);   // :412-415
            PK = ccj_min(PK, RD(T_PK, p) + wp);                 // :189-192
        }
    }
    // ---- R4: X(i,j,k,d), d = k .. l-1 ----
    CCJ_CELL_UNROLL
    for (int d = k; d < l; ++d) {
        const ccj_pos4 p = ccj_pos_of(q, i, j, k, d);
        const ccj_w3v w = ccj_w3_at(c, d + 1, l);
        const int wb = w.wb, wbp = w.wbp;
        const int r00 = RD(T_PRmloop00, p), o00 = RD(T_POmloop00, p);
        PRm00 = ccj_min(PRm00, r00 + wb);                       // :504-507
        PRm01 = ccj_min(PRm01, r00 + wbp);                      // :520-523
        PMm01 = ccj_min(PMm01, RD(T_PMmloop00, p) + wbp);       // :567-570
        POm00 = ccj_min(POm00, o00 + wb);                       // :603-606
        POm01 = ccj_min(POm01, o00 + wbp);                      // :618-621
        if (d > k) {
            const int wp = w.wp;
            PMm10 = ccj_min(PMm10, RD(T_PMmloop10, p) + wb);    // :585-588
            POm10 = ccj_min(POm10, RD(T_POmloop10, p) + wb);    // :636-639
            PfR = ccj_min(PfR, RD(T_PfromR, p) + wp);           // :382-383
            PfO = ccj_min(PfO, RD(T_PfromO, p) + wp);           // :429-432
        }
    }
#undef RD
    // ---- stores, in the reference's in-cell order (:85-127) ----
    ccj_put4(c, T_PLmloop00, idx, PLm00);
    ccj_put4(c, T_PLmloop01, idx, PLm01);
    ccj_put4(c, T_PLmloop10, idx, PLm10);
    ccj_put4(c, T_PRmloop00, idx, PRm00);
    ccj_put4(c, T_PRmloop01, idx, PRm01);
    ccj_put4(c, T_PRmloop10, idx, PRm10);
    ccj_put4(c, T_PMmloop00, idx, PMm00);
    ccj_put4(c, T_PMmloop01, idx, PMm01);
    ccj_put4(c, T_PMmloop10, idx, PMm10);
    ccj_put4(c, T_POmloop00, idx, POm00);
    ccj_put4(c, T_POmloop01, idx, POm01);
    ccj_put4(c, T_POmloop10, idx, POm10);
    int mn;
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
    // ---- PfromL / PfromR / PfromM / PfromMprime / PfromO / PK: the same-cell terms (:354-443, :181-202) ----
    ccj_put4(c, T_PfromL, idx, ccj_min(PfL, ccj_min(vPR + PB, ccj_min(vPM + PB, vPO + PB))));
    ccj_put4(c, T_PfromR, idx, ccj_min(PfR, ccj_min(vPM + PB, vPO + PB)));
    ccj_put4(c, T_PfromM, idx, PfM);
    ccj_put4(c, T_PfromMprime, idx, PfMp);
    ccj_put4(c, T_PfromO, idx, ccj_min(PfO, ccj_min(vPL + PB, vPR + PB)));
    ccj_put4(c, T_PK, idx, ccj_min(PK, ccj_min(ccj_min(vPL + PB, vPM + PB), ccj_min(vPR + PB, vPO + PB))));
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
