// The CCJ recurrences, one function per DP cell.  Shared by every kernel (and by the host-side
// single-thread emulation the unit tests use to debug without a GPU).
//
// Restatement targets (reference file:line):
//   nested tables    src/s_energy_matrix.cc:54-358   (V, WM, WMv, WMp)
//   2D PK tables     src/pseudo_loop.cc:134-179      (WBP, WPP, P) + get_WB/get_WP :647-661
//   4D gap tables    src/pseudo_loop.cc:69-132 (in-cell order), :181-644 (22 recurrences), :663-840 (helpers)
//   exterior         src/W_final.cc:68-77,118-173
// Semantics kept on purpose: two infinities (INF for invalid index, 32767 for a
// valid-but-unset int16 cell), "store only if < INF/2", int16 narrowing, the in-cell evaluation order
// (P?mloop00 sees PX(cell)=32767), no can_pair guard on V, the dead PO window.
#pragma once
#include "ccj_energy.cuh"

struct ccj_cx {
    const ccj_model *M;
    ccj_seq q;
};

// ---------------------------------------------------------------------------------------------
// accessors
// ---------------------------------------------------------------------------------------------
CCJ_HD int16_t *ccj_t4(const ccj_cx &c, int tbl) { return c.q.t4 + (int64_t)tbl * c.q.stride4; }
CCJ_HD int32_t *ccj_t2(const ccj_cx &c, int tbl) { return c.q.t2 + (int64_t)tbl * c.q.stride2; }

// Where a cell's 22 entries live: one index per layout.  Ordinary waves: table `tbl` at t4[tbl*stride4 + off].
// Sharded folds (ccj_types.h, "sharded layout"): a column-read table at shard_rep[off + kind*C], a row-local one at
// shard_loc[owner][aux + (kind-12)*C].
struct ccj_pos4 {
    int64_t off, aux, C;
    int32_t owner;
};
CCJ_HD ccj_pos4 ccj_pos_of(const ccj_seq &q, int i, int j, int k, int l) {
    ccj_pos4 p;
    if (q.shard_G == 0) {
        p.off = ccj_idx4(q.n, i, j, k, l);
        p.aux = 0;
        p.C = q.stride4;
        p.owner = 0;
        return p;
    }
    const int G = q.shard_G, t = (j - i) + (l - k);
    const int64_t L = q.shard_lev[t], C = q.shard_lev[t + 1] - L;
    const int64_t inner = ccj_shard_inner_fast(q.n, G, q.shard_shift, i, j, k, l, p.owner);
    p.C = C;
    p.off = L * G * CCJ_SHARD_NREP + (int64_t)p.owner * CCJ_SHARD_NREP * C + inner;
    p.aux = L * CCJ_SHARD_NLOC + inner;
    return p;
}
CCJ_HD int16_t *ccj_addr4(const ccj_seq &q, int tbl, const ccj_pos4 &p) {
    if (q.shard_G == 0) return q.t4 + (int64_t)tbl * p.C + p.off;
    const int kd = q.shard_kind[tbl];
    return kd < CCJ_SHARD_NREP ? q.shard_rep + p.off + (int64_t)kd * p.C
                               : q.shard_loc[p.owner] + p.aux + (int64_t)(kd - CCJ_SHARD_NREP) * p.C;
}

// Matrix4D::get (src/matrices.hh:177-182)
CCJ_HD int ccj_get4(const ccj_cx &c, int tbl, int i, int j, int k, int l) {
    if (!ccj_valid4(i, j, k, l)) return CCJ_INF;
    if (c.q.shard_G == 0) return (int)ccj_t4(c, tbl)[ccj_idx4(c.q.n, i, j, k, l)];
    return (int)*ccj_addr4(c.q, tbl, ccj_pos_of(c.q, i, j, k, l));
}
// unchecked read of a cell known to be valid
CCJ_HD int ccj_get4u(const ccj_cx &c, int tbl, int i, int j, int k, int l) {
    if (c.q.shard_G == 0) return (int)ccj_t4(c, tbl)[ccj_idx4(c.q.n, i, j, k, l)];
    return (int)*ccj_addr4(c.q, tbl, ccj_pos_of(c.q, i, j, k, l));
}
// "if (min < INF/2) X.set(...)" + Matrix4D::set clamp + int16 narrowing; returns what a later get of
// this cell yields.  Every valid cell is written exactly once, so no table initialisation is needed.
CCJ_HD int ccj_put4(const ccj_cx &c, int tbl, const ccj_pos4 &pos, int mn) {
    int v = CCJ_INTERN_INF;
    if (mn < CCJ_INF / 2) {
        if (mn >= CCJ_INTERN_INF) mn = CCJ_INTERN_INF;
        v = (int)(int16_t)mn;
    }
    *ccj_addr4(c.q, tbl, pos) = (int16_t)v;
    return v;
}
CCJ_HD int ccj_raw2(const ccj_cx &c, int tbl, int i, int j) { return ccj_t2(c, tbl)[ccj_idx2(c.q.n, i, j)]; }
// TriangleMatrix::get with return_val INF (P, WBP, WPP)
CCJ_HD int ccj_tri_get(const ccj_cx &c, int tbl, int i, int j) {
    if (i > j) return CCJ_INF;
    return ccj_raw2(c, tbl, i, j);
}
// s_energy_matrix::get_energy / get_energy_WM / _WMv / _WMp (src/s_energy_matrix.hh:37-41)
CCJ_HD int ccj_V(const ccj_cx &c, int tbl, int i, int j) {
    if (i >= j) return CCJ_INF;
    return ccj_raw2(c, tbl, i, j);
}
// get_WB / get_WP (src/pseudo_loop.cc:647-661).  The stored T2_WB/T2_WP hold the i<=j case.
CCJ_HD int ccj_WBWP(const ccj_cx &c, int tbl, int i, int j) {
    const int n = c.q.n;
    if (i <= 0 || j <= 0 || i > n || j > n) return CCJ_INF;
    if (i > j) return 0;
    return ccj_raw2(c, tbl, i, j);
}
CCJ_HD int ccj_WB(const ccj_cx &c, int i, int j) { return ccj_WBWP(c, T2_WB, i, j); }
CCJ_HD int ccj_WP(const ccj_cx &c, int i, int j) { return ccj_WBWP(c, T2_WP, i, j); }

CCJ_HD int ccj_pt(const ccj_cx &c, int i, int j) { return ccj_ptype(c.M, c.q.S, i, j); }
// pseudo_loop::can_pair (src/pseudo_loop.hh:117-136)
CCJ_HD bool ccj_can_pair(const ccj_cx &c, int i, int j) { return (j - i > CCJ_TURN) && ccj_pt(c, i, j) > 0; }

// ---------------------------------------------------------------------------------------------
// nested part
// ---------------------------------------------------------------------------------------------
// s_energy_matrix::E_MLStem (src/s_energy_matrix.cc:54-112)
CCJ_HD int ccj_E_MLStem(const ccj_cx &c, int vij, int vi1j, int vij1, int vi1j1, int i, int j) {
    const ccj_model *P = c.M;
    const int8_t *S = c.q.S;
    const int n = c.q.n;
    int e = CCJ_INF, en;
    int type = ccj_pt(c, i, j);
    en = vij;
    if (en != CCJ_INF) {
        if (P->dangles == 2) {
            int mm5 = i > 1 ? S[i - 1] : -1;
            int mm3 = j < n ? S[j + 1] : -1;
            en += ccj_E_MLstem(P, type, mm5, mm3);
        } else {
            en += ccj_E_MLstem(P, type, -1, -1);
        }
        e = ccj_min(e, en);
    }
    if (P->dangles == 1) {
        const int mm5 = S[i], mm3 = S[j];
        en = (j - i - 1 > CCJ_TURN) ? vi1j : CCJ_INF;
        if (en != CCJ_INF) {
            en += P->MLbase;
            type = ccj_pt(c, i + 1, j);
            en += ccj_E_MLstem(P, type, mm5, -1);
            e = ccj_min(e, en);
        }
        en = (j - 1 - i > CCJ_TURN) ? vij1 : CCJ_INF;
        if (en != CCJ_INF) {
            en += P->MLbase;
            type = ccj_pt(c, i, j - 1);
            en += ccj_E_MLstem(P, type, -1, mm3);
            e = ccj_min(e, en);
        }
        en = (j - 1 - i - 1 > CCJ_TURN) ? vi1j1 : CCJ_INF;
        if (en != CCJ_INF) {
            en += 2 * P->MLbase;
            type = ccj_pt(c, i + 1, j - 1);
            en += ccj_E_MLstem(P, type, mm5, mm3);
            e = ccj_min(e, en);
        }
    }
    return e;
}
CCJ_HD int ccj_E_MLStem_at(const ccj_cx &c, int i, int j) {
    return ccj_E_MLStem(c, ccj_V(c, T2_V, i, j), ccj_V(c, T2_V, i + 1, j), ccj_V(c, T2_V, i, j - 1),
                        ccj_V(c, T2_V, i + 1, j - 1), i, j);
}

// s_energy_matrix::E_MbLoop (src/s_energy_matrix.cc:122-205)
CCJ_HD int ccj_E_MbLoop(const ccj_cx &c, int WM2ij, int WM2ip1j, int WM2ijm1, int WM2ip1jm1, int i, int j) {
    const ccj_model *P = c.M;
    const int8_t *S = c.q.S;
    int e = CCJ_INF, en = CCJ_INF;
    const int tt = ccj_pt(c, j, i);
    switch (P->dangles) {
        case 2:
            e = WM2ij;
            if (e != CCJ_INF) e += ccj_E_MLstem(P, tt, S[j - 1], S[i + 1]) + P->MLclosing;
            break;
        case 1:
            e = WM2ij;
            if (e != CCJ_INF) e += ccj_E_MLstem(P, tt, -1, -1) + P->MLclosing;
            en = WM2ip1j;
            if (en != CCJ_INF) en += ccj_E_MLstem(P, tt, -1, S[i + 1]) + P->MLclosing + P->MLbase;
            e = ccj_min(e, en);
            en = WM2ijm1;
            if (en != CCJ_INF) en += ccj_E_MLstem(P, tt, S[j - 1], -1) + P->MLclosing + P->MLbase;
            e = ccj_min(e, en);
            en = WM2ip1jm1;
            if (en != CCJ_INF) en += ccj_E_MLstem(P, tt, S[j - 1], S[i + 1]) + P->MLclosing + 2 * P->MLbase;
            e = ccj_min(e, en);
            break;
        case 0:
            e = WM2ij;
            if (e != CCJ_INF) e += ccj_E_MLstem(P, tt, -1, -1) + P->MLclosing;
            break;
    }
    return e;
}

// one k-term of compute_energy_VM (src/s_energy_matrix.cc:243-268)
CCJ_HD int ccj_VM_term(const ccj_cx &c, int i, int j, int k) {
    const int MLb = c.M->MLbase;
    const int wm1 = ccj_V(c, T2_WM, i + 1, k - 1), wm2 = ccj_V(c, T2_WM, i + 2, k - 1);
    const int wmv1 = ccj_V(c, T2_WMv, k, j - 1), wmp1 = ccj_V(c, T2_WMp, k, j - 1);
    const int wmv2 = ccj_V(c, T2_WMv, k, j - 2), wmp2 = ccj_V(c, T2_WMp, k, j - 2);
    int WM2ij = wm1 + wmv1;
    WM2ij = ccj_min(WM2ij, wm1 + wmp1);
    WM2ij = ccj_min(WM2ij, (k - i - 1) * MLb + wmp1);
    int WM2ip1j = wm2 + wmv1;
    WM2ip1j = ccj_min(WM2ip1j, wm2 + ccj_V(c, T2_WMp, k - 1, j - 1)); /* sic: k-1 (s_energy_matrix.cc:254) */
    WM2ip1j = ccj_min(WM2ip1j, (k - (i + 1) - 1) * MLb + wmp1);
    int WM2ijm1 = wm1 + wmv2;
    WM2ijm1 = ccj_min(WM2ijm1, wm1 + wmp2);
    WM2ijm1 = ccj_min(WM2ijm1, (k - i - 1) * MLb + wmp2);
    int WM2ip1jm1 = wm2 + wmv2;
    WM2ip1jm1 = ccj_min(WM2ip1jm1, wm2 + wmp2);
    WM2ip1jm1 = ccj_min(WM2ip1jm1, (k - (i + 1) - 1) * MLb + wmp2);
    return ccj_E_MbLoop(c, WM2ij, WM2ip1j, WM2ijm1, WM2ip1jm1, i, j);
}

// one (k,l) term of compute_internal (src/s_energy_matrix.cc:287-299) == compute_int + V(k,l)
CCJ_HD int ccj_Vint_term(const ccj_cx &c, int i, int j, int k, int l) {
    return ccj_compute_int(c.M, c.q.S, i, j, k, l) + ccj_V(c, T2_V, k, l);
}

// `Par` spreads independent candidates of one cell over cooperating lanes:  lane in [0,nlanes),
// red(v) returns the minimum over the lanes (identity when nlanes==1, e.g. on the host), sync() makes
// the lanes' earlier stores visible to each other.
// (value, position-in-reference-order) of the best candidate so far; lexicographic minimum == the
// reference's "first strictly smaller wins"
struct ccj_best {
    int val, ord;
};
CCJ_HD void ccj_cand(ccj_best &b, int val, int ord) {
    if (val < b.val || (val == b.val && ord < b.ord)) {
        b.val = val;
        b.ord = ord;
    }
}

struct ccj_serial {
    static constexpr int nlanes = 1;
    CCJ_HD int lane() const { return 0; }
    CCJ_HD int red(int v) const { return v; }
    CCJ_HD void sync() const {}
    CCJ_HD ccj_best argmin(ccj_best b) const { return b; }
};

// All nested/2D tables of one (i,j):  V -> (P given) -> WBP, WPP, WB, WP -> WMv, WMp -> WM, i.e. the order
// of W_final::ccj's loop body (src/W_final.cc:62-65) with compute_energies' 2D part (:73-77).
// `p_val` is P(i,j) as compute_P left it (already "stored if < INF/2", else INF+1).
template <class Par>
CCJ_HD void ccj_cell2d(const ccj_cx &c, int i, int j, const Par &par) {
    const ccj_model *M = c.M;
    const int n = c.q.n;
    const int L = par.lane(), NL = Par::nlanes;
    const int ij = ccj_idx2(n, i, j);
    int32_t *t2 = c.q.t2;
    const int64_t s2 = c.q.stride2;

    // ---- V(i,j): s_energy_matrix::compute_energy (src/s_energy_matrix.cc:315-358) ----
    {
        int eH = ccj_HairpinE(M, c.q.S, c.q.seq, i, j);
        int eI = CCJ_INF;
        {
            const int max_k = ccj_min(j - CCJ_TURN - 2, i + CCJ_MAXLOOP + 1);
            // flatten (k,l): k in [i+1,max_k], l in [min_l(k), j-1]
            for (int k = i + 1 + L; k <= max_k; k += NL) {
                const int min_l = ccj_max(k + CCJ_TURN + 1 + CCJ_MAXLOOP + 2, k + j - i) - CCJ_MAXLOOP - 2;
                for (int l = j - 1; l >= min_l; --l) eI = ccj_min(eI, ccj_Vint_term(c, i, j, k, l));
            }
            eI = par.red(eI);
        }
        int eM = CCJ_INF;
        for (int k = i + 1 + L; k <= j - 3; k += NL) eM = ccj_min(eM, ccj_VM_term(c, i, j, k));
        eM = par.red(eM);
        int mn = CCJ_INF / 2, rank = -1;
        if (eH < mn) { mn = eH; rank = 0; }
        if (eI < mn) { mn = eI; rank = 1; }
        if (eM < mn) { mn = eM; rank = 2; }
        if (L == 0 && mn < CCJ_INF / 2) {
            t2[T2_V * s2 + ij] = mn;
            t2[T2_VTYPE * s2 + ij] = rank == 0 ? 'H' : (rank == 1 ? 'I' : 'M');
        }
    }
    par.sync();  // every lane of this cell must see V(i,j) below (WBP's d==i term, WM's k==i term)
    const int p_val = ccj_raw2(c, T2_P, i, j);

    // ---- WBP / WPP (src/pseudo_loop.cc:134-164) ----
    {
        int b1 = CCJ_INF, b2 = CCJ_INF, w1 = CCJ_INF, w2 = CCJ_INF;
        for (int d = i + L; d < j; d += NL) {
            const int wb = ccj_WB(c, i, d - 1), wp = ccj_WP(c, i, d - 1);
            const int v = ccj_V(c, T2_V, d, j);
            const int p = ccj_tri_get(c, T2_P, d, j);
            b1 = ccj_min(b1, wb + v + M->bp_penalty + M->PPS_penalty);
            b2 = ccj_min(b2, wb + p + M->PSM_penalty + M->PPS_penalty);
            w1 = ccj_min(w1, wp + v + M->PPS_penalty);
            w2 = ccj_min(w2, wp + p + M->PSP_penalty + M->PPS_penalty);
        }
        b1 = par.red(b1); b2 = par.red(b2); w1 = par.red(w1); w2 = par.red(w2);
        const int b3 = ccj_tri_get(c, T2_WBP, i, j - 1) + M->cp_penalty;
        const int w3 = ccj_tri_get(c, T2_WPP, i, j - 1) + M->PUP_penalty;
        const int wbp = ccj_min(ccj_min(b1, b2), b3);
        const int wpp = ccj_min(ccj_min(w1, w2), w3);
        if (L == 0) {
            int wbp_s = CCJ_INF + 1, wpp_s = CCJ_INF + 1;
            if (wbp < CCJ_INF / 2) wbp_s = wbp;
            if (wpp < CCJ_INF / 2) wpp_s = wpp;
            t2[T2_WBP * s2 + ij] = wbp_s;
            t2[T2_WPP * s2 + ij] = wpp_s;
            const int wb = ccj_min(M->cp_penalty * (j - i + 1), wbp_s);
            const int wp = ccj_min(M->PUP_penalty * (j - i + 1), wpp_s);
            t2[T2_WB * s2 + ij] = wb;
            t2[T2_WP * s2 + ij] = wp;
            if (c.q.w3) {  // packed {WB,WP,WBP,-} for the tuned 4D kernel
                int32_t *w = c.q.w3 + 4 * (int64_t)ij;
                w[0] = wb; w[1] = wp; w[2] = wbp_s; w[3] = 0;
            }
        }
    }

    if (j - i + 1 < 4) return;

    // ---- WMv / WMp (src/s_energy_matrix.cc:206-217) ----
    if (L == 0) {
        int wmv = ccj_E_MLStem_at(c, i, j);
        int wmp = p_val + M->PSM_penalty + M->b_penalty;
        wmv = ccj_min(wmv, ccj_raw2(c, T2_WMv, i, j - 1) + M->MLbase);
        wmp = ccj_min(wmp, ccj_raw2(c, T2_WMp, i, j - 1) + M->MLbase);
        t2[T2_WMv * s2 + ij] = wmv;
        t2[T2_WMp * s2 + ij] = wmp;
    }
    // ---- WM (src/s_energy_matrix.cc:219-241) ----
    {
        int m1 = CCJ_INF, m2 = CCJ_INF, m3 = CCJ_INF, m4 = CCJ_INF;
        for (int k = j - CCJ_TURN - 1 - L; k >= i; k -= NL) {
            const int wm_kj = ccj_E_MLStem_at(c, k, j);
            const int wmb_kj = ccj_raw2(c, T2_P, k, j) + M->PSM_penalty + M->b_penalty;
            const int wm_ik = ccj_V(c, T2_WM, i, k - 1);
            m1 = ccj_min(m1, (k - i) * M->MLbase + wm_kj);
            m2 = ccj_min(m2, (k - i) * M->MLbase + wmb_kj);
            m3 = ccj_min(m3, wm_ik + wm_kj);
            m4 = ccj_min(m4, wm_ik + wmb_kj);
        }
        m1 = par.red(m1); m2 = par.red(m2); m3 = par.red(m3); m4 = par.red(m4);
        if (L == 0) {
            const int m5 = ccj_raw2(c, T2_WM, i, j - 1) + M->MLbase;
            t2[T2_WM * s2 + ij] = ccj_min(ccj_min(ccj_min(m1, m2), ccj_min(m3, m4)), m5);
        }
    }
}
