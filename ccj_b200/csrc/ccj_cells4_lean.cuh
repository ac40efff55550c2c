// Lean form of ccj_cell4d (ccj_cells4.cuh) for the kernels that run one thread per cell: the row-sharded fold of one
// oversized sequence (k_4d_shard, ccj_shard.cu) and sequences beyond the tuned range (k_4d, ccj_kernels.cu).
//
// Same candidates, same order of the same-cell terms, same stores as ccj_cell4d (src/pseudo_loop.cc:85-127,181-644) --
// what changes is the addressing.  ccj_cell4d resolves every table read through ccj_pos_of / ccj_addr4 (two level-base
// loads, the closed-form row offsets with both the power-of-two and the division form, a kind lookup and a 64-bit
// multiply-add per table: ~350 SASS instructions per split point, profiles/r2_notes.md).  Here a layout policy computes
// ONE position per split point -- level base pointers plus a 32-bit in-level offset -- and every table read at that
// position is `base[kind * C + offset]` with a compile-time kind; the level bases come from a small table the kernel
// keeps in shared memory.  The two forms are checked against each other and against the reference's table hashes on
// the CPU (tests/emu, tests/test_shard_layout.py) and on the GPU (tests/test_gpu_shard.py).
#pragma once
#include "ccj_cells4.cuh"

// kind of a table in the sharded layout as a compile-time constant; must equal ccj_shard_kinds (ccj_lean_kinds_ok)
CCJ_HD constexpr int ccj_lean_kind(int tbl) {
    return tbl == T_PK ? 0 : tbl == T_PL ? 1 : tbl == T_PO ? 2 : tbl == T_PfromL ? 3 : tbl == T_PfromO ? 4
         : tbl == T_PLmloop00 ? 5 : tbl == T_PLmloop01 ? 6 : tbl == T_PLmloop10 ? 7 : tbl == T_PMmloop00 ? 8
         : tbl == T_POmloop00 ? 9 : tbl == T_POmloop01 ? 10 : tbl == T_POmloop10 ? 11
         : tbl == T_PR ? 12 : tbl == T_PM ? 13 : tbl == T_PfromR ? 14 : tbl == T_PfromM ? 15 : tbl == T_PfromMprime ? 16
         : tbl == T_PRmloop00 ? 17 : tbl == T_PRmloop01 ? 18 : tbl == T_PRmloop10 ? 19 : tbl == T_PMmloop01 ? 20 : 21;
}
inline bool ccj_lean_kinds_ok(const int8_t *kind24) {
    for (int t = 0; t < CCJ_NT4; ++t)
        if (kind24[t] != ccj_lean_kind(t)) return false;
    return true;
}

// ---- layout policies ------------------------------------------------------------------------------------------------
// row(i): what is fixed along row i;  pos(row, a', b', kk'): the cell (i', i'+a', i'+a'+2+kk', ...+b') of that row;
// rd<T>(pos) / wr<T>(pos, v): entry of table T there.

// Row-sharded layout (ccj_types.h): rep[12 G lev[t] + (12 r + kind) C(t) + inner], loc_r[10 lev[t] + (kind-12) C(t) + inner].
// Row-local tables are only ever read along the thread's own row (that is what makes them row-local), so `loc` is the
// own rank's base.  In-level offsets are 32-bit: the launcher checks (12 G + 1) C_max < 2^31.  The level bases come
// premultiplied from a table (shared memory on the device).  POW2: G is a power of two (shift / mask instead of
// divisions in the row arithmetic).
struct ccj_lean_lvl {
    int64_t LG, L10;   // 12 G lev[t], 10 lev[t]
    int32_t C, pad;    // cells one table of one rank reserves on level t
};
CCJ_HD void ccj_lean_lvl_fill(ccj_lean_lvl *tab, const int64_t *lev, int n, int G, int first, int step) {
    for (int t = first; t <= n; t += step) {
        ccj_lean_lvl v;
        v.LG = lev[t] * (CCJ_SHARD_NREP * G);
        v.L10 = lev[t] * CCJ_SHARD_NLOC;
        v.C = (int32_t)(lev[t + 1] - lev[t]);
        v.pad = 0;
        tab[t] = v;
    }
}
CCJ_HD int ccj_lean_ld16(const int16_t *p) {
#if defined(__CUDA_ARCH__)
    return (int)__ldg(p);   // sources are cells of lower levels: not written during this launch
#else
    return (int)*p;
#endif
}
template <bool POW2>
struct ccj_lean_shard {
    int16_t *rep, *loc;
    const ccj_lean_lvl *lvl;   // lvl[0..n]
    int n, G, sh;
    struct Row { int r, q, qt; };                 // rank, index among the rank's rows, G q(q-1)/2
    struct Pos { int16_t *rb, *lb; int C, o; };   // level bases (rb already at the row's rank), cells per table, inner
    CCJ_HD Row row(int i) const {
        Row w;
        if (POW2) { w.r = (i - 1) & (G - 1); w.q = (i - 1) >> sh; w.qt = ((w.q * (w.q - 1)) >> 1) << sh; }
        else { w.r = (i - 1) % G; w.q = (i - 1) / G; w.qt = G * (w.q * (w.q - 1) / 2); }
        return w;
    }
    CCJ_HD Pos pos(const Row &w, int ap, int bp, int kk) const {
        const int tt = ap + bp, mr = n - tt - 2 - w.r;
        int Q, slab;
        if (POW2) { Q = (mr + G - 1) >> sh; slab = Q * mr - (((Q * (Q - 1)) >> 1) << sh); }
        else { Q = (mr + G - 1) / G; slab = Q * mr - G * (Q * (Q - 1) / 2); }
        const ccj_lean_lvl lv = lvl[tt];
        Pos p;
        p.C = lv.C;
        p.o = ap * slab + w.q * mr - w.qt + kk;
        p.rb = rep + (lv.LG + (int64_t)(CCJ_SHARD_NREP * w.r) * lv.C);
        p.lb = loc + lv.L10;
        return p;
    }
    template <int TBL> CCJ_HD int16_t *at(const Pos &p) const {
        constexpr int kd = ccj_lean_kind(TBL);
        return kd < CCJ_SHARD_NREP ? p.rb + (kd * p.C + p.o) : p.lb + ((kd - CCJ_SHARD_NREP) * p.C + p.o);
    }
    template <int TBL> CCJ_HD int rd(const Pos &p) const { return ccj_lean_ld16(at<TBL>(p)); }
    template <int TBL> CCJ_HD void wr(const Pos &p, int v) const { *at<TBL>(p) = (int16_t)v; }
    // cells (row, a'+x, b'-x, kk') of ONE level: entry x of table TBL is diag.p[x * diag.stride] (the rank's slab size)
    struct Diag { const int16_t *p; int stride; };
    template <int TBL> CCJ_HD Diag diag(const Row &w, int ap, int bp, int kk) const {
        const int mr = n - (ap + bp) - 2 - w.r;
        int Q;
        if (POW2) Q = (mr + G - 1) >> sh; else Q = (mr + G - 1) / G;
        Diag d;
        d.p = at<TBL>(pos(w, ap, bp, kk));
        d.stride = Q * mr - G * (Q * (Q - 1) / 2);
        return d;
    }
};

// Ordinary layout (ccj_idx4): t4[tbl * stride4 + Cb(b') - Tet(m') + (i'-1)(2m'+2-i')/2 + kk'], Cb / Tet from a table
// (tab[0..n] = Cb(b), tab[n+1 + m] = Tet(m)) so that a position is two table reads and a handful of integer operations.
struct ccj_lean_plain {
    int16_t *t4;
    int64_t st4;
    const int64_t *tab;
    int n;
    struct Row { int i; };
    struct Pos { int16_t *p; };
    CCJ_HD Row row(int i) const { Row w; w.i = i; return w; }
    CCJ_HD Pos pos(const Row &w, int ap, int bp, int kk) const {
        const int m = n - ap - bp - 2;
        Pos p;
        p.p = t4 + (tab[bp] - tab[n + 1 + m] + (((w.i - 1) * (2 * m + 2 - w.i)) >> 1) + kk);
        return p;
    }
    template <int TBL> CCJ_HD int rd(const Pos &p) const { return ccj_lean_ld16(p.p + TBL * st4); }
    template <int TBL> CCJ_HD void wr(const Pos &p, int v) const { p.p[TBL * st4] = (int16_t)v; }
};
// fills tab for ccj_lean_plain; entries x = first, first+step, ...
CCJ_HD void ccj_lean_plain_tab(int64_t *tab, int n, int first, int step) {
    for (int x = first; x <= n; x += step) {
        tab[x] = x <= n - 3 ? ccj_cb(n, x) : 0;
        tab[n + 1 + x] = ccj_tet(x);
    }
}

CCJ_HD ccj_w3v ccj_lean_w3(const int32_t *w3, int idx) {
    ccj_w3v r;
#if defined(__CUDA_ARCH__)
    const int4 v = __ldg(reinterpret_cast<const int4 *>(w3) + idx);
    r.wb = v.x; r.wp = v.y; r.wbp = v.z;
#else
    const int32_t *w = w3 + 4 * (int64_t)idx;
    r.wb = w[0]; r.wp = w[1]; r.wbp = w[2];
#endif
    return r;
}

#ifndef CCJ_LEAN_UNROLL
#define CCJ_LEAN_UNROLL 2
#endif
#define CCJ_LEAN_UNROLL_PRAGMA CCJ_UNROLL_BY(CCJ_LEAN_UNROLL)
// the window loops are chains of two dependent loads per candidate (list entry -> source cell): several in flight
#ifndef CCJ_LEAN_WIN_UNROLL
#define CCJ_LEAN_WIN_UNROLL 8
#endif
#define CCJ_LEAN_WIN_PRAGMA CCJ_UNROLL_BY(CCJ_LEAN_WIN_UNROLL)
#ifndef CCJ_LEAN_ABLATE_WIN   // timing experiments only: 1 = skip the window lists (wrong tables)
#define CCJ_LEAN_ABLATE_WIN 0
#endif
CCJ_HD uint32_t ccj_lean_ldu(const uint32_t *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// Requires c.q.w3 (packed {WB,WP,WBP}) and the partner lists (ccj_lists_ok); callers fall back to ccj_cell4d otherwise.
template <class LY>
CCJ_HD void ccj_cell4d_lean(const ccj_cx &c, const LY &ly, int i, int j, int k, int l) {
    const ccj_model *M = c.M;
    const ccj_seq &q = c.q;
    const int n = q.n, n1 = n + 1;
    const int INF = CCJ_INF;
    const int bp = M->bp_penalty, cp = M->cp_penalty, PB = M->PB_penalty, apbp = M->ap_penalty + M->bp_penalty;
    const int a = j - i, b = l - k, kk = k - j - 2;
    const int32_t *w3 = q.w3;
    const typename LY::Row own = ly.row(i);
    typedef typename LY::Pos Pos;

    int PLm00 = CCJ_INTERN_INF + bp, PLm01 = INF, PLm10 = INF;   // PX(i,j,k,l) is still unset where P?mloop00 reads it
    int PRm00 = CCJ_INTERN_INF + bp, PRm01 = INF, PRm10 = INF;
    int PMm00 = CCJ_INTERN_INF + bp, PMm01 = INF, PMm10 = INF;
    int POm00 = CCJ_INTERN_INF + bp, POm01 = INF, POm10 = INF;
    int PfL = INF, PfR = INF, PfM = INF, PfMp = INF, PfO = INF, PK = INF;
    if (b >= 1) {
        PRm01 = ly.template rd<T_PRmloop01>(ly.pos(own, a, b - 1, kk)) + cp;      // (i,j,k,l-1)   :517-519
        const Pos p = ly.pos(own, a, b - 1, kk + 1);                              // (i,j,k+1,l)
        PRm10 = ly.template rd<T_PRmloop10>(p) + cp;                              // :531-533
        PMm01 = ly.template rd<T_PMmloop01>(p) + cp;                              // :564-566
    } else {
        PRm01 = PRm10 = PMm01 = INF + cp;
    }
    PMm10 = (a >= 1 ? ly.template rd<T_PMmloop10>(ly.pos(own, a - 1, b, kk + 1)) : INF) + cp;   // (i,j-1,k,l)   :578-580

    // ---- L1: X(i,d,k,l) with (d+1,j), d = i .. j-1: a' = d-i, same row, kk' = k-d-2 ----
    if (a >= 1) {
        {   // d = i
            const Pos p = ly.pos(own, 0, b, kk + a);
            const ccj_w3v w = ccj_lean_w3(w3, (a - 1) * n1 + i + 1);
            const int x00 = ly.template rd<T_PLmloop00>(p);
            PLm00 = ccj_min(PLm00, x00 + w.wb);                                  // :455-458
            PLm01 = ccj_min(PLm01, x00 + w.wbp);                                 // :468-471
            PMm00 = ccj_min(PMm00, ly.template rd<T_PMmloop00>(p) + w.wb);       // :548-551
        }
        CCJ_LEAN_UNROLL_PRAGMA
        for (int ap = 1; ap < a; ++ap) {
            const Pos p = ly.pos(own, ap, b, kk + a - ap);
            const ccj_w3v w = ccj_lean_w3(w3, (a - ap - 1) * n1 + i + ap + 1);
            const int x00 = ly.template rd<T_PLmloop00>(p);
            PLm00 = ccj_min(PLm00, x00 + w.wb);
            PLm01 = ccj_min(PLm01, x00 + w.wbp);
            PMm00 = ccj_min(PMm00, ly.template rd<T_PMmloop00>(p) + w.wb);
            PLm10 = ccj_min(PLm10, ly.template rd<T_PLmloop10>(p) + w.wb);       // :484-487
            PfL = ccj_min(PfL, ly.template rd<T_PfromL>(p) + w.wp);              // :360-361
            PfM = ccj_min(PfM, ly.template rd<T_PfromMprime>(p) + w.wp);         // :399-402
            PK = ccj_min(PK, ly.template rd<T_PK>(p) + w.wp);                    // :184-187
        }
        // ---- L2: X(d,j,k,l) with (i,d-1), d = i+1 .. j: row d, a' = j-d, kk' = kk ----
        CCJ_LEAN_UNROLL_PRAGMA
        for (int d = i + 1; d < j; ++d) {
            const Pos p = ly.pos(ly.row(d), j - d, b, kk);
            const ccj_w3v w = ccj_lean_w3(w3, (d - 1 - i) * n1 + i);
            const int x00 = ly.template rd<T_PLmloop00>(p), o00 = ly.template rd<T_POmloop00>(p);
            PLm00 = ccj_min(PLm00, w.wb + x00);                                  // :450-453
            PLm10 = ccj_min(PLm10, w.wbp + x00);                                 // :481-483
            PMm10 = ccj_min(PMm10, w.wbp + ly.template rd<T_PMmloop00>(p));      // :581-584
            POm00 = ccj_min(POm00, w.wb + o00);                                  // :599-602
            POm10 = ccj_min(POm10, w.wbp + o00);                                 // :632-635
            PfL = ccj_min(PfL, ly.template rd<T_PfromL>(p) + w.wp);              // :357-359
            PfO = ccj_min(PfO, ly.template rd<T_PfromO>(p) + w.wp);              // :425-428
        }
        {   // d = j
            const Pos p = ly.pos(ly.row(j), 0, b, kk);
            const ccj_w3v w = ccj_lean_w3(w3, (a - 1) * n1 + i);
            const int x00 = ly.template rd<T_PLmloop00>(p), o00 = ly.template rd<T_POmloop00>(p);
            PLm00 = ccj_min(PLm00, w.wb + x00);
            PLm10 = ccj_min(PLm10, w.wbp + x00);
            PMm10 = ccj_min(PMm10, w.wbp + ly.template rd<T_PMmloop00>(p));
            POm00 = ccj_min(POm00, w.wb + o00);
            POm10 = ccj_min(POm10, w.wbp + o00);
        }
    }
    if (b >= 1) {
        // ---- R3: X(i,j,d,l) with (k,d-1), d = k+1 .. l: b' = l-d, same row, kk' = d-j-2 ----
        CCJ_LEAN_UNROLL_PRAGMA
        for (int d = k + 1; d < l; ++d) {
            const Pos p = ly.pos(own, a, l - d, kk + (d - k));
            const ccj_w3v w = ccj_lean_w3(w3, (d - 1 - k) * n1 + k);
            const int r00 = ly.template rd<T_PRmloop00>(p);
            PRm00 = ccj_min(PRm00, w.wb + r00);                                  // :499-503
            PRm10 = ccj_min(PRm10, w.wbp + r00);                                 // :534-537
            PMm00 = ccj_min(PMm00, ly.template rd<T_PMmloop00>(p) + w.wb);       // :552-555
            PfR = ccj_min(PfR, ly.template rd<T_PfromR>(p) + w.wp);              // :379-381
            // get_PfromMdoubleprime (:663-679); d<l, so its i==j&&k==l base case cannot occur
            PfMp = ccj_min(PfMp, ccj_min(ly.template rd<T_PL>(p) + PB, ly.template rd<T_PR>(p) + PB) + w.wp);   // :412-415
            PK = ccj_min(PK, ly.template rd<T_PK>(p) + w.wp);                    // :189-192
        }
        {   // d = l
            const Pos p = ly.pos(own, a, 0, kk + b);
            const ccj_w3v w = ccj_lean_w3(w3, (b - 1) * n1 + k);
            const int r00 = ly.template rd<T_PRmloop00>(p);
            PRm00 = ccj_min(PRm00, w.wb + r00);
            PRm10 = ccj_min(PRm10, w.wbp + r00);
            PMm00 = ccj_min(PMm00, ly.template rd<T_PMmloop00>(p) + w.wb);
        }
        // ---- R4: X(i,j,k,d) with (d+1,l), d = k .. l-1: b' = d-k, same row, kk' = kk ----
        {   // d = k
            const Pos p = ly.pos(own, a, 0, kk);
            const ccj_w3v w = ccj_lean_w3(w3, (b - 1) * n1 + k + 1);
            const int r00 = ly.template rd<T_PRmloop00>(p), o00 = ly.template rd<T_POmloop00>(p);
            PRm00 = ccj_min(PRm00, r00 + w.wb);                                  // :504-507
            PRm01 = ccj_min(PRm01, r00 + w.wbp);                                 // :520-523
            PMm01 = ccj_min(PMm01, ly.template rd<T_PMmloop00>(p) + w.wbp);      // :567-570
            POm00 = ccj_min(POm00, o00 + w.wb);                                  // :603-606
            POm01 = ccj_min(POm01, o00 + w.wbp);                                 // :618-621
        }
        CCJ_LEAN_UNROLL_PRAGMA
        for (int bq = 1; bq < b; ++bq) {
            const Pos p = ly.pos(own, a, bq, kk);
            const ccj_w3v w = ccj_lean_w3(w3, (b - bq - 1) * n1 + k + bq + 1);
            const int r00 = ly.template rd<T_PRmloop00>(p), o00 = ly.template rd<T_POmloop00>(p);
            PRm00 = ccj_min(PRm00, r00 + w.wb);
            PRm01 = ccj_min(PRm01, r00 + w.wbp);
            PMm01 = ccj_min(PMm01, ly.template rd<T_PMmloop00>(p) + w.wbp);
            POm00 = ccj_min(POm00, o00 + w.wb);
            POm01 = ccj_min(POm01, o00 + w.wbp);
            PMm10 = ccj_min(PMm10, ly.template rd<T_PMmloop10>(p) + w.wb);       // :585-588
            POm10 = ccj_min(POm10, ly.template rd<T_POmloop10>(p) + w.wb);       // :636-639
            PfR = ccj_min(PfR, ly.template rd<T_PfromR>(p) + w.wp);              // :382-383
            PfO = ccj_min(PfO, ly.template rd<T_PfromO>(p) + w.wp);              // :429-432
        }
    }

    // ---- stores, in the reference's in-cell order (:85-127); "if (min < INF/2) set", clamp, int16 narrowing ----
    const Pos self = ly.pos(own, a, b, kk);
#define CCJ_LEAN_PUT(TBL, val) ([&](int mn_) -> int {                                   \
        int v_ = CCJ_INTERN_INF;                                                        \
        if (mn_ < CCJ_INF / 2) { if (mn_ >= CCJ_INTERN_INF) mn_ = CCJ_INTERN_INF; v_ = (int)(int16_t)mn_; }   \
        ly.template wr<TBL>(self, v_);                                                  \
        return v_; })(val)
    CCJ_LEAN_PUT(T_PLmloop00, PLm00);
    CCJ_LEAN_PUT(T_PLmloop01, PLm01);
    CCJ_LEAN_PUT(T_PLmloop10, PLm10);
    CCJ_LEAN_PUT(T_PRmloop00, PRm00);
    CCJ_LEAN_PUT(T_PRmloop01, PRm01);
    CCJ_LEAN_PUT(T_PRmloop10, PRm10);
    CCJ_LEAN_PUT(T_PMmloop00, PMm00);
    CCJ_LEAN_PUT(T_PMmloop01, PMm01);
    CCJ_LEAN_PUT(T_PMmloop10, PMm10);
    CCJ_LEAN_PUT(T_POmloop00, POm00);
    CCJ_LEAN_PUT(T_POmloop01, POm01);
    CCJ_LEAN_PUT(T_POmloop10, POm10);
    const int8_t *S = q.S;
    auto estP = [&](int x, int y) -> int { return q.estP ? q.estP[ccj_idx2(n, x, y)] : ccj_e_stP(M, S, x, y); };
    int mn;
    // ---- PL (:232-253), get_PLiloop (:682-703) over the partner list of (i,j) ----
    mn = INF;
    if (ccj_pt(c, i, j) > 0) {
        const typename LY::Row r1 = ly.row(i + 1);
        if (a > CCJ_TURN) {   // can_pair(i,j)
            if (a > CCJ_TURN + 2) mn = ly.template rd<T_PL>(ly.pos(r1, a - 2, b, kk + 1)) + estP(i, j);
            const int slot = ccj_tri(i, j);
            const uint32_t *lst = q.inlist + (int64_t)slot * CCJ_WIN_IN;
            const int cnt = CCJ_LEAN_ABLATE_WIN ? 0 : q.incnt[slot];
            CCJ_LEAN_WIN_PRAGMA
            for (int e = 0; e < cnt; ++e) {
                const uint32_t en = ccj_lean_ldu(lst + e);
                const int x = (en >> 16) & 0xff, y = en >> 24;   // PL(i+x, j-y, k, l)
                mn = ccj_min(mn, (int)(int16_t)(en & 0xffff) + ly.template rd<T_PL>(ly.pos(ly.row(i + x), a - x - y, b, kk + y)));
            }
        }
        if (a >= 2) {
            const Pos p = ly.pos(r1, a - 2, b, kk + 1);   // (i+1,j-1,k,l)
            mn = ccj_min(mn, ccj_min(ly.template rd<T_PLmloop10>(p), ly.template rd<T_PLmloop01>(p)) + apbp + bp);
            if (a >= CCJ_TURN + 1) mn = ccj_min(mn, ly.template rd<T_PfromL>(p));
        } else {
            mn = ccj_min(mn, INF + apbp + bp);
        }
    }
    const int vPL = CCJ_LEAN_PUT(T_PL, mn);
    // ---- PR (:255-275), get_PRiloop (:717-738) over the partner list of (k,l) ----
    mn = INF;
    if (ccj_pt(c, k, l) > 0) {
        if (b > CCJ_TURN) {
            if (b > CCJ_TURN + 2) mn = ly.template rd<T_PR>(ly.pos(own, a, b - 2, kk + 1)) + estP(k, l);
            const int slot = ccj_tri(k, l);
            const uint32_t *lst = q.inlist + (int64_t)slot * CCJ_WIN_IN;
            const int cnt = CCJ_LEAN_ABLATE_WIN ? 0 : q.incnt[slot];
            CCJ_LEAN_WIN_PRAGMA
            for (int e = 0; e < cnt; ++e) {
                const uint32_t en = ccj_lean_ldu(lst + e);
                const int x = (en >> 16) & 0xff, y = en >> 24;   // PR(i, j, k+x, l-y)
                mn = ccj_min(mn, (int)(int16_t)(en & 0xffff) + ly.template rd<T_PR>(ly.pos(own, a, b - x - y, kk + x)));
            }
        }
        if (b >= 2) {
            const Pos p = ly.pos(own, a, b - 2, kk + 1);   // (i,j,k+1,l-1)
            mn = ccj_min(mn, ccj_min(ly.template rd<T_PRmloop10>(p), ly.template rd<T_PRmloop01>(p)) + apbp + bp);
            if (b >= CCJ_TURN + 1) mn = ccj_min(mn, ly.template rd<T_PfromR>(p));
        } else {
            mn = ccj_min(mn, INF + apbp + bp);
        }
    }
    const int vPR = CCJ_LEAN_PUT(T_PR, mn);
    // ---- PM (:277-300), get_PMiloop (:752-773) over the partners outside (j,k) ----
    mn = INF;
    if (ccj_pt(c, j, k) > 0) {
        if (k - j > CCJ_TURN) {
            if (a >= 1 && b >= 1) mn = ly.template rd<T_PM>(ly.pos(own, a - 1, b - 1, kk + 2)) + estP(j - 1, k + 1);
            const int slot = ccj_tri(j, k);
            const uint32_t *lst = q.outlist + (int64_t)slot * CCJ_WIN_OUT * 2;   // 8-byte entries
            const int cnt = CCJ_LEAN_ABLATE_WIN ? 0 : q.outcnt[slot] & 0xffff;
            CCJ_LEAN_WIN_PRAGMA
            for (int e = 0; e < cnt; ++e) {
                const uint32_t en = ccj_lean_ldu(lst + 2 * e);
                const int x = (en >> 16) & 0xff, y = en >> 24;   // PM(i, j-x, k+y, l), d > i and dp < l
                if (x < a && y < b)
                    mn = ccj_min(mn, (int)(int16_t)(en & 0xffff) + ly.template rd<T_PM>(ly.pos(own, a - x, b - y, kk + x + y)));
            }
        }
        if (a >= 1 && b >= 1) {
            const Pos p = ly.pos(own, a - 1, b - 1, kk + 2);   // (i,j-1,k+1,l)
            mn = ccj_min(mn, ccj_min(ly.template rd<T_PMmloop10>(p), ly.template rd<T_PMmloop01>(p)) + apbp + bp);
            if (k >= j + CCJ_TURN - 1) mn = ccj_min(mn, ly.template rd<T_PfromM>(p));
        } else {
            mn = ccj_min(mn, INF + apbp + bp);
        }
        if (a == 0 && b == 0) mn = ccj_min(mn, 0);
    }
    const int vPM = CCJ_LEAN_PUT(T_PM, mn);
    // ---- PO (:302-322); the window of get_POiloop is dead (:787-808), only its stacking term counts ----
    mn = INF;
    if (ccj_pt(c, i, l) > 0) {
        if (a >= 1 && b >= 1) {
            const Pos p = ly.pos(ly.row(i + 1), a - 1, b - 1, kk);   // (i+1,j,k,l-1)
            if (l - i > CCJ_TURN) mn = ly.template rd<T_PO>(p) + estP(i, l);
            mn = ccj_min(mn, ccj_min(ly.template rd<T_POmloop10>(p), ly.template rd<T_POmloop01>(p)) + apbp + bp);
            if (l >= i + CCJ_TURN + 1) mn = ccj_min(mn, ly.template rd<T_PfromO>(p));
        } else {
            mn = ccj_min(mn, INF + apbp + bp);
        }
    }
    const int vPO = CCJ_LEAN_PUT(T_PO, mn);
    // ---- PfromL / PfromR / PfromM / PfromMprime / PfromO / PK: the same-cell terms (:354-443, :181-202) ----
    CCJ_LEAN_PUT(T_PfromL, ccj_min(PfL, ccj_min(vPR + PB, ccj_min(vPM + PB, vPO + PB))));
    CCJ_LEAN_PUT(T_PfromR, ccj_min(PfR, ccj_min(vPM + PB, vPO + PB)));
    CCJ_LEAN_PUT(T_PfromM, PfM);
    CCJ_LEAN_PUT(T_PfromMprime, PfMp);
    CCJ_LEAN_PUT(T_PfromO, ccj_min(PfO, ccj_min(vPL + PB, vPR + PB)));
    CCJ_LEAN_PUT(T_PK, ccj_min(PK, ccj_min(ccj_min(vPL + PB, vPM + PB), ccj_min(vPR + PB, vPO + PB))));
#undef CCJ_LEAN_PUT
}

// compute_P (src/pseudo_loop.cc:166-179): the share of one (i, j, l) that lane `lane` of warp `wid` evaluates,
//     min over d, k of PK(i,j,d+1,k) + PK(j+1,d,k+1,l),   j < d < k < l.
// Warps take delta = k-d, lanes walk d: the first factors (i,j,d+1,d+delta) are then consecutive entries of row i of slab
// (j-i, delta-1), and the second factors (j+1,d,d+delta+1,l) are the cells (a''=d-j-1, b''=l-d-delta-1, kk=delta-1) of row
// j+1 on the ONE level l-j-delta-2 -- in the sharded layout a constant stride (the rank's slab size) apart.  Two loads,
// one add, one min per term; ccj_P_term resolves both cells from scratch.
template <bool POW2>
CCJ_HD int ccj_P_lean(const ccj_lean_shard<POW2> &ly, int i, int j, int l, int wid, int nwarp, int lane, int nlanes) {
    int mn = CCJ_INF;
    const typename ccj_lean_shard<POW2>::Row r1 = ly.row(i), r2 = ly.row(j + 1);
    for (int dl = 1 + wid; dl <= l - 2 - j; dl += nwarp) {
        const int nd = l - 1 - dl - j;   // d = j+1 .. l-1-dl
        const int16_t *f = ly.template at<T_PK>(ly.pos(r1, j - i, dl - 1, 0));                     // d = j+1: kk = 0
        const typename ccj_lean_shard<POW2>::Diag g = ly.template diag<T_PK>(r2, 0, l - j - dl - 2, dl - 1);   // d = j+1: a'' = 0
        for (int x = lane; x < nd; x += nlanes) mn = ccj_min(mn, ccj_lean_ld16(f + x) + ccj_lean_ld16(g.p + (int64_t)x * g.stride));
    }
    return mn;
}
// ordinary layout: the second factors of one delta are on one level but not equidistant -- one position per term
CCJ_HD int ccj_P_lean(const ccj_lean_plain &ly, int i, int j, int l, int wid, int nwarp, int lane, int nlanes) {
    int mn = CCJ_INF;
    const ccj_lean_plain::Row r1 = ly.row(i), r2 = ly.row(j + 1);
    for (int dl = 1 + wid; dl <= l - 2 - j; dl += nwarp) {
        const int nd = l - 1 - dl - j;
        const int16_t *f = ly.pos(r1, j - i, dl - 1, 0).p + T_PK * ly.st4;
        for (int x = lane; x < nd; x += nlanes)
            mn = ccj_min(mn, ccj_lean_ld16(f + x) + ly.template rd<T_PK>(ly.pos(r2, x, l - j - dl - 2 - x, dl - 1)));
    }
    return mn;
}
