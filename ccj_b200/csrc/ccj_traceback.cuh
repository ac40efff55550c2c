// Traceback over the filled tables: an explicit LIFO stack of seq_interval nodes, every choice the
// FIRST minimum in the reference's source order ("if (tmp < min)").  This restates the reference's
// *traceback* code, which differs from its fill in many places -- all
// of those differences, including the ones that end in exit(EXIT_FAILURE) or in
// printf("Should not be here!"), are kept because they are observable behaviour.
//
//   W_final::backtrack        src/W_final.cc:175-719   (LOOP / FREE / M_WM / M_WMv / M_WMp, dispatch)
//   pseudo_loop::backtrack    src/pseudo_loop.cc:861-2820
//   insert_node               src/W_final.cc:721-731, src/pseudo_loop.cc:2823-2846
//
// Candidates of one node are spread over the lanes of `Par`; each candidate carries its position in
// the reference's evaluation order, and the winner is the lexicographic minimum (value, position),
// which is exactly "first strictly smaller wins".
#pragma once
#include <math.h>
#include "ccj_cells4.cuh"

// node types (src/constants.hh:21-73)
#define TB_LOOP 'V'
#define TB_FREE 'W'
#define TB_M_WM 'B'
#define TB_M_WMv 'v'
#define TB_M_WMp 'p'
#define TB_P_P 'P'
#define TB_P_PK 'k'
#define TB_P_PL 'l'
#define TB_P_PR 'r'
#define TB_P_PM 'm'
#define TB_P_PO 'o'
#define TB_P_PfromL 'f'
#define TB_P_PfromR 'g'
#define TB_P_PfromM 'h'
#define TB_P_PfromMprime '['
#define TB_P_PfromMdoubleprime ']'
#define TB_P_PfromO 'i'
#define TB_P_PLiloop 'j'
#define TB_P_PLiloop5 'b'
#define TB_P_PLmloop 'c'
#define TB_P_PLmloop10 'e'
#define TB_P_PLmloop01 'n'
#define TB_P_PLmloop00 'a'
#define TB_P_PRiloop 'q'
#define TB_P_PRiloop5 's'
#define TB_P_PRmloop 't'
#define TB_P_PRmloop10 'u'
#define TB_P_PRmloop01 '&'
#define TB_P_PRmloop00 '9'
#define TB_P_PMiloop 'w'
#define TB_P_PMiloop5 'x'
#define TB_P_PMmloop 'y'
#define TB_P_PMmloop10 '0'
#define TB_P_PMmloop01 '1'
#define TB_P_PMmloop00 '8'
#define TB_P_POiloop 'z'
#define TB_P_POiloop5 '5'
#define TB_P_POmloop '+'
#define TB_P_POmloop10 '-'
#define TB_P_POmloop01 '='
#define TB_P_POmloop00 '_'
#define TB_P_WB '*'
#define TB_P_WBP '^'
#define TB_P_WP '#'
#define TB_P_WPP '@'

// message prefixes of the reference's std::cerr lines; message id = prefix*256 + node type char
enum { TBM_NONE = 0, TBM_BORDER_CASE, TBM_BORDER_CASES, TBM_BODER_CASES, TBM_IMPOSSIBLE_CASE, TBM_IMPOSSIBLE_CASES,
       TBM_IMPOSSBIBLE_CASES };

// CCJ_HOST_EMU (the g++ build for tests/emu/simt_emu.hpp): the three members that store to global memory stay out of
// line, so that ThreadSanitizer's suppression list can name exactly them (tests/emu/tsan_traceback.supp)
#ifdef CCJ_HOST_EMU
#define CCJ_TB_STORE __attribute__((noinline))
#else
#define CCJ_TB_STORE
#endif
struct ccj_tb {
    const ccj_cx &c;
    int top;   // number of nodes on the stack
    bool stop; // an exit() was reached
    CCJ_HD ccj_tb(const ccj_cx &cx) : c(cx), top(0), stop(false) {}

    CCJ_TB_STORE CCJ_HD void push(int i, int j, int k, int l, int type) {
        if (top >= c.q.tb_cap) {
            fail(CCJ_STACK_OVERFLOW, 0);
            return;
        }
        int32_t *s = c.q.tb_stack + 5 * top++;
        s[0] = i; s[1] = j; s[2] = k; s[3] = l; s[4] = type;
    }
    CCJ_HD void push2(int i, int j, int type) { push(i, j, 0, 0, type); }
    CCJ_TB_STORE CCJ_HD void fail(int status, int msg) {
        if (!stop) {
            c.q.status[0] = status;
            c.q.status[2] = msg;
        }
        stop = true;
    }
    CCJ_HD void die(int prefix, int type) { fail(CCJ_EXIT_FAILURE, prefix * 256 + type); }
    CCJ_TB_STORE CCJ_HD void setpair(int x, int y, int type) {
        c.q.pair_out[x] = y; c.q.pair_out[y] = x;
        c.q.ftype_out[x] = (int8_t)type; c.q.ftype_out[y] = (int8_t)type;
    }
};

// common 4-index guards; the order (border first, then impossible) is the reference's in all cases but
// P_POiloop
CCJ_HD bool ccj_tb_border(int i, int j, int k, int l) { return !(i <= j && j < k - 1 && k <= l); }
CCJ_HD bool ccj_tb_imposs(int n, int i, int j, int k, int l) {
    return i <= 0 || j <= 0 || k <= 0 || l <= 0 || i > n || j > n || k > n || l > n;
}

template <class Par>
CCJ_HD void ccj_tb_node(ccj_tb &T, const Par &par, int ni, int nj, int nk, int nl, int type) {
    const ccj_cx &c = T.c;
    const ccj_model *M = c.M;
    const int8_t *S = c.q.S;
    const int n = c.q.n;
    const int INF = CCJ_INF;
    const int L = par.lane(), NL = Par::nlanes;
    const int PB = M->PB_penalty, bp = M->bp_penalty, cp = M->cp_penalty, ap = M->ap_penalty;
    // 4-index nodes are stored (i, j=l, k=j, l=k) (src/pseudo_loop.cc:2836-2846)
    const int i = ni, l = nj, j = nk, k = nl;

    switch (type) {
        // ------------------------------------------------------------------ W_final::backtrack
        case TB_LOOP: {  // src/W_final.cc:179-342
            const int i = ni, j = nj;
            if (i >= j) return;
            T.setpair(i, j, 'N');  // type overwritten below
            const int vt = ccj_raw2(c, T2_VTYPE, i, j);
            if (vt == 'H') {
                T.setpair(i, j, 'H');
            } else if (vt == 'I') {
                T.setpair(i, j, 'I');
                ccj_best b = {INF, -1};
                const int max_ip = ccj_min(j - CCJ_TURN - 2, i + CCJ_MAXLOOP + 1);
                for (int k = i + 1 + L; k <= max_ip; k += NL) {
                    const int min_l = ccj_max(k + CCJ_TURN + 1 + CCJ_MAXLOOP + 2, k + j - i) - CCJ_MAXLOOP - 2;
                    for (int l = j - 1; l >= min_l; --l)
                        ccj_cand(b, ccj_Vint_term(c, i, j, k, l), (k - i) * 64 + (j - l));
                }
                b = par.argmin(b);
                if (b.ord >= 0) {
                    const int best_ip = i + b.ord / 64, best_jp = j - b.ord % 64;
                    T.push2(best_ip, best_jp, TB_LOOP);  // best_ip < best_jp always holds for a real candidate
                } else {
                    // best_ip=j, best_jp=i: "NOT GOOD RESTR INTER" on stderr, exit(0)
                    T.fail(CCJ_EXIT_ZERO_NOT_GOOD, 0);
                    c.q.status[3] = i;
                    c.q.status[4] = j;
                }
            } else if (vt == 'M') {
                T.setpair(i, j, 'M');
                const int tt = ccj_pt(c, j, i);
                const int e00 = ccj_E_MLstem(M, tt, -1, -1) + M->MLclosing;
                const int e01 = ccj_E_MLstem(M, tt, -1, S[i + 1]) + M->MLclosing + M->MLbase;
                const int e10 = ccj_E_MLstem(M, tt, S[j - 1], -1) + M->MLclosing + M->MLbase;
                const int e11 = ccj_E_MLstem(M, tt, S[j - 1], S[i + 1]) + M->MLclosing + 2 * M->MLbase;
                ccj_best b = {INF, -1};
                for (int k = i + 1 + L; k <= j - 1; k += NL) {
                    const int o = (k - i) * 8;
                    const int wm1 = ccj_V(c, T2_WM, i + 1, k - 1), wm2 = ccj_V(c, T2_WM, i + 2, k - 1);
                    const int a1 = ccj_min(ccj_V(c, T2_WMv, k, j - 1), ccj_V(c, T2_WMp, k, j - 1));
                    const int a2 = ccj_min(ccj_V(c, T2_WMv, k, j - 2), ccj_V(c, T2_WMp, k, j - 2));
                    const int wp1 = ccj_V(c, T2_WMp, k, j - 1), wp2 = ccj_V(c, T2_WMp, k, j - 2);
                    int tmp;
                    tmp = wm1 + a1 + e00; ccj_cand(b, tmp, o + 0);
                    tmp = wm2 + a1 + e01; ccj_cand(b, tmp, o + 1);
                    tmp = wm1 + a2 + e10; ccj_cand(b, tmp, o + 2);
                    tmp = wm2 + a2 + e11; ccj_cand(b, tmp, o + 3);
                    tmp = (k - i - 1) * M->MLbase + wp1 + e00; ccj_cand(b, tmp, o + 4);
                    // rows 6 and 8 keep the previous tmp when k-(i+1)-1 < 0 (src/W_final.cc:286,299): a stale
                    // value can never be strictly smaller than the running minimum, so it is a duplicate
                    if ((k - (i + 1) - 1) >= 0) tmp = (k - (i + 1) - 1) * M->MLbase + wp1 + e01;
                    ccj_cand(b, tmp, o + 5);
                    tmp = (k - i - 1) * M->MLbase + wp2 + e10; ccj_cand(b, tmp, o + 6);
                    if ((k - (i + 1) - 1) >= 0) tmp = (k - (i + 1) - 1) * M->MLbase + wp2 + e11;
                    ccj_cand(b, tmp, o + 7);
                }
                b = par.argmin(b);
                if (b.ord >= 0) {
                    const int best_k = i + b.ord / 8;
                    switch (b.ord % 8 + 1) {
                        case 1: T.push2(i + 1, best_k - 1, TB_M_WM); T.push2(best_k, j - 1, TB_M_WM); break;
                        case 2: T.push2(i + 2, best_k - 1, TB_M_WM); T.push2(best_k, j - 1, TB_M_WM); break;
                        case 3: T.push2(i + 1, best_k - 1, TB_M_WM); T.push2(best_k, j - 2, TB_M_WM); break;
                        case 4: T.push2(i + 2, best_k - 1, TB_M_WM); T.push2(best_k, j - 2, TB_M_WM); break;
                        case 5: case 6: T.push2(best_k, j - 1, TB_M_WM); break;
                        case 7: case 8: T.push2(best_k, j - 2, TB_M_WM); break;
                    }
                }
            }
        } break;

        case TB_FREE: {  // src/W_final.cc:344-539
            const int j = nj;
            if (j == 1) return;
            const int32_t *W = c.q.W;
            ccj_best b = {INF, -1};
            if (L == 0) ccj_cand(b, W[j - 1], 0);  // row 0
            const int dg = M->dangles;
            // first loop (rows 1-4) entirely precedes the second (rows 5-8)
            for (int i = 1 + L; i <= j - 1; i += NL) {
                const int acc = (i > 1) ? W[i - 1] : 0;
                int e = ccj_V(c, T2_V, i, j);
                if (e < INF) {
                    int tmp;
                    if (dg == 2) {
                        const int si1 = i > 1 ? S[i - 1] : -1, sj1 = j < n ? S[j + 1] : -1;
                        tmp = e + ccj_E_ext_stem(M, ccj_pt(c, i, j), si1, sj1) + acc;
                    } else {
                        tmp = e + ccj_E_ext_stem(M, ccj_pt(c, i, j), -1, -1) + acc;
                    }
                    ccj_cand(b, tmp, 8 + i * 8 + 1);
                }
                if (dg == 1) {
                    e = ccj_V(c, T2_V, i + 1, j);
                    if (e < INF) ccj_cand(b, e + ccj_E_ext_stem(M, ccj_pt(c, i + 1, j), S[i], -1) + acc, 8 + i * 8 + 2);
                    e = ccj_V(c, T2_V, i, j - 1);
                    if (e < INF) ccj_cand(b, e + ccj_E_ext_stem(M, ccj_pt(c, i, j - 1), -1, S[j]) + acc, 8 + i * 8 + 3);
                    e = ccj_V(c, T2_V, i + 1, j - 1);
                    if (e < INF)
                        ccj_cand(b, e + ccj_E_ext_stem(M, ccj_pt(c, i + 1, j - 1), S[i], S[j]) + acc, 8 + i * 8 + 4);
                }
            }
            const int base2 = 8 + (n + 2) * 8;
            for (int i = 1 + L; i <= j - 1; i += NL) {
                const int acc = (i - 1 > 0) ? W[i - 1] : 0;
                int e = ccj_tri_get(c, T2_P, i, j);
                if (e < INF) ccj_cand(b, e + M->PS_penalty + acc, base2 + i * 8 + 5);
                if (dg == 1) {
                    e = ccj_tri_get(c, T2_P, i + 1, j);
                    if (e < INF) ccj_cand(b, e + M->PS_penalty + acc, base2 + i * 8 + 6);
                    e = ccj_tri_get(c, T2_P, i, j - 1);
                    if (e < INF) ccj_cand(b, e + M->PS_penalty + acc, base2 + i * 8 + 7);
                    e = ccj_tri_get(c, T2_P, i + 1, j - 1);
                    if (e < INF) ccj_cand(b, e + M->PS_penalty + acc, base2 + i * 8 + 8);
                }
            }
            b = par.argmin(b);
            if (b.ord < 0) return;
            if (b.ord == 0) { T.push2(1, j - 1, TB_FREE); return; }
            int o = b.ord - 8;
            if (o >= (n + 2) * 8) o -= (n + 2) * 8;
            const int row = o % 8 == 0 ? 8 : o % 8;
            const int best_i = (o - row) / 8;
            switch (row) {
                case 1: T.push2(best_i, j, TB_LOOP); if (best_i - 1 > 1) T.push2(1, best_i - 1, TB_FREE); break;
                case 2: T.push2(best_i + 1, j, TB_LOOP); if (best_i >= 1) T.push2(1, best_i, TB_FREE); break;
                case 3: T.push2(best_i, j - 1, TB_LOOP); if (best_i - 1 > 1) T.push2(1, best_i - 1, TB_FREE); break;
                case 4: T.push2(best_i + 1, j - 1, TB_LOOP); if (best_i >= 1) T.push2(1, best_i, TB_FREE); break;
                case 5: T.push2(best_i, j, TB_P_P); if (best_i - 1 > 1) T.push2(1, best_i - 1, TB_FREE); break;
                case 6: T.push2(best_i + 1, j, TB_P_P); if (best_i >= 1) T.push2(1, best_i, TB_FREE); break;
                case 7: T.push2(best_i, j - 1, TB_P_P); if (best_i - 1 > 1) T.push2(1, best_i - 1, TB_FREE); break;
                case 8: T.push2(best_i + 1, j - 1, TB_P_P); if (best_i >= 1) T.push2(1, best_i, TB_FREE); break;
            }
        } break;

        case TB_M_WM: {  // src/W_final.cc:541-595
            const int i = ni, j = nj;
            ccj_best b = {ccj_V(c, T2_WM, i, j - 1) + M->MLbase, 0};  // row 5, unconditional start
            for (int k = i + L; k <= j - CCJ_TURN - 1; k += NL) {
                const int o = (k - i + 1) * 4;
                const int wmv = ccj_V(c, T2_WMv, k, j), wmp = ccj_V(c, T2_WMp, k, j), wm = ccj_V(c, T2_WM, i, k - 1);
                ccj_cand(b, (k - i) * M->MLbase + wmv, o + 0);
                ccj_cand(b, (k - i) * M->MLbase + wmp, o + 1);
                ccj_cand(b, wm + wmv, o + 2);
                ccj_cand(b, wm + wmp, o + 3);
            }
            b = par.argmin(b);
            if (b.ord == 0) { T.push2(i, j - 1, TB_M_WM); return; }
            const int best_k = i + b.ord / 4 - 1;
            switch (b.ord % 4 + 1) {
                case 1: T.push2(best_k, j, TB_M_WMv); break;
                case 2: T.push2(best_k, j, TB_M_WMp); break;
                case 3: T.push2(i, best_k - 1, TB_M_WM); T.push2(best_k, j, TB_M_WMv); break;
                case 4: T.push2(i, best_k - 1, TB_M_WM); T.push2(best_k + 1, j, TB_M_WMp); break; /* sic */
            }
        } break;

        case TB_M_WMv: {  // src/W_final.cc:597-644
            const int i = ni, j = nj;
            const int si = S[i], sj = S[j];
            const int si1 = (i > 1) ? S[i - 1] : -1, sj1 = (j < n) ? S[j + 1] : -1;
            int tt = ccj_pt(c, i, j);
            int mn = ccj_V(c, T2_V, i, j) + ((M->dangles == 2) ? ccj_E_MLstem(M, tt, si1, sj1) : ccj_E_MLstem(M, tt, -1, -1));
            int row = 1, tmp;
            if (M->dangles == 1) {
                tt = ccj_pt(c, i + 1, j);
                tmp = ccj_V(c, T2_V, i + 1, j) + ccj_E_MLstem(M, tt, si, -1) + M->MLbase;
                if (tmp < mn) { mn = tmp; row = 2; }
                tt = ccj_pt(c, i, j - 1);
                tmp = ccj_V(c, T2_V, i, j - 1) + ccj_E_MLstem(M, tt, -1, sj) + M->MLbase;
                if (tmp < mn) { mn = tmp; row = 3; }
                tt = ccj_pt(c, i + 1, j - 1);
                tmp = ccj_V(c, T2_V, i + 1, j - 1) + ccj_E_MLstem(M, tt, si, sj) + 2 * M->MLbase;
                if (tmp < mn) { mn = tmp; row = 4; }
            }
            tmp = ccj_V(c, T2_WMv, i, j - 1) + M->MLbase;
            if (tmp < mn) { mn = tmp; row = 5; }
            switch (row) {
                case 1: T.push2(i, j, TB_LOOP); break;
                case 2: T.push2(i + 1, j, TB_LOOP); break;
                case 3: T.push2(i, j - 1, TB_LOOP); break;
                case 4: T.push2(i + 1, j - 1, TB_LOOP); break;
                case 5: T.push2(i, j - 1, TB_M_WMv); break;
            }
        } break;

        case TB_M_WMp: {  // src/W_final.cc:646-665: row 1 (the pseudoknot itself) pushes nothing
            const int i = ni, j = nj;
            const int mn = ccj_tri_get(c, T2_P, i, j) + M->PSM_penalty + M->b_penalty;
            const int tmp = ccj_V(c, T2_WMp, i, j - 1) + M->MLbase;
            if (tmp < mn) T.push2(i, j - 1, TB_M_WMp);
        } break;

        // ------------------------------------------------------------------ pseudo_loop::backtrack
        case TB_P_P: {  // :867-897
            const int i = ni, l = nj;
            if (i >= l) { T.die(TBM_BORDER_CASE, type); return; }
            // candidates (j,d,k) in lexicographic order; position packed as ((j-i)*s + (d-i))*s + (k-i)
            const int s = l - i + 1;
            ccj_best b = {INF, -1};
            if (c.q.status[6] == 1) {
                // the tuned fill left the two PK copies of compute_P (ccj_types.h): per (j, delta) both factors are rows
                // over d in PKF block (i,j) and PKG block (j+1,l).  Same candidates and the same (value, position)
                // order as the loops below; lanes walk d.
                const int *lay = c.q.lay;
                const int *CF = lay + CCJ_LAY_CF(n), *S2 = lay + CCJ_LAY_S2(n), *EG = lay + CCJ_LAY_EG(n);
                const int16_t *F0 = c.q.pkf + (lay[CCJ_LAY_DF(n) + i] - CF[i]);
                for (int j = i; j <= l - 3; ++j) {
                    const int Lr = l - j - 2, Mf = n - j - 1;
                    const int16_t *Fb = F0 + CF[j], *Gb = c.q.pkg + EG[j + 1] + S2[Lr];
                    for (int dl = 1; dl <= Lr; ++dl) {
                        const int16_t *F = Fb + 8 * ((int)ccj_q8(Mf) - (int)ccj_q8(Mf + 1 - dl));
                        const int16_t *G = Gb + 8 * ((int)ccj_q8(Lr) - (int)ccj_q8(Lr + 1 - dl));
                        for (int x = L; x <= Lr - dl; x += NL) {   // d = j+1+x, k = d+dl
                            const int d = j + 1 + x;
                            ccj_cand(b, (int)F[x] + (int)G[x], ((j - i) * s + (d - i)) * s + (d + dl - i));
                        }
                    }
                }
            } else {
                for (int j = i; j < l; ++j)
                    for (int d = j + 1; d < l; ++d)
                        for (int k = d + 1 + L; k < l; k += NL)
                            ccj_cand(b, ccj_P_term(c, i, l, j, d, k), ((j - i) * s + (d - i)) * s + (k - i));
            }
            b = par.argmin(b);
            int best_j = 0, best_d = 0, best_k = 0;
            if (b.ord >= 0) {
                best_k = i + b.ord % s;
                best_d = i + (b.ord / s) % s;
                best_j = i + b.ord / s / s;
            }
            T.push(i, best_k, best_j, best_d + 1, TB_P_PK);
            T.push(best_j + 1, l, best_d, best_k + 1, TB_P_PK);
        } break;

        case TB_P_PK: {  // :899-997
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            ccj_best b = {INF, -1};
            for (int d = i + 1 + L; d < j; d += NL) ccj_cand(b, ccj_get4(c, T_PK, i, d, k, l) + ccj_WP(c, d + 1, j), d);
            for (int d = k + 1 + L; d < l; d += NL) ccj_cand(b, ccj_get4(c, T_PK, i, j, d, l) + ccj_WP(c, k, d - 1), n + 2 + d);
            b = par.argmin(b);
            const int o = 2 * (n + 2);
            ccj_cand(b, ccj_get4(c, T_PL, i, j, k, l) + PB, o + 3);
            ccj_cand(b, ccj_get4(c, T_PM, i, j, k, l) + PB, o + 4);
            ccj_cand(b, ccj_get4(c, T_PR, i, j, k, l) + PB, o + 5);
            ccj_cand(b, ccj_get4(c, T_PO, i, j, k, l) + PB, o + 6);
            if (b.ord < 0) return;
            if (b.ord < n + 2) { const int d = b.ord; T.push(i, l, d, k, TB_P_PK); T.push2(d + 1, j, TB_P_WP); }
            else if (b.ord < o) { const int d = b.ord - (n + 2); T.push(i, l, j, d, TB_P_PK); T.push2(k, d - 1, TB_P_WP); }
            else if (b.ord == o + 3) T.push(i, l, j, k, TB_P_PL);
            else if (b.ord == o + 4) T.push(i, l, j, k, TB_P_PM);
            else if (b.ord == o + 5) T.push(i, l, j, k, TB_P_PR);
            else T.push(i, l, j, k, TB_P_PO);
        } break;

        case TB_P_PL: {  // :1000-1065
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            int mn = INF, row = -1, tmp;
            if (ccj_pt(c, i, j) > 0) {
                tmp = ccj_PLiloop(c, i, j, k, l);
                if (tmp < mn) { mn = tmp; row = 1; }
                tmp = ccj_PXmloop(c, T_PLmloop10, T_PLmloop01, i + 1, j - 1, k, l) + bp;
                if (tmp < mn) { mn = tmp; row = 2; }
                if (j >= i + CCJ_TURN + 1) {
                    tmp = ccj_get4(c, T_PfromL, i + 1, j - 1, k, l);
                    if (tmp < mn) { mn = tmp; row = 3; }
                }
            }
            if (row == 1) T.push(i, l, j, k, TB_P_PLiloop);
            else if (row == 2) T.push(i, l, j, k, TB_P_PLmloop);
            else if (row == 3) { T.push(i + 1, l, j - 1, k, TB_P_PfromL); T.setpair(i, j, TB_P_PL); }
        } break;

        case TB_P_PR: {  // :1067-1130.  The range test is `>= n` here: l==n exits
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BODER_CASES, type); return; }
            if (i < 0 || j < 0 || k < 0 || l < 0 || i >= n || j >= n || k >= n || l >= n) {
                T.die(TBM_IMPOSSIBLE_CASES, type);
                return;
            }
            int mn = INF, row = -1, tmp;
            if (ccj_pt(c, k, l) > 0) {
                tmp = ccj_PRiloop(c, i, j, k, l);
                if (tmp < mn) { mn = tmp; row = 1; }
                tmp = ccj_PXmloop(c, T_PRmloop10, T_PRmloop01, i, j, k + 1, l - 1) + bp;
                if (tmp < mn) { mn = tmp; row = 2; }
                if (l >= k + CCJ_TURN + 1) {
                    tmp = ccj_get4(c, T_PfromR, i, j, k + 1, l - 1);
                    if (tmp < mn) { mn = tmp; row = 3; }
                }
            }
            if (row == 1) T.push(i, l, j, k, TB_P_PRiloop);
            else if (row == 2) T.push(i, l, j, k, TB_P_PRmloop);
            else if (row == 3) { T.push(i, l - 1, j, k + 1, TB_P_PfromR); T.setpair(k, l, TB_P_PR); }
        } break;

        case TB_P_PM: {  // :1132-1200
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            if (i == j && k == l) { T.setpair(j, k, TB_P_PM); return; }
            int mn = INF, row = -1, tmp;
            if (ccj_pt(c, j, k) > 0) {
                tmp = ccj_PMiloop(c, i, j, k, l);
                if (tmp < mn) { mn = tmp; row = 1; }
                tmp = ccj_PXmloop(c, T_PMmloop10, T_PMmloop01, i, j - 1, k + 1, l) + bp;
                if (tmp < mn) { mn = tmp; row = 2; }
                if (k >= j + CCJ_TURN - 1) {
                    tmp = ccj_get4(c, T_PfromM, i, j - 1, k + 1, l);
                    if (tmp < mn) { mn = tmp; row = 3; }
                }
            }
            if (row == 1) T.push(i, l, j, k, TB_P_PMiloop);
            else if (row == 2) T.push(i, l, j, k, TB_P_PMmloop);
            else if (row == 3) { T.push(i, l, j - 1, k + 1, TB_P_PfromM); T.setpair(j, k, TB_P_PM); }
        } break;

        case TB_P_PO: {  // :1202-1261
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            int mn = INF, row = -1, tmp;
            if (ccj_pt(c, i, l) > 0) {
                tmp = ccj_POiloop(c, i, j, k, l);
                if (tmp < mn) { mn = tmp; row = 1; }
                tmp = ccj_PXmloop(c, T_POmloop10, T_POmloop01, i + 1, j, k, l - 1) + bp;
                if (tmp < mn) { mn = tmp; row = 2; }
                if (l >= i + CCJ_TURN + 1) {
                    tmp = ccj_get4(c, T_PfromO, i + 1, j, k, l - 1);
                    if (tmp < mn) { mn = tmp; row = 3; }
                }
            }
            if (row == 1) T.push(i, l, j, k, TB_P_POiloop);
            else if (row == 2) T.push(i, l, j, k, TB_P_POmloop);
            else if (row == 3) { T.push(i + 1, l - 1, j, k, TB_P_PfromO); T.setpair(i, l, TB_P_PO); }
        } break;

        case TB_P_PfromL: {  // :1263-1354
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_NONE, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_NONE, type); return; }
            if (i == j && k == l) return;
            ccj_best b = {INF, -1};
            for (int d = i + 1 + L; d < j; d += NL) {
                ccj_cand(b, ccj_get4(c, T_PfromL, d, j, k, l) + ccj_WP(c, i, d - 1), 2 * d);
                ccj_cand(b, ccj_get4(c, T_PfromL, i, d, k, l) + ccj_WP(c, d + 1, j), 2 * d + 1);
            }
            b = par.argmin(b);
            const int o = 2 * (n + 2);
            ccj_cand(b, ccj_get4(c, T_PR, i, j, k, l) + PB, o + 3);
            ccj_cand(b, ccj_get4(c, T_PM, i, j, k, l) + PB, o + 4);
            ccj_cand(b, ccj_get4(c, T_PO, i, j, k, l) + PB, o + 5);
            if (b.ord < 0) return;
            if (b.ord < o) {
                const int d = b.ord / 2;
                if (b.ord % 2 == 0) { T.push(d, l, j, k, TB_P_PfromL); T.push2(i, d - 1, TB_P_WP); }
                else { T.push(i, l, d, k, TB_P_PfromL); T.push2(d + 1, j, TB_P_WP); }
            } else if (b.ord == o + 3) T.push(i, l, j, k, TB_P_PR);
            else if (b.ord == o + 4) T.push(i, l, j, k, TB_P_PM);
            else T.push(i, l, j, k, TB_P_PO);
        } break;

        case TB_P_PfromR: {  // :1356-1437
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_NONE, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASE, type); return; }
            if (i == j && k == l) return;
            ccj_best b = {INF, -1};
            for (int d = k + 1 + L; d < l; d += NL) {
                ccj_cand(b, ccj_get4(c, T_PfromR, i, j, d, l) + ccj_WP(c, k, d - 1), 2 * d);
                ccj_cand(b, ccj_get4(c, T_PfromR, i, j, k, d) + ccj_WP(c, d + 1, l), 2 * d + 1);
            }
            b = par.argmin(b);
            const int o = 2 * (n + 2);
            ccj_cand(b, ccj_get4(c, T_PM, i, j, k, l) + PB, o + 3);
            ccj_cand(b, ccj_get4(c, T_PO, i, j, k, l) + PB, o + 4);
            if (b.ord < 0) return;
            if (b.ord < o) {
                const int d = b.ord / 2;
                if (b.ord % 2 == 0) { T.push(i, l, j, d, TB_P_PfromR); T.push2(k, d - 1, TB_P_WP); }
                else { T.push(i, d, j, k, TB_P_PfromR); T.push2(d + 1, l, TB_P_WP); }
            } else if (b.ord == o + 3) T.push(i, l, j, k, TB_P_PM);
            else T.push(i, l, j, k, TB_P_PO);
        } break;

        case TB_P_PfromM: {  // :1439-1480
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_NONE, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_NONE, type); return; }
            if (i == j && k == l) return;
            ccj_best b = {INF, -1};
            for (int d = i + 1 + L; d < j; d += NL)
                ccj_cand(b, ccj_get4(c, T_PfromMprime, i, d, k, l) + ccj_WP(c, d + 1, j), d);
            b = par.argmin(b);
            if (b.ord > -1) { T.push(i, l, b.ord, k, TB_P_PfromMprime); T.push2(b.ord + 1, j, TB_P_WP); }
        } break;

        case TB_P_PfromMprime: {  // :1482-1522 (never reached through W_final's dispatch, kept for completeness)
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_NONE, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_NONE, type); return; }
            if (i == j && k == l) return;
            ccj_best b = {INF, -1};
            for (int d = k + 1 + L; d < l; d += NL) {
                int mpp;  // get_PfromMdoubleprime(i,j,d,l), :663-679
                if (!ccj_valid4(i, j, d, l)) mpp = INF;
                else if (i == j && d == l) mpp = ccj_pt(c, i, l) == 0 ? INF : 0;
                else mpp = ccj_min(ccj_get4(c, T_PL, i, j, d, l) + PB, ccj_get4(c, T_PR, i, j, d, l) + PB);
                ccj_cand(b, mpp + ccj_WP(c, k, d - 1), d);
            }
            b = par.argmin(b);
            if (b.ord > -1) { T.push(i, l, j, b.ord, TB_P_PfromMdoubleprime); T.push2(k, b.ord - 1, TB_P_WP); }
        } break;

        case TB_P_PfromMdoubleprime: {  // :1524-1574
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_NONE, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_NONE, type); return; }
            if (i == j && k == l) return;
            int mn = INF, row = -1, tmp;
            tmp = ccj_get4(c, T_PL, i, j, k, l) + PB;
            if (tmp < mn) { mn = tmp; row = 1; }
            tmp = ccj_get4(c, T_PR, i, j, k, l) + PB;
            if (tmp < mn) { mn = tmp; row = 2; }
            if (row == 1) T.push(i, l, j, k, TB_P_PL);
            else if (row == 2) T.push(i, l, j, k, TB_P_PR);
        } break;

        case TB_P_PfromO: {  // :1576-1659
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASE, type); return; }
            if (i == j && k == l) return;
            ccj_best b = {INF, -1};
            for (int d = i + 1 + L; d < j; d += NL) ccj_cand(b, ccj_get4(c, T_PfromO, d, j, k, l) + ccj_WP(c, i, d - 1), d);
            for (int d = k + 1 + L; d < l; d += NL)
                ccj_cand(b, ccj_get4(c, T_PfromO, i, j, k, d) + ccj_WP(c, d + 1, l), n + 2 + d);
            b = par.argmin(b);
            const int o = 2 * (n + 2);
            ccj_cand(b, ccj_get4(c, T_PL, i, j, k, l) + PB, o + 3);
            ccj_cand(b, ccj_get4(c, T_PR, i, j, k, l) + PB, o + 4);
            if (b.ord < 0) return;
            if (b.ord < n + 2) { const int d = b.ord; T.push(d, l, j, k, TB_P_PfromO); T.push2(i, d - 1, TB_P_WP); }
            else if (b.ord < o) { const int d = b.ord - (n + 2); T.push(i, d, j, k, TB_P_PfromO); T.push2(d + 1, l, TB_P_WP); }
            else if (b.ord == o + 3) T.push(i, l, j, k, TB_P_PL);
            else T.push(i, l, j, k, TB_P_PR);
        } break;

        case TB_P_WB: {  // :1660-1700
            const int i = ni, l = nj;
            if (i <= 0 || l <= 0 || i > n || l > n) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            if (i > l) return;
            int mn = INF, row = -1, tmp;
            tmp = ccj_tri_get(c, T2_WBP, i, l);
            if (tmp < mn) { mn = tmp; row = 1; }
            tmp = cp * (l - i + 1);
            if (tmp < mn) { mn = tmp; row = 2; }
            if (row == 1) T.push2(i, l, TB_P_WBP);
        } break;

        case TB_P_WBP:    // :1701-1756
        case TB_P_WPP: {  // :1799-1853
            const int i = ni, l = nj;
            const bool isB = type == TB_P_WBP;
            if (i > l) { T.die(TBM_BORDER_CASE, type); return; }
            if (i <= 0 || l <= 0 || i > n || l > n) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            ccj_best b = {INF, -1};
            for (int d = i + L; d < l; d += NL) {
                const int w = isB ? ccj_WB(c, i, d - 1) : ccj_WP(c, i, d - 1);
                ccj_cand(b, w + ccj_V(c, T2_V, d, l) + (isB ? bp : 0) + M->PPS_penalty, 2 * d);
                ccj_cand(b, w + ccj_tri_get(c, T2_P, d, l) + (isB ? M->PSM_penalty : M->PSP_penalty) + M->PPS_penalty,
                         2 * d + 1);
            }
            b = par.argmin(b);
            const int o = 2 * (n + 2);
            ccj_cand(b, ccj_tri_get(c, isB ? T2_WBP : T2_WPP, i, l - 1) + (isB ? cp : M->PUP_penalty), o);
            if (b.ord < 0) return;
            if (b.ord == o) { T.push2(i, l - 1, type); return; }
            const int d = b.ord / 2;
            T.push2(i, d - 1, isB ? TB_P_WB : TB_P_WP);
            T.push2(d, l, b.ord % 2 == 0 ? TB_LOOP : TB_P_P);
        } break;

        case TB_P_WP: {  // :1758-1798
            const int i = ni, l = nj;
            if (i <= 0 || l <= 0 || i > n || l > n) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            if (i > l) return;
            int mn = INF, row = -1, tmp;
            tmp = ccj_tri_get(c, T2_WPP, i, l);
            if (tmp < mn) { mn = tmp; row = 1; }
            tmp = M->PUP_penalty * (l - i + 1);
            if (tmp < mn) { mn = tmp; row = 2; }
            if (row == 1) T.push2(i, l, TB_P_WPP);
        } break;

        case TB_P_PLiloop: {  // :1855-1913: strict border test, no can_pair in the window, unconditional stack row
            if (!(i < j && j < k - 1 && k < l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSBIBLE_CASES, type); return; }
            T.setpair(i, j, TB_P_PLiloop);
            ccj_best b = {INF, -1};
            if (ccj_pt(c, i, j) > 0) {
                if (L == 0) ccj_cand(b, ccj_get4(c, T_PL, i + 1, j - 1, k, l) + ccj_e_stP(M, S, i, j), 0);
                const int max_d = ccj_min(j, i + CCJ_MAXLOOP);
                for (int d = i + 1 + L; d < max_d; d += NL) {
                    const int min_dp = ccj_max(d + CCJ_TURN, j - CCJ_MAXLOOP);
                    for (int dp = j - 1; dp > min_dp; --dp)
                        ccj_cand(b, ccj_e_intP(M, S, i, d, dp, j) + ccj_get4(c, T_PL, d, dp, k, l), (d - i) * 64 + (j - dp));
                }
            }
            b = par.argmin(b);
            if (b.ord == 0) T.push(i + 1, l, j - 1, k, TB_P_PL);
            else if (b.ord > 0) T.push(i + b.ord / 64, l, j - b.ord % 64, k, TB_P_PL);
        } break;

        case TB_P_PLmloop: {  // :1915-1954
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            T.setpair(i, j, TB_P_PLmloop);
            const int b1 = ccj_get4(c, T_PLmloop10, i + 1, j - 1, k, l) + ap + bp;
            const int b2 = ccj_get4(c, T_PLmloop01, i + 1, j - 1, k, l) + ap + bp;
            T.push(i + 1, l, j - 1, k, b1 < b2 ? TB_P_PLmloop10 : TB_P_PLmloop01);
        } break;

        case TB_P_PLmloop00: {  // :1955-2010 (starts from the now-set PL(i,j,k,l)+bp)
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            ccj_best b = {ccj_get4(c, T_PL, i, j, k, l) + bp, 0};
            for (int d = i + L; d <= j; d += NL) {
                if (d > i) ccj_cand(b, ccj_WB(c, i, d - 1) + ccj_get4(c, T_PLmloop00, d, j, k, l), 2 * d);
                if (d < j) ccj_cand(b, ccj_get4(c, T_PLmloop00, i, d, k, l) + ccj_WB(c, d + 1, j), 2 * d + 1);
            }
            b = par.argmin(b);
            if (b.ord == 0) T.push(i, l, j, k, TB_P_PL);
            else if (b.ord % 2 == 0) { const int d = b.ord / 2; T.push(d, l, j, k, TB_P_PLmloop00); T.push2(i, d - 1, TB_P_WB); }
            else { const int d = b.ord / 2; T.push(i, l, d, k, TB_P_PLmloop00); T.push2(d + 1, j, TB_P_WB); }
        } break;

        case TB_P_PLmloop01: {  // :2011-2042 (pushes even when nothing was found: best_d = -1)
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            ccj_best b = {INF, -1};
            for (int d = i + L; d < j; d += NL)
                ccj_cand(b, ccj_get4(c, T_PLmloop00, i, d, k, l) + ccj_tri_get(c, T2_WBP, d + 1, j), d);
            b = par.argmin(b);
            T.push(i, l, b.ord, k, TB_P_PLmloop00);
            T.push2(b.ord + 1, j, TB_P_WBP);
        } break;

        case TB_P_PLmloop10: {  // :2043-2090
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            ccj_best b = {INF, -1};
            for (int d = i + 1 + L; d <= j; d += NL) {
                ccj_cand(b, ccj_tri_get(c, T2_WBP, i, d - 1) + ccj_get4(c, T_PLmloop00, d, j, k, l), 2 * d);
                if (d < j) ccj_cand(b, ccj_get4(c, T_PLmloop10, i, d, k, l) + ccj_WB(c, d + 1, j), 2 * d + 1);
            }
            b = par.argmin(b);
            if (b.ord < 0) return;
            const int d = b.ord / 2;
            if (b.ord % 2 == 0) { T.push2(i, d - 1, TB_P_WBP); T.push(d, l, j, k, TB_P_PLmloop00); }
            else { T.push(i, l, d, k, TB_P_PLmloop10); T.push2(d + 1, j, TB_P_WB); }
        } break;

        case TB_P_PRiloop: {  // :2092-2151
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            T.setpair(k, l, TB_P_PRiloop);
            ccj_best b = {INF, -1};
            if (ccj_pt(c, k, l) > 0) {
                if (L == 0) ccj_cand(b, ccj_get4(c, T_PR, i, j, k + 1, l - 1) + ccj_e_stP(M, S, k, l), 0);
                const int max_d = ccj_min(l, k + CCJ_MAXLOOP);
                for (int d = k + 1 + L; d < max_d; d += NL) {
                    const int min_dp = ccj_max(d + CCJ_TURN, l - CCJ_MAXLOOP);
                    for (int dp = l - 1; dp > min_dp; --dp)
                        ccj_cand(b, ccj_e_intP(M, S, k, d, dp, l) + ccj_get4(c, T_PR, i, j, d, dp), (d - k) * 64 + (l - dp));
                }
            }
            b = par.argmin(b);
            if (b.ord == 0) T.push(i, l - 1, j, k + 1, TB_P_PR);
            else if (b.ord > 0) T.push(i, l - b.ord % 64, j, k + b.ord / 64, TB_P_PR);
        } break;

        case TB_P_PRmloop: {  // :2153-2191
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            T.setpair(k, l, TB_P_PRmloop);
            const int b1 = ccj_get4(c, T_PRmloop10, i, j, k + 1, l - 1) + ap + bp;
            const int b2 = ccj_get4(c, T_PRmloop01, i, j, k + 1, l - 1) + ap + bp;
            T.push(i, l - 1, j, k + 1, b1 < b2 ? TB_P_PRmloop10 : TB_P_PRmloop01);
        } break;

        case TB_P_PRmloop00: {  // :2193-2248 -- note insert_node(i,j,...,l,...) argument order (sic)
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            ccj_best b = {ccj_get4(c, T_PR, i, j, k, l) + bp, 0};
            for (int d = k + L; d <= l; d += NL) {
                if (d > k) ccj_cand(b, ccj_WB(c, k, d - 1) + ccj_get4(c, T_PRmloop00, i, j, d, l), 2 * d);
                if (d < l) ccj_cand(b, ccj_get4(c, T_PRmloop00, i, j, k, d) + ccj_WB(c, d + 1, l), 2 * d + 1);
            }
            b = par.argmin(b);
            if (b.ord == 0) T.push(i, j, k, l, TB_P_PR);
            else if (b.ord % 2 == 0) { const int d = b.ord / 2; T.push(i, j, d, l, TB_P_PRmloop00); T.push2(k, d - 1, TB_P_WB); }
            else { const int d = b.ord / 2; T.push(i, j, k, d, TB_P_PRmloop00); T.push2(d + 1, l, TB_P_WB); }
        } break;

        case TB_P_PRmloop01: {  // :2250-2292
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            ccj_best b = {ccj_get4(c, T_PRmloop01, i, j, k, l - 1) + cp, 0};
            for (int d = k + L; d < l; d += NL)
                ccj_cand(b, ccj_get4(c, T_PRmloop00, i, j, k, d) + ccj_tri_get(c, T2_WBP, d + 1, l), d + 1);
            b = par.argmin(b);
            if (b.ord == 0) T.push(i, l - 1, j, k, TB_P_PRmloop01);
            else { const int d = b.ord - 1; T.push2(d + 1, l, TB_P_WBP); T.push(i, d, j, k, TB_P_PRmloop00); }
        } break;

        case TB_P_PRmloop10: {  // :2294-2336
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            ccj_best b = {ccj_get4(c, T_PRmloop10, i, j, k + 1, l) + cp, 0};
            for (int d = k + 1 + L; d <= l; d += NL)
                ccj_cand(b, ccj_tri_get(c, T2_WBP, k, d - 1) + ccj_get4(c, T_PRmloop00, i, j, d, l), d + 1);
            b = par.argmin(b);
            if (b.ord == 0) T.push(i, l, j, k + 1, TB_P_PRmloop10);
            else { const int d = b.ord - 1; T.push2(k, d - 1, TB_P_WBP); T.push(i, l, j, d, TB_P_PRmloop00); }
        } break;

        case TB_P_PMiloop: {  // :2338-2397
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            T.setpair(j, k, TB_P_PMiloop);
            ccj_best b = {INF, -1};
            if (ccj_pt(c, j, k) > 0) {
                if (L == 0) ccj_cand(b, ccj_get4(c, T_PM, i, j - 1, k + 1, l) + ccj_e_stP(M, S, j - 1, k + 1), 0);
                const int max_d = ccj_max(i, j - CCJ_MAXLOOP);
                for (int d = j - 1 - L; d > max_d; d -= NL) {
                    const int min_dp = ccj_min(l, k + CCJ_MAXLOOP);
                    for (int dp = k + 1; dp < min_dp; ++dp)
                        ccj_cand(b, ccj_e_intP(M, S, d, j, k, dp) + ccj_get4(c, T_PM, i, d, dp, l), (j - d) * 64 + (dp - k));
                }
            }
            b = par.argmin(b);
            if (b.ord == 0) T.push(i, l, j - 1, k + 1, TB_P_PM);
            else if (b.ord > 0) T.push(i, l, j - b.ord / 64, k + b.ord % 64, TB_P_PM);
        } break;

        case TB_P_PMmloop: {  // :2398-2436
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            T.setpair(j, k, TB_P_PMmloop);
            const int b1 = ccj_get4(c, T_PMmloop10, i, j - 1, k + 1, l) + ap + bp;
            const int b2 = ccj_get4(c, T_PMmloop01, i, j - 1, k + 1, l) + ap + bp;
            T.push(i, l, j - 1, k + 1, b1 < b2 ? TB_P_PMmloop10 : TB_P_PMmloop01);
        } break;

        case TB_P_PMmloop00: {  // :2437-2497
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            T.setpair(j, k, TB_P_PMmloop);
            ccj_best b = {ccj_get4(c, T_PM, i, j, k, l) + bp, 0};
            for (int d = i + L; d < j; d += NL)
                ccj_cand(b, ccj_WB(c, d + 1, j) + ccj_get4(c, T_PMmloop00, i, d, k, l), 1 + d);
            for (int d = k + 1 + L; d <= l; d += NL)
                ccj_cand(b, ccj_get4(c, T_PMmloop00, i, j, d, l) + ccj_WB(c, k, d - 1), n + 3 + d);
            b = par.argmin(b);
            if (b.ord == 0) T.push(i, l, j, k, TB_P_PM);
            else if (b.ord < n + 3) { const int d = b.ord - 1; T.push(i, l, d, k, TB_P_PMmloop00); T.push2(d + 1, j, TB_P_WB); }
            else { const int d = b.ord - (n + 3); T.push(i, l, j, d, TB_P_PMmloop00); T.push2(k, d - 1, TB_P_WB); }
        } break;

        case TB_P_PMmloop01: {  // :2499-2540 (decomposition differs from the fill's)
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            ccj_best b = {ccj_get4(c, T_PMmloop01, i, j, k + 1, l) + cp, 0};
            for (int d = k + 1 + L; d <= l; d += NL)
                ccj_cand(b, ccj_get4(c, T_PMmloop00, i, j, d, l) + ccj_tri_get(c, T2_WBP, k, d - 1), d);
            b = par.argmin(b);
            if (b.ord == 0) T.push(i, l, j, k + 1, TB_P_PMmloop01);
            else { const int d = b.ord; T.push(i, l, j, d, TB_P_PMmloop00); T.push2(k, d - 1, TB_P_WBP); }
        } break;

        case TB_P_PMmloop10: {  // :2541-2583 (decomposition differs from the fill's)
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            ccj_best b = {ccj_get4(c, T_PMmloop10, i, j - 1, k, l) + cp, 0};
            for (int d = i + 1 + L; d < j; d += NL)
                ccj_cand(b, ccj_tri_get(c, T2_WBP, d, j) + ccj_get4(c, T_PMmloop00, i, d - 1, k, l), d);
            b = par.argmin(b);
            if (b.ord == 0) T.push(i, l, j - 1, k, TB_P_PMmloop10);
            else { const int d = b.ord; T.push(i, l, d - 1, k, TB_P_PMmloop00); T.push2(d, j, TB_P_WBP); }
        } break;

        case TB_P_POiloop: {  // :2584-2644 (impossible-test first; window reads PO.get(d,j,dp,k), always INF)
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            T.setpair(i, l, TB_P_POiloop);
            ccj_best b = {INF, -1};
            if (ccj_pt(c, i, l) > 0) {
                if (L == 0) ccj_cand(b, ccj_get4(c, T_PO, i + 1, j, k, l - 1) + ccj_e_stP(M, S, i, l), 0);
                const int max_d = ccj_min(j, i + CCJ_MAXLOOP);
                for (int d = i + 1 + L; d < max_d; d += NL) {
                    const int min_dp = ccj_max(l - CCJ_MAXLOOP, k);
                    for (int dp = l - 1; dp > min_dp; --dp)
                        ccj_cand(b, ccj_e_intP(M, S, i, d, dp, l) + ccj_get4(c, T_PO, d, j, dp, k), (d - i) * 64 + (l - dp));
                }
            }
            b = par.argmin(b);
            if (b.ord == 0) T.push(i + 1, l - 1, j, k, TB_P_PO);
            else if (b.ord > 0) T.push(i + b.ord / 64, k, j, l - b.ord % 64, TB_P_PO);
        } break;

        case TB_P_POmloop: {  // :2645-2682
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            T.setpair(i, l, TB_P_POmloop);
            const int b1 = ccj_get4(c, T_POmloop10, i + 1, j, k, l - 1) + ap + bp;
            const int b2 = ccj_get4(c, T_POmloop01, i + 1, j, k, l - 1) + ap + bp;
            T.push(i + 1, l - 1, j, k, b1 < b2 ? TB_P_POmloop10 : TB_P_POmloop01);
        } break;

        case TB_P_POmloop00: {  // :2684-2736 (row 2 pushes P_WBP where the fill used WB)
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            ccj_best b = {ccj_get4(c, T_PO, i, j, k, l) + bp, 0};
            for (int d = i + 1 + L; d <= j; d += NL)
                ccj_cand(b, ccj_WB(c, i, d - 1) + ccj_get4(c, T_POmloop00, d, j, k, l), 1 + d);
            for (int d = k + L; d < l; d += NL)
                ccj_cand(b, ccj_get4(c, T_POmloop00, i, j, k, d) + ccj_WB(c, d + 1, l), n + 3 + d);
            b = par.argmin(b);
            if (b.ord == 0) T.push(i, l, j, k, TB_P_PO);
            else if (b.ord < n + 3) { const int d = b.ord - 1; T.push(d, l, j, k, TB_P_POmloop00); T.push2(i, d - 1, TB_P_WBP); }
            else { const int d = b.ord - (n + 3); T.push(i, d, j, k, TB_P_POmloop00); T.push2(d + 1, l, TB_P_WB); }
        } break;

        case TB_P_POmloop01: {  // :2738-2768
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            ccj_best b = {INF, -1};
            for (int d = k + L; d < l; d += NL)
                ccj_cand(b, ccj_get4(c, T_POmloop00, i, j, k, d) + ccj_tri_get(c, T2_WBP, d + 1, l), d);
            b = par.argmin(b);
            T.push(i, b.ord, j, k, TB_P_POmloop00);
            T.push2(b.ord + 1, l, TB_P_WBP);
        } break;

        case TB_P_POmloop10: {  // :2769-2818
            if (ccj_tb_border(i, j, k, l)) { T.die(TBM_BORDER_CASES, type); return; }
            if (ccj_tb_imposs(n, i, j, k, l)) { T.die(TBM_IMPOSSIBLE_CASES, type); return; }
            ccj_best b = {INF, -1};
            for (int d = i + 1 + L; d <= j; d += NL)
                ccj_cand(b, ccj_tri_get(c, T2_WBP, i, d - 1) + ccj_get4(c, T_POmloop00, d, j, k, l), d);
            for (int d = k + 1 + L; d < l; d += NL)
                ccj_cand(b, ccj_get4(c, T_POmloop10, i, j, k, d) + ccj_WB(c, d + 1, l), n + 2 + d);
            b = par.argmin(b);
            if (b.ord < 0) return;
            if (b.ord < n + 2) { const int d = b.ord; T.push(d, l, j, k, TB_P_POmloop00); T.push2(i, d - 1, TB_P_WBP); }
            else { const int d = b.ord - (n + 2); T.push(i, d, j, k, TB_P_POmloop10); T.push2(d + 1, l, TB_P_WB); }
        } break;

        default:
            // P_PLiloop5 & co. are never produced.  P_PfromMprime / P_PfromMdoubleprime are handled above only
            // when reached through pseudo_loop::backtrack; W_final's dispatch list (src/W_final.cc:666-703)
            // lacks them, see ccj_traceback() below.
            break;
    }
}

template <class Par>
CCJ_HD void ccj_traceback(const ccj_cx &c, const Par &par) {
    ccj_tb T(c);
    const int n = c.q.n;
    T.push2(1, n, TB_FREE);
    while (T.top > 0 && !T.stop) {
        const int32_t *s = c.q.tb_stack + 5 * (T.top - 1);
        const int a0 = s[0], a1 = s[1], a2 = s[2], a3 = s[3], ty = s[4];
        --T.top;
        par.sync();
        // W_final::backtrack's switch has no case for these two (src/W_final.cc:666-715): it prints
        // "Should not be here!" on stdout and drops the node.
        if (ty == TB_P_PfromMprime || ty == TB_P_PfromMdoubleprime) {
            if (par.lane() == 0) c.q.status[1] += 1;
            continue;
        }
        ccj_tb_node(T, par, a0, a1, a2, a3, ty);
        par.sync();
    }
}
