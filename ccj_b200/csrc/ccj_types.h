// ccj_b200 -- shared host/device plain-data types, constants and table layouts.
//
// Everything here is POD so it can sit in device global memory and be filled by the host
// loader (energy_model.cpp).  Names follow the reference's domain vocabulary:
//   reference src/matrices.hh      (TriangleMatrix / Matrix4D semantics)
//   reference src/h_globals.hh:7-25 (pseudoknot penalties)
//   reference src/ViennaRNA/params/basic.h:57-118 (vrna_param_t fields we need)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CCJ_HD __host__ __device__ __forceinline__
#else
#define CCJ_HD inline
#endif

#define CCJ_INF 10000000        /* reference src/matrices.hh:10 */
#define CCJ_INTERN_INF 32767    /* Matrix4D::INTERN_INF, src/matrices.hh:150 */
#define CCJ_V_UNSET 10000       /* free_energy_node default, src/h_struct.hh:100 */
#define CCJ_TURN 3
#define CCJ_MAXLOOP 30

// ---- 4D gap tables (order = the reference's in-cell evaluation order is NOT this order; this is
//      the export order shared with oracle/ref_dump.cc) -------------------------------------------
enum ccj_table4 {
    T_PK = 0, T_PL, T_PR, T_PM, T_PO, T_PfromL, T_PfromR, T_PfromM, T_PfromMprime, T_PfromO,
    T_PLmloop00, T_PLmloop01, T_PLmloop10, T_PRmloop00, T_PRmloop01, T_PRmloop10,
    T_PMmloop00, T_PMmloop01, T_PMmloop10, T_POmloop00, T_POmloop01, T_POmloop10,
    CCJ_NT4 = 22,
    CCJ_NT4_STORE = 22   /* min(PL,PR) of a cell (get_PfromMdoubleprime without PB, pseudo_loop.cc:675-678) lives in record g3 only */
};

// ---- 2D tables ----------------------------------------------------------------------------------
enum ccj_table2 {
    T2_V = 0,   // free_energy_node::energy   (init 10000)
    T2_VTYPE,   // free_energy_node::type     (init 'N')
    T2_WM, T2_WMv, T2_WMp,  // init INF+1
    T2_P, T2_WBP, T2_WPP,   // init INF+1
    T2_WB, T2_WP,           // derived: get_WB/get_WP for 1<=i<=j<=n (src/pseudo_loop.cc:647-661)
    CCJ_NT2 = 10,
    CCJ_NT2_EXPORT = 8
};

// ---- energy model as the device sees it ------------------------------------------------------------
// Pair types 0..7 (0 = no pair), bases 0..4 (0 unused for ACGU input), exactly the index space of
// vrna_param_t so that every reference lookup has a 1:1 counterpart.
#define CCJ_MAX_SPECIAL 40
#define CCJ_LOOP_TAB 64      /* bulge / interior sizes 0..63 (PK interior windows reach u=56) */
#define CCJ_HAIRPIN_TAB 4096 /* hairpin sizes 0..4095, >30 extrapolated with the host libm */

struct ccj_model {
    int32_t stack[8][8];
    int32_t hairpin[CCJ_HAIRPIN_TAB];    // hairpin[size], incl. lxc*log extrapolation (hairpin.h:158-161)
    int32_t bulge[CCJ_LOOP_TAB];         // bulge[nl]      (internal.h:505-507)
    int32_t internal_loop[CCJ_LOOP_TAB]; // internal_loop[u] (internal.h:556-560)
    int32_t mismatchExt[8][5][5];
    int32_t mismatchI[8][5][5];
    int32_t mismatch1nI[8][5][5];
    int32_t mismatch23I[8][5][5];
    int32_t mismatchH[8][5][5];
    int32_t mismatchM[8][5][5];
    int32_t dangle5[8][5];
    int32_t dangle3[8][5];
    int32_t int11[8][8][5][5];
    int32_t int21[8][8][5][5][5];
    int32_t int22[8][8][5][5][5][5];
    int32_t ninio2;      // P->ninio[2]
    int32_t max_ninio;   // MAX_NINIO
    int32_t MLbase, MLclosing, TerminalAU;
    int32_t MLintern[8];
    // special hairpins, ASCII, matched against the ASCII sequence exactly like the reference's strstr
    // on P->Tetraloops etc. (hairpin.h:166-193).  Entries are blank separated with a fixed pitch, so a
    // blank-free probe of the entry length can only match at an entry start.
    int32_t n_tetra, n_tri, n_hexa;
    char tetra[CCJ_MAX_SPECIAL][8];
    char tri[CCJ_MAX_SPECIAL][8];
    char hexa[CCJ_MAX_SPECIAL][8];
    int32_t tetra_E[CCJ_MAX_SPECIAL], tri_E[CCJ_MAX_SPECIAL], hexa_E[CCJ_MAX_SPECIAL];
    int32_t special_hp;
    int32_t dangles;     // model_details.dangles (set AFTER scaling, src/W_final.cc:25)
    // pair[][] and rtype[] (src/ViennaRNA/pair_mat.h:20-38,80-155); noGU already applied
    int32_t pair[5][5];
    int32_t rtype[8];
    // pseudoknot penalties (src/h_globals.hh:7-25)
    int32_t PS_penalty, PSM_penalty, PSP_penalty, PB_penalty, PUP_penalty, PPS_penalty;
    int32_t a_penalty, b_penalty, c_penalty, ap_penalty, bp_penalty, cp_penalty;
    double e_stP_penalty, e_intP_penalty;
};

// ---- per-sequence view ----------------------------------------------------------------------------
struct ccj_seq {
    int32_t n;
    int32_t pad_;
    const int8_t *S;     // S[0..n+1]; S[i] base code of nucleotide i (1-based), S[n+1]=S[1], S[0]=S[n]
    const char *seq;     // ASCII, 0-based, n chars (upper case)
    int16_t *t4;         // CCJ_NT4 tables, each `stride4` int16
    int64_t stride4;
    int32_t *t2;         // CCJ_NT2 tables, each `stride2` int32, diagonal-major
    int64_t stride2;
    int32_t *W;          // W[0..n]
    // traceback outputs
    int32_t *pair_out;   // f[1..n].pair  (index 0 unused), -1 = unpaired
    int8_t *ftype_out;   // f[1..n].type
    int32_t *status;     // [0]=status code, [1]=number of "Should not be here!" lines, [2]=message id,
                         // [3],[4]=i,j of a "NOT GOOD RESTR INTER" exit; CCJ_STATUS_INTS ints
    // per-sequence precomputation for the tuned 4D kernel (ccj_fill4.cu)
    int32_t *w3;         // int4 {WB,WP,WBP,0} per 2D index (diagonal-major), written with the 2D tables
    int32_t *estP;       // get_e_stP(i,j) per 2D index
    uint32_t *inlist;    // interior-loop partners INSIDE closing pair (i,j): slot tri(i,j)*CCJ_WIN, see ccj_fill4.cu
    uint32_t *outlist;   // interior-loop partners OUTSIDE inner pair (j,k)
    int32_t *incnt, *outcnt;  // entries per slot
    // "read-group" copies of the gap tables (tuned path): the tables one split-point pattern reads at the SAME
    // cell are interleaved as one record per cell, so a pattern step is one 12/16-byte read per cell instead
    // of 5-7 two-byte reads from as many arrays (ccj_fill4.cu).  Indexed by ccj_idx4 * record length.
    //   g1 (12 B): PK PfromL PfromMprime PLmloop00 PLmloop10 PMmloop00          read as X(i,d,k,l)
    //   g2 (12 B): PfromL PfromO PLmloop00 PMmloop00 POmloop00 -                read as X(d,j,k,l)
    //   g3 (12 B): PK PfromR min(PL,PR) PRmloop00 PMmloop00 -                   read as X(i,j,d,l)
    //   g4 (16 B): PfromR PfromO PRmloop00 PMmloop00 PMmloop10 POmloop00 POmloop10 -   read as X(i,j,k,d)
    int16_t *g1, *g2, *g3, *g4;
    int32_t *lay;        // layout tables (tuned path, n<=448): Tet, Cb, H4, HH4, CbW4 with n+1 ints each, then CF, DF, S2, EG
                         // with n+2 ints each (CCJ_LAY_*)
    int16_t *pkf, *pkg;  // first- and second-factor PK copies for compute_P ("PK copies for compute_P" below)
    // copies of PL / PR / PM for the interior windows, rows padded to 4 entries so that a lane moves 4 cells
    // per 8-byte load ("window layouts" below); written by k_final, read by k_winLR / k_winM only
    int16_t *plw, *prw, *pmw, *pmm;   // pmm: the mask halves of PMW (same quad index, separate array)
    int16_t *wscr;       // window partial minima in the same row layouts: [parity][PL,PR] x wscr_lr, then [parity] x 4*wtot4
    int64_t wscr_lr;     // entries of one PL/PR partial = largest padded level
    int32_t *pmlev4;     // (n+1)^2: quad offset of row (j,k) inside one PM level (ccj_pmw_quad)
    int32_t wtot4;       // quads of one PM level
    int32_t npairs;      // unused on the host
    int32_t *plist;      // plist[s*(n+1)+x], x=0..: the 5' ends i of the pairs (i,i+s), ascending
    int32_t *pcum;       // pcum[s*(n+2)+x]: number of pairs (i,i+s) with i<=x
    int32_t *pmlist;     // all pairs as j | k<<16, ordered by span k-j and then j
    int32_t *pmstart;    // pmstart[s]: number of pairs with span < s, s=0..n+1
    int16_t *scratch;    // per-level partial minima: [partial id][cell of the level], see ccj_fill4.cu
    int64_t scratch_stride;   // cells of the largest level
    int32_t *tb_stack;   // traceback stack, 5 ints per node
    int32_t tb_cap;      // capacity in nodes
    int32_t pad2_;
    // row-sharded fold of ONE oversized sequence ("sharded layout" below; all zero for ordinary waves)
    int32_t shard_G;             // ranks the rows i are dealt to (row i belongs to rank (i-1) mod G); 0 = not sharded
    int32_t shard_rank;          // this rank
    int32_t shard_shift;         // log2(G) if G is a power of two, else -1
    int32_t use_lists;           // the generic cell functions walk inlist / outlist (filled by k_prep) instead of scanning
                                 // the 29 x 29 windows with can_pair tests
    const int64_t *shard_lev;    // lev[t] = sum_{t'<t} C(t'), t = 0..n (device)
    int16_t *shard_rep;          // the 12 column-read tables, every rank's rows (filled by the allgather per level)
    int16_t *const *shard_loc;   // [G] base of each rank's 10 row-local tables; only [shard_rank] is set unless the peers'
                                 // memory was opened (the traceback rank)
    int8_t shard_kind[24];       // table id -> 0..11 (column-read) or 12..21 (row-local)
};
#define CCJ_STATUS_INTS 8
#define CCJ_WIN 841      /* 29*29 window slots of get_P{L,R,M}iloop (src/pseudo_loop.cc:694-700) */
#define CCJ_WIN_IN 848   /* pitch of an inlist slot (4-byte entries, zero-padded to a multiple of 8) */
#define CCJ_WIN_OUT 848  /* pitch of an outlist slot (8-byte entries: packed term, quad offset of the partner's PMW row) */
CCJ_HD int ccj_tri(int i, int j) { return (j - 1) * (j - 2) / 2 + (i - 1); } /* 1<=i<j<=n -> [0, n(n-1)/2) */

// ---- 4D layout ---------------------------------------------------------------------------------------
// A cell is (i,j,k,l) with 1<=i<=j, j<k-1, k<=l<=n  (Matrix4D::get validity, src/matrices.hh:177-182).
// We address it by arm lengths a=j-i, b=l-k and the anchors (i,k):  nesting [b][a][i][k], k fastest.
// One (a,b) "slab" holds the packed triangle {i=1..m, k=i+a+2..n-b}, m=n-a-b-2, row i of length m+1-i.
// All cells of one DP level t=a+b (the wavefront step) with equal (a,b) are therefore ONE contiguous
// run of m(m+1)/2 int16: a warp whose lanes walk that run writes whole 64-byte lines, and the split-point
// reads X(i,d,k,l) / X(d,j,k,l) / X(i,j,d,l) / X(i,j,k,d) of neighbouring lanes are neighbours too.
//   offset = Cb(b) - Tet(m) + (i-1)(2m+2-i)/2 + (k-j-2),   Cb(b) = Pent(n-2) - Pent(n-b-3)
CCJ_HD int64_t ccj_pent(int64_t q) { return q * (q + 1) * (q + 2) * (q + 3) / 24; }
CCJ_HD int64_t ccj_tet(int64_t q) { return q * (q + 1) * (q + 2) / 6; }
CCJ_HD int64_t ccj_cells4(int n) { return n >= 3 ? ccj_pent(n - 2) : 0; } /* == C(n+1,4) */
CCJ_HD int64_t ccj_cb(int n, int b) { return ccj_pent(n - 2) - ccj_pent(n - b - 3); }

CCJ_HD int64_t ccj_idx4(int n, int i, int j, int k, int l) {
    const int a = j - i, b = l - k;
    const int64_t m = n - a - b - 2;
    return ccj_cb(n, b) - ccj_tet(m) + (int64_t)(i - 1) * (2 * m + 2 - i) / 2 + (k - j - 2);
}

// cells of the largest level of a length-n fold: max_t (t+1) * m(m+1)/2, m = n-t-2
CCJ_HD int64_t ccj_level_max(int n) {
    int64_t best = 0;
    for (int t = 0; t <= n - 3; ++t) {
        const int64_t m = n - t - 2, c = (int64_t)(t + 1) * m * (m + 1) / 2;
        if (c > best) best = c;
    }
    return best;
}

// ---- PK copies for compute_P (src/pseudo_loop.cc:166-179) -------------------------------------------------------
// P(i,l) = min_{j,d,k} PK(i,j,d+1,k) + PK(j+1,d,k+1,l).  With delta=k-d, both factors of the terms of one (i,j,l) are
// rows over d=j+1.. : row delta of block (i,j) of the first-factor copy PKF (nesting [i][j][delta][d], entries
// d=j+1..n-delta) and row delta of block (j+1,l) of the second-factor copy PKG ([i2][l][delta][d], d=i2..l-delta-1).
// Every row starts on an 8-entry boundary in both copies, so 8 consecutive terms are one 16-byte load from each.
// A level writes whole rows (delta is fixed per level and block); PKF rows are written coalesced.
//   Q8(x)  = sum_{u<=x} ceil(u/8)  (rows of 1..x entries, in octets);  a block with rows of M, M-1, .., 1 entries
//            has 8*Q8(M) entries and its row delta starts at 8*(Q8(M) - Q8(M+1-delta))
//   PKF: block (i,j), M=n-j-1:   base = DF[i] + CF[j] - CF[i],  CF[j] = sum_{j'<j} 8*Q8(n-j'-1), DF[i] = sum_{i'<i} (CF[n+1]-CF[i'])
//   PKG: block (i2,l), M=l-i2-1: base = EG[i2] + S2[M],         S2[y] = sum_{y'<y} 8*Q8(y'),     EG[i] = sum_{i'<i} S2[n-i']
CCJ_HD int64_t ccj_q8(int64_t x) { const int64_t g = x >> 3, r = x & 7; return x > 0 ? 4 * g * (g + 1) + r * (g + 1) : 0; }
#define CCJ_LAY_CF(n) (5 * ((n) + 1))
#define CCJ_LAY_DF(n) (5 * ((n) + 1) + ((n) + 2))
#define CCJ_LAY_S2(n) (5 * ((n) + 1) + 2 * ((n) + 2))
#define CCJ_LAY_EG(n) (5 * ((n) + 1) + 3 * ((n) + 2))
#define CCJ_LAY_INTS(n) (5 * ((n) + 1) + 4 * ((n) + 2) + 8)
inline int64_t ccj_pkf_total(int n) {   // DF[n+1]
    int64_t cf_total = 0, tot = 0, cf = 0;
    for (int j = 1; j <= n; ++j) cf_total += 8 * ccj_q8(n - j - 1);
    for (int i = 1; i <= n; ++i) { tot += cf_total - cf; cf += 8 * ccj_q8(n - i - 1); }
    return tot > 0 ? tot : 8;
}
inline int64_t ccj_pkg_total(int n) {   // EG[n+1]
    int64_t tot = 0;
    for (int i = 1; i <= n; ++i) {
        int64_t s2 = 0;   // S2[n-i]
        for (int y = 1; y < n - i; ++y) s2 += 8 * ccj_q8(y);
        tot += s2;
    }
    return tot > 0 ? tot : 8;
}

// ---- window layouts ------------------------------------------------------------------------------------
// The interior windows (get_P{L,R,M}iloop, src/pseudo_loop.cc:682-773) read, for a run of cells that share a
// closing pair, the SAME run shifted to another slab.  Their copies store every run with its start on a
// 4-entry (8-byte) boundary and the coordinate that the window leaves unchanged as the position inside the
// run, so that source and target quads line up and a lane moves 4 cells per load:
//   PLW: slab (a,b), row r=i-1 (length m-r), position (n-b)-k       (window keeps k,l; source slab (a-s,b))
//   PRW: slab (a,b), row kr=n-b-k (length m-kr), position i-1       (window keeps i,j; source slab (a,b-s))
//        slab base = 4*(CbW4(b) + HH4(n-b-2) - HH4(m)), row start = 4*(H4(m) - H4(m-r))
//        H4(x)  = quads of a triangle with rows 1..x, each padded to a multiple of 4
//        HH4(x) = sum_{m<=x} H4(m),  CbW4(b) = sum_{b'<b} HH4(n-b'-2)
//   PMW: level-major; row (j,k) of level t holds i=j-a, position (i-1) - 4*((max(j-t,1)-1)>>2)
//        (window keeps i,l; source row (j-x,k+y) of level t-x-y); row pitch W4(j,k) quads, see ccj_pmw_w4.
//        Every PMW quad (4 values, 8 bytes) has a quad of mask halves at the same index of a second array (PMM).
//        For a cell the window may read (a>=1, b>=1) value = PM and mask = -32768; for the end cells and the padding
//        both are 32767.  Candidates with an energy >= 0 only read PMW (32767 + energy saturates by itself); for the
//        few negative ones max(value+energy, mask) masks two cells per instruction.
CCJ_HD int64_t ccj_h4(int64_t x) { const int64_t q = x >> 2, r = x & 3; return (q + 1) * (2 * q + r); }
inline int64_t ccj_hh4(int x) { int64_t s = 0; for (int m = 1; m <= x; ++m) s += ccj_h4(m); return s; }
inline int64_t ccj_winlr_quads(int n) { int64_t s = 0; for (int b = 0; b <= n - 3; ++b) s += ccj_hh4(n - b - 2); return s; }
inline int64_t ccj_winlr_level_max(int n) {  // entries of the largest padded level
    int64_t best = 4;
    for (int t = 0; t <= n - 3; ++t) { const int64_t c = 4 * (int64_t)(t + 1) * ccj_h4(n - t - 2); if (c > best) best = c; }
    return best;
}
CCJ_HD int ccj_pmw_w4(int n, int j, int k) { const int A = j - 1, B = n - k; return (((A < B ? A : B) + 4) >> 2) + 1; }
inline int64_t ccj_pmw_level_quads(int n) {
    int64_t s = 0;
    for (int j = 1; j <= n; ++j) for (int k = j + 2; k <= n; ++k) s += ccj_pmw_w4(n, j, k);
    return s > 0 ? s : 1;
}

CCJ_HD bool ccj_valid4(int i, int j, int k, int l) { return i <= j && j < k - 1 && k <= l; }

// ---- sharded layout (one sequence whose gap tables exceed one GPU; SURVEY.md 8e, src/pseudo_loop.cc:69-132) ---------
// Rows i are dealt cyclically: row i belongs to rank r=(i-1) mod G as its row q=(i-1) div G.  Storage is level-major
// (level t=(j-i)+(l-k), m=n-t-2 rows of m+1-i cells per slab a=j-i): rank r's part of one slab is its rows
// q=0..Q-1, Q=ceil((m-r)/G), row q holding m-r-qG cells, S_r(m) = Q(m-r) - G Q(Q-1)/2 cells in all.  A level reserves
// C(t) = (t+1) S_0(m) cells per table and rank (rank 0 owns the most rows).
//   column-read tables (12: PK PL PO PfromL PfromO PLmloop00/01/10 PMmloop00 POmloop00/01/10 -- read at rows d > i):
//       every rank holds all ranks' rows:   rep[ 12 G lev[t] + (12 r + kind) C(t) + inner ]
//       -> the 12 tables of rank r's rows of one level are ONE contiguous block, and the G blocks of a level are
//          adjacent: the per-level exchange is a single in-place ncclAllGather
//   row-local tables (10: PR PM PfromR PfromM PfromMprime PRmloop00/01/10 PMmloop01/10 -- read along row i only):
//       only the owner holds them:          loc_r[ 10 lev[t] + (kind-12) C(t) + inner ]
//   inner = a S_r(m) + q (m-r) - G q(q-1)/2 + (k-j-2)
#define CCJ_SHARD_NREP 12
#define CCJ_SHARD_NLOC 10
// (all of this fits 32 bits for n < 1600: a S_r(m) < n^3/2; only the level bases need 64 bits)
CCJ_HD int ccj_shard_rows(int m, int r, int G) { return m > r ? (m - r + G - 1) / G : 0; }
CCJ_HD int64_t ccj_shard_slab(int m, int r, int G) {
    const int Q = ccj_shard_rows(m, r, G);
    return Q * (m - r) - G * (Q * (Q - 1) / 2);
}
CCJ_HD int64_t ccj_shard_level_cells(int n, int t, int G) { return n - t - 2 >= 1 ? (int64_t)(t + 1) * ccj_shard_slab(n - t - 2, 0, G) : 0; }
// sh = log2(G) when G is a power of two (the usual 1/2/4/8 ranks: no integer division on the device), else -1
CCJ_HD int ccj_shard_inner_fast(int n, int G, int sh, int i, int j, int k, int l, int &owner) {
    const int a = j - i, m = n - a - (l - k) - 2;
    int r, q, Q;
    if (sh >= 0) {
        r = (i - 1) & (G - 1);
        q = (i - 1) >> sh;
        Q = (m - r + G - 1) >> sh;
    } else {
        r = (i - 1) % G;
        q = (i - 1) / G;
        Q = (m - r + G - 1) / G;
    }
    owner = r;
    const int mr = m - r;
    const int slab = Q * mr - G * (Q * (Q - 1) / 2);
    return a * slab + q * mr - G * (q * (q - 1) / 2) + (k - j - 2);
}
CCJ_HD int64_t ccj_shard_inner(int n, int G, int i, int j, int k, int l) {
    int owner;
    return ccj_shard_inner_fast(n, G, -1, i, j, k, l, owner);
}
// inverse of the row offsets of one rank's part of a slab: cell p (0 <= p < S_r(m)) -> own row q and position kk.
// Row q holds mr - qG cells (mr = m - r) and starts at q mr - G q(q-1)/2.
CCJ_HD void ccj_shard_cell_of(int p, int mr, int G, int Q, int &q, int &kk) {
    const float A = 2.f * (float)mr + (float)G;
#if defined(__CUDA_ARCH__)
    int qq = (int)((A - sqrtf(A * A - 8.f * (float)G * (float)p)) / (2.f * (float)G));
#else
    int qq = (int)((A - __builtin_sqrtf(A * A - 8.f * (float)G * (float)p)) / (2.f * (float)G));
#endif
    if (qq < 0) qq = 0;
    if (qq > Q - 1) qq = Q - 1;
    while (qq > 0 && qq * mr - G * (qq * (qq - 1) / 2) > p) --qq;
    while (qq + 1 < Q && (qq + 1) * mr - G * ((qq + 1) * qq / 2) <= p) ++qq;
    q = qq;
    kk = p - (qq * mr - G * (qq * (qq - 1) / 2));
}
inline void ccj_shard_kinds(int8_t *kind24) {
    static const int rep[CCJ_SHARD_NREP] = {T_PK, T_PL, T_PO, T_PfromL, T_PfromO, T_PLmloop00, T_PLmloop01, T_PLmloop10,
                                            T_PMmloop00, T_POmloop00, T_POmloop01, T_POmloop10};
    for (int t = 0; t < 24; ++t) kind24[t] = -1;
    for (int x = 0; x < CCJ_SHARD_NREP; ++x) kind24[rep[x]] = (int8_t)x;
    int nl = 0;
    for (int t = 0; t < CCJ_NT4; ++t)
        if (kind24[t] < 0) kind24[t] = (int8_t)(CCJ_SHARD_NREP + nl++);
}

// ---- 2D layout: diagonal-major, idx = (j-i)*(n+1) + i, 1<=i<=j<=n ----------------------------------
CCJ_HD int64_t ccj_stride2(int n) { return (int64_t)n * (n + 1) + (n + 1); }
CCJ_HD int32_t ccj_idx2(int n, int i, int j) { return (j - i) * (n + 1) + i; }

// status codes written by the traceback (host maps them back to the reference's messages / exit codes)
enum ccj_status {
    CCJ_OK = 0,
    CCJ_EXIT_FAILURE = 1,       // reference printed a message to stderr and exit(EXIT_FAILURE)
    CCJ_EXIT_ZERO_NOT_GOOD = 2, // "NOT GOOD RESTR INTER" then exit(0) (src/W_final.cc:232-235)
    CCJ_STACK_OVERFLOW = 3      // our traceback stack was too small (never the reference's behaviour)
};
