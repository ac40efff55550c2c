// Tuned level-wavefront kernel for the 22 gap tables (sm_100a).
//
// One launch = one DP level t=(j-i)+(l-k).  blockIdx.y -> a=j-i (b=t-a), blockIdx.z -> sequence,
// threads walk the packed (i,k) triangle of slab (a,b), which is one contiguous run in every table
// (layout in ccj_types.h) -> all 23 stores of a warp are full 64-byte lines.
//
// The reference evaluates, per cell, 22 recurrences one after the other; each is a min over split
// points d of  X(neighbour cell) + 2D-term  (src/pseudo_loop.cc:181-644).  Candidates only read cells of
// LOWER levels, so their order is free.  We regroup them by the 4 access patterns
//     L1: X(i,d,k,l)   L2: X(d,j,k,l)   R3: X(i,j,d,l)   R4: X(i,j,k,d)
// so that one offset computation and one 16-byte {WB,WP,WBP} load serve 5-7 tables, and apply the
// same-cell terms afterwards in the reference's in-cell order (:85-127), which is what fixes the
// "unset = 32767" reads.  The interior-loop windows (get_P{L,R,M}iloop, :682-773) walk per-pair lists of
// pairable partners with pre-rounded energies instead of the 29x29 can_pair-gated scan.
// Checked bit-for-bit against ccj_cell4d (generic version) and the reference's tables.
#include "ccj_kernels.cuh"
#include "ccj_cells4.cuh"

namespace ccj {

// ---------------------------------------------------------------------------------------------
// per-sequence precomputation: e_stP table and window partner lists
// entry = (uint16)energy | x<<16 | y<<24,  x,y in 1..29
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pack_entry(int e, int x, int y, uint32_t &out, int32_t *status) {
    // a term >= INF/2 + 32768 can never bring a minimum below INF/2: dropping it is exact
    if (e >= CCJ_INF / 2 + 40000) return false;
    if (e > 32767 || e < -32768) {
        status[5] = 1;  // energy outside the packed range: the tuned path does not apply to this model
        return false;
    }
    out = (uint32_t)(uint16_t)(int16_t)e | ((uint32_t)x << 16) | ((uint32_t)y << 24);
    return true;
}

__global__ void __launch_bounds__(128) k_prep(const ccj_model *M, const ccj_seq *seqs) {
    ccj_cx c;
    c.M = M;
    c.q = seqs[blockIdx.y];
    const int n = c.q.n;
    const int j = blockIdx.x + 2;
    if (j > n) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int8_t *S = c.q.S;
    for (int i = 1 + wid; i < j; i += nw) {
        const int ij = ccj_idx2(n, i, j);
        if (lane == 0) c.q.estP[ij] = (j - i >= 2) ? ccj_e_stP(M, S, i, j) : CCJ_INF;
        const int slot = ccj_tri(i, j);
        int nin = 0, nout = 0;
        if (ccj_can_pair(c, i, j)) {
            uint32_t *il = c.q.inlist + (int64_t)slot * CCJ_WIN;
            uint32_t *ol = c.q.outlist + (int64_t)slot * CCJ_WIN;
            for (int s0 = 0; s0 < CCJ_WIN; s0 += 32) {
                const int s = s0 + lane;
                const int x = s / 29 + 1, y = s % 29 + 1;
                // inside (i,j): d=i+x, dp=j-y  (get_PLiloop / get_PRiloop window)
                uint32_t ent = 0;
                bool ok = false;
                if (s < CCJ_WIN) {
                    const int d = i + x, dp = j - y;
                    if (x <= j - i - 1 && dp >= d + 4 && ccj_can_pair(c, d, dp))
                        ok = pack_entry(ccj_e_intP(M, S, i, d, dp, j), x, y, ent, c.q.status);
                }
                unsigned bal = __ballot_sync(0xffffffffu, ok);
                if (ok) il[nin + __popc(bal & ((1u << lane) - 1))] = ent;
                nin += __popc(bal);
                // outside (i,j) as inner pair (j,k):=(i,j): d=i-x, dp=j+y  (get_PMiloop window)
                ok = false;
                if (s < CCJ_WIN) {
                    const int d = i - x, dp = j + y;
                    if (d >= 1 && dp <= n && ccj_can_pair(c, d, dp))
                        ok = pack_entry(ccj_e_intP(M, S, d, i, j, dp), x, y, ent, c.q.status);
                }
                bal = __ballot_sync(0xffffffffu, ok);
                if (ok) ol[nout + __popc(bal & ((1u << lane) - 1))] = ent;
                nout += __popc(bal);
            }
        }
        if (lane == 0) {
            c.q.incnt[slot] = nin;
            c.q.outcnt[slot] = nout;
        }
    }
}

// ---------------------------------------------------------------------------------------------
#define K4_THREADS 128
#define K4_MAXN 448  // int32 offsets and the shared tables below

struct T4 {
    const int16_t *__restrict__ p[CCJ_NT4_STORE];
};

__device__ __forceinline__ int ld16(const int16_t *__restrict__ p, int off) { return (int)__ldg(p + off); }
__device__ __forceinline__ int amin(int acc, int x, int w) { return min(acc, x + w); }

__global__ void __launch_bounds__(K4_THREADS) k_4d_v2(const ccj_model *__restrict__ M, const ccj_seq *__restrict__ seqs, int t) {
    __shared__ int s_tet[K4_MAXN + 4];  // Tet(x)
    __shared__ int s_cb[K4_MAXN + 4];   // Cb(b)
    const ccj_seq q = seqs[blockIdx.z];
    const int n = q.n;
    const int m = n - t - 2;
    if (m < 1) return;
    const int ncell = m * (m + 1) / 2;
    if ((int)(blockIdx.x * K4_THREADS) >= ncell) return;
    for (int x = threadIdx.x; x <= n; x += K4_THREADS) {
        s_tet[x] = (int)ccj_tet(x);
        s_cb[x] = x <= n - 3 ? (int)ccj_cb(n, x) : 0;
    }
    __syncthreads();
    const int p = blockIdx.x * K4_THREADS + threadIdx.x;
    if (p >= ncell) return;

    const int a = blockIdx.y, b = t - a;
    // invert p = (i-1)(2m+2-i)/2 + kk  (rows i=1..m of length m+1-i)
    int r = (int)(((2 * m + 1) - sqrtf((float)((2 * m + 1) * (2 * m + 1) - 8 * p))) * 0.5f);
    if (r < 0) r = 0;
    if (r > m - 1) r = m - 1;
    while (r > 0 && r * (2 * m + 1 - r) / 2 > p) --r;
    while ((r + 1) * (2 * m - r) / 2 <= p) ++r;
    const int i = r + 1;
    const int kk = p - r * (2 * m + 1 - r) / 2;
    const int j = i + a, k = j + 2 + kk, l = k + b;

    const int16_t *__restrict__ t4 = q.t4;
    const int64_t st4 = q.stride4;
#define TB(tbl) (t4 + (int64_t)(tbl) * st4)
    const int4 *__restrict__ W3 = reinterpret_cast<const int4 *>(q.w3);
    const int n1 = n + 1;
    // generic offset of cell (ii, ii+aa, kx, kx+bb)
#define OFF(aa, bb, ii, kx) (s_cb[bb] - s_tet[n - (aa) - (bb)-2] + ((((ii)-1) * (2 * (n - (aa) - (bb)-2) + 2 - (ii))) >> 1) + ((kx) - (ii) - (aa)-2))
    const int off0 = OFF(a, b, i, k);
    const int INF = CCJ_INF;
    const int bp = M->bp_penalty, cp = M->cp_penalty, PB = M->PB_penalty, apbp = M->ap_penalty + M->bp_penalty;

    // ---------------- left-arm split points: d = i+ap, ap = 0..a ----------------
    int aPK1 = INF, aPfL2 = INF, aPfM = INF, aPLm00 = INF, aPLm01 = INF, aPLm10 = INF, aPMm00 = INF;
    int aPfL1 = INF, aPfO1 = INF, aPMm10 = INF, aPOm00 = INF, aPOm10 = INF;
    if (a >= 1) {
        const int16_t *__restrict__ pPLm00 = TB(T_PLmloop00), *__restrict__ pPMm00 = TB(T_PMmloop00),
                                    *__restrict__ pPOm00 = TB(T_POmloop00);
        {  // L1 boundary d=i: X(i,i,k,l) with W(i+1,j)
            const int o = OFF(0, b, i, k);
            const int4 w = __ldg(&W3[(a - 1) * n1 + i + 1]);
            const int x = ld16(pPLm00, o);
            aPLm00 = amin(aPLm00, x, w.x);
            aPLm01 = amin(aPLm01, x, w.z);
            aPMm00 = amin(aPMm00, ld16(pPMm00, o), w.x);
        }
        {  // L2 boundary d=j: X(j,j,k,l) with W(i,j-1)
            const int o = OFF(0, b, j, k);
            const int4 w = __ldg(&W3[(a - 1) * n1 + i]);
            const int x = ld16(pPLm00, o);
            aPLm00 = amin(aPLm00, x, w.x);
            aPLm10 = amin(aPLm10, x, w.z);
            aPMm10 = amin(aPMm10, ld16(pPMm00, o), w.z);
            const int y = ld16(pPOm00, o);
            aPOm00 = amin(aPOm00, y, w.x);
            aPOm10 = amin(aPOm10, y, w.z);
        }
        const int16_t *__restrict__ pPK = TB(T_PK), *__restrict__ pPfL = TB(T_PfromL), *__restrict__ pPfMp = TB(T_PfromMprime),
                                    *__restrict__ pPLm10 = TB(T_PLmloop10), *__restrict__ pPfO = TB(T_PfromO);
        const int ub = n - b - 2;
#pragma unroll 2
        for (int ap = 1; ap < a; ++ap) {
            // L1: cell (i, i+ap, k, l), 2D term at (i+ap+1, j)
            const int m1 = ub - ap;
            const int o1 = s_cb[b] - s_tet[m1] + (((i - 1) * (2 * m1 + 2 - i)) >> 1) + (k - i - ap - 2);
            const int4 w1 = __ldg(&W3[(a - ap - 1) * n1 + i + ap + 1]);
            // L2: cell (d, j, k, l) with d=i+ap (arm length a-ap), 2D term at (i, d-1)
            const int d = i + ap;
            const int m2 = ub - (a - ap);
            const int o2 = s_cb[b] - s_tet[m2] + (((d - 1) * (2 * m2 + 2 - d)) >> 1) + (k - j - 2);
            const int4 w2 = __ldg(&W3[(ap - 1) * n1 + i]);
            aPK1 = amin(aPK1, ld16(pPK, o1), w1.y);
            aPfL2 = amin(aPfL2, ld16(pPfL, o1), w1.y);
            aPfM = amin(aPfM, ld16(pPfMp, o1), w1.y);
            const int x1 = ld16(pPLm00, o1);
            aPLm00 = amin(aPLm00, x1, w1.x);
            aPLm01 = amin(aPLm01, x1, w1.z);
            aPLm10 = amin(aPLm10, ld16(pPLm10, o1), w1.x);
            aPMm00 = amin(aPMm00, ld16(pPMm00, o1), w1.x);
            aPfL1 = amin(aPfL1, ld16(pPfL, o2), w2.y);
            aPfO1 = amin(aPfO1, ld16(pPfO, o2), w2.y);
            const int x2 = ld16(pPLm00, o2);
            aPLm00 = amin(aPLm00, x2, w2.x);
            aPLm10 = amin(aPLm10, x2, w2.z);
            aPMm10 = amin(aPMm10, ld16(pPMm00, o2), w2.z);
            const int y2 = ld16(pPOm00, o2);
            aPOm00 = amin(aPOm00, y2, w2.x);
            aPOm10 = amin(aPOm10, y2, w2.z);
        }
    }

    // ---------------- right-arm split points: d = k+bq, bq = 0..b ----------------
    int aPK3 = INF, aPfR1 = INF, aPfMp = INF, aPRm00 = INF, aPRm10 = INF;
    int aPfR2 = INF, aPfO2 = INF, aPRm01 = INF, aPMm01 = INF, aPOm01 = INF;
    if (b >= 1) {
        const int16_t *__restrict__ pPRm00 = TB(T_PRmloop00), *__restrict__ pPMm00 = TB(T_PMmloop00),
                                    *__restrict__ pPOm00 = TB(T_POmloop00);
        {  // R3 boundary d=l: X(i,j,l,l) with W(k,l-1)
            const int o = OFF(a, 0, i, l);
            const int4 w = __ldg(&W3[(b - 1) * n1 + k]);
            const int x = ld16(pPRm00, o);
            aPRm00 = amin(aPRm00, x, w.x);
            aPRm10 = amin(aPRm10, x, w.z);
            aPMm00 = amin(aPMm00, ld16(pPMm00, o), w.x);
        }
        {  // R4 boundary d=k: X(i,j,k,k) with W(k+1,l)
            const int o = OFF(a, 0, i, k);
            const int4 w = __ldg(&W3[(b - 1) * n1 + k + 1]);
            const int x = ld16(pPRm00, o);
            aPRm00 = amin(aPRm00, x, w.x);
            aPRm01 = amin(aPRm01, x, w.z);
            aPMm01 = amin(aPMm01, ld16(pPMm00, o), w.z);
            const int y = ld16(pPOm00, o);
            aPOm00 = amin(aPOm00, y, w.x);
            aPOm01 = amin(aPOm01, y, w.z);
        }
        const int16_t *__restrict__ pPK = TB(T_PK), *__restrict__ pPfR = TB(T_PfromR), *__restrict__ pMpp = TB(T_MPP),
                                    *__restrict__ pPfO = TB(T_PfromO), *__restrict__ pPMm10 = TB(T_PMmloop10),
                                    *__restrict__ pPOm10 = TB(T_POmloop10);
        const int ua = n - a - 2;
        const int ri = i - 1;
#pragma unroll 2
        for (int bq = 1; bq < b; ++bq) {
            // R3: cell (i, j, d, l) with d=k+bq (arm length b-bq), 2D term at (k, d-1)
            const int b3 = b - bq;
            const int m3 = ua - b3;
            const int o3 = s_cb[b3] - s_tet[m3] + ((ri * (2 * m3 + 2 - i)) >> 1) + (k + bq - j - 2);
            const int4 w3 = __ldg(&W3[(bq - 1) * n1 + k]);
            // R4: cell (i, j, k, d) with d=k+bq (arm length bq), 2D term at (d+1, l)
            const int m4 = ua - bq;
            const int o4 = s_cb[bq] - s_tet[m4] + ((ri * (2 * m4 + 2 - i)) >> 1) + (k - j - 2);
            const int4 w4 = __ldg(&W3[(b - bq - 1) * n1 + k + bq + 1]);
            aPK3 = amin(aPK3, ld16(pPK, o3), w3.y);
            aPfR1 = amin(aPfR1, ld16(pPfR, o3), w3.y);
            aPfMp = amin(aPfMp, ld16(pMpp, o3), w3.y);
            const int x3 = ld16(pPRm00, o3);
            aPRm00 = amin(aPRm00, x3, w3.x);
            aPRm10 = amin(aPRm10, x3, w3.z);
            aPMm00 = amin(aPMm00, ld16(pPMm00, o3), w3.x);
            aPfR2 = amin(aPfR2, ld16(pPfR, o4), w4.y);
            aPfO2 = amin(aPfO2, ld16(pPfO, o4), w4.y);
            const int x4 = ld16(pPRm00, o4);
            aPRm00 = amin(aPRm00, x4, w4.x);
            aPRm01 = amin(aPRm01, x4, w4.z);
            aPMm01 = amin(aPMm01, ld16(pPMm00, o4), w4.z);
            aPMm10 = amin(aPMm10, ld16(pPMm10, o4), w4.x);
            const int y4 = ld16(pPOm00, o4);
            aPOm00 = amin(aPOm00, y4, w4.x);
            aPOm01 = amin(aPOm01, y4, w4.z);
            aPOm10 = amin(aPOm10, ld16(pPOm10, o4), w4.x);
        }
    }

    // ---------------- same-cell assembly, reference order (src/pseudo_loop.cc:85-127) ----------------
    int16_t *w4 = q.t4;
#define PUT(tbl, val) ccj_put16(w4 + (int64_t)(tbl) * st4 + off0, (val))
    auto ccj_put16 = [](int16_t *dst, int mn) -> int {
        int v = CCJ_INTERN_INF;
        if (mn < CCJ_INF / 2) {
            if (mn >= CCJ_INTERN_INF) mn = CCJ_INTERN_INF;
            v = (int)(int16_t)mn;
        }
        *dst = (int16_t)v;
        return v;
    };
    const int II = CCJ_INTERN_INF;
    PUT(T_PLmloop00, min(II + bp, aPLm00));
    PUT(T_PLmloop01, aPLm01);
    PUT(T_PLmloop10, aPLm10);
    PUT(T_PRmloop00, min(II + bp, aPRm00));
    {
        // neighbours (i,j,k,l-1) and (i,j,k+1,l): valid iff b>=1
        int e01 = INF, e10 = INF, f01 = INF;
        if (b >= 1) {
            const int oA = OFF(a, b - 1, i, k);      // (i,j,k,l-1)
            const int oB = OFF(a, b - 1, i, k + 1);  // (i,j,k+1,l)
            e01 = ld16(TB(T_PRmloop01), oA) + cp;
            e10 = ld16(TB(T_PRmloop10), oB) + cp;
            f01 = ld16(TB(T_PMmloop01), oB) + cp;
        }
        PUT(T_PRmloop01, min(e01, aPRm01));
        PUT(T_PRmloop10, min(e10, aPRm10));
        PUT(T_PMmloop00, min(II + bp, aPMm00));
        PUT(T_PMmloop01, min(f01, aPMm01));
        int g10 = INF;
        if (a >= 1) g10 = ld16(TB(T_PMmloop10), OFF(a - 1, b, i, k)) + cp;  // (i,j-1,k,l)
        PUT(T_PMmloop10, min(g10, aPMm10));
    }
    PUT(T_POmloop00, min(II + bp, aPOm00));
    PUT(T_POmloop01, aPOm01);
    PUT(T_POmloop10, aPOm10);

    const int8_t *__restrict__ S = q.S;
    const int *__restrict__ estP = q.estP;
    auto ptype = [&](int x, int y) { return __ldg(&M->pair[S[x]][S[y]]); };

    // ---- PL (src/pseudo_loop.cc:232-253, get_PLiloop :682-703, get_PLmloop :705-715) ----
    int vPL;
    {
        int mn = INF;
        if (ptype(i, j) > 0) {
            if (a >= 2) {
                const int o = OFF(a - 2, b, i + 1, k);  // (i+1,j-1,k,l)
                if (a > CCJ_TURN) {  // can_pair(i,j)
                    if (a > CCJ_TURN + 2) mn = ld16(TB(T_PL), o) + __ldg(&estP[a * n1 + i]);
                    const int slot = ccj_tri(i, j);
                    const uint32_t *__restrict__ lst = q.inlist + (int64_t)slot * CCJ_WIN;
                    const int cnt = __ldg(&q.incnt[slot]);
                    const int16_t *__restrict__ pPL = TB(T_PL);
                    for (int e = 0; e < cnt; ++e) {
                        const uint32_t en = __ldg(&lst[e]);
                        const int x = (en >> 16) & 0xff, y = en >> 24;
                        const int mm = m + x + y, ii = i + x;
                        const int o2 = s_cb[b] - s_tet[mm] + (((ii - 1) * (2 * mm + 2 - ii)) >> 1) + (k - j + y - 2);
                        mn = min(mn, (int)(int16_t)(en & 0xffff) + ld16(pPL, o2));
                    }
                }
                mn = min(mn, min(ld16(TB(T_PLmloop10), o), ld16(TB(T_PLmloop01), o)) + apbp + bp);
                if (a >= CCJ_TURN + 1) mn = min(mn, ld16(TB(T_PfromL), o));
            }
            // a<2: get_PLmloop / PfromL read an invalid index -> INF
        }
        vPL = PUT(T_PL, mn);
    }
    // ---- PR (src/pseudo_loop.cc:255-275, get_PRiloop :717-738) ----
    int vPR;
    {
        int mn = INF;
        if (ptype(k, l) > 0) {
            if (b >= 2) {
                const int o = OFF(a, b - 2, i, k + 1);  // (i,j,k+1,l-1)
                if (b > CCJ_TURN) {
                    if (b > CCJ_TURN + 2) mn = ld16(TB(T_PR), o) + __ldg(&estP[b * n1 + k]);
                    const int slot = ccj_tri(k, l);
                    const uint32_t *__restrict__ lst = q.inlist + (int64_t)slot * CCJ_WIN;
                    const int cnt = __ldg(&q.incnt[slot]);
                    const int16_t *__restrict__ pPR = TB(T_PR);
                    const int rr = ((i - 1));
                    for (int e = 0; e < cnt; ++e) {
                        const uint32_t en = __ldg(&lst[e]);
                        const int x = (en >> 16) & 0xff, y = en >> 24;
                        const int mm = m + x + y;
                        const int o2 = s_cb[b - x - y] - s_tet[mm] + ((rr * (2 * mm + 2 - i)) >> 1) + (k + x - j - 2);
                        mn = min(mn, (int)(int16_t)(en & 0xffff) + ld16(pPR, o2));
                    }
                }
                mn = min(mn, min(ld16(TB(T_PRmloop10), o), ld16(TB(T_PRmloop01), o)) + apbp + bp);
                if (b >= CCJ_TURN + 1) mn = min(mn, ld16(TB(T_PfromR), o));
            }
        }
        vPR = PUT(T_PR, mn);
    }
    // ---- PM (src/pseudo_loop.cc:277-300, get_PMiloop :752-773) ----
    int vPM;
    {
        int mn = INF;
        if (ptype(j, k) > 0) {
            if (a >= 1 && b >= 1) {
                const int o = OFF(a - 1, b - 1, i, k + 1);  // (i,j-1,k+1,l)
                if (k - j > CCJ_TURN) {                     // can_pair(j,k)
                    mn = ld16(TB(T_PM), o) + __ldg(&estP[(k - j + 2) * n1 + (j - 1)]);
                    const int slot = ccj_tri(j, k);
                    const uint32_t *__restrict__ lst = q.outlist + (int64_t)slot * CCJ_WIN;
                    const int cnt = __ldg(&q.outcnt[slot]);
                    const int16_t *__restrict__ pPM = TB(T_PM);
                    const int rr = i - 1;
                    for (int e = 0; e < cnt; ++e) {
                        const uint32_t en = __ldg(&lst[e]);
                        const int x = (en >> 16) & 0xff, y = en >> 24;
                        if (x < a && y < b) {
                            const int mm = m + x + y;
                            const int o2 = s_cb[b - y] - s_tet[mm] + ((rr * (2 * mm + 2 - i)) >> 1) + (k + y - j + x - 2);
                            mn = min(mn, (int)(int16_t)(en & 0xffff) + ld16(pPM, o2));
                        }
                    }
                }
                mn = min(mn, min(ld16(TB(T_PMmloop10), o), ld16(TB(T_PMmloop01), o)) + apbp + bp);
                mn = min(mn, ld16(TB(T_PfromM), o));
            }
            if (a == 0 && b == 0) mn = min(mn, 0);
        }
        vPM = PUT(T_PM, mn);
    }
    // ---- PO (src/pseudo_loop.cc:302-322, get_POiloop :787-808: window dead) ----
    int vPO;
    {
        int mn = INF;
        if (ptype(i, l) > 0) {
            if (a >= 1 && b >= 1) {
                const int o = OFF(a - 1, b - 1, i + 1, k);  // (i+1,j,k,l-1)
                mn = ld16(TB(T_PO), o) + __ldg(&estP[(l - i) * n1 + i]);  // l-i>3 always here
                mn = min(mn, min(ld16(TB(T_POmloop10), o), ld16(TB(T_POmloop01), o)) + apbp + bp);
                mn = min(mn, ld16(TB(T_PfromO), o));
            }
        }
        vPO = PUT(T_PO, mn);
    }
    PUT(T_PfromL, min(min(aPfL1, aPfL2), min(vPR + PB, min(vPM + PB, vPO + PB))));
    PUT(T_PfromR, min(min(aPfR1, aPfR2), min(vPM + PB, vPO + PB)));
    PUT(T_PfromM, aPfM);
    PUT(T_PfromMprime, aPfMp + PB);
    PUT(T_PfromO, min(min(aPfO1, aPfO2), min(vPL + PB, vPR + PB)));
    PUT(T_PK, min(min(aPK1, aPK3), min(min(vPL + PB, vPM + PB), min(vPR + PB, vPO + PB))));
    w4[(int64_t)T_MPP * st4 + off0] = (int16_t)min(vPL, vPR);
#undef PUT
#undef OFF
#undef TB
}

void launch_prep(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, cudaStream_t st) {
    if (d.nmax < 2) return;
    k_prep<<<dim3(d.nmax - 1, d.nseq), 128, 0, st>>>(M, seqs);
}

bool fill4_tuned_supported(int nmax) { return nmax <= K4_MAXN; }

void launch_4d_tuned(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int t, cudaStream_t st) {
    const int m = d.nmax - t - 2;
    if (m < 1) return;
    const int ncell = m * (m + 1) / 2;
    k_4d_v2<<<dim3((ncell + K4_THREADS - 1) / K4_THREADS, t + 1, d.nseq), K4_THREADS, 0, st>>>(M, seqs, t);
}

}  // namespace ccj
