// Tuned level-wavefront kernels for the 22 gap tables, compute_P and the tuned traceback scan (sm_100a).
//
// A DP level is t=(j-i)+(l-k).  blockIdx.y -> a=j-i (b=t-a), threads walk the packed (i,k) triangle of slab (a,b),
// which is one contiguous run in every table (layout in ccj_types.h) -> stores of a warp are full lines.
//
// The reference evaluates, per cell, 22 recurrences one after the other; each is a min over split points d of
// X(neighbour cell) + 2D-term  (src/pseudo_loop.cc:181-644).  Candidates only read cells of LOWER levels, so their
// order is free.  They are regrouped by the 4 access patterns
//     L1: X(i,d,k,l)   L2: X(d,j,k,l)   R3: X(i,j,d,l)   R4: X(i,j,k,d)
// ("roles": one record load and one 16-byte {WB,WP,WBP} load serve 5-7 tables), evaluated for three levels per
// launch (k_roles), and the same-cell terms are applied afterwards in the reference's in-cell order (:85-127),
// which is what fixes the "unset = 32767" reads (k_final).  The interior-loop windows (get_P{L,R,M}iloop,
// :682-773) walk per-pair lists of pairable partners with pre-rounded energies instead of the 29x29
// can_pair-gated scan, over dedicated quad-aligned copies of PL/PR/PM with packed int16x2 arithmetic
// (k_winLR, k_winM).  Kernels, in launch order per level: k_roles (every third level), k_winLR, k_winM, k_final;
// per span: k_P_tuned, k_2d (ccj_kernels.cu); once per fill: k_prep_lay, k_fill_pmw, k_prep.
// Checked bit-for-bit against ccj_cell4d (generic version), the reference's tables and the CPU restatement.
// CCJ_HOST_EMU (tests/emu/ccj_emu_tuned.cpp only): g++ compiles these kernels for the SIMT emulator of tests/emu/simt_emu.hpp;
// the inline PTX gets plain C++ equivalents and the <<< >>> launchers are left out.  The CUDA build never defines it.
#include "ccj_kernels.cuh"
#include "ccj_cells4.cuh"

#include <algorithm>
#include <cstdlib>

namespace ccj {

#ifndef K4_THREADS
#define K4_THREADS 128
#endif
#define K4_MAXN 448  // int32 offsets and the shared layout tables
#define CCJ_WRAP_GUARD (-31000)   // see k_final: entries below this make the sequence take the generic kernels

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

// layout tables, PM row offsets and the per-arm lists of pairs; one block per sequence, before k_prep
__global__ void __launch_bounds__(256) k_prep_lay(const ccj_model *M, const ccj_seq *seqs) {
    __shared__ int s_row[K4_MAXN + 4];
    const ccj_seq q = seqs[blockIdx.x];
    const int n = q.n;
    if (n > K4_MAXN || n < 1) return;
    const int n1 = n + 1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int *lay = q.lay;
    for (int x = threadIdx.x; x <= n; x += blockDim.x) {
        lay[x] = (int)ccj_tet(x);
        lay[n1 + x] = x <= n - 3 ? (int)ccj_cb(n, x) : 0;
        lay[2 * n1 + x] = (int)ccj_h4(x);
    }
    // PM rows: quads of row (j,k) inside a level
    for (int j = threadIdx.x + 1; j <= n; j += blockDim.x) {
        int sum = 0;
        for (int k = j + 2; k <= n; ++k) sum += ccj_pmw_w4(n, j, k);
        s_row[j] = sum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int x = 0; x <= n; ++x) { acc += lay[2 * n1 + x]; lay[3 * n1 + x] = acc; }
        acc = 0;
        for (int b = 0; b <= n; ++b) { lay[4 * n1 + b] = acc; if (n - b - 2 >= 0) acc += lay[3 * n1 + n - b - 2]; }
        acc = 0;
        for (int j = 1; j <= n; ++j) { const int v = s_row[j]; s_row[j] = acc; acc += v; }
        // block offsets of the two PK copies of compute_P (ccj_types.h, "PK copies for compute_P")
        int *CF = lay + CCJ_LAY_CF(n), *DF = lay + CCJ_LAY_DF(n), *S2 = lay + CCJ_LAY_S2(n), *EG = lay + CCJ_LAY_EG(n);
        acc = 0;
        CF[0] = 0;
        for (int j = 1; j <= n; ++j) { CF[j] = acc; acc += 8 * (int)ccj_q8(n - j - 1); }
        CF[n + 1] = acc;
        acc = 0;
        DF[0] = 0;
        for (int i = 1; i <= n; ++i) { DF[i] = acc; acc += CF[n + 1] - CF[i]; }
        DF[n + 1] = acc;
        S2[0] = S2[1] = 0;
        for (int y = 1; y <= n; ++y) S2[y + 1] = S2[y] + 8 * (int)ccj_q8(y);
        acc = 0;
        EG[0] = 0;
        for (int i = 1; i <= n; ++i) { EG[i] = acc; acc += S2[n - i]; }
        EG[n + 1] = acc;
    }
    __syncthreads();
    for (int j = threadIdx.x + 1; j <= n; j += blockDim.x) {
        int acc = s_row[j];
        for (int k = j + 2; k <= n; ++k) { q.pmlev4[j * n1 + k] = acc; acc += ccj_pmw_w4(n, j, k); }
    }
    __syncthreads();
    // pairs (i,i+s) per arm length s
    const int8_t *S = q.S;
    for (int s = wid; s <= n; s += nw) {
        int *pl = q.plist + s * n1, *pc = q.pcum + s * (n + 2);
        int cnt = 0;
        if (lane == 0) pc[0] = 0;
        for (int i0 = 1; i0 <= n; i0 += 32) {
            const int i = i0 + lane;
            const bool ok = i + s <= n && s > CCJ_TURN && M->pair[S[i]][S[i + s]] > 0;
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            const int before = __popc(bal & ((1u << lane) - 1));
            if (ok) pl[cnt + before] = i;
            if (i <= n) pc[i] = cnt + before + (ok ? 1 : 0);
            cnt += __popc(bal);
        }
        if (lane == 0) { pc[n + 1] = cnt; s_row[s] = cnt; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int s = 0; s <= n; ++s) { q.pmstart[s] = acc; acc += s_row[s]; }
        q.pmstart[n + 1] = acc;
        q.status[6] = 1;   // the tuned fill is running: layout tables and the second PK copy will be valid for the traceback
    }
    __syncthreads();
    for (int s = wid; s <= n; s += nw) {
        const int base = q.pmstart[s], cnt = s_row[s];
        for (int x = lane; x < cnt; x += 32) {
            const int i = q.plist[s * n1 + x];
            q.pmlist[base + x] = i | ((i + s) << 16);
        }
    }
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
    const int n1 = n + 1;
    for (int i = 1 + wid; i < j; i += nw) {
        const int ij = ccj_idx2(n, i, j);
        if (lane == 0) c.q.estP[ij] = (j - i >= 2) ? ccj_e_stP(M, S, i, j) : CCJ_INF;
        const int slot = ccj_tri(i, j);
        int nin = 0, nout = 0, nneg = 0;
        if (ccj_can_pair(c, i, j)) {
            uint32_t *il = c.q.inlist + (int64_t)slot * CCJ_WIN_IN;
            uint2 *ol = reinterpret_cast<uint2 *>(c.q.outlist) + (int64_t)slot * CCJ_WIN_OUT;
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
                const unsigned bal = __ballot_sync(0xffffffffu, ok);
                if (ok) il[nin + __popc(bal & ((1u << lane) - 1))] = ent;
                nin += __popc(bal);
            }
            // outside (i,j) as inner pair (j,k):=(i,j): d=i-x, dp=j+y  (get_PMiloop window).  Two sweeps: the
            // candidates with a negative energy first (k_winM needs the mask halves of PMW only for those)
            for (int sweep = 0; sweep < 2; ++sweep) {
                for (int s0 = 0; s0 < CCJ_WIN; s0 += 32) {
                    const int s = s0 + lane;
                    const int x = s / 29 + 1, y = s % 29 + 1;
                    uint32_t ent = 0;
                    bool ok = false;
                    int pm4 = 0;
                    if (s < CCJ_WIN) {
                        const int d = i - x, dp = j + y;
                        if (d >= 1 && dp <= n && ccj_can_pair(c, d, dp)) {
                            const int e = ccj_e_intP(M, S, d, i, j, dp);
                            if ((e < 0) == (sweep == 0)) {
                                ok = pack_entry(e, x, y, ent, c.q.status);
                                if (c.q.pmlev4 && n <= K4_MAXN) pm4 = c.q.pmlev4[d * n1 + dp];
                            }
                        }
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, ok);
                    if (ok) ol[nout + __popc(bal & ((1u << lane) - 1))] = make_uint2(ent, (uint32_t)pm4);
                    nout += __popc(bal);
                }
                if (sweep == 0) nneg = nout;
            }
            // zero entries up to the next multiple of 8: the window kernels read whole 8-entry batches
            if (lane < 8) {
                if (nin + lane < ((nin + 7) & ~7)) il[nin + lane] = 0;
                if (nout + lane < ((nout + 7) & ~7)) ol[nout + lane] = make_uint2(0u, 0u);
            }
        }
        if (lane == 0) {
            c.q.incnt[slot] = nin;
            c.q.outcnt[slot] = nout | (nneg << 16);   // all candidates | those with a negative energy (listed first)
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Level kernels.  A cell's work is split over independent pieces -- 4 split-point "roles" (separate thread
// blocks of k_roles) and the interior windows (k_winLR, k_winM, one level ahead on their own stream) -- that
// leave int16-saturated partial minima in per-level scratch buffers; k_final assembles the 22 tables in the
// reference's in-cell order.  Splitting keeps a role at <= 9 accumulators per level, which is what lets one
// thread carry three levels at 7 blocks/SM -- a monolithic cell kernel was latency-bound (84 % long-scoreboard
// stalls, profiles/r1_notes.md).
// Saturating a partial at 32767 is exact: every later operation is +non-negative constant / min, and the
// final store clamps at 32767 anyway (Matrix4D::set, src/matrices.hh:188-191).
// ---------------------------------------------------------------------------------------------
enum {  // partial ids in the scratch
    Q_PK1 = 0, Q_PfL2, Q_PfM, Q_PLm00a, Q_PLm01, Q_PLm10a, Q_PMm00a,              // role L1
    Q_PfL1, Q_PfO1, Q_PLm00b, Q_PLm10b, Q_PMm10a, Q_POm00a, Q_POm10a,               // role L2
    Q_PK3, Q_PfR1, Q_PfMp, Q_PRm00a, Q_PRm10, Q_PMm00b,                              // role R3
    Q_PfR2, Q_PfO2, Q_PRm00b, Q_PRm01, Q_PMm01, Q_PMm10b, Q_POm00b, Q_POm01, Q_POm10b,  // role R4
    Q_COUNT  // the window partials live in ccj_seq::wscr, in the window layouts
};
enum { ROLE_L1 = 0, ROLE_L2, ROLE_R3, ROLE_R4 };

// streaming read of a gap-table entry: read-only path, and ask L2 to fetch the whole 256-byte chunk -- the
// neighbouring warps of the block need the adjacent 64-byte runs of the same slab row
__device__ __forceinline__ int ld16(const int16_t *__restrict__ p, int off) {
#ifdef CCJ_HOST_EMU
    return (int)p[off];
#else
    int v;
    asm("ld.global.nc.L2::256B.s16 %0, [%1];" : "=r"(v) : "l"(p + off));
    return v;
#endif
}
__device__ __forceinline__ int lo16(int w) { return (int)(int16_t)(w & 0xffff); }
#if defined(HI16_PRMT) && !defined(CCJ_HOST_EMU)   // experiment: sign-extend the high half with one PRMT that the three levels of k_roles share
__device__ __forceinline__ int hi16(int w) { int r; asm("prmt.b32 %0, %1, 0, 0xbb32;" : "=r"(r) : "r"(w)); return r; }
#else
__device__ __forceinline__ int hi16(int w) { return w >> 16; }
#endif
struct R12 { int w0, w1, w2; };  // one 12-byte read-group record (6 int16)
__device__ __forceinline__ R12 ldr12(const int *__restrict__ g, int off) {
    const int *p = g + (int64_t)off * 3;
    R12 r;
    r.w0 = __ldg(p); r.w1 = __ldg(p + 1); r.w2 = __ldg(p + 2);
    return r;
}
__device__ __forceinline__ int16_t sat16(int x) { return (int16_t)max(min(x, 32767), -32768); }

struct Cell {
    int n, m, a, b, i, j, k, l, p, c;  // p: index inside slab (a,b); c: index inside the level
};

// shared tables + decode of the thread's cell; returns false for threads past the slab
__device__ __forceinline__ bool cell_setup(const ccj_seq &q, int t, int a, int blk, int *s_tet, int *s_cb, Cell &C) {
    const int n = q.n;
    const int m = n - t - 2;
    for (int x = threadIdx.x; x <= n; x += K4_THREADS) {
        s_tet[x] = __ldg(&q.lay[x]);
        s_cb[x] = __ldg(&q.lay[n + 1 + x]);
    }
    __syncthreads();
    const int ncell = m * (m + 1) / 2;
    const int p = blk * K4_THREADS + threadIdx.x;
    if (p >= ncell) return false;
    // invert p = (i-1)(2m+2-i)/2 + kk  (rows i=1..m of length m+1-i)
    int r = (int)(((2 * m + 1) - sqrtf((float)((2 * m + 1) * (2 * m + 1) - 8 * p))) * 0.5f);
    if (r < 0) r = 0;
    if (r > m - 1) r = m - 1;
    while (r > 0 && r * (2 * m + 1 - r) / 2 > p) --r;
    while ((r + 1) * (2 * m - r) / 2 <= p) ++r;
    C.n = n; C.m = m; C.a = a; C.b = t - a;
    C.i = r + 1;
    const int kk = p - r * (2 * m + 1 - r) / 2;
    C.j = C.i + a; C.k = C.j + 2 + kk; C.l = C.k + C.b;
    C.p = p;
    C.c = a * ncell + p;
    return true;
}

#define TB(tbl) (t4 + (int64_t)(tbl) * st4)
#define OFF(aa, bb, ii, kx) (s_cb[bb] - s_tet[n - (aa) - (bb)-2] + ((((ii)-1) * (2 * (n - (aa) - (bb)-2) + 2 - (ii))) >> 1) + ((kx) - (ii) - (aa)-2))
#ifndef U4
#define U4 2
#endif

// ---- the four split-point roles: accumulators and the add-mins of ONE split point -------------------------------
// *_first is the boundary split point that only some recurrences have (d=i, d=j, d=l, d=k), *_term a regular one.
// r = read-group record of the source cell, w = {WB, WP, WBP} of the 2D interval next to it.
// L1: X(i,d,k,l) with (d+1,j)   [src/pseudo_loop.cc:184-187,357-361,399-402,449-458,468-471,481-487,548-551]
//     record g1 = PK PfromL | PfromMprime PLmloop00 | PLmloop10 PMmloop00
struct AccL1 { int PK, PfL, PfM, PLm00, PLm01, PLm10, PMm00; };
__device__ __forceinline__ void l1_first(AccL1 &A, const R12 &r, const int4 &w) {  // d=i
    const int x = hi16(r.w1);
    A.PLm00 = min(A.PLm00, x + w.x); A.PLm01 = min(A.PLm01, x + w.z); A.PMm00 = min(A.PMm00, hi16(r.w2) + w.x);
}
__device__ __forceinline__ void l1_term(AccL1 &A, const R12 &r, const int4 &w) {
    A.PK = min(A.PK, lo16(r.w0) + w.y); A.PfL = min(A.PfL, hi16(r.w0) + w.y); A.PfM = min(A.PfM, lo16(r.w1) + w.y);
    const int x = hi16(r.w1);
    A.PLm00 = min(A.PLm00, x + w.x); A.PLm01 = min(A.PLm01, x + w.z);
    A.PLm10 = min(A.PLm10, lo16(r.w2) + w.x); A.PMm00 = min(A.PMm00, hi16(r.w2) + w.x);
}
// L2: X(d,j,k,l) with (i,d-1)   [:357-359,425-428,450-453,481-483,581-584,599-602,632-635]
//     record g2 = PfromL PfromO | PLmloop00 PMmloop00 | POmloop00 -
struct AccL2 { int PfL, PfO, PLm00, PLm10, PMm10, POm00, POm10; };
__device__ __forceinline__ void l2_first(AccL2 &A, const R12 &r, const int4 &w) {  // d=j
    const int x = lo16(r.w1);
    A.PLm00 = min(A.PLm00, x + w.x); A.PLm10 = min(A.PLm10, x + w.z); A.PMm10 = min(A.PMm10, hi16(r.w1) + w.z);
    const int y = lo16(r.w2);
    A.POm00 = min(A.POm00, y + w.x); A.POm10 = min(A.POm10, y + w.z);
}
__device__ __forceinline__ void l2_term(AccL2 &A, const R12 &r, const int4 &w) {
    A.PfL = min(A.PfL, lo16(r.w0) + w.y); A.PfO = min(A.PfO, hi16(r.w0) + w.y);
    l2_first(A, r, w);
}
// R3: X(i,j,d,l) with (k,d-1)   [:189-192,379-381,412-415,499-503,534-537,552-555]
//     record g3 = PK PfromR | min(PL,PR) PRmloop00 | PMmloop00 -
struct AccR3 { int PK, PfR, PfMp, PRm00, PRm10, PMm00; };
__device__ __forceinline__ void r3_first(AccR3 &A, const R12 &r, const int4 &w) {  // d=l
    const int x = hi16(r.w1);
    A.PRm00 = min(A.PRm00, x + w.x); A.PRm10 = min(A.PRm10, x + w.z); A.PMm00 = min(A.PMm00, lo16(r.w2) + w.x);
}
__device__ __forceinline__ void r3_term(AccR3 &A, const R12 &r, const int4 &w) {
    A.PK = min(A.PK, lo16(r.w0) + w.y); A.PfR = min(A.PfR, hi16(r.w0) + w.y); A.PfMp = min(A.PfMp, lo16(r.w1) + w.y);
    r3_first(A, r, w);
}
// R4: X(i,j,k,d) with (d+1,l)   [:382-383,429-432,504-507,520-523,567-570,585-588,603-606,618-621,636-639]
//     record g4 = PfromR PfromO | PRmloop00 PMmloop00 | PMmloop10 POmloop00 | POmloop10 -
struct AccR4 { int PfR, PfO, PRm00, PRm01, PMm01, PMm10, POm00, POm01, POm10; };
__device__ __forceinline__ void r4_first(AccR4 &A, const int4 &r, const int4 &w) {  // d=k
    const int x = lo16(r.y);
    A.PRm00 = min(A.PRm00, x + w.x); A.PRm01 = min(A.PRm01, x + w.z); A.PMm01 = min(A.PMm01, hi16(r.y) + w.z);
    const int y = hi16(r.z);
    A.POm00 = min(A.POm00, y + w.x); A.POm01 = min(A.POm01, y + w.z);
}
__device__ __forceinline__ void r4_term(AccR4 &A, const int4 &r, const int4 &w) {
    A.PfR = min(A.PfR, lo16(r.x) + w.y); A.PfO = min(A.PfO, hi16(r.x) + w.y);
    A.PMm10 = min(A.PMm10, lo16(r.z) + w.x); A.POm10 = min(A.POm10, lo16(r.w) + w.x);
    r4_first(A, r, w);
}

// Split-point roles of KF consecutive levels per launch.  A record X(i,d,k,l) that role L1 reads for the cell
// (i,j,k,l) of level t is the very record the cells (i,j+1,k,l), (i,j+2,k,l) of levels t+1, t+2 need at the same
// split point, only with the 2D interval (d+1,j+f) instead of (d+1,j); likewise (i-f,j,k,l) for L2, (i,j,k-f,l)
// for R3 and (i,j,k,l+f) for R4.  Launched for t = 0 mod KF, a thread therefore keeps KF sets of accumulators:
// its own cell (all split points: sources of levels <= t-1) and its partners on levels t+1..t+KF-1 (the same split
// points; their remaining <= f split points, whose sources are cells of levels t..t+f-1, are added by k_final).
// Every record is fetched from HBM once per KF levels.
#ifndef KF
#define KF 3
#endif
#ifndef ROLES_MINB
#define ROLES_MINB 7
#endif
__global__ void __launch_bounds__(K4_THREADS, ROLES_MINB) k_roles(const ccj_model *__restrict__ M, const ccj_seq *__restrict__ seqs, int t) {
    __shared__ int s_tet[K4_MAXN + 4];
    __shared__ int s_cb[K4_MAXN + 4];
    const int role = (int)(blockIdx.z % 4);
    const ccj_seq q = seqs[blockIdx.z / 4];
    const int n = q.n;
    if (n - t - 2 < 1) return;
    {
        const int m0 = n - t - 2;
        if ((int)(blockIdx.x * K4_THREADS) >= m0 * (m0 + 1) / 2) return;
    }
    Cell C;
    if (!cell_setup(q, t, blockIdx.y, blockIdx.x, s_tet, s_cb, C)) return;
    const int a = C.a, b = C.b, i = C.i, j = C.j, k = C.k, l = C.l, m = C.m;
    const int4 *__restrict__ W3 = reinterpret_cast<const int4 *>(q.w3);
    const int n1 = n + 1;
    const int INF = CCJ_INF;
    const int64_t ss = q.scratch_stride;
    const int kk = k - j - 2;
    // scratch slot of the partner on level t+f: plane (t+f) mod KF, slab aa, row ii, position pp of a triangle with m-f rows
    auto slot = [&](int f, int aa, int ii, int pp) -> int16_t * {
        const int mf = m - f;
        return q.scratch + (int64_t)((t + f) % KF) * Q_COUNT * ss + (int64_t)aa * (mf * (mf + 1) / 2) + (((ii - 1) * (2 * mf + 2 - ii)) >> 1) + pp;
    };
#define SAVE(id, v) sc[(int64_t)(id) * ss] = sat16(v)

    if (role == ROLE_L1) {
        AccL1 A[KF];
#pragma unroll
        for (int f = 0; f < KF; ++f) A[f] = {INF, INF, INF, INF, INF, INF, INF};
        if (a >= 1) {
            const int *__restrict__ G = reinterpret_cast<const int *>(q.g1);
            {  // d=i, 2D interval (i+1, j+f)
                const R12 r = ldr12(G, OFF(0, b, i, k));
#pragma unroll
                for (int f = 0; f < KF; ++f) l1_first(A[f], r, __ldg(&W3[(a - 1 + f) * n1 + i + 1]));
            }
            const int ub = n - b - 2, cbb = s_cb[b], ri = i - 1, kc = k - i - 2;
            int ap = 1;
            for (; ap + U4 <= a; ap += U4) {
                R12 r[U4];
                int4 w[KF][U4];
#pragma unroll
                for (int u = 0; u < U4; ++u) {
                    const int m1 = ub - ap - u;
                    r[u] = ldr12(G, cbb - s_tet[m1] + ((ri * (2 * m1 + 2 - i)) >> 1) + (kc - ap - u));
#pragma unroll
                    for (int f = 0; f < KF; ++f) w[f][u] = __ldg(&W3[(a - ap - u - 1 + f) * n1 + i + ap + u + 1]);   // (d+1, j+f)
                }
#pragma unroll
                for (int u = 0; u < U4; ++u)
#pragma unroll
                    for (int f = 0; f < KF; ++f) l1_term(A[f], r[u], w[f][u]);
            }
            for (; ap < a; ++ap) {
                const int m1 = ub - ap;
                const R12 r = ldr12(G, cbb - s_tet[m1] + ((ri * (2 * m1 + 2 - i)) >> 1) + (kc - ap));
#pragma unroll
                for (int f = 0; f < KF; ++f) l1_term(A[f], r, __ldg(&W3[(a - ap - 1 + f) * n1 + i + ap + 1]));
            }
        }
#pragma unroll
        for (int f = 0; f < KF; ++f)
            if (m - f >= 1 && kk >= f) {   // partner (i,j+f,k,l)
                int16_t *sc = slot(f, a + f, i, kk - f);
                SAVE(Q_PK1, A[f].PK); SAVE(Q_PfL2, A[f].PfL); SAVE(Q_PfM, A[f].PfM); SAVE(Q_PLm00a, A[f].PLm00);
                SAVE(Q_PLm01, A[f].PLm01); SAVE(Q_PLm10a, A[f].PLm10); SAVE(Q_PMm00a, A[f].PMm00);
            }
    } else if (role == ROLE_L2) {
        AccL2 A[KF];
#pragma unroll
        for (int f = 0; f < KF; ++f) A[f] = {INF, INF, INF, INF, INF, INF, INF};
        if (a >= 1) {
            const int *__restrict__ G = reinterpret_cast<const int *>(q.g2);
            {  // d=j, 2D interval (i-f, j-1)
                const R12 r = ldr12(G, OFF(0, b, j, k));
#pragma unroll
                for (int f = 0; f < KF; ++f) l2_first(A[f], r, __ldg(&W3[(a - 1 + f) * n1 + i - f]));
            }
            const int ub = n - b - 2, cbb = s_cb[b], kc = k - j - 2;
            int ap = 1;
            for (; ap + U4 <= a; ap += U4) {
                R12 r[U4];
                int4 w[KF][U4];
#pragma unroll
                for (int u = 0; u < U4; ++u) {
                    const int d = i + ap + u, mm = ub - (a - ap - u);
                    r[u] = ldr12(G, cbb - s_tet[mm] + (((d - 1) * (2 * mm + 2 - d)) >> 1) + kc);
#pragma unroll
                    for (int f = 0; f < KF; ++f) w[f][u] = __ldg(&W3[(ap + u - 1 + f) * n1 + i - f]);   // (i-f, d-1)
                }
#pragma unroll
                for (int u = 0; u < U4; ++u)
#pragma unroll
                    for (int f = 0; f < KF; ++f) l2_term(A[f], r[u], w[f][u]);
            }
            for (; ap < a; ++ap) {
                const int d = i + ap, mm = ub - (a - ap);
                const R12 r = ldr12(G, cbb - s_tet[mm] + (((d - 1) * (2 * mm + 2 - d)) >> 1) + kc);
#pragma unroll
                for (int f = 0; f < KF; ++f) l2_term(A[f], r, __ldg(&W3[(ap - 1 + f) * n1 + i - f]));
            }
        }
#pragma unroll
        for (int f = 0; f < KF; ++f)
            if (m - f >= 1 && i - f >= 1) {   // partner (i-f,j,k,l)
                int16_t *sc = slot(f, a + f, i - f, kk);
                SAVE(Q_PfL1, A[f].PfL); SAVE(Q_PfO1, A[f].PfO); SAVE(Q_PLm00b, A[f].PLm00); SAVE(Q_PLm10b, A[f].PLm10);
                SAVE(Q_PMm10a, A[f].PMm10); SAVE(Q_POm00a, A[f].POm00); SAVE(Q_POm10a, A[f].POm10);
            }
    } else if (role == ROLE_R3) {
        AccR3 A[KF];
#pragma unroll
        for (int f = 0; f < KF; ++f) A[f] = {INF, INF, INF, INF, INF, INF};
        if (b >= 1) {
            const int *__restrict__ G = reinterpret_cast<const int *>(q.g3);
            {  // d=l, 2D interval (k-f, l-1)
                const R12 r = ldr12(G, OFF(a, 0, i, l));
#pragma unroll
                for (int f = 0; f < KF; ++f) r3_first(A[f], r, __ldg(&W3[(b - 1 + f) * n1 + k - f]));
            }
            const int ua = n - a - 2, ri = i - 1, kc = k - j - 2;
            int bq = 1;
            for (; bq + U4 <= b; bq += U4) {
                R12 r[U4];
                int4 w[KF][U4];
#pragma unroll
                for (int u = 0; u < U4; ++u) {
                    const int b3 = b - bq - u, m3 = ua - b3;
                    r[u] = ldr12(G, s_cb[b3] - s_tet[m3] + ((ri * (2 * m3 + 2 - i)) >> 1) + (kc + bq + u));
#pragma unroll
                    for (int f = 0; f < KF; ++f) w[f][u] = __ldg(&W3[(bq + u - 1 + f) * n1 + k - f]);   // (k-f, d-1)
                }
#pragma unroll
                for (int u = 0; u < U4; ++u)
#pragma unroll
                    for (int f = 0; f < KF; ++f) r3_term(A[f], r[u], w[f][u]);
            }
            for (; bq < b; ++bq) {
                const int b3 = b - bq, m3 = ua - b3;
                const R12 r = ldr12(G, s_cb[b3] - s_tet[m3] + ((ri * (2 * m3 + 2 - i)) >> 1) + (kc + bq));
#pragma unroll
                for (int f = 0; f < KF; ++f) r3_term(A[f], r, __ldg(&W3[(bq - 1 + f) * n1 + k - f]));
            }
        }
#pragma unroll
        for (int f = 0; f < KF; ++f)
            if (m - f >= 1 && kk >= f) {   // partner (i,j,k-f,l)
                int16_t *sc = slot(f, a, i, kk - f);
                SAVE(Q_PK3, A[f].PK); SAVE(Q_PfR1, A[f].PfR); SAVE(Q_PfMp, A[f].PfMp); SAVE(Q_PRm00a, A[f].PRm00);
                SAVE(Q_PRm10, A[f].PRm10); SAVE(Q_PMm00b, A[f].PMm00);
            }
    } else {
        AccR4 A[KF];
#pragma unroll
        for (int f = 0; f < KF; ++f) A[f] = {INF, INF, INF, INF, INF, INF, INF, INF, INF};
        if (b >= 1) {
            const int4 *__restrict__ G = reinterpret_cast<const int4 *>(q.g4);
            {  // d=k, 2D interval (k+1, l+f)
                const int4 r = __ldg(&G[OFF(a, 0, i, k)]);
#pragma unroll
                for (int f = 0; f < KF; ++f) r4_first(A[f], r, __ldg(&W3[(b - 1 + f) * n1 + k + 1]));
            }
            const int ua = n - a - 2, ri = i - 1, kc = k - j - 2;
            int bq = 1;
            for (; bq + U4 <= b; bq += U4) {
                int4 r[U4], w[KF][U4];
#pragma unroll
                for (int u = 0; u < U4; ++u) {
                    const int b4 = bq + u, m4 = ua - b4;
                    r[u] = __ldg(&G[s_cb[b4] - s_tet[m4] + ((ri * (2 * m4 + 2 - i)) >> 1) + kc]);
#pragma unroll
                    for (int f = 0; f < KF; ++f) w[f][u] = __ldg(&W3[(b - b4 - 1 + f) * n1 + k + b4 + 1]);   // (d+1, l+f)
                }
#pragma unroll
                for (int u = 0; u < U4; ++u)
#pragma unroll
                    for (int f = 0; f < KF; ++f) r4_term(A[f], r[u], w[f][u]);
            }
            for (; bq < b; ++bq) {
                const int m4 = ua - bq;
                const int4 r = __ldg(&G[s_cb[bq] - s_tet[m4] + ((ri * (2 * m4 + 2 - i)) >> 1) + kc]);
#pragma unroll
                for (int f = 0; f < KF; ++f) r4_term(A[f], r, __ldg(&W3[(b - bq - 1 + f) * n1 + k + bq + 1]));
            }
        }
#pragma unroll
        for (int f = 0; f < KF; ++f)
            if (m - f >= 1 && l + f <= n) {   // partner (i,j,k,l+f)
                int16_t *sc = slot(f, a, i, kk);
                SAVE(Q_PfR2, A[f].PfR); SAVE(Q_PfO2, A[f].PfO); SAVE(Q_PRm00b, A[f].PRm00); SAVE(Q_PRm01, A[f].PRm01);
                SAVE(Q_PMm01, A[f].PMm01); SAVE(Q_PMm10b, A[f].PMm10); SAVE(Q_POm00b, A[f].POm00); SAVE(Q_POm01, A[f].POm01);
                SAVE(Q_POm10b, A[f].POm10);
            }
    }
#undef SAVE
}

#define WB 8      // window candidates in flight per lane (FENCE8X assumes 8)
// keep the compiler from sinking the WB loads of a batch to their uses: all of them must be in flight together
#ifdef CCJ_HOST_EMU
#define FENCE8X(v) ((void)0)
#define CCJ_PIN64(p) ((void)0)
#else
#define FENCE8X(v) asm volatile("" : "+r"(v[0].x), "+r"(v[1].x), "+r"(v[2].x), "+r"(v[3].x), "+r"(v[4].x), "+r"(v[5].x), "+r"(v[6].x), "+r"(v[7].x))
#define CCJ_PIN64(p) asm volatile("" : "+l"(p))
#endif
#define WGRP 8    // lanes per run: one pass covers 8 quads = 32 cells
#define WRUNS (K4_THREADS / WGRP)

__device__ __forceinline__ int2 ldq(const int2 *p, int off) {
#ifdef CCJ_HOST_EMU
    return p[off];
#else
    int2 v;
    asm("ld.global.nc.L2::256B.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p + off));
    return v;
#endif
}
__device__ __forceinline__ int pack_sat(int lo, int hi) {
    return (int)((uint32_t)(uint16_t)sat16(lo) | ((uint32_t)(uint16_t)sat16(hi) << 16));
}

// packed int16x2 arithmetic (VIADDMNMX.S16x2 / VIMNMX.S16x2): two cells per instruction, no unpacking
#ifdef CCJ_HOST_EMU   // add.s16x2 wraps, min/max.s16x2 are signed, per 16-bit half
__device__ __forceinline__ int emu_half(int w, int h) { return (int)(int16_t)((unsigned)w >> (16 * h)); }
__device__ __forceinline__ int emu_pack(int lo, int hi) { return (int)(((unsigned)lo & 0xffffu) | ((unsigned)hi << 16)); }
__device__ __forceinline__ int emu_add2(int a, int b) {
    return emu_pack((int)(int16_t)(emu_half(a, 0) + emu_half(b, 0)), (int)(int16_t)(emu_half(a, 1) + emu_half(b, 1)));
}
__device__ __forceinline__ int min2(int a, int b) {
    return emu_pack(std::min(emu_half(a, 0), emu_half(b, 0)), std::min(emu_half(a, 1), emu_half(b, 1)));
}
__device__ __forceinline__ int emu_max2(int a, int b) {
    return emu_pack(std::max(emu_half(a, 0), emu_half(b, 0)), std::max(emu_half(a, 1), emu_half(b, 1)));
}
__device__ __forceinline__ int addmin2(int a, int b, int c) { return min2(emu_add2(a, b), c); }
__device__ __forceinline__ int addmax2(int a, int b, int c) { return emu_max2(emu_add2(a, b), c); }
#else
__device__ __forceinline__ int addmin2(int a, int b, int c) {  // min(a+b, c) per half
    int r, s;
    asm("add.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    asm("min.s16x2 %0, %1, %2;" : "=r"(s) : "r"(r), "r"(c));
    return s;
}
__device__ __forceinline__ int addmax2(int a, int b, int c) {  // max(a+b, c) per half
    int r, s;
    asm("add.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    asm("max.s16x2 %0, %1, %2;" : "=r"(s) : "r"(r), "r"(c));
    return s;
}
__device__ __forceinline__ int min2(int a, int b) {
    int s;
    asm("min.s16x2 %0, %1, %2;" : "=r"(s) : "r"(a), "r"(b));
    return s;
}
#endif
__device__ __forceinline__ int splat16(int e) { return (int)__byte_perm((unsigned)e, 0u, 0x1010); }
// A window candidate is value + energy with the sum saturated at 32767 ("INF or clamped", see k_final).  In
// 16 bits:  min(value, 32767 - max(energy,0)) + energy  -- exact, and it cannot wrap upwards.
#define WIN_INF2 0x7fff7fff

// PL and PR interior windows (get_PLiloop :682-703, get_PRiloop :717-738).
// Cells that share the closing pair share its partner list: for PL the cells of one row of the slab (i,j
// fixed, k running), for PR the cells of one transposed row (k,l fixed, i running).  Both are runs of the
// window layouts PLW / PRW (ccj_types.h), where the run of every candidate starts on the same quad phase:
// a group of 8 lanes takes one run of a PAIRED row (per-arm pair lists from k_prep_lay, so no lane idles on
// an unpairable row) and 32 of its cells.  The 8 lanes decode 8 list entries at a time into (source run,
// energy, clamp) in shared memory; every candidate is then one 8-byte load and four packed int16x2
// instructions per lane (4 cells).  Results go to the wscr partial in the same layout, 8 bytes per lane.
//   blockIdx.x -> (pass, chunk of 16 paired rows), blockIdx.y -> a, blockIdx.z -> (sequence, PL|PR)
// G = lanes per run (8: one pass covers 32 cells of a run, four runs per warp; 16: 64 cells, two runs per warp -- half
// the distinct cache lines per warp request, for levels whose runs are long)
template <bool PIPE, int G>
__global__ void __launch_bounds__(K4_THREADS, PIPE ? 6 : 12) k_winLR(const ccj_model *__restrict__ M, const ccj_seq *__restrict__ seqs, int t, int nchunk) {
    __shared__ int s_T[64];              // quad offset of source slab (a-s,b) resp. (a,b-s), plus H4(m+s)
    __shared__ int s_h4[K4_MAXN + 4];
    __shared__ int2 tile[2][K4_THREADS / G][G];   // double-buffered: (source run start, energy | clamp << 16)
    const int role = blockIdx.z & 1;     // 0: PL   1: PR
    const ccj_seq q = seqs[blockIdx.z >> 1];
    const int n = q.n, n1 = n + 1;
    const int m = n - t - 2;
    if (m < 1) return;
    const int a = blockIdx.y, b = t - a;
    const int arm = role == 0 ? a : b;   // length of the arm whose closing pair carries the window
    if (arm <= CCJ_TURN) return;
    const int chunk = blockIdx.x % nchunk, pass = blockIdx.x / nchunk;
    if (pass * (4 * G) >= m) return;
    const int *__restrict__ pc = q.pcum + arm * (n + 2);
    // paired 5' ends of this slab: PL i in 1..m, PR k in a+3..n-b
    const int first = role == 0 ? 0 : __ldg(&pc[a + 2]);
    const int last = role == 0 ? __ldg(&pc[m]) : __ldg(&pc[n - b]);
    if (first + chunk * (K4_THREADS / G) >= last) return;
    const int *__restrict__ lay = q.lay;
    for (int x = threadIdx.x; x <= n; x += K4_THREADS) s_h4[x] = __ldg(&lay[2 * n1 + x]);
    if (threadIdx.x < 64) {
        const int s = threadIdx.x;
        int v = 0;
        if (m + s <= n - 2) {
            const int bb = role == 0 ? b : b - s;
            if (bb >= 0 && n - bb - 2 >= m + s)
                v = __ldg(&lay[4 * n1 + bb]) + __ldg(&lay[3 * n1 + n - bb - 2]) - __ldg(&lay[3 * n1 + m + s]) + __ldg(&lay[2 * n1 + m + s]);
        }
        s_T[s] = v;
    }
    __syncthreads();
    const int grp = threadIdx.x / G, gl = threadIdx.x & (G - 1);
    const int idx = first + chunk * (K4_THREADS / G) + grp;
    const int p5 = __ldg(&q.plist[arm * n1 + min(idx, last - 1)]);
    const int c = role == 0 ? p5 - 1 : n - b - p5;     // row of the run: i-1 resp. kr=n-b-k
    const int zc = m - c;                               // its length
    const bool gact = idx < last && pass * (4 * G) < zc;   // the group has cells in this pass
    const int qd = pass * G + gl;                    // this lane's quad of the run
    const int2 *src = reinterpret_cast<const int2 *>(role == 0 ? q.plw : q.prw) + qd;
    CCJ_PIN64(src);  // per-lane base in a register pair: one IMAD.WIDE per candidate
    const int own = s_T[0] - s_h4[zc];                  // the run's own (not yet written) quads
    int acc0 = WIN_INF2, acc1 = WIN_INF2;
    if (gact && arm > CCJ_TURN + 2) {  // stacking term, x=y=1 (PL(i+1,j-1,k,l) / PR(i,j,k+1,l-1))
        const int2 w = ldq(src, s_T[2] - s_h4[zc + 1]);
        const int e = __ldg(&q.estP[arm * n1 + p5]);
        const int ee = splat16(e), cc = splat16(32767 - max(e, 0));
        acc0 = addmin2(min2(w.x, cc), ee, acc0);
        acc1 = addmin2(min2(w.y, cc), ee, acc1);
    }
    const int keep0 = acc0, keep1 = acc1;
    const int slot = ccj_tri(p5, p5 + arm);
    const uint32_t *__restrict__ lst = q.inlist + (int64_t)slot * CCJ_WIN_IN;
    const int cnt = gact ? __ldg(&q.incnt[slot]) : 0;
    const int nb = (__reduce_max_sync(0xffffffffu, cnt) + G - 1) / G;   // batches of 8 candidates, warp-uniform
    // PIPE (small waves): software pipeline over the batches -- while batch b is consumed, the loads of batch b+1
    // are in flight and the list entry of batch b+2 is on its way.
    // lane gl fetches entry 8b+gl of its group's list; past the end the last entry again (min is idempotent)
    auto fetch = [&](int bb) -> uint32_t { return cnt > 0 ? __ldg(&lst[min(bb * G + gl, cnt - 1)]) : 0u; };
    auto decode = [&](uint32_t en) -> int2 {
        if (cnt == 0) return make_int2(own, 0x7fff0000);
        const int x = (en >> 16) & 0xff, y = en >> 24, e = (int)(int16_t)(en & 0xffff);
        // PL source (i+x, j-y, k, l): slab (a-s, b), row c+x, length zc+y, same position (n-b)-k
        // PR source (i, j, k+x, l-y): slab (a, b-s), row kr+y, length zc+x, same position i-1
        return make_int2(s_T[x + y] - s_h4[zc + (role == 0 ? y : x)], (e & 0xffff) | ((32767 - max(e, 0)) << 16));
    };
    // one LDS.64 per candidate (shared memory / L1 is the busiest unit of this kernel): the energy and its clamp
    // travel as two halves of one word and are splatted with two PRMTs
#define ISSUE(W_, EC_, buf)  ISSUEH(W_, EC_, buf, 0)
#define ISSUEH(W_, EC_, buf, h0)                                              \
    _Pragma("unroll") for (int u = 0; u < WB; ++u) {                         \
        const int2 d2 = tile[buf][grp][(h0) + u];                            \
        W_[u] = ldq(src, d2.x);                                              \
        EC_[u] = d2.y;                                                       \
    }
#define CONSUME(W_, EC_)                                                      \
    _Pragma("unroll") for (int u = 0; u < WB; ++u) {                         \
        const int ee = (int)__byte_perm((unsigned)EC_[u], 0u, 0x1010), cc = (int)__byte_perm((unsigned)EC_[u], 0u, 0x3232); \
        acc0 = addmin2(min2(W_[u].x, cc), ee, acc0);                          \
        acc1 = addmin2(min2(W_[u].y, cc), ee, acc1);                          \
    }
    if (PIPE) {
        int2 wa[WB], wb[WB];
        int ea[WB], eb[WB];
        uint32_t pre = 0;
        if (nb > 0) {
            tile[0][grp][gl] = decode(fetch(0));
            pre = fetch(1);
            __syncwarp();
            ISSUE(wa, ea, 0);
        }
        for (int bb = 0; bb < nb; bb += 2) {
            if (bb + 1 < nb) {
                tile[1][grp][gl] = decode(pre);
                pre = fetch(bb + 2);
                __syncwarp();
                ISSUE(wb, eb, 1);
            }
            CONSUME(wa, ea);
            __syncwarp();
            if (bb + 1 < nb) {
                if (bb + 2 < nb) {
                    tile[0][grp][gl] = decode(pre);
                    pre = fetch(bb + 3);
                    __syncwarp();
                    ISSUE(wa, ea, 0);
                }
                CONSUME(wb, eb);
                __syncwarp();
            }
        }
    } else {  // waves with many sequences hide the latency with occupancy: plain loop, fewer registers
        uint32_t pre = nb > 0 ? fetch(0) : 0u;
        for (int bb = 0; bb < nb; ++bb) {
            const int2 d = decode(pre);
            pre = fetch(bb + 1);   // the next batch's list entry is on its way while this batch is loaded and consumed
            __syncwarp();
            tile[0][grp][gl] = d;
            __syncwarp();
#pragma unroll
            for (int h = 0; h < G; h += WB) {   // WB candidates in flight at a time
                int2 w[WB];
                int ec[WB];
                ISSUEH(w, ec, 0, h);
                FENCE8X(w);
                CONSUME(w, ec);
            }
        }
    }
#undef ISSUE
#undef ISSUEH
#undef CONSUME
    if (gact && 4 * qd < zc) {
        if (cnt == 0) { acc0 = keep0; acc1 = keep1; }
        int2 *__restrict__ out = reinterpret_cast<int2 *>(q.wscr + (int64_t)((t & 1) * 2 + role) * q.wscr_lr);
        out[a * s_h4[m] + (s_h4[m] - s_h4[zc]) + qd] = make_int2(acc0, acc1);
    }
}

__device__ __forceinline__ int4 ldq4(const int4 *p, int off) {
#ifdef CCJ_HOST_EMU
    return p[off];
#else
    int4 v;
    asm("ld.global.nc.L2::256B.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + off));
    return v;
#endif
}

// every entry of PMW / PMM starts as 32767 ("not a source"): rows are padded and read past their ends
__global__ void __launch_bounds__(256) k_fill_pmw(const ccj_seq *seqs) {
    const ccj_seq q = seqs[blockIdx.y];
    if (q.n > K4_MAXN || q.n < 3) return;
    const int64_t pairs = ((int64_t)q.wtot4 * (q.n - 2) + 16) / 2;   // two 8-byte quads per int4
    int4 *p = reinterpret_cast<int4 *>(q.pmw), *pm = reinterpret_cast<int4 *>(q.pmm);
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < pairs; x += (int64_t)gridDim.x * blockDim.x) {
        p[x] = make_int4(WIN_INF2, WIN_INF2, WIN_INF2, WIN_INF2);
        pm[x] = make_int4(WIN_INF2, WIN_INF2, WIN_INF2, WIN_INF2);
    }
}

// PM interior window (get_PMiloop, src/pseudo_loop.cc:752-773).  The partner list belongs to the INNER pair
// (j,k); the cells of one level that share it are (j-a, j, k, k+t-a), one row of the PMW layout, and a
// candidate (j-x, k+y) keeps i and l, i.e. the position inside the row.  A group of 8 lanes takes 32 cells of
// the row of one PAIRED (j,k) (pairs ordered by span: the rows that exist on level t are a prefix).
// A candidate is valid for j-t+y < i < j-x only (d>i, dp<l in the reference's loops), i.e. for the cells of
// the source row with a'>=1 and b'>=1: PMW stores next to every value a mask half (-32768 for such a cell,
// 32767 for the two end cells and the padding), and the candidate is max(value+energy, mask) -- still two
// cells per instruction.  The 8 lanes of a group decode 8 list entries at a time into shared memory.
//   blockIdx.x -> (pass, chunk of 16 pairs), blockIdx.y -> sequence
template <bool PIPE>
__global__ void __launch_bounds__(K4_THREADS, PIPE ? 4 : 10) k_winM(const ccj_model *__restrict__ M, const ccj_seq *__restrict__ seqs, int t, int nchunk) {
    __shared__ int4 tile[2][WRUNS][WGRP];   // (source row start, energy | clamp << 16, first quad, quads - 1)
    const ccj_seq q = seqs[blockIdx.y];
    const int n = q.n, n1 = n + 1;
    const int m = n - t - 2;
    if (m < 1) return;
    const int chunk = blockIdx.x % nchunk, pass = blockIdx.x / nchunk;
    const int nact = __ldg(&q.pmstart[n - t]);  // pairs with span <= n-1-t: (j-1)+(n-k) >= t
    if (chunk * WRUNS >= nact) return;
    const int grp = threadIdx.x / WGRP, gl = threadIdx.x & (WGRP - 1);
    const int g = min(chunk * WRUNS + grp, nact - 1);
    const int jk = __ldg(&q.pmlist[g]);
    const int j = jk & 0xffff, k = jk >> 16;
    const int B = n - k;
    const int ilo = max(j - t, 1), ihi = j - max(0, t - B);   // i of the row's cells
    const int qlo = (ilo - 1) >> 2, qhi = (ihi - 1) >> 2;
    const int qd = qlo + pass * WGRP + gl;                     // absolute quad of (i-1)
    const bool gact = chunk * WRUNS + grp < nact && qlo + pass * WGRP <= qhi;  // the group has cells in this pass
    const int p0 = 4 * qd;                                     // i-1 of the lane's first cell
    const int wtot4 = q.wtot4;
    const int2 *srcv = reinterpret_cast<const int2 *>(q.pmw) + qd;   // values
    const int2 *srcm = reinterpret_cast<const int2 *>(q.pmm) + qd;   // mask halves, same index
    CCJ_PIN64(srcv);
    const int INF = CCJ_INF;
    const int own = t * wtot4 + __ldg(&q.pmlev4[j * n1 + k]) - qlo;   // the row's own (not yet written) quads
    int acc0 = WIN_INF2, acc1 = WIN_INF2;
    if (gact && t >= 2 && j >= 2 && k < n) {
        // PM(i,j-1,k+1,l) + e_stP(j-1,k+1) for a>=1, b>=1.  Unlike the list candidates this one may read the end cells
        // of its source row (a'=0, b'=0), which PMW blanks out: read the main PM table, once per lane, in 32 bits
        const int e = __ldg(&q.estP[(k - j + 2) * n1 + (j - 1)]);
        const int16_t *__restrict__ pPM = q.t4 + (int64_t)T_PM * q.stride4;
        const int *__restrict__ lay = q.lay;
        int v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = p0 + 1 + u, a = j - i, b = t - a;
            v[u] = INF;
            if (i >= ilo && i <= ihi && a >= 1 && b >= 1) {   // source (i, j-1, k+1, l): slab (a-1, b-1), m' = m+2
                const int mm = m + 2;
                v[u] = ld16(pPM, __ldg(&lay[n1 + b - 1]) - __ldg(&lay[mm]) + (((i - 1) * (2 * mm + 2 - i)) >> 1) + (k - j)) + e;
            }
        }
        acc0 = pack_sat(v[0], v[1]);
        acc1 = pack_sat(v[2], v[3]);
    }
    const int slot = ccj_tri(j, k);
    const uint2 *__restrict__ lst = reinterpret_cast<const uint2 *>(q.outlist) + (int64_t)slot * CCJ_WIN_OUT;
    const int cnts = gact ? __ldg(&q.outcnt[slot]) : 0;
    const int cnt = cnts & 0xffff, nneg = cnts >> 16;   // the first nneg candidates have a negative energy
    // lane gl fetches / decodes entry e of its group's list (entries [lo,hi) belong to the current loop)
    auto fetch = [&](int e, int hi) -> uint2 { return e < hi ? __ldg(&lst[e]) : make_uint2(0u, 0xffffffffu); };
    auto decode = [&](uint2 L) -> int4 {
        if (L.y != 0xffffffffu) {
            const int x = (L.x >> 16) & 0xff, y = L.x >> 24, e = (int)(int16_t)(L.x & 0xffff);
            const int tl = t - x - y;                               // level of the source row (j-x, k+y)
            const int il = max(j - t + y, 1), ih = j - x - max(0, tl - (B - y));   // its cells
            if (tl >= 0 && ih >= il) {
                const int qls = (il - 1) >> 2, qhs = (ih - 1) >> 2;
                return make_int4(tl * wtot4 + (int)L.y - qls, (e & 0xffff) | ((32767 - max(e, 0)) << 16), qls, qhs - qls);
            }
        }
        return make_int4(own, 0x7fff0000, 0x7fffffff, 0);  // no candidate: no quad is inside its (empty) source row
    };
    // one LDS.128 per candidate: (source row start, energy | clamp << 16, first quad, quads - 1).
    // Quads outside the source row belong to other rows: no load, value 32767.
    // (1) candidates with a negative energy: value + energy of a blanked cell (32767) would look finite, so the mask
    //     halves are loaded too (16 bytes per lane) and the candidate is max(value+energy, mask)
    {
        const int nbn = (__reduce_max_sync(0xffffffffu, nneg) + WGRP - 1) / WGRP;
        for (int bb = 0; bb < nbn; ++bb) {
            const int4 d = decode(fetch(bb * WGRP + gl, nneg));
            __syncwarp();
            tile[0][grp][gl] = d;
            __syncwarp();
#pragma unroll
            for (int u = 0; u < WB; ++u) {
                const int4 d2 = tile[0][grp][u];
                const bool ok = (unsigned)(qd - d2.z) <= (unsigned)d2.w;
                const int2 w = ok ? ldq(srcv, d2.x) : make_int2(WIN_INF2, WIN_INF2);
                const int2 mk = ok ? ldq(srcm, d2.x) : make_int2(WIN_INF2, WIN_INF2);
                const int ee = (int)__byte_perm((unsigned)d2.y, 0u, 0x1010), cc = (int)__byte_perm((unsigned)d2.y, 0u, 0x3232);
                acc0 = min2(acc0, addmax2(min2(w.x, cc), ee, mk.x));
                acc1 = min2(acc1, addmax2(min2(w.y, cc), ee, mk.y));
            }
        }
    }
    // (2) all other candidates (energy >= 0, the bulk): 32767 + energy saturates to 32767 by itself, so only the value
    //     halves are loaded (8 bytes per lane) and a candidate is two packed instructions per word
    const int nb = (__reduce_max_sync(0xffffffffu, cnt - nneg) + WGRP - 1) / WGRP;   // batches of 8, warp-uniform
#define ISSUE(W_, EC_, buf)                                                                       \
    _Pragma("unroll") for (int u = 0; u < WB; ++u) {                                              \
        const int4 d2 = tile[buf][grp][u];                                                        \
        const bool ok = (unsigned)(qd - d2.z) <= (unsigned)d2.w;                                  \
        W_[u] = ok ? ldq(srcv, d2.x) : make_int2(WIN_INF2, WIN_INF2);                             \
        EC_[u] = d2.y;                                                                            \
    }
#define CONSUME(W_, EC_)                                                                          \
    _Pragma("unroll") for (int u = 0; u < WB; ++u) {                                              \
        const int ee = (int)__byte_perm((unsigned)EC_[u], 0u, 0x1010), cc = (int)__byte_perm((unsigned)EC_[u], 0u, 0x3232); \
        acc0 = addmin2(min2(W_[u].x, cc), ee, acc0);                                              \
        acc1 = addmin2(min2(W_[u].y, cc), ee, acc1);                                              \
    }
    if (PIPE) {  // see k_winLR
        int2 wa[WB], wb[WB];
        int ea[WB], eb[WB];
        uint2 pre = make_uint2(0u, 0xffffffffu);
        if (nb > 0) {
            tile[0][grp][gl] = decode(fetch(nneg + gl, cnt));
            pre = fetch(nneg + WGRP + gl, cnt);
            __syncwarp();
            ISSUE(wa, ea, 0);
        }
        for (int bb = 0; bb < nb; bb += 2) {
            if (bb + 1 < nb) {
                tile[1][grp][gl] = decode(pre);
                pre = fetch(nneg + (bb + 2) * WGRP + gl, cnt);
                __syncwarp();
                ISSUE(wb, eb, 1);
            }
            CONSUME(wa, ea);
            __syncwarp();
            if (bb + 1 < nb) {
                if (bb + 2 < nb) {
                    tile[0][grp][gl] = decode(pre);
                    pre = fetch(nneg + (bb + 3) * WGRP + gl, cnt);
                    __syncwarp();
                    ISSUE(wa, ea, 0);
                }
                CONSUME(wb, eb);
                __syncwarp();
            }
        }
    } else {
        uint2 pre = fetch(nneg + gl, cnt);
        for (int bb = 0; bb < nb; ++bb) {
            const int4 d = decode(pre);
            pre = fetch(nneg + (bb + 1) * WGRP + gl, cnt);
            __syncwarp();
            tile[0][grp][gl] = d;
            __syncwarp();
            int2 w[WB];
            int ec[WB];
            ISSUE(w, ec, 0);
            CONSUME(w, ec);
        }
    }
#undef ISSUE
#undef CONSUME
    if (gact && qd <= qhi) {
        int2 *__restrict__ out = reinterpret_cast<int2 *>(q.wscr + 4 * q.wscr_lr + (int64_t)(t & 1) * 4 * wtot4);
        out[own - t * wtot4 + qd] = make_int2(acc0, acc1);
    }
}

// same-cell assembly in the reference's order (src/pseudo_loop.cc:85-127) from the partial minima
// layout tables read straight from global memory (L1-resident): k_final needs a dozen entries with block-uniform
// indices, not worth staging in shared memory
struct LayTab {
    const int *p;
    __device__ __forceinline__ int operator[](int x) const { return __ldg(p + x); }
};
#ifndef FINAL_MINB
#define FINAL_MINB 8
#endif
// CCJ_ABLATE: timing experiments only (profiles/exp_ablate.py) -- every bit removes one class of memory operations from
// k_final, so the tables are WRONG when it is non-zero; the shipped library is built without it.
//   1 scattered copies (PRW, PKG, PMW, PMM)   2 plain tables the fill never reads   4 read-group records
//   8 scratch reads   16 late split points   32 lower-level plain-table reads   64 PLW/PKF   128 all plain tables
#ifndef CCJ_ABLATE
#define CCJ_ABLATE 0
#endif
__global__ void __launch_bounds__(K4_THREADS, FINAL_MINB) k_final(const ccj_model *__restrict__ M, const ccj_seq *__restrict__ seqs, int t, int tail) {
    const ccj_seq q = seqs[blockIdx.z];
    const int n = q.n;
    if (n - t - 2 < 1) return;
    const LayTab s_tet{q.lay}, s_cb{q.lay + (n + 1)}, s_hh{q.lay + 3 * (n + 1)}, s_cw{q.lay + 4 * (n + 1)};
    Cell C;
    {   // decode of the thread's cell (cell_setup without the shared tables)
        const int m = n - t - 2, ncell = m * (m + 1) / 2;
        const int p = blockIdx.x * K4_THREADS + threadIdx.x;
        if (p >= ncell) return;
        int r = (int)(((2 * m + 1) - sqrtf((float)((2 * m + 1) * (2 * m + 1) - 8 * p))) * 0.5f);
        r = max(0, min(r, m - 1));
        while (r > 0 && r * (2 * m + 1 - r) / 2 > p) --r;
        while ((r + 1) * (2 * m - r) / 2 <= p) ++r;
        C.n = n; C.m = m; C.a = blockIdx.y; C.b = t - C.a;
        C.i = r + 1;
        C.j = C.i + C.a; C.k = C.j + 2 + (p - r * (2 * m + 1 - r) / 2); C.l = C.k + C.b;
        C.p = p;
        C.c = C.a * ncell + p;
    }
    const int a = C.a, b = C.b, i = C.i, j = C.j, k = C.k, l = C.l;
    const int16_t *__restrict__ t4 = q.t4;
    const int64_t st4 = q.stride4;
    const int n1 = n + 1;
    const int INF = CCJ_INF, II = CCJ_INTERN_INF;
    auto H4 = [](int x) { return ((x >> 2) + 1) * (2 * (x >> 2) + (x & 3)); };
    const int bp = M->bp_penalty, cp = M->cp_penalty, PB = M->PB_penalty, apbp = M->ap_penalty + M->bp_penalty;
    const int64_t ss = q.scratch_stride;
    const int16_t *__restrict__ sc = q.scratch + (int64_t)(t % KF) * Q_COUNT * ss + C.c;
#if CCJ_ABLATE & 8
#define GET(id) (20000 + (id))
#else
#define GET(id) ((int)__ldg(sc + (int64_t)(id) * ss))
#endif
    const int off0 = OFF(a, b, i, k);
#if CCJ_ABLATE & 32
#define ld16(p, o) (1000 + ((o) & 15))
#endif
    // partial minima of the four split-point roles.  tail = t mod KF: k_roles(t-tail) left them for the sources
    // of levels <= t-tail-1 (nothing for a role whose arm is shorter than tail: no partner there); the last `tail`
    // split points of each role, whose sources are cells of levels t-tail..t-1, are added here.
    AccL1 L1 = {INF, INF, INF, INF, INF, INF, INF};
    AccL2 L2 = {INF, INF, INF, INF, INF, INF, INF};
    AccR3 R3 = {INF, INF, INF, INF, INF, INF};
    AccR4 R4 = {INF, INF, INF, INF, INF, INF, INF, INF, INF};
    if (a >= tail) {
        L1 = {GET(Q_PK1), GET(Q_PfL2), GET(Q_PfM), GET(Q_PLm00a), GET(Q_PLm01), GET(Q_PLm10a), GET(Q_PMm00a)};
        L2 = {GET(Q_PfL1), GET(Q_PfO1), GET(Q_PLm00b), GET(Q_PLm10b), GET(Q_PMm10a), GET(Q_POm00a), GET(Q_POm10a)};
    }
    if (b >= tail) {
        R3 = {GET(Q_PK3), GET(Q_PfR1), GET(Q_PfMp), GET(Q_PRm00a), GET(Q_PRm10), GET(Q_PMm00b)};
        R4 = {GET(Q_PfR2), GET(Q_PfO2), GET(Q_PRm00b), GET(Q_PRm01), GET(Q_PMm01), GET(Q_PMm10b), GET(Q_POm00b), GET(Q_POm01), GET(Q_POm10b)};
    }
    if (tail && !(CCJ_ABLATE & 16)) {
        const int4 *__restrict__ W3 = reinterpret_cast<const int4 *>(q.w3);
        // x = arm of the source cell (0 = the boundary split point of the role, its *_first form)
        for (int x = max(0, a - tail); x < a; ++x) {
            const int4 wl = __ldg(&W3[(a - x - 1) * n1 + i + x + 1]);                                   // (d+1, j), d=i+x
            const R12 r1 = ldr12(reinterpret_cast<const int *>(q.g1), OFF(x, b, i, k));              // X(i,i+x,k,l)
            if (x == 0) l1_first(L1, r1, wl); else l1_term(L1, r1, wl);
            const int4 wr = __ldg(&W3[(a - x - 1) * n1 + i]);                                           // (i, d-1), d=j-x
            const R12 r2 = ldr12(reinterpret_cast<const int *>(q.g2), OFF(x, b, j - x, k));          // X(j-x,j,k,l)
            if (x == 0) l2_first(L2, r2, wr); else l2_term(L2, r2, wr);
        }
        for (int x = max(0, b - tail); x < b; ++x) {
            const int4 wl = __ldg(&W3[(b - x - 1) * n1 + k]);                                           // (k, d-1), d=l-x
            const R12 r3 = ldr12(reinterpret_cast<const int *>(q.g3), OFF(a, x, i, l - x));          // X(i,j,l-x,l)
            if (x == 0) r3_first(R3, r3, wl); else r3_term(R3, r3, wl);
            const int4 wr = __ldg(&W3[(b - x - 1) * n1 + k + x + 1]);                                   // (d+1, l), d=k+x
            const int4 r4 = __ldg(reinterpret_cast<const int4 *>(q.g4) + OFF(a, x, i, k));           // X(i,j,k,k+x)
            if (x == 0) r4_first(R4, r4, wr); else r4_term(R4, r4, wr);
        }
    }
    int16_t *w4 = q.t4;
    const int mloc = n - t - 2, kr = n - b - k;
    const int h4m = H4(mloc);
    // entry of (j,k)'s row of level t that holds this cell (ccj_types.h, PMW)
    const int pmrow = 4 * (__ldg(&q.pmlev4[j * n1 + k]) - ((max(j - t, 1) - 1) >> 2)) + (i - 1);
    // Matrix4D::set narrows int32 -> int16 without a lower clamp (src/matrices.hh:188-191): an entry below -32768
    // wraps.  The partial minima and the packed window arithmetic of this path saturate there instead, so a
    // sequence that comes near that range (only designed inputs can: -310 kcal/mol inside one gapped region) is
    // flagged and re-filled by the generic kernels, whose every candidate is evaluated in 32 bits (ccj_batch_fill).
    // The guard band covers the largest single step of the packed paths (a window energy, the PB penalty).
    int32_t *const wrap_flag = q.status + 7;
    auto put16 = [wrap_flag](int16_t *dst, int mn, bool store) -> int {
        int v = CCJ_INTERN_INF;
        if (mn < CCJ_INF / 2) {
            if (mn >= CCJ_INTERN_INF) mn = CCJ_INTERN_INF;
            if (mn < CCJ_WRAP_GUARD) *wrap_flag = 1;
            v = (int)(int16_t)mn;
        }
        if (store) *dst = (int16_t)v;
        return v;
    };
#define PUT_UNREAD(tbl) ((tbl) == T_PK || (tbl) == T_PfromMprime || (tbl) == T_PLmloop00 || (tbl) == T_PMmloop00 || \
                         (tbl) == T_POmloop00 || (tbl) == T_PRmloop00 || (tbl) == T_PL || (tbl) == T_PR)
#define PUT(tbl, val) put16(w4 + (int64_t)(tbl) * st4 + off0, (val), !((CCJ_ABLATE & 128) || ((CCJ_ABLATE & 2) && PUT_UNREAD(tbl))))
    const int vPLm00 = PUT(T_PLmloop00, min(II + bp, min(L1.PLm00, L2.PLm00)));
    PUT(T_PLmloop01, L1.PLm01);
    const int vPLm10 = PUT(T_PLmloop10, min(L1.PLm10, L2.PLm10));
    const int vPRm00 = PUT(T_PRmloop00, min(II + bp, min(R3.PRm00, R4.PRm00)));
    int vPMm00, vPMm10;
    {
        int e01 = INF, e10 = INF, f01 = INF;
        if (b >= 1) {
            const int oA = OFF(a, b - 1, i, k);      // (i,j,k,l-1)
            const int oB = OFF(a, b - 1, i, k + 1);  // (i,j,k+1,l)
            e01 = ld16(TB(T_PRmloop01), oA) + cp;
            e10 = ld16(TB(T_PRmloop10), oB) + cp;
            f01 = ld16(TB(T_PMmloop01), oB) + cp;
        }
        PUT(T_PRmloop01, min(e01, R4.PRm01));
        PUT(T_PRmloop10, min(e10, R3.PRm10));
        vPMm00 = PUT(T_PMmloop00, min(II + bp, min(L1.PMm00, R3.PMm00)));
        PUT(T_PMmloop01, min(f01, R4.PMm01));
        int g10 = INF;
        if (a >= 1) g10 = ld16(TB(T_PMmloop10), OFF(a - 1, b, i, k)) + cp;  // (i,j-1,k,l)
        vPMm10 = PUT(T_PMmloop10, min(g10, min(L2.PMm10, R4.PMm10)));
    }
    const int vPOm00 = PUT(T_POmloop00, min(II + bp, min(L2.POm00, R4.POm00)));
    PUT(T_POmloop01, R4.POm01);
    const int vPOm10 = PUT(T_POmloop10, min(L2.POm10, R4.POm10));

    const int8_t *__restrict__ S = q.S;
    auto ptype = [&](int x, int y) { return __ldg(&M->pair[S[x]][S[y]]); };
    // a window partial of 32767 stands for "INF or clamped": identical after the final clamp
    int vPL, vPR, vPM, vPO;
    {  // PL (src/pseudo_loop.cc:232-253)
        int mn = INF;
        if (ptype(i, j) > 0 && a >= 2) {
            const int o = OFF(a - 2, b, i + 1, k);  // (i+1,j-1,k,l)
            // window partial, PLW layout: row i-1, position (n-b)-k; rows with a<=TURN have no window
            mn = a > CCJ_TURN ? (int)__ldg(q.wscr + (int64_t)((t & 1) * 2) * q.wscr_lr + 4 * (a * h4m + h4m - H4(mloc - i + 1)) + (n - b - k)) : 32767;
            mn = min(mn, min(ld16(TB(T_PLmloop10), o), ld16(TB(T_PLmloop01), o)) + apbp + bp);
            if (a >= CCJ_TURN + 1) mn = min(mn, ld16(TB(T_PfromL), o));
        }
        vPL = PUT(T_PL, mn);
    }
    {  // PR (:255-275)
        int mn = INF;
        if (ptype(k, l) > 0 && b >= 2) {
            const int o = OFF(a, b - 2, i, k + 1);  // (i,j,k+1,l-1)
            // window partial, PRW layout: row kr=n-b-k, position i-1
            mn = b > CCJ_TURN ? (int)__ldg(q.wscr + (int64_t)((t & 1) * 2 + 1) * q.wscr_lr + 4 * (a * h4m + h4m - H4(mloc - kr)) + (i - 1)) : 32767;
            mn = min(mn, min(ld16(TB(T_PRmloop10), o), ld16(TB(T_PRmloop01), o)) + apbp + bp);
            if (b >= CCJ_TURN + 1) mn = min(mn, ld16(TB(T_PfromR), o));
        }
        vPR = PUT(T_PR, mn);
    }
    {  // PM (:277-300)
        int mn = INF;
        if (ptype(j, k) > 0) {
            if (a >= 1 && b >= 1) {
                const int o = OFF(a - 1, b - 1, i, k + 1);  // (i,j-1,k+1,l)
                // window partial, PMW row layout of this level
                mn = k - j > CCJ_TURN ? (int)__ldg(q.wscr + 4 * q.wscr_lr + (int64_t)(t & 1) * 4 * q.wtot4 + pmrow) : 32767;
                mn = min(mn, min(ld16(TB(T_PMmloop10), o), ld16(TB(T_PMmloop01), o)) + apbp + bp);
                mn = min(mn, ld16(TB(T_PfromM), o));
            }
            if (a == 0 && b == 0) mn = min(mn, 0);
        }
        vPM = PUT(T_PM, mn);
    }
    {  // PO (:302-322; the window of get_POiloop is dead, :787-808)
        int mn = INF;
        if (ptype(i, l) > 0 && a >= 1 && b >= 1) {
            const int o = OFF(a - 1, b - 1, i + 1, k);  // (i+1,j,k,l-1)
            mn = ld16(TB(T_PO), o) + __ldg(&q.estP[(l - i) * n1 + i]);
            mn = min(mn, min(ld16(TB(T_POmloop10), o), ld16(TB(T_POmloop01), o)) + apbp + bp);
            mn = min(mn, ld16(TB(T_PfromO), o));
        }
        vPO = PUT(T_PO, mn);
    }
    const int vPfL = PUT(T_PfromL, min(min(L2.PfL, L1.PfL), min(vPR + PB, min(vPM + PB, vPO + PB))));
    const int vPfR = PUT(T_PfromR, min(min(R3.PfR, R4.PfR), min(vPM + PB, vPO + PB)));
    PUT(T_PfromM, L1.PfM);
    const int vPfMp = PUT(T_PfromMprime, R3.PfMp + PB);
    const int vPfO = PUT(T_PfromO, min(min(L2.PfO, R4.PfO), min(vPL + PB, vPR + PB)));
    const int vPK = PUT(T_PK, min(min(L1.PK, R3.PK), min(min(vPL + PB, vPM + PB), min(vPR + PB, vPO + PB))));
    {   // the two PK copies of compute_P (ccj_types.h).  As first factor PK(i,j,d+1,k') this cell is d=k-1, k'=l, i.e.
        // row delta=l-k+1 of block (i,j), position k-j-2: coalesced.  As second factor PK(i2,d,k2,l) it is i2=i, d=j,
        // k2=k: row delta=k-j-1 of block (i,l), position j-i: scattered, one store per cell.
        const int *__restrict__ lay = q.lay;
        const int Mf = n - j - 1, Mg = l - i - 1;
        if (!(CCJ_ABLATE & 64))
        q.pkf[__ldg(&lay[CCJ_LAY_DF(n) + i]) + __ldg(&lay[CCJ_LAY_CF(n) + j]) - __ldg(&lay[CCJ_LAY_CF(n) + i]) +
              8 * ((int)ccj_q8(Mf) - (int)ccj_q8(Mf + 1 - (b + 1))) + (k - j - 2)] = (int16_t)vPK;
        if (!(CCJ_ABLATE & 1))
        q.pkg[__ldg(&lay[CCJ_LAY_EG(n) + i]) + __ldg(&lay[CCJ_LAY_S2(n) + Mg]) +
              8 * ((int)ccj_q8(Mg) - (int)ccj_q8(Mg + 1 - (k - j - 1))) + (j - i)] = (int16_t)vPK;
    }
    {   // window copies (layouts in ccj_types.h); PLW is coalesced, PRW / PMW are one scattered store per cell
        const int slab4 = s_cw[b] + s_hh[n - b - 2] - s_hh[mloc] + h4m;
        if (!(CCJ_ABLATE & 64)) q.plw[4 * (int64_t)(slab4 - H4(mloc - i + 1)) + (n - b - k)] = (int16_t)vPL;
        if (!(CCJ_ABLATE & 1)) q.prw[4 * (int64_t)(slab4 - H4(mloc - kr)) + (i - 1)] = (int16_t)vPR;
        // PMW / PMM: the PM window only reads cells with a>=1, b>=1 whose (j,k) is in some partner list, i.e. can pair;
        // everything else keeps the (32767, "not a source") that k_fill_pmw wrote
        if (!(CCJ_ABLATE & 1) && a >= 1 && b >= 1 && k - j > CCJ_TURN && ptype(j, k) > 0) {
            const int64_t pe = 4 * (int64_t)t * q.wtot4 + pmrow;   // entry of this cell in PMW and PMM
            q.pmw[pe] = (int16_t)vPM;
            q.pmm[pe] = (int16_t)-32768;
        }
    }
    // read-group records (layout in ccj_types.h); consecutive cells -> consecutive records, coalesced
    if (!(CCJ_ABLATE & 4)) {
        const int vMpp = min(vPL, vPR);
        auto pk2 = [](int lo, int hi) { return (int)((uint32_t)(uint16_t)(int16_t)lo | ((uint32_t)(uint16_t)(int16_t)hi << 16)); };
        int *r1 = reinterpret_cast<int *>(q.g1) + (int64_t)off0 * 3;
        r1[0] = pk2(vPK, vPfL); r1[1] = pk2(vPfMp, vPLm00); r1[2] = pk2(vPLm10, vPMm00);
        int *r2 = reinterpret_cast<int *>(q.g2) + (int64_t)off0 * 3;
        r2[0] = pk2(vPfL, vPfO); r2[1] = pk2(vPLm00, vPMm00); r2[2] = pk2(vPOm00, 0);
        int *r3 = reinterpret_cast<int *>(q.g3) + (int64_t)off0 * 3;
        r3[0] = pk2(vPK, vPfR); r3[1] = pk2(vMpp, vPRm00); r3[2] = pk2(vPMm00, 0);
        int4 *r4 = reinterpret_cast<int4 *>(q.g4) + off0;
        *r4 = make_int4(pk2(vPfR, vPfO), pk2(vPRm00, vPMm00), pk2(vPMm10, vPOm00), pk2(vPOm10, 0));
    }
#undef PUT
#undef GET
#if CCJ_ABLATE & 32
#undef ld16
#endif
}

// P(i,l) = min_{i<=j<d<k<l} PK(i,j,d+1,k) + PK(j+1,d,k+1,l)  (src/pseudo_loop.cc:166-179).
// For fixed (i,j,l), L=l-j-2, the terms are rows delta=1..L over d=j+1..l-delta-1: row delta of PKF block (i,j) and row
// delta of PKG block (j+1,l), both starting on 16-byte boundaries (ccj_types.h) -- 8 consecutive terms are one
// 16-byte load from each copy.  blockIdx.x -> i; work items = (j, 32 octets), dealt round-robin to the warps (the
// term sets range from 1 to thousands of terms); a lane takes one octet of one row.
__global__ void __launch_bounds__(256) k_P_tuned(const ccj_seq *__restrict__ seqs, int s, int nj) {
    __shared__ int s_cf[K4_MAXN + 4], s_eg[K4_MAXN + 4], s_s2[K4_MAXN + 4];
    __shared__ int sm[8];
    const ccj_seq q = seqs[blockIdx.z];
    const int n = q.n;
    const int i = 1 + blockIdx.x, l = i + s;
    if (l > n) return;
    const int *__restrict__ lay = q.lay;
    for (int x = threadIdx.x; x <= n + 1; x += 256) {
        s_cf[x] = __ldg(&lay[CCJ_LAY_CF(n) + x]);
        s_eg[x] = __ldg(&lay[CCJ_LAY_EG(n) + x]);
        s_s2[x] = __ldg(&lay[CCJ_LAY_S2(n) + x]);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int16_t *__restrict__ F0 = q.pkf + (__ldg(&lay[CCJ_LAY_DF(n) + i]) - s_cf[i]);   // + CF[j]: block (i,j)
    const int16_t *__restrict__ G0 = q.pkg;                                               // + EG[j+1] + S2[L]: block (j+1,l)
    auto Q8 = [](int x) { const int g = x >> 3, r = x & 7; return 4 * g * (g + 1) + r * (g + 1); };
    int mn = CCJ_INF;
    // the j with more than 32*p octets are a prefix (the term sets shrink with j), which makes the item list a double loop
    const int Lmax = l - i - 2, Omax = Q8(Lmax);
    int Lmin = 1;
    for (int p = 0; p * 32 < Omax; ++p) {
        while (Q8(Lmin) <= p * 32) ++Lmin;   // smallest L with more than 32*p octets
        const int jmax = l - 2 - Lmin;       // L = l-j-2 >= Lmin
        for (int j = i + blockIdx.y + ((wid + p) & 7) * nj; j <= jmax; j += 8 * nj) {  // rotated: every warp gets large and small j
            const int L = l - j - 2, O = Q8(L);
            const int o = p * 32 + lane;
            if (o >= O) continue;
            // octet o of the rows ordered by length: rows of 8g+1..8g+8 entries have g+1 octets each, 4g(g+1) octets before them
            int g = (int)((sqrtf(1.f + (float)o) - 1.f) * 0.5f);
            while (g > 0 && 4 * g * (g + 1) > o) --g;
            while (4 * (g + 1) * (g + 2) <= o) ++g;
            const int rem = o - 4 * g * (g + 1), r = rem / (g + 1), c = rem - r * (g + 1);
            const int len = 8 * g + r + 1;           // entries of the row: delta = L+1-len
            const int Mf = n - j - 1;                // PKF block (i,j): row delta starts at 8*(Q8(Mf) - Q8(Mf+1-delta))
            const int4 f = __ldg(reinterpret_cast<const int4 *>(F0 + s_cf[j] + 8 * (Q8(Mf) - Q8(Mf - L + len) + c)));
            const int4 gg = __ldg(reinterpret_cast<const int4 *>(G0 + s_eg[j + 1] + s_s2[L] + 8 * (O - Q8(len) + c)));
            const int left = len - 8 * c;            // terms of this octet that exist (padding / later d follow)
            int v0 = lo16(f.x) + lo16(gg.x), v1 = hi16(f.x) + hi16(gg.x), v2 = lo16(f.y) + lo16(gg.y), v3 = hi16(f.y) + hi16(gg.y);
            int v4 = lo16(f.z) + lo16(gg.z), v5 = hi16(f.z) + hi16(gg.z), v6 = lo16(f.w) + lo16(gg.w), v7 = hi16(f.w) + hi16(gg.w);
            if (left < 8) {
                if (left < 2) v1 = CCJ_INF;
                if (left < 3) v2 = CCJ_INF;
                if (left < 4) v3 = CCJ_INF;
                if (left < 5) v4 = CCJ_INF;
                if (left < 6) v5 = CCJ_INF;
                if (left < 7) v6 = CCJ_INF;
                v7 = CCJ_INF;
            }
            mn = min(mn, min(min(min(v0, v1), min(v2, v3)), min(min(v4, v5), min(v6, v7))));
        }
    }
    mn = __reduce_min_sync(0xffffffffu, mn);
    if (lane == 0) sm[wid] = mn;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int x = 1; x < 8; ++x) mn = min(mn, sm[x]);
        if (mn < CCJ_INF / 2) atomicMin(&q.t2[T2_P * q.stride2 + ccj_idx2(n, i, l)], mn);
    }
}

#ifndef CCJ_HOST_EMU   // the launchers below are CUDA only; the emulation harness issues the same grids itself
// partner lists + e_stP table only (folds that run the generic cell functions over lists: beyond the tuned range, sharded)
void launch_prep_lists(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, cudaStream_t st) {
    if (d.nmax < 2) return;
    k_prep<<<dim3(d.nmax - 1, d.nseq), 128, 0, st>>>(M, seqs);
}

void launch_prep(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, cudaStream_t st) {
    if (d.nmax < 2) return;
    if (d.nmax <= K4_MAXN) {
        k_prep_lay<<<d.nseq, 256, 0, st>>>(M, seqs);
        k_fill_pmw<<<dim3(296, d.nseq), 256, 0, st>>>(seqs);
    }
    k_prep<<<dim3(d.nmax - 1, d.nseq), 128, 0, st>>>(M, seqs);
}

// K4_MAXN, or less when CCJ_TUNED_MAXN is set (read once; lets the tests drive the beyond-the-range path cheaply)
bool fill4_tuned_supported(int nmax) {
    static const int limit = [] {
        const char *e = getenv("CCJ_TUNED_MAXN");
        const int v = e ? atoi(e) : K4_MAXN;
        return v < K4_MAXN ? v : K4_MAXN;
    }();
    return nmax <= limit;
}

static bool level_dims(LaunchDims d, int t, int &bx) {
    const int m = d.nmax - t - 2;
    if (m < 1) return false;
    bx = (m * (m + 1) / 2 + K4_THREADS - 1) / K4_THREADS;
    return true;
}
void launch_4d_roles(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int t, cudaStream_t st) {
    int bx;
    if (t % KF == 0 && level_dims(d, t, bx)) k_roles<<<dim3(bx, t + 1, d.nseq * 4), K4_THREADS, 0, st>>>(M, seqs, t);  // levels t..t+KF-1
}
void launch_4d_windows(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int t, cudaStream_t st) {
    const int nm = d.nmax, m = nm - t - 2;
    if (m < 1) return;
    // small waves are latency-bound: software-pipelined variants.  Measured break-even (B200): n^2 x sequences ~ 1.2e5
    // (4 x 150 nt, 8 x 100 nt, 2 x 200 nt still gain; 8 x 150 nt, 16 x 100 nt lose)
    const bool pipe = (long long)nm * nm * d.nseq < 120000;
    {   // PL / PR: at most m paired rows per slab, runs of up to m cells
        // wide groups (16 lanes = 64 cells per pass, two runs per warp request instead of four) for levels whose runs
        // reach 32 cells: measured 76.9 -> 72.8 ms for the window kernels of a 32 x 150 nt fill, monotonic in the
        // threshold (profiles/r2_notes.md); CCJ_WINLR_WIDE_FROM overrides the threshold for experiments
        static const int wide_from = [] { const char *e = getenv("CCJ_WINLR_WIDE_FROM"); return e ? atoi(e) : 32; }();
        if (!pipe && m >= wide_from) {
            const int runs = K4_THREADS / 16, nchunk = (m + runs - 1) / runs, npass = (m + 63) / 64;
            k_winLR<false, 16><<<dim3(nchunk * npass, t + 1, d.nseq * 2), K4_THREADS, 0, st>>>(M, seqs, t, nchunk);
        } else {
            const int nchunk = (m + WRUNS - 1) / WRUNS, npass = (m + 4 * WGRP - 1) / (4 * WGRP);
            const dim3 grid(nchunk * npass, t + 1, d.nseq * 2);
            if (pipe) k_winLR<true, 8><<<grid, K4_THREADS, 0, st>>>(M, seqs, t, nchunk);
            else k_winLR<false, 8><<<grid, K4_THREADS, 0, st>>>(M, seqs, t, nchunk);
        }
    }
    {   // PM: pairs (j,k) with TURN < k-j <= n-1-t; a row has at most min(t+1, n-4-t) cells at any quad phase
        long long rows = 0;
        for (int s = CCJ_TURN + 1; s <= nm - 1 - t; ++s) rows += nm - s;
        if (rows < 1) return;
        const int cm = std::max(1, std::min(t + 1, nm - 4 - t));
        const int nq = ((cm + 2) >> 2) + 1;
        const int nchunk = (int)((rows + WRUNS - 1) / WRUNS), npass = (nq + WGRP - 1) / WGRP;
        const dim3 grid(nchunk * npass, d.nseq);
        if (pipe) k_winM<true><<<grid, K4_THREADS, 0, st>>>(M, seqs, t, nchunk);
        else k_winM<false><<<grid, K4_THREADS, 0, st>>>(M, seqs, t, nchunk);
    }
}
void launch_4d_final(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int t, cudaStream_t st) {
    int bx;
    if (level_dims(d, t, bx)) k_final<<<dim3(bx, t + 1, d.nseq), K4_THREADS, 0, st>>>(M, seqs, t, t % KF);
}
void launch_4d_tuned(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int t, cudaStream_t st) {
    launch_4d_roles(M, seqs, d, t, st);
    launch_4d_windows(M, seqs, d, t, st);
    launch_4d_final(M, seqs, d, t, st);
}

int fill4_partials() { return KF * Q_COUNT; }  // KF levels of partial minima (k_roles works on levels t..t+KF-1)
int fill4_fused_levels() { return KF; }

void launch_P_tuned(const ccj_model *, const ccj_seq *seqs, LaunchDims d, int s, cudaStream_t st) {
    if (s < 3 || s > d.nmax - 1) return;
    // one block per (i, residue class of j); enough classes to fill the 148 SMs a few times when the wave is small
    const int per = (d.nmax - s) * d.nseq;
    const int nj = std::max(1, std::min(s - 2, (148 * 6 + per - 1) / per));
    k_P_tuned<<<dim3(d.nmax - s, nj, d.nseq), 256, 0, st>>>(seqs, s, nj);
}
#endif  // CCJ_HOST_EMU

}  // namespace ccj
