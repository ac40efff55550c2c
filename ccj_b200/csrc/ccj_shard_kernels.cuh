// The kernels of the row-sharded fold (ccj_shard.cu includes this file inside its unnamed namespace; tests/emu/ccj_emu_tuned.cpp
// does the same to run them on the host under the SIMT emulator).  Launch sequence and collectives: ccj_shard.cu.
#pragma once
#include <algorithm>
#include <cstdlib>

#include "ccj_cells4_lean.cuh"
#include "ccj_kernels.cuh"

// ---- kernels -------------------------------------------------------------------------------------------------------------
// level t of the rank's rows: blockIdx.y -> a (b=t-a), threads walk the rank's cells of slab (a,b) in storage order
// (rows of decreasing length back to back, so every lane has a cell and a warp's stores are contiguous)
// 12 blocks of 128 threads per SM (<= 42 registers, ~0.3 KB of spills per thread): the kernel is latency-bound -- a serial
// chain of dependent loads per cell -- and occupancy buys more than the spills cost (300-nt fill: 3.23 s at 4 blocks/SM,
// 2.90 / 2.54 / 2.21 / 2.04 / 1.95 / 1.97 s at 5 / 6 / 8 / 10 / 12 / 16)
__global__ void __launch_bounds__(128, 12) k_4d_shard(const ccj_model *M, const ccj_seq *seqs, int t) {
    ccj_cx c;
    c.M = M;
    c.q = seqs[0];
    const int n = c.q.n, G = c.q.shard_G, r = c.q.shard_rank;
    const int m = n - t - 2;
    if (m <= r) return;
    const int mr = m - r, Q = (mr + G - 1) / G;
    const int ncell = Q * mr - G * (Q * (Q - 1) / 2);
    const int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= ncell) return;
    int q, kk;
    ccj_shard_cell_of(p, mr, G, Q, q, kk);
    const int a = blockIdx.y, b = t - a;
    const int i = r + 1 + q * G, k = i + a + 2 + kk;
    ccj_cell4d(c, i, i + a, k, k + b);
}

// The same level with the lean cell function (ccj_cells4_lean.cuh): one position per split point from level bases kept
// in shared memory, compile-time table kinds.  Needs the packed 2D records, the partner lists and 32-bit in-level
// offsets (shard_lean_ok); CCJ_SHARD_LEAN=0 selects k_4d_shard for comparison.
#ifndef SHARD_LEAN_MINB
#define SHARD_LEAN_MINB 8
#endif
template <bool POW2>
__global__ void __launch_bounds__(128, SHARD_LEAN_MINB) k_4d_shard_lean(const ccj_model *M, const ccj_seq *seqs, int t) {
    CCJ_DYN_SHARED(ccj_lean_lvl, s_lvl);
    ccj_cx c;
    c.M = M;
    c.q = seqs[0];
    const int n = c.q.n, G = c.q.shard_G, r = c.q.shard_rank;
    ccj_lean_lvl_fill(s_lvl, c.q.shard_lev, t, G, threadIdx.x, 128);   // sources are cells of levels < t, the cell itself is on t
    __syncthreads();
    const int m = n - t - 2;
    if (m <= r) return;
    const int mr = m - r, Q = (mr + G - 1) / G;
    const int ncell = Q * mr - G * (Q * (Q - 1) / 2);
    const int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= ncell) return;
    int q, kk;
    ccj_shard_cell_of(p, mr, G, Q, q, kk);
    const int a = blockIdx.y, b = t - a;
    const int i = r + 1 + q * G, k = i + a + 2 + kk;
    ccj_lean_shard<POW2> ly;
    ly.rep = c.q.shard_rep;
    ly.loc = c.q.shard_loc[r];
    ly.lvl = s_lvl;
    ly.n = n; ly.G = G; ly.sh = c.q.shard_shift;
    ccj_cell4d_lean(c, ly, i, i + a, k, k + b);
}
// the lean kernel applies: lists + packed 2D records present, kinds as compiled in, in-level offsets fit 32 bits
bool shard_lean_ok(const ccj_seq &q, const int64_t *lev_host, int n, int G) {
    static const bool off = [] { const char *e = getenv("CCJ_SHARD_LEAN"); return e && e[0] == '0'; }();
    if (off || !q.use_lists || !q.w3 || !ccj_lean_kinds_ok(q.shard_kind)) return false;
    int64_t cmax = 0;
    for (int t = 0; t <= n; ++t) cmax = std::max<int64_t>(cmax, lev_host[t + 1] - lev_host[t]);
    return (int64_t)(CCJ_SHARD_NREP * G + 1) * cmax < (int64_t)0x7fffffff && (size_t)(n + 1) * sizeof(ccj_lean_lvl) <= 40000;
}

// P(i,l), l=i+s, for the rank's rows: blockIdx.x -> own row, blockIdx.y -> j (first split point), threads -> (d,k)
__global__ void __launch_bounds__(256) k_P_shard(const ccj_model *M, const ccj_seq *seqs, int s) {
    ccj_cx c;
    c.M = M;
    c.q = seqs[0];
    const int n = c.q.n;
    const int i = c.q.shard_rank + 1 + blockIdx.x * c.q.shard_G, l = i + s;
    if (l > n) return;
    const int j = i + blockIdx.y;
    if (j >= l) return;
    // d in [j+1, l-1], k in [d+1, l-1]: warps take d, lanes walk k (the second factor PK(j+1,d,k+1,l) is contiguous in k)
    int mn = CCJ_INF;
    for (int d = j + 1 + (threadIdx.x >> 5); d < l; d += (int)(blockDim.x >> 5))
        for (int k = d + 1 + (threadIdx.x & 31); k < l; k += 32) mn = ccj_min(mn, ccj_P_term(c, i, l, j, d, k));
    mn = __reduce_min_sync(0xffffffffu, mn);
    __shared__ int sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mn;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int x = 1; x < 8; ++x) mn = ccj_min(mn, sm[x]);
        if (mn < CCJ_INF / 2) atomicMin(&c.q.t2[T2_P * c.q.stride2 + ccj_idx2(n, i, l)], mn);
    }
}

// the same with ccj_P_lean (ccj_cells4_lean.cuh): warps take delta = k-d, lanes walk d -- the first factors are consecutive
// entries of one row, the second factors a constant stride apart on one level
template <bool POW2>
__global__ void __launch_bounds__(256) k_P_shard_lean(const ccj_model *M, const ccj_seq *seqs, int s) {
    CCJ_DYN_SHARED(ccj_lean_lvl, s_lvl);
    __shared__ int sm[8];
    const ccj_seq &q = seqs[0];
    const int n = q.n, G = q.shard_G;
    ccj_lean_lvl_fill(s_lvl, q.shard_lev, s, G, threadIdx.x, 256);   // both factors lie on levels <= s-3
    __syncthreads();
    const int i = q.shard_rank + 1 + blockIdx.x * G, l = i + s;
    if (l > n) return;
    const int j = i + blockIdx.y;
    if (j >= l) return;
    ccj_lean_shard<POW2> ly;
    ly.rep = q.shard_rep;
    ly.loc = nullptr;   // PK is a column-read table
    ly.lvl = s_lvl;
    ly.n = n; ly.G = G; ly.sh = q.shard_shift;
    int mn = ccj_P_lean(ly, i, j, l, threadIdx.x >> 5, 8, threadIdx.x & 31, 32);
    mn = __reduce_min_sync(0xffffffffu, mn);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mn;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int x = 1; x < 8; ++x) mn = ccj_min(mn, sm[x]);
        if (mn < CCJ_INF / 2) atomicMin(&q.t2[T2_P * q.stride2 + ccj_idx2(n, i, l)], mn);
    }
}

// in-process group only: element-wise minimum of `count` int32 at `dst` and `src` into both (allreduce-min by pairs)
__global__ void k_min_into(int32_t *dst, const int32_t *src, int count) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x < count) dst[x] = min(dst[x], src[x]);
}

// one gap table of the sharded fold in the ordinary storage order (ccj_idx4), for export / hashing; runs on a rank
// that can read every rank's row-local tables
__global__ void __launch_bounds__(128) k_shard_export(const ccj_seq *seqs, int table, int t, int ktiles, int16_t *out) {
    ccj_cx c;
    c.M = nullptr;
    c.q = seqs[0];
    const int n = c.q.n, m = n - t - 2;
    if (m < 1) return;
    const int a = blockIdx.y, b = t - a;
    const int ti = blockIdx.x / ktiles, tk = blockIdx.x % ktiles;
    const int i = 1 + ti * 4 + threadIdx.y, kk = tk * 32 + threadIdx.x;
    if (i > m || kk > m - i) return;
    const int j = i + a, k = j + 2 + kk, l = k + b;
    out[ccj_idx4(n, i, j, k, l)] = (int16_t)ccj_get4u(c, table, i, j, k, l);
}

