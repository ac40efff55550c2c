// sm_100a kernels: v1 "generic index" version.  Every kernel is a thin scheduling shell around the
// cell functions of ccj_cells*.cuh / ccj_traceback.cuh (which carry the reference citations).
#include "ccj_kernels.cuh"
#include "ccj_traceback.cuh"
#include "ccj_cells4_lean.cuh"

#include <cstdlib>

namespace ccj {

// warp-cooperative candidate spreading (Par concept of ccj_cells.cuh)
struct WarpPar {
    static constexpr int nlanes = 32;
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ int red(int v) const { return __reduce_min_sync(0xffffffffu, v); }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ ccj_best argmin(ccj_best b) const {
        const int m = __reduce_min_sync(0xffffffffu, b.val);
        const int o = __reduce_min_sync(0xffffffffu, b.val == m ? b.ord : 0x7fffffff);
        ccj_best r;
        r.val = m;
        r.ord = o;
        return r;
    }
};

// block-cooperative candidate spreading for the traceback: TB_THREADS threads of one block work on one sequence.  Every
// thread follows the same control flow (all decisions come from reductions or from memory every thread reads), so the
// block-wide barriers inside red / argmin are reached uniformly.
#define TB_THREADS 128
struct BlockPar {
    static constexpr int nlanes = TB_THREADS;
    int *sm;   // 2 * (TB_THREADS / 32) ints of shared memory
    __device__ __forceinline__ int lane() const { return threadIdx.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ int red(int v) const {
        v = __reduce_min_sync(0xffffffffu, v);
        __syncthreads();   // the scratch may still be read by a slower warp of the previous reduction
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
        __syncthreads();
        int r = sm[0];
#pragma unroll
        for (int w = 1; w < TB_THREADS / 32; ++w) r = min(r, sm[w]);
        return r;
    }
    __device__ __forceinline__ ccj_best argmin(ccj_best b) const {   // lexicographic (value, position)
        const int m = red(b.val);
        const int o = red(b.val == m ? b.ord : 0x7fffffff);
        ccj_best r;
        r.val = m;
        r.ord = o;
        return r;
    }
};

__global__ void k_init(const ccj_model *M, const ccj_seq *seqs) {
    const ccj_seq q = seqs[blockIdx.y];
    const int64_t s2 = q.stride2;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t x = tid; x < s2; x += nth) {
        q.t2[T2_V * s2 + x] = CCJ_V_UNSET;
        q.t2[T2_VTYPE * s2 + x] = 'N';
#pragma unroll
        for (int t = T2_WM; t < CCJ_NT2; ++t) q.t2[t * s2 + x] = CCJ_INF + 1;
    }
    for (int64_t x = tid; x <= q.n + 1; x += nth) {
        if (x <= q.n) q.W[x] = 0;
        q.pair_out[x] = -1;
        q.ftype_out[x] = 'N';
    }
    for (int64_t x = tid; x < CCJ_STATUS_INTS; x += nth) q.status[x] = 0;
}

// P(i,l), l=i+s: blockIdx.x -> i, blockIdx.y -> j (first split point), threads -> (d,k)
__global__ void __launch_bounds__(256) k_P(const ccj_model *M, const ccj_seq *seqs, int s) {
    ccj_cx c;
    c.M = M;
    c.q = seqs[blockIdx.z];
    const int n = c.q.n;
    const int i = 1 + blockIdx.x, l = i + s;
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
        for (int x = 1; x < (int)(blockDim.x >> 5); ++x) mn = ccj_min(mn, sm[x]);
        // "if (min_energy < INF/2) P.set(i,l) = min_energy", table pre-set to INF+1
        if (mn < CCJ_INF / 2) atomicMin(&c.q.t2[T2_P * c.q.stride2 + ccj_idx2(n, i, l)], mn);
    }
}

// the same with ccj_P_lean (ccj_cells4_lean.cuh): warps take delta = k-d, lanes walk d (first factors consecutive in memory)
__global__ void __launch_bounds__(256) k_P_lean(const ccj_model *M, const ccj_seq *seqs, int s) {
    CCJ_DYN_SHARED(int64_t, s_tab);
    __shared__ int sm[8];
    const ccj_seq &q = seqs[blockIdx.z];
    const int n = q.n;
    const int i = 1 + blockIdx.x, l = i + s;
    if (l > n) return;
    const int j = i + blockIdx.y;
    if (j >= l) return;
    ccj_lean_plain_tab(s_tab, n, threadIdx.x, 256);
    __syncthreads();
    ccj_lean_plain ly;
    ly.t4 = q.t4; ly.st4 = q.stride4; ly.tab = s_tab; ly.n = n;
    int mn = ccj_P_lean(ly, i, j, l, threadIdx.x >> 5, 8, threadIdx.x & 31, 32);
    mn = __reduce_min_sync(0xffffffffu, mn);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mn;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int x = 1; x < 8; ++x) mn = ccj_min(mn, sm[x]);
        if (mn < CCJ_INF / 2) atomicMin(&q.t2[T2_P * q.stride2 + ccj_idx2(n, i, l)], mn);
    }
}

__global__ void __launch_bounds__(128) k_2d(const ccj_model *M, const ccj_seq *seqs, int s) {
    ccj_cx c;
    c.M = M;
    c.q = seqs[blockIdx.y];
    const int i = 1 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int j = i + s;
    if (j > c.q.n) return;
    WarpPar par;
    ccj_cell2d(c, i, j, par);
}

// level t: blockIdx.y -> a (b=t-a), blockIdx.z -> sequence, threads walk the packed (i,k) triangle of slab (a,b) in
// storage order (every lane has a cell, a warp's stores are contiguous)
// 12 blocks/SM: see k_4d_shard (ccj_shard.cu) for the measurement behind the occupancy choice
__global__ void __launch_bounds__(128, 12) k_4d(const ccj_model *M, const ccj_seq *seqs, int t) {
    ccj_cx c;
    c.M = M;
    c.q = seqs[blockIdx.z];
    const int n = c.q.n;
    const int m = n - t - 2;  // rows i = 1..m, row i has m+1-i cells
    if (m < 1) return;
    const int ncell = m * (m + 1) / 2;
    const int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= ncell) return;
    int q, kk;
    ccj_shard_cell_of(p, m, 1, m, q, kk);
    const int a = blockIdx.y, b = t - a;
    const int i = 1 + q, k = i + a + 2 + kk;
    ccj_cell4d(c, i, i + a, k, k + b);
}

// The same level with the lean cell function (ccj_cells4_lean.cuh): Cb / Tet of the ordinary layout from a table in
// shared memory, one position per split point.  Sequences without partner lists (CCJ_GENERIC_SCAN) take k_4d.
#ifndef K4D_LEAN_MINB
#define K4D_LEAN_MINB 8
#endif
__global__ void __launch_bounds__(128, K4D_LEAN_MINB) k_4d_lean(const ccj_model *M, const ccj_seq *seqs, int t, int nmax) {
    CCJ_DYN_SHARED(int64_t, s_tab);   // per-block copy for the block's sequence: Cb(0..n), Tet(0..n)
    ccj_cx c;
    c.M = M;
    c.q = seqs[blockIdx.z];
    const int n = c.q.n;
    const int m = n - t - 2;  // rows i = 1..m, row i has m+1-i cells
    if (m < 1 || n > nmax) return;
    const int ncell = m * (m + 1) / 2;
    if ((int)(blockIdx.x * 128) >= ncell) return;
    ccj_lean_plain_tab(s_tab, n, threadIdx.x, 128);
    __syncthreads();
    const int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= ncell) return;
    int q, kk;
    ccj_shard_cell_of(p, m, 1, m, q, kk);
    const int a = blockIdx.y, b = t - a;
    const int i = 1 + q, k = i + a + 2 + kk;
    if (!ccj_lists_ok(c) || !c.q.w3) {   // uniform per sequence
        ccj_cell4d(c, i, i + a, k, k + b);
        return;
    }
    ccj_lean_plain ly;
    ly.t4 = c.q.t4; ly.st4 = c.q.stride4; ly.tab = s_tab; ly.n = n;
    ccj_cell4d_lean(c, ly, i, i + a, k, k + b);
}

__global__ void k_W(const ccj_model *M, const ccj_seq *seqs) {
    ccj_cx c;
    c.M = M;
    c.q = seqs[blockIdx.x];
    WarpPar par;
    for (int j = CCJ_TURN + 1; j <= c.q.n; ++j) {
        const int w = ccj_W_at(c, j, par);
        if (par.lane() == 0) c.q.W[j] = w;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(TB_THREADS) k_traceback(const ccj_model *M, const ccj_seq *seqs) {
    __shared__ int sm[2 * (TB_THREADS / 32)];
    ccj_cx c;
    c.M = M;
    c.q = seqs[blockIdx.x];
    BlockPar par;
    par.sm = sm;
    ccj_traceback(c, par);
}

// one node of the traceback (pseudo_loop::backtrack / W_final::backtrack process one interval per call): the
// nodes it pushes are left in the sequence's traceback stack [0, *out_top), pair/type updates in pair_out/ftype_out
__global__ void k_tb_step(const ccj_model *M, const ccj_seq *seqs, int seq, int a0, int a1, int a2, int a3, int ty, int *out_top) {
    ccj_cx c;
    c.M = M;
    c.q = seqs[seq];
    WarpPar par;
    ccj_tb T(c);
    ccj_tb_node(T, par, a0, a1, a2, a3, ty);
    par.sync();
    if (par.lane() == 0) *out_top = T.top;
}

#ifndef CCJ_HOST_EMU   // the launchers below are CUDA only; the emulation harness issues the same grids itself
// ---------------------------------------------------------------------------------------------
void launch_tb_step(const ccj_model *M, const ccj_seq *seqs, int seq, const int *node, int *out_top, cudaStream_t st) {
    k_tb_step<<<1, 32, 0, st>>>(M, seqs, seq, node[0], node[1], node[2], node[3], node[4], out_top);
}

// CCJ_K4D_LEAN=0 selects the generic-index kernels (k_4d, k_P) for comparison
static bool lean_enabled() {
    static const bool lean = [] { const char *e = getenv("CCJ_K4D_LEAN"); return !(e && e[0] == '0'); }();
    return lean;
}

void launch_init(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, cudaStream_t st) {
    const int64_t s2 = ccj_stride2(d.nmax);
    int bx = (int)((s2 + 255) / 256);
    if (bx > 1024) bx = 1024;
    k_init<<<dim3(bx, d.nseq), 256, 0, st>>>(M, seqs);
}

void launch_P(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int s, cudaStream_t st) {
    if (s < 3 || s > d.nmax - 1) return;  // needs i<=j<d<k<l
    const size_t smem = 2 * (size_t)(d.nmax + 1) * sizeof(int64_t);
    if (lean_enabled() && smem <= 40000)
        k_P_lean<<<dim3(d.nmax - s, s, d.nseq), 256, smem, st>>>(M, seqs, s);
    else
        k_P<<<dim3(d.nmax - s, s, d.nseq), 256, 0, st>>>(M, seqs, s);
}

void launch_2d(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int s, cudaStream_t st) {
    const int rows = d.nmax - s;
    if (rows < 1) return;
    k_2d<<<dim3((rows + 3) / 4, d.nseq), 128, 0, st>>>(M, seqs, s);
}

void launch_4d(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int t, cudaStream_t st) {
    const int m = d.nmax - t - 2;
    if (m < 1) return;
    const int ncell = m * (m + 1) / 2;
    const size_t smem = 2 * (size_t)(d.nmax + 1) * sizeof(int64_t);
    if (lean_enabled() && smem <= 40000)
        k_4d_lean<<<dim3((ncell + 127) / 128, t + 1, d.nseq), 128, smem, st>>>(M, seqs, t, d.nmax);
    else
        k_4d<<<dim3((ncell + 127) / 128, t + 1, d.nseq), 128, 0, st>>>(M, seqs, t);
}

void launch_W(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, cudaStream_t st) {
    k_W<<<d.nseq, 32, 0, st>>>(M, seqs);
}

void launch_traceback(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, cudaStream_t st) {
    k_traceback<<<d.nseq, TB_THREADS, 0, st>>>(M, seqs);
}

int fill_launch_count(int nmax, bool tuned) {
    int c = tuned ? 5 : 3;  // init (+ layout prep + PMW fill) + list prep + W
    for (int s = 0; s < nmax; ++s) {
        if (s >= 3 && s <= nmax - 1) ++c;           // K_P
        ++c;                                        // K_2D
        // level: split roles (every KF-th level; they cover KF), PL/PR windows, PM window (if any pair is short enough), assembly
        if (nmax - s - 2 >= 1) c += tuned ? 2 + (s % fill4_fused_levels() == 0 ? 1 : 0) + (nmax - 1 - s > CCJ_TURN ? 1 : 0) : 1;
    }
    return c;
}
#endif  // CCJ_HOST_EMU

}  // namespace ccj
