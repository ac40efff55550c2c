// sm_100a kernels of the CCJ fill + traceback (declarations of the host-side launchers).
// One DP "level" t = (j-i)+(l-k) of the 4D tables is one launch; see DESIGN.md for the schedule.
#pragma once
#include <cuda_runtime.h>

#include "ccj_types.h"

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges cost nothing unless a profiler is attached

// dynamic shared memory of a kernel; CCJ_HOST_EMU: the g++ build of the kernels for tests/emu/simt_emu.hpp (test tool only)
#ifdef CCJ_HOST_EMU
#define CCJ_DYN_SHARED(T, name) T *name = reinterpret_cast<T *>(simt::dyn_shared())
#else
#define CCJ_DYN_SHARED(T, name) extern __shared__ T name[]
#endif

namespace ccj {

// NVTX range around a phase of the fold (tracing); shows up in Nsight timelines
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

struct LaunchDims {
    int nseq;   // sequences in the wave (blockIdx.z / .y)
    int nmax;   // longest sequence of the wave
};

// sets the 2D tables to their "unset" values, W=0, pair=-1, status=0
void launch_init(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, cudaStream_t st);
// P(i,i+s) for all i, all sequences (src/pseudo_loop.cc:166-179)
void launch_P(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int s, cudaStream_t st);
// V, WBP, WPP, WB, WP, WMv, WMp, WM at span s
void launch_2d(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int s, cudaStream_t st);
// all 22 gap tables of level t
void launch_4d(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int t, cudaStream_t st);
// tuned level kernel (ccj_fill4.cu) + its per-sequence precomputation (e_stP table, window partner lists)
void launch_prep(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, cudaStream_t st);
void launch_prep_lists(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, cudaStream_t st);   // k_prep only
void launch_4d_tuned(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int t, cudaStream_t st);
// the three parts of launch_4d_tuned, for callers that overlap them on different streams:
// windows(t) only reads levels <= t-2, roles(t) levels <= t-1, final(t) needs both
void launch_4d_roles(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int t, cudaStream_t st);
void launch_4d_windows(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int t, cudaStream_t st);
void launch_4d_final(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int t, cudaStream_t st);
void launch_P_tuned(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int s, cudaStream_t st);
bool fill4_tuned_supported(int nmax);
int fill4_partials();  // int16 partial minima per cell in the per-level scratch
int fill4_fused_levels();  // KF: k_roles(t), t = 0 mod KF, covers levels t..t+KF-1 and needs the 2D tables up to span t+KF-2
// exterior W (src/W_final.cc:68-77)
void launch_W(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, cudaStream_t st);
// traceback, one warp per sequence
void launch_traceback(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, cudaStream_t st);

// one traceback node of sequence `seq` (node = i, j, k, l, type as stored on the traceback stack); *out_top = number of
// nodes it pushed (left in that sequence's tb_stack)
void launch_tb_step(const ccj_model *M, const ccj_seq *seqs, int seq, const int *node, int *out_top, cudaStream_t st);

// number of kernel launches the fill of a wave with longest sequence nmax issues (for gpu_launches)
int fill_launch_count(int nmax, bool tuned);

}  // namespace ccj
