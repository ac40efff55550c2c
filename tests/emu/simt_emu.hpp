// TEST TOOL (not part of the product): a small SIMT emulator that lets g++ compile and run the CUDA kernels of
// ccj_b200/csrc/ccj_fill4.cu on the host -- one OS thread per CUDA thread, one thread block at a time, real barriers for
// __syncthreads / __syncwarp, warp collectives (__ballot_sync, __reduce_*_sync) over the live lanes of a warp.
// It exists so that the tuned kernels can be checked against the reference's golden vectors without a GPU and can run
// under AddressSanitizer / UBSan / ThreadSanitizer (compute-sanitizer is closed on the GPU pool): every buffer is an
// exactly sized heap block, __shared__ arrays are real shared storage, and two threads of a block that touch the same
// word without a barrier in between are two OS threads that race.
// Semantics kept: threads that return early no longer take part in later barriers; __shared__ contents survive from
// block to block (uninitialised, as on the device); kernels of one stream run one after the other.
#pragma once
#include <cuda_runtime.h>   // vector types, dim3, make_int4, empty __global__/__device__ for a host compiler

#include <atomic>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)

namespace simt {

struct Barrier {   // live-count barrier: an exiting thread lowers the count instead of arriving
    std::mutex m;                     // guards the counters only; waiters sleep on `gen` (futex), not on the mutex
    int live = 0, waiting = 0;
    std::atomic<unsigned> gen{0};
    void reset(int n) { live = n; waiting = 0; }
    void arrive() {
        m.lock();
        const unsigned my = gen.load(std::memory_order_relaxed);
        if (++waiting == live) {
            waiting = 0;
            gen.store(my + 1, std::memory_order_release);
            m.unlock();
            gen.notify_all();
            return;
        }
        m.unlock();
        while (gen.load(std::memory_order_acquire) == my) gen.wait(my, std::memory_order_acquire);
    }
    void leave() {
        m.lock();
        --live;
        if (live > 0 && waiting == live) {
            waiting = 0;
            gen.fetch_add(1, std::memory_order_release);
            m.unlock();
            gen.notify_all();
            return;
        }
        m.unlock();
    }
};

struct Warp {
    Barrier bar;
    int val[32];
    bool alive[32];
};

struct Block {
    Barrier bar;
    std::vector<Warp> warps;
};

inline Block *&cur_block() { static Block *b = nullptr; return b; }
// dynamic shared memory of the running launch (CCJ_DYN_SHARED in ccj_kernels.cuh)
inline std::vector<long long> &dyn_store() { static std::vector<long long> v; return v; }
inline void *dyn_shared() { return dyn_store().data(); }

}  // namespace simt

// ---- the CUDA built-ins the kernels use --------------------------------------------------------------------------------
inline thread_local uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

inline void __syncthreads() { simt::cur_block()->bar.arrive(); }
inline simt::Warp &simt_warp() { return simt::cur_block()->warps[threadIdx.x >> 5]; }
// SIMT_EMU_DROP_SYNCWARP=1 in the environment: the kernels' __syncwarp() do nothing -- the negative control of the race
// detector (tests/test_emu_tuned.py)
inline bool simt_drop_syncwarp() {
    static const bool drop = [] { const char *e = getenv("SIMT_EMU_DROP_SYNCWARP"); return e && e[0] == '1'; }();
    return drop;
}
inline void __syncwarp(unsigned = 0xffffffffu) {
    if (!simt_drop_syncwarp()) simt_warp().bar.arrive();
}
template <class F>
inline int simt_collective(int v, F fold, int init) {
    simt::Warp &w = simt_warp();
    const int lane = threadIdx.x & 31;
    w.val[lane] = v;
    w.bar.arrive();
    int r = init;
    for (int l = 0; l < 32; ++l)
        if (w.alive[l]) r = fold(r, w.val[l], l);
    w.bar.arrive();   // nobody overwrites val[] before every lane has read it
    return r;
}
inline int __reduce_max_sync(unsigned, int v) { return simt_collective(v, [](int a, int b, int) { return a > b ? a : b; }, INT_MIN); }
inline int __reduce_min_sync(unsigned, int v) { return simt_collective(v, [](int a, int b, int) { return a < b ? a : b; }, INT_MAX); }
inline unsigned __ballot_sync(unsigned, bool p) {
    return (unsigned)simt_collective(p ? 1 : 0, [](int a, int b, int l) { return (int)((unsigned)a | ((unsigned)(b & 1) << l)); }, 0);
}
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const uint64_t src = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int b = 0; b < 4; ++b) r |= (unsigned)((src >> (8 * ((s >> (4 * b)) & 7))) & 0xff) << (8 * b);
    return r;
}
template <class T> inline T __ldg(const T *p) { return *p; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline int atomicMin(int *p, int v) {   // blocks run one after the other and one thread per block calls it
    const int old = *p;
    if (v < old) *p = v;
    return old;
}

namespace simt {

// runs kernel(args...) for every block of `grid`, `block.x` threads each (1-D blocks are all the kernels use).
// The block's threads are created once per launch and walk the blocks together: a full barrier separates two blocks
// (the __shared__ arrays and the barrier state are per block).
template <class K, class... A>
void launch_smem(K kernel, dim3 grid, dim3 block, size_t smem_bytes, A... args) {
    dyn_store().assign(smem_bytes / sizeof(long long) + 2, 0x5555555555555555LL);
    gridDim = grid;
    blockDim = block;
    const int T = (int)block.x, nwarp = (T + 31) / 32;
    Block blk;
    blk.warps = std::vector<Warp>(nwarp);
    cur_block() = &blk;
    Barrier phase;
    phase.reset(T);
    auto reset_block = [&] {
        blk.bar.reset(T);
        for (int w = 0; w < nwarp; ++w) {
            const int lanes = std::min(32, T - 32 * w);
            blk.warps[w].bar.reset(lanes);
            for (int l = 0; l < 32; ++l) blk.warps[w].alive[l] = l < lanes;
        }
    };
    reset_block();
    std::vector<std::thread> th;
    th.reserve(T);
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t] {
            threadIdx = make_uint3((unsigned)t, 0, 0);
            bool first = true;
            for (unsigned bz = 0; bz < grid.z; ++bz)
                for (unsigned by = 0; by < grid.y; ++by)
                    for (unsigned bx = 0; bx < grid.x; ++bx) {
                        if (!first) {
                            phase.arrive();              // every thread is done with the previous block
                            if (t == 0) reset_block();
                            phase.arrive();
                        }
                        first = false;
                        blockIdx = make_uint3(bx, by, bz);
                        kernel(args...);
                        Warp &w = blk.warps[t >> 5];
                        {   // an exited lane is out of every later barrier and collective of this block
                            std::unique_lock<std::mutex> g(w.bar.m);
                            w.alive[t & 31] = false;
                        }
                        w.bar.leave();
                        blk.bar.leave();
                    }
        });
    for (auto &x : th) x.join();
    cur_block() = nullptr;
}
template <class K, class... A>
void launch(K kernel, dim3 grid, dim3 block, A... args) {
    launch_smem(kernel, grid, block, 0, args...);
}

}  // namespace simt
