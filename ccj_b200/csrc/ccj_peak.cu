// Measurement helper: the integer add-min ceiling of this GPU ("INT32 peak must be measured on the
// box with an IADD3/IMNMX micro-kernel").  One min-plus candidate of the fill is one add and one min; sm_100a
// fuses them into VIADDMNMX (and VIADDMNMX.S16x2 for two int16 cells at once), so the ceiling is reported as
// candidates ("add-min pairs") per second for the three instruction forms the fill kernels use.
// Not on the fold path: called by bench.py for the roofline denominators.
#include <cuda_runtime.h>

#include <string>

#include "../../include/ccj_b200.h"

namespace {

#define PEAK_ACC 8      // independent dependency chains per thread
#define PEAK_UNROLL 64  // add-mins per chain and loop trip

// variant 0: min(acc + w, x) in int32            -> VIADDMNMX
// variant 1: the form k_roles uses on a packed int16 record: sign-extend one half, then add-min -> PRMT/SGXT + VIADDMNMX
// variant 2: packed int16x2 (two cells per instruction), add.s16x2 + min.s16x2 -> VIADDMNMX.S16x2 or the pair
template <int VARIANT>
__global__ void __launch_bounds__(256) k_peak(const int *__restrict__ in, int *__restrict__ out, int iters) {
    int acc[PEAK_ACC], x[PEAK_ACC];
    const int w = in[threadIdx.x & 31];
#pragma unroll
    for (int u = 0; u < PEAK_ACC; ++u) {
        acc[u] = in[32 + ((threadIdx.x + u) & 63)];
        x[u] = in[96 + ((threadIdx.x * 3 + u) & 63)];
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < PEAK_UNROLL; ++r) {
#pragma unroll
            for (int u = 0; u < PEAK_ACC; ++u) {
                if (VARIANT == 0) {
                    acc[u] = min(acc[u] + w, x[u]);
                } else if (VARIANT == 1) {
                    acc[u] = min((int)(int16_t)(acc[u] & 0xffff) + w, x[u]);
                } else {
                    int s;
                    asm("add.s16x2 %0, %1, %2;" : "=r"(s) : "r"(acc[u]), "r"(w));
                    asm("min.s16x2 %0, %1, %2;" : "=r"(acc[u]) : "r"(s), "r"(x[u]));
                }
            }
        }
    }
    int r = 0;
#pragma unroll
    for (int u = 0; u < PEAK_ACC; ++u) r ^= acc[u];
    if (r == 0x13572468) out[blockIdx.x * blockDim.x + threadIdx.x] = r;  // practically never: keeps the chains alive
}

}  // namespace

extern "C" int ccj_measure_addmin_peak(ccj_ctx *ctx, int variant, double *pairs_per_s) {
    if (!ctx || !pairs_per_s || variant < 0 || variant > 2) return CCJ_ERR_ARG;
    cudaStream_t st = (cudaStream_t)ccj_stream(ctx);
    cudaDeviceProp prop;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return CCJ_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 256;
    int *d_in = nullptr, *d_out = nullptr;
    if (cudaMalloc((void **)&d_in, 256 * sizeof(int)) != cudaSuccess) return CCJ_ERR_CUDA;
    if (cudaMalloc((void **)&d_out, (size_t)blocks * threads * sizeof(int)) != cudaSuccess) {
        cudaFree(d_in);
        return CCJ_ERR_CUDA;
    }
    int h_in[256];
    for (int x = 0; x < 256; ++x) h_in[x] = variant == 2 ? ((x * 37 - 1000) & 0x7fff) | (((x * 53 - 900) & 0x7fff) << 16) : x * 7919 - 100000;
    for (int x = 0; x < 32; ++x) h_in[x] = -1;   // int32: -1; int16x2: (-1, -1)  // w = -1: the chains keep moving
    cudaMemcpyAsync(d_in, h_in, sizeof h_in, cudaMemcpyHostToDevice, st);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {  // first repetition = warm-up
        cudaEventRecord(e0, st);
        if (variant == 0) k_peak<0><<<blocks, threads, 0, st>>>(d_in, d_out, iters);
        else if (variant == 1) k_peak<1><<<blocks, threads, 0, st>>>(d_in, d_out, iters);
        else k_peak<2><<<blocks, threads, 0, st>>>(d_in, d_out, iters);
        cudaEventRecord(e1, st);
        cudaStreamSynchronize(st);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const cudaError_t err = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_in);
    cudaFree(d_out);
    if (err != cudaSuccess) return CCJ_ERR_CUDA;
    const double instr = (double)blocks * threads * iters * PEAK_UNROLL * PEAK_ACC;
    *pairs_per_s = instr * (variant == 2 ? 2.0 : 1.0) / (best * 1e-3);
    return 0;
}
