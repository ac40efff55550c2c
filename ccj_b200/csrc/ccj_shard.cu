// Row-sharded fold of ONE sequence whose gap tables exceed one GPU (BASELINE config 5; SURVEY.md 8e;
// the loops being sharded: pseudo_loop::compute_energies, src/pseudo_loop.cc:69-132, driven by W_final::ccj,
// src/W_final.cc:58-77).  One process per GPU ("rank"); row i of every gap table belongs to rank (i-1) mod G.
//
// Per DP step s (span of the 2D tables = level of the gap tables; see DESIGN.md "schedule"):
//   k_P_shard(s)      P(i,i+s) for the rank's rows i, from PK of levels <= s-3            (every rank holds all of PK)
//   allreduce-min     the span-s diagonal of P (n+1 int32, contiguous in the diagonal-major 2D layout)
//   k_2d(s)           V / WBP / WPP / WB / WP / WMv / WMp / WM of span s, every rank        (replicated, cheap)
//   k_4d_shard(s)     all 22 gap tables of the rank's cells of level s: the product's cell function, table access through
//                     the sharded layout of ccj_types.h (k_4d_shard_lean / k_P_shard_lean: the lean form of
//                     ccj_cells4_lean.cuh, one position per split point; k_4d_shard / k_P_shard: ccj_cell4d as it stands)
//   allgather         the 12 column-read tables of level s: one in-place ncclAllGather of G adjacent blocks
// then k_W everywhere and the traceback on the rank that opened its peers' memory (the 10 row-local tables of the
// other ranks are read over NVLink through CUDA IPC pointers; the traceback touches O(n) cells per node).
//
// Collectives: NCCL across processes (loaded at run time with dlopen, so that the library itself has no link-time
// dependency), or -- for tests on a single GPU -- an in-process group of G shards on one device whose "collectives"
// are device-to-device copies; both go through the same fill loop.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ccj_b200.h"
#include "ccj_cells4_lean.cuh"
#include "ccj_kernels.cuh"
#include "ccj_render.hpp"
#include "energy_model.hpp"

extern "C" const void *ccj_internal_device_model(ccj_ctx *ctx);   // ccj_abi.cu
extern "C" const void *ccj_internal_host_model(ccj_ctx *ctx);
extern "C" int ccj_internal_device(ccj_ctx *ctx);

namespace {

// ---- NCCL, resolved at run time ---------------------------------------------------------------------------------------
struct Nccl {
    void *lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    bool ok = false;
    std::string err;
};
Nccl &nccl() {
    static Nccl n;
    if (n.lib || !n.err.empty()) return n;
    // NCCL writes its debug lines (NCCL_DEBUG=VERSION/INFO) to stdout by default; the stdout of a fold is the reference's
    // output format and is parsed by callers, so they go to stderr unless the user chose a file
    setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
    // a process that already holds a libnccl (torch bundles its own) must keep using THAT one: loading a second copy
    // under the same soname would shadow it for later imports
    n.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!n.lib && getenv("CCJ_NCCL_LIB")) n.lib = dlopen(getenv("CCJ_NCCL_LIB"), RTLD_NOW | RTLD_LOCAL);
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        if (n.lib) break;
        n.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    }
    if (!n.lib) {
        n.err = "libnccl.so.2 not found";
        return n;
    }
#define SYM(f) n.f = reinterpret_cast<decltype(n.f)>(dlsym(n.lib, "nccl" #f))
    SYM(GetUniqueId); SYM(CommInitRank); SYM(CommDestroy); SYM(AllGather); SYM(AllReduce); SYM(GetErrorString);
    SYM(CommInitAll); SYM(GroupStart); SYM(GroupEnd);
#undef SYM
    n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllGather && n.AllReduce && n.GetErrorString &&
           n.CommInitAll && n.GroupStart && n.GroupEnd;
    if (!n.ok) n.err = "libnccl lacks a required symbol";
    return n;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

struct ccj_shard {
    ccj_ctx *ctx = nullptr;
    int rank = 0, world = 1;
    int device = 0;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    // current sequence
    int n = 0;
    std::string seq;
    char *arena = nullptr;
    size_t arena_bytes = 0;
    size_t off_w3 = 0, off_estp = 0, off_inlist = 0, off_outlist = 0, off_incnt = 0, off_outcnt = 0;
    size_t off_in = 0, off_out = 0, off_t2 = 0, off_ftype = 0, off_tb = 0, off_lev = 0, off_locptr = 0, off_rep = 0, off_loc = 0, off_desc = 0;
    std::vector<int64_t> lev;            // n+2 entries
    ccj_seq h_desc;
    std::vector<void *> opened;          // IPC mappings of the peers' arenas
    std::vector<int16_t *> h_locptr;     // [world]
    bool prepared = false, filled = false, peers = false;
    float ms[4] = {0, 0, 0, 0};          // total, compute, allgather, allreduce
    std::vector<float> level_ms;         // per step s: P kernel, allreduce, 2D + gap-table kernels, allgather
};

namespace {

int sfail(ccj_shard *s, int code, const std::string &msg) {
    if (s) s->err = msg;
    return code;
}
#define SCU(call)                                                                                \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) return sfail(sh, CCJ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

ccj_seq *d_desc(ccj_shard *sh) { return reinterpret_cast<ccj_seq *>(sh->arena + sh->off_desc); }

#include "ccj_shard_kernels.cuh"   // k_4d_shard(_lean), k_P_shard(_lean), shard_lean_ok, k_min_into, k_shard_export

void fnv_add(uint64_t &h, uint64_t v) {
    h ^= v;
    h *= 1099511628211ULL;
}

}  // namespace

extern "C" {

int ccj_shard_unique_id(void *id, size_t bytes) {
    if (!id || bytes < sizeof(ncclUniqueId)) return CCJ_ERR_ARG;
    Nccl &N = nccl();
    if (!N.ok) return CCJ_ERR_STATE;
    ncclUniqueId u;
    if (N.GetUniqueId(&u) != ncclSuccess) return CCJ_ERR_CUDA;
    memcpy(id, &u, sizeof u);
    return 0;
}

size_t ccj_shard_unique_id_bytes(void) { return sizeof(ncclUniqueId); }

int ccj_shard_create(ccj_ctx *ctx, int rank, int world, const void *unique_id, ccj_shard **out) {
    if (!ctx || !out || world < 1 || rank < 0 || rank >= world) return CCJ_ERR_ARG;
    *out = nullptr;
    ccj_shard *sh = new ccj_shard();
    sh->ctx = ctx;
    sh->rank = rank;
    sh->world = world;
    sh->stream = (cudaStream_t)ccj_stream(ctx);
    sh->device = ccj_internal_device(ctx);
    if (unique_id && world > 1) {
        Nccl &N = nccl();
        if (!N.ok) {
            delete sh;
            return CCJ_ERR_STATE;
        }
        ncclUniqueId u;
        memcpy(&u, unique_id, sizeof u);
        const ncclResult_t r = N.CommInitRank(&sh->comm, world, u, rank);
        if (r != ncclSuccess) {
            fprintf(stderr, "ccj_b200: ncclCommInitRank: %s\n", N.GetErrorString(r));
            delete sh;
            return CCJ_ERR_CUDA;
        }
        // NCCL connects its channels lazily inside the first collective of each kind (hundreds of ms): do that here,
        // not inside the first level of a fill
        int32_t *warm = nullptr;
        if (cudaMalloc((void **)&warm, sizeof(int32_t) * 64 * (size_t)world) == cudaSuccess) {
            cudaMemsetAsync(warm, 0, sizeof(int32_t) * 64 * (size_t)world, sh->stream);
            N.AllReduce(warm, warm, 64, ncclInt32, ncclMin, sh->comm, sh->stream);
            N.AllGather(reinterpret_cast<char *>(warm) + 256 * (size_t)rank, warm, 256, ncclInt8, sh->comm, sh->stream);
            cudaStreamSynchronize(sh->stream);
            cudaFree(warm);
        }
    }
    *out = sh;
    return 0;
}

// All ranks inside ONE process, one GPU each (the command line / a host program without a launcher): `world` contexts on
// distinct devices, one NCCL communicator per rank from ncclCommInitAll, peer access enabled between the devices so that
// the traceback rank can read the other ranks' row-local tables directly.
int ccj_shard_create_group(ccj_ctx **ctxs, int world, ccj_shard **out) {
    if (!ctxs || !out || world < 1) return CCJ_ERR_ARG;
    for (int r = 0; r < world; ++r) out[r] = nullptr;
    std::vector<int> devs(world);
    for (int r = 0; r < world; ++r) {
        if (!ctxs[r]) return CCJ_ERR_ARG;
        devs[r] = ccj_internal_device(ctxs[r]);
        for (int q = 0; q < r; ++q)
            if (devs[q] == devs[r]) return CCJ_ERR_ARG;   // one device per rank
    }
    std::vector<ncclComm_t> comms(world, nullptr);
    if (world > 1) {
        Nccl &N = nccl();
        if (!N.ok) return CCJ_ERR_STATE;
        const ncclResult_t r = N.CommInitAll(comms.data(), world, devs.data());
        if (r != ncclSuccess) {
            fprintf(stderr, "ccj_b200: ncclCommInitAll: %s\n", N.GetErrorString(r));
            return CCJ_ERR_CUDA;
        }
        for (int a = 0; a < world; ++a)
            for (int b = 0; b < world; ++b)
                if (a != b) {
                    cudaSetDevice(devs[a]);
                    const cudaError_t e = cudaDeviceEnablePeerAccess(devs[b], 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                        fprintf(stderr, "ccj_b200: no peer access %d -> %d: %s\n", devs[a], devs[b], cudaGetErrorString(e));
                        cudaGetLastError();
                        for (ncclComm_t c : comms) if (c) N.CommDestroy(c);
                        return CCJ_ERR_CUDA;
                    }
                    cudaGetLastError();
                }
    }
    for (int r = 0; r < world; ++r) {
        ccj_shard *sh = new ccj_shard();
        sh->ctx = ctxs[r];
        sh->rank = r;
        sh->world = world;
        sh->device = devs[r];
        sh->stream = (cudaStream_t)ccj_stream(ctxs[r]);
        sh->comm = comms[r];
        out[r] = sh;
    }
    return 0;
}

// One sequence over every GPU the caller hands in: prepare, fill (ccj_shard_fill as an in-process NCCL group),
// traceback on rank 0.  ms4 as ccj_shard_fill.  The contexts must hold the same energy model.
int ccj_shard_fold(ccj_ctx **ctxs, int nctx, const char *seq, int n, ccj_result *result, int32_t *pairs, char *structs,
                   float *ms4) {
    if (!ctxs || nctx < 1 || !seq || n < 1 || !result) return CCJ_ERR_ARG;
    std::vector<ccj_shard *> sh(nctx, nullptr);
    int rc = ccj_shard_create_group(ctxs, nctx, sh.data());
    if (rc) return rc;
    for (int r = 0; r < nctx && !rc; ++r) {
        cudaSetDevice(sh[r]->device);
        rc = ccj_shard_prepare(sh[r], seq, n);
        if (rc) fprintf(stderr, "ccj_b200: %s\n", sh[r]->err.c_str());
    }
    if (!rc) rc = ccj_shard_fill(sh.data(), nctx, ms4);
    if (!rc) {
        cudaSetDevice(sh[0]->device);
        rc = ccj_shard_link_local(sh[0], sh.data(), nctx);
    }
    if (!rc) rc = ccj_shard_traceback(sh[0], result, pairs, structs, nullptr);
    if (rc && sh[0] && !sh[0]->err.empty()) fprintf(stderr, "ccj_b200: %s\n", sh[0]->err.c_str());
    for (ccj_shard *z : sh)
        if (z) {
            cudaSetDevice(z->device);
            ccj_shard_destroy(z);
        }
    return rc;
}

void ccj_shard_destroy(ccj_shard *sh) {
    if (!sh) return;
    for (void *p : sh->opened)
        if (p) cudaIpcCloseMemHandle(p);
    if (sh->arena) cudaFree(sh->arena);
    if (sh->comm) nccl().CommDestroy(sh->comm);
    delete sh;
}

const char *ccj_shard_last_error(const ccj_shard *sh) { return sh ? sh->err.c_str() : "no shard"; }

// bytes of one rank's arena for a length-n sequence dealt to `world` ranks (host only)
int64_t ccj_shard_bytes(int n, int world) {
    if (n < 1 || world < 1) return 0;
    int64_t cells = 0;
    for (int t = 0; t <= n - 3; ++t) cells += ccj_shard_level_cells(n, t, world);
    const int64_t tri = (int64_t)n * (n - 1) / 2 + 1;
    const int64_t small = (int64_t)ccj_stride2(n) * (CCJ_NT2 + 5) * 4 + tri * (CCJ_WIN_IN * 4 + CCJ_WIN_OUT * 8 + 8) + 64 * (int64_t)n + (1 << 16);
    return small + cells * 2 * ((int64_t)CCJ_SHARD_NREP * world + CCJ_SHARD_NLOC);
}

// storage position of (i,j,k,l) in the sharded layout (host only, for the CPU tests): returns the inner index, and
// owner rank, level, the level's reserved cells per table C(t) and lev[t]
int64_t ccj_shard_layout(int n, int world, int i, int j, int k, int l, int32_t *owner, int32_t *level, int64_t *level_cells,
                         int64_t *level_base) {
    if (n < 1 || world < 1 || i < 1 || l > n || !ccj_valid4(i, j, k, l)) return -1;
    const int t = (j - i) + (l - k);
    if (owner) *owner = (i - 1) % world;
    if (level) *level = t;
    if (level_cells) *level_cells = ccj_shard_level_cells(n, t, world);
    if (level_base) {
        int64_t acc = 0;
        for (int x = 0; x < t; ++x) acc += ccj_shard_level_cells(n, x, world);
        *level_base = acc;
    }
    return ccj_shard_inner(n, world, i, j, k, l);
}

int ccj_shard_prepare(ccj_shard *sh, const char *seq, int n) {
    if (!sh || !seq || n < 1) return CCJ_ERR_ARG;
    for (int x = 0; x < n; ++x) {
        const char ch = seq[x];
        if (!(ch == 'G' || ch == 'C' || ch == 'A' || ch == 'U' || ch == 'T')) return sfail(sh, CCJ_ERR_SEQUENCE, "sequence has a character outside GCAUT");
    }
    if (n > CCJ_HAIRPIN_TAB - 2) return sfail(sh, CCJ_ERR_TOO_LARGE, "sequence longer than supported");
    const int G = sh->world;
    sh->prepared = sh->filled = false;
    sh->n = n;
    sh->seq.assign(seq, (size_t)n);
    sh->lev.assign(n + 2, 0);
    for (int t = 0; t <= n; ++t) sh->lev[t + 1] = sh->lev[t] + ccj_shard_level_cells(n, t, G);
    const int64_t cells = sh->lev[n + 1];
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
    sh->off_desc = take(sizeof(ccj_seq));
    sh->off_in = take((size_t)(n + 2) + (size_t)n + 2);
    sh->off_out = take(sizeof(int32_t) * (CCJ_STATUS_INTS + (size_t)(n + 1) + (size_t)(n + 2)));
    sh->off_t2 = take((size_t)ccj_stride2(n) * CCJ_NT2 * sizeof(int32_t));
    sh->off_w3 = take((size_t)ccj_stride2(n) * 4 * sizeof(int32_t));
    const size_t tri = (size_t)n * (n - 1) / 2 + 1;   // per-pair partner lists of the interior windows (k_prep)
    sh->off_estp = take((size_t)ccj_stride2(n) * sizeof(int32_t));
    sh->off_inlist = take(tri * CCJ_WIN_IN * sizeof(uint32_t));
    sh->off_outlist = take(tri * CCJ_WIN_OUT * 2 * sizeof(uint32_t));
    sh->off_incnt = take(tri * sizeof(int32_t));
    sh->off_outcnt = take(tri * sizeof(int32_t));
    sh->off_ftype = take((size_t)n + 2);
    sh->off_tb = take(sizeof(int32_t) * 5 * (size_t)(16 * n + 64));
    sh->off_lev = take(sizeof(int64_t) * (size_t)(n + 2));
    sh->off_locptr = take(sizeof(void *) * (size_t)G);
    sh->off_loc = take((size_t)cells * CCJ_SHARD_NLOC * sizeof(int16_t) + 64);   // same offset on every rank
    sh->off_rep = take((size_t)cells * CCJ_SHARD_NREP * (size_t)G * sizeof(int16_t) + 64);
    size_t fr = 0, tot = 0;
    SCU(cudaMemGetInfo(&fr, &tot));
    if (off > sh->arena_bytes) {
        for (void *p : sh->opened)
            if (p) cudaIpcCloseMemHandle(p);
        sh->opened.clear();
        sh->peers = false;
        if (sh->arena) SCU(cudaFree(sh->arena));
        sh->arena = nullptr;
        sh->arena_bytes = 0;
        SCU(cudaMemGetInfo(&fr, &tot));
        if (off + ((size_t)1 << 30) > fr) return sfail(sh, CCJ_ERR_TOO_LARGE, "the rank's share of the tables does not fit this GPU");
        SCU(cudaMalloc((void **)&sh->arena, off));
        sh->arena_bytes = off;
    }
    // inputs
    std::vector<char> in((size_t)(n + 2) + (size_t)n + 2, 0);
    for (int x = 1; x <= n; ++x) in[x] = (char)ccj::encode_base(seq[x - 1]);
    in[n + 1] = in[1];
    in[0] = in[n];
    memcpy(in.data() + (n + 2), seq, (size_t)n);
    SCU(cudaMemcpyAsync(sh->arena + sh->off_in, in.data(), in.size(), cudaMemcpyHostToDevice, sh->stream));
    SCU(cudaMemcpyAsync(sh->arena + sh->off_lev, sh->lev.data(), sizeof(int64_t) * (size_t)(n + 2), cudaMemcpyHostToDevice, sh->stream));
    sh->h_locptr.assign(G, nullptr);
    sh->h_locptr[sh->rank] = reinterpret_cast<int16_t *>(sh->arena + sh->off_loc);
    SCU(cudaMemcpyAsync(sh->arena + sh->off_locptr, sh->h_locptr.data(), sizeof(void *) * (size_t)G, cudaMemcpyHostToDevice, sh->stream));
    ccj_seq &q = sh->h_desc;
    memset(&q, 0, sizeof q);
    q.n = n;
    q.S = reinterpret_cast<const int8_t *>(sh->arena + sh->off_in);
    q.seq = sh->arena + sh->off_in + (n + 2);
    q.status = reinterpret_cast<int32_t *>(sh->arena + sh->off_out);
    q.W = q.status + CCJ_STATUS_INTS;
    q.pair_out = q.W + (n + 1);
    q.t2 = reinterpret_cast<int32_t *>(sh->arena + sh->off_t2);
    q.stride2 = ccj_stride2(n);
    q.stride4 = 0;
    q.w3 = reinterpret_cast<int32_t *>(sh->arena + sh->off_w3);   // {WB,WP,WBP} packed per interval: one load per split point
    q.estP = reinterpret_cast<int32_t *>(sh->arena + sh->off_estp);
    q.inlist = reinterpret_cast<uint32_t *>(sh->arena + sh->off_inlist);
    q.outlist = reinterpret_cast<uint32_t *>(sh->arena + sh->off_outlist);
    q.incnt = reinterpret_cast<int32_t *>(sh->arena + sh->off_incnt);
    q.outcnt = reinterpret_cast<int32_t *>(sh->arena + sh->off_outcnt);
    { const char *e = getenv("CCJ_GENERIC_SCAN"); q.use_lists = (e && e[0] == '1') ? 0 : 1; }
    q.ftype_out = reinterpret_cast<int8_t *>(sh->arena + sh->off_ftype);
    q.tb_stack = reinterpret_cast<int32_t *>(sh->arena + sh->off_tb);
    q.tb_cap = 16 * n + 64;
    q.shard_G = G;
    q.shard_rank = sh->rank;
    q.shard_shift = -1;
    for (int b = 0; b < 16; ++b)
        if ((1 << b) == G) q.shard_shift = b;
    q.shard_lev = reinterpret_cast<const int64_t *>(sh->arena + sh->off_lev);
    q.shard_rep = reinterpret_cast<int16_t *>(sh->arena + sh->off_rep);
    q.shard_loc = reinterpret_cast<int16_t *const *>(sh->arena + sh->off_locptr);
    ccj_shard_kinds(q.shard_kind);
    SCU(cudaMemcpyAsync(d_desc(sh), &q, sizeof q, cudaMemcpyHostToDevice, sh->stream));
    SCU(cudaStreamSynchronize(sh->stream));
    sh->prepared = true;
    return 0;
}

// CUDA IPC: a handle of this rank's arena (after ccj_shard_prepare), and opening all ranks' handles on the rank that
// runs the traceback / exports tables.  handles = world consecutive cudaIpcMemHandle_t (64 bytes each).
size_t ccj_shard_ipc_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

int ccj_shard_ipc_handle(ccj_shard *sh, void *handle, size_t bytes) {
    if (!sh || !handle || bytes < sizeof(cudaIpcMemHandle_t)) return CCJ_ERR_ARG;
    if (!sh->prepared) return sfail(sh, CCJ_ERR_STATE, "ccj_shard_prepare was not called");
    cudaIpcMemHandle_t h;
    SCU(cudaIpcGetMemHandle(&h, sh->arena));
    memcpy(handle, &h, sizeof h);
    return 0;
}

int ccj_shard_open_peers(ccj_shard *sh, const void *handles, size_t bytes) {
    if (!sh || !handles || bytes < sizeof(cudaIpcMemHandle_t) * (size_t)sh->world) return CCJ_ERR_ARG;
    if (!sh->prepared) return sfail(sh, CCJ_ERR_STATE, "ccj_shard_prepare was not called");
    for (void *p : sh->opened)
        if (p) cudaIpcCloseMemHandle(p);
    sh->opened.assign(sh->world, nullptr);
    for (int r = 0; r < sh->world; ++r) {
        if (r == sh->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + sizeof h * (size_t)r, sizeof h);
        void *base = nullptr;
        SCU(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        sh->opened[r] = base;
        sh->h_locptr[r] = reinterpret_cast<int16_t *>(static_cast<char *>(base) + sh->off_loc);
    }
    SCU(cudaMemcpy(sh->arena + sh->off_locptr, sh->h_locptr.data(), sizeof(void *) * (size_t)sh->world, cudaMemcpyHostToDevice));
    sh->peers = true;
    return 0;
}

// The fill.  `count` == 1: this process is one rank of an NCCL communicator (or the only rank).  `count` == world: an
// in-process group of all ranks on one device (tests), collectives = device copies.
int ccj_shard_fill(ccj_shard **shards, int count, float *ms4) {
    if (!shards || count < 1 || !shards[0]) return CCJ_ERR_ARG;
    ccj_shard *sh = shards[0];
    const int G = sh->world, n = sh->n;
    const bool group = count > 1;
    if (group && count != G) return sfail(sh, CCJ_ERR_ARG, "an in-process group needs every rank");
    if (!group && G > 1 && !sh->comm) return sfail(sh, CCJ_ERR_STATE, "no NCCL communicator (ccj_shard_create without a unique id)");
    for (int x = 0; x < count; ++x)
        if (!shards[x] || !shards[x]->prepared || shards[x]->n != n || shards[x]->world != G || shards[x]->seq != sh->seq)
            return sfail(sh, CCJ_ERR_STATE, "every shard must be prepared with the same sequence");
    // three backends behind one loop: (1) one rank of a multi-process NCCL communicator; (2) an in-process group with one
    // GPU and one NCCL communicator per rank (ccj_shard_create_group): collectives inside ncclGroupStart/End;
    // (3) an in-process group on ONE device without NCCL (tests): collectives are device copies on the shared stream
    const bool nccl_group = group && sh->comm != nullptr;
    const bool use_nccl = (!group && G > 1) || nccl_group;
    static Nccl none;
    Nccl &N = use_nccl ? nccl() : none;
    auto on = [&](ccj_shard *z) { if (nccl_group) cudaSetDevice(z->device); return z->stream; };
    auto model = [](ccj_shard *z) { return static_cast<const ccj_model *>(ccj_internal_device_model(z->ctx)); };
    ccj::LaunchDims d;
    d.nseq = 1;
    d.nmax = n;
    for (int x = 0; x < count; ++x)
        if (!model(shards[x])) return sfail(sh, CCJ_ERR_STATE, "no energy model loaded in a rank's context");
    cudaStream_t st0 = on(sh);
    struct EventSet {   // destroyed on every return path
        std::vector<cudaEvent_t> v;
        ~EventSet() { for (cudaEvent_t e : v) if (e) cudaEventDestroy(e); }
    } evs;
    evs.v.assign((size_t)4 * n + 2, nullptr);
    std::vector<cudaEvent_t> &ev = evs.v;
    for (auto &e : ev) SCU(cudaEventCreate(&e));
    int rc = 0;
    std::string why;
    auto ck = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rc == 0) { rc = CCJ_ERR_CUDA; why = std::string(what) + ": " + cudaGetErrorString(e); }
    };
    auto cn = [&](ncclResult_t r, const char *what) {
        if (r != ncclSuccess && rc == 0) { rc = CCJ_ERR_CUDA; why = std::string(what) + ": " + N.GetErrorString(r); }
    };
    auto mark = [&](int e) { on(sh); ck(cudaEventRecord(ev[e], st0), "event"); };   // timing on rank 0, which owns the most rows
    for (int x = 0; x < count; ++x) {
        cudaStream_t st = on(shards[x]);
        ccj::launch_init(model(shards[x]), d_desc(shards[x]), d, st);
        if (shards[x]->h_desc.use_lists) ccj::launch_prep_lists(model(shards[x]), d_desc(shards[x]), d, st);
    }
    mark(0);
    ccj::NvtxRange nvtx_fill("ccj_shard_fill");
    for (int s = 0; s < n && rc == 0; ++s) {
        const int m = n - s - 2;
        char label[32];
        snprintf(label, sizeof label, "level %d", s);
        ccj::NvtxRange nvtx_level(label);
        // --- P(i,i+s) of the own rows, then the minimum over ranks of the span-s diagonal ---
        if (s >= 3 && s <= n - 1) {
            for (int x = 0; x < count; ++x) {
                ccj_shard *z = shards[x];
                const int rows = (int)ccj_shard_rows(n - s, z->rank, G);
                if (rows <= 0) continue;
                if (!shard_lean_ok(z->h_desc, sh->lev.data(), n, G))
                    k_P_shard<<<dim3(rows, s), 256, 0, on(z)>>>(model(z), d_desc(z), s);
                else if (z->h_desc.shard_shift >= 0)
                    k_P_shard_lean<true><<<dim3(rows, s), 256, (size_t)(s + 1) * sizeof(ccj_lean_lvl), on(z)>>>(model(z), d_desc(z), s);
                else
                    k_P_shard_lean<false><<<dim3(rows, s), 256, (size_t)(s + 1) * sizeof(ccj_lean_lvl), on(z)>>>(model(z), d_desc(z), s);
            }
        }
        mark(4 * s + 1);
        if (s >= 3 && s <= n - 1 && G > 1) {
            const size_t diag = (size_t)T2_P * ccj_stride2(n) + (size_t)s * (n + 1);
            if (use_nccl) {
                if (nccl_group) cn(N.GroupStart(), "ncclGroupStart");
                for (int x = 0; x < count; ++x) {
                    int32_t *dg = shards[x]->h_desc.t2 + diag;
                    cn(N.AllReduce(dg, dg, (size_t)(n + 1), ncclInt32, ncclMin, shards[x]->comm, on(shards[x])), "ncclAllReduce");
                }
                if (nccl_group) cn(N.GroupEnd(), "ncclGroupEnd");
            } else {
                int32_t *d0 = shards[0]->h_desc.t2 + diag;
                for (int x = 1; x < count; ++x) k_min_into<<<(n + 256) / 256, 256, 0, st0>>>(d0, shards[x]->h_desc.t2 + diag, n + 1);
                for (int x = 1; x < count; ++x)
                    ck(cudaMemcpyAsync(shards[x]->h_desc.t2 + diag, d0, sizeof(int32_t) * (size_t)(n + 1), cudaMemcpyDeviceToDevice, st0), "copy");
            }
        }
        mark(4 * s + 2);
        // --- 2D tables of span s (replicated), gap tables of level s (own rows) ---
        for (int x = 0; x < count; ++x) {
            ccj_shard *z = shards[x];
            cudaStream_t st = on(z);
            ccj::launch_2d(model(z), d_desc(z), d, s, st);
            const int64_t ncell = m >= 1 ? ccj_shard_slab(m, z->rank, G) : 0;
            if (ncell < 1) continue;
            const dim3 grid((unsigned)((ncell + 127) / 128), s + 1);
            if (shard_lean_ok(z->h_desc, sh->lev.data(), n, G)) {
                if (z->h_desc.shard_shift >= 0)
                    k_4d_shard_lean<true><<<grid, 128, (size_t)(s + 1) * sizeof(ccj_lean_lvl), st>>>(model(z), d_desc(z), s);
                else
                    k_4d_shard_lean<false><<<grid, 128, (size_t)(s + 1) * sizeof(ccj_lean_lvl), st>>>(model(z), d_desc(z), s);
            } else {
                k_4d_shard<<<grid, 128, 0, st>>>(model(z), d_desc(z), s);
            }
        }
        mark(4 * s + 3);
        // --- the 12 column-read tables of level s to every rank: G adjacent blocks, in place ---
        if (m >= 1 && G > 1) {
            const int64_t C = sh->lev[s + 1] - sh->lev[s];
            const size_t block = (size_t)C * CCJ_SHARD_NREP * sizeof(int16_t);     // bytes one rank contributes
            const size_t base = (size_t)sh->lev[s] * G * CCJ_SHARD_NREP * sizeof(int16_t);
            if (use_nccl) {
                if (nccl_group) cn(N.GroupStart(), "ncclGroupStart");
                for (int x = 0; x < count; ++x) {
                    ccj_shard *z = shards[x];
                    char *recv = reinterpret_cast<char *>(z->h_desc.shard_rep) + base;
                    cn(N.AllGather(recv + block * (size_t)z->rank, recv, block, ncclInt8, z->comm, on(z)), "ncclAllGather");
                }
                if (nccl_group) cn(N.GroupEnd(), "ncclGroupEnd");
            } else {
                for (int x = 0; x < count; ++x)       // rank x's block -> every other rank's copy
                    for (int y = 0; y < count; ++y)
                        if (y != x) {
                            const char *src = reinterpret_cast<const char *>(shards[x]->h_desc.shard_rep) + base + block * (size_t)shards[x]->rank;
                            char *dst = reinterpret_cast<char *>(shards[y]->h_desc.shard_rep) + base + block * (size_t)shards[x]->rank;
                            ck(cudaMemcpyAsync(dst, src, block, cudaMemcpyDeviceToDevice, st0), "copy");
                        }
            }
        }
        mark(4 * s + 4);
        ck(cudaGetLastError(), "launch");
    }
    for (int x = 0; x < count; ++x) ccj::launch_W(model(shards[x]), d_desc(shards[x]), d, on(shards[x]));
    mark(4 * n + 1);
    for (int x = 0; x < count; ++x) {
        cudaStream_t st = on(shards[x]);
        ck(cudaStreamSynchronize(st), "sync");
        ck(cudaGetLastError(), "fill");
    }
    on(sh);
    float total = 0.f, tp = 0.f, tr = 0.f, tc = 0.f, tg = 0.f;
    std::vector<float> lvl((size_t)4 * n, 0.f);
    if (rc == 0) {
        cudaEventElapsedTime(&total, ev[0], ev[4 * n + 1]);
        for (int s = 0; s < n; ++s) {
            float a = 0, b = 0, c2 = 0, g2 = 0;
            cudaEventElapsedTime(&a, ev[4 * s], ev[4 * s + 1]);
            cudaEventElapsedTime(&b, ev[4 * s + 1], ev[4 * s + 2]);
            cudaEventElapsedTime(&c2, ev[4 * s + 2], ev[4 * s + 3]);
            cudaEventElapsedTime(&g2, ev[4 * s + 3], ev[4 * s + 4]);
            tp += a; tr += b; tc += c2; tg += g2;
            lvl[4 * s] = a; lvl[4 * s + 1] = b; lvl[4 * s + 2] = c2; lvl[4 * s + 3] = g2;
        }
    }
    if (rc) return sfail(sh, rc, why);
    for (int x = 0; x < count; ++x) {
        shards[x]->filled = true;
        shards[x]->ms[0] = total; shards[x]->ms[1] = tp + tc; shards[x]->ms[2] = tg; shards[x]->ms[3] = tr;
        shards[x]->level_ms = lvl;
    }
    if (ms4) { ms4[0] = total; ms4[1] = tp + tc; ms4[2] = tg; ms4[3] = tr; }
    return 0;
}

// per-step device times of the last fill: 4 floats per step s = 0..n-1 (P kernel, allreduce, 2D + gap-table kernels,
// allgather), and the bytes this rank contributes to the allgather of level s
int ccj_shard_level_ms(ccj_shard *sh, float *out, int64_t out_len) {
    if (!sh || !out) return CCJ_ERR_ARG;
    if (!sh->filled || out_len < (int64_t)sh->level_ms.size()) return sfail(sh, CCJ_ERR_STATE, "no fill / buffer too small");
    memcpy(out, sh->level_ms.data(), sh->level_ms.size() * sizeof(float));
    return 0;
}
int64_t ccj_shard_level_bytes(int n, int world, int level) {
    return ccj_shard_level_cells(n, level, world) * CCJ_SHARD_NREP * (int64_t)sizeof(int16_t);
}

// in-process groups: give shard `sh` direct pointers to the other shards' row-local tables (same process, same device
// or peer-enabled devices) -- the counterpart of ccj_shard_open_peers
int ccj_shard_link_local(ccj_shard *sh, ccj_shard **all, int count) {
    if (!sh || !all || count != sh->world) return CCJ_ERR_ARG;
    for (int x = 0; x < count; ++x) {
        if (!all[x] || !all[x]->prepared) return sfail(sh, CCJ_ERR_STATE, "every shard must be prepared");
        sh->h_locptr[all[x]->rank] = reinterpret_cast<int16_t *>(all[x]->arena + all[x]->off_loc);
    }
    SCU(cudaMemcpy(sh->arena + sh->off_locptr, sh->h_locptr.data(), sizeof(void *) * (size_t)sh->world, cudaMemcpyHostToDevice));
    sh->peers = true;
    return 0;
}

// W_final::ccj's traceback (src/W_final.cc:84-103) on this rank; needs the peers' row-local tables when world > 1
int ccj_shard_traceback(ccj_shard *sh, ccj_result *res, int32_t *pairs, char *structs, float *ms) {
    if (!sh || !res) return CCJ_ERR_ARG;
    if (!sh->filled) return sfail(sh, CCJ_ERR_STATE, "ccj_shard_fill was not called");
    if (sh->world > 1 && !sh->peers) return sfail(sh, CCJ_ERR_STATE, "the traceback rank must open its peers' memory first");
    const int n = sh->n;
    const ccj_model *M = static_cast<const ccj_model *>(ccj_internal_device_model(sh->ctx));
    if (!M) return sfail(sh, CCJ_ERR_STATE, "no energy model loaded");
    ccj::LaunchDims d;
    d.nseq = 1;
    d.nmax = n;
    cudaEvent_t e0, e1;
    SCU(cudaEventCreate(&e0));
    SCU(cudaEventCreate(&e1));
    cudaEventRecord(e0, sh->stream);
    ccj::launch_traceback(M, d_desc(sh), d, sh->stream);
    cudaEventRecord(e1, sh->stream);
    std::vector<int32_t> out(CCJ_STATUS_INTS + (size_t)(n + 1) + (size_t)(n + 2));
    cudaError_t e = cudaMemcpyAsync(out.data(), sh->arena + sh->off_out, out.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, sh->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(sh->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    float t = 0.f;
    if (e == cudaSuccess) cudaEventElapsedTime(&t, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    SCU(e);
    if (ms) *ms = t;
    const int32_t *stt = out.data(), *W = stt + CCJ_STATUS_INTS, *pr = W + (n + 1);
    res->energy_dcal = W[n];
    res->status = stt[0];
    res->n_should_not_be_here = stt[1];
    res->msg_id = stt[2];
    res->aux_i = stt[3];
    res->aux_j = stt[4];
    if (pairs)
        for (int x = 0; x < n; ++x) pairs[x] = pr[x + 1];
    if (structs) {
        const std::string s = ccj::fill_structure(n, pr);
        memcpy(structs, s.data(), (size_t)n);
    }
    return 0;
}

int ccj_shard_energy(ccj_shard *sh, int32_t *energy_dcal) {   // W[n] after the fill (every rank holds W)
    if (!sh || !energy_dcal) return CCJ_ERR_ARG;
    if (!sh->filled) return sfail(sh, CCJ_ERR_STATE, "ccj_shard_fill was not called");
    SCU(cudaMemcpy(energy_dcal, sh->h_desc.W + sh->n, sizeof(int32_t), cudaMemcpyDeviceToHost));
    return 0;
}

// FNV-1a hash of one gap table in the canonical export order (the hash oracle/ref_dump.cc prints); on a rank that
// can read every rank's rows
int ccj_shard_table4_hash(ccj_shard *sh, int table, uint64_t *hash, int64_t *finite, int32_t *min_value) {
    if (!sh || !hash || table < 0 || table >= CCJ_NT4) return CCJ_ERR_ARG;
    if (!sh->filled) return sfail(sh, CCJ_ERR_STATE, "ccj_shard_fill was not called");
    if (sh->world > 1 && !sh->peers && sh->h_desc.shard_kind[table] >= CCJ_SHARD_NREP)
        return sfail(sh, CCJ_ERR_STATE, "row-local table: open the peers' memory first");
    const int n = sh->n;
    const int64_t cells = ccj_cells4(n);
    int16_t *d_out = nullptr;
    SCU(cudaMalloc((void **)&d_out, (size_t)cells * sizeof(int16_t) + 16));
    for (int t = 0; t <= n - 3; ++t) {
        const int m = n - t - 2, itiles = (m + 3) / 4, ktiles = (m + 31) / 32;
        k_shard_export<<<dim3(itiles * ktiles, t + 1), dim3(32, 4), 0, sh->stream>>>(d_desc(sh), table, t, ktiles, d_out);
    }
    std::vector<int16_t> raw((size_t)cells + 1);
    cudaError_t e = cudaMemcpyAsync(raw.data(), d_out, (size_t)cells * sizeof(int16_t), cudaMemcpyDeviceToHost, sh->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(sh->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(d_out);
    SCU(e);
    uint64_t h = 1469598103934665603ULL;
    int64_t fin = 0;
    int32_t mn = 1 << 30;
    for (int i = 1; i <= n; ++i)
        for (int j = i; j <= n; ++j)
            for (int k = j + 2; k <= n; ++k)
                for (int l = k; l <= n; ++l) {
                    const int16_t v = raw[(size_t)ccj_idx4(n, i, j, k, l)];
                    fnv_add(h, (uint16_t)v);
                    if (v < 32767) {
                        ++fin;
                        if (v < mn) mn = v;
                    }
                }
    *hash = h;
    if (finite) *finite = fin;
    if (min_value) *min_value = fin ? mn : 0;
    return 0;
}

int ccj_shard_table2_hash(ccj_shard *sh, int table, uint64_t *hash, int64_t *finite, int64_t *sum) {
    if (!sh || !hash || table < 0 || table >= CCJ_NT2) return CCJ_ERR_ARG;
    if (!sh->filled) return sfail(sh, CCJ_ERR_STATE, "ccj_shard_fill was not called");
    const int n = sh->n;
    std::vector<int32_t> raw((size_t)ccj_stride2(n));
    SCU(cudaMemcpy(raw.data(), sh->h_desc.t2 + (size_t)table * ccj_stride2(n), raw.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    uint64_t h = 1469598103934665603ULL;
    int64_t fin = 0, sm = 0;
    for (int i = 1; i <= n; ++i)
        for (int j = i; j <= n; ++j) {
            const int32_t v = raw[(size_t)ccj_idx2(n, i, j)];
            fnv_add(h, (uint32_t)v);
            if (v < CCJ_INF / 2) {
                ++fin;
                sm += v;
            }
        }
    *hash = h;
    if (finite) *finite = fin;
    if (sum) *sum = sm;
    return 0;
}

}  // extern "C"
