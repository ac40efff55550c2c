// extern "C" layer of libccj_b200.so: contexts, waves, launch schedule, result read-back.
// See include/ccj_b200.h for the contract and the reference code each entry point replaces.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ccj_b200.h"
#include "ccj_kernels.cuh"
#include "ccj_render.hpp"
#include "energy_model.hpp"

namespace {

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct SeqPlan {
    int n;
    size_t in_off, out_off, tab_off;  // offsets into the three regions
    size_t in_bytes, out_bytes, tab_bytes;
};

// the per-sequence device buffers behind ccj_seq, in arena order
enum { TAB_T4 = 0, TAB_G1, TAB_G2, TAB_G3, TAB_G4, TAB_T2, TAB_W3, TAB_ESTP, TAB_INLIST, TAB_OUTLIST, TAB_INCNT, TAB_OUTCNT, TAB_LAY, TAB_SCRATCH,
       TAB_PLW, TAB_PRW, TAB_PMW, TAB_PMM, TAB_PKF, TAB_PKG, TAB_WSCR, TAB_PMLEV, TAB_PLIST, TAB_PCUM, TAB_PMLIST, TAB_PMSTART, TAB_FTYPE, TAB_TBSTACK,
       TAB_COUNT };
size_t tab_bytes_uncached(int n, int which) {
    const size_t tri = (size_t)n * (n - 1) / 2 + 1;
    // sequences beyond the tuned kernels' range run the generic level kernel, which only needs the 22 tables, the
    // 2D tables and the traceback buffers: without the tuned copies such a sequence still fits one GPU up to n~530
    if (!ccj::fill4_tuned_supported(n)) {
        switch (which) {
            case TAB_T4: return align_up((size_t)ccj_cells4(n) * CCJ_NT4 * sizeof(int16_t) + 16, 256);
            case TAB_T2:
            case TAB_W3:
            case TAB_ESTP:      // the generic cell functions walk the partner lists of k_prep too
            case TAB_INLIST:
            case TAB_OUTLIST:
            case TAB_INCNT:
            case TAB_OUTCNT:
            case TAB_FTYPE:
            case TAB_TBSTACK: break;
            default: return 256;
        }
    }
    switch (which) {
        case TAB_T4: return align_up((size_t)ccj_cells4(n) * CCJ_NT4_STORE * sizeof(int16_t) + 16, 256);
        case TAB_G1:
        case TAB_G2:
        case TAB_G3: return align_up((size_t)ccj_cells4(n) * 6 * sizeof(int16_t) + 64, 256);
        case TAB_G4: return align_up((size_t)ccj_cells4(n) * 8 * sizeof(int16_t) + 64, 256);
        case TAB_T2: return align_up((size_t)ccj_stride2(n) * CCJ_NT2 * sizeof(int32_t), 256);
        case TAB_W3: return align_up((size_t)ccj_stride2(n) * 4 * sizeof(int32_t), 256);
        case TAB_ESTP: return align_up((size_t)ccj_stride2(n) * sizeof(int32_t), 256);
        case TAB_INLIST: return align_up(tri * CCJ_WIN_IN * sizeof(uint32_t), 256);
        case TAB_OUTLIST: return align_up(tri * CCJ_WIN_OUT * 2 * sizeof(uint32_t), 256);
        case TAB_INCNT:
        case TAB_OUTCNT: return align_up(tri * sizeof(int32_t), 256);
        case TAB_LAY: return align_up((size_t)CCJ_LAY_INTS(n) * sizeof(int32_t), 256);
        case TAB_PKF: return align_up((size_t)ccj_pkf_total(n) * sizeof(int16_t) + 64, 256);
        case TAB_PKG: return align_up((size_t)ccj_pkg_total(n) * sizeof(int16_t) + 64, 256);
        case TAB_SCRATCH: return align_up((size_t)ccj_level_max(n) * ccj::fill4_partials() * sizeof(int16_t) + 64, 256);
        case TAB_PLW:
        case TAB_PRW: return align_up((size_t)ccj_winlr_quads(n) * 4 * sizeof(int16_t) + 256, 256);   // lanes read up to 60 cells past a run
        case TAB_PMW:
        case TAB_PMM: return align_up(((size_t)ccj_pmw_level_quads(n) * (size_t)(n > 2 ? n - 2 : 1) + 16) * 8, 256);  // 8 bytes per quad: 4 values resp. 4 masks
        case TAB_WSCR: return align_up(((size_t)ccj_winlr_level_max(n) * 4 + (size_t)ccj_pmw_level_quads(n) * 8) * sizeof(int16_t) + 64, 256);
        case TAB_PMLEV: return align_up((size_t)(n + 1) * (n + 1) * sizeof(int32_t) + 64, 256);
        case TAB_PLIST: return align_up((size_t)(n + 1) * (n + 1) * sizeof(int32_t), 256);
        case TAB_PCUM: return align_up((size_t)(n + 1) * (n + 2) * sizeof(int32_t), 256);
        case TAB_PMLIST: return align_up(tri * sizeof(int32_t), 256);
        case TAB_PMSTART: return align_up((size_t)(n + 4) * sizeof(int32_t), 256);
        case TAB_FTYPE: return align_up((size_t)n + 2, 256);
        case TAB_TBSTACK: return align_up(sizeof(int32_t) * 5 * (size_t)(16 * n + 64), 256);
    }
    return 0;
}
// the window layouts need O(n^2) host loops to size: memoise per length
struct TabSizes { size_t bytes[TAB_COUNT]; size_t off[TAB_COUNT + 1]; int64_t wscr_lr; int32_t wtot4; };
const TabSizes &tab_sizes(int n) {
    static std::mutex mu;
    static std::map<int, TabSizes> cache;
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(n);
    if (it != cache.end()) return it->second;
    TabSizes z;
    z.off[0] = 0;
    for (int x = 0; x < TAB_COUNT; ++x) {
        z.bytes[x] = tab_bytes_uncached(n, x);
        z.off[x + 1] = z.off[x] + z.bytes[x];
    }
    z.wscr_lr = ccj_winlr_level_max(n);
    z.wtot4 = (int32_t)ccj_pmw_level_quads(n);
    return cache.emplace(n, z).first->second;
}
size_t tab_bytes(int n, int which) { return tab_sizes(n).bytes[which]; }
size_t tab_offset(int n, int which) { return tab_sizes(n).off[which]; }

// per-sequence byte needs
void plan_seq(int n, SeqPlan &p) {
    p.n = n;
    // inputs: S (n+2 int8), seq (n chars)
    p.in_bytes = align_up((size_t)(n + 2) + (size_t)n + 2, 16);
    // outputs: status | W[0..n] | pair[0..n+1]
    p.out_bytes = align_up(sizeof(int32_t) * (CCJ_STATUS_INTS + (size_t)(n + 1) + (size_t)(n + 2)), 16);
    // tables: t4, t2, ftype, traceback stack
    p.tab_bytes = 0;
    for (int x = 0; x < TAB_COUNT; ++x) p.tab_bytes += tab_bytes(n, x);
}

}  // namespace

struct ccj_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t s_win = nullptr, s_2d = nullptr;  // side streams of the fill: interior windows, P + 2D tables
    std::vector<cudaEvent_t> dep;                  // dependency events of the captured launch sequence
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    bool model_ok = false;
    ccj_model *h_model = nullptr;  // heap (large)
    ccj_model *d_model = nullptr;

    // arena
    char *d_arena = nullptr;
    size_t arena_bytes = 0;
    ccj_seq *d_seqs = nullptr;
    size_t d_seqs_cap = 0;
    ccj_seq *d_seqs2 = nullptr;    // descriptors of the sequences a fill has to repeat with the generic kernels
    size_t d_seqs2_cap = 0;
    std::vector<ccj_seq> h_seqs;   // host copy of the wave's descriptors
    char *h_stage = nullptr;  // pinned staging for inputs/outputs
    size_t h_stage_bytes = 0;

    // current wave
    std::vector<SeqPlan> plan;
    std::vector<std::string> wave_seqs;
    size_t in_total = 0, out_total = 0, tab_total = 0;
    int nmax = 0;
    bool prepared = false, filled = false, traced = false;
    float fill_ms = 0.f, tb_ms = 0.f;
    int fill_launches = 0;

    // CUDA graphs of the fill, keyed by (nmax, nseq)
    std::map<std::pair<int, int>, cudaGraphExec_t> graphs;
};

namespace {

int fail(ccj_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg;
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(ctx, CCJ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

void drop_graphs(ccj_ctx *ctx) {
    for (auto &g : ctx->graphs) cudaGraphExecDestroy(g.second);
    ctx->graphs.clear();
}

int ensure_arena(ccj_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->arena_bytes) return 0;
    if (ctx->d_arena) {
        CU(cudaFree(ctx->d_arena));
        ctx->d_arena = nullptr;
        ctx->arena_bytes = 0;
    }
    CU(cudaMalloc((void **)&ctx->d_arena, bytes));
    ctx->arena_bytes = bytes;
    return 0;
}

int ensure_stage(ccj_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->h_stage_bytes) return 0;
    if (ctx->h_stage) CU(cudaFreeHost(ctx->h_stage));
    ctx->h_stage = nullptr;
    ctx->h_stage_bytes = 0;
    CU(cudaMallocHost((void **)&ctx->h_stage, bytes));
    ctx->h_stage_bytes = bytes;
    return 0;
}

int ensure_seqs(ccj_ctx *ctx, size_t count) {
    if (count <= ctx->d_seqs_cap) return 0;
    drop_graphs(ctx);  // graphs bake the descriptor pointer
    if (ctx->d_seqs) CU(cudaFree(ctx->d_seqs));
    ctx->d_seqs = nullptr;
    size_t cap = std::max<size_t>(count, 1024);
    CU(cudaMalloc((void **)&ctx->d_seqs, cap * sizeof(ccj_seq)));
    ctx->d_seqs_cap = cap;
    return 0;
}

size_t budget_bytes(ccj_ctx *ctx) {
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) return 0;
    fr += ctx->arena_bytes;  // our own arena can be recycled
    const size_t reserve = (size_t)1 << 30;
    return fr > reserve ? fr - reserve : 0;
}

// the level kernels put the sequence index (times up to 4 roles) into gridDim.y/z, which CUDA limits to 65535
const int kMaxWave = 65535 / 4;

int validate(ccj_ctx *ctx, const char *s, int64_t len) {
    if (len <= 0) return fail(ctx, CCJ_ERR_SEQUENCE, "sequence is missing");
    if (len > CCJ_HAIRPIN_TAB - 2) return fail(ctx, CCJ_ERR_TOO_LARGE, "sequence longer than supported");
    for (int64_t x = 0; x < len; ++x) {
        const char ch = s[x];
        if (!(ch == 'G' || ch == 'C' || ch == 'A' || ch == 'U' || ch == 'T')) {
            char buf[128];
            snprintf(buf, sizeof buf, "Sequence contains character %c that is not G,C,A,U, or T.", ch);
            return fail(ctx, CCJ_ERR_SEQUENCE, buf);
        }
    }
    return 0;
}

// CCJ_GENERIC_SCAN=1: the generic cell functions scan the 29 x 29 interior windows like the reference instead of
// walking k_prep's partner lists (debugging aid; read once)
bool generic_scan() {
    static const bool v = [] { const char *e = getenv("CCJ_GENERIC_SCAN"); return e && e[0] == '1'; }();
    return v;
}

// the fill's launch sequence: per span s  K_P(s) -> K_2D(s) -> K_4D(level s)   (DESIGN.md "schedule")
bool use_tuned(int nmax) {
    const char *g = getenv("CCJ_FILL_GENERIC");  // debugging aid: force the generic-index level kernel
    return !(g && g[0] == '1') && ccj::fill4_tuned_supported(nmax);
}

// Tuned path, three streams (all inside the captured graph); KF = ccj::fill4_fused_levels() = 3, lead = KF-2 = 1:
//   main : k_roles(t) for t = 0 mod KF -- the split points of levels t..t+KF-1 whose sources lie on levels <= t-1;
//          its partner cells on level t+KF-1 read {WB,WP,WBP} of spans <= t+KF-2, hence the wait on evD[t+lead].
//          k_final(t) for every t -- same-cell assembly plus the <= (t mod KF) late split points; reads levels <= t-1
//          (main-stream order), 2D spans <= t (covered by the evD wait of its k_roles) and the window partials (evW[t]).
//   s_win: k_winLR(t), k_winM(t).  A window candidate shortens an arm by x >= 1 on one side and y >= 1 on the other, so
//          level t only reads the window copies of levels <= t-2; it writes the partials of level t into wscr[t & 1],
//          whose previous content k_final(t-2) consumed.  One wait on evF[t-2] covers both -> the windows run one
//          level ahead of the main stream.
//   s_2d : k_P(s) -> k_2d(s) for span s, `lead` spans ahead of the levels.  P(i,i+s) reads PK of levels <= s-3: wait on
//          evF[s-3].
// Every CUDA call is checked; the first error ends the sequence and is returned (a failed launch inside a capture
// is reported by cudaGetLastError, not by the launch statement).
#define EQ(call)                                   \
    do {                                           \
        const cudaError_t e_ = (call);             \
        if (e_ != cudaSuccess) return e_;          \
    } while (0)
#define EQL(launch)                                \
    do {                                           \
        launch;                                    \
        const cudaError_t e_ = cudaGetLastError(); \
        if (e_ != cudaSuccess) return e_;          \
    } while (0)
cudaError_t enqueue_fill(ccj_ctx *ctx, ccj::LaunchDims d) {
    const ccj_model *M = ctx->d_model;
    const ccj_seq *Q = ctx->d_seqs;
    cudaStream_t s0 = ctx->stream;
    if (!use_tuned(d.nmax)) {
        EQL(ccj::launch_init(M, Q, d, s0));
        if (!generic_scan()) EQL(ccj::launch_prep_lists(M, Q, d, s0));
        for (int s = 0; s < d.nmax; ++s) {
            EQL(ccj::launch_P(M, Q, d, s, s0));
            EQL(ccj::launch_2d(M, Q, d, s, s0));
            EQL(ccj::launch_4d(M, Q, d, s, s0));
        }
        EQL(ccj::launch_W(M, Q, d, s0));
        return cudaSuccess;
    }
    const int nm = d.nmax;
    while ((int)ctx->dep.size() < 3 * nm + 2) {
        cudaEvent_t e = nullptr;
        EQ(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->dep.push_back(e);
    }
    cudaEvent_t *evF = ctx->dep.data(), *evD = evF + nm, *evW = evD + nm, evPrep = ctx->dep[3 * nm];
    cudaStream_t sw = ctx->s_win, s2 = ctx->s_2d;
    EQL(ccj::launch_init(M, Q, d, s0));
    EQL(ccj::launch_prep(M, Q, d, s0));
    EQ(cudaEventRecord(evPrep, s0));
    const int last_level = nm - 3;
    if (last_level >= 0) EQ(cudaStreamWaitEvent(sw, evPrep, 0));
    EQ(cudaStreamWaitEvent(s2, evPrep, 0));
    const int KFL = ccj::fill4_fused_levels(), lead = KFL - 2 > 0 ? KFL - 2 : 0;
    auto span_step = [&](int sp) -> cudaError_t {
        if (sp >= nm) return cudaSuccess;
        if (sp >= 3 && sp - 3 <= last_level) EQ(cudaStreamWaitEvent(s2, evF[sp - 3], 0));
        EQL(ccj::launch_P_tuned(M, Q, d, sp, s2));
        EQL(ccj::launch_2d(M, Q, d, sp, s2));
        EQ(cudaEventRecord(evD[sp], s2));
        return cudaSuccess;
    };
    for (int sp = 0; sp < lead; ++sp) EQ(span_step(sp));
    for (int s = 0; s < nm; ++s) {
        EQ(span_step(s + lead));
        if (s <= last_level) {
            if (s >= 2) EQ(cudaStreamWaitEvent(sw, evF[s - 2], 0));
            EQL(ccj::launch_4d_windows(M, Q, d, s, sw));
            EQ(cudaEventRecord(evW[s], sw));
            if (s % KFL == 0) {
                EQ(cudaStreamWaitEvent(s0, evD[std::min(s + lead, nm - 1)], 0));
                EQL(ccj::launch_4d_roles(M, Q, d, s, s0));
            }
            EQ(cudaStreamWaitEvent(s0, evW[s], 0));
            EQL(ccj::launch_4d_final(M, Q, d, s, s0));
            EQ(cudaEventRecord(evF[s], s0));
        }
    }
    EQ(cudaStreamWaitEvent(s0, evD[nm - 1], 0));
    if (last_level >= 0) EQ(cudaStreamWaitEvent(s0, evW[last_level], 0));
    EQL(ccj::launch_W(M, Q, d, s0));
    return cudaSuccess;
}
#undef EQ
#undef EQL

// Sequences whose tables reached the range where the reference's int16 narrowing wraps (flag status[7] of k_final,
// ccj_fill4.cu): the tuned kernels saturate there, so those sequences are filled again by the generic kernels, which
// evaluate every candidate in 32 bits and narrow exactly like Matrix4D::set.  Only designed inputs get here.
int refill_wrapped(ccj_ctx *ctx, ccj::LaunchDims d) {
    CU(cudaMemcpyAsync(ctx->h_stage, ctx->d_arena + ctx->in_total, ctx->out_total, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    std::vector<ccj_seq> redo;
    int nmax = 0;
    for (size_t s = 0; s < ctx->plan.size(); ++s) {
        const int32_t *st = reinterpret_cast<const int32_t *>(ctx->h_stage + ctx->plan[s].out_off);
        if (st[7] != 0) {
            redo.push_back(ctx->h_seqs[s]);
            nmax = std::max(nmax, ctx->plan[s].n);
        }
    }
    if (redo.empty()) return 0;
    if (redo.size() > ctx->d_seqs2_cap) {
        if (ctx->d_seqs2) CU(cudaFree(ctx->d_seqs2));
        ctx->d_seqs2 = nullptr;
        ctx->d_seqs2_cap = 0;
        CU(cudaMalloc((void **)&ctx->d_seqs2, redo.size() * sizeof(ccj_seq)));
        ctx->d_seqs2_cap = redo.size();
    }
    CU(cudaMemcpyAsync(ctx->d_seqs2, redo.data(), redo.size() * sizeof(ccj_seq), cudaMemcpyHostToDevice, ctx->stream));
    ccj::LaunchDims d2;
    d2.nseq = (int)redo.size();
    d2.nmax = nmax;
    const ccj_model *M = ctx->d_model;
    cudaStream_t s0 = ctx->stream;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    cudaEventRecord(e0, s0);
    ccj::launch_init(M, ctx->d_seqs2, d2, s0);   // also clears status[6]: the traceback reads these sequences generically
    if (!generic_scan()) ccj::launch_prep_lists(M, ctx->d_seqs2, d2, s0);
    for (int s = 0; s < d2.nmax; ++s) {
        ccj::launch_P(M, ctx->d_seqs2, d2, s, s0);
        ccj::launch_2d(M, ctx->d_seqs2, d2, s, s0);
        ccj::launch_4d(M, ctx->d_seqs2, d2, s, s0);
    }
    ccj::launch_W(M, ctx->d_seqs2, d2, s0);
    cudaEventRecord(e1, s0);
    cudaError_t e = cudaStreamSynchronize(s0);
    if (e == cudaSuccess) e = cudaGetLastError();
    float ms = 0.f;
    if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CU(e);
    ctx->fill_ms += ms;
    ctx->fill_launches += ccj::fill_launch_count(d2.nmax, false);
    return 0;
}

}  // namespace

extern "C" {

const char *ccj_version(void) { return "ccj_b200 0.1 (sm_100a)"; }

void ccj_ctx_destroy(ccj_ctx *ctx);

int ccj_ctx_create(int device, ccj_ctx **out) {
    if (!out) return CCJ_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) return CCJ_ERR_CUDA;
    ccj_ctx *ctx = new ccj_ctx();
    ctx->device = device;
    // the main stream carries the bandwidth-bound split-point kernels and the assembly (the critical path);
    // the side streams only fill the issue slots it leaves idle -> lower priority
    int prio_lo = 0, prio_hi = 0;
    if (cudaSetDevice(device) == cudaSuccess) cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->s_win, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->s_2d, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
        cudaMalloc((void **)&ctx->d_model, sizeof(ccj_model)) != cudaSuccess) {
        ccj_ctx_destroy(ctx);   // null-checks every handle it releases
        return CCJ_ERR_CUDA;
    }
    ctx->h_model = new ccj_model();
    *out = ctx;
    return 0;
}

void ccj_ctx_destroy(ccj_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    drop_graphs(ctx);
    if (ctx->d_arena) cudaFree(ctx->d_arena);
    if (ctx->d_seqs) cudaFree(ctx->d_seqs);
    if (ctx->d_seqs2) cudaFree(ctx->d_seqs2);
    if (ctx->d_model) cudaFree(ctx->d_model);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t e : ctx->dep) cudaEventDestroy(e);
    if (ctx->s_win) cudaStreamDestroy(ctx->s_win);
    if (ctx->s_2d) cudaStreamDestroy(ctx->s_2d);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx->h_model;
    delete ctx;
}

const char *ccj_last_error(const ccj_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }

// defaults (the reference's compiled-in Turner-2004 tables), then the file or embedded set on top of them
static bool read_params(const char *par_file, const char *embedded, ccj::RawParams &rp, std::string &err) {
    if (!ccj::load_defaults(rp, err)) return false;
    if (embedded) {
        size_t len = 0;
        const char *text = ccj::embedded_par(embedded, &len);
        if (!text) {
            err = std::string("unknown embedded parameter set ") + embedded;
            return false;
        }
        return ccj::load_par_text(text, len, rp, err);
    }
    return ccj::load_par_file(par_file, rp, err);
}

static int model_load_common(ccj_ctx *ctx, const char *par_file, const char *embedded, int dangles, int no_gu) {
    ccj::RawParams *rp = new ccj::RawParams();
    std::string err;
    if (!read_params(par_file, embedded, *rp, err)) {
        delete rp;
        return fail(ctx, CCJ_ERR_PARAMS, err);
    }
    // the reference's reader prints its symmetry warnings while loading (vrna_message_warning, utils.c:180-196)
    for (const std::string &w : rp->warnings) fprintf(stderr, "WARNING: %s\n", w.c_str());
    ccj::build_model(*rp, dangles, no_gu, *ctx->h_model);
    delete rp;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(ctx->d_model, ctx->h_model, sizeof(ccj_model), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->model_ok = true;
    return 0;
}

int ccj_model_load(ccj_ctx *ctx, const char *par_file, int dangles, int no_gu) {
    if (!ctx || !par_file) return CCJ_ERR_ARG;
    return model_load_common(ctx, par_file, nullptr, dangles, no_gu);
}

int ccj_model_load_embedded(ccj_ctx *ctx, const char *name, int dangles, int no_gu) {
    if (!ctx || !name) return CCJ_ERR_ARG;
    return model_load_common(ctx, nullptr, name, dangles, no_gu);
}

int64_t ccj_wave_capacity(ccj_ctx *ctx, int n) {
    if (!ctx || n < 1) return 0;
    cudaSetDevice(ctx->device);
    SeqPlan p;
    plan_seq(n, p);
    const size_t per = p.in_bytes + p.out_bytes + p.tab_bytes + 512;
    return std::min<int64_t>((int64_t)(budget_bytes(ctx) / per), kMaxWave);
}

void *ccj_stream(ccj_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
// library-internal (ccj_shard.cu): the model blob in device memory
const void *ccj_internal_device_model(ccj_ctx *ctx) { return ctx && ctx->model_ok ? ctx->d_model : nullptr; }
int ccj_internal_device(ccj_ctx *ctx) { return ctx ? ctx->device : -1; }
const void *ccj_internal_host_model(ccj_ctx *ctx) { return ctx && ctx->model_ok ? ctx->h_model : nullptr; }
int ccj_device_count(void) {
    int count = 0;
    return cudaGetDeviceCount(&count) == cudaSuccess ? count : 0;
}

int ccj_batch_prepare(ccj_ctx *ctx, const char *seqs, const int64_t *offsets, int nseq) {
    if (!ctx || !seqs || !offsets || nseq < 1) return CCJ_ERR_ARG;
    ccj::NvtxRange nvtx("ccj_batch_prepare");
    if (!ctx->model_ok) return fail(ctx, CCJ_ERR_STATE, "no energy model loaded");
    if (nseq > kMaxWave) return fail(ctx, CCJ_ERR_TOO_LARGE, "more than 16383 sequences in one wave; use ccj_fold_batch");
    CU(cudaSetDevice(ctx->device));
    ctx->prepared = ctx->filled = ctx->traced = false;
    ctx->plan.assign(nseq, SeqPlan());
    ctx->wave_seqs.assign(nseq, std::string());
    size_t in_total = 0, out_total = 0, tab_total = 0;
    int nmax = 0;
    for (int s = 0; s < nseq; ++s) {
        const int64_t len = offsets[s + 1] - offsets[s];
        int rc = validate(ctx, seqs + offsets[s], len);
        if (rc) return rc;
        SeqPlan &p = ctx->plan[s];
        plan_seq((int)len, p);
        p.in_off = in_total;
        p.out_off = out_total;
        p.tab_off = tab_total;
        in_total += p.in_bytes;
        out_total += p.out_bytes;
        tab_total += p.tab_bytes;
        nmax = std::max(nmax, (int)len);
        ctx->wave_seqs[s].assign(seqs + offsets[s], (size_t)len);
    }
    in_total = align_up(in_total, 256);
    out_total = align_up(out_total, 256);
    const size_t need = in_total + out_total + tab_total;
    if (need > budget_bytes(ctx)) return fail(ctx, CCJ_ERR_TOO_LARGE, "batch does not fit GPU memory; use ccj_fold_batch");
    int rc = ensure_arena(ctx, need);
    if (rc) return rc;
    rc = ensure_stage(ctx, std::max(in_total, out_total) + (size_t)nseq * sizeof(ccj_seq));
    if (rc) return rc;
    rc = ensure_seqs(ctx, nseq);
    if (rc) return rc;

    // stage inputs + descriptors
    char *d_in = ctx->d_arena, *d_out = ctx->d_arena + in_total, *d_tab = ctx->d_arena + in_total + out_total;
    memset(ctx->h_stage, 0, in_total);
    ccj_seq *hd = reinterpret_cast<ccj_seq *>(ctx->h_stage + std::max(in_total, out_total));
    for (int s = 0; s < nseq; ++s) {
        const SeqPlan &p = ctx->plan[s];
        const int n = p.n;
        int8_t *S = reinterpret_cast<int8_t *>(ctx->h_stage + p.in_off);
        char *sq = ctx->h_stage + p.in_off + (n + 2);
        const std::string &str = ctx->wave_seqs[s];
        for (int x = 1; x <= n; ++x) S[x] = (int8_t)ccj::encode_base(str[x - 1]);
        S[n + 1] = S[1];
        S[0] = S[n];
        memcpy(sq, str.data(), n);
        ccj_seq &q = hd[s];
        memset(&q, 0, sizeof q);
        q.n = n;
        q.S = reinterpret_cast<const int8_t *>(d_in + p.in_off);
        q.seq = d_in + p.in_off + (n + 2);
        char *o = d_out + p.out_off;
        q.status = reinterpret_cast<int32_t *>(o);
        q.W = q.status + CCJ_STATUS_INTS;
        q.pair_out = q.W + (n + 1);
        char *t = d_tab + p.tab_off;
        q.t4 = reinterpret_cast<int16_t *>(t + tab_offset(n, TAB_T4));
        q.stride4 = ccj_cells4(n);
        q.g1 = reinterpret_cast<int16_t *>(t + tab_offset(n, TAB_G1));
        q.g2 = reinterpret_cast<int16_t *>(t + tab_offset(n, TAB_G2));
        q.g3 = reinterpret_cast<int16_t *>(t + tab_offset(n, TAB_G3));
        q.g4 = reinterpret_cast<int16_t *>(t + tab_offset(n, TAB_G4));
        q.t2 = reinterpret_cast<int32_t *>(t + tab_offset(n, TAB_T2));
        q.stride2 = ccj_stride2(n);
        q.w3 = reinterpret_cast<int32_t *>(t + tab_offset(n, TAB_W3));
        q.estP = reinterpret_cast<int32_t *>(t + tab_offset(n, TAB_ESTP));
        q.inlist = reinterpret_cast<uint32_t *>(t + tab_offset(n, TAB_INLIST));
        q.outlist = reinterpret_cast<uint32_t *>(t + tab_offset(n, TAB_OUTLIST));
        q.incnt = reinterpret_cast<int32_t *>(t + tab_offset(n, TAB_INCNT));
        q.outcnt = reinterpret_cast<int32_t *>(t + tab_offset(n, TAB_OUTCNT));
        q.lay = reinterpret_cast<int32_t *>(t + tab_offset(n, TAB_LAY));
        q.scratch = reinterpret_cast<int16_t *>(t + tab_offset(n, TAB_SCRATCH));
        q.scratch_stride = ccj_level_max(n);
        q.plw = reinterpret_cast<int16_t *>(t + tab_offset(n, TAB_PLW));
        q.prw = reinterpret_cast<int16_t *>(t + tab_offset(n, TAB_PRW));
        q.pmw = reinterpret_cast<int16_t *>(t + tab_offset(n, TAB_PMW));
        q.pmm = reinterpret_cast<int16_t *>(t + tab_offset(n, TAB_PMM));
        q.pkf = reinterpret_cast<int16_t *>(t + tab_offset(n, TAB_PKF));
        q.pkg = reinterpret_cast<int16_t *>(t + tab_offset(n, TAB_PKG));
        q.wscr = reinterpret_cast<int16_t *>(t + tab_offset(n, TAB_WSCR));
        q.wscr_lr = tab_sizes(n).wscr_lr;
        q.wtot4 = tab_sizes(n).wtot4;
        q.pmlev4 = reinterpret_cast<int32_t *>(t + tab_offset(n, TAB_PMLEV));
        q.plist = reinterpret_cast<int32_t *>(t + tab_offset(n, TAB_PLIST));
        q.pcum = reinterpret_cast<int32_t *>(t + tab_offset(n, TAB_PCUM));
        q.pmlist = reinterpret_cast<int32_t *>(t + tab_offset(n, TAB_PMLIST));
        q.pmstart = reinterpret_cast<int32_t *>(t + tab_offset(n, TAB_PMSTART));
        q.ftype_out = reinterpret_cast<int8_t *>(t + tab_offset(n, TAB_FTYPE));
        q.tb_stack = reinterpret_cast<int32_t *>(t + tab_offset(n, TAB_TBSTACK));
        q.tb_cap = 16 * n + 64;
        q.use_lists = generic_scan() ? 0 : 1;
        if (!ccj::fill4_tuned_supported(n)) {   // slim plan: the tuned-only buffers are 256-byte stubs -- nobody may follow them
            q.g1 = q.g2 = q.g3 = q.g4 = nullptr;
            q.lay = nullptr;
            q.scratch = q.plw = q.prw = q.pmw = q.pmm = q.pkf = q.pkg = q.wscr = nullptr;
            q.pmlev4 = q.plist = q.pcum = q.pmlist = q.pmstart = nullptr;
        }
    }
    ctx->h_seqs.assign(hd, hd + nseq);
    CU(cudaMemcpyAsync(d_in, ctx->h_stage, in_total, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_seqs, hd, (size_t)nseq * sizeof(ccj_seq), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->in_total = in_total;
    ctx->out_total = out_total;
    ctx->tab_total = tab_total;
    ctx->nmax = nmax;
    ctx->prepared = true;
    return 0;
}

int ccj_batch_fill(ccj_ctx *ctx) {
    if (!ctx) return CCJ_ERR_ARG;
    if (!ctx->prepared) return fail(ctx, CCJ_ERR_STATE, "ccj_batch_prepare was not called");
    ccj::NvtxRange nvtx("ccj_batch_fill");
    CU(cudaSetDevice(ctx->device));
    ccj::LaunchDims d;
    d.nseq = (int)ctx->plan.size();
    d.nmax = ctx->nmax;
    const std::pair<int, int> key(d.nmax, d.nseq * 2 + (use_tuned(d.nmax) ? 1 : 0));  // the env toggle changes the launch sequence
    auto it = ctx->graphs.find(key);
    if (it == ctx->graphs.end()) {
        // capture the per-level launch sequence once per wave shape
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        const cudaError_t qe = enqueue_fill(ctx, d);
        // always end the capture (also joins the side streams back), never keep a partial graph
        const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
        if (qe != cudaSuccess || ce != cudaSuccess) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            return fail(ctx, CCJ_ERR_CUDA, std::string("fill launch sequence: ") + cudaGetErrorString(qe != cudaSuccess ? qe : ce));
        }
        cudaGraphExec_t ge = nullptr;
        const cudaError_t ie = cudaGraphInstantiate(&ge, g, 0);
        cudaGraphDestroy(g);
        if (ie != cudaSuccess) return fail(ctx, CCJ_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie));
        if (ctx->graphs.size() > 32) drop_graphs(ctx);
        it = ctx->graphs.emplace(key, ge).first;
    }
    {   // a graph whose launch fails is not kept: the next call captures afresh
        cudaError_t le = cudaEventRecord(ctx->ev0, ctx->stream);
        if (le == cudaSuccess) le = cudaGraphLaunch(it->second, ctx->stream);
        if (le == cudaSuccess) le = cudaEventRecord(ctx->ev1, ctx->stream);
        if (le == cudaSuccess) le = cudaStreamSynchronize(ctx->stream);
        if (le == cudaSuccess) le = cudaGetLastError();
        if (le != cudaSuccess) {
            cudaGraphExecDestroy(it->second);
            ctx->graphs.erase(it);
            return fail(ctx, CCJ_ERR_CUDA, std::string("fill graph launch: ") + cudaGetErrorString(le));
        }
    }
    CU(cudaEventElapsedTime(&ctx->fill_ms, ctx->ev0, ctx->ev1));
    ctx->fill_launches = ccj::fill_launch_count(d.nmax, use_tuned(d.nmax));
    if (use_tuned(d.nmax)) {
        const int rc = refill_wrapped(ctx, d);
        if (rc) return rc;
    }
    ctx->filled = true;
    ctx->traced = false;
    return 0;
}

int ccj_batch_fill_profiled(ccj_ctx *ctx, float *kernel_ms) {
    if (!ctx || !kernel_ms) return CCJ_ERR_ARG;
    if (!ctx->prepared) return fail(ctx, CCJ_ERR_STATE, "ccj_batch_prepare was not called");
    CU(cudaSetDevice(ctx->device));
    ccj::LaunchDims d;
    d.nseq = (int)ctx->plan.size();
    d.nmax = ctx->nmax;
    const int nlaunch = ccj::fill_launch_count(d.nmax, true) + 8;
    struct EventSet {   // destroyed on every return path
        std::vector<cudaEvent_t> v;
        ~EventSet() { for (cudaEvent_t e : v) if (e) cudaEventDestroy(e); }
    } evs;
    evs.v.assign((size_t)nlaunch + 1, nullptr);
    std::vector<cudaEvent_t> &ev = evs.v;
    std::vector<int> kind;
    for (auto &e : ev) CU(cudaEventCreate(&e));
    size_t x = 0;
    cudaStream_t st = ctx->stream;
    CU(cudaEventRecord(ev[x++], st));
    cudaError_t mark_err = cudaSuccess;
    auto mark = [&](int k) {
        const cudaError_t e = cudaEventRecord(ev[x++], st);
        if (e != cudaSuccess && mark_err == cudaSuccess) mark_err = e;
        kind.push_back(k);
    };
    const bool tuned = use_tuned(d.nmax);
    ccj::launch_init(ctx->d_model, ctx->d_seqs, d, st); mark(3);
    if (tuned) { ccj::launch_prep(ctx->d_model, ctx->d_seqs, d, st); mark(3); }
    else if (!generic_scan()) { ccj::launch_prep_lists(ctx->d_model, ctx->d_seqs, d, st); mark(3); }
    // same order as enqueue_fill on one stream: the 2D tables run `lead` spans ahead of the levels
    const int lead = tuned ? std::max(ccj::fill4_fused_levels() - 2, 0) : 0;
    auto span_step = [&](int sp) {
        if (sp >= d.nmax) return;
        if (sp >= 3 && sp <= d.nmax - 1) {
            if (tuned) ccj::launch_P_tuned(ctx->d_model, ctx->d_seqs, d, sp, st);
            else ccj::launch_P(ctx->d_model, ctx->d_seqs, d, sp, st);
            mark(1);
        }
        ccj::launch_2d(ctx->d_model, ctx->d_seqs, d, sp, st); mark(2);
    };
    for (int sp = 0; sp < lead; ++sp) span_step(sp);
    for (int s = 0; s < d.nmax; ++s) {
        span_step(s + lead);
        if (d.nmax - s - 2 >= 1) {
            if (tuned) {
                ccj::launch_4d_roles(ctx->d_model, ctx->d_seqs, d, s, st); mark(0);
                ccj::launch_4d_windows(ctx->d_model, ctx->d_seqs, d, s, st); mark(4);
                ccj::launch_4d_final(ctx->d_model, ctx->d_seqs, d, s, st); mark(5);
            } else {
                ccj::launch_4d(ctx->d_model, ctx->d_seqs, d, s, st); mark(0);
            }
        }
    }
    ccj::launch_W(ctx->d_model, ctx->d_seqs, d, st); mark(3);
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    CU(mark_err);
    for (int k = 0; k < 6; ++k) kernel_ms[k] = 0.f;
    for (size_t y = 0; y < kind.size(); ++y) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, ev[y], ev[y + 1]));
        kernel_ms[kind[y]] += ms;
    }
    CU(cudaEventElapsedTime(&ctx->fill_ms, ev[0], ev[x - 1]));
    ctx->fill_launches = (int)kind.size();
    ctx->filled = true;
    ctx->traced = false;
    return 0;
}

int ccj_batch_traceback(ccj_ctx *ctx) {
    if (!ctx) return CCJ_ERR_ARG;
    if (!ctx->filled) return fail(ctx, CCJ_ERR_STATE, "ccj_batch_fill was not called");
    ccj::NvtxRange nvtx("ccj_batch_traceback");
    CU(cudaSetDevice(ctx->device));
    ccj::LaunchDims d;
    d.nseq = (int)ctx->plan.size();
    d.nmax = ctx->nmax;
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    ccj::launch_traceback(ctx->d_model, ctx->d_seqs, d, ctx->stream);
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    CU(cudaEventElapsedTime(&ctx->tb_ms, ctx->ev0, ctx->ev1));
    ctx->traced = true;
    return 0;
}

int ccj_batch_fetch(ccj_ctx *ctx, ccj_result *results, int32_t *pairs, char *structs) {
    if (!ctx || !results) return CCJ_ERR_ARG;
    if (!ctx->traced) return fail(ctx, CCJ_ERR_STATE, "ccj_batch_traceback was not called");
    ccj::NvtxRange nvtx("ccj_batch_fetch");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(ctx->h_stage, ctx->d_arena + ctx->in_total, ctx->out_total, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    size_t off = 0;
    for (size_t s = 0; s < ctx->plan.size(); ++s) {
        const SeqPlan &p = ctx->plan[s];
        const int n = p.n;
        const int32_t *st = reinterpret_cast<const int32_t *>(ctx->h_stage + p.out_off);
        const int32_t *W = st + CCJ_STATUS_INTS;
        const int32_t *pr = W + (n + 1);
        // k_prep met an interior-loop energy outside int16: the packed window lists cannot hold this model (the generic
        // cell functions notice the flag themselves and scan the windows instead)
        if (st[5] != 0 && use_tuned(ctx->nmax))
            return fail(ctx, CCJ_ERR_STATE, "energy model outside the range of the tuned kernels (set CCJ_FILL_GENERIC=1)");
        ccj_result &r = results[s];
        r.energy_dcal = W[n];
        r.status = st[0];
        r.n_should_not_be_here = st[1];
        r.msg_id = st[2];
        r.aux_i = st[3];
        r.aux_j = st[4];
        if (pairs)
            for (int x = 0; x < n; ++x) pairs[off + x] = pr[x + 1];
        if (structs) {
            const std::string sstr = ccj::fill_structure(n, pr);
            memcpy(structs + off, sstr.data(), n);
        }
        off += n;
    }
    return 0;
}

float ccj_last_fill_ms(const ccj_ctx *ctx) { return ctx ? ctx->fill_ms : 0.f; }
float ccj_last_traceback_ms(const ccj_ctx *ctx) { return ctx ? ctx->tb_ms : 0.f; }
int ccj_last_fill_launches(const ccj_ctx *ctx) { return ctx ? ctx->fill_launches : 0; }

int ccj_fold_batch(ccj_ctx *ctx, const char *seqs, const int64_t *offsets, int nseq, ccj_result *results,
                   int32_t *pairs, char *structs) {
    if (!ctx || !seqs || !offsets || nseq < 1 || !results) return CCJ_ERR_ARG;
    if (!ctx->model_ok) return fail(ctx, CCJ_ERR_STATE, "no energy model loaded");
    CU(cudaSetDevice(ctx->device));
    const size_t budget = budget_bytes(ctx);
    int s0 = 0;
    float fill_ms = 0.f, tb_ms = 0.f;
    int launches = 0;
    while (s0 < nseq) {
        // greedy wave: consecutive sequences while they fit
        size_t bytes = 1024;
        int s1 = s0;
        while (s1 < nseq) {
            const int64_t len = offsets[s1 + 1] - offsets[s1];
            int rc = validate(ctx, seqs + offsets[s1], len);
            if (rc) return rc;
            SeqPlan p;
            plan_seq((int)len, p);
            const size_t add = p.in_bytes + p.out_bytes + p.tab_bytes;
            if (bytes + add > budget || s1 - s0 >= kMaxWave) break;
            bytes += add;
            ++s1;
        }
        if (s1 == s0) return fail(ctx, CCJ_ERR_TOO_LARGE, "a sequence does not fit this GPU's memory");
        const int64_t base = offsets[s0];
        std::vector<int64_t> off(s1 - s0 + 1);
        for (int s = s0; s <= s1; ++s) off[s - s0] = offsets[s] - base;
        int rc = ccj_batch_prepare(ctx, seqs + base, off.data(), s1 - s0);
        if (!rc) rc = ccj_batch_fill(ctx);
        if (!rc) rc = ccj_batch_traceback(ctx);
        if (!rc) rc = ccj_batch_fetch(ctx, results + s0, pairs ? pairs + base : nullptr, structs ? structs + base : nullptr);
        if (rc) return rc;
        fill_ms += ctx->fill_ms;
        tb_ms += ctx->tb_ms;
        launches += ctx->fill_launches;
        s0 = s1;
    }
    ctx->fill_ms = fill_ms;
    ctx->tb_ms = tb_ms;
    ctx->fill_launches = launches;
    return 0;
}

// The same over several GPUs of one box inside one process: one host thread per context, sequences dealt dynamically
// in chunks from a shared counter (a GPU that runs slower simply draws fewer chunks).  No collective: sequences are
// independent.  Results land in the caller's arrays at the sequences' own positions.
int ccj_fold_batch_multi(ccj_ctx **ctxs, int nctx, const char *seqs, const int64_t *offsets, int nseq, ccj_result *results,
                         int32_t *pairs, char *structs) {
    if (!ctxs || nctx < 1 || !seqs || !offsets || nseq < 1 || !results) return CCJ_ERR_ARG;
    for (int c = 0; c < nctx; ++c)
        if (!ctxs[c]) return CCJ_ERR_ARG;
    if (nctx == 1) return ccj_fold_batch(ctxs[0], seqs, offsets, nseq, results, pairs, structs);
    int nmax = 1;
    for (int s = 0; s < nseq; ++s) nmax = std::max<int64_t>(nmax, offsets[s + 1] - offsets[s]);
    // chunk: at most one wave of the longest sequence, and small enough that every GPU draws several times
    int64_t chunk = std::max<int64_t>(1, (nseq + 4 * nctx - 1) / (4 * nctx));
    for (int c = 0; c < nctx; ++c) chunk = std::max<int64_t>(1, std::min<int64_t>(chunk, ccj_wave_capacity(ctxs[c], nmax)));
    std::atomic<int64_t> next(0);
    std::vector<int> rcs(nctx, 0);
    std::vector<float> fill(nctx, 0.f), tb(nctx, 0.f);
    std::vector<int> launches(nctx, 0);
    std::vector<std::thread> th;
    for (int c = 0; c < nctx; ++c)
        th.emplace_back([&, c] {
            ccj_ctx *ctx = ctxs[c];
            for (;;) {
                const int64_t lo = next.fetch_add(chunk);
                if (lo >= nseq) break;
                const int64_t hi = std::min<int64_t>(lo + chunk, nseq);
                const int64_t base = offsets[lo];
                std::vector<int64_t> off(hi - lo + 1);
                for (int64_t s = lo; s <= hi; ++s) off[s - lo] = offsets[s] - base;
                const int rc = ccj_fold_batch(ctx, seqs + base, off.data(), (int)(hi - lo), results + lo,
                                              pairs ? pairs + base : nullptr, structs ? structs + base : nullptr);
                if (rc) {
                    rcs[c] = rc;
                    next.store(nseq);   // stop the other threads at their next draw
                    break;
                }
                fill[c] += ctx->fill_ms;
                tb[c] += ctx->tb_ms;
                launches[c] += ctx->fill_launches;
            }
        });
    for (auto &t : th) t.join();
    for (int c = 0; c < nctx; ++c) {
        ctxs[c]->fill_ms = fill[c];
        ctxs[c]->tb_ms = tb[c];
        ctxs[c]->fill_launches = launches[c];
        if (rcs[c]) {
            if (c != 0) ctxs[0]->err = ctxs[c]->err;
            return rcs[c];
        }
    }
    return 0;
}

int64_t ccj_table4_len(int n) { return ccj_cells4(n); }
int64_t ccj_table2_len(int n) { return (int64_t)n * (n + 1) / 2; }

int ccj_export_table4(ccj_ctx *ctx, int seq_index, int table, int16_t *out, int64_t out_len) {
    if (!ctx || !out || table < 0 || table >= CCJ_NT4) return CCJ_ERR_ARG;
    if (!ctx->filled || seq_index < 0 || seq_index >= (int)ctx->plan.size())
        return fail(ctx, CCJ_ERR_STATE, "no filled wave / bad sequence index");
    const SeqPlan &p = ctx->plan[seq_index];
    const int n = p.n;
    const int64_t cells = ccj_cells4(n);
    if (out_len < cells) return fail(ctx, CCJ_ERR_ARG, "output buffer too small");
    CU(cudaSetDevice(ctx->device));
    std::vector<int16_t> raw((size_t)cells);
    const char *d_tab = ctx->d_arena + ctx->in_total + ctx->out_total + p.tab_off;
    CU(cudaMemcpy(raw.data(), d_tab + (size_t)table * cells * sizeof(int16_t), (size_t)cells * sizeof(int16_t),
                  cudaMemcpyDeviceToHost));
    int64_t x = 0;
    for (int i = 1; i <= n; ++i)
        for (int j = i; j <= n; ++j)
            for (int k = j + 2; k <= n; ++k)
                for (int l = k; l <= n; ++l) out[x++] = raw[(size_t)ccj_idx4(n, i, j, k, l)];
    return 0;
}

int ccj_export_table2(ccj_ctx *ctx, int seq_index, int table, int32_t *out, int64_t out_len) {
    if (!ctx || !out || table < 0 || table >= CCJ_NT2) return CCJ_ERR_ARG;
    if (!ctx->filled || seq_index < 0 || seq_index >= (int)ctx->plan.size())
        return fail(ctx, CCJ_ERR_STATE, "no filled wave / bad sequence index");
    const SeqPlan &p = ctx->plan[seq_index];
    const int n = p.n;
    if (out_len < (int64_t)n * (n + 1) / 2) return fail(ctx, CCJ_ERR_ARG, "output buffer too small");
    CU(cudaSetDevice(ctx->device));
    const int64_t s2 = ccj_stride2(n);
    std::vector<int32_t> raw((size_t)s2);
    const char *d_tab = ctx->d_arena + ctx->in_total + ctx->out_total + p.tab_off;
    const size_t t4b = tab_offset(n, TAB_T2);
    CU(cudaMemcpy(raw.data(), d_tab + t4b + (size_t)table * s2 * sizeof(int32_t), (size_t)s2 * sizeof(int32_t),
                  cudaMemcpyDeviceToHost));
    int64_t x = 0;
    for (int i = 1; i <= n; ++i)
        for (int j = i; j <= n; ++j) out[x++] = raw[(size_t)ccj_idx2(n, i, j)];
    return 0;
}

int ccj_table4_get(ccj_ctx *ctx, int seq_index, int table, int i, int j, int k, int l, int32_t *value) {
    if (!ctx || !value || table < 0 || table >= CCJ_NT4) return CCJ_ERR_ARG;
    if (!ctx->filled || seq_index < 0 || seq_index >= (int)ctx->plan.size())
        return fail(ctx, CCJ_ERR_STATE, "no filled wave / bad sequence index");
    const SeqPlan &p = ctx->plan[seq_index];
    const int n = p.n;
    if (!ccj_valid4(i, j, k, l) || i < 1 || l > n) {
        *value = CCJ_INF;
        return 0;
    }
    CU(cudaSetDevice(ctx->device));
    const char *d_tab = ctx->d_arena + ctx->in_total + ctx->out_total + p.tab_off;
    int16_t v = 0;
    CU(cudaMemcpy(&v, d_tab + ((size_t)table * ccj_cells4(n) + (size_t)ccj_idx4(n, i, j, k, l)) * sizeof(int16_t),
                  sizeof v, cudaMemcpyDeviceToHost));
    *value = v;
    return 0;
}

int ccj_table2_get(ccj_ctx *ctx, int seq_index, int table, int i, int j, int32_t *value) {
    if (!ctx || !value || table < 0 || table >= CCJ_NT2) return CCJ_ERR_ARG;
    if (!ctx->filled || seq_index < 0 || seq_index >= (int)ctx->plan.size())
        return fail(ctx, CCJ_ERR_STATE, "no filled wave / bad sequence index");
    const SeqPlan &p = ctx->plan[seq_index];
    const int n = p.n;
    if (i < 1 || j > n || i > j) return fail(ctx, CCJ_ERR_ARG, "need 1 <= i <= j <= n");
    CU(cudaSetDevice(ctx->device));
    const char *d_tab = ctx->d_arena + ctx->in_total + ctx->out_total + p.tab_off + tab_offset(n, TAB_T2);
    CU(cudaMemcpy(value, d_tab + ((size_t)table * ccj_stride2(n) + (size_t)ccj_idx2(n, i, j)) * sizeof(int32_t),
                  sizeof(int32_t), cudaMemcpyDeviceToHost));
    return 0;
}

static void fnv_add(uint64_t &h, uint64_t v) {
    h ^= v;
    h *= 1099511628211ULL;
}

int ccj_table4_hash(ccj_ctx *ctx, int seq_index, int table, uint64_t *hash, int64_t *finite, int32_t *min_value) {
    if (!ctx || !hash) return CCJ_ERR_ARG;
    if (!ctx->filled || seq_index < 0 || seq_index >= (int)ctx->plan.size())
        return fail(ctx, CCJ_ERR_STATE, "no filled wave / bad sequence index");
    const int64_t len = ccj_cells4(ctx->plan[seq_index].n);
    std::vector<int16_t> buf((size_t)len + 1);
    int rc = ccj_export_table4(ctx, seq_index, table, buf.data(), len);
    if (rc) return rc;
    uint64_t h = 1469598103934665603ULL;
    int64_t fin = 0;
    int32_t mn = 1 << 30;
    for (int64_t x = 0; x < len; ++x) {
        fnv_add(h, (uint16_t)buf[x]);
        if (buf[x] < 32767) {
            ++fin;
            if (buf[x] < mn) mn = buf[x];
        }
    }
    *hash = h;
    if (finite) *finite = fin;
    if (min_value) *min_value = fin ? mn : 0;
    return 0;
}

int ccj_table2_hash(ccj_ctx *ctx, int seq_index, int table, uint64_t *hash, int64_t *finite, int64_t *sum) {
    if (!ctx || !hash) return CCJ_ERR_ARG;
    if (!ctx->filled || seq_index < 0 || seq_index >= (int)ctx->plan.size())
        return fail(ctx, CCJ_ERR_STATE, "no filled wave / bad sequence index");
    const int n = ctx->plan[seq_index].n;
    const int64_t len = (int64_t)n * (n + 1) / 2;
    std::vector<int32_t> buf((size_t)len + 1);
    int rc = ccj_export_table2(ctx, seq_index, table, buf.data(), len);
    if (rc) return rc;
    uint64_t h = 1469598103934665603ULL;
    int64_t fin = 0, sm = 0;
    for (int64_t x = 0; x < len; ++x) {
        fnv_add(h, (uint32_t)buf[x]);
        if (buf[x] < CCJ_INF / 2) {
            ++fin;
            sm += buf[x];
        }
    }
    *hash = h;
    if (finite) *finite = fin;
    if (sum) *sum = sm;
    return 0;
}

// SURVEY.md App. D term model + exact can_pair-gated window counts (rows a9-a23 of 8a)
int ccj_count_terms(const char *seq, int n, int no_gu, int64_t *out) {
    if (!seq || n < 1 || !out) return CCJ_ERR_ARG;
    auto m0 = [](int64_t x) { return x > 0 ? x : (int64_t)0; };
    std::vector<int> S(n + 2);
    for (int i = 1; i <= n; ++i) S[i] = ccj::encode_base(seq[i - 1]);
    static const int bp[5][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 5}, {0, 0, 0, 1, 0}, {0, 0, 2, 0, 3}, {0, 6, 0, 4, 0}};
    auto pt = [&](int i, int j) {
        int t = bp[S[i]][S[j]];
        if (no_gu && (t == 3 || t == 4)) t = 0;
        return t;
    };
    auto cp = [&](int i, int j) { return j - i > 3 && pt(i, j) > 0; };
    int64_t cells = 0, split = 0, pterms = 0, iloop = 0;
    // weights: number of (i,gap) placements of arms (a,b): w = sum_{g=2}^{n-1-a-b} (n-a-g-b)
    for (int a = 0; a <= n - 3; ++a)
        for (int b = 0; a + b <= n - 3; ++b) {
            const int64_t m = n - a - b - 2, w = m * (m + 1) / 2;
            const int64_t a1 = m0(a - 1), b1 = m0(b - 1);
            const int64_t sp = (1 + 2 * a) + a + (a + a1) + (1 + 2 * b) + (1 + b) + (1 + b) + (1 + a + b) + (1 + b) +
                               (1 + a + b1) + (1 + a + b) + b + (a + b1) + (2 * a1 + 3) + (2 * b1 + 2) + a1 + b1 +
                               (a1 + b1 + 2) + (a1 + b1 + 4) + 12;
            cells += w;
            split += w * sp;
        }
    for (int s = 3; s <= n - 1; ++s) pterms += (int64_t)(n - s) * ((int64_t)s * (s - 1) * (s - 2) / 6);
    // PL windows depend on (i,j) only; PR on (k,l) only
    for (int i = 1; i <= n; ++i)
        for (int j = i; j <= n; ++j) {
            if (!(pt(i, j) > 0 && cp(i, j))) continue;
            int64_t cnt = 0;
            const int max_d = std::min(j, i + 30);
            for (int d = i + 1; d < max_d; ++d)
                for (int dp = j - 1; dp > std::max(d + 3, j - 30); --dp) cnt += cp(d, dp);
            const int64_t right = (int64_t)(n - j - 1) * (n - j) / 2;  // (k,l) with k>=j+2, k<=l<=n
            const int64_t left = (int64_t)(i - 2) * (i - 1) / 2;       // (i',j') with i'<=j'<=i-2
            iloop += cnt * (m0(right) + m0(left));
        }
    // PM windows: pair (j,k), d in (max(i,j-30), j), dp in (k, min(l,k+30))
    for (int j = 1; j <= n; ++j)
        for (int k = j + 2; k <= n; ++k) {
            if (!(pt(j, k) > 0 && cp(j, k))) continue;
            // cnt[x][y] = pairs with d >= j-x, dp <= k+y
            int64_t pre[31][31];
            for (int x = 0; x <= 30; ++x)
                for (int y = 0; y <= 30; ++y) {
                    int64_t v = 0;
                    if (x > 0 && y > 0) {
                        const int d = j - x, dp = k + y;
                        v = pre[x - 1][y] + pre[x][y - 1] - pre[x - 1][y - 1];
                        if (d >= 1 && dp <= n && x <= 29 && y <= 29) v += cp(d, dp);
                    }
                    pre[x][y] = v;
                }
            // d > max(i, j-30) and dp < min(l, k+30) (:763-766): x = j-d <= min(j-i-1, 29), y = dp-k <= min(l-k-1, 29)
            for (int i = 1; i < j; ++i)
                for (int l = k + 1; l <= n; ++l) {
                    const int x = std::min(j - i - 1, 29), y = std::min(l - k - 1, 29);
                    iloop += pre[x][y];
                }
        }
    out[0] = cells;
    out[1] = split;
    out[2] = pterms;
    out[3] = iloop;
    return 0;
}

// ---- what the C++ class shells (ccj_b200/csrc/{W_final,pseudo_loop,s_energy_matrix}.hh) need beyond a whole fold ----
int ccj_model_build(const char *par_file, int dangles, int no_gu, void *model_out, size_t model_bytes, char *err, size_t err_len) {
    if (!par_file || !model_out || model_bytes != sizeof(ccj_model)) return CCJ_ERR_ARG;
    ccj::RawParams *rp = new ccj::RawParams();
    std::string e;
    if (!read_params(par_file, par_file[0] == '@' ? par_file + 1 : nullptr, *rp, e)) {
        if (err && err_len) snprintf(err, err_len, "%s", e.c_str());
        delete rp;
        return CCJ_ERR_PARAMS;
    }
    for (const std::string &w : rp->warnings) fprintf(stderr, "WARNING: %s\n", w.c_str());
    ccj::build_model(*rp, dangles, no_gu, *static_cast<ccj_model *>(model_out));
    delete rp;
    return 0;
}

int ccj_model_upload(ccj_ctx *ctx, const void *model, size_t model_bytes) {
    if (!ctx || !model || model_bytes != sizeof(ccj_model)) return CCJ_ERR_ARG;
    memcpy(ctx->h_model, model, sizeof(ccj_model));
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(ctx->d_model, ctx->h_model, sizeof(ccj_model), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->model_ok = true;
    return 0;
}

size_t ccj_model_bytes(void) { return sizeof(ccj_model); }

int ccj_copy_table4_raw(ccj_ctx *ctx, int seq_index, int table, int16_t *out, int64_t out_len) {
    if (!ctx || !out || table < 0 || table >= CCJ_NT4) return CCJ_ERR_ARG;
    if (!ctx->filled || seq_index < 0 || seq_index >= (int)ctx->plan.size())
        return fail(ctx, CCJ_ERR_STATE, "no filled wave / bad sequence index");
    const SeqPlan &p = ctx->plan[seq_index];
    const int64_t cells = ccj_cells4(p.n);
    if (out_len < cells) return fail(ctx, CCJ_ERR_ARG, "output buffer too small");
    CU(cudaSetDevice(ctx->device));
    const char *d_tab = ctx->d_arena + ctx->in_total + ctx->out_total + p.tab_off;
    CU(cudaMemcpy(out, d_tab + (size_t)table * cells * sizeof(int16_t), (size_t)cells * sizeof(int16_t), cudaMemcpyDeviceToHost));
    return 0;
}

int ccj_copy_tables2_raw(ccj_ctx *ctx, int seq_index, int32_t *out, int64_t out_len) {
    if (!ctx || !out) return CCJ_ERR_ARG;
    if (!ctx->filled || seq_index < 0 || seq_index >= (int)ctx->plan.size())
        return fail(ctx, CCJ_ERR_STATE, "no filled wave / bad sequence index");
    const SeqPlan &p = ctx->plan[seq_index];
    const int64_t len = ccj_stride2(p.n) * CCJ_NT2;
    if (out_len < len) return fail(ctx, CCJ_ERR_ARG, "output buffer too small");
    CU(cudaSetDevice(ctx->device));
    const char *d_tab = ctx->d_arena + ctx->in_total + ctx->out_total + p.tab_off + tab_offset(p.n, TAB_T2);
    CU(cudaMemcpy(out, d_tab, (size_t)len * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return 0;
}

int ccj_traceback_step(ccj_ctx *ctx, int seq_index, const int32_t *node, int32_t *pushed, int32_t cap, int32_t *n_pushed,
                       int32_t *status3) {
    if (!ctx || !node || !pushed || !n_pushed || cap < 0) return CCJ_ERR_ARG;
    if (!ctx->filled || seq_index < 0 || seq_index >= (int)ctx->plan.size())
        return fail(ctx, CCJ_ERR_STATE, "no filled wave / bad sequence index");
    CU(cudaSetDevice(ctx->device));
    const SeqPlan &p = ctx->plan[seq_index];
    int *d_top = nullptr;
    CU(cudaMalloc((void **)&d_top, sizeof(int)));
    ccj::launch_tb_step(ctx->d_model, ctx->d_seqs, seq_index, node, d_top, ctx->stream);
    int top = 0;
    cudaError_t e = cudaMemcpyAsync(&top, d_top, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(d_top);
    CU(e);
    if (top > cap) return fail(ctx, CCJ_ERR_ARG, "pushed-node buffer too small");
    const char *d_tab = ctx->d_arena + ctx->in_total + ctx->out_total + p.tab_off;
    if (top > 0)
        CU(cudaMemcpy(pushed, d_tab + tab_offset(p.n, TAB_TBSTACK), (size_t)top * 5 * sizeof(int32_t), cudaMemcpyDeviceToHost));
    *n_pushed = top;
    if (status3) {
        int32_t st[CCJ_STATUS_INTS];
        CU(cudaMemcpy(st, ctx->d_arena + ctx->in_total + p.out_off, sizeof st, cudaMemcpyDeviceToHost));
        status3[0] = st[0];
        status3[1] = st[2];
        status3[2] = st[1];
    }
    return 0;
}

int ccj_fetch_fold_state(ccj_ctx *ctx, int seq_index, int32_t *pair, int8_t *type) {
    if (!ctx || !pair) return CCJ_ERR_ARG;
    if (!ctx->filled || seq_index < 0 || seq_index >= (int)ctx->plan.size())
        return fail(ctx, CCJ_ERR_STATE, "no filled wave / bad sequence index");
    CU(cudaSetDevice(ctx->device));
    const SeqPlan &p = ctx->plan[seq_index];
    const int n = p.n;
    const char *o = ctx->d_arena + ctx->in_total + p.out_off;
    CU(cudaMemcpy(pair, o + sizeof(int32_t) * (CCJ_STATUS_INTS + (size_t)(n + 1)), sizeof(int32_t) * (size_t)(n + 1), cudaMemcpyDeviceToHost));
    if (type) {
        const char *d_tab = ctx->d_arena + ctx->in_total + ctx->out_total + p.tab_off;
        CU(cudaMemcpy(type, d_tab + tab_offset(n, TAB_FTYPE), (size_t)(n + 1), cudaMemcpyDeviceToHost));
    }
    return 0;
}

int64_t ccj_layout_index(int n, int i, int j, int k, int l) {
    if (n < 1 || i < 1 || l > n || !ccj_valid4(i, j, k, l)) return -1;
    return ccj_idx4(n, i, j, k, l);
}

int ccj_model_text(const char *par_file, int dangles, int no_gu, const char *out_path, char *err, size_t err_len) {
    if (!par_file || !out_path) return CCJ_ERR_ARG;
    ccj::RawParams *rp = new ccj::RawParams();
    std::string e;
    // "@name" selects a parameter set linked into the library (ccj_model_load_embedded)
    if (!read_params(par_file, par_file[0] == '@' ? par_file + 1 : nullptr, *rp, e)) {
        if (err && err_len) snprintf(err, err_len, "%s", e.c_str());
        delete rp;
        return CCJ_ERR_PARAMS;
    }
    for (const std::string &w : rp->warnings) fprintf(stderr, "WARNING: %s\n", w.c_str());
    ccj_model *m = new ccj_model();
    ccj::build_model(*rp, dangles, no_gu, *m);
    FILE *f = fopen(out_path, "w");
    if (!f) {
        delete rp;
        delete m;
        return CCJ_ERR_ARG;
    }
#define D2(name, l0, l1) for (int a = 0; a < l0; ++a) for (int b = 0; b < l1; ++b) fprintf(f, #name " %d %d %d\n", a, b, m->name[a][b]);
#define D3(name, l0, l1, l2) for (int a = 0; a < l0; ++a) for (int b = 0; b < l1; ++b) for (int c = 0; c < l2; ++c) fprintf(f, #name " %d %d %d %d\n", a, b, c, m->name[a][b][c]);
    D2(stack, 8, 8)
    for (int a = 0; a < 31; ++a) fprintf(f, "hairpin %d %d\n", a, m->hairpin[a]);
    for (int a = 0; a < 31; ++a) fprintf(f, "bulge %d %d\n", a, m->bulge[a]);
    for (int a = 0; a < 31; ++a) fprintf(f, "internal_loop %d %d\n", a, m->internal_loop[a]);
    D3(mismatchExt, 8, 5, 5) D3(mismatchI, 8, 5, 5) D3(mismatch1nI, 8, 5, 5) D3(mismatch23I, 8, 5, 5)
    D3(mismatchH, 8, 5, 5) D3(mismatchM, 8, 5, 5) D2(dangle5, 8, 5) D2(dangle3, 8, 5)
    for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b) for (int c = 0; c < 5; ++c) for (int d = 0; d < 5; ++d)
        fprintf(f, "int11 %d %d %d %d %d\n", a, b, c, d, m->int11[a][b][c][d]);
    for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b) for (int c = 0; c < 5; ++c) for (int d = 0; d < 5; ++d)
        for (int g = 0; g < 5; ++g) fprintf(f, "int21 %d %d %d %d %d %d\n", a, b, c, d, g, m->int21[a][b][c][d][g]);
    for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b) for (int c = 0; c < 5; ++c) for (int d = 0; d < 5; ++d)
        for (int g = 0; g < 5; ++g) for (int h = 0; h < 5; ++h)
            fprintf(f, "int22 %d %d %d %d %d %d %d\n", a, b, c, d, g, h, m->int22[a][b][c][d][g][h]);
#undef D2
#undef D3
    fprintf(f, "ninio 2 %d\n", m->ninio2);
    fprintf(f, "MLbase %d\n", m->MLbase);
    for (int a = 0; a < 8; ++a) fprintf(f, "MLintern %d %d\n", a, m->MLintern[a]);
    fprintf(f, "MLclosing %d\nTerminalAU %d\n", m->MLclosing, m->TerminalAU);
    for (int a = 0; a < m->n_tetra; ++a) if (m->tetra[a][0]) fprintf(f, "Tetraloop_E %.6s %d\n", m->tetra[a], m->tetra_E[a]);
    for (int a = 0; a < m->n_tri; ++a) if (m->tri[a][0]) fprintf(f, "Triloop_E %.5s %d\n", m->tri[a], m->tri_E[a]);
    for (int a = 0; a < m->n_hexa; ++a) if (m->hexa[a][0]) fprintf(f, "Hexaloop_E %.8s %d\n", m->hexa[a], m->hexa_E[a]);
    fprintf(f, "special_hp %d\ndangles %d\n", m->special_hp, m->dangles);
    fclose(f);
    delete rp;
    delete m;
    return 0;
}

}  // extern "C"
