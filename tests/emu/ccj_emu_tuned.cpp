// TEST TOOL (not part of the product): runs the CUDA kernels of ccj_b200/csrc/ccj_fill4.cu (tuned fill: k_prep_lay,
// k_fill_pmw, k_prep, k_roles, k_winLR, k_winM, k_final, k_P_tuned) and of ccj_b200/csrc/ccj_kernels.cu (k_init, k_2d, k_W,
// k_traceback; and with path=generic the one-thread-per-cell fill k_P_lean + k_4d_lean, or k_P + k_4d) on the host,
// compiled by g++ for the SIMT emulator of simt_emu.hpp, in the launch order of ccj_abi.cu (ccj_batch_fill_profiled:
// one stream) and with the launchers' grids.
// Buffers have exactly the sizes ccj_abi.cu plans (tab_bytes_uncached) and are poisoned before the fill, so the tool is
// meaningful under AddressSanitizer / UBSan (out-of-bounds, misaligned vector accesses) and ThreadSanitizer (two threads
// of a block touching the same word without a barrier) -- the stand-in for compute-sanitizer, which is closed on the GPU
// pool.
//
//   ccj_emu_tuned hash <parfile> <dangles> <seq> [noGU] [pipe] [path]  -> same text as `ccj_ref_dump hash` / `ccj_emu hash`
//   ccj_emu_tuned fold <parfile> <dangles> <seq> [noGU] [pipe] [path]  -> stdout / stderr / exit code of the CCJ binary
//   order=<issue|side|main|prio:ABC|rnd:SEED> as 8th argument (tuned path only, needs -DCCJ_ENQUEUE_FILL_INC=<file>): the launch
//         sequence is then NOT restated here -- the text of enqueue_fill() of ccj_abi.cu (extracted by the test into that
//         file) runs against mock streams / events, which turns its launches into a dependency graph (stream order +
//         event record -> wait edges), and the kernels are executed in a legal order of that graph: as issued, side
//         streams (windows, P + 2D) as far ahead as the events allow, main stream first, or random.  Every legal order must
//         give the same tables: a missing dependency of the three-stream graph shows up as a wrong table here.
//   pipe: -1 = the launcher's choice (small waves: software-pipelined window kernels), 0 / 1 = force
//   path: tuned (default) | lean (k_P_lean + k_4d_lean) | generic (k_P + k_4d) | shardG (the row-sharded fold of
//         ccj_shard.cu for G ranks: k_P_shard_lean + k_4d_shard_lean per rank and level; the ranks share the replicated
//         region, which is the state the per-level allgather establishes); suffix +2d: k_init / k_2d / k_W as kernels
#define CCJ_HOST_EMU 1
#include "simt_emu.hpp"

#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../ccj_b200/csrc/energy_model.hpp"
#include "../../ccj_b200/csrc/ccj_fill4.cu"
#include "../../ccj_b200/csrc/ccj_kernels.cu"
#include "../../ccj_b200/csrc/ccj_render.hpp"
namespace {
#include "../../ccj_b200/csrc/ccj_shard_kernels.cuh"   // inside an unnamed namespace, as ccj_shard.cu includes it
}

struct Fnv {
    uint64_t h = 1469598103934665603ULL;
    void add(uint64_t v) { h ^= v; h *= 1099511628211ULL; }
};
static const char *k4dNames[22] = {
    "PK", "PL", "PR", "PM", "PO", "PfromL", "PfromR", "PfromM", "PfromMprime", "PfromO",
    "PLmloop00", "PLmloop01", "PLmloop10", "PRmloop00", "PRmloop01", "PRmloop10",
    "PMmloop00", "PMmloop01", "PMmloop10", "POmloop00", "POmloop01", "POmloop10"};
static const char *k2dNames[8] = {"V", "Vtype", "WM", "WMv", "WMp", "P", "WBP", "WPP"};

// one exactly sized, poisoned heap block per device buffer (sizes: tab_bytes_uncached of ccj_abi.cu without its
// 256-byte rounding, so that the sanitizer sees the first byte past what the plan promises)
// CCJ_EMU_POISON=<byte, hex>: what uninitialised device memory holds (default 55: 0x5555 = 21845 per int16; 80 makes every
// stale entry -32640, which would win any minimum it leaked into; 7f a large finite energy)
static int poison_byte() {
    static const int v = [] { const char *e = getenv("CCJ_EMU_POISON"); return e ? (int)strtol(e, nullptr, 16) & 0xff : 0x55; }();
    return v;
}
template <class T>
static T *buf(std::vector<std::vector<char>> &keep, size_t bytes, int poison = -1) {
    if (poison < 0) poison = poison_byte();
    keep.emplace_back(bytes ? bytes : 1, (char)poison);
    return reinterpret_cast<T *>(keep.back().data());
}

// all device buffers of one sequence, exactly sized and poisoned (ccj_abi.cu: ccj_batch_prepare)
static ccj_seq make_seq(const std::string &seq, std::vector<std::vector<char>> &keep) {
    const int n = (int)seq.size();
    int8_t *S = buf<int8_t>(keep, (size_t)n + 2, 0);
    for (int i = 1; i <= n; ++i) S[i] = (int8_t)ccj::encode_base(seq[i - 1]);
    S[n + 1] = S[1];
    S[0] = S[n];
    const size_t tri = (size_t)n * (n - 1) / 2 + 1;
    const int64_t cells = ccj_cells4(n), s2 = ccj_stride2(n);
    const int64_t wscr_lr = ccj_winlr_level_max(n);
    const int32_t wtot4 = (int32_t)ccj_pmw_level_quads(n);

    ccj_seq q;
    memset(&q, 0, sizeof q);
    q.n = n;
    q.S = S;
    q.seq = seq.c_str();   // the caller keeps the string alive
    q.status = buf<int32_t>(keep, sizeof(int32_t) * CCJ_STATUS_INTS, 0);
    q.W = buf<int32_t>(keep, sizeof(int32_t) * (size_t)(n + 1), 0);
    q.pair_out = buf<int32_t>(keep, sizeof(int32_t) * (size_t)(n + 2), 0);
    q.t4 = buf<int16_t>(keep, (size_t)cells * CCJ_NT4_STORE * sizeof(int16_t) + 16);
    q.stride4 = cells;
    q.g1 = buf<int16_t>(keep, (size_t)cells * 6 * sizeof(int16_t) + 64);
    q.g2 = buf<int16_t>(keep, (size_t)cells * 6 * sizeof(int16_t) + 64);
    q.g3 = buf<int16_t>(keep, (size_t)cells * 6 * sizeof(int16_t) + 64);
    q.g4 = buf<int16_t>(keep, (size_t)cells * 8 * sizeof(int16_t) + 64);
    q.t2 = buf<int32_t>(keep, (size_t)s2 * CCJ_NT2 * sizeof(int32_t));
    q.stride2 = s2;
    q.w3 = buf<int32_t>(keep, (size_t)s2 * 4 * sizeof(int32_t));
    q.estP = buf<int32_t>(keep, (size_t)s2 * sizeof(int32_t));
    q.inlist = buf<uint32_t>(keep, tri * CCJ_WIN_IN * sizeof(uint32_t));
    q.outlist = buf<uint32_t>(keep, tri * CCJ_WIN_OUT * 2 * sizeof(uint32_t));
    q.incnt = buf<int32_t>(keep, tri * sizeof(int32_t));
    q.outcnt = buf<int32_t>(keep, tri * sizeof(int32_t));
    q.lay = buf<int32_t>(keep, (size_t)CCJ_LAY_INTS(n) * sizeof(int32_t));
    q.scratch = buf<int16_t>(keep, (size_t)ccj_level_max(n) * KF * ccj::Q_COUNT * sizeof(int16_t) + 64);
    q.scratch_stride = ccj_level_max(n);
    q.plw = buf<int16_t>(keep, (size_t)ccj_winlr_quads(n) * 4 * sizeof(int16_t) + 256);
    q.prw = buf<int16_t>(keep, (size_t)ccj_winlr_quads(n) * 4 * sizeof(int16_t) + 256);
    q.pmw = buf<int16_t>(keep, ((size_t)ccj_pmw_level_quads(n) * (size_t)(n > 2 ? n - 2 : 1) + 16) * 8);
    q.pmm = buf<int16_t>(keep, ((size_t)ccj_pmw_level_quads(n) * (size_t)(n > 2 ? n - 2 : 1) + 16) * 8);
    q.pkf = buf<int16_t>(keep, (size_t)ccj_pkf_total(n) * sizeof(int16_t) + 64);
    q.pkg = buf<int16_t>(keep, (size_t)ccj_pkg_total(n) * sizeof(int16_t) + 64);
    q.wscr = buf<int16_t>(keep, ((size_t)wscr_lr * 4 + (size_t)wtot4 * 8) * sizeof(int16_t) + 64);
    q.wscr_lr = wscr_lr;
    q.wtot4 = wtot4;
    q.pmlev4 = buf<int32_t>(keep, (size_t)(n + 1) * (n + 1) * sizeof(int32_t) + 64);
    q.plist = buf<int32_t>(keep, (size_t)(n + 1) * (n + 1) * sizeof(int32_t));
    q.pcum = buf<int32_t>(keep, (size_t)(n + 1) * (n + 2) * sizeof(int32_t));
    q.pmlist = buf<int32_t>(keep, tri * sizeof(int32_t));
    q.pmstart = buf<int32_t>(keep, (size_t)(n + 4) * sizeof(int32_t));
    q.ftype_out = buf<int8_t>(keep, (size_t)n + 2, 0);
    q.tb_stack = buf<int32_t>(keep, sizeof(int32_t) * 5 * (size_t)(16 * n + 64));
    q.tb_cap = 16 * n + 64;
    q.use_lists = 1;
    return q;
}

static int print_hashes(const ccj_cx &c, int n) {
    printf("n %d\n", n);
    for (int t = 0; t < 22; ++t) {
        Fnv f;
        long finite = 0;
        int mn = 1 << 30;
        for (int i = 1; i <= n; ++i)
            for (int j = i; j <= n; ++j)
                for (int k = j + 2; k <= n; ++k)
                    for (int l = k; l <= n; ++l) {
                        const int v = ccj_get4(c, t, i, j, k, l);
                        f.add((uint16_t)(int16_t)v);
                        if (v < 32767) { ++finite; if (v < mn) mn = v; }
                    }
        printf("%s %ld %d %016llx\n", k4dNames[t], finite, finite ? mn : 0, (unsigned long long)f.h);
    }
    for (int t = 0; t < 8; ++t) {
        Fnv f;
        long finite = 0;
        long long sum = 0;
        for (int i = 1; i <= n; ++i)
            for (int j = i; j <= n; ++j) {
                const int32_t v = ccj_raw2(c, t, i, j);
                f.add((uint32_t)v);
                if (v < CCJ_INF / 2) { ++finite; sum += v; }
            }
        printf("%s %ld %lld %016llx\n", k2dNames[t], finite, sum, (unsigned long long)f.h);
    }
    return 0;
}


#ifdef CCJ_ENQUEUE_FILL_INC
// ---- enqueue_fill() of ccj_abi.cu against mock streams and events -------------------------------------------------------
#include <functional>
#include <random>
struct Task {
    int stream;
    std::string name;
    std::vector<int> deps;
    std::function<void()> run;
};
static std::vector<Task> g_tasks;
static int g_last[4] = {-1, -1, -1, -1};          // last task of each stream
static std::vector<int> g_pending[4];             // event waits since the stream's last launch
static std::vector<int> g_event_task;             // event id -> task the record captured (-1: nothing before it)
static bool g_kernels2d = false;
static int g_force_pipe = -1;
static std::vector<ccj_seq> *g_qv = nullptr;
static int sid(cudaStream_t s) { return (int)(uintptr_t)s; }          // mock handles: small integers
static int eid(cudaEvent_t e) { return (int)(uintptr_t)e - 1; }
extern "C" cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) {
    g_event_task.push_back(-2);
    *e = (cudaEvent_t)(uintptr_t)g_event_task.size();
    return cudaSuccess;
}
extern "C" cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
    g_event_task[eid(e)] = g_last[sid(s)];
    return cudaSuccess;
}
extern "C" cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) {
    // CCJ_EMU_DROP_WAIT=k: the k-th wait of the sequence is ignored -- the negative control of this checker
    static const int drop = [] { const char *x = getenv("CCJ_EMU_DROP_WAIT"); return x ? atoi(x) : -1; }();
    static int nwait = 0;
    if (nwait++ == drop) {
        fprintf(stderr, "dropped wait %d: stream %d on the event recorded after task %d (%s)\n", drop, sid(s), g_event_task[eid(e)],
                g_event_task[eid(e)] >= 0 ? g_tasks[g_event_task[eid(e)]].name.c_str() : "-");
        return cudaSuccess;
    }
    if (g_event_task[eid(e)] == -2) {
        fprintf(stderr, "wait on an event that was never recorded\n");
        exit(3);
    }
    if (g_event_task[eid(e)] >= 0) g_pending[sid(s)].push_back(g_event_task[eid(e)]);
    return cudaSuccess;
}
extern "C" cudaError_t cudaGetLastError(void) { return cudaSuccess; }
static void add_task(cudaStream_t st, const std::string &name, std::function<void()> run) {
    Task t;
    t.stream = sid(st);
    t.name = name;
    if (g_last[t.stream] >= 0) t.deps.push_back(g_last[t.stream]);
    for (int d : g_pending[t.stream]) t.deps.push_back(d);
    g_pending[t.stream].clear();
    t.run = std::move(run);
    g_tasks.push_back(std::move(t));
    g_last[g_tasks.back().stream] = (int)g_tasks.size() - 1;
}
struct ccj_ctx {   // the members enqueue_fill touches
    const ccj_model *d_model;
    const ccj_seq *d_seqs;
    cudaStream_t stream, s_win, s_2d;
    std::vector<cudaEvent_t> dep;
};
static bool use_tuned(int) { return true; }
static bool generic_scan() { return false; }
namespace ccj {   // the launchers of ccj_kernels.cu / ccj_fill4.cu (left out under CCJ_HOST_EMU): same grids, as graph nodes
void launch_init(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, cudaStream_t st) {
    add_task(st, "init", [=] { simt::launch(k_init, dim3(2, d.nseq), dim3(256), M, seqs); });
}
void launch_prep(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, cudaStream_t st) {
    add_task(st, "prep", [=] {
        simt::launch(k_prep_lay, dim3(d.nseq), dim3(256), M, seqs);
        simt::launch(k_fill_pmw, dim3(2, d.nseq), dim3(256), seqs);
        simt::launch(k_prep, dim3(d.nmax - 1, d.nseq), dim3(128), M, seqs);
    });
}
void launch_prep_lists(const ccj_model *, const ccj_seq *, LaunchDims, cudaStream_t) {}
void launch_P(const ccj_model *, const ccj_seq *, LaunchDims, int, cudaStream_t) {}
void launch_4d(const ccj_model *, const ccj_seq *, LaunchDims, int, cudaStream_t) {}
void launch_P_tuned(const ccj_model *, const ccj_seq *seqs, LaunchDims d, int s, cudaStream_t st) {
    if (s < 3 || s > d.nmax - 1) return;
    const int per = (d.nmax - s) * d.nseq;
    const int nj = std::max(1, std::min(s - 2, (148 * 6 + per - 1) / per));
    add_task(st, "P" + std::to_string(s), [=] { simt::launch(k_P_tuned, dim3(d.nmax - s, nj, d.nseq), dim3(256), seqs, s, nj); });
}
void launch_2d(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int s, cudaStream_t st) {
    if (d.nmax - s < 1) return;
    add_task(st, "2d" + std::to_string(s), [=] {
        if (g_kernels2d) {
            simt::launch(k_2d, dim3((d.nmax - s + 3) / 4, d.nseq), dim3(128), M, seqs, s);
            return;
        }
        ccj_serial serial;
        for (const ccj_seq &z : *g_qv) {
            ccj_cx cz;
            cz.M = M;
            cz.q = z;
            for (int i = 1; i + s <= z.n; ++i) ccj_cell2d(cz, i, i + s, serial);
        }
    });
}
void launch_4d_roles(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int t, cudaStream_t st) {
    const int m = d.nmax - t - 2;
    if (t % KF != 0 || m < 1) return;
    const int bx = (m * (m + 1) / 2 + K4_THREADS - 1) / K4_THREADS;
    add_task(st, "roles" + std::to_string(t), [=] { simt::launch(k_roles, dim3(bx, t + 1, d.nseq * 4), dim3(K4_THREADS), M, seqs, t); });
}
void launch_4d_windows(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int t, cudaStream_t st) {
    const int nm = d.nmax, m = nm - t - 2, nseq = d.nseq;
    if (m < 1) return;
    add_task(st, "win" + std::to_string(t), [=] {
        const bool pipe = g_force_pipe < 0 ? (long long)nm * nm * nseq < 120000 : g_force_pipe != 0;
        if (!pipe && m >= 32) {
            const int runs = K4_THREADS / 16, nchunk = (m + runs - 1) / runs, npass = (m + 63) / 64;
            simt::launch(k_winLR<false, 16>, dim3(nchunk * npass, t + 1, nseq * 2), dim3(K4_THREADS), M, seqs, t, nchunk);
        } else {
            const int nchunk = (m + WRUNS - 1) / WRUNS, npass = (m + 4 * WGRP - 1) / (4 * WGRP);
            const dim3 grid(nchunk * npass, t + 1, nseq * 2);
            if (pipe) simt::launch(k_winLR<true, 8>, grid, dim3(K4_THREADS), M, seqs, t, nchunk);
            else simt::launch(k_winLR<false, 8>, grid, dim3(K4_THREADS), M, seqs, t, nchunk);
        }
        long long rows = 0;
        for (int s = CCJ_TURN + 1; s <= nm - 1 - t; ++s) rows += nm - s;
        if (rows < 1) return;
        const int cm = std::max(1, std::min(t + 1, nm - 4 - t));
        const int nq = ((cm + 2) >> 2) + 1;
        const int nchunk = (int)((rows + WRUNS - 1) / WRUNS), npass = (nq + WGRP - 1) / WGRP;
        if (pipe) simt::launch(k_winM<true>, dim3(nchunk * npass, nseq), dim3(K4_THREADS), M, seqs, t, nchunk);
        else simt::launch(k_winM<false>, dim3(nchunk * npass, nseq), dim3(K4_THREADS), M, seqs, t, nchunk);
    });
}
void launch_4d_final(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, int t, cudaStream_t st) {
    const int m = d.nmax - t - 2;
    if (m < 1) return;
    const int bx = (m * (m + 1) / 2 + K4_THREADS - 1) / K4_THREADS;
    add_task(st, "final" + std::to_string(t), [=] { simt::launch(k_final, dim3(bx, t + 1, d.nseq), dim3(K4_THREADS), M, seqs, t, t % KF); });
}
void launch_W(const ccj_model *M, const ccj_seq *seqs, LaunchDims d, cudaStream_t st) {
    add_task(st, "W", [=] { simt::launch(k_W, dim3(d.nseq), dim3(32), M, seqs); });
}
int fill4_fused_levels() { return KF; }
}  // namespace ccj
#define CU(x) x
#define CCJ_STR2(x) #x
#define CCJ_STR(x) CCJ_STR2(x)
#include CCJ_STR(CCJ_ENQUEUE_FILL_INC)   // cudaError_t enqueue_fill(ccj_ctx *ctx, ccj::LaunchDims d), verbatim from ccj_abi.cu

// runs the graph in a legal order: prio[stream] decides among the ready tasks (lower first), or a seeded random pick
static int run_graph(const std::string &order) {
    const size_t N = g_tasks.size();
    std::vector<char> done(N, 0);
    std::mt19937 rng(order.rfind("rnd:", 0) == 0 ? (unsigned)atoi(order.c_str() + 4) : 0u);
    int prio[4] = {0, 0, 0, 0};   // streams: 1 main, 2 windows, 3 P + 2D
    if (order == "side") { prio[3] = 0; prio[2] = 1; prio[1] = 2; }
    else if (order == "main") { prio[1] = 0; prio[2] = 1; prio[3] = 2; }
    else if (order.rfind("prio:", 0) == 0 && order.size() == 8)   // prio:ABC -- stream A before B before C whenever ready
        for (int x = 0; x < 3; ++x) prio[(order[5 + x] - '0') & 3] = x;
    for (size_t ran = 0; ran < N; ++ran) {
        std::vector<int> ready;
        for (size_t x = 0; x < N; ++x) {
            if (done[x]) continue;
            bool ok = true;
            for (int d : g_tasks[x].deps) ok = ok && done[d];
            if (ok) ready.push_back((int)x);
        }
        if (ready.empty()) { fprintf(stderr, "dependency cycle\n"); return 3; }
        int pick = ready[0];
        if (order.rfind("rnd:", 0) == 0) pick = ready[rng() % ready.size()];
        else if (order != "issue")   // side / main / prio:ABC
            for (int x : ready)
                if (prio[g_tasks[x].stream] < prio[g_tasks[pick].stream]) pick = x;
        g_tasks[pick].run();
        done[pick] = 1;
    }
    return 0;
}
#endif  // CCJ_ENQUEUE_FILL_INC

int main(int argc, char **argv) {
    if (argc < 5 || (std::string(argv[1]) != "hash" && std::string(argv[1]) != "fold")) {
        fprintf(stderr, "usage: ccj_emu_tuned hash|fold <parfile> <dangles> <seq> [noGU] [pipe] [tuned|lean|generic]\n");
        return 2;
    }
    std::string mode = argv[1], seq = argv[4], err;
    std::string path = argc > 7 ? argv[7] : "tuned";
    // "+2d": k_init, k_2d and k_W run as kernels too (warp collectives per 2D cell: slow here); else the same cell
    // functions are swept serially, as tests/emu/ccj_emu.cpp does
    bool kernels2d = false;
    if (path.size() > 3 && path.substr(path.size() - 3) == "+2d") {
        kernels2d = true;
        path = path.substr(0, path.size() - 3);
    }
    const int dangles = atoi(argv[3]);
    const int noGU = argc > 5 ? atoi(argv[5]) : 0;
    const int force_pipe = argc > 6 ? atoi(argv[6]) : -1;
    ccj::RawParams rp;
    if (!ccj::load_par_file(argv[2], rp, err)) {
        fprintf(stderr, "%s\n", err.c_str());
        return 1;
    }
    static ccj_model M;
    ccj::build_model(rp, dangles, noGU, M);
    // several sequences separated by commas = one wave (hash mode, tuned / lean / generic paths): the kernels index the
    // sequence with blockIdx.z / .y and the grids are sized for the longest one, as ccj_batch_fill launches them
    std::vector<std::string> seqv;
    for (size_t a = 0; a <= seq.size();) {
        const size_t b = seq.find(',', a);
        seqv.push_back(seq.substr(a, b == std::string::npos ? std::string::npos : b - a));
        if (b == std::string::npos) break;
        a = b + 1;
    }
    seq = seqv[0];
    int nm = 0;
    for (const std::string &x : seqv) {
        if (x.size() < 4 || x.size() > K4_MAXN) {
            fprintf(stderr, "length outside the tuned range\n");
            return 2;
        }
        nm = std::max(nm, (int)x.size());
    }
    const int n = (int)seq.size(), nseq = (int)seqv.size();
    if (nseq > 1 && (mode != "hash" || argc <= 7 || std::string(argv[7]).rfind("shard", 0) == 0)) {
        fprintf(stderr, "a wave of several sequences: hash mode, explicit path tuned / lean / generic\n");
        return 2;
    }
    std::vector<std::vector<char>> keep;
    std::vector<ccj_seq> qv;
    for (const std::string &x : seqv) qv.push_back(make_seq(x, keep));
    ccj_seq q = qv[0];
    const int64_t s2 = ccj_stride2(n);
    const ccj_seq *seqs = qv.data();
    const ccj_model *Mp = &M;
    ccj_cx c;
    c.M = Mp;
    c.q = q;

    const std::string order = argc > 8 ? argv[8] : "";
    if (!order.empty()) {
#ifdef CCJ_ENQUEUE_FILL_INC
        if (path != "tuned" || mode != "hash") return 2;
        g_kernels2d = kernels2d;
        g_force_pipe = force_pipe;
        g_qv = &qv;
        ccj_ctx ctx;
        ctx.d_model = Mp;
        ctx.d_seqs = seqs;
        ctx.stream = (cudaStream_t)(uintptr_t)1;
        ctx.s_win = (cudaStream_t)(uintptr_t)2;
        ctx.s_2d = (cudaStream_t)(uintptr_t)3;
        ccj::LaunchDims d;
        d.nseq = nseq;
        d.nmax = nm;
        if (enqueue_fill(&ctx, d) != cudaSuccess) return 3;
        size_t cross = 0;
        for (const Task &t : g_tasks)
            for (int dd : t.deps) cross += g_tasks[dd].stream != t.stream;
        fprintf(stderr, "graph: %zu launches, %zu cross-stream dependencies, order %s\n", g_tasks.size(), cross, order.c_str());
        if (run_graph(order.substr(order.rfind('=') == std::string::npos ? 0 : order.rfind('=') + 1))) return 3;
        if (nseq == 1) return print_hashes(c, n);
        for (int x = 0; x < nseq; ++x) {
            printf("# sequence %d\n", x);
            ccj_cx cz;
            cz.M = Mp;
            cz.q = qv[x];
            print_hashes(cz, qv[x].n);
        }
        return 0;
#else
        fprintf(stderr, "built without CCJ_ENQUEUE_FILL_INC\n");
        return 2;
#endif
    }
    if (path.rfind("shard", 0) == 0) {
        const int G = atoi(path.c_str() + 5);
        if (G < 1 || G > 16) return 2;
        std::vector<int64_t> lev(n + 2, 0);
        for (int t = 0; t <= n; ++t) lev[t + 1] = lev[t] + ccj_shard_level_cells(n, t, G);
        int16_t *rep = buf<int16_t>(keep, (size_t)lev[n + 1] * CCJ_SHARD_NREP * G * sizeof(int16_t) + 16);
        std::vector<int16_t *> locptr;
        for (int r = 0; r < G; ++r) locptr.push_back(buf<int16_t>(keep, (size_t)lev[n + 1] * CCJ_SHARD_NLOC * sizeof(int16_t) + 16));
        std::vector<ccj_seq> qs(G, q);
        for (int r = 0; r < G; ++r) {
            ccj_seq &z = qs[r];
            z.t4 = nullptr;   // nothing may touch the ordinary layout
            z.g1 = z.g2 = z.g3 = z.g4 = nullptr;
            z.lay = nullptr;
            z.scratch = z.plw = z.prw = z.pmw = z.pmm = z.pkf = z.pkg = z.wscr = nullptr;
            z.shard_G = G;
            z.shard_rank = r;
            z.shard_shift = -1;
            for (int b = 0; b < 16; ++b)
                if ((1 << b) == G) z.shard_shift = b;
            z.shard_lev = lev.data();
            z.shard_rep = rep;
            z.shard_loc = locptr.data();
            ccj_shard_kinds(z.shard_kind);
            if (!shard_lean_ok(z, lev.data(), n, G)) {
                fprintf(stderr, "shard_lean_ok refuses this fold\n");
                return 2;
            }
        }
        c.q = qs[0];
        ccj_serial serial;
        for (int64_t x = 0; x < s2; ++x) {
            q.t2[T2_V * s2 + x] = CCJ_V_UNSET;
            q.t2[T2_VTYPE * s2 + x] = 'N';
            for (int t = T2_WM; t < CCJ_NT2; ++t) q.t2[t * s2 + x] = CCJ_INF + 1;
        }
        for (int x = 0; x <= n + 1; ++x) {
            if (x <= n) q.W[x] = 0;
            q.pair_out[x] = -1;
            q.ftype_out[x] = 'N';
        }
        memset(q.status, 0, sizeof(int32_t) * CCJ_STATUS_INTS);
        simt::launch(ccj::k_prep, dim3(nm - 1, nseq), dim3(128), Mp, (const ccj_seq *)&qs[0]);   // launch_prep_lists
        const bool pow2 = qs[0].shard_shift >= 0;
        for (int sp = 0; sp < nm; ++sp) {   // the loop of ccj_shard_fill
            const size_t smem = (size_t)(sp + 1) * sizeof(ccj_lean_lvl);
            if (sp >= 3 && sp <= nm - 1)
                for (int r = 0; r < G; ++r) {
                    const int rows = (int)ccj_shard_rows(nm - sp, r, G);
                    if (rows <= 0) continue;
                    if (pow2) simt::launch_smem(k_P_shard_lean<true>, dim3(rows, sp), dim3(256), smem, Mp, (const ccj_seq *)&qs[r], sp);
                    else simt::launch_smem(k_P_shard_lean<false>, dim3(rows, sp), dim3(256), smem, Mp, (const ccj_seq *)&qs[r], sp);
                }   // every rank's atomicMin lands in the one 2D table: the allreduce(min) of the span-sp diagonal
            if (kernels2d) simt::launch(ccj::k_2d, dim3((nm - sp + 3) / 4, nseq), dim3(128), Mp, (const ccj_seq *)&qs[0], sp);
            else for (int i = 1; i + sp <= n; ++i) ccj_cell2d(c, i, i + sp, serial);
            const int m = nm - sp - 2;
            for (int r = 0; r < G && m >= 1; ++r) {
                const int64_t ncell = ccj_shard_slab(m, r, G);
                if (ncell < 1) continue;
                const dim3 grid((unsigned)((ncell + 127) / 128), sp + 1);
                if (pow2) simt::launch_smem(k_4d_shard_lean<true>, grid, dim3(128), smem, Mp, (const ccj_seq *)&qs[r], sp);
                else simt::launch_smem(k_4d_shard_lean<false>, grid, dim3(128), smem, Mp, (const ccj_seq *)&qs[r], sp);
            }
        }
        for (int j = CCJ_TURN + 1; j <= n; ++j) q.W[j] = ccj_W_at(c, j, serial);
        if (mode == "fold") {
            simt::launch(ccj::k_traceback, dim3(nseq), dim3(TB_THREADS), Mp, (const ccj_seq *)&qs[0]);
            return ccj::emit_result(seq, n, q.W[n], q.pair_out, q.status, stdout, stderr);
        }
        return print_hashes(c, n);
    }
    const bool tuned = path == "tuned";
    const size_t lean_smem = 2 * (size_t)(nm + 1) * sizeof(int64_t);
    // launch_init; launch_prep (tuned) / launch_prep_lists (one thread per cell)
    ccj_serial serial;
    if (kernels2d) {
        simt::launch(ccj::k_init, dim3(2, nseq), dim3(256), Mp, seqs);
    } else {
        for (const ccj_seq &z : qv) {
            const int64_t zs2 = z.stride2;
            for (int64_t x = 0; x < zs2; ++x) {
                z.t2[T2_V * zs2 + x] = CCJ_V_UNSET;
                z.t2[T2_VTYPE * zs2 + x] = 'N';
                for (int t = T2_WM; t < CCJ_NT2; ++t) z.t2[t * zs2 + x] = CCJ_INF + 1;
            }
            for (int x = 0; x <= z.n + 1; ++x) {
                if (x <= z.n) z.W[x] = 0;
                z.pair_out[x] = -1;
                z.ftype_out[x] = 'N';
            }
            memset(z.status, 0, sizeof(int32_t) * CCJ_STATUS_INTS);
        }
    }
    if (tuned) {
        simt::launch(ccj::k_prep_lay, dim3(nseq), dim3(256), Mp, seqs);
        simt::launch(ccj::k_fill_pmw, dim3(2, nseq), dim3(256), seqs);   // grid-stride loop: any grid covers the buffer
    }
    simt::launch(ccj::k_prep, dim3(nm - 1, nseq), dim3(128), Mp, seqs);

    auto span_step = [&](int sp) {
        if (sp >= nm) return;
        if (sp >= 3 && sp <= nm - 1) {
            if (tuned) {   // launch_P_tuned
                const int per = (nm - sp) * nseq;
                const int nj = std::max(1, std::min(sp - 2, (148 * 6 + per - 1) / per));
                simt::launch(ccj::k_P_tuned, dim3(nm - sp, nj, nseq), dim3(256), seqs, sp, nj);
            } else if (path == "lean") {   // launch_P
                simt::launch_smem(ccj::k_P_lean, dim3(nm - sp, sp, nseq), dim3(256), lean_smem, Mp, seqs, sp);
            } else {
                simt::launch(ccj::k_P, dim3(nm - sp, sp, nseq), dim3(256), Mp, seqs, sp);
            }
        }
        if (kernels2d) simt::launch(ccj::k_2d, dim3((nm - sp + 3) / 4, nseq), dim3(128), Mp, seqs, sp);   // launch_2d
        else for (const ccj_seq &z : qv) {
            ccj_cx cz;
            cz.M = Mp;
            cz.q = z;
            for (int i = 1; i + sp <= z.n; ++i) ccj_cell2d(cz, i, i + sp, serial);
        }
    };
    const int lead = tuned ? (KF - 2 > 0 ? KF - 2 : 0) : 0;
    for (int sp = 0; sp < lead; ++sp) span_step(sp);
    for (int t = 0; t < nm; ++t) {
        span_step(t + lead);
        const int m = nm - t - 2;
        if (m < 1) continue;
        if (!tuned) {   // launch_4d
            const dim3 grid((m * (m + 1) / 2 + 127) / 128, t + 1, nseq);
            if (path == "lean") simt::launch_smem(ccj::k_4d_lean, grid, dim3(128), lean_smem, Mp, seqs, t, nm);
            else simt::launch(ccj::k_4d, grid, dim3(128), Mp, seqs, t);
            continue;
        }
        const int bx = (m * (m + 1) / 2 + K4_THREADS - 1) / K4_THREADS;
        if (t % KF == 0) simt::launch(ccj::k_roles, dim3(bx, t + 1, nseq * 4), dim3(K4_THREADS), Mp, seqs, t);
        {   // launch_4d_windows
            const bool pipe = force_pipe < 0 ? (long long)nm * nm * nseq < 120000 : force_pipe != 0;
            if (!pipe && m >= 32) {
                const int runs = K4_THREADS / 16, nchunk = (m + runs - 1) / runs, npass = (m + 63) / 64;
                simt::launch(ccj::k_winLR<false, 16>, dim3(nchunk * npass, t + 1, nseq * 2), dim3(K4_THREADS), Mp, seqs, t, nchunk);
            } else {
                const int nchunk = (m + WRUNS - 1) / WRUNS, npass = (m + 4 * WGRP - 1) / (4 * WGRP);
                const dim3 grid(nchunk * npass, t + 1, nseq * 2);
                if (pipe) simt::launch(ccj::k_winLR<true, 8>, grid, dim3(K4_THREADS), Mp, seqs, t, nchunk);
                else simt::launch(ccj::k_winLR<false, 8>, grid, dim3(K4_THREADS), Mp, seqs, t, nchunk);
            }
            long long rows = 0;
            for (int s = CCJ_TURN + 1; s <= nm - 1 - t; ++s) rows += nm - s;
            if (rows >= 1) {
                const int cm = std::max(1, std::min(t + 1, nm - 4 - t));
                const int nq = ((cm + 2) >> 2) + 1;
                const int nchunk = (int)((rows + WRUNS - 1) / WRUNS), npass = (nq + WGRP - 1) / WGRP;
                const dim3 grid(nchunk * npass, nseq);
                if (pipe) simt::launch(ccj::k_winM<true>, grid, dim3(K4_THREADS), Mp, seqs, t, nchunk);
                else simt::launch(ccj::k_winM<false>, grid, dim3(K4_THREADS), Mp, seqs, t, nchunk);
            }
        }
        simt::launch(ccj::k_final, dim3(bx, t + 1, nseq), dim3(K4_THREADS), Mp, seqs, t, t % KF);
    }
    if (kernels2d) simt::launch(ccj::k_W, dim3(nseq), dim3(32), Mp, seqs);   // launch_W
    else for (const ccj_seq &z : qv) {
        ccj_cx cz;
        cz.M = Mp;
        cz.q = z;
        for (int j = CCJ_TURN + 1; j <= z.n; ++j) z.W[j] = ccj_W_at(cz, j, serial);
    }
    for (const ccj_seq &z : qv)
        if (z.status[5] || z.status[7]) fprintf(stderr, "status: list overflow %d, int16 guard %d\n", z.status[5], z.status[7]);

    if (mode == "fold") {
        simt::launch(ccj::k_traceback, dim3(nseq), dim3(TB_THREADS), Mp, seqs);   // launch_traceback
        return ccj::emit_result(seq, n, q.W[n], q.pair_out, q.status, stdout, stderr);
    }
    if (nseq == 1) return print_hashes(c, n);
    for (int x = 0; x < nseq; ++x) {
        printf("# sequence %d\n", x);
        ccj_cx cz;
        cz.M = Mp;
        cz.q = qv[x];
        print_hashes(cz, qv[x].n);
    }
    return 0;
}
