// TEST TOOL (not part of the product): compiles the product's host/device cell functions
// (ccj_b200/csrc/ccj_cells*.cuh) for the CPU and sweeps them single-threaded in the SAME wavefront
// order the CUDA kernels use (2D span s, then 4D level t=s).  It exists so that the recurrences, the
// table layout and the level schedule can be debugged against oracle/_ref without a GPU.
// The shipped library never contains this loop: without CUDA the product fails loudly.
//
//   ccj_emu hash <parfile> <dangles> <seq>   -> same text as `ccj_ref_dump hash`
//   ccj_emu fold <parfile> <dangles> <seq>   -> same stdout/stderr/exit code as the CCJ binary
//   ... [noGU] [G]: with G > 0 the gap tables live in the ROW-SHARDED layout of ccj_types.h (G ranks; the ranks'
//   replicated region is shared here, which is what the per-level allgather establishes on the GPUs)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../ccj_b200/csrc/energy_model.hpp"
#include "../../ccj_b200/csrc/ccj_cells4_lean.cuh"
#ifndef CCJ_EMU_NO_TB
#include "../../ccj_b200/csrc/ccj_traceback.cuh"
#include "../../ccj_b200/csrc/ccj_render.hpp"
#endif

struct Fnv {
    uint64_t h = 1469598103934665603ULL;
    void add(uint64_t v) { h ^= v; h *= 1099511628211ULL; }
};

static const char *k4dNames[22] = {
    "PK", "PL", "PR", "PM", "PO", "PfromL", "PfromR", "PfromM", "PfromMprime", "PfromO",
    "PLmloop00", "PLmloop01", "PLmloop10", "PRmloop00", "PRmloop01", "PRmloop10",
    "PMmloop00", "PMmloop01", "PMmloop10", "POmloop00", "POmloop01", "POmloop10"};
static const char *k2dNames[8] = {"V", "Vtype", "WM", "WMv", "WMp", "P", "WBP", "WPP"};

int main(int argc, char **argv) {
    if (argc < 5) {
        fprintf(stderr, "usage: ccj_emu hash|fold <parfile> <dangles> <seq> [noGU]\n");
        return 2;
    }
    std::string mode = argv[1], seq = argv[4], err;
    int dangles = atoi(argv[3]);
    int noGU = argc > 5 ? atoi(argv[5]) : 0;
    const int shardG = argc > 6 ? atoi(argv[6]) : 0;
    const int packed2d = argc > 7 ? atoi(argv[7]) : 0;   // keep {WB,WP,WBP} packed per interval (ccj_seq::w3) as the GPU folds do
    const int lists = argc > 8 ? atoi(argv[8]) : 0;      // walk per-pair partner lists in the interior windows (k_prep's rule)
    const int lean = argc > 9 ? atoi(argv[9]) : 0;       // sweep ccj_cell4d_lean (the form k_4d / k_4d_shard run) instead of ccj_cell4d
    if (lean && !(packed2d && lists)) {
        fprintf(stderr, "lean needs the packed 2D records and the partner lists\n");
        return 2;
    }
    ccj::RawParams rp;
    if (!ccj::load_par_file(argv[2], rp, err)) {
        fprintf(stderr, "%s\n", err.c_str());
        return 1;
    }
    static ccj_model M;
    ccj::build_model(rp, dangles, noGU, M);

    const int n = (int)seq.size();
    std::vector<int8_t> S(n + 2);
    for (int i = 1; i <= n; ++i) S[i] = (int8_t)ccj::encode_base(seq[i - 1]);
    S[n + 1] = S[1];
    S[0] = S[n];
    const int64_t cells = ccj_cells4(n), s2 = ccj_stride2(n);
    std::vector<int16_t> t4((size_t)(cells * CCJ_NT4 + 8), (int16_t)0x5555);  // poison: every valid cell must be written
    std::vector<int32_t> t2((size_t)(s2 * CCJ_NT2));
    std::vector<int32_t> W(n + 1, 0), pairv(n + 2, -1), st(CCJ_STATUS_INTS, 0), tbs(5 * (16 * n + 64));
    std::vector<int8_t> ftype(n + 2, 'N');
    for (int64_t x = 0; x < s2; ++x) {
        t2[T2_V * s2 + x] = CCJ_V_UNSET;
        t2[T2_VTYPE * s2 + x] = 'N';
        for (int t = T2_WM; t < CCJ_NT2; ++t) t2[t * s2 + x] = CCJ_INF + 1;
    }
    ccj_cx c;
    memset(&c.q, 0, sizeof c.q);
    c.M = &M;
    c.q.n = n;
    c.q.S = S.data();
    c.q.seq = seq.c_str();
    c.q.t4 = t4.data();
    c.q.stride4 = cells;
    c.q.t2 = t2.data();
    c.q.stride2 = s2;
    c.q.W = W.data();
    c.q.pair_out = pairv.data();
    c.q.ftype_out = ftype.data();
    c.q.status = st.data();
    c.q.tb_stack = tbs.data();
    c.q.tb_cap = 16 * n + 64;
    std::vector<int32_t> w3v;
    if (packed2d) {
        w3v.assign((size_t)s2 * 4, 0x55555555);
        c.q.w3 = w3v.data();
    }
    // partner lists, built on the host by the rule of k_prep (ccj_fill4.cu): entry = int16 energy | x << 16 | y << 24
    std::vector<uint32_t> inl, outl;
    std::vector<int32_t> incnt, outcnt;
    if (lists) {
        const size_t tri = (size_t)n * (n - 1) / 2 + 1;
        inl.assign(tri * CCJ_WIN_IN, 0);
        outl.assign(tri * CCJ_WIN_OUT * 2, 0);
        incnt.assign(tri, 0);
        outcnt.assign(tri, 0);
        auto pack = [&](int e, int x, int y, uint32_t &out) {
            if (e >= CCJ_INF / 2 + 40000) return false;
            if (e > 32767 || e < -32768) { st[5] = 1; return false; }
            out = (uint32_t)(uint16_t)(int16_t)e | ((uint32_t)x << 16) | ((uint32_t)y << 24);
            return true;
        };
        for (int j = 2; j <= n; ++j)
            for (int i = 1; i < j; ++i) {
                if (!ccj_can_pair(c, i, j)) continue;
                const size_t slot = (size_t)ccj_tri(i, j);
                int nin = 0, nout = 0, nneg = 0;
                for (int sidx = 0; sidx < CCJ_WIN; ++sidx) {
                    const int x = sidx / 29 + 1, y = sidx % 29 + 1, d = i + x, dp = j - y;
                    uint32_t ent;
                    if (x <= j - i - 1 && dp >= d + 4 && ccj_can_pair(c, d, dp) && pack(ccj_e_intP(&M, S.data(), i, d, dp, j), x, y, ent))
                        inl[slot * CCJ_WIN_IN + nin++] = ent;
                }
                for (int sweep = 0; sweep < 2; ++sweep) {
                    for (int sidx = 0; sidx < CCJ_WIN; ++sidx) {
                        const int x = sidx / 29 + 1, y = sidx % 29 + 1, d = i - x, dp = j + y;
                        if (!(d >= 1 && dp <= n && ccj_can_pair(c, d, dp))) continue;
                        const int e = ccj_e_intP(&M, S.data(), d, i, j, dp);
                        uint32_t ent;
                        if ((e < 0) == (sweep == 0) && pack(e, x, y, ent)) outl[(slot * CCJ_WIN_OUT + nout++) * 2] = ent;
                    }
                    if (sweep == 0) nneg = nout;
                }
                incnt[slot] = nin;
                outcnt[slot] = nout | (nneg << 16);
            }
        c.q.inlist = inl.data();
        c.q.outlist = outl.data();
        c.q.incnt = incnt.data();
        c.q.outcnt = outcnt.data();
        c.q.use_lists = 1;
    }
    std::vector<int64_t> lev(n + 2, 0);
    std::vector<int16_t> rep;
    std::vector<std::vector<int16_t>> loc;
    std::vector<int16_t *> locptr;
    if (shardG > 0) {
        for (int t = 0; t <= n; ++t) lev[t + 1] = lev[t] + ccj_shard_level_cells(n, t, shardG);
        rep.assign((size_t)(lev[n + 1] * CCJ_SHARD_NREP * shardG) + 8, (int16_t)0x5555);
        loc.assign(shardG, std::vector<int16_t>((size_t)(lev[n + 1] * CCJ_SHARD_NLOC) + 8, (int16_t)0x5555));
        for (auto &v : loc) locptr.push_back(v.data());
        c.q.t4 = nullptr;   // nothing may touch the ordinary layout
        c.q.shard_G = shardG;
        c.q.shard_shift = -1;
        for (int b = 0; b < 16; ++b)
            if ((1 << b) == shardG) c.q.shard_shift = b;
        c.q.shard_lev = lev.data();
        c.q.shard_rep = rep.data();
        c.q.shard_loc = locptr.data();
        ccj_shard_kinds(c.q.shard_kind);
    }

    std::vector<int64_t> leantab(2 * (size_t)(n + 1) + 2, 0);
    ccj_lean_plain_tab(leantab.data(), n, 0, 1);
    std::vector<ccj_lean_lvl> leanlvl((size_t)n + 1);
    if (shardG > 0) ccj_lean_lvl_fill(leanlvl.data(), lev.data(), n, shardG, 0, 1);
    if (lean && shardG > 0 && !ccj_lean_kinds_ok(c.q.shard_kind)) {
        fprintf(stderr, "ccj_lean_kind disagrees with ccj_shard_kinds\n");
        return 2;
    }
    ccj_serial par;
    for (int s = 0; s < n; ++s) {
        // K_P + K_2D of span s
        for (int i = 1; i + s <= n; ++i) {
            const int l = i + s;
            int mn = CCJ_INF;
            for (int j = i; j < l; ++j) {
                if (lean && shardG > 0 && c.q.shard_shift >= 0) {
                    ccj_lean_shard<true> ly;
                    ly.rep = rep.data(); ly.loc = nullptr; ly.lvl = leanlvl.data(); ly.n = n; ly.G = shardG; ly.sh = c.q.shard_shift;
                    for (int wid = 0; wid < 3; ++wid)       // the shares of 3 "warps" x 5 "lanes" together are all terms
                        for (int lane = 0; lane < 5; ++lane) mn = ccj_min(mn, ccj_P_lean(ly, i, j, l, wid, 3, lane, 5));
                } else if (lean && shardG > 0) {
                    ccj_lean_shard<false> ly;
                    ly.rep = rep.data(); ly.loc = nullptr; ly.lvl = leanlvl.data(); ly.n = n; ly.G = shardG; ly.sh = -1;
                    for (int wid = 0; wid < 2; ++wid)
                        for (int lane = 0; lane < 3; ++lane) mn = ccj_min(mn, ccj_P_lean(ly, i, j, l, wid, 2, lane, 3));
                } else if (lean) {
                    ccj_lean_plain ly;
                    ly.t4 = c.q.t4; ly.st4 = c.q.stride4; ly.tab = leantab.data(); ly.n = n;
                    for (int wid = 0; wid < 2; ++wid)
                        for (int lane = 0; lane < 4; ++lane) mn = ccj_min(mn, ccj_P_lean(ly, i, j, l, wid, 2, lane, 4));
                } else {
                    for (int d = j + 1; d < l; ++d)
                        for (int k = d + 1; k < l; ++k) mn = ccj_min(mn, ccj_P_term(c, i, l, j, d, k));
                }
            }
            if (mn < CCJ_INF / 2) t2[T2_P * s2 + ccj_idx2(n, i, l)] = mn;
            ccj_cell2d(c, i, l, par);
        }
        // K_4D of level t = s
        const int t = s;
        if (t <= n - 3)
            for (int a = 0; a <= t; ++a) {
                const int b = t - a;
                for (int i = 1; i <= n - t - 2; ++i)
                    for (int k = i + a + 2; k <= n - b; ++k) {
                        if (!lean) {
                            ccj_cell4d(c, i, i + a, k, k + b);
                        } else if (shardG > 0) {
                            if (c.q.shard_shift >= 0) {
                                ccj_lean_shard<true> ly;
                                ly.rep = rep.data();
                                ly.loc = locptr[(i - 1) % shardG];   // the rank that owns row i
                                ly.lvl = leanlvl.data();
                                ly.n = n; ly.G = shardG; ly.sh = c.q.shard_shift;
                                ccj_cell4d_lean(c, ly, i, i + a, k, k + b);
                            } else {
                                ccj_lean_shard<false> ly;
                                ly.rep = rep.data();
                                ly.loc = locptr[(i - 1) % shardG];
                                ly.lvl = leanlvl.data();
                                ly.n = n; ly.G = shardG; ly.sh = -1;
                                ccj_cell4d_lean(c, ly, i, i + a, k, k + b);
                            }
                        } else {
                            ccj_lean_plain ly;
                            ly.t4 = c.q.t4; ly.st4 = c.q.stride4; ly.tab = leantab.data(); ly.n = n;
                            ccj_cell4d_lean(c, ly, i, i + a, k, k + b);
                        }
                    }
            }
    }
    for (int j = CCJ_TURN + 1; j <= n; ++j) W[j] = ccj_W_at(c, j, par);

    if (mode == "hash") {
        printf("n %d\n", n);
        for (int t = 0; t < 22; ++t) {
            Fnv f;
            long finite = 0;
            int mn = 1 << 30;
            for (int i = 1; i <= n; ++i)
                for (int j = i; j <= n; ++j)
                    for (int k = j + 2; k <= n; ++k)
                        for (int l = k; l <= n; ++l) {
                            int v = ccj_get4(c, t, i, j, k, l);
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
                    int32_t v = ccj_raw2(c, t, i, j);
                    f.add((uint32_t)v);
                    if (v < CCJ_INF / 2) { ++finite; sum += v; }
                }
            printf("%s %ld %lld %016llx\n", k2dNames[t], finite, sum, (unsigned long long)f.h);
        }
        return 0;
    }
#ifndef CCJ_EMU_NO_TB
    if (mode == "fold") {
        ccj_traceback(c, par);
        return ccj::emit_result(seq, n, W[n], pairv.data(), st.data(), stdout, stderr);
    }
#endif
    return 2;
}
