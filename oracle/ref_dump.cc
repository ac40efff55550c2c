// TEST INFRASTRUCTURE ONLY -- never linked into, called from, or shipped with the product.
//
// ccj_ref_dump: drives the UNMODIFIED reference classes (compiled from /root/reference by
// oracle/Makefile) and dumps what the reference computes, so that tests can compare the CUDA
// path table-by-table and parameter-by-parameter:
//
//   ccj_ref_dump params <parfile|-> <dangles>            scaled vrna_param_t as "name idx... value" text
//   ccj_ref_dump hash   <parfile|-> <dangles> <seq>      per-table summary + FNV-1a hash (text)
//   ccj_ref_dump bin    <parfile|-> <dangles> <seq> <out> raw tables (layout below)
//   ccj_ref_dump fold   <parfile|-> <dangles> <seq>      same stdout as the CCJ binary's last two lines
//   ccj_ref_dump probe  <parfile|-> <dangles> <seq>      public class surface probe (tests/shell/probe_body.inc)
//
// The reference keeps its tables private; this TU only flips the access specifiers while including
// the reference headers (no reference source is modified or copied).
//
// Canonical export order (shared with ccj_b200's ccj_export_table, see include/ccj_b200.h):
//   4D table : for i=1..n, j=i..n, k=j+2..n, l=k..n  -> int16   (C(n+1,4) values)
//   2D table : for i=1..n, j=i..n                    -> int32   (n(n+1)/2 values)
//   W        : W[0..n]                               -> int32
#define private public
#define protected public
#include "W_final.hh"
#undef private
#undef protected
#include "h_globals.hh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <sys/stat.h>

extern "C" {
#include "ViennaRNA/params/io.h"
}

static const char *k4dNames[22] = {
    "PK", "PL", "PR", "PM", "PO", "PfromL", "PfromR", "PfromM", "PfromMprime", "PfromO",
    "PLmloop00", "PLmloop01", "PLmloop10", "PRmloop00", "PRmloop01", "PRmloop10",
    "PMmloop00", "PMmloop01", "PMmloop10", "POmloop00", "POmloop01", "POmloop10"};

static Matrix4D *table4d(pseudo_loop *P, int t) {
    Matrix4D *tabs[22] = {&P->PK, &P->PL, &P->PR, &P->PM, &P->PO, &P->PfromL, &P->PfromR, &P->PfromM,
                          &P->PfromMprime, &P->PfromO, &P->PLmloop00, &P->PLmloop01, &P->PLmloop10,
                          &P->PRmloop00, &P->PRmloop01, &P->PRmloop10, &P->PMmloop00, &P->PMmloop01,
                          &P->PMmloop10, &P->POmloop00, &P->POmloop01, &P->POmloop10};
    return tabs[t];
}

static void load_params(const char *parfile) {
    if (strcmp(parfile, "-") == 0) return;  // compiled-in defaults
    struct stat sb;
    if (stat(parfile, &sb) != 0) {
        fprintf(stderr, "Not a valid parameter file!\n");
        exit(1);
    }
    vrna_params_load(parfile, VRNA_PARAMETER_FORMAT_DEFAULT);
}

// the reference's fill order: rows from the 3' end, columns ascending (W_final::ccj drives exactly this)
static void run_fill(W_final &w) {
    int n = w.n;
    for (int i = n; i >= 1; --i)
        for (int j = i; j <= n; ++j) {
            w.V->compute_energy(i, j);
            w.P->compute_energies(i, j);
            w.V->compute_WMv_WMp(i, j, w.P->get_energy(i, j));
            w.V->compute_energy_WM(i, j, w.P->P);
        }
}

struct Fnv {
    uint64_t h = 1469598103934665603ULL;
    void add(uint64_t v) { h ^= v; h *= 1099511628211ULL; }
};

template <class F> static void for4d(int n, F f) {
    for (int i = 1; i <= n; ++i)
        for (int j = i; j <= n; ++j)
            for (int k = j + 2; k <= n; ++k)
                for (int l = k; l <= n; ++l) f(i, j, k, l);
}

static int32_t get2d(W_final &w, int t, int i, int j) {
    s_energy_matrix *V = w.V;
    int ij = V->index[i] + j - i;
    switch (t) {
        case 0: return V->nodes[ij].energy;
        case 1: return (int32_t)V->nodes[ij].type;
        case 2: return V->WM[ij];
        case 3: return V->WMv[ij];
        case 4: return V->WMp[ij];
        case 5: return w.P->P[ij];
        case 6: return w.P->WBP[ij];
        default: return w.P->WPP[ij];
    }
}
static const char *k2dNames[8] = {"V", "Vtype", "WM", "WMv", "WMp", "P", "WBP", "WPP"};

#include "../tests/shell/probe_body.inc"

#define DUMP1(name, len) for (int a = 0; a < (len); ++a) printf(#name " %d %d\n", a, p->name[a]);
#define DUMP2(name, l0, l1) for (int a = 0; a < (l0); ++a) for (int b = 0; b < (l1); ++b) printf(#name " %d %d %d\n", a, b, p->name[a][b]);
#define DUMP3(name, l0, l1, l2) for (int a = 0; a < (l0); ++a) for (int b = 0; b < (l1); ++b) for (int c = 0; c < (l2); ++c) printf(#name " %d %d %d %d\n", a, b, c, p->name[a][b][c]);

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: ccj_ref_dump params|hash|bin|fold <parfile|-> <dangles> [seq] [out]\n");
        return 2;
    }
    std::string mode = argv[1];
    load_params(argv[2]);
    int dangles = atoi(argv[3]);

    if (mode == "params") {
        vrna_param_t *p = scale_parameters();
        p->model_details.dangles = dangles;
        DUMP2(stack, 8, 8)
        DUMP1(hairpin, 31)
        DUMP1(bulge, 31)
        DUMP1(internal_loop, 31)
        DUMP3(mismatchExt, 8, 5, 5)
        DUMP3(mismatchI, 8, 5, 5)
        DUMP3(mismatch1nI, 8, 5, 5)
        DUMP3(mismatch23I, 8, 5, 5)
        DUMP3(mismatchH, 8, 5, 5)
        DUMP3(mismatchM, 8, 5, 5)
        DUMP2(dangle5, 8, 5)
        DUMP2(dangle3, 8, 5)
        for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b) for (int c = 0; c < 5; ++c) for (int d = 0; d < 5; ++d)
            printf("int11 %d %d %d %d %d\n", a, b, c, d, p->int11[a][b][c][d]);
        for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b) for (int c = 0; c < 5; ++c) for (int d = 0; d < 5; ++d) for (int e = 0; e < 5; ++e)
            printf("int21 %d %d %d %d %d %d\n", a, b, c, d, e, p->int21[a][b][c][d][e]);
        for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b) for (int c = 0; c < 5; ++c) for (int d = 0; d < 5; ++d) for (int e = 0; e < 5; ++e) for (int f = 0; f < 5; ++f)
            printf("int22 %d %d %d %d %d %d %d\n", a, b, c, d, e, f, p->int22[a][b][c][d][e][f]);
        DUMP1(ninio, 5)
        printf("lxc %.17g\n", p->lxc);
        printf("MLbase %d\n", p->MLbase);
        DUMP1(MLintern, 8)
        printf("MLclosing %d\n", p->MLclosing);
        printf("TerminalAU %d\n", p->TerminalAU);
        printf("DuplexInit %d\n", p->DuplexInit);
        printf("Tetraloops %s|\n", p->Tetraloops);
        for (size_t a = 0; a * 7 < strlen(p->Tetraloops); ++a) printf("Tetraloop_E %zu %d\n", a, p->Tetraloop_E[a]);
        printf("Triloops %s|\n", p->Triloops);
        for (size_t a = 0; a * 6 < strlen(p->Triloops); ++a) printf("Triloop_E %zu %d\n", a, p->Triloop_E[a]);
        printf("Hexaloops %s|\n", p->Hexaloops);
        for (size_t a = 0; a * 9 < strlen(p->Hexaloops); ++a) printf("Hexaloop_E %zu %d\n", a, p->Hexaloop_E[a]);
        printf("special_hp %d\n", p->model_details.special_hp);
        printf("dangles %d\n", p->model_details.dangles);
        return 0;
    }

    if (argc < 5) return 2;
    std::string seq = argv[4];
    if (mode == "probe") {
        run_probe(seq, dangles);
        return 0;
    }
    W_final w(seq, dangles);
    int n = w.n;

    if (mode == "fold") {
        double e = w.ccj();
        printf("%s\n%s (%g)\n", seq.c_str(), w.structure.c_str(), e);
        return 0;
    }

    run_fill(w);
    // exterior W exactly as the reference computes it needs ccj(); recompute through ccj() only in
    // "hash"/"bin" after the tables were captured would re-run the fill, so W is reported by the fold mode
    // of the real binary (energy = W[n]/100) instead.

    if (mode == "hash") {
        printf("n %d\n", n);
        for (int t = 0; t < 22; ++t) {
            Matrix4D *M = table4d(w.P, t);
            Fnv f;
            long finite = 0;
            int mn = 1 << 30;
            for4d(n, [&](int i, int j, int k, int l) {
                int v = M->get(i, j, k, l);
                f.add((uint16_t)(int16_t)v);
                if (v < 32767) { ++finite; if (v < mn) mn = v; }
            });
            printf("%s %ld %d %016llx\n", k4dNames[t], finite, finite ? mn : 0, (unsigned long long)f.h);
        }
        for (int t = 0; t < 8; ++t) {
            Fnv f;
            long finite = 0;
            long long sum = 0;
            for (int i = 1; i <= n; ++i)
                for (int j = i; j <= n; ++j) {
                    int32_t v = get2d(w, t, i, j);
                    f.add((uint32_t)v);
                    if (v < INF / 2) { ++finite; sum += v; }
                }
            printf("%s %ld %lld %016llx\n", k2dNames[t], finite, sum, (unsigned long long)f.h);
        }
        return 0;
    }

    if (mode == "bin") {
        if (argc < 6) return 2;
        FILE *fp = fopen(argv[5], "wb");
        if (!fp) { perror("fopen"); return 1; }
        int32_t hdr[4] = {0x434a4344 /* 'DCJC' */, n, 22, 8};
        fwrite(hdr, 4, 4, fp);
        std::vector<int16_t> buf;
        for (int t = 0; t < 22; ++t) {
            Matrix4D *M = table4d(w.P, t);
            buf.clear();
            for4d(n, [&](int i, int j, int k, int l) { buf.push_back((int16_t)M->get(i, j, k, l)); });
            fwrite(buf.data(), 2, buf.size(), fp);
        }
        std::vector<int32_t> b2;
        for (int t = 0; t < 8; ++t) {
            b2.clear();
            for (int i = 1; i <= n; ++i)
                for (int j = i; j <= n; ++j) b2.push_back(get2d(w, t, i, j));
            fwrite(b2.data(), 4, b2.size(), fp);
        }
        fclose(fp);
        return 0;
    }
    return 2;
}
