// See ccj_shell.hpp / ccj_compat.hh.  Host glue only: the fill and the traceback run on the GPU.
#include "ccj_shell.hpp"
#include "h_externs.hh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>

int noGU = 0;

namespace {
std::string g_param_source;   // "" = compiled-in defaults, "@name" = embedded set, else a file

// the fold whose tables currently sit in the shared context (one wave at a time)
ccj::ShellFold *g_resident = nullptr;
std::vector<std::weak_ptr<ccj::ShellFold>> g_folds;

void die_abi(ccj_ctx *ctx, const char *what) {
    std::cerr << "ccj_b200: " << what << ": " << ccj_last_error(ctx) << std::endl;
    exit(EXIT_FAILURE);
}
}  // namespace

// ---- ViennaRNA entry points the reference's driver code calls (ccj_compat.hh) -----------------------------------------
int vrna_params_load(const char *fname, unsigned int) {
    FILE *f = fname ? fopen(fname, "r") : nullptr;
    if (!f) return 0;
    fclose(f);
    g_param_source = fname;
    return 1;
}

int vrna_params_load_DNA_Mathews2004(void) {
    g_param_source = "@dna_mathews2004";
    return 1;
}

vrna_param_t *scale_parameters(void) {
    ccj_model *m = new ccj_model();
    char err[256] = {0};
    const std::string src = g_param_source.empty() ? "@rna_turner2004" : g_param_source;
    // scaling happens with the default model details: dangles = 2 (src/W_final.cc:20; the caller sets
    // model_details.dangles afterwards, :25)
    if (ccj_model_build(src.c_str(), 2, noGU, m, sizeof(ccj_model), err, sizeof err) != 0) {
        delete m;
        return nullptr;
    }
    vrna_param_t *p = static_cast<vrna_param_t *>(calloc(1, sizeof(vrna_param_t)));
    ccj::vrna_from_model(*m, *p);
    if (!g_param_source.empty() && g_param_source[0] != '@') strncpy(p->param_file, g_param_source.c_str(), 255);
    delete m;
    return p;
}

short *encode_sequence(const char *sequence, short how) {
    const size_t l = strlen(sequence);
    short *S = static_cast<short *>(calloc(l + 2, sizeof(short)));
    for (size_t i = 1; i <= l; ++i) {
        const char c = sequence[i - 1];
        S[i] = (short)(c == 'A' || c == 'a' ? 1 : c == 'C' || c == 'c' ? 2 : c == 'G' || c == 'g' ? 3
                       : c == 'U' || c == 'u' || c == 'T' || c == 't' ? 4 : 0);
    }
    S[l + 1] = S[1];
    S[0] = how == 0 ? (short)l : S[l];
    return S;
}

energy_t TriangleMatrix::get_uc(cand_pos_t i, cand_pos_t j) const { return fold_->raw2(table_, i, j); }

namespace ccj {

ccj_ctx *shell_ctx() {
    static ccj_ctx *ctx = nullptr;
    if (!ctx) {
        const char *d = getenv("CCJ_DEVICE");
        if (ccj_ctx_create(d ? atoi(d) : 0, &ctx) != 0) {
            std::cerr << "ccj_b200: no CUDA device available (this build has no CPU path)" << std::endl;
            exit(EXIT_FAILURE);
        }
    }
    return ctx;
}

// the fields the path reads; the special-loop strings keep the reference's pitch (7/6/9)
void model_from_vrna(const vrna_param_t &p, int no_gu, ccj_model &m) {
    memset(&m, 0, sizeof m);
    memcpy(m.stack, p.stack, sizeof m.stack);
    memcpy(m.mismatchExt, p.mismatchExt, sizeof m.mismatchExt);
    memcpy(m.mismatchI, p.mismatchI, sizeof m.mismatchI);
    memcpy(m.mismatch1nI, p.mismatch1nI, sizeof m.mismatch1nI);
    memcpy(m.mismatch23I, p.mismatch23I, sizeof m.mismatch23I);
    memcpy(m.mismatchH, p.mismatchH, sizeof m.mismatchH);
    memcpy(m.mismatchM, p.mismatchM, sizeof m.mismatchM);
    memcpy(m.dangle5, p.dangle5, sizeof m.dangle5);
    memcpy(m.dangle3, p.dangle3, sizeof m.dangle3);
    memcpy(m.int11, p.int11, sizeof m.int11);
    memcpy(m.int21, p.int21, sizeof m.int21);
    memcpy(m.int22, p.int22, sizeof m.int22);
    for (int s = 0; s < CCJ_HAIRPIN_TAB; ++s)   // hairpin.h:158-161
        m.hairpin[s] = s <= 30 ? p.hairpin[s] : p.hairpin[30] + (int)(p.lxc * log(s / 30.));
    for (int s = 0; s < CCJ_LOOP_TAB; ++s) {     // internal.h:505-507,556-560
        m.bulge[s] = s <= 30 ? p.bulge[s] : p.bulge[30] + (int)(p.lxc * log(s / 30.));
        m.internal_loop[s] = s <= 30 ? p.internal_loop[s] : p.internal_loop[30] + (int)(p.lxc * log(s / 30.));
    }
    m.ninio2 = p.ninio[2];
    m.max_ninio = 300;   // MAX_NINIO, src/ViennaRNA/params/constants.h
    m.MLbase = p.MLbase;
    m.MLclosing = p.MLclosing;
    m.TerminalAU = p.TerminalAU;
    memcpy(m.MLintern, p.MLintern, sizeof m.MLintern);
    auto special = [](const char *names, int pitch, int len, const int *e, char (*out)[8], int32_t *out_e, int32_t &count) {
        count = 0;
        const size_t total = strlen(names);
        for (int i = 0; i < CCJ_MAX_SPECIAL && (size_t)(i * pitch) < total; ++i) {
            memset(out[i], 0, 8);
            memcpy(out[i], names + i * pitch, len);
            out_e[i] = e[i];
            count = i + 1;
        }
    };
    special(p.Tetraloops, 7, 6, p.Tetraloop_E, m.tetra, m.tetra_E, m.n_tetra);
    special(p.Triloops, 6, 5, p.Triloop_E, m.tri, m.tri_E, m.n_tri);
    special(p.Hexaloops, 9, 8, p.Hexaloop_E, m.hexa, m.hexa_E, m.n_hexa);
    m.special_hp = p.model_details.special_hp;
    m.dangles = p.model_details.dangles;
    static const int bp[5][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 5}, {0, 0, 0, 1, 0}, {0, 0, 2, 0, 3}, {0, 6, 0, 4, 0}};
    for (int x = 0; x < 5; ++x)
        for (int y = 0; y < 5; ++y) m.pair[x][y] = bp[x][y];
    if (no_gu) m.pair[3][4] = m.pair[4][3] = 0;
    static const int rt[8] = {0, 2, 1, 4, 3, 6, 5, 7};
    for (int t = 0; t < 8; ++t) m.rtype[t] = rt[t];
    // the reference's recurrences read the penalty GLOBALS (src/h_externs.hh), not the struct fields: whatever the
    // program has set them to by now is what the GPU uses
    m.PS_penalty = ::PS_penalty; m.PSM_penalty = ::PSM_penalty; m.PSP_penalty = ::PSP_penalty; m.PB_penalty = ::PB_penalty;
    m.PUP_penalty = ::PUP_penalty; m.PPS_penalty = ::PPS_penalty; m.e_stP_penalty = ::e_stP_penalty;
    m.e_intP_penalty = ::e_intP_penalty; m.a_penalty = ::a_penalty; m.b_penalty = ::b_penalty; m.c_penalty = ::c_penalty;
    m.ap_penalty = ::ap_penalty; m.bp_penalty = ::bp_penalty; m.cp_penalty = ::cp_penalty;
}

void vrna_from_model(const ccj_model &m, vrna_param_t &p) {
    memcpy(p.stack, m.stack, sizeof p.stack);
    memcpy(p.hairpin, m.hairpin, sizeof p.hairpin);
    memcpy(p.bulge, m.bulge, sizeof p.bulge);
    memcpy(p.internal_loop, m.internal_loop, sizeof p.internal_loop);
    memcpy(p.mismatchExt, m.mismatchExt, sizeof p.mismatchExt);
    memcpy(p.mismatchI, m.mismatchI, sizeof p.mismatchI);
    memcpy(p.mismatch1nI, m.mismatch1nI, sizeof p.mismatch1nI);
    memcpy(p.mismatch23I, m.mismatch23I, sizeof p.mismatch23I);
    memcpy(p.mismatchH, m.mismatchH, sizeof p.mismatchH);
    memcpy(p.mismatchM, m.mismatchM, sizeof p.mismatchM);
    memcpy(p.dangle5, m.dangle5, sizeof p.dangle5);
    memcpy(p.dangle3, m.dangle3, sizeof p.dangle3);
    memcpy(p.int11, m.int11, sizeof p.int11);
    memcpy(p.int21, m.int21, sizeof p.int21);
    memcpy(p.int22, m.int22, sizeof p.int22);
    p.ninio[2] = m.ninio2;
    p.lxc = 107.856;   // src/ViennaRNA/params/default.c:64 (tempf = 1 at 37 C)
    p.MLbase = m.MLbase;
    memcpy(p.MLintern, m.MLintern, sizeof p.MLintern);
    p.MLclosing = m.MLclosing;
    p.TerminalAU = m.TerminalAU;
    p.DuplexInit = 410;
    auto special = [](const char (*in)[8], const int32_t *in_e, int count, int pitch, char *names, int *e) {
        for (int i = 0; i < count; ++i) {
            char *dst = names + i * pitch;
            const size_t l = strnlen(in[i], 8);
            memcpy(dst, in[i], l);
            memset(dst + l, ' ', pitch - l);
            e[i] = in_e[i];
        }
        names[count * pitch] = '\0';
        // the entry the file reader appends for its terminating line is an empty name: keep the string the reference
        // has (names followed by blanks) -- strlen decides how many entries model_from_vrna takes back
        size_t len = strlen(names);
        while (len >= 2 && names[len - 1] == ' ' && names[len - 2] == ' ' && count > 0 && in[count - 1][0] == '\0' &&
               len > (size_t)((count - 1) * pitch + 1))
            names[--len] = '\0';
    };
    special(m.tetra, m.tetra_E, m.n_tetra, 7, p.Tetraloops, p.Tetraloop_E);
    special(m.tri, m.tri_E, m.n_tri, 6, p.Triloops, p.Triloop_E);
    special(m.hexa, m.hexa_E, m.n_hexa, 9, p.Hexaloops, p.Hexaloop_E);
    p.PS_penalty = m.PS_penalty; p.PSM_penalty = m.PSM_penalty; p.PSP_penalty = m.PSP_penalty; p.PB_penalty = m.PB_penalty;
    p.PUP_penalty = m.PUP_penalty; p.PPS_penalty = m.PPS_penalty; p.e_stP_penalty = m.e_stP_penalty;
    p.e_intP_penalty = m.e_intP_penalty; p.ap_penalty = m.ap_penalty; p.bp_penalty = m.bp_penalty; p.cp_penalty = m.cp_penalty;
    p.a_penalty = m.a_penalty; p.b_penalty = m.b_penalty; p.c_penalty = m.c_penalty;
    p.temperature = 37.0;
    vrna_md_t &md = p.model_details;   // src/ViennaRNA/model.c:49-68 defaults
    md.temperature = 37.0;
    md.betaScale = 1.0;
    md.dangles = m.dangles;
    md.special_hp = m.special_hp;
    md.noGU = noGU;
    md.backtrack = 1;
    md.backtrack_type = 'F';
    md.compute_bpp = 1;
    md.max_bp_span = -1;
    md.min_loop_size = TURN;
    md.window_size = -1;
    md.cv_fact = 1.0;
    md.nc_fact = 1.0;
    md.sfact = 1.07;
    memcpy(md.rtype, m.rtype, sizeof md.rtype);
    for (int x = 0; x < 5; ++x) {
        md.alias[x] = (short)x;
        for (int y = 0; y < 5; ++y) md.pair[x][y] = m.pair[x][y];
    }
}

}  // namespace ccj

ccj_shell_fold::~ccj_shell_fold() {
    free(t4);
    if (g_resident == this) g_resident = nullptr;
}

void ccj_shell_fold::ensure_resident() {
    using ccj::shell_ctx;
    if (g_resident == this && filled) return;
    ccj_ctx *ctx = shell_ctx();
    if (ccj_model_upload(ctx, &model, sizeof model) != 0) die_abi(ctx, "model upload");
    const int64_t offsets[2] = {0, (int64_t)n};
    int rc = ccj_batch_prepare(ctx, seq.data(), offsets, 1);
    if (!rc) rc = ccj_batch_fill(ctx);
    if (rc) die_abi(ctx, "fill");
    g_resident = this;
    if (!filled) {
        t2.assign((size_t)(ccj_stride2(n) * CCJ_NT2), 0);
        if (ccj_copy_tables2_raw(ctx, 0, t2.data(), (int64_t)t2.size()) != 0) die_abi(ctx, "table fetch");
        filled = true;
    }
}

void ccj_shell_fold::ensure_filled() {
    if (!filled) ensure_resident();
}

void ccj_shell_fold::need4(int table) {
    using ccj::shell_ctx;
    if (have4[table]) return;
    ensure_resident();
    const int64_t cells = ccj_cells4(n);
    if (!t4) {
        t4 = static_cast<int16_t *>(calloc((size_t)(cells * CCJ_NT4) + 8, sizeof(int16_t)));   // pages commit on first touch
        if (!t4) {
            std::cerr << "ccj_b200: out of host memory for the table mirror" << std::endl;
            exit(EXIT_FAILURE);
        }
    }
    if (cells > 0 && ccj_copy_table4_raw(shell_ctx(), 0, table, t4 + (size_t)table * cells, cells) != 0)
        die_abi(shell_ctx(), "table fetch");
    have4[table] = true;
}

ccj_cx ccj_shell_fold::cx() {
    ensure_filled();
    ccj_cx c;
    memset(&c.q, 0, sizeof c.q);
    c.M = &model;
    c.q.n = n;
    c.q.S = S8.data();
    c.q.seq = seq.data();
    c.q.t4 = t4;
    c.q.stride4 = ccj_cells4(n);
    c.q.t2 = t2.data();
    c.q.stride2 = ccj_stride2(n);
    return c;
}

energy_t ccj_shell_fold::get4(int table, int i, int j, int k, int l) {
    if (!ccj_valid4(i, j, k, l) || i < 1 || l > n) return INF;
    need4(table);
    return t4[(size_t)table * ccj_cells4(n) + ccj_idx4(n, i, j, k, l)];
}

namespace ccj {

std::shared_ptr<ShellFold> shell_fold(const std::string &seq, const vrna_param_t *params) {
    ccj_model m;
    if (params) {
        model_from_vrna(*params, noGU, m);
    } else {
        vrna_param_t *p = scale_parameters();
        if (!p) {
            std::cerr << "Not a valid parameter file!" << std::endl;
            exit(EXIT_FAILURE);
        }
        model_from_vrna(*p, noGU, m);
        free(p);
    }
    for (auto it = g_folds.begin(); it != g_folds.end();) {
        std::shared_ptr<ShellFold> f = it->lock();
        if (!f) {
            it = g_folds.erase(it);
            continue;
        }
        if (f->seq == seq && memcmp(&f->model, &m, sizeof m) == 0) return f;
        ++it;
    }
    std::shared_ptr<ShellFold> f = std::make_shared<ShellFold>();
    f->seq = seq;
    f->n = (int)seq.size();
    f->model = m;
    f->S8.assign(f->n + 2, 0);
    for (int i = 1; i <= f->n; ++i) {
        const char c = seq[i - 1];
        f->S8[i] = (int8_t)(c == 'A' ? 1 : c == 'C' ? 2 : c == 'G' ? 3 : (c == 'U' || c == 'T') ? 4 : 0);
    }
    if (f->n > 0) {
        f->S8[f->n + 1] = f->S8[1];
        f->S8[0] = f->S8[f->n];
    }
    g_folds.push_back(f);
    return f;
}

}  // namespace ccj
