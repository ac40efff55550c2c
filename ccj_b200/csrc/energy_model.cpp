// See energy_model.hpp for the reference locations this file replaces.
#include "energy_model.hpp"

#include <algorithm>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <vector>

namespace ccj {

namespace {

const int kDEF = -50; /* io.c:26 */
const int kNST = 0;   /* io.c:27 */

enum Section {
    SEC_stack, SEC_hairpin, SEC_bulge, SEC_interior, SEC_mm_ext, SEC_mm_hp, SEC_mm_int, SEC_mm_1n, SEC_mm_23,
    SEC_mm_multi, SEC_int11, SEC_int21, SEC_int22, SEC_d5, SEC_d3, SEC_ML, SEC_NINIO, SEC_MISC, SEC_TL, SEC_TRI,
    SEC_HEX, SEC_COUNT
};

struct Reader {
    std::vector<std::string> lines;
    size_t pos = 0;
    std::string err;
    double lxc37;

    bool next_line(std::string &out) {
        if (pos >= lines.size()) return false;
        out = lines[pos++];
        return true;
    }

    // io.c:1100-1121 -- one C comment per line is cut out
    static bool strip_comment(std::string &line, std::string &err) {
        size_t c1 = line.find("/*");
        if (c1 == std::string::npos) return true;
        size_t c2 = line.find("*/", c1);
        if (c2 == std::string::npos) {
            err = "unclosed comment in parameter file";
            return false;
        }
        line.erase(c1, c2 + 2 - c1);
        return true;
    }

    // io.c:713-764 get_array1: consumes whole lines until `size` entries are set; the rest of the last
    // line is dropped.
    bool read_block(int *arr, int size) {
        int i = 0, last = 0;
        while (i < size) {
            std::string line;
            if (!next_line(line)) {
                err = "unexpected end of file in parameter array";
                return false;
            }
            if (!strip_comment(line, err)) return false;
            std::istringstream is(line);
            std::string tok;
            while (i < size && (is >> tok)) {
                if (tok.size() > 15) tok.resize(15); /* %15s */
                int p;
                if (tok[0] == '*') {
                    ++i;
                    continue;
                } else if (tok[0] == 'x') {
                    if (i == 0) {
                        err = "can't extrapolate first value";
                        return false;
                    }
                    p = arr[last] + (int)(0.5 + lxc37 * log(((double)i) / (double)(last)));
                } else if (tok == "DEF") {
                    p = kDEF;
                } else if (tok == "INF") {
                    p = CCJ_INF;
                } else if (tok == "NST") {
                    p = kNST;
                } else {
                    if (sscanf(tok.c_str(), "%d", &p) != 1) {
                        err = "can't interpret `" + tok + "' in parameter file";
                        return false;
                    }
                    last = i;
                }
                arr[i++] = p;
            }
        }
        return true;
    }

    // io.c:767-1007 rd_<N>dim_slice family, expressed once
    bool read_slice(int *array, const int *dim, const int *shift, const int *post, int nd) {
        int delta = 0;
        long total = 1;
        for (int d = 0; d < nd; ++d) {
            delta += shift[d] + post[d];
            total *= dim[d];
        }
        if (delta == 0) return read_block(array, (int)total);
        if (nd == 1) return read_block(array + shift[0], dim[0] - shift[0] - post[0]);
        long sub = total / dim[0];
        for (int i = shift[0]; i < dim[0] - post[0]; ++i)
            if (!read_slice(array + i * sub, dim + 1, shift + 1, post + 1, nd - 1)) return false;
        return true;
    }

    // io.c:1011-1078: "%Ns %d %d" rows until one does not parse; the terminating line is consumed and a
    // blank is appended even for it.
    void read_special(char *names, int pitch, int *e37, size_t names_cap) {
        memset(names, 0, names_cap);
        memset(e37, 0, sizeof(int) * 40);
        int i = 0, r;
        do {
            std::string line;
            if (!next_line(line)) break;
            char name[16] = {0};
            int g = 0, h = 0;
            char fmt[32];
            snprintf(fmt, sizeof fmt, "%%%ds %%d %%d", pitch - 1);
            r = sscanf(line.c_str(), fmt, name, &g, &h);
            if (r >= 1) memcpy(names + pitch * i, name, strlen(name) + 1);
            if (r >= 2) e37[i] = g;
            size_t len = strlen(names);
            if (len + 1 < names_cap) {
                names[len] = ' ';
                names[len + 1] = '\0';
            }
            ++i;
        } while (r == 3 && i < 40);
    }
};

int section_of(const std::string &id, bool &enthalpy) {
    static const struct { const char *name; int sec; } tab[] = {
        {"stack", SEC_stack}, {"hairpin", SEC_hairpin}, {"bulge", SEC_bulge}, {"interior", SEC_interior},
        {"mismatch_exterior", SEC_mm_ext}, {"mismatch_hairpin", SEC_mm_hp}, {"mismatch_interior", SEC_mm_int},
        {"mismatch_interior_1n", SEC_mm_1n}, {"mismatch_interior_23", SEC_mm_23}, {"mismatch_multi", SEC_mm_multi},
        {"int11", SEC_int11}, {"int21", SEC_int21}, {"int22", SEC_int22}, {"dangle5", SEC_d5}, {"dangle3", SEC_d3},
    };
    enthalpy = false;
    std::string base = id;
    const std::string suf = "_enthalpies";
    if (base.size() > suf.size() && base.compare(base.size() - suf.size(), suf.size(), suf) == 0) {
        base.resize(base.size() - suf.size());
        enthalpy = true;
    }
    for (auto &t : tab)
        if (base == t.name) return t.sec;
    if (enthalpy) return -1;
    if (id == "ML_params") return SEC_ML;
    if (id == "NINIO") return SEC_NINIO;
    if (id == "Misc") return SEC_MISC;
    if (id == "Tetraloops") return SEC_TL;
    if (id == "Triloops") return SEC_TRI;
    if (id == "Hexaloops") return SEC_HEX;
    return -1;
}

template <class T> void fill_all(T &arr, int v) {
    int *p = reinterpret_cast<int *>(&arr);
    for (size_t x = 0; x < sizeof(arr) / sizeof(int); ++x) p[x] = v;
}

}  // namespace

// Blank state: pair-type row/column 0 is INF everywhere (src/ViennaRNA/params/default.c), scalar defaults of
// default.c:64-76, arrays INF.  load_defaults() turns it into what the reference holds before any file is read
// (its compiled-in Turner-2004 tables), so that sections a file omits keep those values like in the reference.
RawParams::RawParams() {
    fill_all(stack, CCJ_INF);
    fill_all(hairpin, CCJ_INF);
    fill_all(bulge, CCJ_INF);
    fill_all(internal_loop, CCJ_INF);
    fill_all(mismatchI, CCJ_INF);
    fill_all(mismatchH, CCJ_INF);
    fill_all(mismatchM, CCJ_INF);
    fill_all(mismatch1nI, CCJ_INF);
    fill_all(mismatch23I, CCJ_INF);
    fill_all(mismatchExt, CCJ_INF);
    fill_all(dangle5, CCJ_INF);
    fill_all(dangle3, CCJ_INF);
    fill_all(int11, CCJ_INF);
    fill_all(int21, CCJ_INF);
    fill_all(int22, CCJ_INF);
    ML_BASE = 0;      /* default.c:69 */
    ML_closing = 930; /* default.c:67 */
    ML_intern = -90;  /* default.c:65 */
    ninio = 60;       /* default.c:72 */
    MAX_NINIO = 300;  /* default.c:71 */
    DuplexInit = 410; /* default.c:76 */
    TerminalAU = 50;  /* default.c:74 */
    lxc = 107.856;    /* default.c:64 */
    memset(Tetraloops, 0, sizeof Tetraloops);
    memset(Triloops, 0, sizeof Triloops);
    memset(Hexaloops, 0, sizeof Hexaloops);
    memset(Tetraloop_E, 0, sizeof Tetraloop_E);
    memset(Triloop_E, 0, sizeof Triloop_E);
    memset(Hexaloop_E, 0, sizeof Hexaloop_E);
    present = 0;
}

bool load_par_file(const char *path, RawParams &rp, std::string &err) {
    std::ifstream in(path, std::ios::binary);
    if (!in) {
        err = std::string("cannot open parameter file ") + path;
        return false;
    }
    std::stringstream ss;
    ss << in.rdbuf();
    const std::string text = ss.str();
    return load_par_text(text.data(), text.size(), rp, err);
}

// The reference's compiled-in defaults (src/ViennaRNA/params/default.c) are the Turner-2004 set; its own
// params/rna_Turner04.par holds the same numbers (checked against a dump of the compiled reference after loading
// a file without sections, tests/test_layout_params.py).  The special-loop strings of the defaults lack the extra
// blank entry the file reader appends (io.c:1027-1040).
bool load_defaults(RawParams &rp, std::string &err) {
    size_t len = 0;
    const char *text = embedded_par("rna_turner2004", &len);
    if (!text) {
        err = "embedded default parameter set missing";
        return false;
    }
    if (!load_par_text(text, len, rp, err)) return false;
    for (char *names : {rp.Tetraloops, rp.Triloops, rp.Hexaloops}) {
        const size_t l = strlen(names);
        if (l >= 2 && names[l - 1] == ' ' && names[l - 2] == ' ') names[l - 1] = '\0';
    }
    rp.present = 0;
    rp.warnings.clear();
    return true;
}

bool load_par_text(const char *text, size_t len, RawParams &rp, std::string &err) {
    Reader rd;
    rd.lxc37 = rp.lxc;
    std::string line;
    std::istringstream in(std::string(text, len));
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        rd.lines.push_back(line);
    }
    if (rd.lines.empty()) {
        err = "empty parameter file";
        return false;
    }
    rd.pos = 1;  // header line is skipped whatever it says (io.c:477-481 only warns)

    static const int d_stack[2] = {8, 8}, s_stack[2] = {1, 1}, p0[6] = {0, 0, 0, 0, 0, 0};
    static const int d_mm[3] = {8, 5, 5}, s_mm[3] = {1, 0, 0};
    static const int d_11[4] = {8, 8, 5, 5}, s_11[4] = {1, 1, 0, 0};
    static const int d_21[5] = {8, 8, 5, 5, 5}, s_21[5] = {1, 1, 0, 0, 0};
    static const int d_22[6] = {8, 8, 5, 5, 5, 5}, s_22[6] = {1, 1, 1, 1, 1, 1}, p_22[6] = {1, 1, 0, 0, 0, 0};
    static const int d_dg[2] = {8, 5}, s_dg[2] = {1, 0};

    std::vector<int> scratch(8 * 8 * 5 * 5 * 5 * 5);
    // enthalpy copies that check_symmetry looks at (start = INF everywhere, like row/column 0 of the defaults)
    std::vector<int> stackdH, int11dH, int22dH;
    while (rd.next_line(line)) {
        char ident[256];
        if (sscanf(line.c_str(), "# %255s", ident) != 1) continue;
        bool enth;
        int sec = section_of(ident, enth);
        if (sec < 0) continue;  // "END" and unknown identifiers: nothing to read (io.c:484,664)
        bool ok = true;
        int *dst;
        if (enth) std::fill(scratch.begin(), scratch.end(), CCJ_INF);
#define TARGET(field) (enth ? scratch.data() : &rp.field)
        switch (sec) {
            case SEC_stack: dst = TARGET(stack[0][0]); ok = rd.read_slice(dst, d_stack, s_stack, p0, 2); break;
            case SEC_hairpin: dst = TARGET(hairpin[0]); ok = rd.read_block(dst, 31); break;
            case SEC_bulge: dst = TARGET(bulge[0]); ok = rd.read_block(dst, 31); break;
            case SEC_interior: dst = TARGET(internal_loop[0]); ok = rd.read_block(dst, 31); break;
            case SEC_mm_ext: dst = TARGET(mismatchExt[0][0][0]); ok = rd.read_slice(dst, d_mm, s_mm, p0, 3); break;
            case SEC_mm_hp: dst = TARGET(mismatchH[0][0][0]); ok = rd.read_slice(dst, d_mm, s_mm, p0, 3); break;
            case SEC_mm_int: dst = TARGET(mismatchI[0][0][0]); ok = rd.read_slice(dst, d_mm, s_mm, p0, 3); break;
            case SEC_mm_1n: dst = TARGET(mismatch1nI[0][0][0]); ok = rd.read_slice(dst, d_mm, s_mm, p0, 3); break;
            case SEC_mm_23: dst = TARGET(mismatch23I[0][0][0]); ok = rd.read_slice(dst, d_mm, s_mm, p0, 3); break;
            case SEC_mm_multi: dst = TARGET(mismatchM[0][0][0]); ok = rd.read_slice(dst, d_mm, s_mm, p0, 3); break;
            case SEC_int11: dst = TARGET(int11[0][0][0][0]); ok = rd.read_slice(dst, d_11, s_11, p0, 4); break;
            case SEC_int21: dst = TARGET(int21[0][0][0][0][0]); ok = rd.read_slice(dst, d_21, s_21, p0, 5); break;
            case SEC_int22: dst = TARGET(int22[0][0][0][0][0][0]); ok = rd.read_slice(dst, d_22, s_22, p_22, 6); break;
            case SEC_d5: dst = TARGET(dangle5[0][0]); ok = rd.read_slice(dst, d_dg, s_dg, p0, 2); break;
            case SEC_d3: dst = TARGET(dangle3[0][0]); ok = rd.read_slice(dst, d_dg, s_dg, p0, 2); break;
            case SEC_ML: {
                int v[6] = {rp.ML_BASE, 0, rp.ML_closing, 0, rp.ML_intern, 0};
                ok = rd.read_block(v, 6);
                rp.ML_BASE = v[0];
                rp.ML_closing = v[2];
                rp.ML_intern = v[4];
            } break;
            case SEC_NINIO: {
                int v[3] = {rp.ninio, 0, rp.MAX_NINIO};
                ok = rd.read_block(v, 3);
                rp.ninio = v[0];
                rp.MAX_NINIO = v[2];
            } break;
            case SEC_MISC: {
                int v[4] = {rp.DuplexInit, 0, rp.TerminalAU, 0};
                ok = rd.read_block(v, 4);
                rp.DuplexInit = v[0];
                rp.TerminalAU = v[2];
            } break;
            case SEC_TL: rd.read_special(rp.Tetraloops, 7, rp.Tetraloop_E, sizeof rp.Tetraloops); break;
            case SEC_TRI: rd.read_special(rp.Triloops, 6, rp.Triloop_E, sizeof rp.Triloops); break;
            case SEC_HEX: rd.read_special(rp.Hexaloops, 9, rp.Hexaloop_E, sizeof rp.Hexaloops); break;
        }
#undef TARGET
        if (!ok) {
            err = rd.err + " (section " + ident + ")";
            return false;
        }
        if (!enth) rp.present |= 1u << sec;
        else if (sec == SEC_stack) stackdH.assign(scratch.begin(), scratch.begin() + 8 * 8);
        else if (sec == SEC_int11) int11dH.assign(scratch.begin(), scratch.begin() + 8 * 8 * 5 * 5);
        else if (sec == SEC_int22) int22dH.assign(scratch.begin(), scratch.begin() + 8 * 8 * 5 * 5 * 5 * 5);
    }
    // check_symmetry (io.c:1126-1178), same order and texts.  Sections the file does not hold keep the
    // reference's built-in (symmetric) tables there; the int22 checks cover the entries a file provides
    // (pairs 1..6, bases 1..4), the non-standard ones are derived maxima (update_nst).
    char buf[160];
    if (rp.present & (1u << SEC_stack))
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j)
                if (rp.stack[i][j] != rp.stack[j][i]) rp.warnings.push_back("stacking energies not symmetric");
    if (!stackdH.empty())
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j)
                if (stackdH[i * 8 + j] != stackdH[j * 8 + i]) rp.warnings.push_back("stacking enthalpies not symmetric");
    if (rp.present & (1u << SEC_int11))
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j)
                for (int k = 0; k < 5; ++k)
                    for (int l = 0; l < 5; ++l)
                        if (rp.int11[i][j][k][l] != rp.int11[j][i][l][k]) {
                            snprintf(buf, sizeof buf, "int11 energies not symmetric (%d,%d,%d,%d) (%d vs. %d)", i, j, k, l,
                                     rp.int11[i][j][k][l], rp.int11[j][i][l][k]);
                            rp.warnings.push_back(buf);
                        }
    if (!int11dH.empty()) {
        auto at = [&](int i, int j, int k, int l) { return int11dH[((i * 8 + j) * 5 + k) * 5 + l]; };
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j)
                for (int k = 0; k < 5; ++k)
                    for (int l = 0; l < 5; ++l)
                        if (at(i, j, k, l) != at(j, i, l, k)) rp.warnings.push_back("int11 enthalpies not symmetric");
    }
    if (rp.present & (1u << SEC_int22))
        for (int i = 1; i < 7; ++i)
            for (int j = 1; j < 7; ++j)
                for (int k = 1; k < 5; ++k)
                    for (int l = 1; l < 5; ++l)
                        for (int m = 1; m < 5; ++m)
                            for (int n = 1; n < 5; ++n)
                                if (rp.int22[i][j][k][l][m][n] != rp.int22[j][i][m][n][k][l])
                                    rp.warnings.push_back("int22 energies not symmetric");
    if (!int22dH.empty()) {
        auto at = [&](int i, int j, int k, int l, int m, int n) { return int22dH[((((i * 8 + j) * 5 + k) * 5 + l) * 5 + m) * 5 + n]; };
        for (int i = 1; i < 7; ++i)
            for (int j = 1; j < 7; ++j)
                for (int k = 1; k < 5; ++k)
                    for (int l = 1; l < 5; ++l)
                        for (int m = 1; m < 5; ++m)
                            for (int n = 1; n < 5; ++n)
                                if (at(i, j, k, l, m, n) != at(j, i, m, n, k, l)) {
                                    snprintf(buf, sizeof buf, "int22 enthalpies not symmetric: %d %d %d %d %d %d", i, j, k, l, m, n);
                                    rp.warnings.push_back(buf);
                                }
    }
    return true;
}

int encode_base(char c) {
    switch (c) {
        case 'A': case 'a': return 1;
        case 'C': case 'c': return 2;
        case 'G': case 'g': return 3;
        case 'U': case 'u': case 'T': case 't': return 4;
        default: return 0;
    }
}

static void copy_special(const char *names, int pitch, int len, const int *e, char (*out)[8], int32_t *out_e,
                         int32_t &count) {
    // number of entries the reference scales: (i*pitch) < strlen(names) (params.c:444-451; note it
    // uses 5 for triloops there, but only entries a probe can hit matter: an entry must be `len` chars)
    count = 0;
    size_t total = strlen(names);
    for (int i = 0; i < CCJ_MAX_SPECIAL && (size_t)(i * pitch) < total; ++i) {
        memset(out[i], 0, 8);
        memcpy(out[i], names + i * pitch, len);
        out_e[i] = e[i];
        count = i + 1;
    }
}

void build_model(const RawParams &rp, int dangles, int noGU, ccj_model &m) {
    memset(&m, 0, sizeof m);
    memcpy(m.stack, rp.stack, sizeof m.stack);
    memcpy(m.mismatchI, rp.mismatchI, sizeof m.mismatchI);
    memcpy(m.mismatchH, rp.mismatchH, sizeof m.mismatchH);
    memcpy(m.mismatch1nI, rp.mismatch1nI, sizeof m.mismatch1nI);
    memcpy(m.mismatch23I, rp.mismatch23I, sizeof m.mismatch23I);
    memcpy(m.int11, rp.int11, sizeof m.int11);
    memcpy(m.int21, rp.int21, sizeof m.int21);
    memcpy(m.int22, rp.int22, sizeof m.int22);
    // params.c:486-512: scaling happens with md->dangles = 2 (non-zero), so mismatchM/Ext and the
    // dangles are clamped to <= 0
    for (int t = 0; t < 8; ++t)
        for (int x = 0; x < 5; ++x) {
            for (int y = 0; y < 5; ++y) {
                m.mismatchM[t][x][y] = rp.mismatchM[t][x][y] > 0 ? 0 : rp.mismatchM[t][x][y];
                m.mismatchExt[t][x][y] = rp.mismatchExt[t][x][y] > 0 ? 0 : rp.mismatchExt[t][x][y];
            }
            m.dangle5[t][x] = rp.dangle5[t][x] > 0 ? 0 : rp.dangle5[t][x];
            m.dangle3[t][x] = rp.dangle3[t][x] > 0 ? 0 : rp.dangle3[t][x];
        }
    const double lxc = rp.lxc;  // lxc37 * tempf, tempf == 1 at 37 C
    for (int s = 0; s < CCJ_HAIRPIN_TAB; ++s)
        m.hairpin[s] = s <= 30 ? rp.hairpin[s] : rp.hairpin[30] + (int)(lxc * log((s) / 30.));
    for (int s = 0; s < CCJ_LOOP_TAB; ++s) {
        m.bulge[s] = s <= 30 ? rp.bulge[s] : rp.bulge[30] + (int)(lxc * log(s / 30.));
        m.internal_loop[s] = s <= 30 ? rp.internal_loop[s] : rp.internal_loop[30] + (int)(lxc * log((s) / 30.));
    }
    m.ninio2 = rp.ninio;
    m.max_ninio = rp.MAX_NINIO;
    m.MLbase = rp.ML_BASE;
    m.MLclosing = rp.ML_closing;
    m.TerminalAU = rp.TerminalAU;
    for (int t = 0; t < 8; ++t) m.MLintern[t] = rp.ML_intern;
    copy_special(rp.Tetraloops, 7, 6, rp.Tetraloop_E, m.tetra, m.tetra_E, m.n_tetra);
    copy_special(rp.Triloops, 6, 5, rp.Triloop_E, m.tri, m.tri_E, m.n_tri);
    copy_special(rp.Hexaloops, 9, 8, rp.Hexaloop_E, m.hexa, m.hexa_E, m.n_hexa);
    m.special_hp = 1; /* model.c default */
    m.dangles = dangles;

    // pair_mat.h:20-29 BP_pair restricted to _ACGU; noGU clears GU/UG (pair_mat.h:94-95)
    static const int bp[5][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 5}, {0, 0, 0, 1, 0}, {0, 0, 2, 0, 3}, {0, 6, 0, 4, 0}};
    for (int x = 0; x < 5; ++x)
        for (int y = 0; y < 5; ++y) m.pair[x][y] = bp[x][y];
    if (noGU) m.pair[3][4] = m.pair[4][3] = 0;
    static const int rt[8] = {0, 2, 1, 4, 3, 6, 5, 7};
    for (int t = 0; t < 8; ++t) m.rtype[t] = rt[t];

    /* src/h_globals.hh:7-25 */
    m.PS_penalty = -138;
    m.PSM_penalty = 1007;
    m.PSP_penalty = 1500;
    m.PB_penalty = 246;
    m.PUP_penalty = 6;
    m.PPS_penalty = 96;
    m.e_stP_penalty = 0.89;
    m.e_intP_penalty = 0.74;
    m.a_penalty = 339;
    m.b_penalty = 3;
    m.c_penalty = 2;
    m.ap_penalty = 341;
    m.bp_penalty = 56;
    m.cp_penalty = 12;
}

}  // namespace ccj
