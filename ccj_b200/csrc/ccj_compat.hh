// ccj_compat.hh -- the plain types, constants and ViennaRNA entry points that code written against the reference's
// headers uses around W_final / pseudo_loop / s_energy_matrix, so that such code compiles against the B200 shells:
//   src/base_types.hh:6-15          energy_t, cand_pos_t, ...
//   src/matrices.hh:10,14-79        INF, TriangleMatrix (here: a read-only view of a device-resident table)
//   src/constants.hh:21-73          node type characters
//   src/h_struct.hh:9-103           minimum_fold, seq_interval, free_energy_node
//   src/ViennaRNA/params/basic.h:57-118, src/ViennaRNA/model.h:170-236   vrna_param_t / vrna_md_t, field for field
//   src/ViennaRNA/params/io.h, params/basic.h, pair_mat.h   vrna_params_load, scale_parameters, make_pair_matrix,
//                                                           encode_sequence
// vrna_param_t mirrors the reference's struct member by member and in order (tests/shell/layout_probe.cc compares
// every offset with the reference's own header), so `params_->stack[..]`, `params->model_details.dangles` etc. mean
// the same thing.  The GPU does not read this struct directly: the shells convert it to the library's model blob
// (ccj_model_upload), so edits a caller makes to a vrna_param_t before constructing the tables are honoured.
#ifndef CCJ_B200_COMPAT_HH
#define CCJ_B200_COMPAT_HH
#include <cstdint>
#include <string>
#include <vector>

#include "ccj_b200.h"
#include "ccj_types.h"

typedef int_least32_t energy_t;
typedef int_least16_t energy_16t;
typedef int_least32_t cand_pos_t;
typedef uint_least32_t cand_pos_tu;
typedef int_least16_t pair_type;
typedef int_least16_t base_type;
typedef double pf_t;

#ifndef INF
#define INF CCJ_INF
#endif
#ifndef TURN
#define TURN CCJ_TURN
#endif
#ifndef MAXLOOP
#define MAXLOOP CCJ_MAXLOOP
#endif
#define NBPAIRS 7
#define MAXALPHA 20
#define VRNA_PARAMETER_FORMAT_DEFAULT 0
#define VRNA_GQUAD_MAX_STACK_SIZE 7
#define VRNA_GQUAD_MAX_LINKER_LENGTH 15

// ---- src/constants.hh:21-73 ------------------------------------------------------------------------------------
#define NONE 'N'
#define HAIRP 'H'
#define INTER 'I'
#define MULTI 'M'
#define M_WM 'B'
#define M_WMv 'v'
#define M_WMp 'p'
#define FREE 'W'
#define LOOP 'V'
#define P_P 'P'
#define P_PK 'k'
#define P_PL 'l'
#define P_PR 'r'
#define P_PM 'm'
#define P_PO 'o'
#define P_PfromL 'f'
#define P_PfromR 'g'
#define P_PfromM 'h'
#define P_PfromMprime '['
#define P_PfromMdoubleprime ']'
#define P_PfromO 'i'
#define P_PLiloop 'j'
#define P_PLiloop5 'b'
#define P_PLmloop 'c'
#define P_PLmloop10 'e'
#define P_PLmloop01 'n'
#define P_PLmloop00 'a'
#define P_PRiloop 'q'
#define P_PRiloop5 's'
#define P_PRmloop 't'
#define P_PRmloop10 'u'
#define P_PRmloop01 '&'
#define P_PRmloop00 '9'
#define P_PMiloop 'w'
#define P_PMiloop5 'x'
#define P_PMmloop 'y'
#define P_PMmloop10 '0'
#define P_PMmloop01 '1'
#define P_PMmloop00 '8'
#define P_POiloop 'z'
#define P_POiloop5 '5'
#define P_POmloop '+'
#define P_POmloop10 '-'
#define P_POmloop01 '='
#define P_POmloop00 '_'
#define P_WB '*'
#define P_WBP '^'
#define P_WP '#'
#define P_WPP '@'

// ---- src/h_struct.hh ---------------------------------------------------------------------------------------------
typedef struct minimum_fold {
    cand_pos_t pair;
    char type;
    minimum_fold() : pair(-1), type(NONE) {}
} minimum_fold;

struct seq_interval {
    cand_pos_t i;
    cand_pos_t j;
    energy_t energy;
    char type;
    seq_interval *next = nullptr;
    cand_pos_t k;   // the gapped region is [i,k] U [l,j] (src/h_struct.hh:72-76)
    cand_pos_t l;
    cand_pos_t asym;
    void copy(seq_interval *other) {
        other->i = i; other->j = j; other->energy = energy; other->type = type;
        other->k = k; other->l = l; other->asym = asym;
    }
};

struct free_energy_node {
    int energy;
    char type;
    free_energy_node() : energy(10000), type(NONE) {}
};

// ---- src/ViennaRNA/model.h:170-236 ------------------------------------------------------------------------------
struct vrna_md_s {
    double temperature;
    double betaScale;
    int pf_smooth;
    int dangles;
    int special_hp;
    int noLP;
    int noGU;
    int noGUclosure;
    int logML;
    int circ;
    int gquad;
    int uniq_ML;
    int energy_set;
    int backtrack;
    char backtrack_type;
    int compute_bpp;
    char nonstandards[64];
    int max_bp_span;
    int min_loop_size;
    int window_size;
    int oldAliEn;
    int ribo;
    double cv_fact;
    double nc_fact;
    double sfact;
    int rtype[8];
    short alias[MAXALPHA + 1];
    int pair[MAXALPHA + 1][MAXALPHA + 1];
};
typedef struct vrna_md_s vrna_md_t;

// ---- src/ViennaRNA/params/basic.h:57-118 -------------------------------------------------------------------------
struct vrna_param_s {
    int id;
    int stack[NBPAIRS + 1][NBPAIRS + 1];
    int hairpin[31];
    int bulge[MAXLOOP + 1];
    int internal_loop[MAXLOOP + 1];
    int mismatchExt[NBPAIRS + 1][5][5];
    int mismatchI[NBPAIRS + 1][5][5];
    int mismatch1nI[NBPAIRS + 1][5][5];
    int mismatch23I[NBPAIRS + 1][5][5];
    int mismatchH[NBPAIRS + 1][5][5];
    int mismatchM[NBPAIRS + 1][5][5];
    int dangle5[NBPAIRS + 1][5];
    int dangle3[NBPAIRS + 1][5];
    int int11[NBPAIRS + 1][NBPAIRS + 1][5][5];
    int int21[NBPAIRS + 1][NBPAIRS + 1][5][5][5];
    int int22[NBPAIRS + 1][NBPAIRS + 1][5][5][5][5];
    int ninio[5];
    double lxc;
    int MLbase;
    int MLintern[NBPAIRS + 1];
    int MLclosing;
    int PS_penalty;
    int PSM_penalty;
    int PSP_penalty;
    int PB_penalty;
    int PUP_penalty;
    int PPS_penalty;
    double e_stP_penalty;
    double e_intP_penalty;
    int ap_penalty;
    int bp_penalty;
    int cp_penalty;
    int a_penalty;
    int b_penalty;
    int c_penalty;
    int TerminalAU;
    int DuplexInit;
    int Tetraloop_E[200];
    char Tetraloops[1401];
    int Triloop_E[40];
    char Triloops[241];
    int Hexaloop_E[40];
    char Hexaloops[1801];
    int TripleC;
    int MultipleCA;
    int MultipleCB;
    int gquad[VRNA_GQUAD_MAX_STACK_SIZE + 1][3 * VRNA_GQUAD_MAX_LINKER_LENGTH + 1];
    int gquadLayerMismatch;
    int gquadLayerMismatchMax;
    double temperature;
    vrna_md_t model_details;
    char param_file[256];
};
typedef struct vrna_param_s vrna_param_t;
typedef struct vrna_param_s paramT;

// ---- process-global configuration, as in the reference ----------------------------------------------------------------
extern int noGU;   // ViennaRNA global read by make_pair_matrix (src/CCJ.cc:77)

// vrna_params_load (src/ViennaRNA/params/io.c:252): remembers the file for the next scale_parameters(); non-zero on success
int vrna_params_load(const char *fname, unsigned int options);
// vrna_params_load_DNA_Mathews2004: the set linked into libccj_b200.so
int vrna_params_load_DNA_Mathews2004(void);
// scale_parameters (src/ViennaRNA/params/params.c:961): a malloc'ed vrna_param_t with the 37 C values of the loaded
// set (default: the compiled-in Turner-2004 set), dangles=2 at scaling time; the caller frees it with free()
vrna_param_t *scale_parameters(void);
// encode_sequence (src/ViennaRNA/pair_mat.h:158-183): malloc'ed short[n+2]; how=0: S[0]=n, how=1: S1 with S1[0]=S1[n],
// S1[n+1]=S1[1]
short *encode_sequence(const char *sequence, short how);

// pair[][] / rtype[] are per-translation-unit statics in the reference (src/ViennaRNA/pair_mat.h:33-38,80-155)
static int rtype[8] = {0, 2, 1, 4, 3, 6, 5, 7};
static int pair[MAXALPHA + 1][MAXALPHA + 1];
static inline void make_pair_matrix(void) {
    static const int bp[5][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 5}, {0, 0, 0, 1, 0}, {0, 0, 2, 0, 3}, {0, 6, 0, 4, 0}};
    for (int x = 0; x <= MAXALPHA; ++x)
        for (int y = 0; y <= MAXALPHA; ++y) pair[x][y] = (x < 5 && y < 5) ? bp[x][y] : 0;
    if (noGU) pair[3][4] = pair[4][3] = 0;
    (void)rtype;
}

// ---- TriangleMatrix: the public face of src/matrices.hh:14-79 over a device-resident 2D table ------------------------
// (pseudo_loop::P is a public member of the reference and W_final::ccj passes it to compute_energy_WM)
struct ccj_shell_fold;   // the shared bulk fold behind the shell objects (ccj_shell.hpp)
class TriangleMatrix {
public:
    TriangleMatrix() : fold_(nullptr), table_(0), return_val_(INF) {}
    void bind(ccj_shell_fold *fold, int table, energy_t return_val = INF) { fold_ = fold; table_ = table; return_val_ = return_val; }
    energy_t get_uc(cand_pos_t i, cand_pos_t j) const;
    energy_t get(cand_pos_t i, cand_pos_t j) const { return i > j ? return_val_ : get_uc(i, j); }

private:
    ccj_shell_fold *fold_;
    int table_;
    energy_t return_val_;
};
#endif
