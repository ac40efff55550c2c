// Host-side rendering of a finished traceback: f[].pair -> dot-bracket, and the reference's
// stdout/stderr/exit-code conventions.
//   W_final::fill_structure  src/W_final.cc:764-819
//   CCJ.cc:107-108           "SEQ\nSTRUCT (ENERGY)\n", energy = W[n]/100.0 through operator<<(double)
//   pseudo_loop.cc std::cerr messages (see ccj_traceback.cuh for the id encoding)
#pragma once
#include <cstdio>
#include <string>
#include <vector>

#include "ccj_types.h"

namespace ccj {

// pair[1..n] (-1 = unpaired) -> structure of length n
inline std::string fill_structure(int n, const int32_t *pair) {
    struct Band {
        char open, close;
        int outer_start, outer_end, inner_start, inner_end;
    };
    std::string structure(n + 1, '.');
    // bracket stack, bottom -> top: <> {} [] ()
    std::vector<std::pair<char, char>> st = {{'<', '>'}, {'{', '}'}, {'[', ']'}, {'(', ')'}};
    std::vector<Band> bands;
    bands.push_back(Band{'|', '|', 0, 0, 0, 0});
    for (int i = 1; i <= n; i++) {
        const int j = pair[i];
        if (j == -1) {
            structure[i] = '.';
        } else if (i < j) {
            bool in_band = false;
            for (Band &b : bands) {
                if (i > b.inner_start && j < b.inner_end) {
                    b.inner_start = i;
                    b.inner_end = j;
                    structure[i] = b.open;
                    structure[j] = b.close;
                    in_band = true;
                    break;
                }
            }
            if (!in_band) {
                // the reference pops without an emptiness check (a fifth open band family is UB there)
                std::pair<char, char> e = st.empty() ? std::make_pair('?', '?') : st.back();
                if (!st.empty()) st.pop_back();
                bands.push_back(Band{e.first, e.second, i, j, i, j});
                structure[i] = e.first;
                structure[j] = e.second;
            }
        } else {
            for (Band &b : bands) {
                if (i == b.outer_end) {
                    st.push_back({b.open, b.close});
                    break;
                }
            }
        }
    }
    return structure.substr(1, n);
}

inline std::string traceback_message(int msg) {
    static const char *prefix[] = {"", "border case: ", "border cases: ", "boder cases: ", "impossible case: ",
                                   "impossible cases: ", "impossbible cases: "};
    struct Name { char type; const char *name; };
    static const Name names[] = {
        {'P', "P_P"}, {'k', "P_PK"}, {'l', "P_PL"}, {'r', "P_PR"}, {'m', "P_PM"}, {'o', "P_PO"}, {'f', "P_PfromL"},
        {'g', "P_PfromR"}, {'h', "P_PfromM"}, {'[', "P_PfromMprime"}, {']', "P_PfromMdoubleprime"}, {'i', "P_PfromO"},
        {'j', "P_PLiloop"}, {'c', "P_PLmloop"}, {'e', "P_PLmloop10"}, {'n', "P_PLmloop01"}, {'a', "P_PLmloop00"},
        {'q', "P_PRiloop"}, {'t', "P_PRmloop"}, {'u', "P_PRmloop10"}, {'&', "P_PRmloop01"}, {'9', "P_PRmloop00"},
        {'w', "P_PMiloop"}, {'y', "P_PMmloop"}, {'0', "P_PMmloop10"}, {'1', "P_PMmloop01"}, {'8', "P_PMmloop00"},
        {'z', "P_POiloop"}, {'+', "P_POmloop"}, {'-', "P_POmloop10"}, {'=', "P_POmloop01"}, {'_', "P_POmloop00"},
        {'*', "P_WB"}, {'^', "P_WBP"}, {'#', "P_WP"}, {'@', "P_WPP"}};
    const int p = msg / 256;
    const char t = (char)(msg % 256);
    std::string s = (p >= 0 && p < 7) ? prefix[p] : "";
    s += "This should not have happened!, ";
    for (const Name &nm : names)
        if (nm.type == t) return s + nm.name;
    return s + "?";
}

// energy text exactly as `std::cout << double` prints it (precision 6, %g rules)
inline std::string energy_text(int32_t Wn) {
    char buf[64];
    snprintf(buf, sizeof buf, "%g", Wn / 100.0);
    return buf;
}

// Writes what the reference binary would have written for this fold and returns its exit code.
inline int emit_result(const std::string &seq, int n, int32_t Wn, const int32_t *pair, const int32_t *status,
                       FILE *out, FILE *err) {
    for (int x = 0; x < status[1]; ++x) fputs("Should not be here!\n", out);
    if (status[0] == CCJ_EXIT_FAILURE) {
        fprintf(err, "%s\n", traceback_message(status[2]).c_str());
        return 1;
    }
    if (status[0] == CCJ_EXIT_ZERO_NOT_GOOD) {
        fprintf(err, "NOT GOOD RESTR INTER, i=%d, j=%d, best_ip=%d, best_jp=%d\n", status[3], status[4], status[4],
                status[3]);
        return 0;
    }
    if (status[0] != CCJ_OK) {
        fprintf(err, "ccj_b200: internal traceback error %d\n", status[0]);
        return 70;
    }
    fprintf(out, "%s\n%s (%s)\n", seq.c_str(), fill_structure(n, pair).c_str(), energy_text(Wn).c_str());
    return 0;
}

}  // namespace ccj
