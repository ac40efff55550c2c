// TEST TOOL: ccj_shard_cell_of (ccj_types.h) is the exact inverse of the row offsets of a rank's part of a slab, for every
// cell of every slab size up to 700 rows and 1, 2, 3, 4, 8 ranks (the flat thread->cell mapping of k_4d / k_4d_shard).
#include <initializer_list>
#include <cstdio>
#include "ccj_types.h"
int main() {
    long bad = 0, tot = 0;
    for (int G : {1, 2, 3, 4, 8})
        for (int mr = 1; mr <= 700; mr += (mr < 40 ? 1 : 7)) {
            const int Q = (mr + G - 1) / G;
            int p = 0;
            for (int q = 0; q < Q; ++q)
                for (int kk = 0; kk < mr - q * G; ++kk, ++p) {
                    int q2, k2;
                    ccj_shard_cell_of(p, mr, G, Q, q2, k2);
                    ++tot;
                    if (q2 != q || k2 != kk) { if (bad++ < 5) printf("BAD G=%d mr=%d p=%d want (%d,%d) got (%d,%d)\n", G, mr, p, q, kk, q2, k2); }
                }
            if (p != Q * mr - G * (Q * (Q - 1) / 2)) { printf("size mismatch\n"); ++bad; }
        }
    printf("checked %ld cells, %ld bad\n", tot, bad);
    return bad != 0;
}
