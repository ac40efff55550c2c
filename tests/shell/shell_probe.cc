// TEST TOOL: the probe of tests/shell/probe_body.inc against the B200 class shells (needs a GPU).
//   shell_probe <parfile|-> <dangles> <seq>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "W_final.hh"
#include "h_globals.hh"

#include "probe_body.inc"

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    if (strcmp(argv[1], "-") != 0 && !vrna_params_load(argv[1], VRNA_PARAMETER_FORMAT_DEFAULT)) {
        fprintf(stderr, "Not a valid parameter file!\n");
        return 1;
    }
    run_probe(argv[3], atoi(argv[2]));
    return 0;
}
