// See cmdline.hh.  Behaviour checked against the reference binary by tests/test_cli.py.
#include "cmdline.hh"

#include <getopt.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

void cmdline_parser_print_version(void) { printf("%s %s\n", CMDLINE_PARSER_PACKAGE, CMDLINE_PARSER_VERSION); }

void cmdline_parser_print_help(void) {
    fputs("Usage: CCJ [options] [sequence]\n"
          "Pseudoknotted minimum free energy folding of RNAs\n"
          "\n"
          "Read RNA sequence from stdin or cmdline; predict minimum\n"
          "free energy and optimum structure\n"
          "\n"
          "  -h, --help               Print help and exit\n"
          "  -V, --version            Print version and exit\n"
          "  -i, --input-file=STRING  Give a path to an input file containing the sequence\n"
          "                             (and input structure if known)\n"
          "  -d, --dangles=INT        Specify the dangle model to be used (base is 2)\n"
          "                             (default=`2')\n"
          "  -P, --paramFile=STRING   Read energy parameters from paramfile, instead of\n"
          "                             using the default parameter set.\n"
          "      --noConv             Do not convert DNA into RNA. This will use the\n"
          "                             Matthews 2004 parameters for DNA  (default=off)\n"
          "      --noGU               Turn off G-U and U-G (and G-T and T-G) base pairing\n"
          "                             (default=off)\n"
          "\n"
          "The input sequence is read from standard input, unless it is\n"
          "given on the command line.\n"
          "\n",
          stdout);
}

void cmdline_parser_free(struct args_info *a) {
    free(a->input_file_arg);
    free(a->paramFile_arg);
    free(a->batch_file_arg);
    for (unsigned i = 0; i < a->inputs_num; ++i) free(a->inputs[i]);
    free(a->inputs);
    memset(a, 0, sizeof *a);
}

static bool twice(const char *prog, unsigned seen, const char *lng, char sht) {
    if (!seen) return false;
    if (sht != '-') fprintf(stderr, "%s: `--%s' (`-%c') option given more than once\n", prog, lng, sht);
    else fprintf(stderr, "%s: `--%s' option given more than once\n", prog, lng);
    return true;
}

int cmdline_parser(int argc, char **argv, struct args_info *a) {
    memset(a, 0, sizeof *a);
    a->dangles_arg = 2;
    static struct option longopts[] = {{"help", 0, nullptr, 'h'},        {"version", 0, nullptr, 'V'},
                                       {"input-file", 1, nullptr, 'i'},  {"dangles", 1, nullptr, 'd'},
                                       {"paramFile", 1, nullptr, 'P'},   {"noConv", 0, nullptr, 0},
                                       {"noGU", 0, nullptr, 0},          {"batch-file", 1, nullptr, 0},
                                       {nullptr, 0, nullptr, 0}};
    const char *prog = argv[0];
    optarg = nullptr;
    optind = 0;
    opterr = 1;
    bool fail = false;
    for (;;) {
        int idx = 0;
        const int c = getopt_long(argc, argv, "hVi:d:P:", longopts, &idx);
        if (c == -1) break;
        switch (c) {
            case 'h': cmdline_parser_print_help(); exit(EXIT_SUCCESS);
            case 'V': cmdline_parser_print_version(); exit(EXIT_SUCCESS);
            case 'i':
                if (twice(prog, a->input_file_given, "input-file", 'i')) { fail = true; break; }
                a->input_file_given = 1;
                a->input_file_arg = strdup(optarg);
                break;
            case 'd': {
                if (twice(prog, a->dangles_given, "dangles", 'd')) { fail = true; break; }
                a->dangles_given = 1;
                char *stop = nullptr;
                a->dangles_arg = (int)strtol(optarg, &stop, 0);
                if (!(stop && *stop == '\0')) {
                    fprintf(stderr, "%s: invalid numeric value: %s\n", prog, optarg);
                    fail = true;
                }
            } break;
            case 'P':
                if (twice(prog, a->paramFile_given, "paramFile", 'P')) { fail = true; break; }
                a->paramFile_given = 1;
                a->paramFile_arg = strdup(optarg);
                break;
            case 0:
                if (strcmp(longopts[idx].name, "noConv") == 0) {
                    if (twice(prog, a->noConv_given, "noConv", '-')) { fail = true; break; }
                    a->noConv_given = 1;
                    a->noConv_flag = !a->noConv_flag;
                } else if (strcmp(longopts[idx].name, "noGU") == 0) {
                    if (twice(prog, a->noGU_given, "noGU", '-')) { fail = true; break; }
                    a->noGU_given = 1;
                    a->noGU_flag = !a->noGU_flag;
                } else if (strcmp(longopts[idx].name, "batch-file") == 0) {
                    if (twice(prog, a->batch_file_given, "batch-file", '-')) { fail = true; break; }
                    a->batch_file_given = 1;
                    a->batch_file_arg = strdup(optarg);
                }
                break;
            default:  // '?': getopt_long printed the message
                fail = true;
        }
        if (fail) break;
    }
    if (fail) {
        cmdline_parser_free(a);
        exit(EXIT_FAILURE);
    }
    if (optind < argc) {
        a->inputs_num = argc - optind;
        a->inputs = (char **)malloc(sizeof(char *) * a->inputs_num);
        for (unsigned i = 0; i < a->inputs_num; ++i) a->inputs[i] = strdup(argv[optind + i]);
    }
    return 0;
}
