// Command-line options of CCJ -- hand-written getopt_long parser with the option set, help/version text,
// error messages and exit codes of the reference's gengetopt-generated parser (src/ccj.ggo:1-33,
// src/cmdline.cc; gengetopt itself is not available in this image).
#ifndef CCJ_B200_CMDLINE_HH
#define CCJ_B200_CMDLINE_HH

#define CMDLINE_PARSER_PACKAGE "CCJ"
#define CMDLINE_PARSER_VERSION "1.0"

struct args_info {
    char *input_file_arg;       // -i, --input-file (parsed, never read: src/CCJ.cc:71)
    int dangles_arg;            // -d, --dangles (default 2)
    char *paramFile_arg;        // -P, --paramFile
    int noConv_flag;            // --noConv
    int noGU_flag;              // --noGU
    char *batch_file_arg;       // --batch-file (ccj_b200 extension, SURVEY.md 8f: FASTA or one sequence per line)
    unsigned int help_given, version_given, input_file_given, dangles_given, paramFile_given, noConv_given,
        noGU_given, batch_file_given;
    char **inputs;              // unnamed options: [sequence]
    unsigned inputs_num;
};

// 0 on success; prints the reference's diagnostics and exits with its codes otherwise
int cmdline_parser(int argc, char **argv, struct args_info *args_info);
void cmdline_parser_print_help(void);
void cmdline_parser_print_version(void);
void cmdline_parser_free(struct args_info *args_info);

#endif
