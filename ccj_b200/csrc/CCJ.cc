// CCJ command line -- drop-in for the reference driver (src/CCJ.cc:16-115): same options, same sequence
// handling (stdin/argv, toupper, T->U unless --noConv, validation), same parameter-file selection, same
// output "SEQ\nSTRUCT (ENERGY)\n".  The fold itself runs on the GPU through W_final.
#include <sys/stat.h>

#include <algorithm>
#include <iostream>
#include <string>

#include "W_final.hh"
#include "cmdline.hh"

static bool exists(const std::string &path) {
    struct stat buffer;
    return stat(path.c_str(), &buffer) == 0;
}

static void validateSequence(const std::string &sequence) {
    if (sequence.length() == 0) {
        std::cout << "sequence is missing" << std::endl;
        exit(EXIT_FAILURE);
    }
    for (char c : sequence)
        if (!(c == 'G' || c == 'C' || c == 'A' || c == 'U' || c == 'T')) {
            std::cout << "Sequence contains character " << c << " that is not G,C,A,U, or T." << std::endl;
            exit(EXIT_FAILURE);
        }
}

static std::string ccj(const std::string &seq, double &energy, int dangle) {
    W_final min_fold(seq, dangle);
    energy = min_fold.ccj();
    return min_fold.structure;
}

int main(int argc, char *argv[]) {
    args_info args;
    if (cmdline_parser(argc, argv, &args) != 0) exit(1);

    std::string seq;
    if (args.inputs_num > 0) seq = args.inputs[0];
    else if (!args.input_file_given) std::getline(std::cin, seq);

    std::transform(seq.begin(), seq.end(), seq.begin(), ::toupper);
    if (!args.noConv_flag)
        for (char &c : seq)
            if (c == 'T') c = 'U';

    noGU = args.noGU_given;
    validateSequence(seq);

    std::string file;
    if (args.paramFile_given) {
        file = args.paramFile_arg;
    } else if (seq.find('T') != std::string::npos) {
        // the reference switches to its embedded Mathews-2004 DNA set here (vrna_params_load_DNA_Mathews2004);
        // that hex-encoded set is not bundled -- the equivalent file is params/dna_Matthews04.par
        noGU = 1;
        file = "params/dna_Matthews04.par";
    } else {
        file = "params/rna_DirksPierce09.par";  // cwd-relative, as in the reference (src/CCJ.cc:92)
    }
    if (!exists(file) || !ccj_params_load(file.c_str())) {
        std::cerr << "Not a valid parameter file!" << std::endl;
        exit(EXIT_FAILURE);
    }

    double energy;
    std::string structure = ccj(seq, energy, args.dangles_arg);
    std::cout << seq << std::endl;
    std::cout << structure << " (" << energy << ")" << std::endl;
    cmdline_parser_free(&args);
    return 0;
}
