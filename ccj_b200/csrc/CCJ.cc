// CCJ command line -- drop-in for the reference driver (src/CCJ.cc:16-115): same options, same sequence
// handling (stdin/argv, toupper, T->U unless --noConv, validation), same parameter-file selection, same
// output "SEQ\nSTRUCT (ENERGY)\n".  The fold itself runs on the GPU through W_final.
#include <sys/stat.h>

#include <algorithm>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "W_final.hh"
#include "ccj_render.hpp"
#include "cmdline.hh"
#include "h_globals.hh"

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

static std::string fold_one(const std::string &seq, double &energy, int dangle) {  // ccj() of src/CCJ.cc:44-49
    W_final min_fold(seq, dangle);
    energy = min_fold.ccj();
    return min_fold.structure;
}

// --batch-file (ccj_b200 extension; the reference folds one sequence per process): every record of a FASTA file, or
// of a file with one sequence per line, goes through the same conversion and validation as a command-line
// sequence and all valid records are folded as ONE GPU batch (ccj_fold_batch).  Per record the output is the
// header line, if any, followed by exactly what the reference prints for that sequence; a record the reference
// would have rejected or aborted on prints that message instead and makes the exit status 1.
static int fold_batch_file(const args_info &args) {
    std::ifstream in(args.batch_file_arg);
    if (!in) {
        std::cerr << "Cannot read batch file " << args.batch_file_arg << std::endl;
        return EXIT_FAILURE;
    }
    struct Record { std::string header, seq; bool valid; };
    std::vector<Record> recs;
    std::string line;
    bool fasta = false;
    while (std::getline(in, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
        if (line.empty()) continue;
        if (line[0] == '>') {
            fasta = true;
            recs.push_back({line, "", true});
        } else if (fasta && !recs.empty()) {
            recs.back().seq += line;   // FASTA sequences may span lines
        } else {
            recs.push_back({"", line, true});
        }
    }
    noGU = args.noGU_given;
    bool dna = false;
    for (Record &r : recs) {
        std::transform(r.seq.begin(), r.seq.end(), r.seq.begin(), ::toupper);
        if (!args.noConv_flag)
            for (char &c : r.seq)
                if (c == 'T') c = 'U';
        for (char c : r.seq)
            if (!(c == 'G' || c == 'C' || c == 'A' || c == 'U' || c == 'T')) r.valid = false;
        if (r.seq.empty()) r.valid = false;
        if (r.valid && r.seq.find('T') != std::string::npos) dna = true;
    }
    std::string file;
    if (args.paramFile_given) file = args.paramFile_arg;
    else if (dna) { noGU = 1; file = "@dna_mathews2004"; }   // vrna_params_load_DNA_Mathews2004 (src/CCJ.cc:88-90)
    else file = "params/rna_DirksPierce09.par";
    ccj_ctx *ctx = nullptr;
    const char *dev = getenv("CCJ_DEVICE");
    if (ccj_ctx_create(dev ? atoi(dev) : 0, &ctx) != 0) {
        std::cerr << "ccj_b200: no CUDA device available (this build has no CPU path)" << std::endl;
        return EXIT_FAILURE;
    }
    const int mrc = file[0] == '@' ? ccj_model_load_embedded(ctx, file.c_str() + 1, args.dangles_arg, noGU)
                                   : (exists(file) ? ccj_model_load(ctx, file.c_str(), args.dangles_arg, noGU) : -1);
    if (mrc != 0) {
        std::cerr << "Not a valid parameter file!" << std::endl;
        return EXIT_FAILURE;
    }
    std::string all;
    std::vector<int64_t> offs(1, 0);
    std::vector<size_t> which;
    for (size_t x = 0; x < recs.size(); ++x)
        if (recs[x].valid) {
            all += recs[x].seq;
            offs.push_back((int64_t)all.size());
            which.push_back(x);
        }
    std::vector<ccj_result> res(which.size());
    std::vector<int32_t> pairs(all.size() + 1, -1);
    std::string dots(all.size() + 1, '.');
    if (!which.empty() && ccj_fold_batch(ctx, all.data(), offs.data(), (int)which.size(), res.data(), pairs.data(), &dots[0]) != 0) {
        std::cerr << "ccj_b200: " << ccj_last_error(ctx) << std::endl;
        return EXIT_FAILURE;
    }
    int rc = 0;
    size_t w = 0;
    for (size_t x = 0; x < recs.size(); ++x) {
        const Record &r = recs[x];
        if (!r.header.empty()) std::cout << r.header << std::endl;
        if (!r.valid) {
            if (r.seq.empty()) std::cout << "sequence is missing" << std::endl;
            else
                for (char c : r.seq)
                    if (!(c == 'G' || c == 'C' || c == 'A' || c == 'U' || c == 'T')) {
                        std::cout << "Sequence contains character " << c << " that is not G,C,A,U, or T." << std::endl;
                        break;
                    }
            rc = EXIT_FAILURE;
            continue;
        }
        const ccj_result &q = res[w];
        const int n = (int)r.seq.size();
        const int32_t status[5] = {q.status, q.n_should_not_be_here, q.msg_id, q.aux_i, q.aux_j};
        std::cout.flush();
        // pair[] of sequence w starts at its offset and is 0-based there; emit_result wants pair[1..n]
        std::vector<int32_t> pr(n + 2, -1);
        for (int y = 0; y < n; ++y) pr[y + 1] = pairs[offs[w] + y];
        if (ccj::emit_result(r.seq, n, q.energy_dcal, pr.data(), status, stdout, stderr) != 0) rc = EXIT_FAILURE;
        fflush(stdout);
        ++w;
    }
    ccj_ctx_destroy(ctx);
    return rc;
}

int main(int argc, char *argv[]) {
    args_info args;
    if (cmdline_parser(argc, argv, &args) != 0) exit(1);
    if (args.batch_file_given) {
        const int rc = fold_batch_file(args);
        cmdline_parser_free(&args);
        return rc;
    }

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
        // the reference switches to its embedded Mathews-2004 DNA set here (src/CCJ.cc:88-90); the same set is linked
        // into libccj_b200.so, so this works from any directory like the reference does
        noGU = 1;
        ccj_params_load_DNA_Mathews2004();
    } else {
        file = "params/rna_DirksPierce09.par";  // cwd-relative, as in the reference (src/CCJ.cc:92)
    }
    if (!file.empty() && (!exists(file) || !ccj_params_load(file.c_str()))) {
        std::cerr << "Not a valid parameter file!" << std::endl;
        exit(EXIT_FAILURE);
    }

    double energy;
    std::string structure = fold_one(seq, energy, args.dangles_arg);
    std::cout << seq << std::endl;
    std::cout << structure << " (" << energy << ")" << std::endl;
    cmdline_parser_free(&args);
    return 0;
}
