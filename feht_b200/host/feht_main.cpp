// feht_main.cpp -- `feht-b200`: the feht command line (src/UserInput.hs:41-94 under /root/reference) on the
// CUDA hot path.  Same options, same defaults, same output; see include/feht_host.h.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <string>

#include "../../include/feht_cuda.h"
#include "../../include/feht_host.h"

namespace {

const char *kUsage =
    "feht - predictive marker discovery\n\n"
    "Usage: feht-b200 (-i|--infoFile FILE) (-d|--datafile FILE)\n"
    "            [--one \"Group1Name Group1Item Group1Item Group1Item\"]\n"
    "            [--two \"Group2Name Group2Item Group2Item Group2Item\"]\n"
    "            [-l|--delimiter [',', '\\t' ...], DEFAULT='\\t']\n"
    "            [-m|--mode ['binary', 'snp'], DEFAULT='binary']\n"
    "            [-c|--correction ['none', 'bonferroni'], DEFAULT='bonferroni']\n"
    "            [-f|--ratioFilter [Filter results by ratio (0.00-1.00), DEFAULT=0]]\n"
    "            [--device N] [--engine auto|popc|gemm]\n"
    "  Predictive marker discovery for groups; binary data, genomic data (single\n"
    "  nucleotide variants), and arbitrary character data.\n";

bool read_file(const std::string &path, std::string *out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    out->assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    return true;
}

// optparse `auto` for Char reads Haskell character syntax: '\t', ',', ... (UserInput.hs:63-68)
bool parse_delimiter(const std::string &s, char *d) {
    if (s.size() == 3 && s.front() == '\'' && s.back() == '\'') { *d = s[1]; return true; }
    if (s == "'\\t'") { *d = '\t'; return true; }
    if (s.size() == 1) { *d = s[0]; return true; }   // convenience: a bare character
    if (s == "\\t") { *d = '\t'; return true; }
    return false;
}

}  // namespace

int main(int argc, char **argv) {
    std::string info, data, one = "all all", two = "all all", mode = "binary", corr = "bonferroni", engine = "auto";
    char delim = '\t';
    double ratio_filter = 0.0;
    int device = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto need = [&](const char *what) -> const char * {
            if (i + 1 >= argc) { std::fprintf(stderr, "The option `%s` expects an argument.\n\n%s", what, kUsage); std::exit(1); }
            return argv[++i];
        };
        if (a == "-i" || a == "--infoFile") info = need("-i");
        else if (a == "-d" || a == "--datafile") data = need("-d");
        else if (a == "--one") one = need("--one");
        else if (a == "--two") two = need("--two");
        else if (a == "-l" || a == "--delimiter") {
            if (!parse_delimiter(need("-l"), &delim)) { std::fprintf(stderr, "option -l: cannot parse value\n\n%s", kUsage); return 1; }
        } else if (a == "-m" || a == "--mode") mode = need("-m");
        else if (a == "-c" || a == "--correction") corr = need("-c");
        else if (a == "-f" || a == "--ratioFilter") ratio_filter = std::atof(need("-f"));
        else if (a == "--device") device = std::atoi(need("--device"));
        else if (a == "--engine") engine = need("--engine");
        else if (a == "-h" || a == "--help") { std::fputs(kUsage, stdout); return 0; }
        else { std::fprintf(stderr, "Invalid option `%s'\n\n%s", a.c_str(), kUsage); return 1; }
    }
    if (info.empty() || data.empty()) { std::fprintf(stderr, "Missing: (-i|--infoFile FILE) (-d|--datafile FILE)\n\n%s", kUsage); return 1; }
    if (mode != "binary" && mode != "snp") { std::fprintf(stderr, "option -m: Incorrect data mode type.\n\n%s", kUsage); return 1; }
    if (corr != "none" && corr != "bonferroni") { std::fprintf(stderr, "option -c: Incorrect multiple testing correction supplied.\n\n%s", kUsage); return 1; }
    const int eng = engine == "popc" ? FEHT_ENGINE_POPC : engine == "gemm" ? FEHT_ENGINE_GEMM : FEHT_ENGINE_AUTO;
    std::string meta_bytes, data_bytes;
    if (!read_file(info, &meta_bytes)) { std::fprintf(stderr, "feht: %s: openBinaryFile: does not exist (No such file or directory)\n", info.c_str()); return 1; }
    if (!read_file(data, &data_bytes)) { std::fprintf(stderr, "feht: %s: openBinaryFile: does not exist (No such file or directory)\n", data.c_str()); return 1; }
    char *out = nullptr;
    int64_t len = 0;
    char err[1024] = {0};
    const int rc = feht_host_run(meta_bytes.data(), int64_t(meta_bytes.size()), data_bytes.data(),
                                 int64_t(data_bytes.size()), one.c_str(), two.c_str(), delim, mode == "snp",
                                 corr == "bonferroni" ? FEHT_CORRECTION_BONFERRONI : FEHT_CORRECTION_NONE,
                                 ratio_filter, device, eng, &out, &len, err, int(sizeof(err)));
    if (rc != 0) { std::fprintf(stderr, "feht: %s\n", err); return 1; }
    std::fwrite(out, 1, size_t(len), stdout);
    feht_host_free(out);
    return 0;
}
