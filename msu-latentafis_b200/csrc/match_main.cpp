// Drop-in command line of the reference matcher (matching/main.cpp:35-87):
//   match [-c codebook.dat] [-s scoredir/] [-g gallerydir] (-l latent.dat | -ldir latentdir) [-d device] [-gpus N]
// Flags take precedence over ../afis.config (relative to the working directory).  Unlike the
// reference the config file is optional when every needed flag is given.
// -gpus N (or LAFIS_GPUS=N) shards the gallery over devices 0..N-1 of this process (contiguous index ranges,
// one NCCL all-gather per match, score rows gathered to device 0); the files written are the same, byte for byte.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "latentafis_b200.h"

namespace fs = std::filesystem;

namespace {

// matching/argparser.h:4-25: exact-token lookup, the value is the next token
struct ArgParser {
    std::vector<std::string> tokens;
    ArgParser(int argc, char** argv) {
        for (int i = 1; i < argc; ++i) tokens.push_back(argv[i]);
    }
    bool exists(const std::string& o) const { return std::find(tokens.begin(), tokens.end(), o) != tokens.end(); }
    std::string get(const std::string& o) const {
        auto it = std::find(tokens.begin(), tokens.end(), o);
        if (it != tokens.end() && ++it != tokens.end()) return *it;
        return "";
    }
};

// afis.config is a flat JSON object of string values (afis.config:1-17); a key lookup suffices
std::string config_value(const std::string& text, const std::string& key) {
    const std::string pat = "\"" + key + "\"";
    size_t p = text.find(pat);
    if (p == std::string::npos) return "";
    p = text.find(':', p + pat.size());
    if (p == std::string::npos) return "";
    p = text.find('"', p);
    if (p == std::string::npos) return "";
    std::string out;
    for (size_t i = p + 1; i < text.size() && text[i] != '"'; ++i) {
        if (text[i] == '\\' && i + 1 < text.size()) ++i;
        out.push_back(text[i]);
    }
    return out;
}

}  // namespace

int main(int argc, char** argv) {
    // LAFIS_INGEST_TIMING: where a cold process spends its time (stderr, next to the ingest line of the library)
    const bool timing = std::getenv("LAFIS_INGEST_TIMING") != nullptr;
    const auto t_main = std::chrono::steady_clock::now();
    auto stamp = [&](const char* what) {
        if (timing)
            std::cerr << "lafis cli: " << what << " at "
                      << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_main).count() << " ms"
                      << std::endl;
    };
    std::string config;
    {
        std::ifstream in((fs::current_path().parent_path() / "afis.config").string());
        std::stringstream ss;
        ss << in.rdbuf();
        config = ss.str();
    }
    ArgParser args(argc, argv);
    const std::string codebook = args.exists("-c") ? args.get("-c") : config_value(config, "CodebookPath");
    const int device = args.exists("-d") ? std::atoi(args.get("-d").c_str()) : 0;

    int n_gpus = 1;
    if (args.exists("-gpus")) n_gpus = std::atoi(args.get("-gpus").c_str());
    else if (const char* e = std::getenv("LAFIS_GPUS")) n_gpus = std::atoi(e);
    lafis_ctx* ctx = nullptr;
    lafis_group* group = nullptr;
    int rc = n_gpus > 1 ? lafis_group_create(codebook.c_str(), nullptr, n_gpus, &group) : lafis_create(codebook.c_str(), device, &ctx);
    stamp("matcher created (CUDA context, module, codebook)");
    if (rc != LAFIS_OK) {
        std::cerr << "match: cannot create matcher (status " << rc << "): " << lafis_last_error(nullptr) << "; codebook '"
                  << codebook << "', CUDA device " << device << ", " << n_gpus << " GPU(s) (sm_100 GPUs are required)"
                  << std::endl;
        return 2;
    }
    std::string score_path;
    if (args.exists("-s")) score_path = args.get("-s");
    else {
        std::cout << "Missing argument for score directory. Using default from afis.config" << std::endl;
        score_path = config_value(config, "ScorePath");
    }
    std::error_code ec;
    fs::create_directory(fs::path(score_path), ec);
    std::string gallery_path;
    if (args.exists("-g")) gallery_path = args.get("-g");
    else {
        std::cout << "Missing argument for gallery directory. Using default from afis.config" << std::endl;
        gallery_path = config_value(config, "GalleryTemplateDirectory");
    }
    if (args.exists("-l")) {
        rc = group ? lafis_group_one2list_matching(group, args.get("-l").c_str(), gallery_path.c_str(), score_path.c_str())
                   : lafis_one2list_matching(ctx, args.get("-l").c_str(), gallery_path.c_str(), score_path.c_str());
    } else if (args.exists("-ldir")) {
        rc = group ? lafis_group_list2list_matching(group, args.get("-ldir").c_str(), gallery_path.c_str(), score_path.c_str())
                   : lafis_list2list_matching(ctx, args.get("-ldir").c_str(), gallery_path.c_str(), score_path.c_str());
    } else {
        std::cout << "Missing argument for latent template or directory. Assuming batch matching, using default "
                     "directory from afis.config"
                  << std::endl;
        const std::string ldir = config_value(config, "LatentTemplateDirectory");
        rc = group ? lafis_group_list2list_matching(group, ldir.c_str(), gallery_path.c_str(), score_path.c_str())
                   : lafis_list2list_matching(ctx, ldir.c_str(), gallery_path.c_str(), score_path.c_str());
    }
    stamp("driver returned (ingest, match, files written)");
    if (rc < LAFIS_ERR_NO_TEMPLATES)
        std::cerr << "match: " << (group ? lafis_group_last_error(group) : lafis_last_error(ctx)) << " (status " << rc << ")"
                  << std::endl;
    if (group) lafis_group_destroy(group);
    else lafis_destroy(ctx);
    stamp("matcher destroyed");
    return 0;  // matching/main.cpp:86 ignores the drivers' return codes
}
