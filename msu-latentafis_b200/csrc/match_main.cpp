// Drop-in command line of the reference matcher (matching/main.cpp:35-87):
//   match [-c codebook.dat] [-s scoredir/] [-g gallerydir] (-l latent.dat | -ldir latentdir) [-d device]
// Flags take precedence over ../afis.config (relative to the working directory).  Unlike the
// reference the config file is optional when every needed flag is given.
#include <algorithm>
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

    lafis_ctx* ctx = nullptr;
    int rc = lafis_create(codebook.c_str(), device, &ctx);
    if (rc != LAFIS_OK) {
        std::cerr << "match: cannot create matcher (status " << rc << "); codebook '" << codebook
                  << "', CUDA device " << device << " (an sm_100 GPU is required)" << std::endl;
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
        rc = lafis_one2list_matching(ctx, args.get("-l").c_str(), gallery_path.c_str(), score_path.c_str());
    } else if (args.exists("-ldir")) {
        rc = lafis_list2list_matching(ctx, args.get("-ldir").c_str(), gallery_path.c_str(), score_path.c_str());
    } else {
        std::cout << "Missing argument for latent template or directory. Assuming batch matching, using default "
                     "directory from afis.config"
                  << std::endl;
        rc = lafis_list2list_matching(ctx, config_value(config, "LatentTemplateDirectory").c_str(), gallery_path.c_str(),
                                      score_path.c_str());
    }
    if (rc < LAFIS_ERR_NO_TEMPLATES) std::cerr << "match: " << lafis_last_error(ctx) << " (status " << rc << ")" << std::endl;
    lafis_destroy(ctx);
    return 0;  // matching/main.cpp:86 ignores the drivers' return codes
}
