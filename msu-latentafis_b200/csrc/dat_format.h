// Readers for the matcher's on-disk formats: template .dat files and the PQ codebook.
// Behavioural mirror of PQ::Matcher::load_FP_template (matching/matcher.cpp:785-884 latent,
// :886-983 rolled) and of the codebook read in PQ::Matcher::Matcher (:58-93).  Only what the
// matcher later reads is retained: for a rolled print the first non-empty minutiae
// template and the first non-empty texture template (matcher.cpp:406, :413 use index 0 only); for a
// latent print the non-empty minutiae templates at positions 26, 2, 11 (:380) and texture
// template 0 (:411).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace lafis {

constexpr int kDesLen = 96;        // descriptor length the matcher is built for
constexpr int kSubs = 16;          // PQ sub-quantizers
constexpr int kClusters = 256;     // centroids per sub-quantizer
constexpr int kSubDim = 6;
constexpr int kMaxMinutiae = 2000; // matcher.cpp:788 / :889 Max_Nrof_Minutiae
constexpr int kMaxTexture = 1000;  // matcher.h:31-32, applied at matcher.cpp:544-547
constexpr int kSelected[3] = {26, 2, 11};  // matcher.cpp:380

struct PointSet {
    std::vector<int16_t> x, y;
    std::vector<float> ori;
    std::vector<float> des;     // [n][96] (minutiae, latent texture)
    std::vector<uint8_t> codes; // [n][16] (rolled texture)
    int n() const { return (int)x.size(); }
};

struct RolledTemplate {
    int status = 0;  // LAFIS_TPL_*
    int n_minu_templates = 0, n_tex_templates = 0;
    PointSet minu, tex;
};

struct LatentTemplate {
    int load_rc = 0;  // what the reference loader would have returned
    int n_minu_templates = 0, n_tex_templates = 0;
    PointSet minu[3];  // slots for template positions 26, 2, 11
    PointSet tex;
};

// Returns the reference loader's return code (0, 1, 2, 4, -1) and fills `out`; -3 on I/O error.
int read_rolled_dat(const std::string& path, RolledTemplate& out);
int read_latent_dat(const std::string& path, LatentTemplate& out);

// Writes a rolled template in the layout of Template2Bin_Byte_PQ_rolled (extraction/descriptor_PQ.py:178-272):
// version-1 header, h, w, block sizes clamped to 50, one minutiae template (x, y as u16, ori f32, des f32[n][96]) and
// one texture template (block coordinates u16, ori f32, 16 PQ codes per point).  At most 2000 points per template
// are written (:213-214, :245-246).  Returns 0, or -3 on I/O error.
int write_rolled_dat(const std::string& path, int h, int w, int blkH, int blkW, const PointSet& minu, const PointSet& tex);

// Writes a latent template in the layout of Template2Bin_Byte_latent (extraction/descriptor_PQ.py:80-175): version-1
// header, h, w, block sizes clamped to 50, `minu.size()` minutiae templates (an empty one is just its zero count,
// :113-117) and `tex.size()` texture templates with block coordinates and f32 descriptors [n][96].  At most 2000
// points per template are written (:111-112, :143-144).  Returns 0, or -3 on I/O error.
int write_latent_dat(const std::string& path, int h, int w, int blkH, int blkW, const std::vector<PointSet>& minu,
                     const std::vector<PointSet>& tex);

// u16 subs, u16 clusters, u16 sub_dim, f32[subs][clusters][sub_dim]; returns 0 or a negative LAFIS_ERR_*
int read_codebook(const std::string& path, std::vector<float>& codewords, int& subs, int& clusters, int& sub_dim);

// A path as the reference prints it into its score files: boost::filesystem's stream inserter
// (matcher.cpp:202 `output << rolled_template_files[j]`), i.e. boost::io::quoted with '&' as the escape character -
// double quotes around the string, '&' in front of every '"' and '&'.
std::string quoted_path(const std::string& p);
// One row of the N-vs-N score file (matcher.cpp:198-205): <quoted path>,<score, fixed, 3 decimals>\n - what
// `output << path << "," << std::setprecision(3) << std::fixed << score << std::endl` writes, without the flush.
void append_score_row(std::string& out, const std::string& quoted, float score);

// *.dat entries of a directory in directory-iteration order (matcher.cpp:103-110, :122-130, :227-234)
std::vector<std::string> list_dat_files(const std::string& dir);

}  // namespace lafis
