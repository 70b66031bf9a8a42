// 1-vs-N and N-vs-N drivers with the reference's score-file formats, on top of the C ABI.
//
//   reference: PQ::Matcher::One2List_matching   matching/matcher.cpp:216-337
//              PQ::Matcher::List2List_matching  matching/matcher.cpp:96-214
//
// Differences that are deliberate: the gallery is parsed ONCE into HBM instead of once per
// (latent, rolled) pair (matcher.cpp:173, :278), and all latents of a directory are scored as
// batches.  File names, CSV layouts, return codes and console messages follow the reference.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <numeric>
#include <string>
#include <vector>

#include "dat_format.h"
#include "lafis_internal.h"

namespace fs = std::filesystem;
using namespace lafis;

namespace {

// What the drivers need from "a matcher": one context, or a group of contexts over a sharded gallery.
struct Engine {
    virtual ~Engine() {}
    virtual int load_files(const std::vector<const char*>& ptrs) = 0;
    virtual bool has_dir(const std::string& dir) const = 0;  // the resident gallery was loaded from `dir`
    virtual void set_dir(const std::string& dir) = 0;
    virtual int gallery_size() const = 0;
    virtual const char* gallery_path(int j) const = 0;
    virtual int gallery_status(int j) const = 0;
    virtual int gallery_minutiae(int j) = 0;                 // minutiae of template j (0: no minutiae template)
    virtual lafis_ctx* latents_ctx() = 0;
    virtual int match_scores(lafis_latents* L, float* scores) = 0;
    virtual int correspondences(lafis_latents* L, int q, int j, int16_t* xy, int* counts) = 0;
    virtual const char* last_error() const = 0;
};

struct SingleEngine : Engine {
    lafis_ctx* c;
    explicit SingleEngine(lafis_ctx* ctx) : c(ctx) {}
    int load_files(const std::vector<const char*>& p) override { return lafis_gallery_load_files(c, p.data(), (int)p.size(), 0, 1); }
    // the association lives in the context and is dropped by everything that replaces the gallery (free_gallery)
    bool has_dir(const std::string& dir) const override {
        return c->gallery_set && c->gallery_dir == dir && c->gallery_dir_shards == 1 && c->gal.n > 0;
    }
    void set_dir(const std::string& dir) override {
        c->gallery_dir = dir;
        c->gallery_dir_shard = 0;
        c->gallery_dir_shards = 1;
    }
    int gallery_size() const override { return lafis_gallery_size(c); }
    const char* gallery_path(int j) const override { return lafis_gallery_path(c, j); }
    int gallery_status(int j) const override { return lafis_gallery_status(c, j); }
    int gallery_minutiae(int j) override { return (j >= 0 && j < (int)c->h_minu_n.size()) ? c->h_minu_n[j] : 0; }
    lafis_ctx* latents_ctx() override { return c; }
    int match_scores(lafis_latents* L, float* scores) override { return lafis_match(c, L, 0, nullptr, scores, nullptr); }
    int correspondences(lafis_latents* L, int q, int j, int16_t* xy, int* counts) override {
        return lafis_correspondences(c, L, q, j, xy, counts);
    }
    const char* last_error() const override { return lafis_last_error(c); }
};

struct GroupEngine : Engine {
    lafis_group* g;
    explicit GroupEngine(lafis_group* grp) : g(grp) {}
    int shard_of(int j) const {
        int r = 0;
        while (r + 1 < (int)g->ctx.size() && (uint32_t)j >= g->base[r + 1]) ++r;
        return r;
    }
    int load_files(const std::vector<const char*>& p) override { return lafis_group_gallery_load_files(g, p.data(), (int)p.size()); }
    bool has_dir(const std::string& dir) const override {
        const int w = (int)g->ctx.size();
        for (int i = 0; i < w; ++i) {
            const lafis_ctx* c = g->ctx[i];
            if (!c->gallery_set || c->gallery_dir != dir || c->gallery_dir_shard != i || c->gallery_dir_shards != w) return false;
        }
        return lafis_group_gallery_size(g) > 0;
    }
    void set_dir(const std::string& dir) override {
        for (size_t i = 0; i < g->ctx.size(); ++i) {
            g->ctx[i]->gallery_dir = dir;
            g->ctx[i]->gallery_dir_shard = (int)i;
            g->ctx[i]->gallery_dir_shards = (int)g->ctx.size();
        }
    }
    int gallery_size() const override { return lafis_group_gallery_size(g); }
    const char* gallery_path(int j) const override {
        const int r = shard_of(j);
        return lafis_gallery_path(g->ctx[r], j - (int)g->base[r]);
    }
    int gallery_status(int j) const override {
        const int r = shard_of(j);
        return lafis_gallery_status(g->ctx[r], j - (int)g->base[r]);
    }
    int gallery_minutiae(int j) override {
        const int r = shard_of(j), l = j - (int)g->base[r];
        const lafis_ctx* c = g->ctx[r];
        return (l >= 0 && l < (int)c->h_minu_n.size()) ? c->h_minu_n[l] : 0;
    }
    lafis_ctx* latents_ctx() override { return g->ctx[0]; }
    int match_scores(lafis_latents* L, float* scores) override { return lafis_group_match(g, L, 0, nullptr, scores); }
    int correspondences(lafis_latents* L, int q, int j, int16_t* xy, int* counts) override {
        const int r = shard_of(j);  // on the device that holds the template
        const int rc = lafis_correspondences(g->ctx[r], L, q, j - (int)g->base[r], xy, counts);
        if (rc < 0) g->err = lafis_last_error(g->ctx[r]);
        return rc;
    }
    const char* last_error() const override { return lafis_group_last_error(g); }
};

int ensure_gallery(Engine& E, const std::string& dir, bool announce) {
    if (E.has_dir(dir)) return LAFIS_OK;
    std::vector<std::string> files = list_dat_files(dir);
    if (announce)
        for (const std::string& f : files) std::cout << "rolled template file" << fs::path(f) << std::endl;
    if (files.empty()) {
        std::cout << "No rolled templates found in directory: " << dir << std::endl;
        return LAFIS_ERR_NO_TEMPLATES;
    }
    std::vector<const char*> ptrs(files.size());
    for (size_t i = 0; i < files.size(); ++i) ptrs[i] = files[i].c_str();
    const int rc = E.load_files(ptrs);
    if (rc == LAFIS_OK) E.set_dir(dir);
    return rc;
}

int one2list(Engine& E, const char* latent_template_file, const char* rolled_dir, const char* score_path);
int list2list(Engine& E, const char* latent_dir, const char* rolled_dir, const char* score_path);

}  // namespace

extern "C" {

LAFIS_API void lafis_forget_gallery_dir(lafis_ctx* ctx) {
    if (ctx) ctx->gallery_dir.clear();
}

LAFIS_API int lafis_one2list_matching(lafis_ctx* ctx, const char* latent_template_file, const char* rolled_dir,
                                      const char* score_path) {
    if (!ctx || !latent_template_file || !rolled_dir || !score_path) return LAFIS_ERR_ARG;
    SingleEngine E(ctx);
    return one2list(E, latent_template_file, rolled_dir, score_path);
}

LAFIS_API int lafis_list2list_matching(lafis_ctx* ctx, const char* latent_dir, const char* rolled_dir,
                                       const char* score_path) {
    if (!ctx || !latent_dir || !rolled_dir || !score_path) return LAFIS_ERR_ARG;
    SingleEngine E(ctx);
    return list2list(E, latent_dir, rolled_dir, score_path);
}

LAFIS_API int lafis_group_one2list_matching(lafis_group* group, const char* latent_template_file, const char* rolled_dir,
                                            const char* score_path) {
    if (!group || !latent_template_file || !rolled_dir || !score_path) return LAFIS_ERR_ARG;
    GroupEngine E(group);
    return one2list(E, latent_template_file, rolled_dir, score_path);
}

LAFIS_API int lafis_group_list2list_matching(lafis_group* group, const char* latent_dir, const char* rolled_dir,
                                             const char* score_path) {
    if (!group || !latent_dir || !rolled_dir || !score_path) return LAFIS_ERR_ARG;
    GroupEngine E(group);
    return list2list(E, latent_dir, rolled_dir, score_path);
}

LAFIS_API int lafis_enroll_rolled(lafis_ctx* ctx, const lafis_rolled_features* F, const char* out_path) {
    if (!ctx || !F || !out_path || F->n_minu < 0 || F->n_tex < 0 || (F->n_minu > 0 && (!F->minu_xyo || !F->minu_des)) ||
        (F->n_tex > 0 && (!F->tex_xyo || !F->tex_des)))
        return LAFIS_ERR_ARG;
    PointSet minu, tex;
    const int nm = std::min(F->n_minu, kMaxMinutiae), nt = std::min(F->n_tex, kMaxMinutiae);
    minu.x.resize(nm);
    minu.y.resize(nm);
    minu.ori.resize(nm);
    for (int i = 0; i < nm; ++i) {  // np.uint16(x): truncation (descriptor_PQ.py:221-229)
        minu.x[i] = (int16_t)(uint16_t)F->minu_xyo[3 * i];
        minu.y[i] = (int16_t)(uint16_t)F->minu_xyo[3 * i + 1];
        minu.ori[i] = F->minu_xyo[3 * i + 2];
    }
    // raw 192-d descriptors: CompNet + re-normalisation first (descriptor_DR.py:141-165)
    if (F->des_len != 0 && F->des_len != kDesLen && F->des_len != 2 * kDesLen) return LAFIS_ERR_ARG;
    const bool raw = F->des_len == 2 * kDesLen;
    std::vector<float> tex_small;
    const float* tex_des = F->tex_des;
    if (raw) {
        minu.des.resize((size_t)nm * kDesLen);
        tex_small.resize((size_t)nt * kDesLen);
        int rc = nm > 0 ? lafis_compress_descriptors(ctx, F->minu_des, nm, minu.des.data(), 1, 0) : LAFIS_OK;
        if (rc == LAFIS_OK && nt > 0) rc = lafis_compress_descriptors(ctx, F->tex_des, nt, tex_small.data(), 1, 0);
        if (rc != LAFIS_OK) return rc;
        tex_des = tex_small.data();
    } else {
        minu.des.assign(F->minu_des, F->minu_des + (size_t)nm * kDesLen);
    }
    tex.x.resize(nt);
    tex.y.resize(nt);
    tex.ori.resize(nt);
    for (int i = 0; i < nt; ++i) {  // block units (descriptor_PQ.py:249-256)
        tex.x[i] = (int16_t)(uint16_t)((F->tex_xyo[3 * i] - 24.0f) / 16.0f);
        tex.y[i] = (int16_t)(uint16_t)((F->tex_xyo[3 * i + 1] - 24.0f) / 16.0f);
        tex.ori[i] = F->tex_xyo[3 * i + 2];
    }
    tex.codes.resize((size_t)nt * kSubs);
    if (nt > 0) {
        const int rc = lafis_pq_encode(ctx, tex_des, nt, tex.codes.data(), 0);
        if (rc != LAFIS_OK) return rc;
    }
    return write_rolled_dat(out_path, F->h, F->w, F->blkH, F->blkW, minu, tex) == 0 ? LAFIS_OK : LAFIS_ERR_IO;
}

LAFIS_API int lafis_enroll_latent(lafis_ctx* ctx, const lafis_latent_features* F, const char* out_path) {
    if (!ctx || !F || !out_path || F->n_minu_templates < 0 || F->n_minu_templates > 255 || F->n_tex_templates < 0 ||
        F->n_tex_templates > 255 || (F->n_minu_templates > 0 && !F->minu) || (F->n_tex_templates > 0 && !F->tex))
        return LAFIS_ERR_ARG;
    if (F->des_len != 0 && F->des_len != kDesLen && F->des_len != 2 * kDesLen) return LAFIS_ERR_ARG;
    const int in_len = F->des_len == 2 * kDesLen ? 2 * kDesLen : kDesLen;
    const int nmt = F->n_minu_templates, ntt = F->n_tex_templates;
    std::vector<PointSet> sets(nmt + ntt);
    std::vector<float> raw;  // 192-d descriptors of the whole print, template after template
    size_t total = 0;
    for (int t = 0; t < nmt + ntt; ++t) {
        const lafis_point_set& in = t < nmt ? F->minu[t] : F->tex[t - nmt];
        if (in.n < 0 || (in.n > 0 && (!in.xyo || !in.des))) return LAFIS_ERR_ARG;
        total += (size_t)std::min(in.n, kMaxMinutiae);
    }
    if (in_len != kDesLen) raw.reserve(total * in_len);
    for (int t = 0; t < nmt + ntt; ++t) {
        const bool is_tex = t >= nmt;
        const lafis_point_set& in = is_tex ? F->tex[t - nmt] : F->minu[t];
        PointSet& s = sets[t];
        const int n = std::min(in.n, kMaxMinutiae);
        s.x.resize(n);
        s.y.resize(n);
        s.ori.resize(n);
        for (int i = 0; i < n; ++i) {
            if (is_tex) {  // block units (descriptor_PQ.py:149-156)
                s.x[i] = (int16_t)(uint16_t)((in.xyo[3 * i] - 24.0f) / 16.0f);
                s.y[i] = (int16_t)(uint16_t)((in.xyo[3 * i + 1] - 24.0f) / 16.0f);
            } else {  // np.uint16(x): truncation (:118-123)
                s.x[i] = (int16_t)(uint16_t)in.xyo[3 * i];
                s.y[i] = (int16_t)(uint16_t)in.xyo[3 * i + 1];
            }
            s.ori[i] = in.xyo[3 * i + 2];
        }
        if (in_len == kDesLen) s.des.assign(in.des, in.des + (size_t)n * kDesLen);
        else raw.insert(raw.end(), in.des, in.des + (size_t)n * in_len);
    }
    if (in_len != kDesLen && total > 0) {
        std::vector<float> small(total * kDesLen);
        const int rc = lafis_compress_descriptors(ctx, raw.data(), (int64_t)total, small.data(), 1, 0);
        if (rc != LAFIS_OK) return rc;
        size_t at = 0;
        for (PointSet& s : sets) {
            s.des.assign(small.begin() + at * kDesLen, small.begin() + (at + s.n()) * kDesLen);
            at += s.n();
        }
    }
    const std::vector<PointSet> minu(sets.begin(), sets.begin() + nmt), tex(sets.begin() + nmt, sets.end());
    return write_latent_dat(out_path, F->h, F->w, F->blkH, F->blkW, minu, tex) == 0 ? LAFIS_OK : LAFIS_ERR_IO;
}

}  // extern "C"

namespace {

int one2list(Engine& E, const char* latent_template_file, const char* rolled_dir, const char* score_path) {
    lafis_ctx* ctx = E.latents_ctx();
    const fs::path latent_file(latent_template_file);
    const std::string score_file = std::string(score_path) + latent_file.stem().string() + ".csv";
    int rc = ensure_gallery(E, rolled_dir, false);
    if (rc != LAFIS_OK) return rc;
    const int G = E.gallery_size();
    const auto t0 = std::chrono::high_resolution_clock::now();
    std::cout << "Latent Query: " << latent_file << std::endl;
    std::cout << "Gallery size: " << G << std::endl;

    lafis_latents* L = nullptr;
    const char* lp = latent_template_file;
    rc = lafis_latents_load_files(ctx, &lp, 1, &L);
    if (rc != LAFIS_OK) return rc;
    const int st = lafis_latents_status(L, 0);
    if (st == LAFIS_LATENT_EMPTY) {
        // matcher.cpp:260-267 writes "0" when the latent holds no template at all; every pair then
        // returns 1 and the driver exits with 1 (:296-299).  A latent with 1..26 minutiae templates
        // and no texture template takes the same exit without the file.
        LatentTemplate T;
        read_latent_dat(latent_template_file, T);
        if (T.n_minu_templates <= 0 && T.n_tex_templates <= 0) {
            std::ofstream output(score_file);
            output << 0 << std::endl;
        }
        std::cout << "Matching failed: latent template is empty. Exiting." << std::endl;
        lafis_latents_free(L);
        return LAFIS_LATENT_EMPTY;
    }
    if (st != LAFIS_OK) {
        lafis_latents_free(L);
        return st;
    }
    std::vector<float> scores((size_t)G, -1.0f);
    rc = E.match_scores(L, scores.data());
    if (rc != LAFIS_OK) {
        lafis_latents_free(L);
        return rc;
    }
    for (int j = 0; j < G; ++j)
        if (scores[j] == -1.0f && E.gallery_status(j) != LAFIS_TPL_OK && E.gallery_status(j) != LAFIS_TPL_TRUNCATED)
            std::cout << "Comparison failed: rolled template is empty. Skipping." << std::endl;

    // rank list: the reference's own call, std::sort with (scores[a] > scores[b]) (matcher.cpp:306-309)
    std::vector<int> ind((size_t)G);
    std::iota(ind.begin(), ind.end(), 0);
    std::sort(ind.begin(), ind.end(), [&](const int& a, const int& b) { return scores[a] > scores[b]; });
    std::ofstream output(score_file);
    output << "filename,score" << std::endl;
    std::cout << "Match Results" << std::endl;
    std::cout << "----------------" << std::endl;
    std::cout << "Rank     Filename      Score" << std::endl;
    // correspondence files of the 24 best (matcher.cpp:322-327, written by :497-505); the reference's prefix
    // "/LatentAFIS/scores/" is configurable here
    const char* corr_env = std::getenv("LAFIS_CORR_PATH");
    const std::string corr_prefix = corr_env ? std::string(corr_env) : std::string(score_path);
    std::vector<int16_t> corr_xy((size_t)3 * LAFIS_MAX_CORR * 4);
    const int latent_nm = lafis_latents_minu_templates(L, 0);
    for (int j = 0; j < 24 && j < G; ++j) {
        const fs::path rolled(E.gallery_path(ind[j]));
        output << std::to_string(j + 1) << quoted_path(rolled.string()) << "," << scores[ind[j]] << std::endl;
        const int gst = E.gallery_status(ind[j]);
        // the reference returns before matching an empty print (:388-391) and only enters the minutiae loop when the
        // rolled print has a minutiae template (:400)
        if ((gst == LAFIS_TPL_OK || gst == LAFIS_TPL_TRUNCATED) && E.gallery_minutiae(ind[j]) > 0) {
            int counts[3] = {0, 0, 0};
            rc = E.correspondences(L, 0, ind[j], corr_xy.data(), counts);
            if (rc != LAFIS_OK) {
                lafis_latents_free(L);
                return rc;
            }
            const std::string corr_file = corr_prefix + "corr" + latent_file.stem().string() + "_" + rolled.stem().string();
            for (int i = 0; i < 3; ++i) {
                if (latent_nm <= kSelected[i]) continue;  // :403-404: no such latent template, no file
                std::ofstream co(corr_file + "_" + std::to_string(i) + ".csv");
                for (int k = 0; k < counts[i]; ++k) {
                    const int16_t* e = &corr_xy[((size_t)i * LAFIS_MAX_CORR + k) * 4];
                    co << e[0] << "," << e[1] << "," << e[2] << "," << e[3] << std::endl;
                }
            }
        }
        std::cout << std::to_string(j + 1) << "        " << rolled.filename() << "       " << scores[ind[j]] << std::endl;
    }
    lafis_latents_free(L);
    output.close();
    const std::chrono::duration<double, std::milli> span = std::chrono::high_resolution_clock::now() - t0;
    std::cout << "Total matching duration (ms): " << span.count() << std::endl;
    if (getenv("LAFIS_INGEST_TIMING")) std::cerr << "lafis cli: matching (gallery resident) took " << span.count() << " ms" << std::endl;
    return LAFIS_OK;
}

int list2list(Engine& E, const char* latent_dir, const char* rolled_dir, const char* score_path) {
    lafis_ctx* ctx = E.latents_ctx();
    std::vector<std::string> latent_files = list_dat_files(latent_dir);
    for (const std::string& f : latent_files) std::cout << "latent template file" << fs::path(f) << std::endl;
    if (latent_files.empty()) {
        std::cout << "No latent templates found in directory: " << latent_dir << std::endl;
        return LAFIS_ERR_NO_TEMPLATES;
    }
    int rc = ensure_gallery(E, rolled_dir, true);
    if (rc != LAFIS_OK) return rc;
    const int G = E.gallery_size();
    std::cout << "Gallery size: " << G << std::endl;
    const auto t0 = std::chrono::high_resolution_clock::now();

    const int n = (int)latent_files.size();
    // batches bounded by the size of the host score matrix (256 MB)
    const int batch = std::max(1, std::min(n, (int)(((size_t)64 << 20) / (size_t)std::max(G, 1))));
    std::vector<float> scores;
    std::vector<std::string> quoted;  // gallery paths as the score files print them, built once
    std::string rows;
    for (int b0 = 0; b0 < n; b0 += batch) {
        const int nb = std::min(batch, n - b0);
        std::vector<const char*> ptrs(nb);
        for (int i = 0; i < nb; ++i) ptrs[i] = latent_files[b0 + i].c_str();
        lafis_latents* L = nullptr;
        rc = lafis_latents_load_files(ctx, ptrs.data(), nb, &L);
        if (rc != LAFIS_OK) return rc;
        scores.assign((size_t)nb * G, -1.0f);
        bool any = false;
        for (int i = 0; i < nb; ++i) any = any || lafis_latents_status(L, i) == LAFIS_OK;
        if (any) {
            rc = E.match_scores(L, scores.data());
            if (rc != LAFIS_OK) {
                lafis_latents_free(L);
                return rc;
            }
        }
        for (int i = 0; i < nb; ++i) {
            const fs::path lf(latent_files[b0 + i]);
            std::cout << lf << std::endl;
            const std::string out_name = std::string(score_path) + lf.stem().string() + ".csv";
            const int st = lafis_latents_status(L, i);
            if (st == LAFIS_LATENT_EMPTY) {
                LatentTemplate T;
                read_latent_dat(latent_files[b0 + i], T);
                std::cout << "Latent minutiae templates: " << T.n_minu_templates << std::endl;
                std::cout << "Latent texture templates: " << T.n_tex_templates << std::endl;
                if (T.n_minu_templates <= 0 && T.n_tex_templates <= 0) {  // matcher.cpp:153-163
                    std::cout << "No minutiae or texture templates found" << std::endl;
                    std::ofstream output(out_name);
                    output << 0 << std::endl;
                } else {  // matcher.cpp:191-194
                    std::cout << "Matching failed: latent template is empty. Skipping." << std::endl;
                }
                continue;
            }
            if (st != LAFIS_OK) {
                std::cout << "Matching failed: latent template layout is outside the matcher's domain. Skipping."
                          << std::endl;
                continue;
            }
            // one buffered write per latent (the reference flushes every row: 10^6 system calls per file at scale)
            if (quoted.empty() && G > 0) {
                quoted.resize(G);
                for (int j = 0; j < G; ++j) quoted[j] = quoted_path(E.gallery_path(j));
            }
            rows.clear();
            for (int j = 0; j < G; ++j) append_score_row(rows, quoted[j], scores[(size_t)i * G + j]);
            std::ofstream output(out_name, std::ios::binary);
            output.write(rows.data(), (std::streamsize)rows.size());
        }
        lafis_latents_free(L);
    }
    const std::chrono::duration<double, std::milli> span = std::chrono::high_resolution_clock::now() - t0;
    std::cout << "Total matching duration (ms): " << span.count() << std::endl;
    if (getenv("LAFIS_INGEST_TIMING")) std::cerr << "lafis cli: matching (gallery resident) took " << span.count() << " ms" << std::endl;
    return LAFIS_OK;
}

}  // namespace
