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

#include "../../include/latentafis_b200.h"
#include "dat_format.h"

namespace fs = std::filesystem;
using namespace lafis;

namespace {

// the gallery already resident on `ctx` is reused when it was loaded from the same directory
struct DirCache {
    lafis_ctx* ctx = nullptr;
    std::string dir;
};
DirCache g_cache;

int ensure_gallery(lafis_ctx* ctx, const std::string& dir, bool announce) {
    if (g_cache.ctx == ctx && g_cache.dir == dir && lafis_gallery_size(ctx) > 0) return LAFIS_OK;
    std::vector<std::string> files = list_dat_files(dir);
    if (announce)
        for (const std::string& f : files) std::cout << "rolled template file" << fs::path(f) << std::endl;
    if (files.empty()) {
        std::cout << "No rolled templates found in directory: " << dir << std::endl;
        return LAFIS_ERR_NO_TEMPLATES;
    }
    std::vector<const char*> ptrs(files.size());
    for (size_t i = 0; i < files.size(); ++i) ptrs[i] = files[i].c_str();
    int rc = lafis_gallery_load_files(ctx, ptrs.data(), (int)ptrs.size(), 0, 1);
    if (rc == LAFIS_OK) {
        g_cache.ctx = ctx;
        g_cache.dir = dir;
    }
    return rc;
}

}  // namespace

extern "C" {

LAFIS_API void lafis_forget_gallery_dir(lafis_ctx* ctx) {
    if (g_cache.ctx == ctx) g_cache = DirCache();
}

LAFIS_API int lafis_one2list_matching(lafis_ctx* ctx, const char* latent_template_file, const char* rolled_dir,
                                      const char* score_path) {
    if (!ctx || !latent_template_file || !rolled_dir || !score_path) return LAFIS_ERR_ARG;
    const fs::path latent_file(latent_template_file);
    const std::string score_file = std::string(score_path) + latent_file.stem().string() + ".csv";
    int rc = ensure_gallery(ctx, rolled_dir, false);
    if (rc != LAFIS_OK) return rc;
    const int G = lafis_gallery_size(ctx);
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
    rc = lafis_match(ctx, L, 0, nullptr, scores.data(), nullptr);
    if (rc != LAFIS_OK) {
        lafis_latents_free(L);
        return rc;
    }
    for (int j = 0; j < G; ++j)
        if (scores[j] == -1.0f && lafis_gallery_status(ctx, j) != LAFIS_TPL_OK &&
            lafis_gallery_status(ctx, j) != LAFIS_TPL_TRUNCATED)
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
    for (int j = 0; j < 24 && j < G; ++j) {
        const fs::path rolled(lafis_gallery_path(ctx, ind[j]));
        output << std::to_string(j + 1) << quoted_path(rolled.string()) << "," << scores[ind[j]] << std::endl;
        const int gst = lafis_gallery_status(ctx, ind[j]);
        if (gst == LAFIS_TPL_OK || gst == LAFIS_TPL_TRUNCATED) {  // the reference returns before matching an empty print (:388-391)
            int counts[3] = {0, 0, 0};
            rc = lafis_correspondences(ctx, L, 0, ind[j], corr_xy.data(), counts);
            if (rc != LAFIS_OK) {
                lafis_latents_free(L);
                return rc;
            }
            const std::string corr_file = corr_prefix + "corr" + latent_file.stem().string() + "_" + rolled.stem().string();
            for (int i = 0; i < 3; ++i) {
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
    return LAFIS_OK;
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

LAFIS_API int lafis_list2list_matching(lafis_ctx* ctx, const char* latent_dir, const char* rolled_dir,
                                       const char* score_path) {
    if (!ctx || !latent_dir || !rolled_dir || !score_path) return LAFIS_ERR_ARG;
    std::vector<std::string> latent_files = list_dat_files(latent_dir);
    for (const std::string& f : latent_files) std::cout << "latent template file" << fs::path(f) << std::endl;
    if (latent_files.empty()) {
        std::cout << "No latent templates found in directory: " << latent_dir << std::endl;
        return LAFIS_ERR_NO_TEMPLATES;
    }
    int rc = ensure_gallery(ctx, rolled_dir, true);
    if (rc != LAFIS_OK) return rc;
    const int G = lafis_gallery_size(ctx);
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
            rc = lafis_match(ctx, L, 0, nullptr, scores.data(), nullptr);
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
                for (int j = 0; j < G; ++j) quoted[j] = quoted_path(lafis_gallery_path(ctx, j));
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
    return LAFIS_OK;
}

}  // extern "C"
