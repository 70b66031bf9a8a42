// C-callable harness around the UNMODIFIED reference matcher — TEST INFRASTRUCTURE ONLY.
//
// Built by oracle/build_ref.sh into oracle/_ref/libref_matcher.so together with the reference's
// own matching/matcher.cpp (compiled from /root/reference where it lies; the only edit is the
// missing `return 0;` at matcher.cpp:374 and :417, applied to a throw-away copy in a temp
// directory, see build_ref.sh).  Nothing here re-implements the algorithm: every entry point
// forwards to a PQ::Matcher member (matching/matcher.h:34-75).
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <numeric>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>
#include <omp.h>

// The graph-pruning members are private in matcher.h:55-61; the stage-level parity tests need
// them.  All standard headers matcher.h pulls in are already included above, so this only
// affects the reference's own class.
#define private public
#include "matcher.h"
#undef private

using namespace PQ;

namespace {
typedef std::vector<std::tuple<float, int, int> > CorrVec;

CorrVec to_corr(const float* v, const int* li, const int* rj, int n) {
    CorrVec c;
    c.reserve(n);
    for (int i = 0; i < n; ++i) c.push_back(std::make_tuple(v[i], li[i], rj[i]));
    return c;
}
int from_corr(const CorrVec& c, float* v, int* li, int* rj) {
    for (size_t i = 0; i < c.size(); ++i) {
        v[i] = std::get<0>(c[i]);
        li[i] = std::get<1>(c[i]);
        rj[i] = std::get<2>(c[i]);
    }
    return static_cast<int>(c.size());
}
}  // namespace

extern "C" {

void* ref_matcher_new(const char* codebook) { return new Matcher(std::string(codebook)); }
void ref_matcher_free(void* m) { delete static_cast<Matcher*>(m); }

void* ref_load_latent(void* m, const char* path, int* rc) {
    LatentFPTemplate* t = new LatentFPTemplate();
    int r = static_cast<Matcher*>(m)->load_FP_template(std::string(path), *t);
    if (rc) *rc = r;
    return t;
}
void* ref_load_rolled(void* m, const char* path, int* rc) {
    RolledFPTemplate* t = new RolledFPTemplate();
    int r = static_cast<Matcher*>(m)->load_FP_template(std::string(path), *t);
    if (rc) *rc = r;
    // matcher.cpp:173-177: a negative return empties the template
    if (r < 0) {
        t->m_nrof_minu_templates = 0;
        t->m_nrof_texture_templates = 0;
    }
    return t;
}
void ref_free_latent(void* t) { delete static_cast<LatentFPTemplate*>(t); }
void ref_free_rolled(void* t) { delete static_cast<RolledFPTemplate*>(t); }

int ref_latent_counts(void* t, int* n_minu_templates, int* n_tex_templates) {
    LatentFPTemplate* l = static_cast<LatentFPTemplate*>(t);
    *n_minu_templates = l->m_nrof_minu_templates;
    *n_tex_templates = l->m_nrof_texture_templates;
    return 0;
}

// One (latent, rolled) comparison exactly as the drivers do it (matcher.cpp:179-189, 284-294).
// comp[0..2] = the three minutiae-template scores, comp[3] = score[28]; *final_score is the fused
// value or -1 when the reference would have left the slot untouched.
int ref_score_pair(void* m, void* latent, void* rolled, float* comp, float* final_score) {
    LatentFPTemplate* l = static_cast<LatentFPTemplate*>(latent);
    RolledFPTemplate* r = static_cast<RolledFPTemplate*>(rolled);
    std::vector<float> score;
    int result = static_cast<Matcher*>(m)->One2One_matching_selected_templates(*l, *r, score);
    comp[0] = comp[1] = comp[2] = comp[3] = 0.f;
    *final_score = -1.f;
    if (result == 1 || result == 2) return result;
    if (score.size() < 29) return -100;  // score[28] would be out of bounds (SURVEY §7 hard parts)
    comp[0] = score[0];
    comp[1] = score[1];
    comp[2] = score[2];
    comp[3] = score[28];
    float fs = score[0] + score[1] + score[2] + score[28] * 0.3;  // matcher.cpp:188 verbatim expression
    *final_score = fs;
    return 0;
}

// The reference's OpenMP loop over a pre-loaded gallery (the "preloaded harness" of
// BASELINE.md §3): same schedule clause as matcher.cpp:168, thread count chosen by the caller.
int ref_score_gallery(void* m, void* latent, void** rolled, int n, float* finals, float* comps, int nthreads) {
    Matcher* mm = static_cast<Matcher*>(m);
    LatentFPTemplate* l = static_cast<LatentFPTemplate*>(latent);
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    int bad = 0;
#pragma omp parallel for num_threads(nthreads) schedule(static, 16)
    for (int j = 0; j < n; ++j) {
        float c[4], f;
        int rc = ref_score_pair(mm, l, rolled[j], c, &f);
        if (rc == -100) {
#pragma omp atomic
            bad++;
        }
        finals[j] = f;
        if (comps) std::memcpy(comps + 4 * j, c, sizeof(c));
    }
    return bad ? -100 : 0;
}

// The save_corr branch (matcher.cpp:497-505) as One2List_matching drives it for its 24 best (:322-327): writes
// "<corr_file>_<i>.csv", i = 0..2, one "lx,ly,rx,ry" row per surviving correspondence.
int ref_save_corr(void* m, void* latent, void* rolled, const char* corr_file) {
    LatentFPTemplate* l = static_cast<LatentFPTemplate*>(latent);
    RolledFPTemplate* r = static_cast<RolledFPTemplate*>(rolled);
    std::vector<float> score;
    return static_cast<Matcher*>(m)->One2One_matching_selected_templates(*l, *r, score, true, std::string(corr_file));
}

int ref_one2list(void* m, const char* latent_file, const char* gallery_dir, const char* score_dir) {
    return static_cast<Matcher*>(m)->One2List_matching(latent_file, gallery_dir, score_dir);
}
int ref_list2list(void* m, const char* latent_dir, const char* gallery_dir, const char* score_dir) {
    return static_cast<Matcher*>(m)->List2List_matching(latent_dir, gallery_dir, score_dir);
}

// ---- stage-level access (private members of PQ::Matcher) ----
float ref_minutiae_score(void* m, void* latent, int latent_tpl, void* rolled) {
    LatentFPTemplate* l = static_cast<LatentFPTemplate*>(latent);
    RolledFPTemplate* r = static_cast<RolledFPTemplate*>(rolled);
    return static_cast<Matcher*>(m)->One2One_minutiae_matching(l->m_minu_templates[latent_tpl], r->m_minu_templates[0]);
}
float ref_texture_score(void* m, void* latent, void* rolled) {
    LatentFPTemplate* l = static_cast<LatentFPTemplate*>(latent);
    RolledFPTemplate* r = static_cast<RolledFPTemplate*>(rolled);
    return static_cast<Matcher*>(m)->One2One_texture_matching(l->m_texture_templates[0], r->m_texture_templates[0]);
}
// which: 0 = LSS_R_Fast2_Dist_lookup on the texture templates (matcher.cpp:1225),
//        1 = LSS_R_Fast2_Dist_eigen on minutiae template latent_tpl (matcher.cpp:1350),
//        2 = LSS_R_Fast2 on the texture templates, 3 = LSS_R_Fast2 on minutiae (matcher.cpp:1471)
int ref_prune(void* m, int which, void* latent, int latent_tpl, void* rolled, const float* v, const int* li,
              const int* rj, int n, float* ov, int* oli, int* orj) {
    Matcher* mm = static_cast<Matcher*>(m);
    LatentFPTemplate* l = static_cast<LatentFPTemplate*>(latent);
    RolledFPTemplate* r = static_cast<RolledFPTemplate*>(rolled);
    CorrVec c = to_corr(v, li, rj, n), out;
    switch (which) {
        case 0: out = mm->LSS_R_Fast2_Dist_lookup(c, l->m_texture_templates[0], r->m_texture_templates[0], 30); break;
        case 1: out = mm->LSS_R_Fast2_Dist_eigen(c, l->m_minu_templates[latent_tpl], r->m_minu_templates[0], 30); break;
        case 2: out = mm->LSS_R_Fast2(c, l->m_texture_templates[0], r->m_texture_templates[0], 30); break;
        case 3: out = mm->LSS_R_Fast2(c, l->m_minu_templates[latent_tpl], r->m_minu_templates[0], 30); break;
        default: return -1;
    }
    return from_corr(out, ov, oli, orj);
}

// libstdc++ std::sort with the reference's comparator shape (matcher.cpp:475-476, 740-741,
// 1300-1301, 1422-1423, 1589-1590): indices sorted by key descending.  Used to pin the
// permutation emulation in oracle/lafis_oracle.c.
void ref_std_sort_desc(const float* key, int* idx, int n) {
    std::vector<int> y(n);
    std::iota(y.begin(), y.end(), 0);
    auto comparator = [key](int a, int b) { return key[a] > key[b]; };
    std::sort(y.begin(), y.end(), comparator);
    std::copy(y.begin(), y.end(), idx);
}

float ref_atan2f(float y, float x) { return atan2(y, x); }  // resolves to the float overload, as in matcher.cpp:1516

int ref_omp_max_threads(void) { return omp_get_max_threads(); }
}
