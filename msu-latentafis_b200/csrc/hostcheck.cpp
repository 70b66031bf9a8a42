// Host build of the bit-exact helper headers and the .dat parsers for the CPU unit tests
// (tests/test_host_math.py, tests/test_dat_format.py).  Not part of the GPU library.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

#include "dat_format.h"
#include "minu_plan.h"
#include "exact_math.h"
#include "stdsort_emul.h"

using namespace lafis;

extern "C" {

float hc_atan2f(float y, float x) { return atan2f_fdlibm(y, x); }
void hc_atan2f_many(const float* y, const float* x, float* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = atan2f_fdlibm(y[i], x[i]);
}

// first `need` positions of the emulated std::sort permutation
void hc_sort_prefix(const float* key, int n, int need, int* out) {
    std::vector<int> y((size_t)std::max(n, 1));
    std_sort_desc_prefix(DenseKey<float>{key}, y.data(), n, need);
    std::copy(y.begin(), y.begin() + std::min(n, need), out);
}
void hc_sort_prefix_u16(const float* key, int n, int need, int* out) {
    std::vector<uint16_t> y((size_t)std::max(n, 1));
    std_sort_desc_prefix(DenseKey<float>{key}, y.data(), n, need);
    for (int i = 0; i < std::min(n, need); ++i) out[i] = y[i];
}
// the same replay with every partition step resolved by the closed form the block-cooperative device version uses
void hc_sort_prefix_closed(const float* key, int n, int need, int* out) {
    std::vector<int> y((size_t)std::max(n, 1)), ls((size_t)std::max(n, 1)), rs((size_t)std::max(n, 1));
    for (int i = 0; i < n; ++i) y[i] = i;
    StdSortEmu<DenseKey<float>, int> s{DenseKey<float>{key}, y.data()};
    s.sort_prefix(n, need, ls.data(), rs.data());
    std::copy(y.begin(), y.begin() + std::min(n, need), out);
}
// libstdc++'s own answer, with the reference's comparator shape (matcher.cpp:475-476)
void hc_std_sort(const float* key, int n, int* out) {
    std::vector<int> y((size_t)n);
    std::iota(y.begin(), y.end(), 0);
    std::sort(y.begin(), y.end(), [key](int a, int b) { return key[a] > key[b]; });
    std::copy(y.begin(), y.end(), out);
}

// parsers: counts and a checksum of what was kept
int hc_read_rolled(const char* path, int* status, int* n_minu_t, int* n_tex_t, int* n_minu, int* n_tex) {
    RolledTemplate R;
    int rc = read_rolled_dat(path, R);
    *status = R.status;
    *n_minu_t = R.n_minu_templates;
    *n_tex_t = R.n_tex_templates;
    *n_minu = R.minu.n();
    *n_tex = R.tex.n();
    return rc;
}
int hc_read_latent(const char* path, int* n_minu_t, int* n_tex_t, int* slot_n, int* n_tex) {
    LatentTemplate T;
    int rc = read_latent_dat(path, T);
    *n_minu_t = T.n_minu_templates;
    *n_tex_t = T.n_tex_templates;
    for (int s = 0; s < 3; ++s) slot_n[s] = T.minu[s].n();
    *n_tex = T.tex.n();
    return rc;
}
int hc_rolled_arrays(const char* path, short* mx, short* my, float* mori, float* mdes, short* tx, short* ty, float* tori,
                     unsigned char* codes) {
    RolledTemplate R;
    int rc = read_rolled_dat(path, R);
    auto cp = [](void* d, const void* s, size_t b) {
        if (d && b) std::memcpy(d, s, b);
    };
    cp(mx, R.minu.x.data(), 2 * R.minu.x.size());
    cp(my, R.minu.y.data(), 2 * R.minu.y.size());
    cp(mori, R.minu.ori.data(), 4 * R.minu.ori.size());
    cp(mdes, R.minu.des.data(), 4 * R.minu.des.size());
    cp(tx, R.tex.x.data(), 2 * R.tex.x.size());
    cp(ty, R.tex.y.data(), 2 * R.tex.y.size());
    cp(tori, R.tex.ori.data(), 4 * R.tex.ori.size());
    cp(codes, R.tex.codes.data(), R.tex.codes.size());
    return rc;
}
// write_latent_dat from flat arrays: counts[t] points per template (minutiae templates first), arrays concatenated
int hc_write_latent(const char* path, int h, int w, int blkH, int blkW, int n_minu_t, int n_tex_t, const int* counts,
                    const short* x, const short* y, const float* ori, const float* des) {
    std::vector<PointSet> minu(n_minu_t), tex(n_tex_t);
    size_t at = 0;
    for (int t = 0; t < n_minu_t + n_tex_t; ++t) {
        PointSet& s = t < n_minu_t ? minu[t] : tex[t - n_minu_t];
        const size_t n = (size_t)counts[t];
        s.x.assign(x + at, x + at + n);
        s.y.assign(y + at, y + at + n);
        s.ori.assign(ori + at, ori + at + n);
        s.des.assign(des + at * kDesLen, des + (at + n) * kDesLen);
        at += n;
    }
    return write_latent_dat(path, h, w, blkH, blkW, minu, tex);
}
// the per-call tile plan of the fast minutiae kernels (minu_plan.h)
void hc_plan_minu(int max_slot_n, int max_nR, const unsigned short* h_n, long n, long* out) {
    const MinuPlan p = plan_minu(max_slot_n, max_nR, h_n, (size_t)n);
    out[0] = p.l_cap;
    out[1] = p.r_cap;
    out[2] = p.b_double;
    out[3] = p.efficient ? 1 : 0;
    out[4] = p.slow_dense ? 1 : 0;
    out[5] = (long)p.sim_smem;
    out[6] = (long)p.sel_smem;
    out[7] = (long)p.slow_smem;
    out[8] = (long)p.job_stride;
}
// rows of the N-vs-N score file as the driver writes them; returns the length (the buffer holds at most cap bytes)
long hc_format_score_rows(const char* const* paths, const float* scores, int n, char* out, long cap) {
    std::string buf;
    for (int i = 0; i < n; ++i) append_score_row(buf, quoted_path(paths[i]), scores[i]);
    std::memcpy(out, buf.data(), std::min<size_t>(buf.size(), (size_t)cap));
    return (long)buf.size();
}
}
