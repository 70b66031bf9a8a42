// Shared-memory budgets of the minutiae kernels and the per-call plan that sizes their tiles - plain C++ so that the
// host library, the CUDA kernels and the CPU unit tests (lib/libhostcheck.so) share one definition.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <algorithm>

#if defined(__CUDACC__)
#define LAFIS_PLAN_HD __host__ __device__
#else
#define LAFIS_PLAN_HD
#endif

namespace lafis {

constexpr int kMaxDynSmem = 227 * 1024 - 1024;  // kernels also hold a little static shared memory
constexpr int kSelMaxCand = 512;  // indices + keys in the histogram's 4 KB (2 x 512 x 4 B)
constexpr int kSelBins = 1024;  // 64 bins per binade over [2^-16, 1): float bits >> 17, offset; smaller values share bin 0

LAFIS_PLAN_HD inline size_t minu_sim_smem_bytes(int a_slot_stride, int b_buf_stride, int b_double) {
    return sizeof(float) * ((size_t)3 * a_slot_stride + (size_t)(b_double ? 2 : 1) * b_buf_stride);
}

// Row stride (floats) of minu_select_kernel's copy of a similarity matrix with np (a multiple of 4) padded columns:
// the next multiple of 4 that is 4 mod 8 - rows stay 16-byte aligned and eight consecutive rows start in eight different
// 16-byte bank groups.
LAFIS_PLAN_HD inline int sel_ld(int np) { return (np & 4) ? np : np + 4; }

// the copy of S, row and column sums, and 4 KB that hold the 1024-bin histogram first and the candidate list
// (kSelMaxCand indices + kSelMaxCand keys) afterwards
LAFIS_PLAN_HD inline size_t minu_select_smem_bytes(int max_nL, int max_np) {
    return sizeof(float) * ((size_t)max_nL * sel_ld(max_np) + ((max_nL + 3) & ~3) + max_np) + sizeof(int) * kSelBins + 16;
}

LAFIS_PLAN_HD inline size_t minu_select_slow_smem_bytes(int max_nL, int max_np, bool dense) {
    return sizeof(float) * ((size_t)max_nL * (max_np + 1) + max_nL + max_np) +
           (sizeof(uint16_t) + (dense ? sizeof(uint32_t) : 0)) * (size_t)max_nL * max_np + 16;
}

// Shared-memory geometry of the fast minutiae kernels (minu_sim.cuh) for a latent batch and a gallery.  They take
// latent slots of <= l_cap and gallery templates of <= r_cap minutiae; everything larger goes to minu_big.cuh.
struct MinuPlan {
    int l_cap = 1, r_cap = 4;
    int maxL = 1, maxNp = 4;
    int a_slot_stride = 0, b_buf_stride = 0, b_double = 1;
    size_t sim_smem = 0, sel_smem = 0, slow_smem = 0, job_stride = 0;
    bool slow_dense = false;
    bool efficient = false;  // double-buffered similarity kernel and four selection CTAs per SM
};
constexpr int kFastMaxL = 128;  // three resident latent blocks of 128 minutiae are 148 KB of shared memory

inline bool minu_plan_fill(MinuPlan& p, int L, int Np) {
    p.maxL = L;
    p.maxNp = Np;
    const int maxLp = (L + 3) & ~3;
    p.a_slot_stride = 96 * maxLp + 16;
    p.b_buf_stride = 96 * Np + 160;
    p.b_double = minu_sim_smem_bytes(p.a_slot_stride, p.b_buf_stride, 1) <= (size_t)kMaxDynSmem ? 1 : 0;
    p.sim_smem = minu_sim_smem_bytes(p.a_slot_stride, p.b_buf_stride, p.b_double);
    p.sel_smem = minu_select_smem_bytes(L, Np);
    // the dense key copy enables the block-parallel introsort replay; one CTA per SM is enough for this rare path
    p.slow_dense = minu_select_slow_smem_bytes(L, Np, true) <= (size_t)kMaxDynSmem;
    p.slow_smem = minu_select_slow_smem_bytes(L, Np, p.slow_dense);
    p.job_stride = (size_t)L * Np;
    p.efficient = p.b_double && 4 * (p.sel_smem + 2048) <= (size_t)228 * 1024;
    p.l_cap = L;
    p.r_cap = Np;
    return p.sim_smem <= (size_t)kMaxDynSmem && p.slow_smem <= (size_t)kMaxDynSmem && p.sel_smem <= (size_t)kMaxDynSmem &&
           (size_t)L * Np < 65536;
}

// max_nR: largest template among the n templates whose minutiae counts are in h_n (may be NULL)
inline MinuPlan plan_minu(int max_slot_n, int max_nR, const uint16_t* h_n, size_t n) {
    MinuPlan p;
    const int L = std::min(std::max(1, max_slot_n), kFastMaxL);
    int Np = std::max(4, (max_nR + 3) & ~3);
    while (Np > 4 && !minu_plan_fill(p, L, Np)) Np -= 4;
    minu_plan_fill(p, L, Np);
    if (!p.efficient && h_n && n > 0) {
        // one outsized template must not push the whole gallery onto the slow geometry: when at most 0.5 % of the
        // templates exceed the largest efficient tile, they are the ones that go to the big-pair kernels
        MinuPlan e;
        int Ne = Np;
        while (Ne > 4) {
            if (minu_plan_fill(e, L, Ne) && e.efficient) break;
            Ne -= 4;
        }
        if (Ne > 4 && Ne < Np) {
            size_t big = 0;
            for (size_t i = 0; i < n; ++i) big += h_n[i] > Ne;
            if (big * 200 <= n) p = e;
        }
    }
    return p;
}

}  // namespace lafis
