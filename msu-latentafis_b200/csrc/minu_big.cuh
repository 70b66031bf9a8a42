// Minutiae stages K5 + K6 + K7 for pairs that do not fit the shared-memory tiles of minu_sim.cuh: a latent
// template of more than 128 minutiae, or a gallery template beyond what the fast kernels were sized for (the
// reference accepts up to 2000 minutiae per template, matcher.cpp:788-841).  Same arithmetic, same order, same
// outputs (top-120 candidates of matcher.cpp:440-488 in std::sort order) - but every matrix lives in HBM / L2:
//
//   minu_big_sim_kernel     S = max(0, A.B^T), k ascending, unfused multiply and add          (:440-452)
//   minu_big_select_kernel  row / column sums in order, fp32 estimate -> histogram -> candidates -> the reference's
//                           double-precision value for candidates only -> bitonic sort; ties -> slow list (:455-488)
//   minu_big_slow_kernel    every value in double, libstdc++'s introsort replayed by one warp over keys and indices
//                           in HBM (32-bit indices: up to 2000 x 2000 entries)
//
// One CTA per job of a host-built work list; rare by construction (the host sizes the fast path so that at most
// 0.5 % of a gallery comes here), so these kernels are written for clarity, not for the last cycle.
#pragma once
#include "device_common.cuh"
#include "minu_sim.cuh"
#include "stdsort_emul.h"

namespace lafis {

struct MinuBigParams {
    // latent side
    const int* slot_n;
    const uint32_t* slot_off;
    const float* lat_desT;
    const int* lat_status;
    // gallery side
    const uint32_t* minu_off;
    const uint16_t* minu_n;
    const float* minu_desT;
    int g0, n_chunk;
    // work list: job = (q * n_chunk + tl) * 3 + slot, scratch offset of its matrices (in elements)
    const int* jobs;
    const unsigned long long* s_off;
    int n_jobs;
    float* S;         // [nL][np] per job
    uint32_t* keys;   // [nL * nR] per job (slow path)
    uint32_t* order;  // [nL * nR] per job (slow path)
    int max_nL, max_np;  // shared-memory geometry of the sums
    // outputs, indexed by job like the fast path's
    float* corr_v;
    uint32_t* corr_ij;
    int* corr_n;
    int* slow_count;
    int* slow_list;   // positions in the work list
    unsigned long long* replay_count;
};

struct BigJob {
    int job, q, tl, slot, nL, nR, np, npL;
};
__device__ __forceinline__ BigJob big_job(const MinuBigParams& P, int w) {
    BigJob b;
    b.job = P.jobs[w];
    b.slot = b.job % 3;
    const int pair = b.job / 3;
    b.tl = pair % P.n_chunk;
    b.q = pair / P.n_chunk;
    b.nR = P.minu_n[P.g0 + b.tl];
    b.nL = (P.lat_status[b.q] == 0) ? P.slot_n[b.q * 3 + b.slot] : 0;
    b.np = (b.nR + 3) & ~3;
    b.npL = (b.nL + 3) & ~3;
    return b;
}

__global__ void __launch_bounds__(256) minu_big_sim_kernel(MinuBigParams P) {
    const BigJob b = big_job(P, blockIdx.x);
    if (b.nL <= 0 || b.nR <= 0) return;
    const float* A = P.lat_desT + (size_t)96 * P.slot_off[b.q * 3 + b.slot];  // [96][npL]
    const float* B = P.minu_desT + (size_t)96 * P.minu_off[P.g0 + b.tl];       // [96][np]
    float* S = P.S + P.s_off[blockIdx.x];
    const int nq = b.np >> 2;
    for (int e = threadIdx.x; e < b.nL * nq; e += blockDim.x) {
        const int i = e / nq, j = (e - i * nq) * 4;
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int k = 0; k < 96; ++k) {
            const float a = __ldg(A + (size_t)k * b.npL + i);
            const float4 bv = __ldg(reinterpret_cast<const float4*>(B + (size_t)k * b.np + j));
            acc[0] = f_add(acc[0], f_mul(a, bv.x));
            acc[1] = f_add(acc[1], f_mul(a, bv.y));
            acc[2] = f_add(acc[2], f_mul(a, bv.z));
            acc[3] = f_add(acc[3], f_mul(a, bv.w));
        }
        float4 v;
        v.x = acc[0] < 0.0f ? 0.0f : acc[0];  // matcher.cpp:449-450
        v.y = acc[1] < 0.0f ? 0.0f : acc[1];
        v.z = acc[2] < 0.0f ? 0.0f : acc[2];
        v.w = acc[3] < 0.0f ? 0.0f : acc[3];
        *reinterpret_cast<float4*>(S + (size_t)i * b.np + j) = v;
    }
}

// floats of the row / column sums, rounded so that the histogram behind them (re-used for 64-bit sort keys) stays
// 16-byte aligned
__host__ __device__ inline int minu_big_sum_floats(int max_nL, int max_np) { return (max_nL + max_np + 3) & ~3; }
__host__ __device__ inline size_t minu_big_select_smem_bytes(int max_nL, int max_np) {
    return sizeof(float) * (size_t)minu_big_sum_floats(max_nL, max_np) + sizeof(int) * kSelBins + sizeof(int) * kSelMaxCand + 16;
}

// row / column sums of S in the reference's order (matcher.cpp:455-456 through the Eigen stand-in: ascending index)
__device__ __forceinline__ void big_sums(const float* __restrict__ S, int nL, int nR, int np, float* lsum, float* rsum) {
    for (int j = threadIdx.x; j < nR; j += blockDim.x) {
        float acc = S[j];
        for (int i = 1; i < nL; ++i) acc = f_add(acc, S[(size_t)i * np + j]);
        rsum[j] = acc;
    }
    for (int i = threadIdx.x; i < nL; i += blockDim.x) {
        const float* row = S + (size_t)i * np;
        float acc = row[0];
        for (int j = 1; j < nR; ++j) acc = f_add(acc, row[j]);
        lsum[i] = acc;
    }
}

__global__ void __launch_bounds__(kSelThreads) minu_big_select_kernel(MinuBigParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kSelThreads / 32;
    const BigJob b = big_job(P, blockIdx.x);
    const int nL = b.nL, nR = b.nR, np = b.np;
    if (nL <= 0 || nR <= 0) {
        if (tid == 0) P.corr_n[b.job] = 0;
        return;
    }
    float* lsum = reinterpret_cast<float*>(smem);          // [max_nL]
    float* rsum = lsum + P.max_nL;                          // [max_np]
    int* hist = reinterpret_cast<int*>(lsum + minu_big_sum_floats(P.max_nL, P.max_np));  // [1024]
    int* cand_e = hist + kSelBins;                          // [kSelMaxCand]
    __shared__ int s_ncand, s_npos, s_flag;
    __shared__ float s_thr;
    __shared__ int s_order[kTopCorrMinu];
    const float* S = P.S + P.s_off[blockIdx.x];

    for (int e = tid; e < kSelBins; e += kSelThreads) hist[e] = 0;
    if (tid == 0) {
        s_ncand = 0;
        s_npos = 0;
        s_flag = 0;
    }
    big_sums(S, nL, nR, np, lsum, rsum);
    __syncthreads();

    const long long M = (long long)nL * nR;
    const int K = M < kTopCorrMinu ? (int)M : kTopCorrMinu;
    // pass 1: histogram of the fp32 estimates (see minu_select_kernel for the margins)
    for (int i = warp; i < nL; i += NW) {
        const float l = lsum[i];
        for (int j = lane; j < nR; j += 32) {
            const float s = S[(size_t)i * np + j];
            if (s > 0.0f) {
                const uint32_t hb = __float_as_uint(approx_key(s, l, rsum[j])) >> 17;
                atomicAdd(&hist[hb > kSelBinBase ? min(hb - kSelBinBase, (uint32_t)(kSelBins - 1)) : 0u], 1);
            }
        }
    }
    __syncthreads();
    if (warp == 0) {
        int tot = 0;
        for (int e = lane; e < kSelBins; e += 32) tot += hist[e];
        tot = __reduce_add_sync(0xffffffffu, tot);
        int above = 0, bin = 0;
        if (tot >= K) {
            for (int c = kSelBins / 32 - 1; c >= 0; --c) {
                const int h = hist[c * 32 + lane];
                const int t = __reduce_add_sync(0xffffffffu, h);
                if (above + t >= K) {
                    int suf = h;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int o = __shfl_down_sync(0xffffffffu, suf, d);
                        if (lane + d < 32) suf += o;
                    }
                    const bool mine = above + (suf - h) < K && K <= above + suf;
                    const unsigned ball = __ballot_sync(0xffffffffu, mine);
                    bin = c * 32 + (31 - __clz(ball));
                    break;
                }
                above += t;
            }
        }
        if (lane == 0) {
            s_npos = tot;
            s_thr = bin > 0 ? __uint_as_float(((uint32_t)bin + kSelBinBase) << 17) * (1.0f - 4e-6f) : 0.0f;
        }
    }
    __syncthreads();
    if (s_npos < K) {  // the 120th value is a zero: ties among zeros decide the order
        if (tid == 0) P.slow_list[atomicAdd(P.slow_count, 1)] = blockIdx.x;
        return;
    }
    // pass 2: candidates
    {
        const float thr = s_thr;
        for (int i = warp; i < nL; i += NW) {
            const float l = lsum[i];
            for (int j = lane; j < nR; j += 32) {
                const float s = S[(size_t)i * np + j];
                if (s > 0.0f && approx_key(s, l, rsum[j]) >= thr) {
                    const int pos = atomicAdd(&s_ncand, 1);
                    if (pos < kSelMaxCand) cand_e[pos] = i * nR + j;
                }
            }
        }
    }
    __syncthreads();
    const int nc = s_ncand;
    if (nc > kSelMaxCand) {
        if (tid == 0) P.slow_list[atomicAdd(P.slow_count, 1)] = blockIdx.x;
        return;
    }
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(hist);  // 512 keys = the histogram's 4 KB
    int np2 = 128;
    while (np2 < nc) np2 <<= 1;
    __syncthreads();
    for (int c = tid; c < np2; c += kSelThreads) {
        unsigned long long k = 0ull;
        if (c < nc) {
            const int e = cand_e[c];
            const int i = e / nR, j = e - i * nR;
            const uint32_t key = exact_key(S[(size_t)i * np + j], lsum[i], rsum[j]);
            k = ((unsigned long long)key << 32) | (unsigned long long)(0xffffffffu - (uint32_t)e);
        }
        skey[c] = k;
    }
    __syncthreads();
    block_bitonic_desc<kSelThreads, (kSelMaxCand + kSelThreads - 1) / kSelThreads>(skey, np2);
    if (tid < K && tid + 1 < nc && (skey[tid] >> 32) == (skey[tid + 1] >> 32)) s_flag = 1;
    __syncthreads();
    if (s_flag) {
        if (tid == 0) P.slow_list[atomicAdd(P.slow_count, 1)] = blockIdx.x;
        return;
    }
    if (tid < K) s_order[tid] = (int)(0xffffffffu - (uint32_t)(skey[tid] & 0xffffffffull));
    __syncthreads();
    if (tid < K) {
        const int e = s_order[tid];
        const int i = e / nR, j = e - i * nR;
        P.corr_v[(size_t)b.job * kTopCorrMinu + tid] = S[(size_t)i * np + j];
        P.corr_ij[(size_t)b.job * kTopCorrMinu + tid] = ((uint32_t)i << 16) | (uint32_t)j;
    }
    if (tid == 0) P.corr_n[b.job] = K;
}

__global__ void __launch_bounds__(kSelThreads) minu_big_slow_kernel(MinuBigParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kSelThreads / 32;
    float* lsum = reinterpret_cast<float*>(smem);
    float* rsum = lsum + P.max_nL;
    const int n_slow = *P.slow_count;
    for (int sj = blockIdx.x; sj < n_slow; sj += gridDim.x) {
        const int w = P.slow_list[sj];
        const BigJob b = big_job(P, w);
        const int nL = b.nL, nR = b.nR, np = b.np;
        const float* S = P.S + P.s_off[w];
        uint32_t* keys = P.keys + P.s_off[w];
        uint32_t* y = P.order + P.s_off[w];
        __syncthreads();
        big_sums(S, nL, nR, np, lsum, rsum);
        __syncthreads();
        for (int i = warp; i < nL; i += NW)
            for (int j = lane; j < nR; j += 32) {
                const float s = S[(size_t)i * np + j];
                keys[(size_t)i * nR + j] = (s != 0.0f) ? exact_key(s, lsum[i], rsum[j]) : 0u;
            }
        __syncthreads();
        const int M = nL * nR;
        const int K = M < kTopCorrMinu ? M : kTopCorrMinu;
        if (warp == 0) {
            warp_std_sort_desc_prefix(DenseKey<uint32_t>{keys}, y, M, K);
            if (lane == 0) atomicAdd(P.replay_count, 1ull);
        }
        __syncthreads();
        if (tid < K) {
            const int e = (int)y[tid];
            const int i = e / nR, j = e - i * nR;
            P.corr_v[(size_t)b.job * kTopCorrMinu + tid] = S[(size_t)i * np + j];
            P.corr_ij[(size_t)b.job * kTopCorrMinu + tid] = ((uint32_t)i << 16) | (uint32_t)j;
        }
        if (tid == 0) P.corr_n[b.job] = K;
    }
}

}  // namespace lafis
