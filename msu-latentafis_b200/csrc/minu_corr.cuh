// Minutiae-template similarity, normalisation and top-120 candidate selection — stages K5 + K6 + K7.
//
//   reference: One2One_minutiae_matching steps 1-3, matching/matcher.cpp:440-488
//
// One CTA per gallery template.  Its k-major descriptor block B_T[96][nR] is staged once in shared
// memory and reused for every latent of the batch and each of the three selected latent templates
// (matcher.cpp:380, :400-408).  Per latent template:
//   K5  S = max(0, A.B^T): every output accumulates k = 0..95 in order with an unfused multiply and
//       add (the Eigen stand-in's order, oracle/shim/Eigen/Dense).  Warp tile 32x32, thread tile 8x4;
//       per k a warp issues three conflict-free LDS.128 (two broadcast A rows groups, one B group).
//   K6  column sums (i ascending) and row sums (j ascending) by one thread per column / row; the
//       normalised value S/(l_i + r_j - S + 1e-6) is evaluated in double exactly like :467 (the
//       literal is a double) and narrowed to float.  S is held with an odd row stride so that both
//       traversals are bank-conflict free.
//   K7  the 120 largest normalised values in std::sort order (:473-488): a three-pass radix select on
//       the float bit patterns finds the 120th value, the survivors are rank-sorted with the total
//       order (value desc, index asc).  If two survivors are equal, or the 120th value is tied with a
//       non-survivor, the permutation libstdc++'s introsort would produce is no longer implied by the
//       values, and thread 0 replays that introsort over all nL*nR indices (stdsort_emul.h).  The
//       output carries the RAW similarity (:486), recomputed from A_T/B_T in the same k order.
#pragma once
#include "device_common.cuh"
#include "stdsort_emul.h"

namespace lafis {

constexpr int kCorrThreads = 512;
constexpr int kHistBins = 4096;
constexpr int kCorrMaxDynSmem = 227 * 1024 - 2048;  // the kernel also holds ~1 KB of static shared memory

struct MinuCorrParams {
    // latent side
    const int* slot_n;          // [3Q]
    const uint32_t* slot_off;   // [3Q]
    const float* lat_desT;
    const int* lat_status;      // [Q]
    int Q;
    // gallery side
    const uint32_t* minu_off;
    const uint16_t* minu_n;
    const float* minu_desT;
    int g0, n_chunk;
    // shared-memory geometry chosen by the host: row counts rounded up to 32
    int nLs, nRs;
    // outputs, all [Q][n_chunk][3]
    float* corr_v;        // x120
    uint32_t* corr_ij;    // x120, (i << 16) | j
    int* corr_n;
    unsigned long long* slow_path_count;  // statistics
};

__host__ __device__ inline size_t minu_corr_smem_bytes(int nLs, int nRs) {
    // A_T [96][nLs] + B_T [96][nRs] + S/keys [nLs][nRs+1] + sums + histogram + survivors
    return sizeof(float) * ((size_t)96 * nLs + (size_t)96 * nRs + (size_t)nLs * (nRs + 1) + nLs + nRs) +
           sizeof(int) * kHistBins + (sizeof(uint32_t) + sizeof(int)) * 128 + 64;
}

// suffix-count search over a histogram held in shared memory, executed by warp 0:
// finds the highest bin b with (number of entries in bins > b) < need <= (entries in bins >= b).
// Returns b and writes the count of entries strictly above b.
__device__ __forceinline__ int find_bin_warp(const int* hist, int nbins, int need, int lane, int* above_out) {
    const int per = nbins / 32;
    int local = 0;
    for (int k = 0; k < per; ++k) local += hist[lane * per + k];
    // inclusive suffix sum over lanes: suf[lane] = sum_{l >= lane} local[l]
    int suf = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_down_sync(0xffffffffu, suf, d);
        if (lane + d < 32) suf += o;
    }
    const int above_lane = suf - local;  // entries in lanes > lane
    const bool mine = above_lane < need && need <= suf;
    const unsigned ball = __ballot_sync(0xffffffffu, mine);
    const int owner = ball ? (31 - __clz(ball)) : 0;
    int bin = 0, above = 0;
    if (lane == owner) {
        int run = above_lane;
        bin = lane * per;
        above = run;
        for (int k = per - 1; k >= 0; --k) {
            const int h = hist[lane * per + k];
            if (run + h >= need) {
                bin = lane * per + k;
                above = run;
                break;
            }
            run += h;
        }
    }
    bin = __shfl_sync(0xffffffffu, bin, owner);
    above = __shfl_sync(0xffffffffu, above, owner);
    *above_out = above;
    return bin;
}

__global__ void __launch_bounds__(kCorrThreads, 1) minu_corr_kernel(MinuCorrParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int nLs = P.nLs, nRs = P.nRs, ldS = nRs + 1;
    float* A_T = reinterpret_cast<float*>(smem);          // [96][nLs]
    float* B_T = A_T + 96 * nLs;                           // [96][nRs]
    float* S = B_T + 96 * nRs;                             // [nLs][ldS]  raw similarity, then keys
    float* lsum = S + (size_t)nLs * ldS;                   // [nLs]
    float* rsum = lsum + nLs;                              // [nRs]
    int* hist = reinterpret_cast<int*>(rsum + nRs);        // [4096]
    uint32_t* cand_key = reinterpret_cast<uint32_t*>(hist + kHistBins);  // [128]
    int* cand_idx = reinterpret_cast<int*>(cand_key + 128);              // [128]
    __shared__ int s_cnt, s_flag, s_bin, s_above;
    __shared__ int s_order[128];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kCorrThreads / 32;
    const int tl = blockIdx.x;
    const int g = P.g0 + tl;
    const int nR = P.minu_n[g];
    uint32_t* keys = reinterpret_cast<uint32_t*>(S);

    // ---- stage B_T (zero padded to nRs columns) ----
    auto stage_B = [&]() {
        if (nR <= 0) return;
        const int nRp = (nR + 3) & ~3;
        const float* src = P.minu_desT + (size_t)96 * P.minu_off[g];
        for (int e = tid; e < 96 * nRs; e += kCorrThreads) {
            const int k = e / nRs, j = e - k * nRs;
            B_T[e] = (j < nR) ? __ldg(src + (size_t)k * nRp + j) : 0.0f;
        }
    };
    auto stage_A = [&](int q, int slot, int nL) {
        const int nLp = (nL + 3) & ~3;
        const float* src = P.lat_desT + (size_t)96 * P.slot_off[q * 3 + slot];
        for (int e = tid; e < 96 * nLs; e += kCorrThreads) {
            const int k = e / nLs, i = e - k * nLs;
            A_T[e] = (i < nL) ? __ldg(src + (size_t)k * nLp + i) : 0.0f;
        }
    };
    stage_B();

    for (int q = 0; q < P.Q; ++q) {
        for (int slot = 0; slot < 3; ++slot) {
            const size_t oidx = ((size_t)q * P.n_chunk + tl) * 3 + slot;
            const int nL = (P.lat_status[q] == 0) ? P.slot_n[q * 3 + slot] : 0;
            if (nL <= 0 || nR <= 0) {  // rolled without minutiae / latent slot absent: score stays 0
                if (tid == 0) P.corr_n[oidx] = 0;
                continue;
            }
            __syncthreads();  // previous slot done with A_T / S
            stage_A(q, slot, nL);
            __syncthreads();

            // ---- K5 ----
            const int tiles_i = (nL + 31) >> 5, tiles_j = (nR + 31) >> 5;
            const int li = lane >> 3, lj = lane & 7;
            for (int wt = warp; wt < tiles_i * tiles_j; wt += NW) {
                const int i0 = (wt / tiles_j) * 32 + li * 8, j0 = (wt % tiles_j) * 32 + lj * 4;
                float acc[8][4];
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
                const float* ap = A_T + i0;
                const float* bp = B_T + j0;
#pragma unroll 4
                for (int k = 0; k < 96; ++k) {
                    const float4 a0 = *reinterpret_cast<const float4*>(ap + k * nLs);
                    const float4 a1 = *reinterpret_cast<const float4*>(ap + k * nLs + 4);
                    const float4 b4 = *reinterpret_cast<const float4*>(bp + k * nRs);
                    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                    for (int a = 0; a < 8; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc[a][b] = f_add(acc[a][b], f_mul(av[a], bv[b]));
                }
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        float v = acc[a][b];
                        if (v < 0.0f) v = 0.0f;  // matcher.cpp:449-450
                        S[(size_t)(i0 + a) * ldS + j0 + b] = v;
                    }
            }
            __syncthreads();

            // ---- K6: sums ----
            if (tid < nR) {  // column sums, i ascending
                float acc = S[tid];
                for (int i = 1; i < nL; ++i) acc = f_add(acc, S[(size_t)i * ldS + tid]);
                rsum[tid] = acc;
            } else if (tid >= 256 && tid - 256 < nL) {  // row sums, j ascending
                const float* row = S + (size_t)(tid - 256) * ldS;
                float acc = row[0];
                for (int j = 1; j < nR; ++j) acc = f_add(acc, row[j]);
                lsum[tid - 256] = acc;
            }
            for (int b = tid; b < kHistBins; b += kCorrThreads) hist[b] = 0;
            if (tid == 0) {
                s_cnt = 0;
                s_flag = 0;
            }
            __syncthreads();

            // ---- K6: normalised keys (in place) + first histogram ----
            const int M = nL * nR;
            const int K = M < kTopCorrMinu ? M : kTopCorrMinu;
            for (int e = tid; e < M; e += kCorrThreads) {
                const int i = e / nR, j = e - i * nR;
                const float s = S[(size_t)i * ldS + j];
                uint32_t key = 0;
                if (s != 0.0f) {
                    const float den = f_sub(f_add(lsum[i], rsum[j]), s);
                    const double qd = (double)s / ((double)den + 0.000001);
                    key = __float_as_uint((float)qd);
                    if (key == 0x80000000u) key = 0;
                }
                keys[(size_t)i * ldS + j] = key;
                atomicAdd(&hist[key >> 20], 1);
            }
            __syncthreads();

            // ---- K7: radix select of the K-th largest key ----
            uint32_t prefix = 0;
            int above = 0;
            {
                if (warp == 0) {
                    int ab;
                    const int b = find_bin_warp(hist, kHistBins, K, lane, &ab);
                    if (lane == 0) {
                        s_bin = b;
                        s_above = ab;
                    }
                }
                __syncthreads();
                prefix = (uint32_t)s_bin << 20;
                above = s_above;
                __syncthreads();
                for (int b = tid; b < 1024; b += kCorrThreads) hist[b] = 0;
                __syncthreads();
                for (int e = tid; e < M; e += kCorrThreads) {
                    const int i = e / nR, j = e - i * nR;
                    const uint32_t key = keys[(size_t)i * ldS + j];
                    if ((key >> 20) == (prefix >> 20)) atomicAdd(&hist[(key >> 10) & 1023], 1);
                }
                __syncthreads();
                if (warp == 0) {
                    int ab;
                    const int b = find_bin_warp(hist, 1024, K - above, lane, &ab);
                    if (lane == 0) {
                        s_bin = b;
                        s_above = above + ab;
                    }
                }
                __syncthreads();
                prefix |= (uint32_t)s_bin << 10;
                above = s_above;
                __syncthreads();
                for (int b = tid; b < 1024; b += kCorrThreads) hist[b] = 0;
                __syncthreads();
                for (int e = tid; e < M; e += kCorrThreads) {
                    const int i = e / nR, j = e - i * nR;
                    const uint32_t key = keys[(size_t)i * ldS + j];
                    if ((key >> 10) == (prefix >> 10)) atomicAdd(&hist[key & 1023], 1);
                }
                __syncthreads();
                if (warp == 0) {
                    int ab;
                    const int b = find_bin_warp(hist, 1024, K - above, lane, &ab);
                    if (lane == 0) {
                        s_bin = b;
                        s_above = above + ab;
                    }
                }
                __syncthreads();
                prefix |= (uint32_t)s_bin;
                above = s_above;
            }
            const uint32_t T = prefix;                 // the K-th largest key
            const int n_eq = hist[T & 1023];           // keys equal to T
            const bool boundary_tie = (above + n_eq) != K;

            // ---- survivors ----
            if (!boundary_tie) {
                for (int e = tid; e < M; e += kCorrThreads) {
                    const int i = e / nR, j = e - i * nR;
                    const uint32_t key = keys[(size_t)i * ldS + j];
                    if (key >= T) {
                        const int pos = atomicAdd(&s_cnt, 1);
                        if (pos < 128) {
                            cand_key[pos] = key;
                            cand_idx[pos] = e;
                        }
                    }
                }
                __syncthreads();
                if (tid < K) {  // rank sort with the total order (key desc, index asc)
                    const uint32_t mk = cand_key[tid];
                    const int me = cand_idx[tid];
                    int rank = 0;
                    bool tie = false;
                    for (int d = 0; d < K; ++d) {
                        const uint32_t ok = cand_key[d];
                        const int oe = cand_idx[d];
                        rank += (ok > mk) || (ok == mk && oe < me);
                        tie |= (ok == mk && d != tid);
                    }
                    s_order[rank] = me;
                    if (tie) s_flag = 1;
                }
            }
            __syncthreads();
            if (boundary_tie || s_flag) {
                // Replay libstdc++'s introsort (rare).  The index array (u16, M < 65536) borrows the
                // A_T/B_T staging area, which is re-staged afterwards.
                if (tid == 0) {
                    uint16_t* y = reinterpret_cast<uint16_t*>(A_T);
                    const uint32_t* kk = keys;
                    const int nRr = nR, ld = ldS;
                    auto keyfn = [kk, nRr, ld](int e) -> uint32_t {
                        const int i = e / nRr;
                        return kk[i * ld + (e - i * nRr)];
                    };
                    std_sort_desc_prefix(keyfn, y, M, K);
                    for (int r = 0; r < K; ++r) s_order[r] = (int)y[r];
                    atomicAdd(P.slow_path_count, 1ull);
                }
                __syncthreads();
                stage_B();
                stage_A(q, slot, nL);
                __syncthreads();
            }

            // ---- output: raw similarity recomputed in reference order ----
            if (tid < K) {
                const int e = s_order[tid];
                const int i = e / nR, j = e - i * nR;
                float acc = 0.0f;
                for (int k = 0; k < 96; ++k) acc = f_add(acc, f_mul(A_T[k * nLs + i], B_T[k * nRs + j]));
                if (acc < 0.0f) acc = 0.0f;
                P.corr_v[oidx * kTopCorrMinu + tid] = acc;
                P.corr_ij[oidx * kTopCorrMinu + tid] = ((uint32_t)i << 16) | (uint32_t)j;
            }
            if (tid == 0) P.corr_n[oidx] = K;
        }
    }
}

}  // namespace lafis
