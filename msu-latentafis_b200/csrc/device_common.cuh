// Shared device-side declarations: packed gallery / latent layouts in HBM and small helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "exact_math.h"

namespace lafis {

constexpr int kDesLenD = 96;
constexpr int kTopCorrMinu = 120;  // matcher.cpp:479 topN
constexpr int kTopCorrTex = 200;   // matcher.cpp:33 N
constexpr int kTableN = 50;        // matcher.cpp:45 dist_N

// Gallery resident in HBM (structure of arrays, see DESIGN.md "Data layout").
struct DeviceGallery {
    int n = 0;
    // minutiae template 0 of every rolled print
    uint32_t* minu_off = nullptr;   // [n+1] offsets in minutiae units, every template padded to x4
    uint16_t* minu_n = nullptr;     // [n]
    short2* minu_xy = nullptr;      // [tot_minu_padded] pixels
    float* minu_ori = nullptr;      // [tot_minu_padded]
    float* minu_desT = nullptr;     // template g: [96][np_g] at 96*minu_off[g], np_g = padded count
    // texture template 0 of every rolled print
    uint32_t* tex_off = nullptr;    // [n+1]
    short2* tex_xy = nullptr;       // [tot_tex] block units
    float* tex_ori = nullptr;       // [tot_tex]
    uint4* tex_codes = nullptr;     // [tot_tex + 16] 16 PQ codes per point
    int8_t* status = nullptr;       // [n]
};

// A latent batch resident in HBM.
struct DeviceLatents {
    int n = 0;
    int lt_stride = 0;              // padded max texture points per latent (multiple of 8)
    int* slot_n = nullptr;          // [3n] minutiae per selected template slot, 0 when absent
    uint32_t* slot_off = nullptr;   // [3n] padded offsets (multiples of 4)
    short2* minu_xy = nullptr;
    float* minu_ori = nullptr;
    float* minu_desT = nullptr;     // slot s: [96][np_s] at 96*slot_off[s]
    int* tex_n = nullptr;           // [n] texture points used (<= 1000), 0 when no texture template
    short2* tex_xy = nullptr;       // [n][lt_stride]
    float* tex_ori = nullptr;       // [n][lt_stride]
    float* tex_des = nullptr;       // [n][lt_stride][96] row-major, zero padded
    int* tex_weighted = nullptr;    // [n] 1 when score[28] is the texture score (28 minutiae templates)
    int* status = nullptr;          // [n] LAFIS_OK / LAFIS_LATENT_EMPTY / LAFIS_ERR_LATENT_LAYOUT
};

constexpr int kMergeCap = 4096;    // rank-list entries one merge CTA sorts (= kTopkChunk, misc_kernels.cuh)

// one rank-list entry in HBM; same layout as lafis_hit
struct HitDev {
    float score;
    uint32_t index;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Per-(latent, template) kernels run on a grid of (jobs per latent, latents): blockIdx.x is the job inside the latent,
// the latent is blockIdx.y, with blockIdx.z carrying batches beyond the 65,535 rows of a grid.
constexpr int kGridLatents = 32768;
inline dim3 job_grid(unsigned jobs_per_latent, int Q) {
    return dim3(jobs_per_latent, (unsigned)(Q < kGridLatents ? Q : kGridLatents), (unsigned)((Q + kGridLatents - 1) / kGridLatents));
}
__device__ __forceinline__ int job_latent() { return (int)(blockIdx.z * (unsigned)kGridLatents + blockIdx.y); }

// Block-wide bitonic sort of np2 (a power of two, >= 32, <= NT * KPT) 64-bit keys in shared memory, descending.
// Compare-exchange distances below 32 stay inside a warp: each thread keeps its key(s) in registers and
// exchanges them by shuffle, so only the distances >= 32 go through shared memory and a block barrier
// (6 barriers instead of 28 for 128 keys).  Every thread of the block must call; ends with a barrier.
template <int NT, int KPT>
__device__ __forceinline__ void block_bitonic_desc(unsigned long long* skey, int np2) {
    const int tid = threadIdx.x;
    unsigned long long mine[KPT];
    auto load_keys = [&]() {
#pragma unroll
        for (int e = 0; e < KPT; ++e) {
            const int i = tid + e * NT;
            mine[e] = (i < np2) ? skey[i] : 0ull;
        }
    };
    auto store_keys = [&]() {
#pragma unroll
        for (int e = 0; e < KPT; ++e) {
            const int i = tid + e * NT;
            if (i < np2) skey[i] = mine[e];
        }
    };
    auto exchange = [&](int k, int j) {  // one compare-exchange step at distance j < 32 of merge size k
#pragma unroll
        for (int e = 0; e < KPT; ++e) {
            const int i = tid + e * NT;
            if ((i & ~31) >= np2) continue;  // warp-uniform: this warp holds no key of slot e
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, mine[e], j);
            const bool take_max = ((i & j) == 0) == ((i & k) == 0);
            mine[e] = ((mine[e] > other) == take_max) ? mine[e] : other;
        }
    };
    load_keys();
    for (int k = 2; k <= 32; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) exchange(k, j);
    store_keys();
    __syncthreads();
    for (int k = 64; k <= np2; k <<= 1) {
        for (int j = k >> 1; j >= 32; j >>= 1) {
            for (int t = tid; t < np2 / 2; t += NT) {
                const int lo = ((t / j) * (j << 1)) + (t % j), hi = lo + j;
                const bool desc = ((lo & k) == 0);
                const unsigned long long a = skey[lo], b = skey[hi];
                if ((a < b) == desc) {
                    skey[lo] = b;
                    skey[hi] = a;
                }
            }
            __syncthreads();
        }
        load_keys();
        for (int j = 16; j > 0; j >>= 1) exchange(k, j);
        store_keys();
        __syncthreads();
    }
}

// Block-wide descending rank of n 32-bit keys by ONE bucket pass instead of a sorting network: the keys fall into NB
// buckets that are uniform in the key's integer value between the block's smallest and largest key (for sortable float
// bit patterns: linear inside a binade, logarithmic across binades - narrow clusters and heavy tails both spread), a
// suffix scan over the bucket counts gives every bucket's first rank, and a key's rank inside its bucket comes from
// comparing it with the bucket's other keys (about n / NB of them).  O(n) instructions and 5 barriers, against
// O(n log^2 n) compare-exchanges and a barrier per long-distance stage of the bitonic network.
//   key e of thread tid has id = tid + e * NT and takes part when id < n.
//   rank[e]  = position in the order (key descending, id ascending)
//   first[e] = number of keys strictly larger (the first rank of the key's group of equal keys)
//   tied[e]  = another key has the same value
// Shared memory: start[NB + 1], buck[n], red[2 * NT / 32].  Every thread of the block must call.
template <int NT, int KPT, int NB>
__device__ __forceinline__ void block_rank_desc(const uint32_t (&u)[KPT], int n, int* __restrict__ start, uint2* __restrict__ buck,
                                                uint32_t* __restrict__ red, int (&rank)[KPT], int (&first)[KPT], bool (&tied)[KPT]) {
    static_assert(NB % NT == 0 && (NB & (NB - 1)) == 0, "whole buckets per thread");
    constexpr int NW = NT / 32, BPT = NB / NT;
    constexpr int LOG_NB = 31 - __builtin_clz((unsigned)NB);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t lo = 0xffffffffu, hi = 0u;
#pragma unroll
    for (int e = 0; e < KPT; ++e)
        if (tid + e * NT < n) {
            lo = min(lo, u[e]);
            hi = max(hi, u[e]);
        }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if (lane == 0) {
        red[warp] = lo;
        red[NW + warp] = hi;
    }
    for (int b = tid; b <= NB; b += NT) start[b] = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        lo = min(lo, red[k]);
        hi = max(hi, red[NW + k]);
    }
    const uint32_t span = hi - lo;
    const int sh = span < (uint32_t)NB ? 0 : (32 - __clz(span)) - LOG_NB;  // (span >> sh) < NB
    int bin[KPT], slot[KPT];
#pragma unroll
    for (int e = 0; e < KPT; ++e) {
        bin[e] = 0;
        slot[e] = 0;
        if (tid + e * NT < n) {
            bin[e] = (int)((u[e] - lo) >> sh);
            slot[e] = atomicAdd(&start[bin[e]], 1);
        }
    }
    __syncthreads();
    // suffix scan in place: start[b + 1] = number of keys in buckets above b, start[0] = n
    int h[BPT], tot = 0;
#pragma unroll
    for (int k = 0; k < BPT; ++k) {
        h[k] = start[NB - 1 - (tid * BPT + k)];
        tot += h[k];
    }
    int inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) red[warp] = (uint32_t)inc;  // (the minima / maxima in red were consumed before the last barrier)
    __syncthreads();
    int run = inc - tot;
#pragma unroll
    for (int k = 0; k < NW; ++k) run += (k < warp) ? (int)red[k] : 0;
#pragma unroll
    for (int k = 0; k < BPT; ++k) {
        start[NB - (tid * BPT + k)] = run;  // bucket b = NB - 1 - (tid * BPT + k): entry b + 1
        run += h[k];
    }
    if (tid == NT - 1) start[0] = run;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < KPT; ++e)
        if (tid + e * NT < n) buck[start[bin[e] + 1] + slot[e]] = make_uint2(u[e], (uint32_t)(tid + e * NT));
    __syncthreads();
#pragma unroll
    for (int e = 0; e < KPT; ++e) {
        rank[e] = first[e] = 0;
        tied[e] = false;
        if (tid + e * NT < n) {
            const int base = start[bin[e] + 1], cnt = start[bin[e]] - base;
            const uint32_t id = (uint32_t)(tid + e * NT);
            int gt = 0, before = 0;
#pragma unroll 1
            for (int p = 0; p < cnt; ++p) {
                const uint2 o = buck[base + p];
                gt += o.x > u[e];
                before += (o.x == u[e]) & (o.y < id);
                tied[e] |= (o.x == u[e]) & (o.y != id);
            }
            first[e] = base + gt;
            rank[e] = base + gt + before;
        }
    }
}

}  // namespace lafis
