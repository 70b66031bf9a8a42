// Score fusion, rank-list selection, PQ encoding and gallery re-layout kernels.
//
//   reference: score fusion                      matching/matcher.cpp:188 / :293            (K10)
//              rank list                         matching/matcher.cpp:306-309               (K11)
//              TrainedPQEncoder.encode_multi     extraction/descriptor_PQ.py:19-27          (§8f.3)
#pragma once
#include <math_constants.h>

#include "device_common.cuh"

namespace lafis {

// ---- K10: final = score[0] + score[1] + score[2] + score[28]*0.3 -------------------------------
// The three float additions happen in float, "score[28]*0.3" and the last addition in double
// (0.3 is a double literal), the result is narrowed to float.  Slots the reference never writes
// keep -1: latent not matchable (One2One_matching_selected_templates returns 1) or rolled template
// empty / failed to load (returns 2, matcher.cpp:173-177, :184-187).
struct FuseParams {
    float* comp;               // [Q][G][4]: minutiae slots 0..2 and the texture score, rewritten as score[0], score[1], score[2], score[28]
    const int* lat_status;     // [Q]
    const int* tex_weighted;   // [Q] 0: texture score not read / 1: it is score[28] (weight 0.3) / 2 + k: it is score[k], k = 0..2 (weight 1)
    const int8_t* gal_status;  // [G]
    int Q, G;
    float* final_scores;       // [Q][G]
};

__global__ void fuse_kernel(FuseParams P) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)P.Q * P.G) return;
    const int q = (int)(e / P.G), g = (int)(e % P.G);
    float out = -1.0f;
    const int8_t gs = P.gal_status[g];
    if (P.lat_status[q] == 0 && (gs == 0 || gs == 2)) {
        float4 c = *reinterpret_cast<const float4*>(P.comp + e * 4);
        const int mode = P.tex_weighted[q];
        if (mode != 1) {
            // <= 2 minutiae templates: the texture score IS score[n_minu_templates] (matcher.cpp:414), one of the three
            // unweighted terms (the minutiae scores are all 0 then); any other count: it is never read
            const float t = c.w;
            c.w = 0.0f;
            if (mode == 2) c.x = t;
            else if (mode == 3) c.y = t;
            else if (mode == 4) c.z = t;
            *reinterpret_cast<float4*>(P.comp + e * 4) = c;
        }
        const float s012 = f_add(f_add(c.x, c.y), c.z);
        out = (float)((double)s012 + (double)c.w * 0.3);
    }
    P.final_scores[e] = out;
}

// ---- K11: rank lists -----------------------------------------------------------------------------
// Rank order is (score descending, gallery index ascending) on the unrounded fp32 scores
// (SURVEY.md §8d).  Both are folded into one 64-bit key whose unsigned order is the rank order.
__device__ __forceinline__ unsigned long long rank_key(float score, uint32_t index) {
    uint32_t u = __float_as_uint(score);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned long long)(~index);
}
__device__ __forceinline__ void rank_unkey(unsigned long long k, float* score, uint32_t* index) {
    uint32_t u = (uint32_t)(k >> 32);
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    *score = __uint_as_float(u);
    *index = ~(uint32_t)(k & 0xffffffffull);
}

constexpr int kTopkThreads = 512;
constexpr int kTopkChunk = 4096;
static_assert(kMergeCap == kTopkChunk, "the multi-GPU merge sorts one chunk");

// in-place bitonic sort (descending) of kTopkChunk keys in shared memory
__device__ __forceinline__ void bitonic_desc(unsigned long long* s) {
    for (int k = 2; k <= kTopkChunk; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < kTopkChunk / 2; t += kTopkThreads) {
                const int lo = ((t / j) * (j << 1)) + (t % j), hi = lo + j;
                const bool desc = ((lo & k) == 0);
                const unsigned long long a = s[lo], b = s[hi];
                if ((a < b) == desc) {
                    s[lo] = b;
                    s[hi] = a;
                }
            }
            __syncthreads();
        }
    }
}

// level 0: keys from scores.  grid = (chunks, Q).  out[q][chunk][k]
__global__ void __launch_bounds__(kTopkThreads) topk_scores_kernel(const float* scores, int G, uint32_t index_base,
                                                                  int k, unsigned long long* out) {
    __shared__ unsigned long long s[kTopkChunk];
    const int q = blockIdx.y, chunk = blockIdx.x;
    const int base = chunk * kTopkChunk;
    for (int t = threadIdx.x; t < kTopkChunk; t += kTopkThreads) {
        const int g = base + t;
        s[t] = (g < G) ? rank_key(scores[(size_t)q * G + g], index_base + (uint32_t)g) : 0ull;
    }
    __syncthreads();
    bitonic_desc(s);
    for (int t = threadIdx.x; t < k; t += kTopkThreads) out[((size_t)q * gridDim.x + chunk) * k + t] = s[t];
}

// level >= 1: keys from keys.  in[q][n_in] -> out[q][chunk][k]
__global__ void __launch_bounds__(kTopkThreads) topk_keys_kernel(const unsigned long long* in, int n_in, int k,
                                                                unsigned long long* out) {
    __shared__ unsigned long long s[kTopkChunk];
    const int q = blockIdx.y, chunk = blockIdx.x;
    const int base = chunk * kTopkChunk;
    for (int t = threadIdx.x; t < kTopkChunk; t += kTopkThreads) {
        const int g = base + t;
        s[t] = (g < n_in) ? in[(size_t)q * n_in + g] : 0ull;
    }
    __syncthreads();
    bitonic_desc(s);
    for (int t = threadIdx.x; t < k; t += kTopkThreads) out[((size_t)q * gridDim.x + chunk) * k + t] = s[t];
}

__global__ void keys_to_hits_kernel(const unsigned long long* keys, size_t n, HitDev* hits) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const unsigned long long k = keys[e];
    HitDev h;
    if (k == 0ull) {
        h.score = -CUDART_INF_F;
        h.index = 0xffffffffu;
    } else {
        rank_unkey(k, &h.score, &h.index);
    }
    hits[e] = h;
}

// merge of gathered per-shard rank lists: in[list][latent][k] (the layout an all-gather of per-rank
// [latent][k] blocks produces) -> out[latent][k].  One CTA per latent, n_lists * k <= kTopkChunk.
__global__ void __launch_bounds__(kTopkThreads) merge_hits_kernel(const HitDev* in, int n_latents, int n_lists, int k,
                                                                 HitDev* out) {
    __shared__ unsigned long long s[kTopkChunk];
    const int q = blockIdx.x;
    const int tot = n_lists * k;
    for (int t = threadIdx.x; t < kTopkChunk; t += kTopkThreads) {
        unsigned long long key = 0ull;
        if (t < tot) {
            const HitDev h = in[((size_t)(t / k) * n_latents + q) * k + (t % k)];
            if (h.index != 0xffffffffu) key = rank_key(h.score, h.index);
        }
        s[t] = key;
    }
    __syncthreads();
    bitonic_desc(s);
    for (int t = threadIdx.x; t < k; t += kTopkThreads) {
        HitDev h;
        if (s[t] == 0ull) {
            h.score = -CUDART_INF_F;
            h.index = 0xffffffffu;
        } else {
            rank_unkey(s[t], &h.score, &h.index);
        }
        out[(size_t)q * k + t] = h;
    }
}

// ---- PQ encoder ----------------------------------------------------------------------------------
// One thread per (point, sub-quantizer): nearest of 256 centroids in 6-d, first minimum wins
// (scipy.cluster.vq.vq).  fp32 squared distances accumulated in dimension order.
__global__ void pq_encode_kernel(const float* des, long long n, const float* codebook, uint8_t* codes) {
    __shared__ float cw[16 * 256 * 6 / 4];  // one quarter of the codebook per pass (24 KB)
    const long long point = (long long)blockIdx.x * (blockDim.x / 16) + threadIdx.x / 16;
    const int m = threadIdx.x & 15;
    float d[6] = {0, 0, 0, 0, 0, 0};
    if (point < n)
        for (int k = 0; k < 6; ++k) d[k] = des[point * kDesLenD + m * 6 + k];
    int best = 0;
    for (int pass = 0; pass < 4; ++pass) {  // sub-quantizers 4*pass .. 4*pass+3
        __syncthreads();
        for (int e = threadIdx.x; e < 4 * 256 * 6; e += blockDim.x) cw[e] = codebook[pass * 4 * 256 * 6 + e];
        __syncthreads();
        if ((m >> 2) == pass && point < n) {
            const float* w = cw + (m & 3) * 256 * 6;
            float bd = CUDART_INF_F;
            for (int c = 0; c < 256; ++c) {
                float dist = 0.0f;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const float t = f_sub(d[k], w[c * 6 + k]);
                    dist = f_add(dist, f_mul(t, t));
                }
                if (dist < bd) {
                    bd = dist;
                    best = c;
                }
            }
        }
    }
    if (point < n) codes[point * 16 + m] = (uint8_t)best;
}

// ---- gallery re-layout (ingest) --------------------------------------------------------------------
// row-major descriptors [n][96] of template g -> k-major [96][np] (np = n rounded up to 4, zero
// padded); coordinates -> short2.  grid = templates.
struct RelayoutParams {
    int n_templates;
    const uint32_t* src_off;   // [n+1] unpadded offsets
    const uint32_t* dst_off;   // [n+1] padded offsets
    const int16_t* x;
    const int16_t* y;
    const float* ori;
    const float* des;          // row-major
    short2* out_xy;
    float* out_ori;
    float* out_desT;
};

__global__ void relayout_minutiae_kernel(RelayoutParams P) {
    const int g = blockIdx.x;
    const uint32_t s0 = P.src_off[g], n = P.src_off[g + 1] - s0;
    const uint32_t d0 = P.dst_off[g], np = P.dst_off[g + 1] - d0;
    for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) {
        if (i < n) {
            P.out_xy[d0 + i] = make_short2(P.x[s0 + i], P.y[s0 + i]);
            P.out_ori[d0 + i] = P.ori[s0 + i];
        } else {
            P.out_xy[d0 + i] = make_short2(0, 0);
            P.out_ori[d0 + i] = 0.0f;
        }
    }
    float* dst = P.out_desT + (size_t)96 * d0;
    const float* src = P.des + (size_t)96 * s0;
    for (uint32_t e = threadIdx.x; e < 96 * np; e += blockDim.x) {
        const uint32_t k = e / np, i = e - k * np;
        dst[e] = (i < n) ? src[(size_t)i * 96 + k] : 0.0f;
    }
}

// texture template g: points [src_off[g], src_off[g] + n) with n = dst_off[g+1] - dst_off[g]
// (<= 1000, matcher.cpp:546-547) -> packed coordinates, orientations and 16-byte code words.
struct TexCopyParams {
    const uint32_t* src_off;
    const uint32_t* dst_off;
    const int16_t* x;
    const int16_t* y;
    const float* ori;
    const uint4* codes;
    bool codes_aligned;
    short2* out_xy;
    float* out_ori;
    uint4* out_codes;
};

__global__ void copy_texture_kernel(TexCopyParams P) {
    const int g = blockIdx.x;
    const uint32_t s0 = P.src_off[g], d0 = P.dst_off[g], n = P.dst_off[g + 1] - d0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        P.out_xy[d0 + i] = make_short2(P.x[s0 + i], P.y[s0 + i]);
        P.out_ori[d0 + i] = P.ori[s0 + i];
        uint4 c;
        if (P.codes_aligned) {
            c = P.codes[s0 + i];
        } else {
            const uint8_t* b = reinterpret_cast<const uint8_t*>(P.codes) + (size_t)(s0 + i) * 16;
            uint32_t w[4];
            for (int k = 0; k < 4; ++k)
                w[k] = (uint32_t)b[4 * k] | ((uint32_t)b[4 * k + 1] << 8) | ((uint32_t)b[4 * k + 2] << 16) |
                       ((uint32_t)b[4 * k + 3] << 24);
            c = make_uint4(w[0], w[1], w[2], w[3]);
        }
        P.out_codes[d0 + i] = c;
    }
}

// ---- ingest from the parser threads' device pool -----------------------------------------------------
// lafis_gallery_load_files: every parser thread ships blocks of parsed templates through pinned ring buffers into one
// device pool; a template's pieces lie back to back at 16-byte aligned offsets:
//   minutiae: x[n] i16 | y[n] i16 | ori[n] f32 | des[n][96] f32 (row-major, as in the file)
//   texture : x[n] i16 | y[n] i16 | ori[n] f32 | codes[n][16] u8
// The two kernels below produce the resident layout (k-major descriptors, short2 coordinates, uint4 code words)
// straight from the pool.  grid = templates.
struct PoolRec {
    unsigned long long minu_at, tex_at;  // byte offsets into the pool
    uint32_t n_minu, n_tex;              // points kept (texture: <= 1000, matcher.cpp:546-547)
};
__host__ __device__ inline size_t pool_align(size_t v) { return (v + 15) & ~(size_t)15; }
__host__ __device__ inline size_t pool_minu_bytes(size_t n) { return pool_align(2 * n) * 2 + pool_align(4 * n) + pool_align(4 * 96 * n); }
__host__ __device__ inline size_t pool_tex_bytes(size_t n) { return pool_align(2 * n) * 2 + pool_align(4 * n) + pool_align(16 * n); }

struct PoolRelayoutParams {
    const unsigned char* pool;
    const PoolRec* rec;
    const uint32_t* minu_off;  // resident (padded) offsets
    const uint32_t* tex_off;
    short2* minu_xy;
    float* minu_ori;
    float* minu_desT;
    short2* tex_xy;
    float* tex_ori;
    uint4* tex_codes;
};

__global__ void relayout_minutiae_pool_kernel(PoolRelayoutParams P) {
    const int g = blockIdx.x;
    const PoolRec r = P.rec[g];
    const uint32_t n = r.n_minu, d0 = P.minu_off[g], np = P.minu_off[g + 1] - d0;
    const unsigned char* base = P.pool + r.minu_at;
    const int16_t* x = reinterpret_cast<const int16_t*>(base);
    const int16_t* y = reinterpret_cast<const int16_t*>(base + pool_align(2 * (size_t)n));
    const float* ori = reinterpret_cast<const float*>(base + 2 * pool_align(2 * (size_t)n));
    const float* des = reinterpret_cast<const float*>(base + 2 * pool_align(2 * (size_t)n) + pool_align(4 * (size_t)n));
    for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) {
        P.minu_xy[d0 + i] = i < n ? make_short2(x[i], y[i]) : make_short2(0, 0);
        P.minu_ori[d0 + i] = i < n ? ori[i] : 0.0f;
    }
    float* dst = P.minu_desT + (size_t)96 * d0;
    for (uint32_t e = threadIdx.x; e < 96 * np; e += blockDim.x) {
        const uint32_t k = e / np, i = e - k * np;
        dst[e] = (i < n) ? des[(size_t)i * 96 + k] : 0.0f;
    }
}

__global__ void copy_texture_pool_kernel(PoolRelayoutParams P) {
    const int g = blockIdx.x;
    const PoolRec r = P.rec[g];
    const uint32_t n = r.n_tex, d0 = P.tex_off[g];
    const unsigned char* base = P.pool + r.tex_at;
    const int16_t* x = reinterpret_cast<const int16_t*>(base);
    const int16_t* y = reinterpret_cast<const int16_t*>(base + pool_align(2 * (size_t)n));
    const float* ori = reinterpret_cast<const float*>(base + 2 * pool_align(2 * (size_t)n));
    const uint4* codes = reinterpret_cast<const uint4*>(base + 2 * pool_align(2 * (size_t)n) + pool_align(4 * (size_t)n));
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        P.tex_xy[d0 + i] = make_short2(x[i], y[i]);
        P.tex_ori[d0 + i] = ori[i];
        P.tex_codes[d0 + i] = codes[i];
    }
}

}  // namespace lafis
