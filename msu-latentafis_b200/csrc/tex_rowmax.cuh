// Texture-template similarity by PQ look-up and its per-row maximum — stages K1 + K2 + K3(first half).
//
//   reference: LatentTextureTemplate::compute_dist_to_codewords   matching/include.h:327-359   (K1)
//              One2One_texture_matching, method 1                 matching/matcher.cpp:566-594 (K2)
//              row-wise std::max_element                          matching/matcher.cpp:727-735 (K3a)
//
// What the matcher consumes of the nLt x nRt similarity matrix is, per latent row, only its maximum and
// the first column attaining it.  The kernels below find exactly that without evaluating every entry in
// fp32:
//
//  tex_lut_kernel        K1 verbatim (fp32, k ascending, unfused) into HBM, LUT[q][row][m][code], plus a
//                        per-row scale = 4095 / max entry.
//  tex_rowmax_kernel     A persistent CTA owns 16 latent rows.  It keeps a 12-BIT QUANTISED copy of their
//                        LUT rows in 128 KB of shared memory, [s][code][b][row] (sub-quantizer m = 4s + b),
//                        and streams the gallery's PQ code words through it: one LDS.128 returns the
//                        quantised entries of 8 rows, which are summed as packed 16-bit integers (16
//                        entries of <= 4095 cannot overflow 16 bits).  With q = floor(LUT * scale) computed
//                        in fp32, q - 1 < LUT * scale < q + 2, so for the integer sums
//                        Dq - 16 < scale * D < Dq + 32: a column whose Dq exceeds the row's minimum Dq by
//                        more than 48 (+ slack for the fp32 rounding of the reference's own summation;
//                        kWindow = 64 in total) has a strictly smaller similarity than the column holding
//                        that minimum and can neither be the maximum nor tie with it.  Columns inside the
//                        window (typically 1-3 per row) are queued and re-evaluated EXACTLY - fp32 LUT from
//                        HBM/L2, the reference's four running values and summation order - and the row
//                        maximum / first arg-max is taken over those exact values.
//
// The quantised pass moves half the shared-memory bytes per (row, column, sub-quantizer) of an fp32 pass,
// which is the resource that bounds this path; the exact pass touches < 1 % of the
// entries.  Shared-memory gather layout as before: lanes of a pair share the rolled point
// j = 16*batch + 4*quarter + pair and split the 16 rows 8|8; at step t of group s, pair p gathers
// sub-quantizer m = 4s + ((t+p)&3), so the four pairs of an LDS.128 phase hit the four 32-byte b-slices of
// one 128-byte line: conflict-free.
//
// Degenerate rows (all LUT entries ~0) get scale 0: every column is then inside the window and the row is
// evaluated exactly in full, like a template whose candidate queue overflows.
#pragma once
#include "device_common.cuh"

namespace lafis {

constexpr int kRowTile = 16;
constexpr int kRowmaxThreads = 512;
constexpr int kRowmaxRegs = 96;  // leaves 16K registers per SM for the selection / graph CTAs that run beside it
constexpr int kLutBytes = 4 * 256 * 128;  // [s][code][b][row] u16
constexpr int kWindow = 64;               // see the bound above: 48 + 16 slack
constexpr int kQueueCap = 480;            // candidate queue entries per warp and gallery template
constexpr int kQLevels = 4095;

struct TexWarpState {
    unsigned long long best[kRowTile];  // (sortable similarity << 32) | ~column: max = largest value, first column
    uint32_t rmin[kRowTile];            // running minimum of the quantised distance per row
    uint32_t queue[kQueueCap];          // row | column << 4 | Dq << 14
    int count, overflow;
};
constexpr int kRowmaxSmem = kLutBytes + (kRowmaxThreads / 32) * (int)sizeof(TexWarpState);

// ---- K1 ---------------------------------------------------------------------------------------------
struct TexLutParams {
    const float* lat_des;   // [Q][lt_stride][96]
    const int* lat_nt;      // [Q]
    int lt_stride, Q;
    const float* codebook;  // [16][256][6]
    float* lut;             // [Q][lt_stride][16][256]
    float* row_scale;       // [Q][lt_stride]
};

__global__ void __launch_bounds__(256) tex_lut_kernel(TexLutParams P) {
    const int row = blockIdx.x, q = blockIdx.y;
    __shared__ float des[kDesLenD];
    __shared__ float wmax[8];
    const int tid = threadIdx.x;
    if (row >= P.lat_nt[q]) {
        if (tid == 0) P.row_scale[(size_t)q * P.lt_stride + row] = -1.0f;  // padding row
        return;
    }
    const float* d = P.lat_des + ((size_t)q * P.lt_stride + row) * kDesLenD;
    if (tid < kDesLenD) des[tid] = d[tid];
    __syncthreads();
    float* out = P.lut + ((size_t)q * P.lt_stride + row) * 4096;
    float mx = 0.0f;
    for (int mc = tid; mc < 4096; mc += 256) {
        const int m = mc >> 8;
        const float2* w2 = reinterpret_cast<const float2*>(P.codebook + (size_t)mc * 6);
        const float2 wa = __ldg(w2), wb = __ldg(w2 + 1), wc = __ldg(w2 + 2);
        const float w[6] = {wa.x, wa.y, wb.x, wb.y, wc.x, wc.y};
        float dist = 0.0f;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const float t = f_sub(des[m * 6 + k], w[k]);
            dist = f_add(dist, f_mul(t, t));
        }
        out[mc] = dist;
        mx = fmaxf(mx, dist);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) wmax[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
        for (int k = 1; k < 8; ++k) mx = fmaxf(mx, wmax[k]);
        // rows whose entries are all tiny are not quantised (scale 0: every column becomes a candidate)
        P.row_scale[(size_t)q * P.lt_stride + row] = (mx >= 0.01f) ? (float)kQLevels / mx : 0.0f;
    }
}

// ---- K2 + K3a ------------------------------------------------------------------------------------------
struct TexRowmaxParams {
    const float* lut;          // [Q][lt_stride][16][256] fp32, exact
    const float* row_scale;    // [Q][lt_stride]; < 0 for padding rows
    const int* lat_nt;         // [Q]
    int lt_stride;
    int Q;
    const uint32_t* tex_off;   // gallery
    const uint4* codes;
    int g0, n_chunk;           // gallery templates [g0, g0+n_chunk)
    int slices;                // the chunk is cut into this many slices
    float* rowmax_val;         // [Q][n_chunk][lt_stride]
    uint16_t* rowmax_j;        // [Q][n_chunk][lt_stride]
    int* job_counter;          // zeroed before launch
    unsigned long long* counters;  // [4] statistics: queued, exact evaluations, overflowed templates, templates
};

__device__ __forceinline__ float tex_exact_sim(const float* __restrict__ lut_row, uint4 c) {
    // matcher.cpp:577-592: four running values, value b fed by sub-quantizers b, b+4, b+8, b+12
    const uint32_t w[4] = {c.x, c.y, c.z, c.w};
    float d[4] = {6.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t code = (w[s] >> (8 * b)) & 0xffu;
            d[b] = f_sub(d[b], __ldg(lut_row + (4 * s + b) * 256 + code));
        }
    return f_add(f_add(d[0], d[1]), f_add(d[2], d[3]));
}

__device__ __forceinline__ unsigned long long tex_best_key(float sim, int j) {
    uint32_t u = __float_as_uint(sim);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (uint32_t)j);
}

__global__ void __maxnreg__(kRowmaxRegs) tex_rowmax_kernel(TexRowmaxParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint16_t* lut16 = reinterpret_cast<uint16_t*>(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kRowmaxThreads / 32;
    TexWarpState& ws = reinterpret_cast<TexWarpState*>(smem + kLutBytes)[warp];
    __shared__ int s_job;
    __shared__ float s_scale[kRowTile];

    const int n_rowtiles = P.lt_stride / kRowTile;
    const int n_jobs = P.Q * n_rowtiles * P.slices;
    const int slice_len = (P.n_chunk + P.slices - 1) / P.slices;

    // lane roles
    const int pr = (lane >> 1) & 3, hf = lane & 1;
    const int jl = (lane >> 3) * 4 + pr;  // rolled point within a batch of 16
    uint32_t sel[4], off[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int b = (t + pr) & 3;
        sel[t] = 0x4440u | (uint32_t)b;  // __byte_perm selector: byte b of the word, zero extended
        off[t] = smem_u32(lut16) + (uint32_t)(b * 32 + hf * 16);
    }
    unsigned long long n_queued = 0, n_exact = 0, n_over = 0, n_tpl = 0;

    for (;;) {
        __syncthreads();  // previous job's LUT no longer in use
        if (tid == 0) s_job = atomicAdd(P.job_counter, 1);
        __syncthreads();
        const int job = s_job;
        if (job >= n_jobs) break;
        const int slice = job % P.slices;
        const int rt = (job / P.slices) % n_rowtiles;
        const int q = job / (P.slices * n_rowtiles);
        const int nLt = P.lat_nt[q];
        if (rt * kRowTile >= nLt) continue;  // uniform across the CTA

        // ---- quantised LUT of rows [rt*16, rt*16+16) ----
        const size_t row0 = (size_t)q * P.lt_stride + (size_t)rt * kRowTile;
        if (tid < kRowTile) s_scale[tid] = P.row_scale[row0 + tid];
        __syncthreads();
        for (int r = 0; r < kRowTile; ++r) {
            const float sc = s_scale[r];
            const float* src = P.lut + (row0 + r) * 4096;
            for (int mc = tid; mc < 4096; mc += kRowmaxThreads) {
                const int m = mc >> 8, code = mc & 255;
                uint32_t qv = 0;
                if (sc > 0.0f) qv = (uint32_t)min((int)floorf(__ldg(src + mc) * sc), kQLevels);
                lut16[(((m >> 2) * 256 + code) * 4 + (m & 3)) * kRowTile + r] = (uint16_t)qv;
            }
        }
        __syncthreads();
        // rows this lane accumulates: hf*8 .. hf*8+7; padding rows never produce candidates
        uint32_t row_live = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (s_scale[hf * 8 + r] >= 0.0f) row_live |= 1u << r;

        // ---- stream the slice ----
        const int t_begin = slice * slice_len;
        const int t_end = min(P.n_chunk, t_begin + slice_len);
        for (int tl = t_begin + warp; tl < t_end; tl += NW) {
            const int g = P.g0 + tl;
            const uint32_t base = P.tex_off[g];
            const int n = (int)(P.tex_off[g + 1] - base);
            if (n <= 0) continue;
            ++n_tpl;
            if (lane < kRowTile) ws.best[lane] = 0ull;
            __syncwarp();
            const uint4* cp = P.codes + base + jl;

            // quantised distances of my 8 rows to the point of batch j0
            auto batch = [&](uint4 c, uint32_t* dq) {
                uint32_t acc[4] = {0u, 0u, 0u, 0u};
                const uint32_t words[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                for (int s = 0; s < 4; ++s) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const uint32_t code = __byte_perm(words[s], 0u, sel[t]);
                        const uint32_t addr = off[t] + code * 128u + (uint32_t)(s * 32768);
                        uint32_t v0, v1, v2, v3;
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                     : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                                     : "r"(addr));
                        acc[0] += v0;  // packed 16-bit sums: 16 entries of <= 4095 stay below 2^16
                        acc[1] += v1;
                        acc[2] += v2;
                        acc[3] += v3;
                    }
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    dq[2 * r] = acc[r] & 0xffffu;
                    dq[2 * r + 1] = acc[r] >> 16;
                }
            };

            // Per-row thresholds live in registers and are refreshed from the warp's shared minima every four
            // batches.  A threshold is (smallest Dq seen so far) + kWindow, so a stale value is only ever
            // too large: it can queue too much, never miss a candidate.
            const unsigned hf_mask = hf ? 0xaaaaaaaau : 0x55555555u;
            uint32_t thr[8];
            {   // warm-up: minima over the first 32 points, nothing queued
                uint32_t lmin[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) lmin[r] = 0xffffu;
                for (int j0 = 0; j0 < n && j0 < 32; j0 += 16) {
                    uint32_t dq[8];
                    batch(__ldg(cp + j0), dq);
                    if (j0 + jl < n) {
#pragma unroll
                        for (int r = 0; r < 8; ++r) lmin[r] = min(lmin[r], dq[r]);
                    }
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const uint32_t mn = __reduce_min_sync(hf_mask, lmin[r]);
                    thr[r] = mn + kWindow;
                    if (lane < 2) ws.rmin[hf * 8 + r] = mn;  // lanes 0 / 1 publish rows 0..7 / 8..15
                }
            }
            if (lane == 0) ws.count = 0;
            __syncwarp();
            uint4 cnext = __ldg(cp);
            int since_refresh = 0;
            for (int j0 = 0; j0 < n; j0 += 16) {
                const uint4 c = cnext;
                if (j0 + 16 < n) cnext = __ldg(cp + j0 + 16);
                uint32_t dq[8];
                batch(c, dq);
                const int j = j0 + jl;
                if (j < n) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        if (dq[r] <= thr[r] && ((row_live >> r) & 1u)) {  // rare: ~6 per row and template
                            const int slot = atomicAdd(&ws.count, 1);
                            if (slot < kQueueCap) ws.queue[slot] = (uint32_t)(hf * 8 + r) | ((uint32_t)j << 4) | (dq[r] << 14);
                            if (dq[r] + kWindow < thr[r]) {  // a new minimum: share it through shared memory
                                thr[r] = dq[r] + kWindow;
                                atomicMin(&ws.rmin[hf * 8 + r], dq[r]);
                            }
                        }
                    }
                }
                if (++since_refresh == 4) {  // pick up minima found by the other lanes
                    since_refresh = 0;
#pragma unroll
                    for (int r = 0; r < 8; ++r) thr[r] = min(thr[r], ws.rmin[hf * 8 + r] + kWindow);
                }
            }
            __syncwarp();
            // queue state
            if (lane == 0) {
                const int count = ws.count;
                ws.overflow = count > kQueueCap;
                ws.count = min(count, kQueueCap);
            }
            __syncwarp();

            // ---- exact re-evaluation of the candidates ----
            const float* lut_rows = P.lut + row0 * 4096;
            if (ws.overflow) {
                ++n_over;
                for (int e = lane; e < kRowTile * n; e += 32) {
                    const int r = e / n, j = e - r * n;
                    if (s_scale[r] < 0.0f) continue;
                    const float sim = tex_exact_sim(lut_rows + (size_t)r * 4096, __ldg(P.codes + base + j));
                    atomicMax(&ws.best[r], tex_best_key(sim, j));
                    ++n_exact;
                }
            } else {
                const int cnt = ws.count;
                n_queued += (lane == 0) ? cnt : 0;
                for (int e = lane; e < cnt; e += 32) {
                    const uint32_t ent = ws.queue[e];
                    const int r = (int)(ent & 15u), j = (int)((ent >> 4) & 1023u);
                    const uint32_t dqv = ent >> 14;
                    if (dqv > ws.rmin[r] + kWindow) continue;  // outside the final window
                    const float sim = tex_exact_sim(lut_rows + (size_t)r * 4096, __ldg(P.codes + base + j));
                    atomicMax(&ws.best[r], tex_best_key(sim, j));
                    ++n_exact;
                }
            }
            __syncwarp();
            if (lane < kRowTile && s_scale[lane] >= 0.0f) {
                const unsigned long long k = ws.best[lane];
                uint32_t u = (uint32_t)(k >> 32);
                u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
                const size_t o = ((size_t)q * P.n_chunk + tl) * P.lt_stride + (size_t)rt * kRowTile + lane;
                P.rowmax_val[o] = __uint_as_float(u);
                P.rowmax_j[o] = (uint16_t)(0xffffffffu - (uint32_t)(k & 0xffffffffull));
            }
            __syncwarp();
        }
    }
    // statistics, one atomic per warp
    n_exact = __reduce_add_sync(0xffffffffu, (unsigned)n_exact);
    if (lane == 0) {
        atomicAdd(P.counters + 0, n_queued);
        atomicAdd(P.counters + 1, n_exact);
        atomicAdd(P.counters + 2, n_over);
        atomicAdd(P.counters + 3, n_tpl);
    }
}

}  // namespace lafis
