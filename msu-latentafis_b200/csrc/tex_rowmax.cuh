// Texture-template similarity by PQ look-up and its per-row maximum — stages K1 + K2 + K3(first half).
//
//   reference: LatentTextureTemplate::compute_dist_to_codewords   matching/include.h:327-359   (K1)
//              One2One_texture_matching, method 1                 matching/matcher.cpp:566-594 (K2)
//              row-wise std::max_element                          matching/matcher.cpp:727-735 (K3a)
//
// A CTA owns 8 latent texture points ("row tile") of one latent.  It builds their distance table
// LUT[s][code][b][row] (sub-quantizer m = 4s+b) in 128 KB of shared memory, straight from the
// latent descriptors and the codebook, and keeps it there while it streams a slice of the gallery's
// PQ codes through it.  The similarity matrix is never materialised: per (gallery template, row)
// only max_j sim[row][j] and the first j attaining it leave the SM (that is all K3 reads).
//
// Shared-memory gather layout.  One LDS.128 serves 8 lanes per bank phase; a phase is conflict-free
// when its 8 lanes cover all 32 banks.  Lane = (quarter q: 2 bits, pair p: 2 bits, half h: 1 bit).
//   * lanes of a pair share the rolled point j = 16*batch + 4q + p and split the 8 rows 4|4;
//   * at step t of group s, pair p gathers sub-quantizer m = 4s + ((t+p)&3), so the four pairs of a
//     phase always hit the four different b-slices (32 B each) of one 128 B LUT line: no conflicts,
//     no cross-lane reduction per element.
// Arithmetic order is the reference's: four running values, value b fed by sub-quantizers
// b, b+4, b+8, b+12 in that order, value 0 starting from 6; result (d1+d2)+(d3+d4).  Each lane keeps
// the four values in registers indexed by t; the value for b sits at t = (b-p)&3, and because fp32
// addition is commutative the final expression needs only two selects on the parity of p.
#pragma once
#include "device_common.cuh"

namespace lafis {

constexpr int kRowTile = 8;
constexpr int kRowmaxThreads = 512;
constexpr int kLutBytes = 4 * 256 * 128;                        // [s][code][b][row] floats
constexpr int kRowmaxSmem = kLutBytes + kRowTile * kDesLenD * 4;  // + descriptor tile

struct TexRowmaxParams {
    const float* lat_des;      // [Q][lt_stride][96]
    const int* lat_nt;         // [Q]
    int lt_stride;
    int Q;
    const float* codebook;     // [16][256][6]
    const uint32_t* tex_off;   // gallery
    const uint4* codes;
    int g0, n_chunk;           // gallery templates [g0, g0+n_chunk)
    int slices;                // the chunk is cut into this many slices
    float* rowmax_val;         // [Q][n_chunk][lt_stride]
    uint16_t* rowmax_j;        // [Q][n_chunk][lt_stride]
    int* job_counter;          // zeroed before launch
};

__global__ void __launch_bounds__(kRowmaxThreads, 1) tex_rowmax_kernel(TexRowmaxParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* lut = reinterpret_cast<float*>(smem);
    float* des_tile = reinterpret_cast<float*>(smem + kLutBytes);
    __shared__ int s_job;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kRowmaxThreads / 32;
    const int n_rowtiles = P.lt_stride / kRowTile;
    const int n_jobs = P.Q * n_rowtiles * P.slices;
    const int slice_len = (P.n_chunk + P.slices - 1) / P.slices;

    // lane roles
    const int pr = (lane >> 1) & 3, hf = lane & 1;
    const int jl = (lane >> 3) * 4 + pr;  // rolled point within a batch of 16
    uint32_t sel[4], off[4];
    float init[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int b = (t + pr) & 3;
        sel[t] = 0x4440u | (uint32_t)b;  // __byte_perm selector: byte b of the word, zero extended
        off[t] = smem_u32(lut) + (uint32_t)(b * 32 + hf * 16);
        init[t] = (b == 0) ? 6.0f : 0.0f;
    }
    const bool odd = (pr & 1) != 0;

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

        // ---- K1: LUT for rows [rt*8, rt*8+8) ----
        const float* dsrc = P.lat_des + ((size_t)q * P.lt_stride + (size_t)rt * kRowTile) * kDesLenD;
        for (int e = tid; e < kRowTile * kDesLenD; e += kRowmaxThreads) des_tile[e] = dsrc[e];
        __syncthreads();
        for (int mc = tid; mc < 16 * 256; mc += kRowmaxThreads) {
            const int m = mc >> 8, code = mc & 255;
            const float2* w2 = reinterpret_cast<const float2*>(P.codebook + (size_t)mc * 6);
            const float2 wa = __ldg(w2), wb = __ldg(w2 + 1), wc = __ldg(w2 + 2);
            const float w[6] = {wa.x, wa.y, wb.x, wb.y, wc.x, wc.y};
            float out[kRowTile];
#pragma unroll
            for (int r = 0; r < kRowTile; ++r) {
                const float* d = des_tile + r * kDesLenD + m * 6;
                float dist = 0.0f;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const float t = f_sub(d[k], w[k]);
                    dist = f_add(dist, f_mul(t, t));
                }
                out[r] = dist;
            }
            float4* dst = reinterpret_cast<float4*>(lut + (((m >> 2) * 256 + code) * 4 + (m & 3)) * 8);
            dst[0] = make_float4(out[0], out[1], out[2], out[3]);
            dst[1] = make_float4(out[4], out[5], out[6], out[7]);
        }
        __syncthreads();

        // ---- K2 + K3a: stream the slice ----
        const int t_begin = slice * slice_len;
        const int t_end = min(P.n_chunk, t_begin + slice_len);
        for (int tl = t_begin + warp; tl < t_end; tl += NW) {
            const int g = P.g0 + tl;
            const uint32_t base = P.tex_off[g];
            const int n = (int)(P.tex_off[g + 1] - base);
            float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            int bestj[4] = {0, 0, 0, 0};
            const uint4* cp = P.codes + base + jl;
            uint4 cnext = (n > 0) ? __ldg(cp) : make_uint4(0, 0, 0, 0);
            for (int j0 = 0; j0 < n; j0 += 16) {
                const uint4 c = cnext;
                if (j0 + 16 < n) cnext = __ldg(cp + j0 + 16);
                float acc[4][4];
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[t][r] = init[t];
                const uint32_t words[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                for (int s = 0; s < 4; ++s) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const uint32_t code = __byte_perm(words[s], 0u, sel[t]);
                        const uint32_t addr = off[t] + code * 128u + (uint32_t)(s * 32768);
                        float4 v;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                                     : "r"(addr));
                        acc[t][0] = f_sub(acc[t][0], v.x);
                        acc[t][1] = f_sub(acc[t][1], v.y);
                        acc[t][2] = f_sub(acc[t][2], v.z);
                        acc[t][3] = f_sub(acc[t][3], v.w);
                    }
                }
                const int j = j0 + jl;
                const bool valid = j < n;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float x = f_add(acc[0][r], odd ? acc[3][r] : acc[1][r]);
                    const float y = f_add(acc[2][r], odd ? acc[1][r] : acc[3][r]);
                    const float sim = f_add(x, y);
                    if (valid && sim > best[r]) {
                        best[r] = sim;
                        bestj[r] = j;
                    }
                }
            }
            // first maximum over the 16 lanes that hold the same rows (lane bits 1..4)
#pragma unroll
            for (int mask = 2; mask <= 16; mask <<= 1) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float ov = __shfl_xor_sync(0xffffffffu, best[r], mask);
                    const int oj = __shfl_xor_sync(0xffffffffu, bestj[r], mask);
                    if (ov > best[r] || (ov == best[r] && oj < bestj[r])) {
                        best[r] = ov;
                        bestj[r] = oj;
                    }
                }
            }
            if (lane < 2) {
                const size_t o = ((size_t)q * P.n_chunk + tl) * P.lt_stride + (size_t)rt * kRowTile + hf * 4;
                *reinterpret_cast<float4*>(P.rowmax_val + o) = make_float4(best[0], best[1], best[2], best[3]);
                ushort4 js = make_ushort4((unsigned short)bestj[0], (unsigned short)bestj[1],
                                          (unsigned short)bestj[2], (unsigned short)bestj[3]);
                *reinterpret_cast<ushort4*>(P.rowmax_j + o) = js;
            }
        }
    }
}

}  // namespace lafis
