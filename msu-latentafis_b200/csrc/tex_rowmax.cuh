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
//                        per-row scale = 255 / max entry.
//  tex_rowmax_kernel     A persistent CTA owns 32 latent rows.  It keeps an 8-BIT QUANTISED copy of their
//                        LUT rows in 128 KB of shared memory, [s][code][b][row] (sub-quantizer m = 4s + b),
//                        and streams the gallery's PQ code words through it: one LDS.128 returns the
//                        quantised entries of 16 rows (4 words x 4 bytes).  Sixteen such words are summed
//                        per column without unpacking: `all` = plain 32-bit sum of the words (carries run
//                        into the neighbouring byte lanes) and `odd` = sum of the words' odd bytes spread to
//                        16-bit fields; then  odd  holds the sums of rows 1|3 and  all - (odd << 8)  those of
//                        rows 0|2, each below 2^12.  With q = rint(LUT * scale) in fp32,
//                        |q - LUT*scale| < 0.5001, so for the integer sums |Dq - scale*D| < 8.001: a column
//                        whose Dq exceeds the row's minimum Dq by more than 16 (+ < 0.4 for the fp32
//                        rounding of the reference's own summation; kWindow = 18 in total) has a strictly
//                        smaller similarity than the column holding that minimum and can neither be the
//                        maximum nor tie with it.
//                        The stream loop is branch-free: per (lane, row) it tracks, as packed 16-bit fields
//                        (Dq >> 2) << 6 | batch, the three smallest of the lane's columns (a lane sees every
//                        16th column; the batch index names the column).  After the template, the row
//                        minimum is a warp reduction; every lane whose best column is inside the window
//                        contributes its best and second best column as candidates; a lane whose THIRD best
//                        is inside the window too (< 1 % of the rows) has its ~50 columns re-scanned.  The
//                        candidates (~1.9 per row) are re-evaluated EXACTLY - fp32 LUT from L2, the
//                        reference's four running values and summation order - and the row maximum /
//                        first arg-max is taken over those exact values.
//
// The quantised pass moves a quarter of the shared-memory bytes per (row, column, sub-quantizer) of an fp32
// pass, which is the resource that bounds this path.  Shared-memory gather layout: lanes
// of a pair share the rolled point j = 16*batch + 4*quarter + pair and split the 32 rows 16|16; at step t
// of group s, pair p gathers sub-quantizer m = 4s + ((t+p)&3), so the four pairs of an LDS.128 phase hit
// the four 32-byte b-slices of one 128-byte line: conflict-free.
//
// Degenerate rows (all LUT entries ~0) get scale 0 and are evaluated exactly in full, like the rows of a
// template whose candidate list overflows.
#pragma once
#include <type_traits>

#include "device_common.cuh"

namespace lafis {

constexpr int kRowTile = 32;
constexpr int kRowmaxThreads = 512;
constexpr int kRowmaxRegs = 128;  // the whole register file: at 96 ptxas serialised the 16 gathers of a batch and sank the code-word look-ahead
constexpr int kLutBytes = 4 * 256 * 128;  // [s][code][b][row] u8
constexpr int kWindow = 18;               // see the bound above: 16.002 + 0.4, rounded up with slack
constexpr int kWindowQ = (kWindow + 3) / 4 + 1;  // the same window on distances tracked as Dq >> 2: floor((x + W) / 4) <= floor(x / 4) + ceil(W / 4) + 1
constexpr int kCandCap = 480;             // exact re-evaluations per warp and gallery template (beyond: every row in full, ~100 x the cost)
constexpr int kAmbCap = 128;              // (lane, row) re-scans per warp and gallery template
constexpr int kQLevels = 255;

struct TexWarpState {
    unsigned long long best[kRowTile];  // (sortable similarity << 32) | ~column: max = largest value, first column
    uint32_t cand[kCandCap];            // row | column << 5
    uint32_t amb[kAmbCap];              // row | lane << 5 | limit << 10
    uint32_t full_rows;                 // rows that need every column evaluated exactly
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
    unsigned char* lut8;       // [Q][row tiles][kLutBytes]: the tiles' 8-bit tables as the kernel holds them (tex_lut8_kernel)
    const int* lat_nt;         // [Q]
    int lt_stride;
    int Q;
    const uint32_t* tex_off;   // gallery
    const uint4* codes;
    int g0, n_chunk;           // gallery templates [g0, g0+n_chunk)
    int slices;                // the chunk is cut into this many slices
    uint32_t one;              // = 1, opaque to the compiler: x * one + y keeps an addition on the FMA pipe (IMAD)
    float* rowmax_val;         // [Q][n_chunk][lt_stride]
    uint16_t* rowmax_j;        // [Q][n_chunk][lt_stride]
    int* job_counter;          // zeroed before launch
    unsigned long long* counters;  // [4] statistics: candidates, exact evaluations, overflowed templates, templates
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

// row (0..15 within the lane's half of the tile) held by 16-bit field f of the packed sums:
// words 0,2,4,6 = all - (odd << 8) of LDS word w = f >> 2 (rows 4w | 4w+2), words 1,3,5,7 = odd (rows 4w+1 | 4w+3)
__device__ __forceinline__ constexpr int tex_field_row(int f) { return 4 * (f >> 2) + ((f >> 1) & 1) + 2 * (f & 1); }

// quantised distance of ONE row to one rolled point (the re-scan of an ambiguous lane)
__device__ __forceinline__ uint32_t tex_quant_dist(const unsigned char* __restrict__ lut8, int row, uint4 c) {
    const uint32_t w[4] = {c.x, c.y, c.z, c.w};
    uint32_t d = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t code = (w[s] >> (8 * b)) & 0xffu;
            d += lut8[((s * 256 + code) * 4 + b) * kRowTile + row];
        }
    return d;
}

// The 8-bit table of one (latent, 32-row tile) in the layout tex_rowmax_kernel gathers from, [s][code][b][row] (sub-quantizer
// m = 4s + b): q = min(rint(LUT * scale), 255) per row, 0 for rows that are not quantised (padding, degenerate).  A tile with
// at most 16 live rows is laid out for the kernel's half mode: both 16-row halves hold rows 0..15.
__global__ void __launch_bounds__(512) tex_lut8_kernel(TexRowmaxParams P) {
    const int rt = blockIdx.x, q = blockIdx.y, tid = threadIdx.x;
    const int nLt = P.lat_nt[q];
    if (rt * kRowTile >= nLt) return;
    __shared__ float s_scale[kRowTile];
    const int n_rowtiles = (P.lt_stride + kRowTile - 1) / kRowTile;
    const size_t row0 = (size_t)q * P.lt_stride + (size_t)rt * kRowTile;
    const bool half = nLt - rt * kRowTile <= kRowTile / 2;
    const int src_of_slot = half ? (tid & 15) : tid;  // LUT slot -> row of the tile
    if (tid < kRowTile) s_scale[tid] = (rt * kRowTile + src_of_slot < P.lt_stride) ? P.row_scale[row0 + src_of_slot] : -1.0f;
    __syncthreads();
    unsigned char* out = P.lut8 + ((size_t)q * n_rowtiles + rt) * kLutBytes;
    // a thread quantises 16 rows of one (sub-quantizer, code) and stores them as one 16-byte vector
    for (int e = tid; e < 2 * 4096; e += 512) {
        const int mc = e & 4095, hslot = e >> 12;
        const int m = mc >> 8, code = mc & 255;
        uint32_t wv[4] = {0u, 0u, 0u, 0u};
        const float* src = P.lut + (row0 + (size_t)(half ? 0 : hslot * 16)) * 4096 + mc;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float sc = s_scale[hslot * 16 + r];
            uint32_t qv = 0;
            if (sc > 0.0f) qv = (uint32_t)min((int)rintf(f_mul(__ldg(src + (size_t)r * 4096), sc)), kQLevels);
            wv[r >> 2] |= qv << (8 * (r & 3));
        }
        *reinterpret_cast<uint4*>(out + (((m >> 2) * 256 + code) * 4 + (m & 3)) * kRowTile + hslot * 16) =
            make_uint4(wv[0], wv[1], wv[2], wv[3]);
    }
}

__global__ void __maxnreg__(kRowmaxRegs) tex_rowmax_kernel(TexRowmaxParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* lut8 = smem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    TexWarpState& ws = reinterpret_cast<TexWarpState*>(smem + kLutBytes)[warp];
    __shared__ int s_job, s_next;
    __shared__ float s_scale[kRowTile];

    const int n_rowtiles = (P.lt_stride + kRowTile - 1) / kRowTile;
    const int n_jobs = P.Q * n_rowtiles * P.slices;
    const int slice_len = (P.n_chunk + P.slices - 1) / P.slices;

    // lane roles
    const int pr = (lane >> 1) & 3, hf = lane & 1;
    const int jl = (lane >> 3) * 4 + pr;  // rolled point within a batch of 16
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned hf_mask = hf ? 0xaaaaaaaau : 0x55555555u;
    uint32_t sel[4], off[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int b = (t + pr) & 3;
        sel[t] = 0x4440u | (uint32_t)b;  // __byte_perm selector: byte b of the word, zero extended
        off[t] = smem_u32(lut8) + (uint32_t)(b * 32 + hf * 16);
    }
    unsigned long long n_cand = 0, n_exact = 0, n_over = 0, n_tpl = 0;
    const uint32_t one = P.one;

    for (;;) {
        __syncthreads();  // previous job's LUT no longer in use
        if (tid == 0) {
            s_job = atomicAdd(P.job_counter, 1);
            s_next = 0;
        }
        __syncthreads();
        const int job = s_job;
        if (job >= n_jobs) break;
        // slice-major job order: the (latent, row tile) jobs of one gallery slice are drawn back to back, so the CTAs
        // working at any time share a handful of slices whose code words stay in L2 - the gallery is read from HBM
        // about once per launch instead of once per row tile and latent
        const int per_slice = P.Q * n_rowtiles;
        const int slice = job / per_slice;
        const int q = (job - slice * per_slice) / n_rowtiles;
        const int rt = job - slice * per_slice - q * n_rowtiles;
        const int nLt = P.lat_nt[q];
        if (rt * kRowTile >= nLt) continue;  // uniform across the CTA

        // ---- quantised LUT of rows [rt*32, rt*32+32) ----
        const size_t row0 = (size_t)q * P.lt_stride + (size_t)rt * kRowTile;
        // A tile with at most 16 live rows (the last one of a latent: 400 texture points are 12.5 tiles) runs in HALF
        // mode: the 16 rows are laid into both halves of the LUT, and the two lanes of a pair - which otherwise share a
        // rolled point and split the 32 rows - take a rolled point each, 32 per batch instead of 16.  Same gathers, same
        // conflict-free pattern (a phase's eight lanes still hit the eight 16-byte groups of their 128-byte lines), half
        // the batches: the tile costs half instead of a full tile's time for half a tile's rows.
        const bool half = nLt - rt * kRowTile <= kRowTile / 2;
        const int src_of_slot = half ? (tid & 15) : tid;  // LUT slot -> row of the tile
        if (tid < kRowTile) s_scale[tid] = (rt * kRowTile + src_of_slot < P.lt_stride) ? P.row_scale[row0 + src_of_slot] : -1.0f;
        __syncthreads();
        {   // the tile's 8-bit table, quantised once per match by tex_lut8_kernel: 128 KB from L2 (a per-job build from the
            // fp32 table took ~25 us, 2.7 % of this kernel at two slices per SM).  Plain 16-byte loads, eight in flight:
            // the form of this copy decides which schedule ptxas finds for the stream loop below (cp.async or other
            // unroll factors: 24.7 ms against 24.2 ms, tools/run_variants.sh on one box).
            const uint4* src = reinterpret_cast<const uint4*>(P.lut8 + ((size_t)q * n_rowtiles + rt) * kLutBytes);
#pragma unroll 8
            for (int e = tid; e < kLutBytes / 16; e += kRowmaxThreads) reinterpret_cast<uint4*>(lut8)[e] = __ldg(src + e);
        }
        __syncthreads();
        // rows of this lane's half: live (take part), degenerate (scale 0: every column is a candidate)
        uint32_t row_live = 0, row_degen = 0;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float sc = s_scale[hf * 16 + r];
            if (sc > 0.0f) row_live |= 1u << r;
            if (sc == 0.0f) row_degen |= 1u << r;
        }
        const uint32_t degen_all =
            half ? __shfl_sync(0xffffffffu, row_degen, 0)
                 : __shfl_sync(0xffffffffu, row_degen, 0) | (__shfl_sync(0xffffffffu, row_degen, 1) << 16);
        // rolled points per batch and this lane's point inside a batch; the lanes that hold the same rows
        const int cpb = half ? 32 : 16;
        const int jcol = half ? 2 * jl + hf : jl;
        const unsigned row_mask = half ? 0xffffffffu : hf_mask;
        const uint32_t slot_mask = half ? 0xffffu : 0xffffffffu;  // LUT slots that stand for rows of the tile

        // ---- stream the slice: warps draw templates from a shared counter (template sizes vary 600..1000 points) ----
        const int t_begin = slice * slice_len;
        const int t_end = min(P.n_chunk, t_begin + slice_len);
        // A warp always holds its next template too and asks L2 for that template's code words (12.8 KB) while it
        // works on the current one: the stream loop's own look-ahead of one batch then never waits for DRAM.
        auto draw = [&]() {
            int t = 0;
            if (lane == 0) t = t_begin + atomicAdd(&s_next, 1);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t < t_end) {
                const uint32_t b = P.tex_off[P.g0 + t];
                const int bytes = (int)(P.tex_off[P.g0 + t + 1] - b) * 16;
                const char* p = reinterpret_cast<const char*>(P.codes + b);
                for (int o = lane * 128; o < bytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
            }
            return t;
        };
        int tl_next = draw();
        for (;;) {
            const int tl = tl_next;
            if (tl >= t_end) break;
            tl_next = draw();
            const int g = P.g0 + tl;
            const uint32_t base = P.tex_off[g];
            const int n = (int)(P.tex_off[g + 1] - base);
            if (n <= 0) continue;
            ++n_tpl;
            ws.best[lane] = 0ull;
            const uint4* cp = P.codes + base + jcol;

            // per (lane, row): the two smallest tracking fields of this lane's columns, as packed 16-bit fields
            // (Dq >> 2) << 6 | batch: 10 bits of distance (Dq <= 4080), 6 bits of batch index (<= 62)
            uint32_t f1[8], f2[8], f3[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) f1[u] = f2[u] = f3[u] = 0xffffffffu;

            // code words two batches ahead: wherever ptxas places the load inside the body, a full batch of work
            // lies between it and its first use
            auto stream = [&](auto cpb_tag) {
            constexpr int CPB = decltype(cpb_tag)::value;  // rolled points per batch: 16, or 32 in half mode
            uint4 cnext = __ldg(cp), cnext2 = cnext;
            if (CPB < n) cnext2 = __ldg(cp + CPB);
            uint32_t bp = 0;  // batch index in both fields
            for (int j0 = 0; j0 < n; j0 += CPB, bp += 0x00010001u) {
                const uint4 c = cnext;
                cnext = cnext2;
                if (j0 + 2 * CPB < n) cnext2 = __ldg(cp + j0 + 2 * CPB);
                // The integer ALU pipe (PRMT, IADD3, VIMNMX: 16 lanes/clk per scheduler) would bound this loop, so the
                // additions go to the FMA pipe (IMAD with a run-time multiplier of 1); the two pipes end up level.
                uint32_t all[4] = {0u, 0u, 0u, 0u}, odd[4] = {0u, 0u, 0u, 0u};
                const uint32_t words[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                for (int s = 0; s < 4; ++s) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const uint32_t code = __byte_perm(words[s], 0u, sel[t]);
                        const uint32_t addr = off[t] + code * 128u + (uint32_t)(s * 32768);
                        uint32_t v[4];
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                                     : "r"(addr));
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            all[w] = v[w] * one + all[w];                  // bytes carry into their neighbours: resolved below
                            const uint32_t o = __byte_perm(v[w], 0u, 0x4341u);  // rows 1 | 3 as 16-bit fields
                            if (s == 3) odd[w] += o;  // a quarter stays on the ALU pipe (pairs fuse into IADD3)
                            else odd[w] = o * one + odd[w];
                        }
                    }
                }
                if (j0 + jcol < n) {
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const uint32_t dq[2] = {all[w] - (odd[w] << 8), odd[w]};  // rows 4w | 4w+2, rows 4w+1 | 4w+3
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int u = 2 * w + h;
                            const uint32_t t = ((dq[h] << 4) & 0xffc0ffc0u) | bp;
                            f3[u] = __vminu2(f3[u], __vmaxu2(f2[u], t));
                            f2[u] = __vminu2(f2[u], __vmaxu2(f1[u], t));
                            f1[u] = __vminu2(f1[u], t);
                        }
                    }
                }
            }
            };
            if (half) stream(std::integral_constant<int, 32>{});
            else stream(std::integral_constant<int, 16>{});

            // ---- candidates: columns whose quantised distance is within kWindow of the row minimum ----
            int cnt = 0, acnt = 0;
#pragma unroll
            for (int f = 0; f < 16; ++f) {
                const int rl = tex_field_row(f);
                const uint32_t mine = (f & 1) ? (f1[f >> 1] >> 16) : (f1[f >> 1] & 0xffffu);
                const uint32_t second = (f & 1) ? (f2[f >> 1] >> 16) : (f2[f >> 1] & 0xffffu);
                const uint32_t third = (f & 1) ? (f3[f >> 1] >> 16) : (f3[f >> 1] & 0xffffu);
                // Dq_j <= Dq_min + kWindow  =>  (Dq_j >> 2) <= (Dq_min >> 2) + kWindowQ
                const uint32_t limq = (__reduce_min_sync(row_mask, mine) >> 6) + kWindowQ;
                const bool live = (row_live >> rl) & 1u;
                const bool amb = live && (third >> 6) <= limq;                 // >= 3 of my columns inside the window: re-scan
                const bool cand = live && !amb && (mine >> 6) <= limq;         // my best column
                const bool cand2 = cand && (second >> 6) <= limq;              // and my second best
                const unsigned mc = __ballot_sync(0xffffffffu, cand), mc2 = __ballot_sync(0xffffffffu, cand2);
                const unsigned ma = __ballot_sync(0xffffffffu, amb);
                const int row = half ? rl : hf * 16 + rl;
                if (cand) {
                    const int pos = cnt + __popc(mc & lt_mask);
                    if (pos < kCandCap) ws.cand[pos] = (uint32_t)row | ((uint32_t)(cpb * (int)(mine & 63u) + jcol) << 5);
                }
                cnt += __popc(mc);
                if (cand2) {
                    const int pos = cnt + __popc(mc2 & lt_mask);
                    if (pos < kCandCap) ws.cand[pos] = (uint32_t)row | ((uint32_t)(cpb * (int)(second & 63u) + jcol) << 5);
                }
                cnt += __popc(mc2);
                if (amb) {
                    const int pos = acnt + __popc(ma & lt_mask);
                    if (pos < kAmbCap) ws.amb[pos] = (uint32_t)row | ((uint32_t)lane << 5) | (limq << 10);
                }
                acnt += __popc(ma);
            }
            __syncwarp();
            bool overflow = acnt > kAmbCap;
            // re-scan of the ambiguous (lane, row)s: all 32 lanes share the ~n/16 columns of that lane
            for (int a = 0; a < acnt && !overflow; ++a) {
                const uint32_t ent = ws.amb[a];
                const int row = (int)(ent & 31u), al = (int)((ent >> 5) & 31u);
                const uint32_t lim = ent >> 10;
                const int ajl = (al >> 3) * 4 + ((al >> 1) & 3);
                const int ajcol = half ? 2 * ajl + (al & 1) : ajl;
                for (int b0 = 0; b0 * cpb < n; b0 += 32) {
                    const int j = (b0 + lane) * cpb + ajcol;
                    bool hit = false;
                    if (j < n) hit = (tex_quant_dist(lut8, row, __ldg(P.codes + base + j)) >> 2) <= lim;
                    const unsigned mh = __ballot_sync(0xffffffffu, hit);
                    if (hit) {
                        const int pos = cnt + __popc(mh & lt_mask);
                        if (pos < kCandCap) ws.cand[pos] = (uint32_t)row | ((uint32_t)j << 5);
                    }
                    cnt += __popc(mh);
                }
            }
            overflow = overflow || cnt > kCandCap;
            __syncwarp();

            // ---- exact re-evaluation of the candidates ----
            const float* lut_rows = P.lut + row0 * 4096;
            uint32_t full = degen_all;  // rows evaluated in full
            if (overflow) {
                ++n_over;
                full = slot_mask;
            } else {
                n_cand += (lane == 0) ? cnt : 0;
                for (int e = lane; e < cnt; e += 32) {
                    const uint32_t ent = ws.cand[e];
                    const int r = (int)(ent & 31u), j = (int)(ent >> 5);
                    const float sim = tex_exact_sim(lut_rows + (size_t)r * 4096, __ldg(P.codes + base + j));
                    atomicMax(&ws.best[r], tex_best_key(sim, j));
                    ++n_exact;
                }
            }
            while (full) {  // rare: degenerate rows / overflowed templates
                const int r = __ffs(full) - 1;
                full &= full - 1;
                if (s_scale[r] < 0.0f) continue;
                for (int j = lane; j < n; j += 32) {
                    const float sim = tex_exact_sim(lut_rows + (size_t)r * 4096, __ldg(P.codes + base + j));
                    atomicMax(&ws.best[r], tex_best_key(sim, j));
                    ++n_exact;
                }
            }
            __syncwarp();
            if (((slot_mask >> lane) & 1u) && s_scale[lane] >= 0.0f) {
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
        atomicAdd(P.counters + 0, n_cand);
        atomicAdd(P.counters + 1, n_exact);
        atomicAdd(P.counters + 2, n_over);
        atomicAdd(P.counters + 3, n_tpl);
    }
}

}  // namespace lafis
