// Minutiae-template similarity, normalisation and top-120 candidate selection — stages K5 + K6 + K7,
// as two kernels (the one-kernel version in minu_corr.cuh spent ~70 % of its time at block barriers).
//
//   reference: One2One_minutiae_matching steps 1-3, matching/matcher.cpp:440-488
//
// minu_sim_kernel (K5).  Persistent CTAs, one per SM.  S = max(0, A.B^T) for the three selected latent
// templates of a latent (matcher.cpp:380) against one gallery template per step.  Every output
// accumulates k = 0..95 in order with an unfused fp32 multiply and add (the Eigen stand-in's order,
// oracle/shim/Eigen/Dense).  The gallery's k-major descriptor blocks stream through a double buffer
// with cp.async while the previous block is being multiplied; the latent's blocks stay resident
// (single latent) or are re-staged per latent (batches; minu_sim_jobs_kernel stages them once per
// job instead).  Warp tile 20 x 16*NC, thread tile 10 x NC with NC = 6..10 chosen per template
// (96..160 columns): per k five loads feed 10*NC multiply-adds.  S goes to HBM ([nL][np] per job) -
// ~115 KB per pair written once
// and read once, far below what the path's fp32 issue rate lets HBM see.
//
// minu_select_kernel (K6 + K7).  One 384-thread CTA per (latent, template, slot), ~55 KB of shared
// memory so that four CTAs share an SM and hide each other's barriers.  S arrives with cp.async into a copy
// whose row stride is 4 mod 8 floats: rows are 16-byte aligned and 16-byte loads along a row as well as down
// eight consecutive rows are conflict-free.  Column sums (i ascending) and row sums (j ascending) by one thread
// per column / row.  The 120 largest normalised values S/(l_i + r_j - S + 1e-6) are found in two passes over an
// fp32 estimate of that value (relative error < 1e-6), a thread owning one group of four consecutive columns
// and every RP-th row (its four column sums stay in registers): the K-th largest of the <= 384 thread maxima
// bounds the K-th largest estimate from below, a 1024-bin histogram of the thread maxima' float bit patterns
// locates its bin, and everything at or above that bin's lower edge (with a 4e-6 relative safety margin, which
// provably contains the exact top-120) becomes a candidate; only candidates get the reference's double-precision
// evaluation (:467).  Candidates are ranked by counting (rank = number of candidates with a larger value).  If
// two of the selected values tie, or fewer than 120 values are positive, the permutation libstdc++'s introsort
// would produce is not implied by the values and the job goes to minu_select_slow_kernel, which evaluates every
// value in double and replays the introsort (stdsort_emul.h).  The output carries the RAW similarity (:486).
#pragma once
#include "device_common.cuh"
#include "minu_plan.h"
#include "stdsort_emul.h"

namespace lafis {

constexpr int kSimThreads = 384;   // 12 warps: the 3 x 80 latent rows make 12 tiles of 20 rows, 3 per scheduler
constexpr int kSimTileRows = 20;   // rows of a warp tile (10 per thread, two thread rows)
constexpr int kSelThreads = 384;  // 12 warps per job, 4 jobs per SM (shared memory): 48 of 64 warp slots
constexpr uint32_t kSelBinBase = (127u - 16u) << 6;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct MinuSimParams {
    // latent side
    const int* slot_n;         // [3Q]
    const uint32_t* slot_off;  // [3Q] padded offsets
    const float* lat_desT;
    const int* lat_status;     // [Q]
    int Q;
    // gallery side
    const uint32_t* minu_off;
    const uint16_t* minu_n;
    const float* minu_desT;
    int g0, n_chunk;
    // shared-memory geometry (floats)
    int a_slot_stride;  // one latent slot: 96 * max padded slot count + 16
    int b_buf_stride;   // one gallery block: 96 * max padded template count + 160
    int b_double;       // two gallery buffers
    int parts;          // jobs per latent (see minu_sim_kernel)
    int l_cap, r_cap;   // latent slots / gallery templates with more minutiae are skipped here (minu_big.cuh takes them)
    // output: S[job][i * np + j], job = (q * n_chunk + tl) * 3 + slot
    float* S;
    size_t job_stride;
};


// One warp tile of S = max(0, A.B^T): 20 rows x 16*NC columns.  A thread owns 10 rows - 8 consecutive ones at
// tile + li*8 and 2 at tile + 16 + li*2 (li = lane / 16), so that its A operands are two 16-byte and one 8-byte shared
// memory load per k - and NC columns, 16 lanes across: NC = 8 is the 128-column tile (columns jbase + {lj*4..+3,
// 64+lj*4..+3}); the other widths (96, 112, 144, 160 columns for NC = 6, 7, 9, 10) let one tile span a whole template
// with at most 15 padding columns: 4 @ lj*4, then 4 @ 64+lj*4 (NC >= 8) or 2 @ 64+lj*2 (NC < 8), then the remaining 1
// or 2 columns.  k ascending, unfused.  With 3 x 80 latent rows the 12 tiles keep all 12 warps (3 per scheduler)
// equally busy; the former 16-row tiles left one of 16 warps idle and one scheduler a quarter short.
template <int NC>
__device__ __forceinline__ void sim_tile(const float* __restrict__ As, int tile0, int li, int npL, const float* __restrict__ Bt,
                                         int npR, int lj, int jbase, int nL, float* __restrict__ out) {
    constexpr int N2 = NC >= 8 ? 4 : 2;          // width of the second column group
    constexpr int N3 = NC - 4 - N2;              // width of the third (0, 1 or 2)
    constexpr int J3 = 64 + 16 * N2;             // its first column: 128 or 96
    constexpr int RT = 10;
    const int ja = jbase + lj * 4, jb = jbase + 64 + lj * N2, jc = jbase + J3 + lj * N3;
    const float* bpa = Bt + ja;
    const float* bpb = Bt + jb;
    const float* bpc = Bt + jc;
    const int i8 = tile0 + li * 8, i2 = tile0 + 16 + li * 2;
    // operand pointers advance by one k-row per step (one add each instead of a multiply-add per load)
    const float* pa8 = As + i8;
    const float* pa2 = As + i2;
    float acc[RT][NC];
#pragma unroll
    for (int a = 0; a < RT; ++a)
#pragma unroll
        for (int b = 0; b < NC; ++b) acc[a][b] = 0.0f;
#pragma unroll 4
    for (int k = 0; k < 96; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(pa8);
        const float4 a1 = *reinterpret_cast<const float4*>(pa8 + 4);
        const float2 a2 = *reinterpret_cast<const float2*>(pa2);
        pa8 += npL;
        pa2 += npL;
        float bv[NC];
        {
            const float4 b0 = *reinterpret_cast<const float4*>(bpa);
            bv[0] = b0.x, bv[1] = b0.y, bv[2] = b0.z, bv[3] = b0.w;
        }
        if (N2 == 4) {
            const float4 b1 = *reinterpret_cast<const float4*>(bpa + 64);  // jb = ja + 64
            bv[4] = b1.x, bv[5] = b1.y, bv[6] = b1.z, bv[7] = b1.w;
        } else {
            const float2 b1 = *reinterpret_cast<const float2*>(bpb);
            bv[4] = b1.x, bv[5] = b1.y;
            bpb += npR;
        }
        if (N3 == 2) {
            const float2 b2 = *reinterpret_cast<const float2*>(bpc);
            bv[NC - 2] = b2.x, bv[NC - 1] = b2.y;
            bpc += npR;
        } else if (N3 == 1) {
            bv[NC - 1] = *bpc;
            bpc += npR;
        }
        bpa += npR;
        const float av[RT] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y};
#pragma unroll
        for (int a = 0; a < RT; ++a)
#pragma unroll
            for (int b = 0; b < NC; ++b) acc[a][b] = f_add(acc[a][b], f_mul(av[a], bv[b]));
    }
#pragma unroll
    for (int a = 0; a < RT; ++a) {
        const int i = a < 8 ? i8 + a : i2 + (a - 8);
        if (i >= nL) continue;
        float v[NC];
#pragma unroll
        for (int b = 0; b < NC; ++b) v[b] = acc[a][b] < 0.0f ? 0.0f : acc[a][b];  // matcher.cpp:449-450
        float* row = out + (size_t)i * npR;
        if (ja < npR) *reinterpret_cast<float4*>(row + ja) = make_float4(v[0], v[1], v[2], v[3]);
        if (N2 == 4) {
            if (jb < npR) *reinterpret_cast<float4*>(row + jb) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
            if (jb < npR) *reinterpret_cast<float2*>(row + jb) = make_float2(v[4], v[5]);
        }
        if (N3 == 2) {
            if (jc < npR) *reinterpret_cast<float2*>(row + jc) = make_float2(v[NC - 2], v[NC - 1]);
        } else if (N3 == 1) {
            if (jc < npR) row[jc] = v[NC - 1];
        }
    }
}

__global__ void __launch_bounds__(kSimThreads, 1) minu_sim_kernel(MinuSimParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* A = reinterpret_cast<float*>(smem);        // [3][a_slot_stride]
    float* B = A + 3 * (size_t)P.a_slot_stride;       // [1|2][b_buf_stride]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kSimThreads / 32;
    const int li = lane >> 4, lj = lane & 15;

    // a template's size and offset are read one step before they are needed (two templates ahead of the one being
    // multiplied): the global-load latency of these two words would otherwise stall the whole CTA once per template
    auto head = [&](int tl, int& nR, uint32_t& off) {
        nR = 0;
        off = 0;
        if (tl < P.n_chunk) {
            nR = P.minu_n[P.g0 + tl];
            off = P.minu_off[P.g0 + tl];
        }
    };
    auto load_B = [&](int nR, uint32_t off, int buf) {
        if (nR <= 0 || nR > P.r_cap) return;
        const int np = (nR + 3) & ~3;
        const float4* src = reinterpret_cast<const float4*>(P.minu_desT + (size_t)96 * off);
        float4* dst = reinterpret_cast<float4*>(B + (size_t)buf * P.b_buf_stride);
        for (int e = tid; e < 24 * np; e += kSimThreads) cp_async16(dst + e, src + e);
    };
    auto load_A = [&](int q) {
        for (int s = 0; s < 3; ++s) {
            const int nL = P.slot_n[q * 3 + s];
            if (nL <= 0 || nL > P.l_cap) continue;
            const int np = (nL + 3) & ~3;
            const float4* src = reinterpret_cast<const float4*>(P.lat_desT + (size_t)96 * P.slot_off[q * 3 + s]);
            float4* dst = reinterpret_cast<float4*>(A + (size_t)s * P.a_slot_stride);
            for (int e = tid; e < 24 * np; e += kSimThreads) cp_async16(dst + e, src + e);
        }
    };

    int tl = blockIdx.x;
    if (tl >= P.n_chunk) return;
    int nR, nR_next, nR_next2;
    uint32_t off, off_next, off_next2;
    head(tl, nR, off);
    head(tl + gridDim.x, nR_next, off_next);
    load_B(nR, off, 0);
    if (P.Q == 1) load_A(0);
    cp_async_commit();

    for (int it = 0; tl < P.n_chunk; tl += gridDim.x, ++it) {
        const int cur = P.b_double ? (it & 1) : 0;
        const int tl_next = tl + gridDim.x;
        head(tl_next + gridDim.x, nR_next2, off_next2);  // consumed in the next iteration
        bool prefetched = false;
        if (P.b_double && tl_next < P.n_chunk) {
            load_B(nR_next, off_next, cur ^ 1);
            prefetched = true;
        }
        cp_async_commit();
        const int npR = (nR + 3) & ~3;
        const float* Bt = B + (size_t)cur * P.b_buf_stride;
        const int tiles_j = (nR + 127) >> 7;

        for (int q = 0; q < P.Q; ++q) {
            const bool live = P.lat_status[q] == 0 && nR > 0 && nR <= P.r_cap;
            if (P.Q > 1) {
                __syncthreads();  // previous latent's tiles are done with A
                if (live) load_A(q);
                cp_async_commit();
                cp_async_wait<0>();
            } else {
                if (prefetched) cp_async_wait<1>();
                else cp_async_wait<0>();
            }
            __syncthreads();
            if (!live) continue;

            // tile list over the three slots.  A warp tile is 16 rows x 16*NC columns (8 x NC per thread); templates
            // of up to 160 minutiae are spanned by ONE tile of the smallest sufficient width, larger ones by
            // 128-column tiles.
            const int nc = nR <= 160 ? max(6, (nR + 15) >> 4) : 8;
            const int tiles_jj = nR <= 160 ? 1 : tiles_j;
            int t0[4];
            t0[0] = 0;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int rows = P.slot_n[q * 3 + s] <= P.l_cap ? P.slot_n[q * 3 + s] : 0;
                t0[s + 1] = t0[s] + ((rows + kSimTileRows - 1) / kSimTileRows) * tiles_jj;
            }
            for (int t = warp; t < t0[3]; t += NW) {
                const int s = (t >= t0[2]) ? 2 : (t >= t0[1]) ? 1 : 0;
                const int tt = t - t0[s];
                const int ti = tt / tiles_jj, tj = tt - ti * tiles_jj;
                const int nL = P.slot_n[q * 3 + s];
                const int npL = (nL + 3) & ~3;
                const float* As = A + (size_t)s * P.a_slot_stride;
                float* out = P.S + ((size_t)((size_t)q * P.n_chunk + tl) * 3 + s) * P.job_stride;
                switch (nc) {
                    case 6: sim_tile<6>(As, ti * kSimTileRows, li, npL, Bt, npR, lj, 0, nL, out); break;
                    case 7: sim_tile<7>(As, ti * kSimTileRows, li, npL, Bt, npR, lj, 0, nL, out); break;
                    case 9: sim_tile<9>(As, ti * kSimTileRows, li, npL, Bt, npR, lj, 0, nL, out); break;
                    case 10: sim_tile<10>(As, ti * kSimTileRows, li, npL, Bt, npR, lj, 0, nL, out); break;
                    default: sim_tile<8>(As, ti * kSimTileRows, li, npL, Bt, npR, lj, tj * 128, nL, out); break;
                }
            }
        }
        __syncthreads();  // everyone is done with this gallery block before it is overwritten
        if (!P.b_double && tl_next < P.n_chunk) {
            load_B(nR_next, off_next, 0);
            cp_async_commit();
        }
        nR = nR_next;
        off = off_next;
        nR_next = nR_next2;
        off_next = off_next2;
    }
    cp_async_wait<0>();
}

// The same tiles scheduled as jobs = (latent, part of the chunk): the latent's blocks are staged once per job instead
// of once per (template, latent), and the job count is a multiple of the grid.  Measured: better when a batch leaves
// only a few templates per CTA and chunk (256 latents, 620-template chunks: 1,474 -> 1,261 ms per 5.1 M pairs), slightly
// worse otherwise (23.6 against 23.3 ms for a single latent, 24.7 against 24.0 ms per latent for 27) - the host picks.
__global__ void __launch_bounds__(kSimThreads, 1) minu_sim_jobs_kernel(MinuSimParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* A = reinterpret_cast<float*>(smem);        // [3][a_slot_stride]
    float* B = A + 3 * (size_t)P.a_slot_stride;       // [1|2][b_buf_stride]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kSimThreads / 32;
    const int li = lane >> 4, lj = lane & 15;

    auto load_B = [&](int tl, int buf) {
        const int g = P.g0 + tl;
        const int nR = P.minu_n[g];
        if (nR <= 0 || nR > P.r_cap) return;
        const int np = (nR + 3) & ~3;
        const float4* src = reinterpret_cast<const float4*>(P.minu_desT + (size_t)96 * P.minu_off[g]);
        float4* dst = reinterpret_cast<float4*>(B + (size_t)buf * P.b_buf_stride);
        for (int e = tid; e < 24 * np; e += kSimThreads) cp_async16(dst + e, src + e);
    };
    auto load_A = [&](int q) {
        for (int s = 0; s < 3; ++s) {
            const int nL = P.slot_n[q * 3 + s];
            if (nL <= 0 || nL > P.l_cap) continue;
            const int np = (nL + 3) & ~3;
            const float4* src = reinterpret_cast<const float4*>(P.lat_desT + (size_t)96 * P.slot_off[q * 3 + s]);
            float4* dst = reinterpret_cast<float4*>(A + (size_t)s * P.a_slot_stride);
            for (int e = tid; e < 24 * np; e += kSimThreads) cp_async16(dst + e, src + e);
        }
    };

    // Jobs = (latent, part of the chunk): the latent's blocks are staged once per job and stay resident while the
    // job's templates (tl = part, part + parts, ...) stream through the double buffer.  A single latent has as
    // many parts as CTAs; a batch gets parts = SMs / gcd(Q, SMs) so that the job count is a multiple of the grid.
    const int parts = P.parts;
    const int n_jobs = P.Q * parts;
    for (int job = blockIdx.x; job < n_jobs; job += gridDim.x) {
        const int q = job / parts, part = job - q * parts;
        __syncthreads();  // the previous job is done with A and B
        if (part >= P.n_chunk || P.lat_status[q] != 0) continue;  // uniform across the CTA
        load_A(q);
        load_B(part, 0);
        cp_async_commit();

        int tl = part;
        for (int it = 0; tl < P.n_chunk; tl += parts, ++it) {
            const int cur = P.b_double ? (it & 1) : 0;
            const int tl_next = tl + parts;
            bool prefetched = false;
            if (P.b_double && tl_next < P.n_chunk) {
                load_B(tl_next, cur ^ 1);
                prefetched = true;
            }
            cp_async_commit();
            const int g = P.g0 + tl;
            const int nR = P.minu_n[g];
            const int npR = (nR + 3) & ~3;
            const float* Bt = B + (size_t)cur * P.b_buf_stride;
            const int tiles_j = (nR + 127) >> 7;
            if (prefetched) cp_async_wait<1>();
            else cp_async_wait<0>();
            __syncthreads();
            if (nR > 0 && nR <= P.r_cap) {
                // tile list over the three slots.  A warp tile is 16 rows x 16*NC columns (8 x NC per thread);
                // templates of up to 160 minutiae are spanned by ONE tile of the smallest sufficient width, larger
                // ones by 128-column tiles.
                const int nc = nR <= 160 ? max(6, (nR + 15) >> 4) : 8;
                const int tiles_jj = nR <= 160 ? 1 : tiles_j;
                int t0[4];
                t0[0] = 0;
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const int rows = P.slot_n[q * 3 + s] <= P.l_cap ? P.slot_n[q * 3 + s] : 0;
                    t0[s + 1] = t0[s] + ((rows + kSimTileRows - 1) / kSimTileRows) * tiles_jj;
                }
                for (int t = warp; t < t0[3]; t += NW) {
                    const int s = (t >= t0[2]) ? 2 : (t >= t0[1]) ? 1 : 0;
                    const int tt = t - t0[s];
                    const int ti = tt / tiles_jj, tj = tt - ti * tiles_jj;
                    const int nL = P.slot_n[q * 3 + s];
                    const int npL = (nL + 3) & ~3;
                    const float* As = A + (size_t)s * P.a_slot_stride;
                    float* out = P.S + ((size_t)((size_t)q * P.n_chunk + tl) * 3 + s) * P.job_stride;
                    switch (nc) {
                        case 6: sim_tile<6>(As, ti * kSimTileRows, li, npL, Bt, npR, lj, 0, nL, out); break;
                        case 7: sim_tile<7>(As, ti * kSimTileRows, li, npL, Bt, npR, lj, 0, nL, out); break;
                        case 9: sim_tile<9>(As, ti * kSimTileRows, li, npL, Bt, npR, lj, 0, nL, out); break;
                        case 10: sim_tile<10>(As, ti * kSimTileRows, li, npL, Bt, npR, lj, 0, nL, out); break;
                        default: sim_tile<8>(As, ti * kSimTileRows, li, npL, Bt, npR, lj, tj * 128, nL, out); break;
                    }
                }
            }
            __syncthreads();  // everyone is done with this gallery block before it is overwritten
            if (!P.b_double && tl_next < P.n_chunk) {
                load_B(tl_next, 0);
                cp_async_commit();
            }
        }
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
struct MinuSelectParams {
    const int* slot_n;
    const int* lat_status;
    int Q;
    const uint16_t* minu_n;
    int g0, n_chunk;
    const float* S;
    size_t job_stride;
    int max_nL, max_np;  // shared-memory geometry
    int l_cap, r_cap;    // larger latent slots / gallery templates are left to minu_big.cuh (corr_n = 0 here)
    int slow_dense;      // minu_select_slow_kernel keeps a dense copy of the keys
    float* corr_v;       // [job][120]
    uint32_t* corr_ij;   // [job][120] (i << 16) | j
    int* corr_n;         // [job]
    int* slow_count;
    int* slow_jobs;
};


// the reference's normalised similarity, matcher.cpp:467: float sums, "+0.000001" promotes the
// denominator and the division to double, the quotient is narrowed to float
__device__ __forceinline__ uint32_t exact_key(float s, float l, float r) {
    const float den = f_sub(f_add(l, r), s);
    const double qd = (double)s / ((double)den + 0.000001);
    uint32_t key = __float_as_uint((float)qd);
    if (key == 0x80000000u) key = 0;
    return key;
}
// fp32 estimate of the same value; relative error < 1e-6
__device__ __forceinline__ float approx_key(float s, float l, float r) {
    const float den = f_sub(f_add(l, r), s) + 0.000001f;
    float rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(den));  // 1 ulp
    return s * rc;
}

__device__ __forceinline__ float rcp_approx(float x) {
    float rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(x));  // 1 ulp
    return rc;
}

__global__ void __launch_bounds__(kSelThreads, 4) minu_select_kernel(MinuSelectParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // grid = (3 * n_chunk, latents): no run-time division in the job decode (job_grid, device_common.cuh)
    const int q = job_latent();
    if (q >= P.Q) return;
    const int tl = (int)(blockIdx.x / 3u), slot = (int)(blockIdx.x - 3u * (unsigned)tl);
    const size_t job = (size_t)q * gridDim.x + blockIdx.x;  // (q * n_chunk + tl) * 3 + slot
    const int nR = P.minu_n[P.g0 + tl];
    const int nL = (P.lat_status[q] == 0) ? P.slot_n[q * 3 + slot] : 0;
    if (nL <= 0 || nR <= 0 || nL > P.l_cap || nR > P.r_cap) {
        // rolled without minutiae / latent slot absent: score stays 0; oversized pairs are filled in by minu_big.cuh
        if (tid == 0) P.corr_n[job] = 0;
        return;
    }
    const int np = (nR + 3) & ~3, G = np >> 2, ld = sel_ld(np);
    float* Ssm = reinterpret_cast<float*>(smem);                       // [nL][ld]
    float* lsum = Ssm + (size_t)P.max_nL * sel_ld(P.max_np);           // [nL]
    float* rsum = lsum + ((P.max_nL + 3) & ~3);                        // [np], 16-byte aligned
    int* hist = reinterpret_cast<int*>(rsum + P.max_np);               // [1024]; later: candidates and their keys
    uint32_t* cand_ij = reinterpret_cast<uint32_t*>(hist);             // [kSelMaxCand] (i << 16) | j
    uint32_t* cand_key = cand_ij + kSelMaxCand;                        // [kSelMaxCand] exact keys
    __shared__ int s_ncand, s_bad;
    __shared__ float s_thr;
    __shared__ unsigned char s_slot[kTopCorrMinu + 8];

    // a thread owns one group of four consecutive columns (g) and every RP-th row: its four column sums stay in
    // registers through both passes, and every access to the copy of S is a 16-byte one
    const int RP = kSelThreads / G;  // rows in flight (G <= 90: r_cap <= 360)
    const int g = tid % G, r0 = tid / G;
    const bool active = r0 < RP;

    const float* Sg = P.S + job * P.job_stride;
    if (active)
        for (int i = r0; i < nL; i += RP) cp_async16(Ssm + i * ld + 4 * g, Sg + (size_t)i * np + 4 * g);
    cp_async_commit();
    for (int b = tid; b < kSelBins; b += kSelThreads) hist[b] = 0;
    if (tid < kTopCorrMinu + 8) s_slot[tid] = 0;
    if (tid == 0) {
        s_ncand = 0;
        s_bad = 0;
    }
    cp_async_wait<0>();
    __syncthreads();

    // ---- K6: sums.  first half of the CTA: columns (i ascending); second half: rows (j ascending; the padding
    //      columns hold +0, which leaves a non-negative fp32 accumulator unchanged) ----
    if (tid < kSelThreads / 2) {
        for (int j = tid; j < np; j += kSelThreads / 2) {
            float acc = Ssm[j];
#pragma unroll 8
            for (int i = 1; i < nL; ++i) acc = f_add(acc, Ssm[i * ld + j]);
            rsum[j] = acc;
        }
    } else {
        for (int i = tid - kSelThreads / 2; i < nL; i += kSelThreads / 2) {
            const float4* row = reinterpret_cast<const float4*>(Ssm + i * ld);
            float4 v = row[0];
            float acc = f_add(f_add(f_add(v.x, v.y), v.z), v.w);
#pragma unroll 4
            for (int c = 1; c < G; ++c) {
                v = row[c];
                acc = f_add(f_add(f_add(f_add(acc, v.x), v.y), v.z), v.w);
            }
            lsum[i] = acc;
        }
    }
    __syncthreads();

    // ---- pass 1: fp32 estimates (written over the copy of S; the raw value is re-read from HBM for the few
    //      candidates).  Every thread keeps the largest of its ~28 estimates: the K-th largest of those <= 384 thread
    //      maxima (distinct elements) is a lower bound of the K-th largest estimate overall and sits at the ~1.5 %
    //      quantile, so only the thread maxima go through the histogram.  Small matrices (fewer than two rows per
    //      thread: too few thread maxima for a sharp bound) put every positive estimate into the histogram instead. ----
    const int M = nL * nR;
    const int K = M < kTopCorrMinu ? M : kTopCorrMinu;
    const bool full_hist = nL < 2 * RP;
    auto bin_of = [](float a) -> uint32_t {
        const uint32_t hb = __float_as_uint(a) >> 17;
        return hb > kSelBinBase ? min(hb - kSelBinBase, (uint32_t)(kSelBins - 1)) : 0u;
    };
    if (active) {
        const float4 r4 = *reinterpret_cast<const float4*>(rsum + 4 * g);
        const float2 rp0 = make_float2(r4.x + 0.000001f, r4.y + 0.000001f);
        const float2 rp1 = make_float2(r4.z + 0.000001f, r4.w + 0.000001f);
        const float2 m1 = make_float2(-1.0f, -1.0f);
        float mymax = 0.0f;
#pragma unroll 2
        for (int i = r0; i < nL; i += RP) {
            const float l = lsum[i];
            float4* sp = reinterpret_cast<float4*>(Ssm + i * ld) + g;
            const float4 s = *sp;
            // den = (l + (r + 1e-6)) - s: within 3 ulp of the reference's denominator (s <= l, r: no cancellation)
            const float2 d0 = __ffma2_rn(make_float2(s.x, s.y), m1, __fadd2_rn(make_float2(l, l), rp0));
            const float2 d1 = __ffma2_rn(make_float2(s.z, s.w), m1, __fadd2_rn(make_float2(l, l), rp1));
            float4 a;
            a.x = s.x * rcp_approx(d0.x);
            a.y = s.y * rcp_approx(d0.y);
            a.z = s.z * rcp_approx(d1.x);
            a.w = s.w * rcp_approx(d1.y);
            *sp = a;
            if (full_hist) {
                if (a.x > 0.0f) atomicAdd(&hist[bin_of(a.x)], 1);
                if (a.y > 0.0f) atomicAdd(&hist[bin_of(a.y)], 1);
                if (a.z > 0.0f) atomicAdd(&hist[bin_of(a.z)], 1);
                if (a.w > 0.0f) atomicAdd(&hist[bin_of(a.w)], 1);
            } else {
                mymax = fmaxf(fmaxf(mymax, fmaxf(a.x, a.y)), fmaxf(a.z, a.w));
            }
        }
        if (mymax > 0.0f) atomicAdd(&hist[bin_of(mymax)], 1);
    }
    __syncthreads();
    if (warp == 0) {  // bin of the K-th largest histogram entry, scanning 32-bin chunks from the top
        int above = 0, bin = 0;  // fewer than K entries: bin 0, every positive estimate is a candidate
        for (int c = kSelBins / 32 - 1; c >= 0; --c) {
            const int h = hist[c * 32 + lane];
            const int tot = __reduce_add_sync(0xffffffffu, h);
            if (above + tot >= K) {
                // suffix sums inside the chunk: lanes above me
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
            above += tot;
        }
        // lower edge of the bin, lowered by 4e-6 relative (estimate error < 1e-6 on either side): at least K
        // estimates lie at or above the edge, so the exact K-th largest value does too.  Bin 0 collects everything
        // below 2^-16: every positive estimate is a candidate then (the smallest denormal as the threshold).
        if (lane == 0) s_thr = bin > 0 ? __uint_as_float(((uint32_t)bin + kSelBinBase) << 17) * (1.0f - 4e-6f) : __uint_as_float(1u);
    }
    __syncthreads();

    // ---- pass 2: candidates (the histogram's storage becomes the candidate list) ----
    if (active) {
        const float thr = s_thr;
        for (int i = r0; i < nL; i += RP) {
            const float4 a = *(reinterpret_cast<const float4*>(Ssm + i * ld) + g);
            if (fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)) >= thr) {  // one reservation for the thread's (1..4) candidates
                const float av[4] = {a.x, a.y, a.z, a.w};
                int pos = atomicAdd(&s_ncand, (int)(a.x >= thr) + (int)(a.y >= thr) + (int)(a.z >= thr) + (int)(a.w >= thr));
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (av[c] >= thr) {
                        if (pos < kSelMaxCand) cand_ij[pos] = ((uint32_t)i << 16) | (uint32_t)(4 * g + c);
                        ++pos;
                    }
            }
        }
    }
    __syncthreads();
    const int nc = s_ncand;
    if (nc > kSelMaxCand || nc < K) {
        // pathological value distribution, or the K-th value is not positive (ties among zeros decide the order):
        // the slow kernel sorts everything
        if (tid == 0) P.slow_jobs[atomicAdd(P.slow_count, 1)] = (int)job;
        return;
    }
    // ---- the reference's double-precision value (:467) of every candidate ----
    uint32_t my_ij[2], my_key[2];
    float my_s[2];
    const int nc4 = (nc + 3) & ~3;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int c = tid + r * kSelThreads;
        my_key[r] = 0u;
        if (c < nc) {
            my_ij[r] = cand_ij[c];
            const int i = (int)(my_ij[r] >> 16), j = (int)(my_ij[r] & 0xffffu);
            my_s[r] = __ldg(Sg + (size_t)i * np + j);
            my_key[r] = exact_key(my_s[r], lsum[i], rsum[j]);
        }
        if (c < nc4) cand_key[c] = my_key[r];  // zero padding to a multiple of four keys
    }
    __syncthreads();
    // ---- rank = number of candidates with a larger value.  The candidates contain every element that can be among
    //      the K largest, so ranks below K are global ranks.  Equal values share a rank and leave the next one empty:
    //      an empty rank among 0..K means that a tie reaches into the first K positions (or straddles position K), where
    //      the permutation is introsort-specific (slow kernel). ----
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int c = tid + r * kSelThreads;
        if (c >= nc) break;
        const uint32_t mine = my_key[r];
        int gt = 0;
        const uint4* k4 = reinterpret_cast<const uint4*>(cand_key);
#pragma unroll 4
        for (int e = 0; e < nc4 / 4; ++e) {
            const uint4 k = k4[e];
            gt += (k.x > mine) + (k.y > mine) + (k.z > mine) + (k.w > mine);
        }
        if (gt <= K) s_slot[gt] = 1;
        if (gt < K) {
            P.corr_v[job * kTopCorrMinu + gt] = my_s[r];  // the RAW similarity (:486)
            P.corr_ij[job * kTopCorrMinu + gt] = my_ij[r];
        }
    }
    __syncthreads();
    if (tid < K + (nc > K ? 1 : 0) && s_slot[tid] == 0) s_bad = 1;
    __syncthreads();
    if (tid == 0) {
        if (s_bad) P.slow_jobs[atomicAdd(P.slow_count, 1)] = (int)job;
        else P.corr_n[job] = K;
    }
}

// Jobs whose order depends on how libstdc++'s introsort permutes equal keys.
// `dense`: a second, densely packed copy of the keys (no index arithmetic in the replay's comparator); large
// templates that cannot afford it compare through the strided copy.

__global__ void __launch_bounds__(kSelThreads) minu_select_slow_kernel(MinuSelectParams P, unsigned long long* replay_count) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kSelThreads / 32;
    float* Ssm = reinterpret_cast<float*>(smem);
    float* lsum = Ssm + (size_t)P.max_nL * (P.max_np + 1);
    float* rsum = lsum + P.max_nL;
    uint32_t* skeys = reinterpret_cast<uint32_t*>(Ssm);                              // strided [nL][ld], in place
    uint32_t* dkeys = reinterpret_cast<uint32_t*>(rsum + P.max_np);                  // dense [nL * nR] (slow_dense)
    uint16_t* y = P.slow_dense ? reinterpret_cast<uint16_t*>(dkeys + (size_t)P.max_nL * P.max_np)
                               : reinterpret_cast<uint16_t*>(dkeys);                 // [nL * nR]
    const int n_jobs = *P.slow_count;
    for (int jb = blockIdx.x; jb < n_jobs; jb += gridDim.x) {
        const size_t job = (size_t)P.slow_jobs[jb];
        const int slot = (int)(job % 3);
        const size_t pair = job / 3;
        const int tl = (int)(pair % P.n_chunk), q = (int)(pair / P.n_chunk);
        const int nR = P.minu_n[P.g0 + tl];
        const int nL = P.slot_n[q * 3 + slot];
        const int np = (nR + 3) & ~3, ld = np + 1;
        const float* Sg = P.S + job * P.job_stride;
        __syncthreads();
        for (int i = warp; i < nL; i += NW)
            for (int j = lane; j < nR; j += 32) Ssm[i * ld + j] = Sg[(size_t)i * np + j];
        __syncthreads();
        for (int j = tid; j < nR; j += kSelThreads) {
            float acc = Ssm[j];
            for (int i = 1; i < nL; ++i) acc = f_add(acc, Ssm[i * ld + j]);
            rsum[j] = acc;
        }
        for (int i = tid; i < nL; i += kSelThreads) {
            float acc = Ssm[i * ld];
            for (int j = 1; j < nR; ++j) acc = f_add(acc, Ssm[i * ld + j]);
            lsum[i] = acc;
        }
        __syncthreads();
        for (int i = warp; i < nL; i += NW)
            for (int j = lane; j < nR; j += 32) {
                const float s = Ssm[i * ld + j];
                const uint32_t key = (s != 0.0f) ? exact_key(s, lsum[i], rsum[j]) : 0u;
                if (P.slow_dense) dkeys[i * nR + j] = key;
                else skeys[i * ld + j] = key;
            }
        __syncthreads();
        const int M = nL * nR;
        const int K = M < kTopCorrMinu ? M : kTopCorrMinu;
        if (P.slow_dense) {
            // replay of libstdc++'s introsort with block-parallel partition steps (stdsort_emul.h); the strided copy
            // of S is dead once the dense keys exist and serves as the two stop lists
            __shared__ int s_part[2 * NW + 2];
            uint16_t* Ls = reinterpret_cast<uint16_t*>(Ssm);
            BlockStdSortEmu<DenseKey<uint32_t>, uint16_t, kSelThreads> bs{DenseKey<uint32_t>{dkeys}, y, Ls, Ls + M, s_part};
            bs.sort_prefix(M, K);
            if (tid == 0) atomicAdd(replay_count, 1ull);
        } else if (warp == 0) {  // large templates: warp-cooperative replay through the strided keys
            {
                const uint32_t* kk = skeys;
                const int nRr = nR, ldd = ld;
                auto keyfn = [kk, nRr, ldd](int e) -> uint32_t {
                    const int i = e / nRr;
                    return kk[i * ldd + (e - i * nRr)];
                };
                warp_std_sort_desc_prefix(keyfn, y, M, K);
            }
            if (lane == 0) atomicAdd(replay_count, 1ull);
        }
        __syncthreads();
        if (tid < K) {
            const int e = y[tid];
            const int i = e / nR, j = e - i * nR;
            P.corr_v[job * kTopCorrMinu + tid] = Sg[(size_t)i * np + j];
            P.corr_ij[job * kTopCorrMinu + tid] = ((uint32_t)i << 16) | (uint32_t)j;
        }
        if (tid == 0) P.corr_n[job] = K;
    }
}

}  // namespace lafis
