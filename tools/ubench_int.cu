// Micro-benchmark: issue rates of the integer instructions the texture kernel is built from (sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned u32;

template <int MODE>
__global__ void k(u32* out, int iters, u32 x, u32 y) {
    u32 a[16];
    for (int i = 0; i < 16; ++i) a[i] = x + i * 7 + threadIdx.x;
    u32 b = y + threadIdx.x, c = x ^ y;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) a[i] = a[i] + b + c;                                   // IADD3
            if (MODE == 1) a[i] = (a[i] & b) ^ c;                                 // LOP3
            if (MODE == 2) a[i] = __byte_perm(a[i], b, 0x4341 + (c & 0));         // PRMT
            if (MODE == 3) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));  // IMAD
            if (MODE == 4) a[i] = min(a[i], b + i);                               // IMNMX (+add folded?)
            if (MODE == 5) a[i] = __vminu2(a[i], b);                              // packed u16 min
            if (MODE == 6) a[i] = __vmaxu2(__vminu2(a[i], b), c);                 // packed min+max
            if (MODE == 7) { a[i] = a[i] + b + c; asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[(i + 8) & 15]) : "r"(b), "r"(c)); }  // IADD3 + IMAD mix
            if (MODE == 8) { a[i] = __byte_perm(a[i], b, 0x4341); asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[(i + 8) & 15]) : "r"(b), "r"(c)); }  // PRMT + IMAD
            if (MODE == 9) a[i] = __funnelshift_r(a[i], b, 8);                    // SHF
            if (MODE == 10) a[i] = min(min(a[i], b), c);                          // 3-input min?
            if (MODE == 11) a[i] = __vaddus2(a[i], b);                            // packed saturating add
            if (MODE == 12) a[i] = __vadd2(a[i], b);                              // packed add
            if (MODE == 13) a[i] = __vcmpleu2(a[i], b);                           // packed compare
        }
        b += c;
    }
    u32 s = 0;
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// LDS.128 gather + integer mix, as in the texture kernel: NALU extra ALU ops per LDS.128
template <int NALU, int NIMAD>
__global__ void lds_mix(u32* out, int iters, u32 x) {
    extern __shared__ uint4 sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = make_uint4(i, i * 3, i * 5, i * 7);
    __syncthreads();
    u32 acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    u32 idx = threadIdx.x * 37 + x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint4 v = sm[(idx + u * 1031) & 8191];
            if (NALU >= 1) acc[0] = acc[0] + v.x + v.y;
            if (NALU >= 2) acc[1] = acc[1] + v.z + v.w;
            if (NALU >= 3) acc[2] += __byte_perm(v.x, 0, 0x4341);
            if (NALU >= 5) acc[3] += __byte_perm(v.y, 0, 0x4341);
            if (NALU >= 7) acc[4] += __byte_perm(v.z, 0, 0x4341);
            if (NALU >= 9) acc[5] += __byte_perm(v.w, 0, 0x4341);
            if (NIMAD >= 1) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[6]) : "r"(v.x), "r"(x));
            if (NIMAD >= 2) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[7]) : "r"(v.y), "r"(x));
            if (NALU == 0 && NIMAD == 0) acc[0] ^= v.x ^ v.y ^ v.z ^ v.w;
        }
        idx = idx * 5 + acc[0];
    }
    u32 s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(u32* out, const char* name, int nops) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<148 * 4, 512>>>(out, iters, 12345u, 777u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    const double winst = 148.0 * 4 * 16 * iters * 16.0 * nops;
    const double clk = ms * 1e-3 * 1.965e9;
    printf("%-28s %7.3f ms  %.2f source-op warps/clk/SM (%s)\n", name, ms, winst / clk / 148, cudaGetErrorString(cudaGetLastError()));
}
template <int NALU, int NIMAD>
void run_lds(u32* out, const char* name) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    float ms = 0;
    cudaFuncSetAttribute(lds_mix<NALU, NIMAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        lds_mix<NALU, NIMAD><<<148, 512, 131072>>>(out, iters, 3u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    const double lds = 148.0 * 16 * iters * 8.0;
    const double clk = ms * 1e-3 * 1.965e9;
    printf("%-40s %7.3f ms  %.3f LDS.128 warps/clk/SM = %.0f%% of 128 B/clk (%s)\n", name, ms, lds / clk / 148, 100.0 * lds / clk / 148 * 4, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    u32* out;
    cudaMalloc(&out, 148 * 4 * 512 * 4);
    run<0>(out, "IADD3", 1);
    run<1>(out, "LOP3", 1);
    run<2>(out, "PRMT", 1);
    run<3>(out, "IMAD", 1);
    run<4>(out, "IMNMX", 1);
    run<5>(out, "vminu2", 1);
    run<6>(out, "vminu2+vmaxu2", 2);
    run<7>(out, "IADD3 + IMAD", 2);
    run<8>(out, "PRMT + IMAD", 2);
    run<9>(out, "SHF", 1);
    run<10>(out, "min3", 1);
    run<11>(out, "vaddus2", 1);
    run<12>(out, "vadd2", 1);
    run<13>(out, "vcmpleu2", 1);
    run_lds<0, 0>(out, "LDS.128 random gather + xor");
    run_lds<2, 0>(out, "LDS.128 + 2 IADD3");
    run_lds<10, 0>(out, "LDS.128 + 2 IADD3 + 4 (PRMT+IADD)");
    run_lds<6, 2>(out, "LDS.128 + 2 IADD3 + 2 (PRMT+IADD) + 2 IMAD");
    return 0;
}
