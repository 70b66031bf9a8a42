// Micro-benchmark: which issue pipes the scalar / packed fp32 instructions of sm_100a share.
// Each mode runs a fixed instruction mix per loop iteration; the table printed is warp-instructions/clk/SM.
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fmul(float a, float b) { float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fadd(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned iadd(unsigned a, unsigned b) { unsigned r; asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

// NP packed-mul (as fma2 with -0), NQ packed-add, NM scalar mul, NA scalar add, NI integer add per iteration
template <int NP, int NQ, int NM, int NA, int NI>
__global__ void k(float* out, int iters, float x, float y, u64 nz) {
    u64 p[16], q[16];
    float m[16], a[16];
    unsigned ii[16];
    for (int i = 0; i < 16; ++i) { p[i] = pk(x + i, y + threadIdx.x); q[i] = pk(y + i, x); m[i] = x + i; a[i] = y + i; ii[i] = i + threadIdx.x; }
    u64 pb = pk(x, y);
    float bx = x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i < NP) p[i] = fma2(p[i], pb, nz);
            if (i < NQ) q[i] = add2(q[i], pb);
            if (i < NM) m[i] = fmul(m[i], bx);
            if (i < NA) a[i] = fadd(a[i], bx);
            if (i < NI) ii[i] = iadd(ii[i], 3u);
        }
    }
    float s = 0;
    for (int i = 0; i < 16; ++i) s += a[i] + m[i] + (float)(p[i] & 0xffff) + (float)(q[i] & 0xffff) + (float)ii[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NP, int NQ, int NM, int NA, int NI>
void run(float* out, const char* name) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k<NP, NQ, NM, NA, NI><<<148 * 4, 512>>>(out, iters, 1.0f, 2.0f, 0x8000000080000000ull);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    const double winst = 148.0 * 4 * 16 * iters * (NP + NQ + NM + NA + NI);
    const double clk = ms * 1e-3 * 1.965e9;
    printf("%-44s fma2=%2d add2=%2d fmul=%2d fadd=%2d iadd=%2d  %7.3f ms  %.2f warp-inst/clk/SM  clk/iter/SMSP=%.1f (%s)\n", name, NP, NQ, NM, NA, NI, ms,
           winst / clk / 148, clk / iters / 16.0 , cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 4 * 512 * 4);
    run<16, 0, 0, 0, 0>(out, "packed mul only");
    run<0, 16, 0, 0, 0>(out, "packed add only");
    run<0, 0, 16, 0, 0>(out, "scalar mul only");
    run<0, 0, 0, 16, 0>(out, "scalar add only");
    run<0, 0, 0, 0, 16>(out, "int add only");
    run<0, 0, 16, 16, 0>(out, "scalar mul+add (current kernel)");
    run<8, 8, 0, 0, 0>(out, "packed mul + packed add (16 MAC)");
    run<8, 0, 0, 16, 0>(out, "packed mul + scalar add (16 MAC)");
    run<8, 4, 0, 8, 0>(out, "packed mul + half packed add (16 MAC)");
    run<8, 3, 0, 10, 0>(out, "packed mul + 3 packed/10 scalar add (16 MAC)");
    run<8, 2, 0, 12, 0>(out, "packed mul + 2 packed/12 scalar add (16 MAC)");
    run<8, 6, 0, 4, 0>(out, "packed mul + 6 packed/4 scalar add (16 MAC)");
    run<0, 8, 16, 0, 0>(out, "scalar mul + packed add (16 MAC)");
    run<0, 4, 16, 8, 0>(out, "scalar mul + 4 packed/8 scalar add");
    run<8, 8, 0, 0, 8>(out, "packed mul+add + 8 iadd");
    run<0, 0, 16, 16, 8>(out, "scalar mul+add + 8 iadd");
    run<4, 4, 8, 8, 0>(out, "half packed, half scalar (16 MAC)");
    run<6, 6, 4, 4, 0>(out, "3/4 packed, 1/4 scalar (16 MAC)");
    return 0;
}
