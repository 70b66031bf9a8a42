// Micro-benchmark: throughput of unfused fp32 multiply+add (what the exact-order dot products need)
// scalar vs packed f32x2 (Blackwell mul.rn.f32x2 / add.rn.f32x2) vs FFMA, and LDS.128 gather rate.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

template <int MODE>
__global__ void k(float* out, int iters, float x, float y) {
    float a[16];
    unsigned long long p[8];
    for (int i = 0; i < 16; ++i) a[i] = x + i + threadIdx.x;
    for (int i = 0; i < 8; ++i) p[i] = pk(x + i, y + threadIdx.x);
    float bx = x, by = y;
    unsigned long long pb = pk(x, y), pc = pk(y, x);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = __fadd_rn(a[i], __fmul_rn(bx, by + (float)i));
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(bx, by, a[i]);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = add2(p[i], mul2(pb, pc));
        } else if (MODE == 3) {
            const unsigned long long nz = 0x8000000080000000ull;
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = add2(p[i], fma2(pb, pc + i, nz));
        }
        bx += 1e-9f;
        pb = add2(pb, pc);
    }
    float s = 0;
    for (int i = 0; i < 16; ++i) s += a[i];
    for (int i = 0; i < 8; ++i) s += (float)(p[i] & 0xffff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 4 * 512 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    const char* names[4] = {"FMUL+FADD scalar (16 MAC/iter)", "FFMA scalar (16 MAC/iter)", "mul.f32x2+add.f32x2 (ptxas fuses!)", "fma.f32x2(a,b,-0)+add.f32x2 exact"};
    for (int mode = 0; mode < 4; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 4, 512>>>(out, iters, 1.0f, 2.0f);
            if (mode == 1) k<1><<<148 * 4, 512>>>(out, iters, 1.0f, 2.0f);
            if (mode == 2) k<2><<<148 * 4, 512>>>(out, iters, 1.0f, 2.0f);
            if (mode == 3) k<3><<<148 * 4, 512>>>(out, iters, 1.0f, 2.0f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            double macs = 148.0 * 4 * 512 * iters * 16;
            if (rep) printf("%-36s %8.3f ms  %.2f TMAC/s  = %.1f MAC/clk/SM @1.965GHz (%s)\n", names[mode], ms, macs / ms / 1e9,
                            macs / (ms * 1e-3) / 148 / 1.965e9, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
