// Descriptor compression 192 -> 96: the reference's CompNet in eval mode
// (extraction/models/net_compress.py:33-53 with BasicBlock :7-31) followed by the re-normalisation of
// extraction/descriptor_DR.py:150-152 (row / ||row|| * 1.73):
//
//   h1 = lrelu(bn0(x  W0^T + b0))                       layer1   (192 -> 96)
//   t  = lrelu(bn1(h1 W1^T + b1))                       layer2.layers[0..2]
//   u  = lrelu(bn2(t  W2^T + b2) + h1)                  layer2.layers[3..4] + residual (:24-30)
//   y  = bn3(u W3^T + b3)                               layer3
//   out = y / sqrt(sum y^2) * 1.73f
//
// Floating-point work: fp32 FMA on the CUDA cores (the reference runs fp32 torch; parity is a tolerance,
// tests/test_gpu_compnet.py), BatchNorm folded with the bias into one scale/shift per output on the host.
// Two persistent kernels so that the weights in shared memory leave room for enough warps to hide the
// shared-memory latency (one kernel holding all 184 KB ran 6 warps per SM and reached 39 % of the FMA rate):
//   compnet_l1_kernel    W0 (72 KB, k-major) resident, 16 warps/SM, x -> h1 (HBM, 384 B per descriptor)
//   compnet_l234_kernel  W1..W3 (108 KB) resident, 12 warps/SM, h1 -> out
// A warp owns 8 descriptors at a time and carries them through its layers in a private activation buffer, so the
// only block barrier is the one after the weight load.  Thread tile: 8 descriptors x 3 outputs (lane, lane+32,
// lane+64): per 4 k eight broadcast LDS.128 and 12 conflict-free weight loads feed 96 FFMA.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace lafis {

constexpr int kCompIn = 192, kCompOut = 96;
constexpr int kCompPts = 8;                  // descriptors per warp and pass
constexpr int kCompWarpsL1 = 16, kCompWarpsL234 = 12;
constexpr int kCompW0Floats = kCompIn * kCompOut;                                // 18,432
constexpr int kCompWeightFloats = kCompW0Floats + 3 * kCompOut * kCompOut;       // 46,080
constexpr int kCompAffFloats = 4 * 2 * kCompOut;                                 // scale, shift per layer

struct CompNetParams {
    const float* x;      // [n][192]
    long long n;
    float* h1;           // [n][96] scratch: output of layer1
    float* out;          // [n][96]
    const float* wt;     // k-major: W0t [192][96], W1t, W2t, W3t [96][96]
    const float* aff;    // [4][2][96]: y = acc * scale + shift  (scale = gamma / sqrt(var + eps),
                         //             shift = (bias - mean) * scale + beta)
    int normalise;       // descriptor_DR.py:150-152
};

__host__ __device__ inline size_t compnet_l1_smem_bytes() {
    return sizeof(float) * ((size_t)kCompW0Floats + 2 * kCompOut + (size_t)kCompWarpsL1 * kCompPts * kCompIn);
}
__host__ __device__ inline size_t compnet_l234_smem_bytes() {
    return sizeof(float) * ((size_t)3 * kCompOut * kCompOut + 6 * kCompOut + (size_t)kCompWarpsL234 * kCompPts * 2 * kCompOut);
}

__device__ __forceinline__ float lrelu02(float v) { return v > 0.0f ? v : v * 0.2f; }

// acc[p][c] = sum_k in[p][k] * Wt[k][lane + 32 c], k ascending
template <int K, int STRIDE>
__device__ __forceinline__ void compnet_layer(const float* __restrict__ in, const float* __restrict__ Wt, int lane,
                                              float (&acc)[kCompPts][3]) {
#pragma unroll
    for (int p = 0; p < kCompPts; ++p) acc[p][0] = acc[p][1] = acc[p][2] = 0.0f;
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        float4 xv[kCompPts];
#pragma unroll
        for (int p = 0; p < kCompPts; ++p) xv[p] = *reinterpret_cast<const float4*>(in + p * STRIDE + k);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const float w0 = Wt[(k + kk) * kCompOut + lane], w1 = Wt[(k + kk) * kCompOut + lane + 32],
                        w2 = Wt[(k + kk) * kCompOut + lane + 64];
#pragma unroll
            for (int p = 0; p < kCompPts; ++p) {
                const float xs = kk == 0 ? xv[p].x : kk == 1 ? xv[p].y : kk == 2 ? xv[p].z : xv[p].w;
                acc[p][0] = fmaf(xs, w0, acc[p][0]);
                acc[p][1] = fmaf(xs, w1, acc[p][1]);
                acc[p][2] = fmaf(xs, w2, acc[p][2]);
            }
        }
    }
}

// the warp's 8 descriptors of ROW floats each are contiguous in HBM: coalesced 16-byte loads into its buffer,
// rows beyond n read as zeros
template <int ROW>
__device__ __forceinline__ void compnet_stage(float* __restrict__ dst_f, const float* __restrict__ src_f, int np, int lane) {
    const float4* src = reinterpret_cast<const float4*>(src_f);
    float4* dst = reinterpret_cast<float4*>(dst_f);
    const int valid = np * ROW / 4;
#pragma unroll
    for (int e = 0; e < kCompPts * ROW / 4 / 32; ++e) {
        const int idx = lane + 32 * e;
        dst[idx] = idx < valid ? __ldcs(src + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__global__ void __launch_bounds__(kCompWarpsL1 * 32, 1) compnet_l1_kernel(CompNetParams P) {
    extern __shared__ __align__(16) float csm[];
    float* W0 = csm;
    float* aff = W0 + kCompW0Floats;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* A = aff + 2 * kCompOut + warp * kCompPts * kCompIn;  // the warp's descriptors [8][192]
    {
        const float4* src = reinterpret_cast<const float4*>(P.wt);
        float4* dst = reinterpret_cast<float4*>(W0);
        for (int e = tid; e < kCompW0Floats / 4; e += blockDim.x) dst[e] = __ldg(src + e);
        for (int e = tid; e < 2 * kCompOut; e += blockDim.x) aff[e] = __ldg(P.aff + e);
    }
    __syncthreads();
    float sc[3], sh[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        sc[c] = aff[lane + 32 * c];
        sh[c] = aff[kCompOut + lane + 32 * c];
    }
    const long long n_groups = (P.n + kCompPts - 1) / kCompPts;
    for (long long g = (long long)blockIdx.x * kCompWarpsL1 + warp; g < n_groups; g += (long long)gridDim.x * kCompWarpsL1) {
        const long long p0 = g * kCompPts;
        const int np = (int)((P.n - p0) < kCompPts ? (P.n - p0) : kCompPts);
        __syncwarp();
        compnet_stage<kCompIn>(A, P.x + p0 * kCompIn, np, lane);
        __syncwarp();
        float acc[kCompPts][3];
        compnet_layer<kCompIn, kCompIn>(A, W0, lane, acc);
#pragma unroll
        for (int p = 0; p < kCompPts; ++p)
            if (p < np) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    P.h1[(p0 + p) * kCompOut + lane + 32 * c] = lrelu02(fmaf(acc[p][c], sc[c], sh[c]));
            }
    }
}

__global__ void __launch_bounds__(kCompWarpsL234 * 32, 1) compnet_l234_kernel(CompNetParams P) {
    extern __shared__ __align__(16) float csm[];
    float* W = csm;  // W1t, W2t, W3t
    float* aff = W + 3 * kCompOut * kCompOut;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* A = aff + 6 * kCompOut + warp * kCompPts * 2 * kCompOut;  // [8][96]: h1, then u
    float* B = A + kCompPts * kCompOut;                              // [8][96]: t
    {
        const float4* src = reinterpret_cast<const float4*>(P.wt + kCompW0Floats);
        float4* dst = reinterpret_cast<float4*>(W);
        for (int e = tid; e < 3 * kCompOut * kCompOut / 4; e += blockDim.x) dst[e] = __ldg(src + e);
        for (int e = tid; e < 6 * kCompOut; e += blockDim.x) aff[e] = __ldg(P.aff + 2 * kCompOut + e);
    }
    __syncthreads();
    const float* W1 = W;
    const float* W2 = W1 + kCompOut * kCompOut;
    const float* W3 = W2 + kCompOut * kCompOut;
    float sc[3][3], sh[3][3];
#pragma unroll
    for (int l = 0; l < 3; ++l)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            sc[l][c] = aff[(l * 2 + 0) * kCompOut + lane + 32 * c];
            sh[l][c] = aff[(l * 2 + 1) * kCompOut + lane + 32 * c];
        }
    const long long n_groups = (P.n + kCompPts - 1) / kCompPts;
    for (long long g = (long long)blockIdx.x * kCompWarpsL234 + warp; g < n_groups;
         g += (long long)gridDim.x * kCompWarpsL234) {
        const long long p0 = g * kCompPts;
        const int np = (int)((P.n - p0) < kCompPts ? (P.n - p0) : kCompPts);
        __syncwarp();
        compnet_stage<kCompOut>(A, P.h1 + p0 * kCompOut, np, lane);
        __syncwarp();
        float acc[kCompPts][3];
        float h1[kCompPts][3];
#pragma unroll
        for (int p = 0; p < kCompPts; ++p)
#pragma unroll
            for (int c = 0; c < 3; ++c) h1[p][c] = A[p * kCompOut + lane + 32 * c];
        compnet_layer<kCompOut, kCompOut>(A, W1, lane, acc);
#pragma unroll
        for (int p = 0; p < kCompPts; ++p)
#pragma unroll
            for (int c = 0; c < 3; ++c) B[p * kCompOut + lane + 32 * c] = lrelu02(fmaf(acc[p][c], sc[0][c], sh[0][c]));
        __syncwarp();
        compnet_layer<kCompOut, kCompOut>(B, W2, lane, acc);
#pragma unroll
        for (int p = 0; p < kCompPts; ++p)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                A[p * kCompOut + lane + 32 * c] = lrelu02(fmaf(acc[p][c], sc[1][c], sh[1][c]) + h1[p][c]);
        __syncwarp();
        compnet_layer<kCompOut, kCompOut>(A, W3, lane, acc);
#pragma unroll
        for (int p = 0; p < kCompPts; ++p) {
            float y[3];
            float ss = 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                y[c] = fmaf(acc[p][c], sc[2][c], sh[2][c]);
                ss = fmaf(y[c], y[c], ss);
            }
            if (P.normalise) {
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
                const float nrm = sqrtf(ss);
#pragma unroll
                for (int c = 0; c < 3; ++c) y[c] = y[c] / nrm * 1.73f;
            }
            if (p < np) {
#pragma unroll
                for (int c = 0; c < 3; ++c) __stcs(P.out + (p0 + p) * kCompOut + lane + 32 * c, y[c]);
            }
        }
    }
}

}  // namespace lafis
