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

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

}  // namespace lafis
