// Private definitions shared by the translation units of liblatentafis_b200.so: the context and latent-batch
// objects behind the opaque handles of include/latentafis_b200.h.  Host-side counterpart of the members of
// PQ::Matcher (matching/matcher.h:34-75).  Nothing here is part of the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/latentafis_b200.h"
#include "device_common.cuh"

namespace lafis {

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) cap = std::max<size_t>(n, 1);
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// pinned host staging buffer that only grows (enrollment calls, ingest)
struct PinnedBuf {
    unsigned char* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost(&p, std::max<size_t>(n, 256));
        if (e == cudaSuccess) cap = std::max<size_t>(n, 256);
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct CommState;  // sharded.cu: NCCL communicator of a context

}  // namespace lafis

struct lafis_latents {
    int n = 0;
    std::vector<int> status;            // LAFIS_OK / LAFIS_LATENT_EMPTY / LAFIS_ERR_LATENT_LAYOUT
    std::vector<int> tex_weighted;      // 0: texture score not read, 1: it is score[28] (weight 0.3), 2 + k: it is score[k], k = 0..2 (weight 1)
    std::vector<int> n_minu_templates;  // as in the file (slot presence for the drivers)
    // host staging, already in device layout
    int lt_stride = 8;
    int max_slot_n = 0;
    std::vector<int> slot_n;            // [3n]
    std::vector<uint32_t> slot_off;     // [3n] padded offsets
    uint32_t tot_minu_padded = 0;
    std::vector<short2> minu_xy;
    std::vector<float> minu_ori;
    std::vector<float> minu_desT;
    std::vector<int> tex_n;             // [n]
    std::vector<short2> tex_xy;         // [n][lt_stride]
    std::vector<float> tex_ori;
    std::vector<float> tex_des;         // [n][lt_stride][96]
    // one pinned arena holding everything above back to back (single H2D copy)
    unsigned char* pinned = nullptr;
    size_t arena_bytes = 0;
    size_t o_slot_n = 0, o_slot_off = 0, o_minu_xy = 0, o_minu_ori = 0, o_minu_desT = 0, o_tex_n = 0, o_tex_xy = 0,
           o_tex_ori = 0, o_tex_des = 0, o_weighted = 0, o_status = 0;
    // device residency: the batch remembers the device and stream it lives on, so that it can be released after
    // its context has been destroyed
    lafis_ctx* owner = nullptr;
    int device = 0;
    unsigned char* d_arena = nullptr;
    bool resident = false;
};

struct lafis_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream_b = nullptr;  // texture chain runs here, concurrently with the minutiae chain on `stream`
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // rare-path kernels (introsort replays of the selection, dense texture graphs: a few hundred long jobs on <= 2 CTAs per
    // SM) run here, next to the following many-CTA kernel of the main stream instead of holding it up
    cudaStream_t stream_c = nullptr;
    cudaEvent_t ev_sel = nullptr, ev_slow = nullptr, ev_gtex = nullptr, ev_gtexd = nullptr;
    bool two_streams = true;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> stage_ev;  // 16 per pipeline chunk + 2 for the tail, grown on demand
    int stage_chunks = 0;               // chunks of the last match
    std::string err;
    int sm_count = 148;
    size_t work_budget = (size_t)24 << 30;
    bool work_budget_fixed = false;  // LAFIS_WORK_BYTES given: otherwise sized from the free device memory per match

    float* d_codebook = nullptr;  // [16][256][6]
    float* d_table = nullptr;     // [50*50]
    float* d_compnet = nullptr;   // CompNet: k-major weights [46080] + folded BatchNorm scale/shift [4][2][96]

    // gallery
    lafis::DeviceGallery gal;
    bool gallery_set = false;          // a gallery (possibly an empty shard) has been made resident
    uint64_t gallery_generation = 0;   // bumped whenever the resident gallery changes
    std::string gallery_dir;           // directory the drivers loaded it from ("" otherwise); cleared with the gallery
    int gallery_dir_shard = 0, gallery_dir_shards = 1;
    uint32_t index_base = 0;
    std::vector<std::string> paths;
    std::vector<int8_t> h_status;
    std::vector<uint32_t> h_minu_off;  // padded
    std::vector<uint16_t> h_minu_n;
    std::vector<uint32_t> h_tex_off;
    int max_nR = 0, max_nRt = 0;
    uint64_t algo_bytes = 0;

    // work buffers
    lafis::DevBuf<float> tex_lut, tex_scale;  // fp32 PQ distance tables of the latent batch (K1) + per-row quantiser scales
    lafis::DevBuf<unsigned char> tex_lut8;    // the row tiles' 8-bit tables in tex_rowmax_kernel's shared-memory layout (128 KB per 32 rows)
    lafis::DevBuf<float> rowmax_val;
    lafis::DevBuf<uint16_t> rowmax_j;
    lafis::DevBuf<float> corr_v;
    lafis::DevBuf<uint32_t> corr_ij;
    lafis::DevBuf<int> corr_n;
    lafis::DevBuf<float> sim;            // S matrices of the current chunk
    lafis::DevBuf<int> slow_jobs;        // selection jobs that need the introsort replay
    int* d_slow_count = nullptr;
    lafis::DevBuf<int> ov_minu, ov_tex;  // overflow job lists of the sparse graph kernels
    lafis::DevBuf<int> ov_minu2;         // ... of the second-chance minutiae kernel (graph_minu_mid_kernel): the dense kernel's list
    int* d_ov_count = nullptr;           // [4]: 0 minutiae, 1 texture, 2 minutiae second chance
    lafis::DevBuf<float> comp;
    lafis::DevBuf<float> final_scores;
    lafis::DevBuf<short4> corr_xy;       // lafis_correspondences: surviving correspondences of the 3 minutiae components
    lafis::DevBuf<int> corr_xy_n;
    lafis::DevBuf<unsigned long long> keys_a, keys_b;
    lafis::DevBuf<lafis::HitDev> hits;
    lafis::DevBuf<unsigned char> lat_arena;  // for non-resident latent batches
    lafis::DevBuf<float> compnet_h1;         // CompNet: output of layer1, [n][96]
    // persistent staging of the enrollment calls (lafis_pq_encode / lafis_compress_descriptors with host pointers)
    lafis::DevBuf<unsigned char> enroll_in, enroll_out;
    lafis::PinnedBuf enroll_pin;
    // oversized minutiae pairs (minu_big.cuh): work list and matrices in HBM
    lafis::DevBuf<int> big_jobs, big_slow;
    lafis::DevBuf<unsigned long long> big_soff;
    lafis::DevBuf<float> big_S;
    lafis::DevBuf<uint32_t> big_keys, big_order;
    int* d_job_counter = nullptr;
    unsigned long long* d_slow = nullptr;  // [12] counters (8: jobs of the second-chance minutiae graph kernel): 0 minutiae introsort replays, 1 texture top-200 replays,
                                           //     4..7 texture row-max: queued, exact evaluations, overflowed, templates

    // multi-GPU (sharded.cu)
    lafis::CommState* comm = nullptr;
    lafis::DevBuf<lafis::HitDev> gathered, merged;  // all-gathered per-shard rank lists, merged global lists
    lafis::DevBuf<float> gather_scores;             // root: per-rank score blocks of the N-vs-N gather

    lafis_stats stats{};
};

// Several contexts of ONE process, one per device, joined by an NCCL communicator: the gallery is sharded over them
// (contiguous index ranges) and every match ends in the all-gather + merge of lafis_match_sharded.
struct lafis_group {
    std::vector<lafis_ctx*> ctx;
    std::string err;
    std::vector<uint32_t> base;  // [n+1] first global gallery index of every shard
};

namespace lafis {

// records a message on the context (or for lafis_last_error(NULL) before one exists) and returns `code`
int fail(lafis_ctx* c, int code, const char* fmt, ...);

#define LAFIS_CUDA(c, expr)                                                                                  \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            return lafis::fail((c), LAFIS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                               __FILE__, __LINE__);                                                          \
    } while (0)

// lafis_api.cu: enqueue a whole match on the context's stream(s); results stay in c->final_scores / c->comp / c->hits
int run_match(lafis_ctx* c, lafis_latents* L, int topk);
void collect_times(lafis_ctx* c);
void comm_release(lafis_ctx* c);  // sharded.cu, called by lafis_destroy

}  // namespace lafis
