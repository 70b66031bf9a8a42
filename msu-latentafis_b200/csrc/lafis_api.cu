// C-ABI implementation (include/latentafis_b200.h): context, gallery ingest into HBM, latent
// staging, the chunked match pipeline and rank lists.  Host-side counterpart of PQ::Matcher
// (matching/matcher.h:34-75, matching/matcher.cpp:31-337); all arithmetic of the hot path happens in
// the kernels included below.  There is no CPU implementation of any scoring stage in this file.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "lafis_internal.h"
#include "compnet.cuh"
#include "dat_format.h"
#include "graph_prune.cuh"
#include "graph_sparse.cuh"
#include "minu_big.cuh"
#include "minu_plan.h"
#include "minu_sim.cuh"
#include "misc_kernels.cuh"
#include "tex_rowmax.cuh"

using namespace lafis;

// ---------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------
namespace {

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

std::string g_create_error;  // failures before a context exists (lafis_last_error(NULL))

}  // namespace

namespace lafis {
int fail(lafis_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    else g_create_error = buf;
    return code;
}
}  // namespace lafis

namespace {


// Work buffers of a pipeline chunk: as much as the device can spare once the gallery is resident (the similarity
// matrices alone are ~146 KB per (latent, template) pair, so a 256-latent batch wants tens of GB to keep its chunks above
// a few thousand templates).  Sized when the gallery changes, not per match: cudaMemGetInfo takes a driver lock that
// monitoring tools (nvidia-smi polling) hold for milliseconds at a time.
void size_work_budget(lafis_ctx* c) {
    if (c->work_budget_fixed) return;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return;
    const size_t held = sizeof(float) * (c->sim.cap + c->rowmax_val.cap + c->corr_v.cap) + 2 * c->rowmax_j.cap +
                        4 * (c->corr_ij.cap + c->corr_n.cap + c->slow_jobs.cap + c->ov_minu.cap + c->ov_minu2.cap + c->ov_tex.cap);
    c->work_budget = std::min<size_t>(std::max<size_t>((free_b + held) / 2, (size_t)1 << 30), (size_t)64 << 30);
}

void free_gallery(lafis_ctx* c) {
    DeviceGallery& g = c->gal;
    cudaFree(g.minu_off);
    cudaFree(g.minu_n);
    cudaFree(g.minu_xy);
    cudaFree(g.minu_ori);
    cudaFree(g.minu_desT);
    cudaFree(g.tex_off);
    cudaFree(g.tex_xy);
    cudaFree(g.tex_ori);
    cudaFree(g.tex_codes);
    cudaFree(g.status);
    g = DeviceGallery();
    c->paths.clear();
    c->h_status.clear();
    c->h_minu_off.clear();
    c->h_minu_n.clear();
    c->h_tex_off.clear();
    c->max_nR = c->max_nRt = 0;
    c->algo_bytes = 0;
    c->gallery_set = false;
    c->gallery_dir.clear();
    ++c->gallery_generation;
}

int create_common(const float* codewords, int device, lafis_ctx** out) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev)
        return fail(nullptr, LAFIS_ERR_CUDA, "no CUDA device %d (%s)", device, cudaGetErrorString(e));
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, LAFIS_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)  // built for sm_100a only; no other code path exists
        return fail(nullptr, LAFIS_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    lafis_ctx* c = new lafis_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->work_budget_fixed = false;
    if (const char* s = getenv("LAFIS_WORK_BYTES")) {
        c->work_budget = (size_t)strtoull(s, nullptr, 10);
        c->work_budget_fixed = true;
    }
    cudaError_t last = cudaSuccess;
    const char* what = "";
#define TRY(expr) (ok = ok && ((last = (expr)) == cudaSuccess || (what = #expr, false)))
    bool ok = true;
    TRY(cudaSetDevice(device));
    ok = ok && cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&c->stream_b, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) == cudaSuccess;
    {
        int lo_pri = 0, hi_pri = 0;  // the few rare-path CTAs should take free SM resources ahead of the main stream's
        cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
        ok = ok && cudaStreamCreateWithPriority(&c->stream_c, cudaStreamNonBlocking, hi_pri) == cudaSuccess;
        for (cudaEvent_t* ev : {&c->ev_sel, &c->ev_slow, &c->ev_gtex, &c->ev_gtexd})
            ok = ok && cudaEventCreateWithFlags(ev, cudaEventDisableTiming) == cudaSuccess;
    }
    if (const char* e2 = getenv("LAFIS_STREAMS")) c->two_streams = atoi(e2) >= 2;
    ok = ok && cudaEventCreate(&c->ev0) == cudaSuccess && cudaEventCreate(&c->ev1) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_codebook, sizeof(float) * kSubs * kClusters * kSubDim) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_table, sizeof(float) * kTableN * kTableN) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_job_counter, sizeof(int)) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_ov_count, 4 * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_slow_count, sizeof(int)) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_slow, 12 * sizeof(unsigned long long)) == cudaSuccess;
    if (ok) {
        // matcher.cpp:49-56: table[i*50+j] = (float)sqrt((16 i)^2 + (16 j)^2), square root in double
        std::vector<float> table(kTableN * kTableN);
        for (int i = 0; i < kTableN; ++i)
            for (int j = 0; j < kTableN; ++j)
                table[i * kTableN + j] = (float)std::sqrt((i * 16.0) * (i * 16.0) + (j * 16.0) * (j * 16.0));
        ok = cudaMemcpy(c->d_table, table.data(), sizeof(float) * table.size(), cudaMemcpyHostToDevice) == cudaSuccess;
        ok = ok && cudaMemcpy(c->d_codebook, codewords, sizeof(float) * kSubs * kClusters * kSubDim,
                              cudaMemcpyHostToDevice) == cudaSuccess;
        ok = ok && cudaMemset(c->d_slow, 0, 12 * sizeof(unsigned long long)) == cudaSuccess;
        TRY(cudaFuncSetAttribute(tex_rowmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRowmaxSmem));
        TRY(cudaFuncSetAttribute(minu_sim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
        TRY(cudaFuncSetAttribute(minu_sim_jobs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
        TRY(cudaFuncSetAttribute(minu_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
        TRY(cudaFuncSetAttribute(minu_select_slow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
        TRY(cudaFuncSetAttribute(graph_minu_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGraphMinuSmem));
        TRY(cudaFuncSetAttribute(graph_tex_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGraphTexSmem));
        TRY(cudaFuncSetAttribute(graph_minu_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(SparseWork<false>)));
        TRY(cudaFuncSetAttribute(graph_tex_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(SparseWork<true>)));
        TRY(cudaFuncSetAttribute(compnet_l1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)compnet_l1_smem_bytes()));
        TRY(cudaFuncSetAttribute(compnet_l234_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)compnet_l234_smem_bytes()));
    }
#undef TRY
    if (!ok) {
        if (last == cudaSuccess) last = cudaGetLastError();
        fail(nullptr, LAFIS_ERR_CUDA, "context setup failed: %s %s", what, cudaGetErrorString(last));
        lafis_destroy(c);
        return LAFIS_ERR_CUDA;
    }
    *out = c;
    return LAFIS_OK;
}

// Everything a packed gallery needs on the host before the device re-layout.
struct IngestPlan {
    std::vector<uint32_t> dst_minu_off, dst_tex_off;
    std::vector<uint16_t> minu_n;
    std::vector<int8_t> status;
    bool tex_truncated = false;
};

}  // namespace

// ---------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------
extern "C" {

const char* lafis_version(void) { return "latentafis_b200 0.1 (sm_100a)"; }

int lafis_create_from_codebook(const float* codewords, int subs, int clusters, int sub_dim, int device,
                               lafis_ctx** out) {
    if (!codewords || !out) return LAFIS_ERR_ARG;
    if (subs != kSubs || clusters != kClusters || sub_dim != kSubDim) return LAFIS_ERR_CODEBOOK;
    return create_common(codewords, device, out);
}

int lafis_create(const char* codebook_path, int device, lafis_ctx** out) {
    if (!codebook_path || !out) return LAFIS_ERR_ARG;
    std::vector<float> cw;
    int subs = 0, clusters = 0, sub_dim = 0;
    int rc = read_codebook(codebook_path, cw, subs, clusters, sub_dim);
    if (rc == -3) return LAFIS_ERR_IO;
    if (rc != 0) return LAFIS_ERR_CODEBOOK;
    return lafis_create_from_codebook(cw.data(), subs, clusters, sub_dim, device, out);
}

void lafis_destroy(lafis_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    free_gallery(c);
    c->rowmax_val.release();
    c->tex_lut.release();
    c->tex_scale.release();
    c->tex_lut8.release();
    c->rowmax_j.release();
    c->corr_v.release();
    c->corr_ij.release();
    c->corr_n.release();
    c->sim.release();
    c->slow_jobs.release();
    cudaFree(c->d_slow_count);
    c->ov_minu.release();
    c->ov_minu2.release();
    c->ov_tex.release();
    cudaFree(c->d_ov_count);
    c->comp.release();
    c->final_scores.release();
    c->keys_a.release();
    c->keys_b.release();
    c->hits.release();
    c->lat_arena.release();
    c->compnet_h1.release();
    c->corr_xy.release();
    c->corr_xy_n.release();
    c->enroll_in.release();
    c->enroll_out.release();
    c->enroll_pin.release();
    comm_release(c);
    c->gathered.release();
    c->merged.release();
    c->gather_scores.release();
    c->big_jobs.release();
    c->big_slow.release();
    c->big_soff.release();
    c->big_S.release();
    c->big_keys.release();
    c->big_order.release();
    cudaFree(c->d_codebook);
    cudaFree(c->d_table);
    cudaFree(c->d_compnet);
    cudaFree(c->d_job_counter);
    cudaFree(c->d_slow);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (cudaEvent_t e : c->stage_ev) cudaEventDestroy(e);
    if (c->stream_b) cudaStreamDestroy(c->stream_b);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->stream_c) cudaStreamDestroy(c->stream_c);
    for (cudaEvent_t ev : {c->ev_sel, c->ev_slow, c->ev_gtex, c->ev_gtexd})
        if (ev) cudaEventDestroy(ev);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* lafis_last_error(const lafis_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

void* lafis_stream(const lafis_ctx* c) { return c ? (void*)c->stream : nullptr; }

int lafis_set_streams(lafis_ctx* c, int n_streams) {
    if (!c || n_streams < 1 || n_streams > 2) return fail(c, LAFIS_ERR_ARG, "n_streams must be 1 or 2");
    c->two_streams = n_streams == 2;
    return LAFIS_OK;
}

int lafis_get_stats(const lafis_ctx* c, lafis_stats* out) {
    if (!c || !out) return LAFIS_ERR_ARG;
    *out = c->stats;
    return LAFIS_OK;
}

// ---------------------------------------------------------------------------------------------------
// gallery ingest
// ---------------------------------------------------------------------------------------------------
int lafis_gallery_set_packed(lafis_ctx* c, const lafis_packed_gallery* g, uint32_t index_base) {
    if (!c || !g || g->n_templates < 0 || !g->minu_off || !g->tex_off) return fail(c, LAFIS_ERR_ARG, "bad gallery argument");
    LAFIS_CUDA(c, cudaSetDevice(c->device));
    LAFIS_CUDA(c, cudaStreamSynchronize(c->stream));
    free_gallery(c);
    const int n = g->n_templates;
    c->index_base = index_base;
    c->gallery_set = true;  // an empty shard is a gallery too: matches return empty rank lists
    if (n == 0) return LAFIS_OK;

    // ---- host plan ----
    std::vector<uint32_t> dst_minu_off(n + 1), dst_tex_off(n + 1);
    std::vector<uint16_t> minu_n(n);
    std::vector<int8_t> status(n);
    bool tex_truncated = false;
    int max_nR = 0, max_nRt = 0;
    uint64_t algo = 0;
    dst_minu_off[0] = dst_tex_off[0] = 0;
    for (int t = 0; t < n; ++t) {
        const uint32_t nm = g->minu_off[t + 1] - g->minu_off[t];
        uint32_t nt = g->tex_off[t + 1] - g->tex_off[t];
        if (nm > (uint32_t)kMaxMinutiae) return fail(c, LAFIS_ERR_UNSUPPORTED_SIZE, "template %d: %u minutiae", t, nm);
        if (nt > (uint32_t)kMaxTexture) {  // matcher.cpp:546-547
            nt = kMaxTexture;
            tex_truncated = true;
        }
        minu_n[t] = (uint16_t)nm;
        dst_minu_off[t + 1] = dst_minu_off[t] + ((nm + 3u) & ~3u);
        dst_tex_off[t + 1] = dst_tex_off[t] + nt;
        max_nR = std::max(max_nR, (int)nm);
        max_nRt = std::max(max_nRt, (int)nt);
        int8_t st = g->status ? g->status[t] : (int8_t)LAFIS_TPL_OK;
        if ((st == LAFIS_TPL_OK || st == LAFIS_TPL_TRUNCATED) && nm == 0 && nt == 0) st = LAFIS_TPL_EMPTY;
        status[t] = st;
        algo += 392ull * nm + 24ull * nt;
    }
    const uint32_t src_minu_tot = g->minu_off[n], src_tex_tot = g->tex_off[n];
    const uint32_t dst_minu_tot = dst_minu_off[n], dst_tex_tot = dst_tex_off[n];

    // ---- source arrays on the device ----
    const int16_t *sx = g->minu_x, *sy = g->minu_y, *tx = g->tex_x, *ty = g->tex_y;
    const float *sori = g->minu_ori, *sdes = g->minu_des, *tori = g->tex_ori;
    const uint8_t* tcodes = g->tex_codes;
    std::vector<void*> temps;
    auto cleanup = [&]() {
        for (void* p : temps) cudaFree(p);
        temps.clear();
    };
    auto up = [&](const void* src, size_t bytes, const void** dst) -> cudaError_t {
        void* d = nullptr;
        cudaError_t e = cudaMalloc(&d, std::max<size_t>(bytes, 16));
        if (e != cudaSuccess) return e;
        temps.push_back(d);
        if (bytes) e = cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, c->stream);
        *dst = d;
        return e;
    };
#define LAFIS_CUDA_T(expr)                                                                              \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            cleanup();                                                                                  \
            free_gallery(c);                                                                            \
            return fail(c, LAFIS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),    \
                        __FILE__, __LINE__);                                                            \
        }                                                                                               \
    } while (0)

    // texture truncation (> 1000 points) needs a per-template gather; do it on the host side of the
    // offsets only: the copy kernels below take src and dst offsets separately.
    if (!g->on_device) {
        LAFIS_CUDA_T(up(g->minu_x, 2ull * src_minu_tot, (const void**)&sx));
        LAFIS_CUDA_T(up(g->minu_y, 2ull * src_minu_tot, (const void**)&sy));
        LAFIS_CUDA_T(up(g->minu_ori, 4ull * src_minu_tot, (const void**)&sori));
        LAFIS_CUDA_T(up(g->minu_des, 4ull * kDesLen * src_minu_tot, (const void**)&sdes));
        LAFIS_CUDA_T(up(g->tex_x, 2ull * src_tex_tot, (const void**)&tx));
        LAFIS_CUDA_T(up(g->tex_y, 2ull * src_tex_tot, (const void**)&ty));
        LAFIS_CUDA_T(up(g->tex_ori, 4ull * src_tex_tot, (const void**)&tori));
        LAFIS_CUDA_T(up(g->tex_codes, 16ull * src_tex_tot, (const void**)&tcodes));
    }
    const uint32_t *d_src_minu_off = nullptr, *d_src_tex_off = nullptr;
    LAFIS_CUDA_T(up(g->minu_off, 4ull * (n + 1), (const void**)&d_src_minu_off));
    LAFIS_CUDA_T(up(g->tex_off, 4ull * (n + 1), (const void**)&d_src_tex_off));

    // ---- resident arrays ----
    DeviceGallery& G = c->gal;
    G.n = n;
    LAFIS_CUDA_T(cudaMalloc(&G.minu_off, 4ull * (n + 1)));
    LAFIS_CUDA_T(cudaMalloc(&G.minu_n, 2ull * n));
    LAFIS_CUDA_T(cudaMalloc(&G.minu_xy, sizeof(short2) * std::max<size_t>(dst_minu_tot, 4)));
    LAFIS_CUDA_T(cudaMalloc(&G.minu_ori, 4ull * std::max<size_t>(dst_minu_tot, 4)));
    LAFIS_CUDA_T(cudaMalloc(&G.minu_desT, 4ull * kDesLen * std::max<size_t>(dst_minu_tot, 4)));
    LAFIS_CUDA_T(cudaMalloc(&G.tex_off, 4ull * (n + 1)));
    LAFIS_CUDA_T(cudaMalloc(&G.tex_xy, sizeof(short2) * std::max<size_t>(dst_tex_tot, 4)));
    LAFIS_CUDA_T(cudaMalloc(&G.tex_ori, 4ull * std::max<size_t>(dst_tex_tot, 4)));
    LAFIS_CUDA_T(cudaMalloc(&G.tex_codes, 16ull * ((size_t)dst_tex_tot + 32)));
    LAFIS_CUDA_T(cudaMalloc(&G.status, (size_t)n));
    LAFIS_CUDA_T(cudaMemcpyAsync(G.minu_off, dst_minu_off.data(), 4ull * (n + 1), cudaMemcpyHostToDevice, c->stream));
    LAFIS_CUDA_T(cudaMemcpyAsync(G.minu_n, minu_n.data(), 2ull * n, cudaMemcpyHostToDevice, c->stream));
    LAFIS_CUDA_T(cudaMemcpyAsync(G.tex_off, dst_tex_off.data(), 4ull * (n + 1), cudaMemcpyHostToDevice, c->stream));
    LAFIS_CUDA_T(cudaMemcpyAsync(G.status, status.data(), (size_t)n, cudaMemcpyHostToDevice, c->stream));
    LAFIS_CUDA_T(cudaMemsetAsync(G.tex_codes + dst_tex_tot, 0, 16ull * 32, c->stream));

    RelayoutParams rp;
    rp.n_templates = n;
    rp.src_off = d_src_minu_off;
    rp.dst_off = G.minu_off;
    rp.x = sx;
    rp.y = sy;
    rp.ori = sori;
    rp.des = sdes;
    rp.out_xy = G.minu_xy;
    rp.out_ori = G.minu_ori;
    rp.out_desT = G.minu_desT;
    relayout_minutiae_kernel<<<n, 256, 0, c->stream>>>(rp);
    TexCopyParams tp;
    tp.src_off = d_src_tex_off;
    tp.dst_off = G.tex_off;
    tp.x = tx;
    tp.y = ty;
    tp.ori = tori;
    tp.codes = reinterpret_cast<const uint4*>(tcodes);
    tp.codes_aligned = ((uintptr_t)tcodes & 15u) == 0;
    tp.out_xy = G.tex_xy;
    tp.out_ori = G.tex_ori;
    tp.out_codes = G.tex_codes;
    copy_texture_kernel<<<n, 256, 0, c->stream>>>(tp);
    c->stats.kernel_launches += 2;
    LAFIS_CUDA_T(cudaGetLastError());
    LAFIS_CUDA_T(cudaStreamSynchronize(c->stream));
    cleanup();
#undef LAFIS_CUDA_T
    (void)tex_truncated;

    c->h_status.swap(status);
    c->h_minu_off.swap(dst_minu_off);
    c->h_minu_n.swap(minu_n);
    c->h_tex_off.swap(dst_tex_off);
    c->max_nR = max_nR;
    c->max_nRt = max_nRt;
    c->algo_bytes = algo;
    c->paths.assign(n, std::string());
    size_work_budget(c);
    return LAFIS_OK;
}

int lafis_gallery_load_files(lafis_ctx* c, const char* const* paths, int n, int shard_rank, int shard_count) {
    if (!c || (!paths && n > 0) || n < 0 || shard_count <= 0 || shard_rank < 0 || shard_rank >= shard_count)
        return fail(c, LAFIS_ERR_ARG, "bad argument");
    if (n == 0) return fail(c, LAFIS_ERR_NO_TEMPLATES, "no rolled templates");
    const long long lo = (long long)n * shard_rank / shard_count, hi = (long long)n * (shard_rank + 1) / shard_count;
    const int m = (int)(hi - lo);
    // The reference re-parses every rolled file for every latent (matcher.cpp:173/:278); here every file is parsed once,
    // by all host cores.  Each parser thread owns a contiguous range of files and a ring of pinned staging buffers: it
    // packs the templates it has parsed into the current buffer and, when that is full, reserves a region of ONE device
    // pool and ships the buffer there with an asynchronous copy on its own stream while it goes on parsing into the next
    // buffer.  No pageable copy, no consolidated host copy of the gallery, and the upload overlaps the parsing.
    LAFIS_CUDA(c, cudaSetDevice(c->device));
    LAFIS_CUDA(c, cudaStreamSynchronize(c->stream));
    free_gallery(c);
    c->index_base = (uint32_t)lo;
    c->gallery_set = true;
    if (m == 0) return LAFIS_OK;  // an empty shard (more ranks than files)

    const bool timing = getenv("LAFIS_INGEST_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double, std::milli>(now() - t).count();
    };
    const auto t_all = now();
    const int n_threads = std::max(1, std::min({(int)std::thread::hardware_concurrency(), 64, m / 32 + 1}));
    auto range_lo = [&](int i) { return (int)((long long)m * i / n_threads); };
    auto run_parallel = [&](auto&& fn) {
        std::vector<std::thread> pool;
        for (int i = 1; i < n_threads; ++i) pool.emplace_back(fn, i);
        fn(0);
        for (std::thread& th : pool) th.join();
    };

    // pool capacity: what a file can contribute is bounded by its size (+ alignment padding per piece)
    std::vector<unsigned long long> part_bytes(n_threads, 0);
    run_parallel([&](int i) {
        unsigned long long tot = 0;
        for (int t = range_lo(i); t < range_lo(i + 1); ++t) {
            std::error_code ec;
            const auto sz = std::filesystem::file_size(paths[lo + t], ec);
            tot += (ec ? 0ull : (unsigned long long)sz) + 8 * 16;
        }
        part_bytes[i] = tot;
    });
    unsigned long long pool_cap = 0;
    for (unsigned long long b : part_bytes) pool_cap += b;
    // staging slots: 1/8 of a thread's share, between 1 MB (> the largest template: 2000 minutiae + 1000 texture points
    // = 0.8 MB) and 8 MB - pinning memory costs time too, and a 2,000-file gallery should not pin 400 MB for it
    const size_t kSlotBytes = std::min<size_t>((size_t)8 << 20, std::max<size_t>((size_t)1 << 20, (size_t)(pool_cap / n_threads / 8 + 0xfffff) & ~(size_t)0xfffff));
    constexpr int kRing = 3;
    unsigned char* d_pool = nullptr;
    LAFIS_CUDA(c, cudaMalloc(&d_pool, std::max<unsigned long long>(pool_cap, 16)));
    std::atomic<unsigned long long> pool_used{0};

    std::vector<PoolRec> rec(m);
    std::vector<int8_t> status(m);
    std::vector<std::string> kept(m);
    std::vector<int> thread_err(n_threads, 0);
    run_parallel([&](int i) {
        cudaStream_t s_ = nullptr;
        unsigned char* slot[kRing] = {nullptr, nullptr, nullptr};
        cudaEvent_t done[kRing] = {nullptr, nullptr, nullptr};
        bool ok = cudaSetDevice(c->device) == cudaSuccess && cudaStreamCreateWithFlags(&s_, cudaStreamNonBlocking) == cudaSuccess;
        for (int k = 0; k < kRing && ok; ++k)
            ok = cudaMallocHost(&slot[k], kSlotBytes) == cudaSuccess && cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming) == cudaSuccess;
        int cur = 0, first_in_slot = range_lo(i);
        size_t used = 0;
        auto flush = [&](int t_end) {  // ship templates [first_in_slot, t_end) packed in slot[cur]
            if (!ok || used == 0) {
                first_in_slot = t_end;
                return;
            }
            const unsigned long long at = pool_used.fetch_add(used);
            if (at + used > pool_cap) {
                ok = false;
                return;
            }
            ok = cudaMemcpyAsync(d_pool + at, slot[cur], used, cudaMemcpyHostToDevice, s_) == cudaSuccess &&
                 cudaEventRecord(done[cur], s_) == cudaSuccess;
            for (int t = first_in_slot; t < t_end; ++t) {
                rec[t].minu_at += at;
                rec[t].tex_at += at;
            }
            first_in_slot = t_end;
            cur = (cur + 1) % kRing;
            used = 0;
            if (ok) ok = cudaEventSynchronize(done[cur]) == cudaSuccess;  // the buffer about to be reused (no-op when never used)
        };
        for (int t = range_lo(i); t < range_lo(i + 1) && ok; ++t) {
            RolledTemplate R;
            read_rolled_dat(paths[lo + t], R);
            kept[t] = paths[lo + t];
            status[t] = (int8_t)R.status;
            const size_t nm = R.minu.x.size();
            const size_t nt = std::min<size_t>(R.tex.x.size(), (size_t)kMaxTexture);  // matcher.cpp:546-547
            const size_t need = pool_minu_bytes(nm) + pool_tex_bytes(nt);
            if (need > kSlotBytes) {
                ok = false;
                break;
            }
            if (used + need > kSlotBytes) flush(t);
            unsigned char* p = slot[cur] + used;
            rec[t].minu_at = used;
            rec[t].n_minu = (uint32_t)nm;
            std::memcpy(p, R.minu.x.data(), 2 * nm);
            std::memcpy(p + pool_align(2 * nm), R.minu.y.data(), 2 * nm);
            std::memcpy(p + 2 * pool_align(2 * nm), R.minu.ori.data(), 4 * nm);
            std::memcpy(p + 2 * pool_align(2 * nm) + pool_align(4 * nm), R.minu.des.data(), 4 * (size_t)kDesLen * nm);
            p += pool_minu_bytes(nm);
            rec[t].tex_at = used + pool_minu_bytes(nm);
            rec[t].n_tex = (uint32_t)nt;
            std::memcpy(p, R.tex.x.data(), 2 * nt);
            std::memcpy(p + pool_align(2 * nt), R.tex.y.data(), 2 * nt);
            std::memcpy(p + 2 * pool_align(2 * nt), R.tex.ori.data(), 4 * nt);
            std::memcpy(p + 2 * pool_align(2 * nt) + pool_align(4 * nt), R.tex.codes.data(), 16 * nt);
            used += need;
        }
        flush(range_lo(i + 1));
        if (s_) ok = cudaStreamSynchronize(s_) == cudaSuccess && ok;
        for (int k = 0; k < kRing; ++k) {
            if (done[k]) cudaEventDestroy(done[k]);
            if (slot[k]) cudaFreeHost(slot[k]);
        }
        if (s_) cudaStreamDestroy(s_);
        thread_err[i] = ok ? 0 : 1;
    });
    const double stream_ms = ms_since(t_all);
    for (int e : thread_err)
        if (e) {
            cudaFree(d_pool);
            free_gallery(c);
            return fail(c, LAFIS_ERR_CUDA, "gallery ingest: staging or host-to-device copy failed: %s",
                        cudaGetErrorString(cudaGetLastError()));
        }

    // ---- resident arrays + re-layout from the pool ----
    const auto t_set = now();
    std::vector<uint32_t> dst_minu_off(m + 1), dst_tex_off(m + 1);
    std::vector<uint16_t> minu_n(m);
    int max_nR = 0, max_nRt = 0;
    uint64_t algo = 0;
    dst_minu_off[0] = dst_tex_off[0] = 0;
    for (int t = 0; t < m; ++t) {
        const uint32_t nm = rec[t].n_minu, nt = rec[t].n_tex;
        minu_n[t] = (uint16_t)nm;
        dst_minu_off[t + 1] = dst_minu_off[t] + ((nm + 3u) & ~3u);
        dst_tex_off[t + 1] = dst_tex_off[t] + nt;
        max_nR = std::max(max_nR, (int)nm);
        max_nRt = std::max(max_nRt, (int)nt);
        if ((status[t] == LAFIS_TPL_OK || status[t] == LAFIS_TPL_TRUNCATED) && nm == 0 && nt == 0) status[t] = LAFIS_TPL_EMPTY;
        algo += 392ull * nm + 24ull * nt;
        if ((size_t)dst_minu_off[t + 1] > 0x7fffffffu || (size_t)dst_tex_off[t + 1] > 0x7fffffffu) {
            cudaFree(d_pool);
            free_gallery(c);
            return fail(c, LAFIS_ERR_ARG, "gallery shard too large for 32-bit offsets");
        }
    }
    const uint32_t dst_minu_tot = dst_minu_off[m], dst_tex_tot = dst_tex_off[m];
    DeviceGallery& G = c->gal;
    PoolRec* d_rec = nullptr;
    auto bail = [&](cudaError_t e, const char* what) {
        cudaFree(d_pool);
        cudaFree(d_rec);
        free_gallery(c);
        return fail(c, LAFIS_ERR_CUDA, "gallery ingest: %s failed: %s", what, cudaGetErrorString(e));
    };
#define LAFIS_CUDA_I(expr)                           \
    do {                                             \
        cudaError_t e__ = (expr);                    \
        if (e__ != cudaSuccess) return bail(e__, #expr); \
    } while (0)
    G.n = m;
    LAFIS_CUDA_I(cudaMalloc(&d_rec, sizeof(PoolRec) * (size_t)m));
    LAFIS_CUDA_I(cudaMalloc(&G.minu_off, 4ull * (m + 1)));
    LAFIS_CUDA_I(cudaMalloc(&G.minu_n, 2ull * m));
    LAFIS_CUDA_I(cudaMalloc(&G.minu_xy, sizeof(short2) * std::max<size_t>(dst_minu_tot, 4)));
    LAFIS_CUDA_I(cudaMalloc(&G.minu_ori, 4ull * std::max<size_t>(dst_minu_tot, 4)));
    LAFIS_CUDA_I(cudaMalloc(&G.minu_desT, 4ull * kDesLen * std::max<size_t>(dst_minu_tot, 4)));
    LAFIS_CUDA_I(cudaMalloc(&G.tex_off, 4ull * (m + 1)));
    LAFIS_CUDA_I(cudaMalloc(&G.tex_xy, sizeof(short2) * std::max<size_t>(dst_tex_tot, 4)));
    LAFIS_CUDA_I(cudaMalloc(&G.tex_ori, 4ull * std::max<size_t>(dst_tex_tot, 4)));
    LAFIS_CUDA_I(cudaMalloc(&G.tex_codes, 16ull * ((size_t)dst_tex_tot + 32)));
    LAFIS_CUDA_I(cudaMalloc(&G.status, (size_t)m));
    LAFIS_CUDA_I(cudaMemcpyAsync(d_rec, rec.data(), sizeof(PoolRec) * (size_t)m, cudaMemcpyHostToDevice, c->stream));
    LAFIS_CUDA_I(cudaMemcpyAsync(G.minu_off, dst_minu_off.data(), 4ull * (m + 1), cudaMemcpyHostToDevice, c->stream));
    LAFIS_CUDA_I(cudaMemcpyAsync(G.minu_n, minu_n.data(), 2ull * m, cudaMemcpyHostToDevice, c->stream));
    LAFIS_CUDA_I(cudaMemcpyAsync(G.tex_off, dst_tex_off.data(), 4ull * (m + 1), cudaMemcpyHostToDevice, c->stream));
    LAFIS_CUDA_I(cudaMemcpyAsync(G.status, status.data(), (size_t)m, cudaMemcpyHostToDevice, c->stream));
    LAFIS_CUDA_I(cudaMemsetAsync(G.tex_codes + dst_tex_tot, 0, 16ull * 32, c->stream));
    PoolRelayoutParams rp;
    rp.pool = d_pool;
    rp.rec = d_rec;
    rp.minu_off = G.minu_off;
    rp.tex_off = G.tex_off;
    rp.minu_xy = G.minu_xy;
    rp.minu_ori = G.minu_ori;
    rp.minu_desT = G.minu_desT;
    rp.tex_xy = G.tex_xy;
    rp.tex_ori = G.tex_ori;
    rp.tex_codes = G.tex_codes;
    relayout_minutiae_pool_kernel<<<m, 256, 0, c->stream>>>(rp);
    copy_texture_pool_kernel<<<m, 256, 0, c->stream>>>(rp);
    c->stats.kernel_launches += 2;
    LAFIS_CUDA_I(cudaGetLastError());
    LAFIS_CUDA_I(cudaStreamSynchronize(c->stream));
#undef LAFIS_CUDA_I
    cudaFree(d_pool);
    cudaFree(d_rec);
    c->h_status.swap(status);
    c->h_minu_off.swap(dst_minu_off);
    c->h_minu_n.swap(minu_n);
    c->h_tex_off.swap(dst_tex_off);
    c->max_nR = max_nR;
    c->max_nRt = max_nRt;
    c->algo_bytes = algo;
    c->paths.swap(kept);
    size_work_budget(c);
    if (timing)
        fprintf(stderr,
                "lafis ingest: %d files, %d threads: parse + pinned upload %.1f ms (%.2f GB staged), re-layout %.1f ms, total %.1f ms "
                "= %.0f templates/s\n",
                m, n_threads, stream_ms, (double)pool_used.load() / 1e9, ms_since(t_set), ms_since(t_all),
                m / (ms_since(t_all) / 1e3));
    return LAFIS_OK;
}

int lafis_gallery_load_dir(lafis_ctx* c, const char* dir, int shard_rank, int shard_count) {
    if (!c || !dir) return fail(c, LAFIS_ERR_ARG, "bad argument");
    std::vector<std::string> files = list_dat_files(dir);
    if (files.empty()) return fail(c, LAFIS_ERR_NO_TEMPLATES, "No rolled templates found in directory: %s", dir);
    std::vector<const char*> ptrs(files.size());
    for (size_t i = 0; i < files.size(); ++i) ptrs[i] = files[i].c_str();
    return lafis_gallery_load_files(c, ptrs.data(), (int)ptrs.size(), shard_rank, shard_count);
}

int lafis_gallery_size(const lafis_ctx* c) { return c ? c->gal.n : 0; }
const char* lafis_gallery_path(const lafis_ctx* c, int i) {
    return (c && i >= 0 && i < (int)c->paths.size()) ? c->paths[i].c_str() : "";
}
int lafis_gallery_status(const lafis_ctx* c, int i) {
    return (c && i >= 0 && i < (int)c->h_status.size()) ? c->h_status[i] : LAFIS_TPL_FAILED;
}
uint64_t lafis_gallery_bytes(const lafis_ctx* c) { return c ? c->algo_bytes : 0; }

int lafis_gallery_get_template(const lafis_ctx* cc, int t, int* n_minu, int16_t* mx, int16_t* my, float* mori,
                               float* mdes, int* n_tex, int16_t* tx, int16_t* ty, float* tori, uint8_t* tcodes) {
    lafis_ctx* c = const_cast<lafis_ctx*>(cc);
    if (!c || t < 0 || t >= c->gal.n) return fail(c, LAFIS_ERR_ARG, "template index out of range");
    LAFIS_CUDA(c, cudaSetDevice(c->device));
    const int nm = c->h_minu_n[t];
    const int np = (int)(c->h_minu_off[t + 1] - c->h_minu_off[t]);
    const uint32_t mo = c->h_minu_off[t], to = c->h_tex_off[t];
    const int nt = (int)(c->h_tex_off[t + 1] - to);
    if (n_minu) *n_minu = nm;
    if (n_tex) *n_tex = nt;
    LAFIS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (nm > 0 && (mx || my)) {
        std::vector<short2> xy(nm);
        LAFIS_CUDA(c, cudaMemcpy(xy.data(), c->gal.minu_xy + mo, sizeof(short2) * nm, cudaMemcpyDeviceToHost));
        for (int i = 0; i < nm; ++i) {
            if (mx) mx[i] = xy[i].x;
            if (my) my[i] = xy[i].y;
        }
    }
    if (nm > 0 && mori) LAFIS_CUDA(c, cudaMemcpy(mori, c->gal.minu_ori + mo, 4ull * nm, cudaMemcpyDeviceToHost));
    if (nm > 0 && mdes) {
        std::vector<float> T((size_t)kDesLen * np);
        LAFIS_CUDA(c, cudaMemcpy(T.data(), c->gal.minu_desT + (size_t)kDesLen * mo, 4ull * kDesLen * np,
                                 cudaMemcpyDeviceToHost));
        for (int i = 0; i < nm; ++i)
            for (int k = 0; k < kDesLen; ++k) mdes[(size_t)i * kDesLen + k] = T[(size_t)k * np + i];
    }
    if (nt > 0 && (tx || ty)) {
        std::vector<short2> xy(nt);
        LAFIS_CUDA(c, cudaMemcpy(xy.data(), c->gal.tex_xy + to, sizeof(short2) * nt, cudaMemcpyDeviceToHost));
        for (int i = 0; i < nt; ++i) {
            if (tx) tx[i] = xy[i].x;
            if (ty) ty[i] = xy[i].y;
        }
    }
    if (nt > 0 && tori) LAFIS_CUDA(c, cudaMemcpy(tori, c->gal.tex_ori + to, 4ull * nt, cudaMemcpyDeviceToHost));
    if (nt > 0 && tcodes) LAFIS_CUDA(c, cudaMemcpy(tcodes, c->gal.tex_codes + to, 16ull * nt, cudaMemcpyDeviceToHost));
    return LAFIS_OK;
}

// ---------------------------------------------------------------------------------------------------
// latents
// ---------------------------------------------------------------------------------------------------
static int finish_latents(lafis_ctx* c, lafis_latents* L) {
    // lay the staging vectors out in one pinned arena
    const int n = L->n;
    auto place = [](size_t& cur, size_t bytes) {
        size_t at = cur;
        cur = (cur + bytes + 255) & ~(size_t)255;
        return at;
    };
    size_t cur = 0;
    L->o_slot_n = place(cur, sizeof(int) * 3 * n);
    L->o_slot_off = place(cur, sizeof(uint32_t) * 3 * n);
    L->o_minu_xy = place(cur, sizeof(short2) * L->minu_xy.size());
    L->o_minu_ori = place(cur, sizeof(float) * L->minu_ori.size());
    L->o_minu_desT = place(cur, sizeof(float) * L->minu_desT.size());
    L->o_tex_n = place(cur, sizeof(int) * n);
    L->o_tex_xy = place(cur, sizeof(short2) * L->tex_xy.size());
    L->o_tex_ori = place(cur, sizeof(float) * L->tex_ori.size());
    L->o_tex_des = place(cur, sizeof(float) * L->tex_des.size());
    L->o_weighted = place(cur, sizeof(int) * n);
    L->o_status = place(cur, sizeof(int) * n);
    L->arena_bytes = cur;
    if (cudaSetDevice(c->device) != cudaSuccess || cudaMallocHost(&L->pinned, std::max<size_t>(cur, 256)) != cudaSuccess)
        return fail(c, LAFIS_ERR_CUDA, "cudaMallocHost(%zu) failed", cur);
    auto put = [&](size_t off, const void* src, size_t bytes) {
        if (bytes) std::memcpy(L->pinned + off, src, bytes);
    };
    put(L->o_slot_n, L->slot_n.data(), sizeof(int) * L->slot_n.size());
    put(L->o_slot_off, L->slot_off.data(), sizeof(uint32_t) * L->slot_off.size());
    put(L->o_minu_xy, L->minu_xy.data(), sizeof(short2) * L->minu_xy.size());
    put(L->o_minu_ori, L->minu_ori.data(), sizeof(float) * L->minu_ori.size());
    put(L->o_minu_desT, L->minu_desT.data(), sizeof(float) * L->minu_desT.size());
    put(L->o_tex_n, L->tex_n.data(), sizeof(int) * L->tex_n.size());
    put(L->o_tex_xy, L->tex_xy.data(), sizeof(short2) * L->tex_xy.size());
    put(L->o_tex_ori, L->tex_ori.data(), sizeof(float) * L->tex_ori.size());
    put(L->o_tex_des, L->tex_des.data(), sizeof(float) * L->tex_des.size());
    put(L->o_weighted, L->tex_weighted.data(), sizeof(int) * L->tex_weighted.size());
    put(L->o_status, L->status.data(), sizeof(int) * L->status.size());
    // the vectors are no longer needed once the arena is filled
    std::vector<short2>().swap(L->minu_xy);
    std::vector<float>().swap(L->minu_ori);
    std::vector<float>().swap(L->minu_desT);
    std::vector<short2>().swap(L->tex_xy);
    std::vector<float>().swap(L->tex_ori);
    std::vector<float>().swap(L->tex_des);
    return LAFIS_OK;
}

int lafis_latents_from_packed(lafis_ctx* c, const lafis_packed_latents* p, lafis_latents** out) {
    if (!c || !p || !out || p->n_latents < 0) return fail(c, LAFIS_ERR_ARG, "bad latent argument");
    const int n = p->n_latents;
    lafis_latents* L = new lafis_latents();
    L->n = n;
    L->owner = c;
    L->device = c->device;
    L->n_minu_templates.assign(p->n_minu_templates, p->n_minu_templates + n);
    L->status.resize(n);
    L->tex_weighted.resize(n);
    L->slot_n.assign(3 * n, 0);
    L->slot_off.assign(3 * n, 0);
    L->tex_n.assign(n, 0);
    int max_t = 0;
    for (int q = 0; q < n; ++q) {
        const int nm = p->n_minu_templates[q], nt = p->n_tex_templates[q];
        // One2One_matching_selected_templates, matcher.cpp:383-386, and the score[28] read of :188
        int st = LAFIS_OK;
        if (nm <= kSelected[0] && nt <= 0) st = LAFIS_LATENT_EMPTY;
        else if (nm + nt < 29) st = LAFIS_ERR_LATENT_LAYOUT;
        L->status[q] = st;
        // The texture score lands in score[n_minu_templates] (matcher.cpp:414).  The fusion (:188, :293) reads
        // score[0] + score[1] + score[2] + score[28] * 0.3: with 28 minutiae templates the texture score is the
        // weighted term (mode 1); with k = 0, 1 or 2 minutiae templates (and enough texture templates for score[28] to
        // exist) it sits in the unweighted slot score[k], the minutiae scores being all 0 then (mode 2 + k).  With
        // any other template count it is never read, so it is not computed.
        L->tex_weighted[q] = (nt >= 1 && st == LAFIS_OK) ? (nm == 28 ? 1 : (nm >= 0 && nm <= 2) ? 2 + nm : 0) : 0;
        int ntp = nt > 0 ? (int)(p->tex_off[q + 1] - p->tex_off[q]) : 0;
        if (ntp > kMaxTexture) ntp = kMaxTexture;  // matcher.cpp:544-545
        if (!L->tex_weighted[q]) ntp = 0;
        L->tex_n[q] = ntp;
        max_t = std::max(max_t, ntp);
    }
    L->lt_stride = std::max(kRowTile, round_up(max_t, kRowTile));
    uint32_t off = 0;
    for (int s = 0; s < 3 * n; ++s) {
        const int q = s / 3, slot = s % 3;
        int cnt = (int)(p->minu_off[s + 1] - p->minu_off[s]);
        if (p->n_minu_templates[q] <= kSelected[slot]) cnt = 0;  // matcher.cpp:403-404
        if (cnt > kMaxMinutiae) {
            delete L;
            return fail(c, LAFIS_ERR_UNSUPPORTED_SIZE, "latent %d: %d minutiae", q, cnt);
        }
        L->slot_n[s] = cnt;
        L->slot_off[s] = off;
        L->max_slot_n = std::max(L->max_slot_n, cnt);
        off += (uint32_t)((cnt + 3) & ~3);
    }
    L->tot_minu_padded = off;
    L->minu_xy.assign(std::max<uint32_t>(off, 4), make_short2(0, 0));
    L->minu_ori.assign(std::max<uint32_t>(off, 4), 0.0f);
    L->minu_desT.assign((size_t)kDesLen * std::max<uint32_t>(off, 4), 0.0f);
    for (int s = 0; s < 3 * n; ++s) {
        const int cnt = L->slot_n[s], np = (cnt + 3) & ~3;
        const uint32_t so = p->minu_off[s], dof = L->slot_off[s];
        for (int i = 0; i < cnt; ++i) {
            L->minu_xy[dof + i] = make_short2(p->minu_x[so + i], p->minu_y[so + i]);
            L->minu_ori[dof + i] = p->minu_ori[so + i];
            const float* d = p->minu_des + (size_t)(so + i) * kDesLen;
            float* T = L->minu_desT.data() + (size_t)kDesLen * dof;
            for (int k = 0; k < kDesLen; ++k) T[(size_t)k * np + i] = d[k];
        }
    }
    const size_t lt = (size_t)L->lt_stride;
    L->tex_xy.assign(lt * std::max(n, 1), make_short2(0, 0));
    L->tex_ori.assign(lt * std::max(n, 1), 0.0f);
    L->tex_des.assign(lt * std::max(n, 1) * kDesLen, 0.0f);
    for (int q = 0; q < n; ++q) {
        const int cnt = L->tex_n[q];
        const uint32_t so = p->tex_off[q];
        for (int i = 0; i < cnt; ++i) {
            L->tex_xy[q * lt + i] = make_short2(p->tex_x[so + i], p->tex_y[so + i]);
            L->tex_ori[q * lt + i] = p->tex_ori[so + i];
        }
        if (cnt)
            std::memcpy(L->tex_des.data() + (size_t)q * lt * kDesLen, p->tex_des + (size_t)so * kDesLen,
                        sizeof(float) * (size_t)cnt * kDesLen);
    }
    int rc = finish_latents(c, L);
    if (rc != LAFIS_OK) {
        lafis_latents_free(L);
        return rc;
    }
    *out = L;
    return LAFIS_OK;
}

int lafis_latents_load_files(lafis_ctx* c, const char* const* paths, int n, lafis_latents** out) {
    if (!c || !out || (!paths && n > 0) || n < 0) return fail(c, LAFIS_ERR_ARG, "bad argument");
    if (n == 0) return fail(c, LAFIS_ERR_NO_TEMPLATES, "no latent templates");
    std::vector<int32_t> nm(n), nt(n);
    std::vector<uint32_t> minu_off(3 * n + 1, 0), tex_off(n + 1, 0);
    std::vector<int16_t> mx, my, tx, ty;
    std::vector<float> mori, mdes, tori, tdes;
    for (int q = 0; q < n; ++q) {
        LatentTemplate T;
        read_latent_dat(paths[q], T);  // the drivers ignore the loader's return code (matcher.cpp:150, :259)
        nm[q] = T.n_minu_templates;
        nt[q] = T.n_tex_templates;
        for (int s = 0; s < 3; ++s) {
            const PointSet& P = T.minu[s];
            mx.insert(mx.end(), P.x.begin(), P.x.end());
            my.insert(my.end(), P.y.begin(), P.y.end());
            mori.insert(mori.end(), P.ori.begin(), P.ori.end());
            mdes.insert(mdes.end(), P.des.begin(), P.des.end());
            minu_off[3 * q + s + 1] = (uint32_t)mx.size();
        }
        tx.insert(tx.end(), T.tex.x.begin(), T.tex.x.end());
        ty.insert(ty.end(), T.tex.y.begin(), T.tex.y.end());
        tori.insert(tori.end(), T.tex.ori.begin(), T.tex.ori.end());
        tdes.insert(tdes.end(), T.tex.des.begin(), T.tex.des.end());
        tex_off[q + 1] = (uint32_t)tx.size();
    }
    lafis_packed_latents p{};
    p.n_latents = n;
    p.n_minu_templates = nm.data();
    p.n_tex_templates = nt.data();
    p.minu_off = minu_off.data();
    p.minu_x = mx.data();
    p.minu_y = my.data();
    p.minu_ori = mori.data();
    p.minu_des = mdes.data();
    p.tex_off = tex_off.data();
    p.tex_x = tx.data();
    p.tex_y = ty.data();
    p.tex_ori = tori.data();
    p.tex_des = tdes.data();
    return lafis_latents_from_packed(c, &p, out);
}

int lafis_latents_count(const lafis_latents* l) { return l ? l->n : 0; }
uint64_t lafis_latents_bytes(const lafis_latents* l) { return l ? (uint64_t)l->arena_bytes : 0; }
int lafis_latents_status(const lafis_latents* l, int q) {
    return (l && q >= 0 && q < l->n) ? l->status[q] : LAFIS_ERR_ARG;
}
int lafis_latents_minu_templates(const lafis_latents* l, int q) {
    return (l && q >= 0 && q < l->n) ? l->n_minu_templates[q] : 0;
}

void lafis_latents_free(lafis_latents* l) {
    if (!l) return;
    // the batch may outlive the context it was created on: only its own device number is used here
    // (cudaFree waits for outstanding work on that memory)
    cudaSetDevice(l->device);
    if (l->d_arena) cudaFree(l->d_arena);
    if (l->pinned) cudaFreeHost(l->pinned);
    delete l;
}

int lafis_latents_make_resident(lafis_ctx* c, lafis_latents* l) {
    if (!c || !l) return fail(c, LAFIS_ERR_ARG, "bad argument");
    LAFIS_CUDA(c, cudaSetDevice(c->device));
    if (!l->d_arena) LAFIS_CUDA(c, cudaMalloc(&l->d_arena, std::max<size_t>(l->arena_bytes, 256)));
    LAFIS_CUDA(c, cudaMemcpyAsync(l->d_arena, l->pinned, l->arena_bytes, cudaMemcpyHostToDevice, c->stream));
    LAFIS_CUDA(c, cudaStreamSynchronize(c->stream));
    l->resident = true;
    l->owner = c;
    l->device = c->device;
    return LAFIS_OK;
}

// ---------------------------------------------------------------------------------------------------
// the hot path
// ---------------------------------------------------------------------------------------------------
// Oversized minutiae pairs: the work list `jobs` (sizes[k] = nL * padded nR elements of job k) in sub-batches whose
// matrices fit a bounded scratch area; results land in the fast path's corr_* arrays before the graph stage reads them.
static int run_big_jobs(lafis_ctx* c, cudaStream_t st, MinuBigParams B, const std::vector<int>& jobs,
                        const std::vector<unsigned long long>& sizes) {
    constexpr unsigned long long kBatchElems = 48ull << 20;  // 192 MB per scratch array
    const size_t n = jobs.size();
    c->stats.minu_big_jobs += n;
    size_t at = 0;
    while (at < n) {
        std::vector<unsigned long long> off;
        unsigned long long tot = 0;
        size_t end = at;
        while (end < n && (end == at || tot + sizes[end] <= kBatchElems)) {
            off.push_back(tot);
            tot += (sizes[end] + 3ull) & ~3ull;  // keep every matrix 16-byte aligned
            ++end;
        }
        const int nb = (int)(end - at);
        LAFIS_CUDA(c, c->big_jobs.reserve(nb));
        LAFIS_CUDA(c, c->big_slow.reserve(nb));
        LAFIS_CUDA(c, c->big_soff.reserve(nb));
        LAFIS_CUDA(c, c->big_S.reserve(tot));
        LAFIS_CUDA(c, c->big_keys.reserve(tot));
        LAFIS_CUDA(c, c->big_order.reserve(tot));
        LAFIS_CUDA(c, cudaMemcpyAsync(c->big_jobs.p, jobs.data() + at, sizeof(int) * nb, cudaMemcpyHostToDevice, st));
        LAFIS_CUDA(c, cudaMemcpyAsync(c->big_soff.p, off.data(), sizeof(unsigned long long) * nb, cudaMemcpyHostToDevice, st));
        LAFIS_CUDA(c, cudaStreamSynchronize(st));  // `off` is a pageable temporary
        LAFIS_CUDA(c, cudaMemsetAsync(c->d_slow_count, 0, sizeof(int), st));
        B.jobs = c->big_jobs.p;
        B.s_off = c->big_soff.p;
        B.n_jobs = nb;
        B.S = c->big_S.p;
        B.keys = c->big_keys.p;
        B.order = c->big_order.p;
        B.slow_count = c->d_slow_count;
        B.slow_list = c->big_slow.p;
        B.replay_count = c->d_slow;
        const size_t smem = minu_big_select_smem_bytes(B.max_nL, B.max_np);
        minu_big_sim_kernel<<<nb, 256, 0, st>>>(B);
        minu_big_select_kernel<<<nb, kSelThreads, smem, st>>>(B);
        minu_big_slow_kernel<<<std::min(nb, 2 * c->sm_count), kSelThreads, smem, st>>>(B);
        c->stats.kernel_launches += 3;
        LAFIS_CUDA(c, cudaGetLastError());
        at = end;
    }
    return LAFIS_OK;
}

}  // extern "C"

// An empty shard (more ranks than gallery files): empty rank lists, no scores.
__global__ void fill_empty_hits_kernel(HitDev* hits, size_t n) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) {
        hits[e].score = -CUDART_INF_F;
        hits[e].index = 0xffffffffu;
    }
}

int lafis::run_match(lafis_ctx* c, lafis_latents* L, int topk) {
    const int Q = L->n, G = c->gal.n;
    if (G <= 0 && c->gallery_set && Q > 0) {
        if (topk < 0 || topk > kTopkChunk / 2) return fail(c, LAFIS_ERR_ARG, "topk must be in [0, %d]", kTopkChunk / 2);
        LAFIS_CUDA(c, cudaSetDevice(c->device));
        LAFIS_CUDA(c, cudaEventRecord(c->ev0, c->stream));
        c->stage_chunks = 0;
        while (c->stage_ev.size() < 2) {
            cudaEvent_t e;
            LAFIS_CUDA(c, cudaEventCreate(&e));
            c->stage_ev.push_back(e);
        }
        cudaEventRecord(c->stage_ev[0], c->stream);
        if (topk > 0) {
            const size_t nh = (size_t)Q * topk;
            LAFIS_CUDA(c, c->hits.reserve(nh));
            fill_empty_hits_kernel<<<(unsigned)((nh + 255) / 256), 256, 0, c->stream>>>(c->hits.p, nh);
            c->stats.kernel_launches += 1;
        }
        cudaEventRecord(c->stage_ev[1], c->stream);
        LAFIS_CUDA(c, cudaEventRecord(c->ev1, c->stream));
        return LAFIS_OK;
    }
    if (G <= 0) return fail(c, LAFIS_ERR_NO_GALLERY, "no gallery resident");
    if (Q <= 0) return fail(c, LAFIS_ERR_ARG, "empty latent batch");
    if (topk < 0 || topk > kTopkChunk / 2) return fail(c, LAFIS_ERR_ARG, "topk must be in [0, %d]", kTopkChunk / 2);
    LAFIS_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;

    // ---- latent batch in HBM ----
    unsigned char* A = nullptr;
    if (L->resident && L->owner == c) {
        A = L->d_arena;
    } else {
        LAFIS_CUDA(c, c->lat_arena.reserve(L->arena_bytes));
        A = c->lat_arena.p;
    }
    LAFIS_CUDA(c, cudaEventRecord(c->ev0, st));
    if (A != L->d_arena) LAFIS_CUDA(c, cudaMemcpyAsync(A, L->pinned, L->arena_bytes, cudaMemcpyHostToDevice, st));
    DeviceLatents D;
    D.n = Q;
    D.lt_stride = L->lt_stride;
    D.slot_n = reinterpret_cast<int*>(A + L->o_slot_n);
    D.slot_off = reinterpret_cast<uint32_t*>(A + L->o_slot_off);
    D.minu_xy = reinterpret_cast<short2*>(A + L->o_minu_xy);
    D.minu_ori = reinterpret_cast<float*>(A + L->o_minu_ori);
    D.minu_desT = reinterpret_cast<float*>(A + L->o_minu_desT);
    D.tex_n = reinterpret_cast<int*>(A + L->o_tex_n);
    D.tex_xy = reinterpret_cast<short2*>(A + L->o_tex_xy);
    D.tex_ori = reinterpret_cast<float*>(A + L->o_tex_ori);
    D.tex_des = reinterpret_cast<float*>(A + L->o_tex_des);
    D.tex_weighted = reinterpret_cast<int*>(A + L->o_weighted);
    D.status = reinterpret_cast<int*>(A + L->o_status);

    // ---- geometry ----
    const MinuPlan plan = plan_minu(L->max_slot_n, c->max_nR, c->h_minu_n.data(), c->h_minu_n.size());
    const int maxL = plan.maxL, maxNp = plan.maxNp;
    const int a_slot_stride = plan.a_slot_stride, b_buf_stride = plan.b_buf_stride, b_double = plan.b_double;
    const size_t sim_smem = plan.sim_smem, sel_smem = plan.sel_smem, slow_smem = plan.slow_smem;
    const bool slow_dense = plan.slow_dense;
    const size_t job_stride = plan.job_stride;
    // pairs beyond the fast kernels' tiles (minu_big.cuh)
    const bool has_big = L->max_slot_n > plan.l_cap || c->max_nR > plan.r_cap;
    const size_t lt = (size_t)L->lt_stride;
    const size_t per_tpl = (size_t)Q * (lt * 6 + 3 * (kTopCorrMinu * 8 + 4) + 3 * job_stride * 4 + 32);
    size_t chunk_sz = std::max<size_t>(1, c->work_budget / std::max<size_t>(per_tpl, 1));
    chunk_sz = std::min<size_t>(chunk_sz, (size_t)G);
    // keep every grid dimension and flat index comfortably inside 31 bits
    chunk_sz = std::min<size_t>(chunk_sz, (size_t)0x7fffffff / ((size_t)Q * 3 * kTopCorrMinu));
    chunk_sz = std::max<size_t>(chunk_sz, 1);
    const int n_chunk_max = (int)chunk_sz;

    LAFIS_CUDA(c, c->tex_lut.reserve((size_t)Q * lt * 4096));
    LAFIS_CUDA(c, c->tex_scale.reserve((size_t)Q * lt));
    LAFIS_CUDA(c, c->tex_lut8.reserve((size_t)Q * ((lt + kRowTile - 1) / kRowTile) * kLutBytes));
    LAFIS_CUDA(c, c->rowmax_val.reserve((size_t)Q * n_chunk_max * lt));
    LAFIS_CUDA(c, c->rowmax_j.reserve((size_t)Q * n_chunk_max * lt));
    LAFIS_CUDA(c, c->corr_v.reserve((size_t)Q * n_chunk_max * 3 * kTopCorrMinu));
    LAFIS_CUDA(c, c->corr_ij.reserve((size_t)Q * n_chunk_max * 3 * kTopCorrMinu));
    LAFIS_CUDA(c, c->corr_n.reserve((size_t)Q * n_chunk_max * 3));
    LAFIS_CUDA(c, c->sim.reserve((size_t)Q * n_chunk_max * 3 * job_stride));
    LAFIS_CUDA(c, c->slow_jobs.reserve((size_t)Q * n_chunk_max * 3));
    LAFIS_CUDA(c, c->ov_minu.reserve((size_t)Q * n_chunk_max * 3));
    LAFIS_CUDA(c, c->ov_minu2.reserve((size_t)Q * n_chunk_max * 3));
    LAFIS_CUDA(c, c->ov_tex.reserve((size_t)Q * n_chunk_max));
    LAFIS_CUDA(c, c->comp.reserve((size_t)Q * G * 4));
    LAFIS_CUDA(c, c->final_scores.reserve((size_t)Q * G));

    // stage time stamps: a (start, end) pair for each of the 5 stages of a chunk, recorded on the stage's own
    // stream, and 2 for the tail
    const int n_chunks = (G + n_chunk_max - 1) / n_chunk_max;
    while ((int)c->stage_ev.size() < 16 * n_chunks + 2) {
        cudaEvent_t e;
        LAFIS_CUDA(c, cudaEventCreate(&e));
        c->stage_ev.push_back(e);
    }
    c->stage_chunks = n_chunks;
    int chunk_id = 0;
    // The texture chain could run on a second stream next to the minutiae chain; it does not pay: the persistent kernels
    // take whole SMs, and the smaller kernels of the two chains only slow each other down (1 latent x 100,000 prints in one
    // chunk: 68.7 against 68.1 ms; 27 latents in seven chunks: 1,901 against 1,840 ms; 256 latents x 20,000: 3,630 against
    // 3,523 ms).  Kept as an experiment switch.
    static const bool force_two = getenv("LAFIS_FORCE_TWO_STREAMS") != nullptr;
    const bool two = c->two_streams && force_two;
    cudaStream_t sb = two ? c->stream_b : st;
    // Rare-path kernels on their own (high-priority) stream: the selection's introsort replays run next to the texture
    // graph kernel, the dense texture graphs next to the minutiae graph kernel, instead of a few long jobs holding up the
    // 300,000-CTA kernel behind them (0.8 ms of 68 per 100,000 pairs).  Not with oversized pairs (their host-built work
    // lists follow the replay kernel on the main stream) and not in the serialised mode (lafis_set_streams(ctx, 1)).
    const bool tails = c->two_streams && !has_big;
    cudaStream_t sc = tails ? c->stream_c : st;
    bool tail_pending = false;  // an earlier chunk's dense texture graphs may still read the overflow list
    auto begin = [&](int stage, cudaStream_t s_) { cudaEventRecord(c->stage_ev[16 * chunk_id + 2 * stage], s_); };
    auto end = [&](int stage, cudaStream_t s_) { cudaEventRecord(c->stage_ev[16 * chunk_id + 2 * stage + 1], s_); };
    if (two) {  // the texture chain starts once the latent batch is in HBM
        LAFIS_CUDA(c, cudaEventRecord(c->ev_fork, st));
        LAFIS_CUDA(c, cudaStreamWaitEvent(sb, c->ev_fork, 0));
    }

    {   // ---- K1: fp32 distance tables of the batch, once per match ----
        TexLutParams P;
        P.lat_des = D.tex_des;
        P.lat_nt = D.tex_n;
        P.lt_stride = L->lt_stride;
        P.Q = Q;
        P.codebook = c->d_codebook;
        P.lut = c->tex_lut.p;
        P.row_scale = c->tex_scale.p;
        tex_lut_kernel<<<dim3(L->lt_stride, Q), 256, 0, sb>>>(P);
        // ... and their 8-bit copies, one 128 KB tile per 32 rows in the layout tex_rowmax_kernel gathers from
        TexRowmaxParams T;
        T.lut = c->tex_lut.p;
        T.row_scale = c->tex_scale.p;
        T.lut8 = c->tex_lut8.p;
        T.lat_nt = D.tex_n;
        T.lt_stride = L->lt_stride;
        T.Q = Q;
        tex_lut8_kernel<<<dim3((L->lt_stride + kRowTile - 1) / kRowTile, Q), 512, 0, sb>>>(T);
        c->stats.kernel_launches += 2;
    }
    for (int g0 = 0; g0 < G; g0 += n_chunk_max) {
        const int n_chunk = std::min(n_chunk_max, G - g0);
        begin(1, st);
        // ---- K5, then K6 + K7 ----
        {
            MinuSimParams P;
            P.slot_n = D.slot_n;
            P.slot_off = D.slot_off;
            P.lat_desT = D.minu_desT;
            P.lat_status = D.status;
            P.Q = Q;
            P.minu_off = c->gal.minu_off;
            P.minu_n = c->gal.minu_n;
            P.minu_desT = c->gal.minu_desT;
            P.g0 = g0;
            P.n_chunk = n_chunk;
            P.a_slot_stride = a_slot_stride;
            P.b_buf_stride = b_buf_stride;
            P.b_double = b_double;
            P.S = c->sim.p;
            P.job_stride = job_stride;
            P.l_cap = plan.l_cap;
            P.r_cap = plan.r_cap;
            {
                int g = c->sm_count, r = Q;  // gcd(Q, SMs)
                while (r) {
                    const int t = g % r;
                    g = r;
                    r = t;
                }
                P.parts = std::max(1, std::min(c->sm_count / g, n_chunk));
            }
            if (Q > 1 && n_chunk < 16 * c->sm_count)  // few templates per CTA and chunk: balanced (latent, part) jobs
                minu_sim_jobs_kernel<<<std::min(Q * P.parts, c->sm_count), kSimThreads, sim_smem, st>>>(P);
            else
                minu_sim_kernel<<<std::min(n_chunk, c->sm_count), kSimThreads, sim_smem, st>>>(P);
            end(1, st);
            begin(2, st);
            MinuSelectParams R;
            R.slot_n = D.slot_n;
            R.lat_status = D.status;
            R.Q = Q;
            R.minu_n = c->gal.minu_n;
            R.g0 = g0;
            R.n_chunk = n_chunk;
            R.S = c->sim.p;
            R.job_stride = job_stride;
            R.max_nL = maxL;
            R.max_np = maxNp;
            R.l_cap = plan.l_cap;
            R.r_cap = plan.r_cap;
            R.slow_dense = slow_dense ? 1 : 0;
            R.corr_v = c->corr_v.p;
            R.corr_ij = c->corr_ij.p;
            R.corr_n = c->corr_n.p;
            R.slow_count = c->d_slow_count;
            R.slow_jobs = c->slow_jobs.p;
            LAFIS_CUDA(c, cudaMemsetAsync(c->d_slow_count, 0, sizeof(int), st));
            const unsigned jobs = (unsigned)((size_t)Q * n_chunk * 3);
            minu_select_kernel<<<job_grid(3u * (unsigned)n_chunk, Q), kSelThreads, sel_smem, st>>>(R);
            end(2, st);
            if (tails) {
                LAFIS_CUDA(c, cudaEventRecord(c->ev_sel, st));
                LAFIS_CUDA(c, cudaStreamWaitEvent(sc, c->ev_sel, 0));
            }
            begin(6, sc);
            minu_select_slow_kernel<<<std::min<unsigned>(jobs, 2u * c->sm_count), kSelThreads, slow_smem, sc>>>(R, c->d_slow);
            if (tails) LAFIS_CUDA(c, cudaEventRecord(c->ev_slow, sc));
            if (has_big) {
                // work list of this chunk: every (latent, slot) against the oversized templates, and oversized
                // latent slots against every template
                std::vector<int> big_tl;
                for (int tl = 0; tl < n_chunk; ++tl)
                    if (c->h_minu_n[g0 + tl] > plan.r_cap) big_tl.push_back(tl);
                std::vector<int> bj;
                std::vector<unsigned long long> bsz;
                for (int q = 0; q < Q; ++q) {
                    if (L->status[q] != LAFIS_OK) continue;
                    for (int slot = 0; slot < 3; ++slot) {
                        const int nLs = L->slot_n[q * 3 + slot];
                        if (nLs <= 0) continue;
                        auto push = [&](int tl) {
                            const int nR = c->h_minu_n[g0 + tl];
                            if (nR <= 0) return;
                            bj.push_back((int)(((size_t)q * n_chunk + tl) * 3 + slot));
                            bsz.push_back((unsigned long long)nLs * ((nR + 3) & ~3));
                        };
                        if (nLs > plan.l_cap) {
                            for (int tl = 0; tl < n_chunk; ++tl) push(tl);
                        } else {
                            for (int tl : big_tl) push(tl);
                        }
                    }
                }
                MinuBigParams B;
                B.slot_n = D.slot_n;
                B.slot_off = D.slot_off;
                B.lat_desT = D.minu_desT;
                B.lat_status = D.status;
                B.minu_off = c->gal.minu_off;
                B.minu_n = c->gal.minu_n;
                B.minu_desT = c->gal.minu_desT;
                B.g0 = g0;
                B.n_chunk = n_chunk;
                B.max_nL = std::max(1, L->max_slot_n);
                B.max_np = std::max(4, (c->max_nR + 3) & ~3);
                B.corr_v = c->corr_v.p;
                B.corr_ij = c->corr_ij.p;
                B.corr_n = c->corr_n.p;
                const int rc_big = run_big_jobs(c, st, B, bj, bsz);
                if (rc_big != LAFIS_OK) return rc_big;
            }
        }
        end(6, sc);
        // ---- texture chain (stream sb): K2 + K3a, then K3b + K4 + K9 ----
        if (tail_pending)  // the previous chunk's dense texture graphs read the row maxima and the overflow list
            LAFIS_CUDA(c, cudaStreamWaitEvent(sb, c->ev_gtexd, 0));
        begin(0, sb);
        {
            TexRowmaxParams P;
            P.lut = c->tex_lut.p;
            P.row_scale = c->tex_scale.p;
            P.lut8 = c->tex_lut8.p;
            P.lat_nt = D.tex_n;
            P.lt_stride = L->lt_stride;
            P.Q = Q;
            P.tex_off = c->gal.tex_off;
            P.codes = c->gal.tex_codes;
            P.g0 = g0;
            P.n_chunk = n_chunk;
            const int n_rowtiles = L->lt_stride / kRowTile;
            // Jobs = (slice of the chunk, latent, row tile), drawn dynamically and slice-major by one persistent CTA per SM.
            // Slices of ~352 templates (~4.5 MB of code words): the few slices in flight at any time stay in L2, so the
            // gallery's code words come from HBM about once per launch, whatever the number of row tiles and latents.
            // With at least one slice per SM the count is rounded to a multiple of the SM count, so that the number of
            // (equal-sized) jobs is a multiple of the grid; chunks too small for ~4 jobs per SM are cut into slices of
            // >= 64 templates.
            int slices = std::max(1, n_chunk / 352);
            if (slices >= c->sm_count) slices = (slices + c->sm_count / 2) / c->sm_count * c->sm_count;
            if ((long long)slices * Q * n_rowtiles < 4ll * c->sm_count) {
                slices = (4 * c->sm_count + Q * n_rowtiles - 1) / (Q * n_rowtiles);
                slices = std::max(1, std::min(slices, (n_chunk + 63) / 64));
            }
            P.slices = slices;
            P.one = 1u;
            P.rowmax_val = c->rowmax_val.p;
            P.rowmax_j = c->rowmax_j.p;
            P.job_counter = c->d_job_counter;
            P.counters = c->d_slow + 4;
            LAFIS_CUDA(c, cudaMemsetAsync(c->d_job_counter, 0, sizeof(int), sb));
            const int grid = std::min(c->sm_count, Q * n_rowtiles * slices);
            tex_rowmax_kernel<<<grid, kRowmaxThreads, kRowmaxSmem, sb>>>(P);
        }
        end(0, sb);
        begin(4, sb);
        // ---- K3b + K4 + K9 (texture) ----
        {
            LAFIS_CUDA(c, cudaMemsetAsync(c->d_ov_count + 1, 0, sizeof(int), sb));
            GraphTexParams P;
            P.rowmax_val = c->rowmax_val.p;
            P.rowmax_j = c->rowmax_j.p;
            P.lt_stride = L->lt_stride;
            P.lat_nt = D.tex_n;
            P.lat_status = D.status;
            P.lat_xy = D.tex_xy;
            P.lat_ori = D.tex_ori;
            P.tex_off = c->gal.tex_off;
            P.gal_xy = c->gal.tex_xy;
            P.gal_ori = c->gal.tex_ori;
            P.table = c->d_table;
            P.g0 = g0;
            P.n_chunk = n_chunk;
            P.G = G;
            P.Q = Q;
            P.comp = c->comp.p;
            P.slow_path_count = c->d_slow + 1;
            P.dense_jobs_total = c->d_slow + 3;
            const unsigned grid = (unsigned)((size_t)Q * n_chunk);
            graph_tex_sparse_kernel<<<job_grid((unsigned)n_chunk, Q), SparseGeom<true>::NT, sizeof(SparseWork<true>), sb>>>(
                P, OverflowList{c->d_ov_count + 1, c->ov_tex.p});
            cudaStream_t sd = sb;
            if (tails) {  // the few dense jobs run next to the minutiae graph kernel
                LAFIS_CUDA(c, cudaEventRecord(c->ev_gtex, sb));
                LAFIS_CUDA(c, cudaStreamWaitEvent(sc, c->ev_gtex, 0));
                sd = sc;
            }
            graph_tex_dense_kernel<<<std::min<unsigned>(grid, (unsigned)c->sm_count), kGraphTexThreads, kGraphTexSmem, sd>>>(
                P, c->d_ov_count + 1, c->ov_tex.p);
            end(4, sd);
            if (tails) {
                LAFIS_CUDA(c, cudaEventRecord(c->ev_gtexd, sc));
                tail_pending = true;
            }
        }
        if (tails) LAFIS_CUDA(c, cudaStreamWaitEvent(st, c->ev_slow, 0));  // the replayed jobs' candidate lists
        begin(3, st);
        // ---- K8 + K9 (minutiae) ----
        {
            GraphMinuParams P;
            P.corr_v = c->corr_v.p;
            P.corr_ij = c->corr_ij.p;
            P.corr_n = c->corr_n.p;
            P.slot_off = D.slot_off;
            P.lat_xy = D.minu_xy;
            P.lat_ori = D.minu_ori;
            P.minu_off = c->gal.minu_off;
            P.gal_xy = c->gal.minu_xy;
            P.gal_ori = c->gal.minu_ori;
            P.g0 = g0;
            P.n_chunk = n_chunk;
            P.G = G;
            P.Q = Q;
            P.comp = c->comp.p;
            P.dense_jobs_total = c->d_slow + 2;
            const unsigned grid = (unsigned)((size_t)Q * n_chunk * 3);
            LAFIS_CUDA(c, cudaMemsetAsync(c->d_ov_count, 0, sizeof(int), st));
            LAFIS_CUDA(c, cudaMemsetAsync(c->d_ov_count + 2, 0, sizeof(int), st));
            graph_minu_sparse_kernel<<<job_grid(3u * (unsigned)n_chunk, Q), SparseGeom<false>::NT, sizeof(SparseWork<false>), st>>>(
                P, OverflowList{c->d_ov_count, c->ov_minu.p});
            end(3, st);
            begin(7, st);
            // overflowed jobs: second chance with a larger CSR, then the dense kernel for what is left
            graph_minu_mid_kernel<<<std::min<unsigned>(grid, 5u * c->sm_count), SparseGeom<false, 1>::NT, sizeof(SparseWork<false, 1>), st>>>(
                P, c->d_ov_count, c->ov_minu.p, OverflowList{c->d_ov_count + 2, c->ov_minu2.p}, c->d_slow + 8);
            graph_minu_dense_kernel<<<std::min<unsigned>(grid, 2u * c->sm_count), kGraphMinuThreads, kGraphMinuSmem, st>>>(
                P, c->d_ov_count + 2, c->ov_minu2.p);
        }
        end(7, st);
        c->stats.kernel_launches += 9;
        LAFIS_CUDA(c, cudaGetLastError());
        ++chunk_id;
    }

    if (two) {  // join: the fusion needs both chains
        LAFIS_CUDA(c, cudaEventRecord(c->ev_join, sb));
        LAFIS_CUDA(c, cudaStreamWaitEvent(st, c->ev_join, 0));
    }
    if (tail_pending) LAFIS_CUDA(c, cudaStreamWaitEvent(st, c->ev_gtexd, 0));
    // ---- K10 ----
    cudaEventRecord(c->stage_ev[16 * n_chunks], st);
    {
        FuseParams P;
        P.comp = c->comp.p;
        P.lat_status = D.status;
        P.tex_weighted = D.tex_weighted;
        P.gal_status = c->gal.status;
        P.Q = Q;
        P.G = G;
        P.final_scores = c->final_scores.p;
        const size_t tot = (size_t)Q * G;
        fuse_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(P);
        c->stats.kernel_launches += 1;
    }
    // ---- K11: rank lists ----
    if (topk > 0) {
        int chunks = (G + kTopkChunk - 1) / kTopkChunk;
        LAFIS_CUDA(c, c->keys_a.reserve((size_t)Q * chunks * topk));
        LAFIS_CUDA(c, c->hits.reserve((size_t)Q * topk));
        topk_scores_kernel<<<dim3(chunks, Q), kTopkThreads, 0, st>>>(c->final_scores.p, G, c->index_base, topk,
                                                                      c->keys_a.p);
        c->stats.kernel_launches += 1;
        unsigned long long* cur = c->keys_a.p;
        int n_in = chunks * topk;
        bool use_b = true;
        while (n_in > topk) {
            const int chunks2 = (n_in + kTopkChunk - 1) / kTopkChunk;
            DevBuf<unsigned long long>& dst = use_b ? c->keys_b : c->keys_a;
            LAFIS_CUDA(c, dst.reserve((size_t)Q * chunks2 * topk));
            topk_keys_kernel<<<dim3(chunks2, Q), kTopkThreads, 0, st>>>(cur, n_in, topk, dst.p);
            c->stats.kernel_launches += 1;
            cur = dst.p;
            n_in = chunks2 * topk;
            use_b = !use_b;
            if (chunks2 == 1) break;
        }
        const size_t nh = (size_t)Q * topk;
        keys_to_hits_kernel<<<(unsigned)((nh + 255) / 256), 256, 0, st>>>(cur, nh, c->hits.p);
        c->stats.kernel_launches += 1;
    }
    cudaEventRecord(c->stage_ev[16 * n_chunks + 1], st);
    LAFIS_CUDA(c, cudaGetLastError());
    LAFIS_CUDA(c, cudaEventRecord(c->ev1, st));
    c->stats.pairs_scored += (uint64_t)Q * G;
    return LAFIS_OK;
}

extern "C" {

// ---------------------------------------------------------------------------------------------------
// correspondences of one (latent, gallery template) pair - the reference's save_corr output
// (matcher.cpp:322-327 for the 24 best of a 1-vs-N search, written by :497-505)
// ---------------------------------------------------------------------------------------------------
static int run_correspondences(lafis_ctx* c, lafis_latents* L, int q, int gi, short4* h_xy, int* h_n) {
    const int Q = L->n, G = c->gal.n;
    if (G <= 0) return fail(c, LAFIS_ERR_NO_GALLERY, "no gallery resident");
    if (q < 0 || q >= Q || gi < 0 || gi >= G) return fail(c, LAFIS_ERR_ARG, "latent %d / gallery template %d out of range", q, gi);
    LAFIS_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    unsigned char* A = nullptr;
    if (L->resident && L->owner == c) {
        A = L->d_arena;
    } else {
        LAFIS_CUDA(c, c->lat_arena.reserve(L->arena_bytes));
        A = c->lat_arena.p;
        LAFIS_CUDA(c, cudaMemcpyAsync(A, L->pinned, L->arena_bytes, cudaMemcpyHostToDevice, st));
    }
    const int nR_gi = c->h_minu_n[gi];
    const MinuPlan plan = plan_minu(L->max_slot_n, nR_gi, nullptr, 0);
    const int maxL = plan.maxL, maxNp = plan.maxNp;
    const int a_slot_stride = plan.a_slot_stride, b_buf_stride = plan.b_buf_stride, b_double = plan.b_double;
    const size_t sim_smem = plan.sim_smem, sel_smem = plan.sel_smem, slow_smem = plan.slow_smem;
    const bool slow_dense = plan.slow_dense;
    const size_t job_stride = plan.job_stride;
    const unsigned jobs = (unsigned)Q * 3u;
    LAFIS_CUDA(c, c->corr_v.reserve((size_t)jobs * kTopCorrMinu));
    LAFIS_CUDA(c, c->corr_ij.reserve((size_t)jobs * kTopCorrMinu));
    LAFIS_CUDA(c, c->corr_n.reserve(jobs));
    LAFIS_CUDA(c, c->sim.reserve((size_t)jobs * job_stride));
    LAFIS_CUDA(c, c->slow_jobs.reserve(jobs));
    LAFIS_CUDA(c, c->ov_minu.reserve(jobs));
    LAFIS_CUDA(c, c->comp.reserve((size_t)Q * G * 4));
    LAFIS_CUDA(c, c->corr_xy.reserve((size_t)jobs * kTopCorrMinu));
    LAFIS_CUDA(c, c->corr_xy_n.reserve(jobs));

    MinuSimParams P;
    P.slot_n = reinterpret_cast<int*>(A + L->o_slot_n);
    P.slot_off = reinterpret_cast<uint32_t*>(A + L->o_slot_off);
    P.lat_desT = reinterpret_cast<float*>(A + L->o_minu_desT);
    P.lat_status = reinterpret_cast<int*>(A + L->o_status);
    P.Q = Q;
    P.minu_off = c->gal.minu_off;
    P.minu_n = c->gal.minu_n;
    P.minu_desT = c->gal.minu_desT;
    P.g0 = gi;
    P.n_chunk = 1;
    P.a_slot_stride = a_slot_stride;
    P.b_buf_stride = b_buf_stride;
    P.b_double = b_double;
    P.S = c->sim.p;
    P.job_stride = job_stride;
    P.parts = 1;
    P.l_cap = plan.l_cap;
    P.r_cap = plan.r_cap;
    minu_sim_kernel<<<1, kSimThreads, sim_smem, st>>>(P);
    MinuSelectParams R;
    R.slot_n = P.slot_n;
    R.lat_status = P.lat_status;
    R.Q = Q;
    R.minu_n = c->gal.minu_n;
    R.g0 = gi;
    R.n_chunk = 1;
    R.S = c->sim.p;
    R.job_stride = job_stride;
    R.max_nL = maxL;
    R.max_np = maxNp;
    R.l_cap = plan.l_cap;
    R.r_cap = plan.r_cap;
    R.slow_dense = slow_dense ? 1 : 0;
    R.corr_v = c->corr_v.p;
    R.corr_ij = c->corr_ij.p;
    R.corr_n = c->corr_n.p;
    R.slow_count = c->d_slow_count;
    R.slow_jobs = c->slow_jobs.p;
    LAFIS_CUDA(c, cudaMemsetAsync(c->d_slow_count, 0, sizeof(int), st));
    minu_select_kernel<<<job_grid(3u, Q), kSelThreads, sel_smem, st>>>(R);
    minu_select_slow_kernel<<<std::min<unsigned>(jobs, (unsigned)c->sm_count), kSelThreads, slow_smem, st>>>(R, c->d_slow);
    if (nR_gi > 0 && L->status[q] == LAFIS_OK && (nR_gi > plan.r_cap || L->max_slot_n > plan.l_cap)) {
        std::vector<int> bj;
        std::vector<unsigned long long> bsz;
        for (int slot = 0; slot < 3; ++slot) {
            const int nLs = L->slot_n[q * 3 + slot];
            if (nLs <= 0 || (nLs <= plan.l_cap && nR_gi <= plan.r_cap)) continue;
            bj.push_back(q * 3 + slot);
            bsz.push_back((unsigned long long)nLs * ((nR_gi + 3) & ~3));
        }
        MinuBigParams B;
        B.slot_n = P.slot_n;
        B.slot_off = P.slot_off;
        B.lat_desT = P.lat_desT;
        B.lat_status = P.lat_status;
        B.minu_off = c->gal.minu_off;
        B.minu_n = c->gal.minu_n;
        B.minu_desT = c->gal.minu_desT;
        B.g0 = gi;
        B.n_chunk = 1;
        B.max_nL = std::max(1, L->max_slot_n);
        B.max_np = std::max(4, (nR_gi + 3) & ~3);
        B.corr_v = c->corr_v.p;
        B.corr_ij = c->corr_ij.p;
        B.corr_n = c->corr_n.p;
        const int rc_big = run_big_jobs(c, st, B, bj, bsz);
        if (rc_big != LAFIS_OK) return rc_big;
    }
    // every job of latent q goes through the dense graph kernel, which can list its survivors
    const int h_jobs[3] = {q * 3 + 0, q * 3 + 1, q * 3 + 2};
    const int h_count = 3;
    LAFIS_CUDA(c, cudaMemcpyAsync(c->ov_minu.p, h_jobs, sizeof h_jobs, cudaMemcpyHostToDevice, st));
    LAFIS_CUDA(c, cudaMemcpyAsync(c->d_ov_count, &h_count, sizeof(int), cudaMemcpyHostToDevice, st));
    GraphMinuParams Gp;
    Gp.corr_v = c->corr_v.p;
    Gp.corr_ij = c->corr_ij.p;
    Gp.corr_n = c->corr_n.p;
    Gp.slot_off = P.slot_off;
    Gp.lat_xy = reinterpret_cast<short2*>(A + L->o_minu_xy);
    Gp.lat_ori = reinterpret_cast<float*>(A + L->o_minu_ori);
    Gp.minu_off = c->gal.minu_off;
    Gp.gal_xy = c->gal.minu_xy;
    Gp.gal_ori = c->gal.minu_ori;
    Gp.g0 = gi;
    Gp.n_chunk = 1;
    Gp.G = G;
    Gp.Q = Q;
    Gp.comp = c->comp.p;
    Gp.corr_out = c->corr_xy.p;
    Gp.corr_out_n = c->corr_xy_n.p;
    graph_minu_dense_kernel<<<3, kGraphMinuThreads, kGraphMinuSmem, st>>>(Gp, c->d_ov_count, c->ov_minu.p);
    c->stats.kernel_launches += 4;
    LAFIS_CUDA(c, cudaGetLastError());
    LAFIS_CUDA(c, cudaMemcpyAsync(h_n, c->corr_xy_n.p + q * 3, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
    LAFIS_CUDA(c, cudaMemcpyAsync(h_xy, c->corr_xy.p + (size_t)q * 3 * kTopCorrMinu, 3 * kTopCorrMinu * sizeof(short4),
                                  cudaMemcpyDeviceToHost, st));
    LAFIS_CUDA(c, cudaStreamSynchronize(st));
    return LAFIS_OK;
}

}  // extern "C"

// after the stream has been synchronised: device times of the last match
void lafis::collect_times(lafis_ctx* c) {
    cudaEventElapsedTime(&c->stats.last_match_ms, c->ev0, c->ev1);
    float ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < c->stage_chunks; ++k)
        for (int s = 0; s < 8; ++s) {
            if (s == 5) continue;
            float t = 0;
            cudaEventElapsedTime(&t, c->stage_ev[16 * k + 2 * s], c->stage_ev[16 * k + 2 * s + 1]);
            ms[s] += t;
        }
    float t = 0;
    cudaEventElapsedTime(&t, c->stage_ev[16 * c->stage_chunks], c->stage_ev[16 * c->stage_chunks + 1]);
    ms[5] = t;
    std::memcpy(c->stats.last_stage_ms, ms, sizeof ms);
    unsigned long long cnt[12];
    if (cudaMemcpy(cnt, c->d_slow, sizeof cnt, cudaMemcpyDeviceToHost) == cudaSuccess) {
        c->stats.graph_minu_mid_jobs = cnt[8];
        c->stats.minu_replays = cnt[0];
        c->stats.tex_replays = cnt[1];
        c->stats.graph_minu_dense_jobs = cnt[2];
        c->stats.graph_tex_dense_jobs = cnt[3];
        c->stats.tex_queued = cnt[4];
        c->stats.tex_exact = cnt[5];
        c->stats.tex_overflow = cnt[6];
        c->stats.tex_templates = cnt[7];
    }
}

extern "C" {

int lafis_match_device(lafis_ctx* c, lafis_latents* L, int topk, const void** d_hits, const float** d_all_scores) {
    if (!c || !L) return fail(c, LAFIS_ERR_ARG, "bad argument");
    int rc = run_match(c, L, topk);
    if (rc != LAFIS_OK) return rc;
    LAFIS_CUDA(c, cudaStreamSynchronize(c->stream));
    collect_times(c);
    if (d_hits) *d_hits = topk > 0 ? (const void*)c->hits.p : nullptr;
    if (d_all_scores) *d_all_scores = c->final_scores.p;
    return LAFIS_OK;
}

int lafis_match(lafis_ctx* c, lafis_latents* L, int topk, lafis_hit* hits, float* all_scores, float* components) {
    if (!c || !L) return fail(c, LAFIS_ERR_ARG, "bad argument");
    if (!hits) topk = 0;
    int rc = run_match(c, L, topk);
    if (rc != LAFIS_OK) return rc;
    const size_t QG = (size_t)L->n * c->gal.n;
    if (hits && topk > 0)
        LAFIS_CUDA(c, cudaMemcpyAsync(hits, c->hits.p, sizeof(lafis_hit) * (size_t)L->n * topk, cudaMemcpyDeviceToHost,
                                      c->stream));
    if (all_scores)
        LAFIS_CUDA(c, cudaMemcpyAsync(all_scores, c->final_scores.p, sizeof(float) * QG, cudaMemcpyDeviceToHost, c->stream));
    if (components)
        LAFIS_CUDA(c, cudaMemcpyAsync(components, c->comp.p, sizeof(float) * QG * 4, cudaMemcpyDeviceToHost, c->stream));
    LAFIS_CUDA(c, cudaStreamSynchronize(c->stream));
    collect_times(c);
    return LAFIS_OK;
}

int lafis_correspondences(lafis_ctx* c, lafis_latents* L, int q, int gallery_index, int16_t* xy_out, int* counts_out) {
    if (!c || !L || !xy_out || !counts_out) return fail(c, LAFIS_ERR_ARG, "bad argument");
    static_assert(sizeof(short4) == 4 * sizeof(int16_t), "short4 layout");
    for (int s = 0; s < 3; ++s) counts_out[s] = 0;
    if (q >= 0 && q < L->n && L->status[q] != LAFIS_OK) return L->status[q];
    return run_correspondences(c, L, q, gallery_index, reinterpret_cast<short4*>(xy_out), counts_out);
}

int lafis_merge_hits_device(lafis_ctx* c, const void* d_gathered, int n_latents, int n_lists, int topk, void* d_out) {
    if (!c || !d_gathered || !d_out || n_latents <= 0 || n_lists <= 0 || topk <= 0 || n_lists * topk > kTopkChunk)
        return fail(c, LAFIS_ERR_ARG, "merge_hits_device: n_lists * topk must be in [1, %d]", kTopkChunk);
    LAFIS_CUDA(c, cudaSetDevice(c->device));
    merge_hits_kernel<<<n_latents, kTopkThreads, 0, c->stream>>>(reinterpret_cast<const HitDev*>(d_gathered), n_latents,
                                                                  n_lists, topk, reinterpret_cast<HitDev*>(d_out));
    c->stats.kernel_launches += 1;
    LAFIS_CUDA(c, cudaGetLastError());
    return LAFIS_OK;
}

int lafis_merge_hits(const lafis_hit* shard_hits, int n_latents, int n_lists, int topk, lafis_hit* out) {
    if (!shard_hits || !out || n_latents < 0 || n_lists <= 0 || topk <= 0) return LAFIS_ERR_ARG;
    std::vector<lafis_hit> all((size_t)n_lists * topk);
    for (int q = 0; q < n_latents; ++q) {
        const lafis_hit* src = shard_hits + (size_t)q * n_lists * topk;
        all.assign(src, src + (size_t)n_lists * topk);
        std::sort(all.begin(), all.end(), [](const lafis_hit& a, const lafis_hit& b) {
            const bool ae = a.index == 0xffffffffu, be = b.index == 0xffffffffu;  // empty slots last
            if (ae != be) return be;
            if (a.score != b.score) return a.score > b.score;
            return a.index < b.index;
        });
        std::copy(all.begin(), all.begin() + topk, out + (size_t)q * topk);
    }
    return LAFIS_OK;
}

int lafis_pq_encode(lafis_ctx* c, const float* des, int64_t n, uint8_t* codes, int on_device) {
    if (!c || !des || !codes || n < 0) return fail(c, LAFIS_ERR_ARG, "bad argument");
    if (n == 0) return LAFIS_OK;
    LAFIS_CUDA(c, cudaSetDevice(c->device));
    const float* d_des = des;
    uint8_t* d_codes = codes;
    const size_t in_bytes = sizeof(float) * kDesLen * (size_t)n, out_bytes = (size_t)n * kSubs;
    if (!on_device) {
        // staging buffers owned by the context: they only grow, so a stream of enrollment calls performs no
        // allocation; the descriptors pass through pinned memory so that the copies are truly asynchronous
        LAFIS_CUDA(c, c->enroll_in.reserve(in_bytes));
        LAFIS_CUDA(c, c->enroll_out.reserve(out_bytes));
        LAFIS_CUDA(c, c->enroll_pin.reserve(in_bytes + out_bytes));
        std::memcpy(c->enroll_pin.p, des, in_bytes);
        LAFIS_CUDA(c, cudaMemcpyAsync(c->enroll_in.p, c->enroll_pin.p, in_bytes, cudaMemcpyHostToDevice, c->stream));
        d_des = reinterpret_cast<const float*>(c->enroll_in.p);
        d_codes = c->enroll_out.p;
    }
    const int pts_per_block = 256 / 16;
    pq_encode_kernel<<<(unsigned)((n + pts_per_block - 1) / pts_per_block), 256, 0, c->stream>>>(d_des, (long long)n,
                                                                                               c->d_codebook, d_codes);
    c->stats.kernel_launches += 1;
    LAFIS_CUDA(c, cudaGetLastError());
    if (!on_device)
        LAFIS_CUDA(c, cudaMemcpyAsync(c->enroll_pin.p + in_bytes, d_codes, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    LAFIS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (!on_device) std::memcpy(codes, c->enroll_pin.p + in_bytes, out_bytes);
    return LAFIS_OK;
}

int lafis_compnet_load(lafis_ctx* c, const lafis_compnet_weights* w) {
    if (!c || !w) return fail(c, LAFIS_ERR_ARG, "bad argument");
    for (int l = 0; l < 4; ++l)
        if (!w->weight[l] || !w->bias[l] || !w->bn_weight[l] || !w->bn_bias[l] || !w->bn_mean[l] || !w->bn_var[l])
            return fail(c, LAFIS_ERR_ARG, "CompNet layer %d: missing tensor", l);
    LAFIS_CUDA(c, cudaSetDevice(c->device));
    // k-major copies of the torch [out][in] weights; Linear bias and eval-mode BatchNorm1d
    // ((v - mean) / sqrt(var + eps) * gamma + beta) folded into one scale and shift per output
    std::vector<float> h((size_t)kCompWeightFloats + kCompAffFloats);
    size_t off = 0;
    for (int l = 0; l < 4; ++l) {
        const int in = l == 0 ? kCompIn : kCompOut;
        for (int k = 0; k < in; ++k)
            for (int o = 0; o < kCompOut; ++o) h[off + (size_t)k * kCompOut + o] = w->weight[l][(size_t)o * in + k];
        off += (size_t)in * kCompOut;
    }
    for (int l = 0; l < 4; ++l)
        for (int o = 0; o < kCompOut; ++o) {
            const double scale = (double)w->bn_weight[l][o] / std::sqrt((double)w->bn_var[l][o] + (double)w->bn_eps);
            h[off + (size_t)(l * 2 + 0) * kCompOut + o] = (float)scale;
            h[off + (size_t)(l * 2 + 1) * kCompOut + o] =
                (float)(((double)w->bias[l][o] - (double)w->bn_mean[l][o]) * scale + (double)w->bn_bias[l][o]);
        }
    if (!c->d_compnet) LAFIS_CUDA(c, cudaMalloc(&c->d_compnet, sizeof(float) * h.size()));
    LAFIS_CUDA(c, cudaMemcpyAsync(c->d_compnet, h.data(), sizeof(float) * h.size(), cudaMemcpyHostToDevice, c->stream));
    LAFIS_CUDA(c, cudaStreamSynchronize(c->stream));
    return LAFIS_OK;
}

int lafis_compress_descriptors(lafis_ctx* c, const float* des_in, int64_t n, float* des_out, int normalise,
                               int on_device) {
    if (!c || !des_in || !des_out || n < 0) return fail(c, LAFIS_ERR_ARG, "bad argument");
    if (!c->d_compnet) return fail(c, LAFIS_ERR_ARG, "lafis_compnet_load has not been called on this context");
    if (on_device && (((uintptr_t)des_in | (uintptr_t)des_out) & 15u))
        return fail(c, LAFIS_ERR_ARG, "device descriptors must be 16-byte aligned");
    if (n == 0) return LAFIS_OK;
    LAFIS_CUDA(c, cudaSetDevice(c->device));
    const float* d_in = des_in;
    float* d_out = des_out;
    const size_t in_bytes = sizeof(float) * kCompIn * (size_t)n, out_bytes = sizeof(float) * kCompOut * (size_t)n;
    LAFIS_CUDA(c, c->compnet_h1.reserve((size_t)n * kCompOut));
    if (!on_device) {  // persistent staging, as in lafis_pq_encode
        LAFIS_CUDA(c, c->enroll_in.reserve(in_bytes));
        LAFIS_CUDA(c, c->enroll_out.reserve(out_bytes));
        LAFIS_CUDA(c, c->enroll_pin.reserve(in_bytes + out_bytes));
        std::memcpy(c->enroll_pin.p, des_in, in_bytes);
        LAFIS_CUDA(c, cudaMemcpyAsync(c->enroll_in.p, c->enroll_pin.p, in_bytes, cudaMemcpyHostToDevice, c->stream));
        d_in = reinterpret_cast<const float*>(c->enroll_in.p);
        d_out = reinterpret_cast<float*>(c->enroll_out.p);
    }
    CompNetParams P;
    P.x = d_in;
    P.n = (long long)n;
    P.h1 = c->compnet_h1.p;
    P.out = d_out;
    P.wt = c->d_compnet;
    P.aff = c->d_compnet + kCompWeightFloats;
    P.normalise = normalise;
    const long long groups = (n + kCompPts - 1) / kCompPts;
    const unsigned grid1 = (unsigned)std::min<long long>(c->sm_count, (groups + kCompWarpsL1 - 1) / kCompWarpsL1);
    const unsigned grid2 = (unsigned)std::min<long long>(c->sm_count, (groups + kCompWarpsL234 - 1) / kCompWarpsL234);
    compnet_l1_kernel<<<grid1, kCompWarpsL1 * 32, compnet_l1_smem_bytes(), c->stream>>>(P);
    compnet_l234_kernel<<<grid2, kCompWarpsL234 * 32, compnet_l234_smem_bytes(), c->stream>>>(P);
    c->stats.kernel_launches += 2;
    LAFIS_CUDA(c, cudaGetLastError());
    if (!on_device)
        LAFIS_CUDA(c, cudaMemcpyAsync(c->enroll_pin.p + in_bytes, d_out, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    LAFIS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (!on_device) std::memcpy(des_out, c->enroll_pin.p + in_bytes, out_bytes);
    return LAFIS_OK;
}

}  // extern "C"
