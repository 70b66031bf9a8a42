// Multi-GPU: the gallery sharded over ranks (contiguous index ranges), every rank scores its shard, ONE
// ncclAllGather of the per-shard rank lists followed by the device-side merge gives every rank the global list
// (SURVEY.md §8e).  The N-vs-N drivers additionally gather the full score matrix to one rank (grouped
// ncclSend / ncclRecv, unequal shard sizes).
//
//   reference: the gallery loop that is sharded      matching/matcher.cpp:168-190 (N-vs-N), :273-295 (1-vs-N)
//              rank list over the whole gallery      matching/matcher.cpp:306-309
//              score rows over the whole gallery     matching/matcher.cpp:198-205
//
// Two ways to use it, same code underneath:
//   * one process per GPU (torchrun-style): every process creates its context, rank 0 calls
//     lafis_comm_unique_id(), the id travels by whatever the launcher offers, every rank calls lafis_comm_init();
//   * one process driving several GPUs (`match -gpus N`): lafis_group_create() makes one context per device and
//     joins them with the same calls from one host thread per device.
// NCCL is bound at run time (dlopen of libnccl.so.2): a process that already carries a copy - PyTorch bundles its
// own - keeps using that one, and single-GPU users need no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <thread>

#include "dat_format.h"
#include "lafis_internal.h"

using namespace lafis;

namespace lafis {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string why;
};

static NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
            return;
        }
        auto sym = [&](const char* n) {
            void* p = dlsym(api.handle, n);
            if (!p && api.why.empty()) api.why = std::string("libnccl: missing symbol ") + n;
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
        api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    });
    return api;
}

struct CommState {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    uint32_t* d_sizes = nullptr;        // [world][2] (index_base, n) of every shard, all-gathered
    std::vector<uint32_t> sizes;        // host copy
};

void comm_release(lafis_ctx* c) {
    if (!c || !c->comm) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm->comm && nccl().CommDestroy) nccl().CommDestroy(c->comm->comm);
    cudaFree(c->comm->d_sizes);
    delete c->comm;
    c->comm = nullptr;
}

}  // namespace lafis

namespace {

#define LAFIS_NCCL(c, expr)                                                                                   \
    do {                                                                                                      \
        ncclResult_t r__ = (expr);                                                                            \
        if (r__ != ncclSuccess)                                                                               \
            return fail((c), LAFIS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, nccl().GetErrorString(r__), __FILE__, \
                        __LINE__);                                                                            \
    } while (0)

// (index_base, n) of every shard, exchanged with one small all-gather; every rank must call
int exchange_sizes(lafis_ctx* c) {
    CommState* cs = c->comm;
    const uint32_t mine[2] = {c->index_base, (uint32_t)c->gal.n};
    LAFIS_CUDA(c, cudaMemcpyAsync(cs->d_sizes + 2 * cs->rank, mine, sizeof mine, cudaMemcpyHostToDevice, c->stream));
    LAFIS_CUDA(c, cudaStreamSynchronize(c->stream));  // `mine` is a stack temporary
    LAFIS_NCCL(c, nccl().AllGather(cs->d_sizes + 2 * cs->rank, cs->d_sizes, 2, ncclUint32, cs->comm, c->stream));
    cs->sizes.resize(2 * (size_t)cs->world);
    LAFIS_CUDA(c, cudaMemcpyAsync(cs->sizes.data(), cs->d_sizes, sizeof(uint32_t) * cs->sizes.size(), cudaMemcpyDeviceToHost,
                                  c->stream));
    LAFIS_CUDA(c, cudaStreamSynchronize(c->stream));
    return LAFIS_OK;
}

// run_match on the local shard, then the exchange; results: c->merged (global rank lists, every rank) and, on `root`
// when gather_scores is set, the per-rank score blocks in c->gather_scores
int sharded_enqueue(lafis_ctx* c, lafis_latents* L, int topk, int gather_scores, int root, bool* have_scores) {
    CommState* cs = c->comm;
    *have_scores = false;
    if (!c->gallery_set) return fail(c, LAFIS_ERR_NO_GALLERY, "no gallery shard resident on rank %d", cs ? cs->rank : 0);
    const int world = cs ? cs->world : 1;
    if (topk > 0 && (size_t)world * topk > (size_t)kMergeCap)
        return fail(c, LAFIS_ERR_ARG, "world size x topk must not exceed %d", kMergeCap);
    int rc = run_match(c, L, topk);
    if (rc != LAFIS_OK) return rc;
    const int Q = L->n;
    if (topk > 0) {
        const size_t per = (size_t)Q * topk;
        LAFIS_CUDA(c, c->merged.reserve(per));
        if (world > 1) {
            LAFIS_CUDA(c, c->gathered.reserve(per * world));
            // the ONE collective of the path: per-shard rank lists, 8 bytes per entry
            LAFIS_NCCL(c, nccl().AllGather(c->hits.p, c->gathered.p, per, ncclUint64, cs->comm, c->stream));
            rc = lafis_merge_hits_device(c, c->gathered.p, Q, world, topk, c->merged.p);
            if (rc != LAFIS_OK) return rc;
        } else {
            LAFIS_CUDA(c, cudaMemcpyAsync(c->merged.p, c->hits.p, per * sizeof(HitDev), cudaMemcpyDeviceToDevice, c->stream));
        }
    }
    if (gather_scores && world > 1) {
        if (root < 0 || root >= world) return fail(c, LAFIS_ERR_ARG, "root %d outside [0, %d)", root, world);
        rc = exchange_sizes(c);
        if (rc != LAFIS_OK) return rc;
        const size_t own = (size_t)Q * c->gal.n;
        if (cs->rank == root) {
            size_t tot = 0;
            for (int r = 0; r < world; ++r)
                if (r != root) tot += (size_t)Q * cs->sizes[2 * r + 1];
            LAFIS_CUDA(c, c->gather_scores.reserve(tot));
            LAFIS_NCCL(c, nccl().GroupStart());
            size_t at = 0;
            for (int r = 0; r < world; ++r) {
                if (r == root) continue;
                const size_t cnt = (size_t)Q * cs->sizes[2 * r + 1];
                if (cnt) LAFIS_NCCL(c, nccl().Recv(c->gather_scores.p + at, cnt, ncclFloat, r, cs->comm, c->stream));
                at += cnt;
            }
            LAFIS_NCCL(c, nccl().GroupEnd());
        } else if (own) {
            LAFIS_NCCL(c, nccl().Send(c->final_scores.p, own, ncclFloat, root, cs->comm, c->stream));
        }
        *have_scores = cs->rank == root;
    } else if (gather_scores) {
        *have_scores = true;
    }
    LAFIS_CUDA(c, cudaEventRecord(c->ev1, c->stream));  // the exchange is part of the match time
    return LAFIS_OK;
}

// root: rows [q][G_total] assembled from the shard blocks
int download_scores(lafis_ctx* c, int Q, float* all_scores) {
    CommState* cs = c->comm;
    const int world = cs ? cs->world : 1;
    if (world == 1) {
        if ((size_t)Q * c->gal.n)
            LAFIS_CUDA(c, cudaMemcpyAsync(all_scores, c->final_scores.p, sizeof(float) * (size_t)Q * c->gal.n,
                                          cudaMemcpyDeviceToHost, c->stream));
        return LAFIS_OK;
    }
    size_t g_total = 0;
    for (int r = 0; r < world; ++r) g_total += cs->sizes[2 * r + 1];
    size_t at = 0;
    for (int r = 0; r < world; ++r) {
        const size_t gr = cs->sizes[2 * r + 1], base = cs->sizes[2 * r];
        if (base + gr > g_total) return fail(c, LAFIS_ERR_ARG, "shard %d [%zu, %zu) lies outside the gallery of %zu", r, base, base + gr, g_total);
        const float* src = (r == cs->rank) ? c->final_scores.p : c->gather_scores.p + at;
        if (r != cs->rank) at += (size_t)Q * gr;
        if (gr == 0) continue;
        LAFIS_CUDA(c, cudaMemcpy2DAsync(all_scores + base, sizeof(float) * g_total, src, sizeof(float) * gr, sizeof(float) * gr,
                                        (size_t)Q, cudaMemcpyDeviceToHost, c->stream));
    }
    return LAFIS_OK;
}

template <typename F>
void for_each_rank(int n, F&& fn) {  // one host thread per device (collectives block until every rank has joined)
    std::vector<std::thread> pool;
    for (int i = 1; i < n; ++i) pool.emplace_back(fn, i);
    fn(0);
    for (std::thread& t : pool) t.join();
}

int group_fail(lafis_group* g, int code, const std::string& msg) {
    if (g) g->err = msg;
    return code;
}

// first failing rank's status and message
int group_status(lafis_group* g, const std::vector<int>& rcs, const char* what) {
    for (size_t i = 0; i < rcs.size(); ++i)
        if (rcs[i] != LAFIS_OK)
            return group_fail(g, rcs[i], std::string(what) + " failed on device " + std::to_string(g->ctx[i]->device) + ": " +
                                             lafis_last_error(g->ctx[i]));
    return LAFIS_OK;
}

}  // namespace

extern "C" {

int lafis_comm_unique_id(void* id_out) {
    if (!id_out) return LAFIS_ERR_ARG;
    static_assert(sizeof(ncclUniqueId) == LAFIS_COMM_ID_BYTES, "LAFIS_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");
    NcclApi& n = nccl();
    if (!n.handle || !n.why.empty()) return fail(nullptr, LAFIS_ERR_CUDA, "NCCL unavailable: %s", n.why.c_str());
    ncclUniqueId id;
    ncclResult_t r = n.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, LAFIS_ERR_CUDA, "ncclGetUniqueId failed: %s", n.GetErrorString(r));
    std::memcpy(id_out, &id, sizeof id);
    return LAFIS_OK;
}

int lafis_comm_init(lafis_ctx* c, const void* id, int rank, int world) {
    if (!c || !id || world < 1 || rank < 0 || rank >= world) return fail(c, LAFIS_ERR_ARG, "bad communicator argument");
    NcclApi& n = nccl();
    if (!n.handle || !n.why.empty()) return fail(c, LAFIS_ERR_CUDA, "NCCL unavailable: %s", n.why.c_str());
    comm_release(c);
    LAFIS_CUDA(c, cudaSetDevice(c->device));
    CommState* cs = new CommState();
    cs->rank = rank;
    cs->world = world;
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof uid);
    ncclResult_t r = n.CommInitRank(&cs->comm, world, uid, rank);
    if (r != ncclSuccess) {
        delete cs;
        return fail(c, LAFIS_ERR_CUDA, "ncclCommInitRank(rank %d of %d) failed: %s", rank, world, n.GetErrorString(r));
    }
    if (cudaMalloc(&cs->d_sizes, sizeof(uint32_t) * 2 * (size_t)world) != cudaSuccess) {
        n.CommDestroy(cs->comm);
        delete cs;
        return fail(c, LAFIS_ERR_CUDA, "cudaMalloc failed");
    }
    c->comm = cs;
    return LAFIS_OK;
}

int lafis_comm_rank(const lafis_ctx* c) { return (c && c->comm) ? c->comm->rank : -1; }
int lafis_comm_world(const lafis_ctx* c) { return (c && c->comm) ? c->comm->world : 1; }
void lafis_comm_destroy(lafis_ctx* c) { comm_release(c); }

int lafis_comm_nccl_version(void) {
    NcclApi& n = nccl();
    int v = 0;
    if (n.handle && n.GetVersion) n.GetVersion(&v);
    return v;
}

int lafis_gallery_total(lafis_ctx* c, uint32_t* shard_base_out, uint32_t* shard_n_out) {
    if (!c) return LAFIS_ERR_ARG;
    if (!c->comm || c->comm->world == 1) {
        if (shard_base_out) shard_base_out[0] = c->index_base;
        if (shard_n_out) shard_n_out[0] = (uint32_t)c->gal.n;
        return c->gal.n;
    }
    LAFIS_CUDA(c, cudaSetDevice(c->device));
    const int rc = exchange_sizes(c);
    if (rc != LAFIS_OK) return rc;
    long long tot = 0;
    for (int r = 0; r < c->comm->world; ++r) {
        if (shard_base_out) shard_base_out[r] = c->comm->sizes[2 * r];
        if (shard_n_out) shard_n_out[r] = c->comm->sizes[2 * r + 1];
        tot += c->comm->sizes[2 * r + 1];
    }
    return (int)tot;
}

int lafis_match_sharded_device(lafis_ctx* c, lafis_latents* L, int topk, const void** d_hits) {
    if (!c || !L) return fail(c, LAFIS_ERR_ARG, "bad argument");
    bool have = false;
    const int rc = sharded_enqueue(c, L, topk, 0, 0, &have);
    if (rc != LAFIS_OK) return rc;
    LAFIS_CUDA(c, cudaStreamSynchronize(c->stream));
    collect_times(c);
    if (d_hits) *d_hits = topk > 0 ? (const void*)c->merged.p : nullptr;
    return LAFIS_OK;
}

int lafis_match_sharded(lafis_ctx* c, lafis_latents* L, int topk, lafis_hit* hits, int gather_scores, float* all_scores,
                        int root) {
    if (!c || !L) return fail(c, LAFIS_ERR_ARG, "bad argument");
    if (!hits) topk = 0;
    bool have = false;
    int rc = sharded_enqueue(c, L, topk, gather_scores, root, &have);
    if (rc != LAFIS_OK) return rc;
    if (hits && topk > 0)
        LAFIS_CUDA(c, cudaMemcpyAsync(hits, c->merged.p, sizeof(lafis_hit) * (size_t)L->n * topk, cudaMemcpyDeviceToHost,
                                      c->stream));
    if (have && all_scores) {
        rc = download_scores(c, L->n, all_scores);
        if (rc != LAFIS_OK) return rc;
    }
    LAFIS_CUDA(c, cudaStreamSynchronize(c->stream));
    collect_times(c);
    return LAFIS_OK;
}

// ---------------------------------------------------------------------------------------------------
// one process, several devices
// ---------------------------------------------------------------------------------------------------
int lafis_group_create(const char* codebook_path, const int* devices, int n_devices, lafis_group** out) {
    if (!codebook_path || !out || n_devices < 1 || n_devices > 64) return LAFIS_ERR_ARG;
    lafis_group* g = new lafis_group();
    for (int i = 0; i < n_devices; ++i) {
        lafis_ctx* c = nullptr;
        const int rc = lafis_create(codebook_path, devices ? devices[i] : i, &c);
        if (rc != LAFIS_OK) {
            lafis_group_destroy(g);
            return rc;
        }
        g->ctx.push_back(c);
    }
    if (n_devices > 1) {
        unsigned char id[LAFIS_COMM_ID_BYTES];
        int rc = lafis_comm_unique_id(id);
        if (rc == LAFIS_OK) {
            std::vector<int> rcs(n_devices, LAFIS_OK);
            for_each_rank(n_devices, [&](int i) { rcs[i] = lafis_comm_init(g->ctx[i], id, i, n_devices); });
            rc = group_status(g, rcs, "lafis_comm_init");
            if (rc != LAFIS_OK) fail(nullptr, rc, "%s", g->err.c_str());
        }
        if (rc != LAFIS_OK) {
            lafis_group_destroy(g);
            return rc;
        }
    }
    g->base.assign(n_devices + 1, 0);
    *out = g;
    return LAFIS_OK;
}

void lafis_group_destroy(lafis_group* g) {
    if (!g) return;
    // communicators first (ncclCommDestroy of one rank may wait for its peers), then the contexts
    const int n = (int)g->ctx.size();
    if (n > 1) for_each_rank(n, [&](int i) { lafis_comm_destroy(g->ctx[i]); });
    for (lafis_ctx* c : g->ctx) lafis_destroy(c);
    delete g;
}

int lafis_group_size(const lafis_group* g) { return g ? (int)g->ctx.size() : 0; }
lafis_ctx* lafis_group_ctx(lafis_group* g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[i] : nullptr; }
const char* lafis_group_last_error(const lafis_group* g) { return g ? g->err.c_str() : lafis_last_error(nullptr); }

int lafis_group_gallery_load_files(lafis_group* g, const char* const* paths, int n) {
    if (!g || (!paths && n > 0) || n < 0) return group_fail(g, LAFIS_ERR_ARG, "bad argument");
    if (n == 0) return group_fail(g, LAFIS_ERR_NO_TEMPLATES, "no rolled templates");
    const int w = (int)g->ctx.size();
    std::vector<int> rcs(w, LAFIS_OK);
    for_each_rank(w, [&](int i) { rcs[i] = lafis_gallery_load_files(g->ctx[i], paths, n, i, w); });
    const int rc = group_status(g, rcs, "lafis_gallery_load_files");
    if (rc != LAFIS_OK) return rc;
    for (int i = 0; i <= w; ++i) g->base[i] = (uint32_t)((long long)n * i / w);
    return LAFIS_OK;
}

int lafis_group_gallery_load_dir(lafis_group* g, const char* dir) {
    if (!g || !dir) return group_fail(g, LAFIS_ERR_ARG, "bad argument");
    std::vector<std::string> files = list_dat_files(dir);
    if (files.empty()) return group_fail(g, LAFIS_ERR_NO_TEMPLATES, std::string("No rolled templates found in directory: ") + dir);
    std::vector<const char*> ptrs(files.size());
    for (size_t i = 0; i < files.size(); ++i) ptrs[i] = files[i].c_str();
    const int rc = lafis_group_gallery_load_files(g, ptrs.data(), (int)ptrs.size());
    if (rc == LAFIS_OK)
        for (size_t i = 0; i < g->ctx.size(); ++i) {
            g->ctx[i]->gallery_dir = dir;
            g->ctx[i]->gallery_dir_shard = (int)i;
            g->ctx[i]->gallery_dir_shards = (int)g->ctx.size();
        }
    return rc;
}

int lafis_group_gallery_size(const lafis_group* g) { return g ? (int)g->base.back() : 0; }

int lafis_group_match(lafis_group* g, lafis_latents* L, int topk, lafis_hit* hits, float* all_scores) {
    if (!g || !L) return group_fail(g, LAFIS_ERR_ARG, "bad argument");
    const int w = (int)g->ctx.size();
    if (!hits) topk = 0;
    std::vector<int> rcs(w, LAFIS_OK);
    // every rank needs a list buffer for the collective to be entered; only rank 0's goes back to the caller
    std::vector<std::vector<lafis_hit>> scratch(w);
    for_each_rank(w, [&](int i) {
        lafis_hit* h = hits;
        if (i != 0 && topk > 0) {
            scratch[i].resize((size_t)lafis_latents_count(L) * topk);
            h = scratch[i].data();
        }
        rcs[i] = lafis_match_sharded(g->ctx[i], L, topk, h, all_scores ? 1 : 0, i == 0 ? all_scores : nullptr, 0);
    });
    return group_status(g, rcs, "lafis_match_sharded");
}

}  // extern "C"
