/*
 * latentafis_b200 — C ABI of the B200-native MSU-LatentAFIS 1-vs-N gallery matcher.
 *
 * The reference has no FFI: its in-process surface is the C++ class PQ::Matcher
 * (matching/matcher.h:34-75) driven by matching/main.cpp.  This header is the C-ABI equivalent of
 * that surface for the hot path; every entry point names the reference member it
 * replaces.  Plain pointers and sizes only; no torch / CUDA types.  All functions return an
 * LAFIS_* status; lafis_last_error() gives a human-readable message for the last failure on a
 * context.  A context may be used from one host thread at a time.  The library requires a CUDA
 * device of compute capability 10.x; there is no CPU fallback — every entry point that computes
 * fails with LAFIS_ERR_CUDA when the device path is unavailable.
 */
#ifndef LATENTAFIS_B200_H
#define LATENTAFIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAFIS_API __attribute__((visibility("default")))

/* ---- status codes.  Positive values mirror the reference's own return conventions
 *      (matcher.cpp:296-299 driver "latent empty" = 1; loaders :798-801, :837-846, :865-869). ---- */
enum {
    LAFIS_OK = 0,
    LAFIS_LATENT_EMPTY = 1,        /* One2List_matching returns 1: latent has nothing to match with */
    LAFIS_ERR_NO_TEMPLATES = -1,   /* drivers return -1 when a directory holds no *.dat */
    LAFIS_ERR_ARG = -2,
    LAFIS_ERR_IO = -3,
    LAFIS_ERR_CODEBOOK = -4,       /* not a 16 x 256 x 6 codebook */
    LAFIS_ERR_CUDA = -5,
    LAFIS_ERR_UNSUPPORTED_SIZE = -6, /* more than 2000 points in a template: the reference's loaders reject it too
                                        (matcher.cpp:837-841, :865-869) */
    LAFIS_ERR_LATENT_LAYOUT = -7,  /* score[28] would be read out of bounds by the reference
                                      (matcher.cpp:188): latent has < 29 template slots */
    LAFIS_ERR_NO_GALLERY = -8
};

/* per-template load status kept for every gallery entry (matcher.cpp:886-983 return values) */
enum {
    LAFIS_TPL_OK = 0,
    LAFIS_TPL_EMPTY = 1,      /* file <= 10 bytes, or nothing usable inside: score stays -1 */
    LAFIS_TPL_TRUNCATED = 2,  /* loader returned 2 or 4 part-way: templates read so far are kept */
    LAFIS_TPL_FAILED = -1     /* loader returned < 0: the drivers empty the template, score -1 */
};

typedef struct lafis_ctx lafis_ctx;

/* one rank-list entry; ordering is (score descending, gallery index ascending) */
typedef struct {
    float score;
    uint32_t index; /* global gallery index (shard base + local index) */
} lafis_hit;

/* A gallery in structure-of-arrays form, host or device memory.  Template g owns minutiae
 * [minu_off[g], minu_off[g+1]) and texture points [tex_off[g], tex_off[g+1]).  Descriptors are
 * row-major as in the .dat files (matcher.cpp:948-953, :971-975); the library re-lays them out for
 * the kernels.  Texture point counts above 1000 are truncated like matcher.cpp:546-547. */
typedef struct {
    int32_t n_templates;
    const uint32_t* minu_off;  /* [n_templates+1] */
    const int16_t* minu_x;     /* pixels */
    const int16_t* minu_y;
    const float* minu_ori;
    const float* minu_des;     /* [total_minu][96] */
    const uint32_t* tex_off;   /* [n_templates+1] */
    const int16_t* tex_x;      /* block units */
    const int16_t* tex_y;
    const float* tex_ori;
    const uint8_t* tex_codes;  /* [total_tex][16] */
    const int8_t* status;      /* [n_templates] LAFIS_TPL_*, may be NULL (= all OK) */
    int32_t on_device;         /* 0: host pointers; 1: device pointers on the context's GPU
                                  (minu_off / tex_off / status are always host pointers) */
} lafis_packed_gallery;

/* A batch of latent prints, host memory.  Only what the matcher reads is carried: the three
 * selected minutiae templates {26, 2, 11} (matcher.cpp:380) and texture template 0 (:411).
 * Slot s of latent q owns minutiae [minu_off[3q+s], minu_off[3q+s+1]). */
typedef struct {
    int32_t n_latents;
    const int32_t* n_minu_templates; /* [n] non-empty minutiae templates in the file (28 expected) */
    const int32_t* n_tex_templates;  /* [n] */
    const uint32_t* minu_off;        /* [3n+1] */
    const int16_t* minu_x;
    const int16_t* minu_y;
    const float* minu_ori;
    const float* minu_des;           /* [total][96] */
    const uint32_t* tex_off;         /* [n+1] */
    const int16_t* tex_x;
    const int16_t* tex_y;
    const float* tex_ori;
    const float* tex_des;            /* [total][96] */
} lafis_packed_latents;

/* ---- context: replaces PQ::Matcher::Matcher(code_file), matcher.cpp:31-94 ---- */
LAFIS_API int lafis_create(const char* codebook_path, int device, lafis_ctx** out);
/* same, from an in-memory 16 x 256 x 6 float codebook */
LAFIS_API int lafis_create_from_codebook(const float* codewords, int subs, int clusters, int sub_dim, int device,
                                         lafis_ctx** out);
LAFIS_API void lafis_destroy(lafis_ctx* ctx);
LAFIS_API const char* lafis_last_error(const lafis_ctx* ctx);
LAFIS_API const char* lafis_version(void);

/* ---- gallery ingest: replaces the per-pair load_FP_template(rolled) inside the hot loop,
 *      matcher.cpp:173 / :278 / :886-983.  The gallery stays resident in HBM across matches.
 *      shard_rank / shard_count select a contiguous slice [floor(G*r/c), floor(G*(r+1)/c)) of
 *      the template list for multi-GPU runs; indices reported in hits stay global. ---- */
LAFIS_API int lafis_gallery_load_dir(lafis_ctx* ctx, const char* dir, int shard_rank, int shard_count);
LAFIS_API int lafis_gallery_load_files(lafis_ctx* ctx, const char* const* paths, int n, int shard_rank,
                                       int shard_count);
LAFIS_API int lafis_gallery_set_packed(lafis_ctx* ctx, const lafis_packed_gallery* g, uint32_t index_base);
LAFIS_API int lafis_gallery_size(const lafis_ctx* ctx);          /* templates resident on this context */
LAFIS_API const char* lafis_gallery_path(const lafis_ctx* ctx, int local_index); /* "" for packed input */
LAFIS_API int lafis_gallery_status(const lafis_ctx* ctx, int local_index);
LAFIS_API uint64_t lafis_gallery_bytes(const lafis_ctx* ctx);    /* algorithmic bytes resident (392*nRm + 24*nRt) */

/* read one resident template back in .dat form (tests / oracle sampling); arrays may be NULL to
 * query the counts only.  des is row-major [n_minu][96], codes [n_tex][16]. */
LAFIS_API int lafis_gallery_get_template(const lafis_ctx* ctx, int local_index, int* n_minu, int16_t* mx, int16_t* my,
                                         float* mori, float* mdes, int* n_tex, int16_t* tx, int16_t* ty, float* tori,
                                         uint8_t* tcodes);

/* ---- latent parsing: replaces load_FP_template(latent), matcher.cpp:785-884 (without the LUT,
 *      which the device builds).  The returned handle owns a host-side packed batch. ---- */
typedef struct lafis_latents lafis_latents;
LAFIS_API int lafis_latents_load_files(lafis_ctx* ctx, const char* const* paths, int n, lafis_latents** out);
LAFIS_API int lafis_latents_from_packed(lafis_ctx* ctx, const lafis_packed_latents* p, lafis_latents** out);
LAFIS_API int lafis_latents_count(const lafis_latents* l);
LAFIS_API uint64_t lafis_latents_bytes(const lafis_latents* l); /* bytes one host-to-device staging copy moves */
LAFIS_API int lafis_latents_status(const lafis_latents* l, int q); /* LAFIS_OK, LAFIS_LATENT_EMPTY, LAFIS_ERR_LATENT_LAYOUT */
LAFIS_API int lafis_latents_minu_templates(const lafis_latents* l, int q); /* non-empty minutiae templates in the file
                                                                              (m_nrof_minu_templates, matcher.cpp:403) */
LAFIS_API void lafis_latents_free(lafis_latents* l);
/* pre-stage the batch in HBM so that a following lafis_match() performs no host-to-device copy */
LAFIS_API int lafis_latents_make_resident(lafis_ctx* ctx, lafis_latents* l);
/* A latent batch may be used with any context of the process (non-resident batches are staged per match) and may
 * be freed before or after the context it was created on. */

/* ---- the hot path: replaces the omp loop of One2List_matching / List2List_matching
 *      (matcher.cpp:168-190, :273-295) = One2One_matching_selected_templates (:376-417) for every
 *      (latent, gallery[i]) plus the fusion of :188, followed by the rank list of :306-309.
 *
 *      all_scores : optional [n_latents * gallery_size] fused scores, -1 where the reference
 *                   leaves the slot untouched; host memory.
 *      components : optional [n_latents * gallery_size * 4] = score[0], score[1], score[2], score[28].
 *      hits       : optional [n_latents * topk] rank lists of this shard; entries beyond the
 *                   gallery size are {-inf, UINT32_MAX}.
 *      Latents whose status is not LAFIS_OK get all_scores = -1 and empty hit lists. ---- */
LAFIS_API int lafis_match(lafis_ctx* ctx, lafis_latents* latents, int topk, lafis_hit* hits, float* all_scores,
                          float* components);
/* same computation, results left in HBM (device-resident timing / multi-GPU gather).  The
 * returned device pointers stay valid until the next match on this context. */
LAFIS_API int lafis_match_device(lafis_ctx* ctx, lafis_latents* latents, int topk, const void** d_hits,
                                 const float** d_all_scores);
/* ---- surviving minutiae correspondences of ONE (latent, gallery template) pair.
 *      Replaces the save_corr branch of One2One_minutiae_matching (matcher.cpp:497-505), which
 *      One2List_matching runs for the 24 best gallery templates (:322-327): for each of the three selected
 *      minutiae templates (:380) the list corr3 of LSS_R_Fast2 (:495), in the order it is written to
 *      "<corr_file>_<i>.csv".  xy_out[(slot * LAFIS_MAX_CORR + k) * 4 + {0,1,2,3}] = latent x, latent y,
 *      rolled x, rolled y of correspondence k; counts_out[slot] = number of correspondences (their
 *      similarities sum to component score[slot]).  Host memory; gallery_index is local to this context.
 *      Returns the latent's status when it is not LAFIS_OK (no output then). ---- */
#define LAFIS_MAX_CORR 120
LAFIS_API int lafis_correspondences(lafis_ctx* ctx, lafis_latents* latents, int q, int gallery_index,
                                    int16_t* xy_out /* [3][LAFIS_MAX_CORR][4] */, int* counts_out /* [3] */);

/* merge per-shard rank lists (n_lists lists of topk entries per latent, concatenated per latent)
 * into global ones with the (score desc, index asc) rule; host memory. */
LAFIS_API int lafis_merge_hits(const lafis_hit* shard_hits, int n_latents, int n_lists, int topk, lafis_hit* out);

/* same merge on the device, for the multi-GPU exchange: d_gathered is the result of an all-gather of
 * every rank's [n_latents][topk] list, i.e. [n_lists][n_latents][topk]; d_out is [n_latents][topk].
 * Enqueued on the context's stream; n_lists * topk <= 4096. */
LAFIS_API int lafis_merge_hits_device(lafis_ctx* ctx, const void* d_gathered, int n_latents, int n_lists, int topk,
                                      void* d_out);

/* ---- multi-GPU: the gallery loop of One2List_matching / List2List_matching (matcher.cpp:168-190,
 *      :273-295) sharded over ranks by contiguous gallery index ranges (lafis_gallery_load_*(rank, world)), the rank
 *      list of :306-309 rebuilt from the shards' lists with ONE ncclAllGather + lafis_merge_hits_device, and - for the
 *      score rows of :198-205 - the shards' score blocks gathered to one rank.  A context joins a communicator with
 *      lafis_comm_init: rank 0 obtains an id with lafis_comm_unique_id and hands it to the other ranks (one process
 *      per GPU, any launcher), or lafis_group_create does all of it for several devices of one process.  NCCL is
 *      bound at run time (libnccl.so.2); these calls fail with LAFIS_ERR_CUDA when it is absent. ---- */
#define LAFIS_COMM_ID_BYTES 128
LAFIS_API int lafis_comm_unique_id(void* id_out /* [LAFIS_COMM_ID_BYTES] */);
LAFIS_API int lafis_comm_init(lafis_ctx* ctx, const void* id, int rank, int world); /* collective: every rank calls */
LAFIS_API int lafis_comm_rank(const lafis_ctx* ctx);  /* -1 without a communicator */
LAFIS_API int lafis_comm_world(const lafis_ctx* ctx); /* 1 without a communicator */
LAFIS_API void lafis_comm_destroy(lafis_ctx* ctx);
LAFIS_API int lafis_comm_nccl_version(void);          /* e.g. 22703; 0 when NCCL is not loadable */
/* collective: gallery templates over all shards (>= 0) or a negative status; the optional arrays [world] receive
 * every shard's first global index and size */
LAFIS_API int lafis_gallery_total(lafis_ctx* ctx, uint32_t* shard_base_out, uint32_t* shard_n_out);
/* The sharded hot path; collective, every rank passes the same latent batch, topk, gather_scores and root.
 *   hits          every rank: [n_latents * topk] GLOBAL rank lists (may be NULL: no lists)
 *   gather_scores != 0: the fused scores of all shards are gathered to `root`
 *   all_scores    root only: [n_latents * G_total] rows in global gallery order; ignored elsewhere */
LAFIS_API int lafis_match_sharded(lafis_ctx* ctx, lafis_latents* latents, int topk, lafis_hit* hits, int gather_scores,
                                  float* all_scores, int root);
/* same, the merged global rank lists stay in HBM (valid until the next match on this context) */
LAFIS_API int lafis_match_sharded_device(lafis_ctx* ctx, lafis_latents* latents, int topk, const void** d_hits);

/* several devices of one process: one context per device, joined by a communicator; a latent batch created on any
 * of the contexts serves all of them */
typedef struct lafis_group lafis_group;
LAFIS_API int lafis_group_create(const char* codebook_path, const int* devices /* NULL: 0..n-1 */, int n_devices,
                                 lafis_group** out);
LAFIS_API void lafis_group_destroy(lafis_group* group);
LAFIS_API int lafis_group_size(const lafis_group* group);
LAFIS_API lafis_ctx* lafis_group_ctx(lafis_group* group, int i);
LAFIS_API const char* lafis_group_last_error(const lafis_group* group);
LAFIS_API int lafis_group_gallery_load_dir(lafis_group* group, const char* dir);
LAFIS_API int lafis_group_gallery_load_files(lafis_group* group, const char* const* paths, int n);
LAFIS_API int lafis_group_gallery_size(const lafis_group* group); /* over all shards */
/* lafis_match over the sharded gallery: hits [n_latents * topk] global lists, all_scores [n_latents * G_total] */
LAFIS_API int lafis_group_match(lafis_group* group, lafis_latents* latents, int topk, lafis_hit* hits, float* all_scores);
/* the two drivers below over the sharded gallery: same files, byte for byte */
LAFIS_API int lafis_group_one2list_matching(lafis_group* group, const char* latent_template_file, const char* rolled_dir,
                                            const char* score_path);
LAFIS_API int lafis_group_list2list_matching(lafis_group* group, const char* latent_dir, const char* rolled_dir,
                                             const char* score_path);

/* ---- drivers with the reference's score-file formats (SURVEY.md §8b "Score files") ----
 *      lafis_one2list_matching  replaces PQ::Matcher::One2List_matching,  matcher.cpp:216-337:
 *        writes <score_path><latent stem>.csv = "filename,score" + up to 24 rows
 *        <rank>"<rolled path>",<score>; returns 0, -1 (no *.dat in rolled_dir), 1 (latent empty).
 *        For those (up to) 24 templates it also writes the correspondence files of :322-327,
 *        "<corr_prefix>corr<latent stem>_<rolled stem>_<i>.csv" (i = 0, 1, 2; rows "lx,ly,rx,ry"), where
 *        corr_prefix is the environment variable LAFIS_CORR_PATH if set, else score_path (the reference
 *        hard-codes "/LatentAFIS/scores/").
 *      lafis_list2list_matching replaces PQ::Matcher::List2List_matching, matcher.cpp:96-214:
 *        one CSV per latent with one row "<rolled path>",<score %.3f> per gallery file in
 *        directory order, -1.000 for entries the reference leaves unscored.
 *      score_path is a prefix (the reference concatenates strings: it needs a trailing '/').
 *      The gallery parsed for a directory stays resident and is reused by later calls with the
 *      same directory; lafis_forget_gallery_dir() drops that association. ---- */
LAFIS_API int lafis_one2list_matching(lafis_ctx* ctx, const char* latent_template_file, const char* rolled_dir,
                                      const char* score_path);
LAFIS_API int lafis_list2list_matching(lafis_ctx* ctx, const char* latent_dir, const char* rolled_dir,
                                       const char* score_path);
LAFIS_API void lafis_forget_gallery_dir(lafis_ctx* ctx);

/* ---- enrollment helper: PQ-encode descriptors with the context's codebook,
 *      TrainedPQEncoder.encode_multi, extraction/descriptor_PQ.py:19-27.  des/codes are device
 *      pointers when on_device != 0. ---- */
LAFIS_API int lafis_pq_encode(lafis_ctx* ctx, const float* des, int64_t n, uint8_t* codes, int on_device);

/* ---- descriptor compression 192 -> 96: the reference's CompNet in eval mode
 *      (extraction/models/net_compress.py:33-53, BasicBlock :7-31) as run by template_compression
 *      (extraction/descriptor_DR.py:141-152), including the re-normalisation row / ||row|| * 1.73 (:150-152).
 *      The reference ships no weights: the caller passes the state_dict tensors (torch layouts, fp32) of
 *      layer l = 0..3 = layer1.0/.1, layer2.layers.0/.1, layer2.layers.3/.4, layer3.0/.1 —
 *      weight[l] is the Linear weight [96][192] (l = 0) or [96][96], bias[l] its bias [96], bn_*[l] the
 *      BatchNorm1d parameters / running statistics [96].  Host pointers; the network stays resident on the
 *      context.  lafis_compress_descriptors maps des_in [n][192] to des_out [n][96] (device pointers when
 *      on_device != 0; 16-byte aligned); normalise = 0 returns the raw network output. ---- */
typedef struct {
    const float* weight[4];
    const float* bias[4];
    const float* bn_weight[4];
    const float* bn_bias[4];
    const float* bn_mean[4];
    const float* bn_var[4];
    float bn_eps; /* torch default 1e-5 */
} lafis_compnet_weights;
LAFIS_API int lafis_compnet_load(lafis_ctx* ctx, const lafis_compnet_weights* w);
LAFIS_API int lafis_compress_descriptors(lafis_ctx* ctx, const float* des_in, int64_t n, float* des_out, int normalise,
                                         int on_device);

/* ---- enrollment of one rolled print: the tail of the reference's extraction pipeline,
 *      TrainedPQEncoder.encode_multi (descriptor_PQ.py:19-27) on the texture descriptors followed by
 *      Template2Bin_Byte_PQ_rolled (:178-272).  Coordinates arrive as the reference holds them, rows of
 *      {x, y, orientation} in pixels; minutiae x, y are truncated to u16, texture points are written in block
 *      units (u16)((x - 24) / 16) (:249-256).  The PQ codes are computed on the GPU with the context's codebook.
 *      The file is readable by the reference matcher and by lafis_gallery_load_*. ---- */
typedef struct {
    int h, w, blkH, blkW;     /* image size and ridge-flow block grid (block sizes are clamped to 50) */
    int n_minu;               /* <= 2000 are written */
    const float* minu_xyo;    /* [n_minu][3] */
    const float* minu_des;    /* [n_minu][96] */
    int n_tex;
    const float* tex_xyo;     /* [n_tex][3], pixels */
    const float* tex_des;     /* [n_tex][96], PQ-encoded on the device */
    int des_len;              /* 96 (or 0): descriptors are final.  192: raw descriptors; minutiae and texture
                                 descriptors first go through lafis_compress_descriptors on the device
                                 (needs lafis_compnet_load), arrays are then [n][192] */
} lafis_rolled_features;
LAFIS_API int lafis_enroll_rolled(lafis_ctx* ctx, const lafis_rolled_features* features, const char* out_path);

/* ---- enrollment of one latent print: Template2Bin_Byte_latent (extraction/descriptor_PQ.py:80-175), the writer
 *      of the files lafis_latents_load_files / the reference's load_FP_template(latent) read.  The reference
 *      pipeline produces 28 minutiae templates and one texture template; all are written in order (an empty
 *      minutiae template is written as its zero count, :113-117, and shifts the reader's template indices as in
 *      matcher.cpp:834-836).  Coordinates as in lafis_rolled_features: {x, y, orientation} rows in pixels,
 *      texture points are written in block units.  des_len 192: every descriptor of the print goes through
 *      lafis_compress_descriptors in ONE device call first (descriptor_DR.py:196-230). ---- */
typedef struct {
    int n;             /* points; at most 2000 are written */
    const float* xyo;  /* [n][3] */
    const float* des;  /* [n][des_len] */
} lafis_point_set;
typedef struct {
    int h, w, blkH, blkW;
    int n_minu_templates;         /* <= 255 */
    const lafis_point_set* minu;  /* [n_minu_templates] */
    int n_tex_templates;          /* <= 255 */
    const lafis_point_set* tex;   /* [n_tex_templates] */
    int des_len;                  /* 96 (or 0), or 192 (needs lafis_compnet_load) */
} lafis_latent_features;
LAFIS_API int lafis_enroll_latent(lafis_ctx* ctx, const lafis_latent_features* features, const char* out_path);

/* ---- instrumentation ---- */
typedef struct {
    uint64_t kernel_launches; /* kernels launched by this library since the context was created */
    uint64_t pairs_scored;    /* (latent, gallery) pairs scored */
    float last_match_ms;      /* device time of the last lafis_match*, CUDA events */
    float last_stage_ms[8];   /* per-kernel device times of the last match (CUDA events around the launches):
                                 0 tex_rowmax, 1 minu_sim, 2 minu_select, 3 graph_minu_sparse, 4 graph_tex (sparse +
                                 dense), 5 fuse + rank lists, 6 minu_select_slow (+ the oversized-pair kernels),
                                 7 graph_minu second chance + dense.  With the default two streams the rare-path kernels
                                 (6 and the dense part of 4) overlap the large kernels; lafis_set_streams(ctx, 1) makes
                                 the intervals exclusive. */
    /* cumulative exactness bookkeeping since the context was created */
    uint64_t minu_replays;    /* top-120 selections that needed the introsort replay (ties) */
    uint64_t tex_replays;     /* top-200 row selections that needed the introsort replay */
    uint64_t tex_queued;      /* texture row-max: (row, column) candidates the 8-bit integer filter let through */
    uint64_t tex_exact;       /* ... of which re-evaluated exactly in fp32 */
    uint64_t tex_overflow;    /* ... (warp, template) visits whose queue overflowed: evaluated exactly in full */
    uint64_t tex_templates;   /* ... (warp, template) visits in total */
    uint64_t minu_big_jobs;   /* (latent, template, minutiae slot) jobs too large for the shared-memory tiles, scored
                                 by the HBM-resident kernels (same results, slower) */
    uint64_t graph_minu_dense_jobs; /* minutiae pruning jobs whose consistency graph did not fit the sparse kernel
                                       (mated or near-duplicate prints): dense kernel, same result, ~100 us each */
    uint64_t graph_tex_dense_jobs;  /* the same for the texture component */
    uint64_t graph_minu_mid_jobs;   /* minutiae pruning jobs beyond the first sparse kernel's 2,560 non-zeros, retried with
                                       6,656 (clustered minutiae); graph_minu_dense_jobs counts what that left over */
} lafis_stats;
LAFIS_API int lafis_get_stats(const lafis_ctx* ctx, lafis_stats* out);
/* 2 (default): the rare-path kernels (introsort replays of the selection, dense texture graphs: a few hundred long jobs)
 * run on a second, high-priority CUDA stream next to the following large kernel of the main stream instead of holding it
 * up; 1: every kernel on one stream - per-kernel times in lafis_stats are then exclusive.  (Running the whole texture
 * chain beside the minutiae chain was measured and does not pay; LAFIS_FORCE_TWO_STREAMS=1 re-enables it for experiments.) */
LAFIS_API int lafis_set_streams(lafis_ctx* ctx, int n_streams);
LAFIS_API void* lafis_stream(const lafis_ctx* ctx); /* the cudaStream_t all work is enqueued on */

#ifdef __cplusplus
}
#endif
#endif /* LATENTAFIS_B200_H */
