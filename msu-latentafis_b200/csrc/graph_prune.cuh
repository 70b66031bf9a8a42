// Second-order graph pruning of candidate correspondences and the component score —
// stages K3(second half), K4 / K8, K9, K9b.
//
//   reference: top-N rows by similarity             matching/matcher.cpp:736-749   (K3b, texture)
//              LSS_R_Fast2_Dist_lookup              matching/matcher.cpp:1225-1348 (K4, texture)
//              LSS_R_Fast2_Dist_eigen               matching/matcher.cpp:1350-1469 (K8, minutiae)
//              LSS_R_Fast2 + adjust_angle           matching/matcher.cpp:1471-1647 (K9, both)
//              score = sum of surviving similarities :508-514, :775-781            (K9b)
//
// DENSE variant (fallback for jobs the sparse fast path of graph_sparse.cuh cannot hold: mated pairs,
// whose consistency graphs are dense).  One CTA per job.  All intermediate state lives in shared
// memory: the pairwise compatibility matrix H (fp32 for the distance graph, bytes for the angle
// graph), the power-iteration vectors and the candidate lists.  Every floating-point step uses the
// reference's operation order with unfused fp32 / genuine fp64 where the reference's literals force
// double evaluation, so that the discrete decisions (thresholds, sort order, greedy selection) come
// out identical:
//   * H rows are accumulated k ascending (Eigen stand-in order), sums are sequential, the
//     normalisation factor is (float)(1.0 / ((double)sum + 1e-5));
//   * candidates are ordered with a parallel rank sort on the total order (value desc, index asc);
//     std::sort is only stable for n <= 16, so when n > 16 and two candidates above the stop
//     threshold are equal, thread 0 replays libstdc++'s introsort (stdsort_emul.h);
//   * the greedy pass of the reference (accept a candidate iff its latent and rolled minutiae are
//     unused and it is compatible with everything accepted so far) is evaluated as "repeatedly take
//     the first unblocked candidate, then block everything that conflicts with it", which visits
//     and accepts exactly the same candidates in the same order;
//   * line angles use atan2f_fdlibm (exact_math.h), bit-identical to the libm call at :1516.
#pragma once
#include "device_common.cuh"
#include "stdsort_emul.h"

namespace lafis {

#define LAFIS_PI_D 3.1415926 /* matching/matcher.h:26: a double literal */

__device__ __forceinline__ float adjust_angle_ref(float angle) {  // matcher.cpp:1638-1647
    if ((double)angle > LAFIS_PI_D) angle = (float)((double)angle - 2 * LAFIS_PI_D);
    else if ((double)angle < -LAFIS_PI_D) angle = (float)((double)angle + 2 * LAFIS_PI_D);
    return angle;
}
__device__ __forceinline__ float angle_gap_ref(float a1, float a2) {  // matcher.cpp:1501-1504
    float d = fabsf(f_sub(a1, a2));
    if ((double)d > LAFIS_PI_D) d = (float)(2 * LAFIS_PI_D - (double)d);
    return d;
}

// Work arrays of one pruning job, all in shared memory.
template <int MAXN>
struct GraphWork {
    float v[MAXN];            // similarity of each candidate
    int li[MAXN], rj[MAXN];   // latent / rolled minutia index of each candidate
    int lx[MAXN], ly[MAXN], rx[MAXN], ry[MAXN];
    float lo[MAXN], ro[MAXN];
    float b[MAXN], c[MAXN];   // power iteration
    int y[MAXN];              // sorted order
    int sel[MAXN];            // accepted candidates
    float v2[MAXN];
    int li2[MAXN], rj2[MAXN];
    int lx2[MAXN], ly2[MAXN], rx2[MAXN], ry2[MAXN];
    float lo2[MAXN], ro2[MAXN];
    float f;
    int next, flag, nsel;
};

// sort candidates [0,num) by key descending into w.y with std::sort's permutation
template <int MAXN, int NT>
__device__ __forceinline__ void sort_candidates(GraphWork<MAXN>& w, const float* key, int num, double stop_thr) {
    const int tid = threadIdx.x;
    if (tid == 0) w.flag = 0;
    __syncthreads();
    if (tid < num) {
        const float mk = key[tid];
        int rank = 0;
        bool tie = false;
        for (int k = 0; k < num; ++k) {
            const float ok = key[k];
            rank += (ok > mk) || (ok == mk && k < tid);
            tie |= (ok == mk && k != tid);
        }
        w.y[rank] = tid;
        if (tie && num > 16 && !((double)mk < stop_thr)) w.flag = 1;
    }
    __syncthreads();
    if (w.flag) {
        if (tid == 0) std_sort_desc_emulate<float, int>(key, w.y, num);
        __syncthreads();
    }
}

// Greedy selection over the sorted candidates.  compat(a, b) is the pairwise predicate.
// On return w.sel[0..w.nsel) holds the accepted candidate indices in acceptance order.
template <int MAXN, int NT, typename Compat>
__device__ __forceinline__ void greedy_select(GraphWork<MAXN>& w, const float* key, int num, double stop_thr,
                                              const int* cli, const int* crj, Compat compat) {
    const int p = threadIdx.x;
    int ind = 0;
    bool open = false;
    if (p < num) {
        ind = w.y[p];
        open = !((double)key[ind] < stop_thr);  // sorted, so the open positions form a prefix
    }
    int nsel = 0;
    for (;;) {
        if (p == 0) w.next = 0x7fffffff;
        __syncthreads();
        if (open) atomicMin(&w.next, p);
        __syncthreads();
        const int c = w.next;
        if (c == 0x7fffffff) break;
        const int s = w.y[c];
        if (p == c) {
            w.sel[nsel] = s;
            open = false;
        } else if (open) {
            if (cli[ind] == cli[s] || crj[ind] == crj[s] || !compat(s, ind)) open = false;
        }
        ++nsel;
        __syncthreads();
    }
    if (p == 0) w.nsel = nsel;
    __syncthreads();
}

// The whole pruning cascade for one candidate list held in w (v, li, rj, coordinates, orientations).
// LOOKUP selects the texture flavour (table distances, 3 iterations) or the minutiae flavour
// (Euclidean distances, 5 iterations).  Returns the component score (valid in thread 0).
// *n_final (thread 0) = number of surviving correspondences; they are w.sel[0..n) into the *2 arrays, in the
// order the reference pushes them into corr3 (and writes them to its correspondence file, matcher.cpp:497-505).
template <int MAXN, int NT, bool LOOKUP>
__device__ float prune_cascade(GraphWork<MAXN>& w, int num, float* H, const float* table, int* n_final = nullptr) {
    const int tid = threadIdx.x;
    constexpr int LD = MAXN;
    if (n_final) *n_final = 0;
    if (num <= 0) return 0.0f;

    // ---- distance-consistency graph ----
    for (int e = tid; e < num * LD; e += NT) H[e] = 0.0f;
    __syncthreads();
    for (int e = tid; e < num * num; e += NT) {
        const int i = e / num, j = e - i * num;
        if (i >= j) continue;
        float d1, d2;
        if (LOOKUP) {  // matcher.cpp:1246-1262
            const int dx1 = abs(w.lx[i] - w.lx[j]), dx2 = abs(w.rx[i] - w.rx[j]);
            const int dy1 = abs(w.ly[i] - w.ly[j]), dy2 = abs(w.ry[i] - w.ry[j]);
            if (dx1 >= kTableN || dx2 >= kTableN || dy1 >= kTableN || dy2 >= kTableN) continue;
            d1 = table[dx1 * kTableN + dy1];
            d2 = table[dx2 * kTableN + dy2];
        } else {  // matcher.cpp:1372-1384
            const float dx1 = (float)(w.lx[i] - w.lx[j]), dx2 = (float)(w.rx[i] - w.rx[j]);
            const float dy1 = (float)(w.ly[i] - w.ly[j]), dy2 = (float)(w.ry[i] - w.ry[j]);
            d1 = __fsqrt_rn(f_add(f_mul(dx1, dx1), f_mul(dy1, dy1)));
            d2 = __fsqrt_rn(f_add(f_mul(dx2, dx2), f_mul(dy2, dy2)));
        }
        const float dist = fabsf(f_sub(d1, d2));
        if (dist > 30.0f) continue;
        // (30 - dist) / 25.0 is a double division narrowed to float; for a float numerator and the
        // exactly representable 25 that equals the correctly rounded float division
        float h = f_div(f_sub(30.0f, dist), 25.0f);
        if (h > 1.0f) h = 1.0f;
        else if (h < 0.0f) h = 0.0f;
        H[i * LD + j] = h;
        H[j * LD + i] = h;
    }
    if (tid < num) w.b[tid] = w.v[tid];
    __syncthreads();
    constexpr int ITERS = LOOKUP ? 3 : 5;  // matcher.cpp:1284 / :1406
    for (int it = 0; it < ITERS; ++it) {
        if (tid < num) {
            float acc = 0.0f;
            for (int k = 0; k < num; ++k) acc = f_add(acc, f_mul(H[k * LD + tid], w.b[k]));
            w.c[tid] = acc;
        }
        __syncthreads();
        if (tid == 0) {
            float sum = w.c[0];
            for (int i = 1; i < num; ++i) sum = f_add(sum, w.c[i]);
            w.f = (float)(1.0 / ((double)sum + 0.00001));
        }
        __syncthreads();
        if (tid < num) w.b[tid] = f_mul(w.c[tid], w.f);
        __syncthreads();
    }
    sort_candidates<MAXN, NT>(w, w.b, num, 0.0001);
    greedy_select<MAXN, NT>(w, w.b, num, 0.0001, w.li, w.rj,
                            [&](int a, int bb) { return !((double)H[a * LD + bb] < 0.00001); });
    const int n2 = w.nsel;
    if (n2 <= 0) return 0.0f;
    if (tid < n2) {
        const int s = w.sel[tid];
        w.v2[tid] = w.v[s];
        w.li2[tid] = w.li[s];
        w.rj2[tid] = w.rj[s];
        w.lx2[tid] = w.lx[s];
        w.ly2[tid] = w.ly[s];
        w.rx2[tid] = w.rx[s];
        w.ry2[tid] = w.ry[s];
        w.lo2[tid] = w.lo[s];
        w.ro2[tid] = w.ro[s];
    }
    __syncthreads();

    // ---- orientation-consistency graph, matcher.cpp:1486-1554 ----
    unsigned char* Hb = reinterpret_cast<unsigned char*>(H);
    for (int e = tid; e < n2 * n2; e += NT) Hb[e] = 0;
    __syncthreads();
    for (int e = tid; e < n2 * n2; e += NT) {
        const int i = e / n2, j = e - i * n2;
        if (i >= j) continue;
        float a1 = adjust_angle_ref(f_sub(w.lo2[i], w.lo2[j]));
        float a2 = adjust_angle_ref(f_sub(w.ro2[i], w.ro2[j]));
        if ((double)angle_gap_ref(a1, a2) > LAFIS_PI_D / 4.) continue;
        const float dx1 = (float)(w.lx2[i] - w.lx2[j]), dy1 = (float)(w.ly2[i] - w.ly2[j]);
        const float line1 = -atan2f_fdlibm(dy1, dx1);
        const float dx2 = (float)(w.rx2[i] - w.rx2[j]), dy2 = (float)(w.ry2[i] - w.ry2[j]);
        const float line2 = -atan2f_fdlibm(dy2, dx2);
        a1 = adjust_angle_ref(f_sub(w.lo2[i], line1));
        a2 = adjust_angle_ref(f_sub(w.ro2[i], line2));
        if ((double)angle_gap_ref(a1, a2) > LAFIS_PI_D / 6.) continue;
        a1 = adjust_angle_ref(f_sub(w.lo2[j], line1));
        a2 = adjust_angle_ref(f_sub(w.ro2[j], line2));
        if ((double)angle_gap_ref(a1, a2) > LAFIS_PI_D / 6.) continue;
        Hb[i * n2 + j] = 1;
        Hb[j * n2 + i] = 1;
    }
    if (tid < n2) w.b[tid] = (float)(1.0 / (double)n2);  // matcher.cpp:1558
    __syncthreads();
    for (int it = 0; it < 5; ++it) {  // matcher.cpp:1563-1581
        if (tid < n2) {
            float acc = 0.0f;
            for (int k = 0; k < n2; ++k)
                if (Hb[k * n2 + tid]) acc = f_add(acc, w.b[k]);
            w.c[tid] = acc;
        }
        __syncthreads();
        if (tid == 0) {
            float sum = 0.0f;
            for (int j = 0; j < n2; ++j) sum = f_add(sum, w.c[j]);
            w.f = (float)(1.0 / ((double)sum + 0.00001));
        }
        __syncthreads();
        if (tid < n2) w.b[tid] = f_mul(w.c[tid], w.f);
        __syncthreads();
    }
    sort_candidates<MAXN, NT>(w, w.b, n2, 0.001);
    greedy_select<MAXN, NT>(w, w.b, n2, 0.001, w.li2, w.rj2,
                            [&](int a, int bb) { return Hb[a * n2 + bb] != 0; });
    float score = 0.0f;
    if (tid == 0)
        for (int s = 0; s < w.nsel; ++s) score = f_add(score, w.v2[w.sel[s]]);
    if (n_final) *n_final = w.nsel;
    return score;
}

// ------------------------------------------------------------------------------------------------
// minutiae components: grid = Q * n_chunk * 3, block = 128
// ------------------------------------------------------------------------------------------------
struct GraphMinuParams {
    const float* corr_v;
    const uint32_t* corr_ij;
    const int* corr_n;
    // latent minutiae
    const uint32_t* slot_off;
    const short2* lat_xy;
    const float* lat_ori;
    // gallery minutiae
    const uint32_t* minu_off;
    const short2* gal_xy;
    const float* gal_ori;
    int g0, n_chunk;
    int Q;        // latents of the batch
    int G;        // templates resident on this device
    float* comp;  // [Q][G][4] = score[0], score[1], score[2], score[28]
    // optional (dense kernel only): the surviving correspondences of every job, [job][kTopCorrMinu] x
    // {latent x, latent y, rolled x, rolled y} and their count - the reference's save_corr output
    short4* corr_out = nullptr;
    int* corr_out_n = nullptr;
    unsigned long long* dense_jobs_total = nullptr;  // statistics: jobs the dense kernel took over
};

constexpr int kGraphMinuThreads = 128;
constexpr size_t kGraphMinuSmem = sizeof(float) * kTopCorrMinu * kTopCorrMinu + sizeof(GraphWork<kTopCorrMinu>);

// Dense fallback: processes the jobs the sparse kernel (graph_sparse.cuh) could not hold.
__global__ void __launch_bounds__(kGraphMinuThreads) graph_minu_dense_kernel(GraphMinuParams P, const int* job_count,
                                                                             const int* jobs) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* H = reinterpret_cast<float*>(smem);
    GraphWork<kTopCorrMinu>& w = *reinterpret_cast<GraphWork<kTopCorrMinu>*>(smem + sizeof(float) * kTopCorrMinu * kTopCorrMinu);
    const int tid = threadIdx.x;
    const int n_jobs = *job_count;
    if (blockIdx.x == 0 && tid == 0 && P.dense_jobs_total && !P.corr_out) atomicAdd(P.dense_jobs_total, (unsigned long long)n_jobs);
    for (int jb = blockIdx.x; jb < n_jobs; jb += gridDim.x) {
        const size_t oidx = (size_t)jobs[jb];  // (q * n_chunk + tl) * 3 + slot
        const int slot = (int)(oidx % 3);
        const size_t pair = oidx / 3;
        const int tl = (int)(pair % P.n_chunk), q = (int)(pair / P.n_chunk);
        const int num = P.corr_n[oidx];
        __syncthreads();  // previous job's shared state no longer in use
        if (num > 0 && tid < num) {
            const uint32_t ij = P.corr_ij[oidx * kTopCorrMinu + tid];
            const int i = (int)(ij >> 16), j = (int)(ij & 0xffffu);
            w.v[tid] = P.corr_v[oidx * kTopCorrMinu + tid];
            w.li[tid] = i;
            w.rj[tid] = j;
            const uint32_t lo = P.slot_off[q * 3 + slot] + i, go = P.minu_off[P.g0 + tl] + j;
            const short2 a = P.lat_xy[lo], b = P.gal_xy[go];
            w.lx[tid] = a.x;
            w.ly[tid] = a.y;
            w.rx[tid] = b.x;
            w.ry[tid] = b.y;
            w.lo[tid] = P.lat_ori[lo];
            w.ro[tid] = P.gal_ori[go];
        }
        __syncthreads();
        int n_final = 0;
        const float score = prune_cascade<kTopCorrMinu, kGraphMinuThreads, false>(w, num, H, nullptr, &n_final);
        if (tid == 0) P.comp[((size_t)q * P.G + P.g0 + tl) * 4 + slot] = score;
        if (P.corr_out) {  // n_final is uniform across the block
            if (tid == 0) P.corr_out_n[oidx] = n_final;
            if (tid < n_final) {
                const int s = w.sel[tid];
                P.corr_out[oidx * kTopCorrMinu + tid] =
                    make_short4((short)w.lx2[s], (short)w.ly2[s], (short)w.rx2[s], (short)w.ry2[s]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// texture component: grid = Q * n_chunk, block = 256
// ------------------------------------------------------------------------------------------------
struct GraphTexParams {
    const float* rowmax_val;   // [Q][n_chunk][lt_stride]
    const uint16_t* rowmax_j;
    int lt_stride;
    const int* lat_nt;         // [Q]
    const int* lat_status;     // [Q]
    const short2* lat_xy;      // [Q][lt_stride]
    const float* lat_ori;
    const uint32_t* tex_off;
    const short2* gal_xy;
    const float* gal_ori;
    const float* table;        // [2500]
    int g0, n_chunk;
    int Q;
    int G;
    float* comp;               // [Q][G][4]; slot 3
    unsigned long long* slow_path_count;
    unsigned long long* dense_jobs_total = nullptr;  // statistics: jobs the dense kernel took over
};

constexpr int kGraphTexThreads = 256;
constexpr int kMaxTexPts = 1000;
struct TexRowWork {
    float rv[kMaxTexPts];
    int ry[kMaxTexPts];
    unsigned short rj[kMaxTexPts];
    float table[kTableN * kTableN];
    int flag;
};
constexpr size_t kGraphTexSmem =
    sizeof(float) * kTopCorrTex * kTopCorrTex + sizeof(GraphWork<kTopCorrTex>) + sizeof(TexRowWork);

__global__ void __launch_bounds__(kGraphTexThreads) graph_tex_dense_kernel(GraphTexParams P, const int* job_count,
                                                                           const int* jobs) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* H = reinterpret_cast<float*>(smem);
    GraphWork<kTopCorrTex>& w = *reinterpret_cast<GraphWork<kTopCorrTex>*>(smem + sizeof(float) * kTopCorrTex * kTopCorrTex);
    TexRowWork& r = *reinterpret_cast<TexRowWork*>(smem + sizeof(float) * kTopCorrTex * kTopCorrTex + sizeof(GraphWork<kTopCorrTex>));
    const int tid = threadIdx.x;
    const int n_jobs = *job_count;
    if (blockIdx.x == 0 && tid == 0 && P.dense_jobs_total) atomicAdd(P.dense_jobs_total, (unsigned long long)n_jobs);
    for (int e = tid; e < kTableN * kTableN; e += kGraphTexThreads) r.table[e] = P.table[e];
    for (int jb = blockIdx.x; jb < n_jobs; jb += gridDim.x) {
        const size_t pair = (size_t)jobs[jb];
        const int tl = (int)(pair % P.n_chunk), q = (int)(pair / P.n_chunk);
        const int g = P.g0 + tl;
        const int nLt = (P.lat_status[q] == 0) ? P.lat_nt[q] : 0;
        const uint32_t gbase = P.tex_off[g];
        const int nRt = (int)(P.tex_off[g + 1] - gbase);
        __syncthreads();  // previous job's shared state no longer in use
        if (nLt <= 0 || nRt <= 0) {  // matcher.cpp:411: no texture template on one side, score stays 0
            if (tid == 0) P.comp[((size_t)q * P.G + g) * 4 + 3] = 0.0f;
            continue;
        }
        const size_t rbase = pair * (size_t)P.lt_stride;
        for (int i = tid; i < nLt; i += kGraphTexThreads) {
            r.rv[i] = P.rowmax_val[rbase + i];
            r.rj[i] = P.rowmax_j[rbase + i];
        }
        if (tid == 0) r.flag = 0;
        __syncthreads();

        // ---- K3b: the N best rows in std::sort order (matcher.cpp:736-749) ----
        int num;
        if (nLt > kTopCorrTex) {
            for (int i = tid; i < nLt; i += kGraphTexThreads) {
                const float mk = r.rv[i];
                int rank = 0;
                bool tie = false;
                for (int k = 0; k < nLt; ++k) {
                    const float ok = r.rv[k];
                    rank += (ok > mk) || (ok == mk && k < i);
                    tie |= (ok == mk && k != i);
                }
                r.ry[rank] = i;
                // a tie group matters when it reaches into the first N positions
                if (tie && rank <= kTopCorrTex) r.flag = 1;
            }
            __syncthreads();
            if (r.flag) {
                if (tid == 0) {
                    std_sort_desc_prefix(DenseKey<float>{r.rv}, r.ry, nLt, kTopCorrTex);
                    atomicAdd(P.slow_path_count, 1ull);
                }
                __syncthreads();
            }
            num = kTopCorrTex;
        } else {
            for (int i = tid; i < nLt; i += kGraphTexThreads) r.ry[i] = i;
            __syncthreads();
            num = nLt;
        }
        if (tid < num) {
            const int i = r.ry[tid];
            const int j = r.rj[i];
            w.v[tid] = r.rv[i];
            w.li[tid] = i;
            w.rj[tid] = j;
            const short2 a = P.lat_xy[(size_t)q * P.lt_stride + i], b = P.gal_xy[gbase + j];
            w.lx[tid] = a.x;
            w.ly[tid] = a.y;
            w.rx[tid] = b.x;
            w.ry[tid] = b.y;
            w.lo[tid] = P.lat_ori[(size_t)q * P.lt_stride + i];
            w.ro[tid] = P.gal_ori[gbase + j];
        }
        __syncthreads();
        const float score = prune_cascade<kTopCorrTex, kGraphTexThreads, true>(w, num, H, r.table);
        if (tid == 0) P.comp[((size_t)q * P.G + g) * 4 + 3] = score;
    }
}

}  // namespace lafis
