/*
 * CPU oracle for the MSU-LatentAFIS 1-vs-N matcher hot path — TEST INFRASTRUCTURE ONLY.
 *
 * This file is a plain-C restatement of the algorithm in the reference's matching/matcher.cpp and
 * matching/include.h.  It exists so that tests can compare the CUDA path against an independent CPU
 * implementation, stage by stage.  It must never be linked, imported or called by the product
 * (msu-latentafis_b200/): only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may use it.
 *
 * Parity status: PINNED against the reference itself.  The reference ships no golden vectors, so the
 * pin is oracle/_ref/libref_matcher.so (the reference's own matcher.cpp compiled by
 * oracle/build_ref.sh); tests/test_oracle_vs_reference.py requires bit-identical component and
 * fused scores on synthetic templates, and tests/golden/ holds vectors produced by that library.
 * At the Eigen boundary (GEMM / GEMV / reductions) the summation order is the one defined in
 * oracle/shim/Eigen/Dense, because the reference pins no Eigen version.
 *
 * Every function cites the reference lines it follows (paths relative to the reference root).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LO_PI 3.1415926 /* matching/include.h:22, matching/matcher.h:26 — a double literal */
#define LO_TABLE_N 50   /* matching/matcher.cpp:45 dist_N */
#define LO_MAX_TEX 1000 /* matching/matcher.h:31-32 MaxNRolledMinu / MaxNLatentMinu */

/* element count for malloc/calloc: never zero, never negative */
static size_t lo_cnt(long n) { return n > 0 ? (size_t)n : (size_t)1; }

typedef struct {
    int n;
    const short *x, *y;
    const float* ori;
    int des_len;
    const float* des;           /* float descriptors [n][des_len] (minutiae, latent texture) */
    const unsigned char* codes; /* PQ codes [n][des_len] (rolled texture) */
} lo_template_t;

/* ------------------------------------------------------------------------------------------------
 * std::sort permutation (libstdc++ 13, bits/stl_algo.h __sort / __introsort_loop /
 * __final_insertion_sort, bits/stl_heap.h).  The reference sorts index vectors with an UNSTABLE
 * std::sort at matcher.cpp:476, :741, :1301, :1423, :1590; LSS_R_Fast2 produces structural ties, so
 * the exact permutation is part of the result.  `idx` holds 0..n-1 on entry; comparator is
 * key[a] > key[b] (descending), exactly the lambdas at those call sites.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const float* key;
} lo_cmp_t;

static inline int lo_before(const lo_cmp_t* c, int a, int b) { return c->key[a] > c->key[b]; }
static inline void lo_swap(int* a, int* b) {
    int t = *a;
    *a = *b;
    *b = t;
}

static void lo_unguarded_linear_insert(int* last, const lo_cmp_t* c) {
    int val = *last;
    int* next = last - 1;
    while (lo_before(c, val, *next)) {
        *last = *next;
        last = next;
        --next;
    }
    *last = val;
}

static void lo_insertion_sort(int* first, int* last, const lo_cmp_t* c) {
    if (first == last) return;
    for (int* i = first + 1; i != last; ++i) {
        if (lo_before(c, *i, *first)) {
            int val = *i;
            memmove(first + 1, first, (size_t)(i - first) * sizeof(int));
            *first = val;
        } else {
            lo_unguarded_linear_insert(i, c);
        }
    }
}

static void lo_push_heap(int* first, long hole, long top, int value, const lo_cmp_t* c) {
    long parent = (hole - 1) / 2;
    while (hole > top && lo_before(c, first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

static void lo_adjust_heap(int* first, long hole, long len, int value, const lo_cmp_t* c) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (lo_before(c, first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    lo_push_heap(first, hole, top, value, c);
}

static void lo_heapsort(int* first, int* last, const lo_cmp_t* c) {
    /* __partial_sort(first, last, last): __heap_select degenerates to __make_heap, then __sort_heap */
    long len = last - first;
    if (len >= 2) {
        long parent = (len - 2) / 2;
        for (;;) {
            int value = first[parent];
            lo_adjust_heap(first, parent, len, value, c);
            if (parent == 0) break;
            parent--;
        }
    }
    while (last - first > 1) {
        --last;
        int value = *last;
        *last = *first;
        lo_adjust_heap(first, 0, last - first, value, c);
    }
}

static void lo_introsort_loop(int* first, int* last, long depth_limit, const lo_cmp_t* c) {
    while (last - first > 16) {
        if (depth_limit == 0) {
            lo_heapsort(first, last, c);
            return;
        }
        --depth_limit;
        /* __unguarded_partition_pivot */
        int* mid = first + (last - first) / 2;
        int *a = first + 1, *b = mid, *cc = last - 1;
        if (lo_before(c, *a, *b)) {
            if (lo_before(c, *b, *cc)) lo_swap(first, b);
            else if (lo_before(c, *a, *cc)) lo_swap(first, cc);
            else lo_swap(first, a);
        } else if (lo_before(c, *a, *cc)) lo_swap(first, a);
        else if (lo_before(c, *b, *cc)) lo_swap(first, cc);
        else lo_swap(first, b);
        int *lo = first + 1, *hi = last;
        for (;;) {
            while (lo_before(c, *lo, *first)) ++lo;
            --hi;
            while (lo_before(c, *first, *hi)) --hi;
            if (!(lo < hi)) break;
            lo_swap(lo, hi);
            ++lo;
        }
        lo_introsort_loop(lo, last, depth_limit, c);
        last = lo;
    }
}

void lo_std_sort_desc(const float* key, int* idx, int n) {
    for (int i = 0; i < n; ++i) idx[i] = i; /* std::iota */
    if (n == 0) return;
    lo_cmp_t c = {key};
    long lg = 0;
    for (long m = n; m > 1; m >>= 1) ++lg; /* std::__lg */
    lo_introsort_loop(idx, idx + n, lg * 2, &c);
    if (n > 16) {
        lo_insertion_sort(idx, idx + 16, &c);
        for (int* i = idx + 16; i != idx + n; ++i) lo_unguarded_linear_insert(i, &c);
    } else {
        lo_insertion_sort(idx, idx + n, &c);
    }
}

/* ------------------------------------------------------------------------------------------------
 * F0: distance table, matcher.cpp:45-56.  table[i*50+j] = (float)sqrt((16 i)^2 + (16 j)^2), the
 * square root taken in double.
 * ---------------------------------------------------------------------------------------------- */
void lo_make_table(float* table) {
    for (int i = 0; i < LO_TABLE_N; ++i)
        for (int j = 0; j < LO_TABLE_N; ++j)
            table[i * LO_TABLE_N + j] = (float)sqrt((i * 16.0) * (i * 16.0) + (j * 16.0) * (j * 16.0));
}

/* ------------------------------------------------------------------------------------------------
 * K1: PQ distance look-up table of a latent texture template, include.h:327-359.
 * lut[i][m][q] = sum_{k<sub_dim} (des[i][m*sub_dim+k] - cw[m][q][k])^2, k ascending, fp32, separate
 * multiply and add.
 * ---------------------------------------------------------------------------------------------- */
void lo_build_lut(const float* des, int n, int des_len, const float* cw, int subs, int clusters, int sub_dim,
                  float* lut) {
    for (int i = 0; i < n; ++i)
        for (int m = 0; m < subs; ++m)
            for (int q = 0; q < clusters; ++q) {
                const float* d = des + (size_t)i * des_len + m * sub_dim;
                const float* w = cw + ((size_t)m * clusters + q) * sub_dim;
                float dist = 0.0f;
                for (int k = 0; k < sub_dim; ++k) {
                    float t = d[k] - w[k];
                    float t2 = t * t;
                    dist = dist + t2;
                }
                lut[((size_t)i * subs + m) * clusters + q] = dist;
            }
}

/* ------------------------------------------------------------------------------------------------
 * K2: texture similarity by LUT gather, matcher.cpp:566-594 (method 1).  Four running values, the
 * first starting at 6, sub-quantizer m feeding value (m mod 4); result (d1+d2)+(d3+d4).
 * ---------------------------------------------------------------------------------------------- */
void lo_texture_similarity(const float* lut, int nL, const unsigned char* codes, int nR, int subs, int clusters,
                           int code_stride, float* sim) {
    size_t n = 0;
    for (int i = 0; i < nL; ++i) {
        const float* row = lut + (size_t)i * subs * clusters;
        for (int j = 0; j < nR; ++j) {
            const unsigned char* c = codes + (size_t)j * code_stride;
            float d[4] = {6.0f, 0.0f, 0.0f, 0.0f};
            for (int m = 0; m < subs; m += 4)
                for (int u = 0; u < 4; ++u) d[u] -= row[(size_t)(m + u) * clusters + c[m + u]];
            sim[n++] = (d[0] + d[1]) + (d[2] + d[3]);
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * K3: initial texture correspondences, matcher.cpp:723-749.  Per latent row the FIRST maximum
 * (std::max_element); if more than N rows, the N best rows in std::sort order.
 * ---------------------------------------------------------------------------------------------- */
int lo_texture_initial_corr(const float* sim, int nL, int nR, int N, float* v, int* li, int* rj) {
    float* rmax = (float*)calloc(lo_cnt(nL), sizeof(float));
    int* rarg = (int*)calloc(lo_cnt(nL), sizeof(int));
    for (int i = 0; i < nL; ++i) {
        const float* p = sim + (size_t)i * nR;
        int best = 0;
        for (int j = 1; j < nR; ++j)
            if (p[best] < p[j]) best = j; /* max_element keeps the first of equal maxima */
        rmax[i] = nR > 0 ? p[best] : 0.0f;
        rarg[i] = best;
    }
    int num;
    if (nL > N) {
        int* y = (int*)malloc(sizeof(int) * (size_t)nL);
        lo_std_sort_desc(rmax, y, nL);
        for (int i = 0; i < N; ++i) {
            v[i] = rmax[y[i]];
            li[i] = y[i];
            rj[i] = rarg[y[i]];
        }
        free(y);
        num = N;
    } else {
        for (int i = 0; i < nL; ++i) {
            v[i] = rmax[i];
            li[i] = i;
            rj[i] = rarg[i];
        }
        num = nL;
    }
    free(rmax);
    free(rarg);
    return num;
}

/* ------------------------------------------------------------------------------------------------
 * Shared tail of the two distance-graph routines: power iteration, sort, greedy selection.
 * matcher.cpp:1277-1347 (lookup, 3 iterations) and :1399-1468 (eigen, 5 iterations).
 * ---------------------------------------------------------------------------------------------- */
static int lo_dist_graph_select(const float* H, int num, int iters, const float* v, const int* li, const int* rj,
                                int nL, int nR, float* ov, int* oli, int* orj) {
    if (num <= 0) return 0;
    float* b = (float*)malloc(sizeof(float) * (size_t)num);
    float* c = (float*)malloc(sizeof(float) * (size_t)num);
    for (int i = 0; i < num; ++i) b[i] = v[i];
    for (int it = 0; it < iters; ++it) {
        for (int i = 0; i < num; ++i) { /* c = H * b (oracle/shim/Eigen/Dense: k ascending) */
            float acc = 0.0f;
            for (int k = 0; k < num; ++k) {
                float p = H[(size_t)i * num + k] * b[k];
                acc = acc + p;
            }
            c[i] = acc;
        }
        float sum = c[0];
        for (int i = 1; i < num; ++i) sum += c[i];
        float f = (float)(1. / (sum + 0.00001)); /* scalar converted to float before the product */
        for (int i = 0; i < num; ++i) b[i] = c[i] * f;
    }
    int* y = (int*)malloc(sizeof(int) * (size_t)num);
    lo_std_sort_desc(b, y, num);
    short* fl = (short*)calloc((size_t)(nL > 0 ? nL : 1), sizeof(short));
    short* fr = (short*)calloc((size_t)(nR > 0 ? nR : 1), sizeof(short));
    int* sel = (int*)malloc(sizeof(int) * (size_t)num);
    int nsel = 0;
    for (int i = 0; i < num; ++i) {
        int ind = y[i];
        if (b[ind] < 0.0001) break;
        if (fl[li[ind]] == 1 || fr[rj[ind]] == 1) continue;
        int ok = 1;
        if (i != 0)
            for (int s = 0; s < nsel; ++s)
                if (H[(size_t)ind * num + sel[s]] < 0.00001) {
                    ok = 0;
                    break;
                }
        if (ok) {
            sel[nsel] = ind;
            ov[nsel] = v[ind];
            oli[nsel] = li[ind];
            orj[nsel] = rj[ind];
            nsel++;
            fl[li[ind]] = 1;
            fr[rj[ind]] = 1;
        }
    }
    free(b); free(c); free(y); free(fl); free(fr); free(sel);
    return nsel;
}

/* K4: LSS_R_Fast2_Dist_lookup, matcher.cpp:1225-1348 (block coordinates, table look-up). */
int lo_prune_dist_lookup(const float* v, const int* li, const int* rj, int num, const lo_template_t* L,
                         const lo_template_t* R, const float* table, float* ov, int* oli, int* orj) {
    float* H = (float*)calloc((size_t)(num > 0 ? num : 1) * (size_t)(num > 0 ? num : 1), sizeof(float));
    for (int i = 0; i < num - 1; ++i)
        for (int j = i + 1; j < num; ++j) {
            int dx1 = abs((int)L->x[li[i]] - (int)L->x[li[j]]);
            int dx2 = abs((int)R->x[rj[i]] - (int)R->x[rj[j]]);
            int dy1 = abs((int)L->y[li[i]] - (int)L->y[li[j]]);
            int dy2 = abs((int)R->y[rj[i]] - (int)R->y[rj[j]]);
            if (dx1 >= LO_TABLE_N || dx2 >= LO_TABLE_N || dy1 >= LO_TABLE_N || dy2 >= LO_TABLE_N) continue;
            float d1 = table[dx1 * LO_TABLE_N + dy1];
            float d2 = table[dx2 * LO_TABLE_N + dy2];
            float dist = fabsf(d1 - d2);
            if (dist > 30.0f) continue;
            float h = (30 - dist) / (25.0);
            if (h > 1) h = 1.0;
            else if (h < 0) h = 0.0;
            H[(size_t)i * num + j] = h;
            H[(size_t)j * num + i] = h;
        }
    int k = lo_dist_graph_select(H, num, 3, v, li, rj, L->n, R->n, ov, oli, orj);
    free(H);
    return k;
}

/* K8: LSS_R_Fast2_Dist_eigen, matcher.cpp:1350-1469 (pixel coordinates, Euclidean distance). */
int lo_prune_dist_euclid(const float* v, const int* li, const int* rj, int num, const lo_template_t* L,
                         const lo_template_t* R, float* ov, int* oli, int* orj) {
    float* H = (float*)calloc((size_t)(num > 0 ? num : 1) * (size_t)(num > 0 ? num : 1), sizeof(float));
    for (int i = 0; i < num - 1; ++i)
        for (int j = i + 1; j < num; ++j) {
            float dx1 = (float)((int)L->x[li[i]] - (int)L->x[li[j]]);
            float dx2 = (float)((int)R->x[rj[i]] - (int)R->x[rj[j]]);
            float dy1 = (float)((int)L->y[li[i]] - (int)L->y[li[j]]);
            float dy2 = (float)((int)R->y[rj[i]] - (int)R->y[rj[j]]);
            float a1 = dx1 * dx1, b1 = dy1 * dy1;
            float d1 = sqrtf(a1 + b1);
            float a2 = dx2 * dx2, b2 = dy2 * dy2;
            float d2 = sqrtf(a2 + b2);
            float dist = fabsf(d1 - d2);
            if (dist > 30.0f) continue;
            float h = (30 - dist) / (25.0);
            if (h > 1) h = 1.0;
            else if (h < 0) h = 0.0;
            H[(size_t)i * num + j] = h;
            H[(size_t)j * num + i] = h;
        }
    int k = lo_dist_graph_select(H, num, 5, v, li, rj, L->n, R->n, ov, oli, orj);
    free(H);
    return k;
}

/* matcher.cpp:1638-1647.  The comparisons and the +-2*PI happen in double (PI is a double literal),
 * the result is narrowed back to float. */
static float lo_adjust_angle(float angle) {
    if (angle > LO_PI) angle -= 2 * LO_PI;
    else if (angle < -LO_PI) angle += 2 * LO_PI;
    return angle;
}

/* |a1 - a2| folded into [0, PI], matcher.cpp:1501-1504 (and :1528-1531, :1544-1547). */
static float lo_angle_gap(float a1, float a2) {
    float d = fabsf(a1 - a2);
    if (d > LO_PI) d = 2 * LO_PI - d;
    return d;
}

/* K9: LSS_R_Fast2, matcher.cpp:1471-1636 (orientation-consistency graph, boolean). */
int lo_prune_angle(const float* v, const int* li, const int* rj, int num, const lo_template_t* L,
                   const lo_template_t* R, float* ov, int* oli, int* orj) {
    if (num <= 0) return 0;
    unsigned char* H = (unsigned char*)calloc((size_t)num * (size_t)num, 1);
    for (int i = 0; i < num - 1; ++i)
        for (int j = i + 1; j < num; ++j) {
            const int l1 = li[i], l2 = li[j], r1 = rj[i], r2 = rj[j];
            float a1 = lo_adjust_angle(L->ori[l1] - L->ori[l2]);
            float a2 = lo_adjust_angle(R->ori[r1] - R->ori[r2]);
            if (lo_angle_gap(a1, a2) > LO_PI / 4.) continue;

            float dx1 = (float)((int)L->x[l1] - (int)L->x[l2]);
            float dy1 = (float)((int)L->y[l1] - (int)L->y[l2]);
            float line1 = -atan2f(dy1, dx1);
            float dx2 = (float)((int)R->x[r1] - (int)R->x[r2]);
            float dy2 = (float)((int)R->y[r1] - (int)R->y[r2]);
            float line2 = -atan2f(dy2, dx2);

            a1 = lo_adjust_angle(L->ori[l1] - line1);
            a2 = lo_adjust_angle(R->ori[r1] - line2);
            if (lo_angle_gap(a1, a2) > LO_PI / 6.) continue;

            a1 = lo_adjust_angle(L->ori[l2] - line1);
            a2 = lo_adjust_angle(R->ori[r2] - line2);
            if (lo_angle_gap(a1, a2) > LO_PI / 6.) continue;

            H[(size_t)i * num + j] = 1;
            H[(size_t)j * num + i] = 1;
        }
    float* S = (float*)malloc(sizeof(float) * (size_t)num);
    float* S1 = (float*)malloc(sizeof(float) * (size_t)num);
    float s0 = 1.0 / num;
    for (int i = 0; i < num; ++i) S[i] = s0;
    for (int it = 0; it < 5; ++it) { /* matcher.cpp:1563-1581 */
        float sum = 0.0f;
        for (int j = 0; j < num; ++j) {
            float acc = 0;
            for (int k = 0; k < num; ++k)
                if (H[(size_t)j * num + k]) acc += S[k];
            S1[j] = acc;
            sum += acc;
        }
        sum = 1.0 / (sum + 0.00001);
        for (int j = 0; j < num; ++j) S[j] = S1[j] * sum;
    }
    int* y = (int*)malloc(sizeof(int) * (size_t)num);
    lo_std_sort_desc(S, y, num);
    short* fl = (short*)calloc((size_t)(L->n > 0 ? L->n : 1), sizeof(short));
    short* fr = (short*)calloc((size_t)(R->n > 0 ? R->n : 1), sizeof(short));
    int* sel = (int*)malloc(sizeof(int) * (size_t)num);
    int nsel = 0;
    for (int i = 0; i < num; ++i) { /* matcher.cpp:1596-1633 */
        int ind = y[i];
        if (S[ind] < 0.001) break;
        if (fl[li[ind]] == 1 || fr[rj[ind]] == 1) continue;
        int ok = 1;
        if (i != 0)
            for (int s = 0; s < nsel; ++s)
                if (!H[(size_t)ind * num + sel[s]]) {
                    ok = 0;
                    break;
                }
        if (ok) {
            sel[nsel] = ind;
            ov[nsel] = v[ind];
            oli[nsel] = li[ind];
            orj[nsel] = rj[ind];
            nsel++;
            fl[li[ind]] = 1;
            fr[rj[ind]] = 1;
        }
    }
    free(H); free(S); free(S1); free(y); free(fl); free(fr); free(sel);
    return nsel;
}

/* ------------------------------------------------------------------------------------------------
 * K5-K7: minutiae similarity, normalisation and the top-120 candidates, matcher.cpp:440-488.
 * S_out / nrm_out (each nL*nR, optional) expose the intermediate matrices to the stage tests.
 * ---------------------------------------------------------------------------------------------- */
int lo_minutiae_initial_corr(const float* A, int nL, const float* B, int nR, int des_len, float* v, int* li,
                             int* rj, float* S_out, float* nrm_out) {
    size_t tot = (size_t)nL * (size_t)nR;
    float* S = (float*)malloc(sizeof(float) * lo_cnt((long)tot));
    float* nrm = (float*)malloc(sizeof(float) * lo_cnt((long)tot));
    for (int i = 0; i < nL; ++i)
        for (int j = 0; j < nR; ++j) {
            float acc = 0.0f;
            for (int k = 0; k < des_len; ++k) {
                float p = A[(size_t)i * des_len + k] * B[(size_t)j * des_len + k];
                acc = acc + p;
            }
            if (acc < 0) acc = 0;
            S[(size_t)i * nR + j] = acc;
        }
    float* rsum = (float*)malloc(sizeof(float) * lo_cnt(nR));
    float* lsum = (float*)malloc(sizeof(float) * lo_cnt(nL));
    for (int j = 0; j < nR; ++j) {
        float acc = nL > 0 ? S[j] : 0.0f;
        for (int i = 1; i < nL; ++i) acc += S[(size_t)i * nR + j];
        rsum[j] = acc;
    }
    for (int i = 0; i < nL; ++i) {
        float acc = nR > 0 ? S[(size_t)i * nR] : 0.0f;
        for (int j = 1; j < nR; ++j) acc += S[(size_t)i * nR + j];
        lsum[i] = acc;
    }
    for (int i = 0; i < nL; ++i)
        for (int j = 0; j < nR; ++j) {
            float s = S[(size_t)i * nR + j];
            /* matcher.cpp:467 — float sums, then "+0.000001" promotes the denominator and the
             * division to double; the quotient is narrowed to float on assignment */
            float q = s / (lsum[i] + rsum[j] - s + 0.000001);
            nrm[(size_t)i * nR + j] = q;
        }
    int* y = (int*)malloc(sizeof(int) * lo_cnt((long)tot));
    lo_std_sort_desc(nrm, y, (int)tot);
    int topN = 120;
    if ((long)tot < topN) topN = (int)tot;
    for (int t = 0; t < topN; ++t) {
        int i = y[t] / nR, j = y[t] - i * nR;
        v[t] = S[(size_t)i * nR + j]; /* the RAW similarity, matcher.cpp:486 */
        li[t] = i;
        rj[t] = j;
    }
    if (S_out) memcpy(S_out, S, sizeof(float) * tot);
    if (nrm_out) memcpy(nrm_out, nrm, sizeof(float) * tot);
    free(S); free(nrm); free(rsum); free(lsum); free(y);
    return topN;
}

/* One2One_minutiae_matching, matcher.cpp:420-516. */
float lo_minutiae_score(const lo_template_t* L, const lo_template_t* R) {
    float v[120], v2[120], v3[120];
    int li[120], rj[120], li2[120], rj2[120], li3[120], rj3[120];
    int n1 = lo_minutiae_initial_corr(L->des, L->n, R->des, R->n, R->des_len, v, li, rj, NULL, NULL);
    int n2 = lo_prune_dist_euclid(v, li, rj, n1, L, R, v2, li2, rj2);
    int n3 = lo_prune_angle(v2, li2, rj2, n2, L, R, v3, li3, rj3);
    float score = 0.0f;
    for (int i = 0; i < n3; ++i) score += v3[i];
    return score;
}

/* One2One_texture_matching, matcher.cpp:531-783.  `lut` is K1's output for L. */
float lo_texture_score(const lo_template_t* L, const float* lut, const lo_template_t* R, const float* table,
                       int subs, int clusters) {
    int nL = L->n > LO_MAX_TEX ? LO_MAX_TEX : L->n; /* matcher.cpp:544-547 */
    int nR = R->n > LO_MAX_TEX ? LO_MAX_TEX : R->n;
    float* sim = (float*)malloc(sizeof(float) * lo_cnt((long)nL * nR));
    lo_texture_similarity(lut, nL, R->codes, nR, subs, clusters, R->des_len, sim);
    const int N = 200; /* matcher.cpp:33 */
    int cap = nL > N ? N : nL;
    if (cap < 1) cap = 1;
    float *v = (float*)malloc(sizeof(float) * 3 * (size_t)cap), *v2 = v + cap, *v3 = v2 + cap;
    int *li = (int*)malloc(sizeof(int) * 6 * (size_t)cap), *rj = li + cap, *li2 = rj + cap, *rj2 = li2 + cap,
        *li3 = rj2 + cap, *rj3 = li3 + cap;
    lo_template_t Lc = *L, Rc = *R;
    Lc.n = nL;
    Rc.n = nR;
    int n1 = lo_texture_initial_corr(sim, nL, nR, N, v, li, rj);
    int n2 = lo_prune_dist_lookup(v, li, rj, n1, &Lc, &Rc, table, v2, li2, rj2);
    int n3 = lo_prune_angle(v2, li2, rj2, n2, &Lc, &Rc, v3, li3, rj3);
    float score = 0.0f;
    for (int i = 0; i < n3; ++i) score += v3[i];
    free(sim); free(v); free(li);
    return score;
}

/* Score fusion, matcher.cpp:188 / :293: float sums, then "+ score[28]*0.3" in double. */
float lo_fuse(float s0, float s1, float s2, float s28) {
    float final_score = s0 + s1 + s2 + s28 * 0.3;
    return final_score;
}

/* One2One_matching_selected_templates + fusion, matcher.cpp:376-417 and :179-189.
 * latent_minu: the latent's NON-EMPTY minutiae templates in file order (the loader drops empty
 * ones, matcher.cpp:834-836).  Returns 1 / 2 like the reference; -100 when score[28] would be read
 * out of bounds (fewer than 29 template slots), which is outside the parity domain.
 * comp = {score[0], score[1], score[2], score[28]} of the reference's score vector: minutiae score i
 * lands in score[i] (:406), the texture score in score[n_latent_minu] (:414) - i.e. in score[28] for
 * the regular 28-template latent, in one of score[0..2] for a latent with at most 2 minutiae
 * templates, and in a slot the fusion never reads otherwise. */
int lo_score_pair(const lo_template_t* latent_minu, int n_latent_minu, const lo_template_t* latent_tex,
                  int n_latent_tex, const float* latent_lut, const lo_template_t* rolled_minu, int n_rolled_minu,
                  const lo_template_t* rolled_tex, int n_rolled_tex, const float* table, int subs, int clusters,
                  float* comp, float* final_score) {
    static const int selected[3] = {27 - 1, 3 - 1, 12 - 1}; /* matcher.cpp:380 */
    comp[0] = comp[1] = comp[2] = comp[3] = 0.0f;
    *final_score = -1.0f;
    if (n_latent_minu <= selected[0] && n_latent_tex <= 0) return 1;
    if (n_rolled_minu <= 0 && n_rolled_tex <= 0) return 2;
    if (n_latent_minu + n_latent_tex < 29) return -100;
    for (int i = 0; i < 3 && n_rolled_minu > 0; ++i) {
        if (n_latent_minu <= selected[i]) continue;
        comp[i] = lo_minutiae_score(&latent_minu[selected[i]], &rolled_minu[0]);
    }
    if (n_latent_tex > 0 && n_rolled_tex > 0 && (n_latent_minu == 28 || n_latent_minu <= 2)) {
        const float t = lo_texture_score(&latent_tex[0], latent_lut, &rolled_tex[0], table, subs, clusters);
        comp[n_latent_minu == 28 ? 3 : n_latent_minu] = t;
    }
    *final_score = lo_fuse(comp[0], comp[1], comp[2], comp[3]);
    return 0;
}
