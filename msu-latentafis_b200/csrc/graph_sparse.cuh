// Second-order graph pruning, sparse fast path — stages K3b, K4 / K8, K9, K9b for the overwhelmingly
// common case (impostor pairs), with the dense kernels of graph_prune.cuh as the fallback.
//
//   reference: top-N rows by similarity             matching/matcher.cpp:736-749   (K3b, texture)
//              LSS_R_Fast2_Dist_lookup              matching/matcher.cpp:1225-1348 (K4, texture)
//              LSS_R_Fast2_Dist_eigen               matching/matcher.cpp:1350-1469 (K8, minutiae)
//              LSS_R_Fast2 + adjust_angle           matching/matcher.cpp:1471-1647 (K9, both)
//              score = sum of surviving similarities :508-514, :775-781            (K9b)
//
// Why sparse.  H[a][b] of the distance-consistency graph is non-zero only when the two candidate
// correspondences a, b preserve their mutual distance to within 30 px; for a non-mated pair that holds
// for ~9 % of the entries (measured: at most 28 of 120 per row).  The reference's dense mat-vec adds
// H[i][k]*b[k] for k ascending; a zero entry contributes +-0, which leaves an fp32 accumulator
// unchanged, so summing only the non-zero entries in ascending k is bit-identical and ~10x less work,
// and the graph fits in ~20 KB of shared memory instead of 57-160 KB: 11 (minutiae) / 4 (texture) CTAs
// per SM instead of 3 / 1.  Rows are built by one warp each (lanes = columns, ballot compaction keeps
// ascending column order) into that warp's private slice of a CSR buffer - no atomics.
//
// The orientation graph runs on the survivors of the first stage (typically < 10) inside ONE warp with
// the graph as 32-bit row masks; no block barrier after the distance stage.
//
// Anything that does not fit - a row slice overflowing (mated pairs have dense graphs) or more than 32
// survivors - is appended to an overflow list and recomputed from scratch by the dense kernel.
#pragma once
#include "graph_prune.cuh"

namespace lafis {

template <bool LOOKUP>
struct SparseGeom;
template <>
struct SparseGeom<false> {  // minutiae: <= 120 candidates
    static constexpr int MAXN = kTopCorrMinu, MAXP = 128, NT = 128, CAP = 3072;
};
template <>
struct SparseGeom<true> {  // texture: <= 200 candidates
    static constexpr int MAXN = kTopCorrTex, MAXP = 224, NT = 256, CAP = 5632;
};

template <bool LOOKUP>
struct SparseWork {
    using G = SparseGeom<LOOKUP>;
    float vals[G::CAP];  // CSR values; before the graph is built the texture kernel sorts row maxima here
    float v[G::MAXP];
    float lo[G::MAXP], ro[G::MAXP];
    float b[G::MAXP], c[G::MAXP];
    short2 lxy[G::MAXP], rxy[G::MAXP];
    // texture graph only: the same coordinates as floats (exact), its pre-test runs on the fp32 pipes
    float2 lxyf[LOOKUP ? G::MAXP : 1], rxyf[LOOKUP ? G::MAXP : 1];
    unsigned short li[G::MAXP], rj[G::MAXP];
    unsigned short row_start[G::MAXP], row_len[G::MAXP];
    unsigned short y[G::MAXP];    // candidates in std::sort order
    unsigned short sel[G::MAXP];  // accepted candidates, acceptance order
    unsigned char cols[G::CAP];
    unsigned char rowj[G::NT / 32][G::MAXP];  // per-warp scratch: columns that passed the cheap test
    unsigned long long skeys[G::MAXP <= 128 ? 128 : 256];  // (b, index) keys of the candidate ranking
    float f;
    int overflow, tie, nsel;
    // orientation stage (<= 32 survivors), only touched when the introsort replay is needed
    float s2[32];
    unsigned short y2[32];
};

// Cheap, conservative pre-test of "H[a][b] may be non-zero" evaluated for ALL pairs; the exact entry is then
// computed only for the survivors (~9 % of the pairs of a non-mated print), compacted so that the expensive
// correctly-rounded square roots and divisions run on full warps.
// Coordinates arrive as floats (small integers, exact) so that no integer->float conversion and no 64-bit
// integer product is needed per pair.
template <bool LOOKUP>
__device__ __forceinline__ bool pair_may_connect(float2 la, float2 lb, float2 ra, float2 rb) {
    const float dx1 = la.x - lb.x, dx2 = ra.x - rb.x, dy1 = la.y - lb.y, dy2 = ra.y - rb.y;  // exact
    const float s1 = fmaf(dx1, dx1, dy1 * dy1), s2 = fmaf(dx2, dx2, dy2 * dy2);               // exact integers < 2^24
    if (LOOKUP) {
        // matcher.cpp:1246-1266 with table[dx][dy] = fl(16 sqrt(dx^2 + dy^2)): |d1 - d2| <= 30 needs
        // (sqrt(s1) - sqrt(s2))^2 <= (30/16)^2 = 3.515625, i.e. s1 + s2 - 3.515625 <= 2 sqrt(s1 s2), with s < 5000
        // once every |dx|, |dy| is below 50 (:1257).  Tested with 3.625 (the table entries are rounded to fp32,
        // relative 6e-8 on values <= 1110; t*t and 4*s1*s2 round with relative 6e-8 on values whose exact
        // versions differ by > 1e-5 relative whenever the two constants matter): never rejects a connected pair.
        // (the |dx|, |dy| < 50 condition of :1257 is left to the exact entry: pair_h returns 0 for such pairs, and on
        //  the reference's <= 50 x 50 block grids it never fails, so testing it here only costs instructions)
        const float u = fmaxf(fmaf(s1 + s2, 0.5f, -1.8125f), 0.0f);  // (s1 + s2 - 3.625) / 2, exact
        return u * u <= s1 * s2;
    } else {
        // |d1 - d2| <= 30.04  <=>  s1 + s2 - 30.04^2 <= 2 sqrt(s1 s2), with s = d^2 an exact integer below
        // 2^23 (coordinates below 2048: jobs with larger ones go to the dense kernel).  Rounding moves the
        // comparison by < 0.2 squared pixels, the margin over 30^2 is 2.4: nothing with |d1 - d2| <= 30 is
        // rejected here.
        const float u = fmaxf(fmaf(s1 + s2, 0.5f, -451.2f), 0.0f);
        return u * u <= s1 * s2;
    }
}

// The minutiae graph keeps its candidates' pixel coordinates as short2 in registers (4 chunks) and forms the
// squared distances in integers; measured faster there than the float-coordinate form above (14.4 vs 13.9 ms).
// |d1 - d2| <= 30.04  <=>  (s1 + s2 - 902.4) / 2 <= sqrt(s1 s2); coordinates are below 2048 (larger ones: dense kernel).
__device__ __forceinline__ bool pair_may_connect_px(short2 la, short2 lb, short2 ra, short2 rb) {
    const int dx1 = (int)la.x - (int)lb.x, dx2 = (int)ra.x - (int)rb.x;
    const int dy1 = (int)la.y - (int)lb.y, dy2 = (int)ra.y - (int)rb.y;
    const float s1 = (float)(dx1 * dx1 + dy1 * dy1), s2 = (float)(dx2 * dx2 + dy2 * dy2);  // exact below 2^23
    const float u = fmaxf(fmaf(s1 + s2, 0.5f, -451.2f), 0.0f);                              // (s1 + s2 - 902.4) / 2
    return u * u <= s1 * s2;
}

// H entry of the distance-consistency graph for candidates a, b (symmetric in a, b), exact.
template <bool LOOKUP>
__device__ __forceinline__ float pair_h(short2 la, short2 lb, short2 ra, short2 rb, const float* __restrict__ table) {
    float d1, d2;
    if (LOOKUP) {  // matcher.cpp:1246-1262
        const int dx1 = abs((int)la.x - (int)lb.x), dx2 = abs((int)ra.x - (int)rb.x);
        const int dy1 = abs((int)la.y - (int)lb.y), dy2 = abs((int)ra.y - (int)rb.y);
        if ((dx1 >= kTableN) | (dx2 >= kTableN) | (dy1 >= kTableN) | (dy2 >= kTableN)) return 0.0f;
        d1 = __ldg(table + dx1 * kTableN + dy1);
        d2 = __ldg(table + dx2 * kTableN + dy2);
    } else {  // matcher.cpp:1372-1384
        const float dx1 = (float)((int)la.x - (int)lb.x), dx2 = (float)((int)ra.x - (int)rb.x);
        const float dy1 = (float)((int)la.y - (int)lb.y), dy2 = (float)((int)ra.y - (int)rb.y);
        d1 = __fsqrt_rn(f_add(f_mul(dx1, dx1), f_mul(dy1, dy1)));
        d2 = __fsqrt_rn(f_add(f_mul(dx2, dx2), f_mul(dy2, dy2)));
    }
    const float dist = fabsf(f_sub(d1, d2));
    if (dist > 30.0f) return 0.0f;
    // (30 - dist) / 25.0: a double division narrowed to float == the correctly rounded float division
    float h = f_div(f_sub(30.0f, dist), 25.0f);
    if (h > 1.0f) h = 1.0f;
    else if (h < 0.0f) h = 0.0f;
    return h;
}

// orientation compatibility of survivors i < j (matcher.cpp:1495-1549); "1" is i, "2" is j
__device__ __forceinline__ bool angle_compatible(short2 l1, short2 l2, short2 r1, short2 r2, float lo1, float lo2,
                                                 float ro1, float ro2) {
    float a1 = adjust_angle_ref(f_sub(lo1, lo2));
    float a2 = adjust_angle_ref(f_sub(ro1, ro2));
    if ((double)angle_gap_ref(a1, a2) > LAFIS_PI_D / 4.) return false;
    const float dx1 = (float)((int)l1.x - (int)l2.x), dy1 = (float)((int)l1.y - (int)l2.y);
    const float line1 = -atan2f_fdlibm(dy1, dx1);
    const float dx2 = (float)((int)r1.x - (int)r2.x), dy2 = (float)((int)r1.y - (int)r2.y);
    const float line2 = -atan2f_fdlibm(dy2, dx2);
    a1 = adjust_angle_ref(f_sub(lo1, line1));
    a2 = adjust_angle_ref(f_sub(ro1, line2));
    if ((double)angle_gap_ref(a1, a2) > LAFIS_PI_D / 6.) return false;
    a1 = adjust_angle_ref(f_sub(lo2, line1));
    a2 = adjust_angle_ref(f_sub(ro2, line2));
    if ((double)angle_gap_ref(a1, a2) > LAFIS_PI_D / 6.) return false;
    return true;
}

// The cascade on the candidate list held in w (v, li, rj, coordinates, orientations).  Returns true
// when the score (thread 0) is valid, false when the job must go to the dense kernel.
template <bool LOOKUP>
__device__ bool sparse_cascade(SparseWork<LOOKUP>& w, int num, const float* __restrict__ table, float* score_out) {
    using G = SparseGeom<LOOKUP>;
    constexpr int NT = G::NT, NW = NT / 32, CAPW = G::CAP / NW, CH = G::MAXP / 32;
    constexpr int ITERS = LOOKUP ? 3 : 5;  // matcher.cpp:1284 / :1406
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    *score_out = 0.0f;
    if (num <= 0) return true;

    // ---- CSR rows of the distance-consistency graph ----
    {
        // column coordinates of the pre-test: registers for the minutiae graph (4 chunks), shared memory for the
        // texture graph (7 chunks would cost 28 registers and an occupancy step)
        constexpr int CHR = LOOKUP ? 1 : CH;
        short2 clxy[CHR], crxy[CHR];
        if (!LOOKUP) {
#pragma unroll
            for (int c = 0; c < CHR; ++c) {
                const int j = lane + 32 * c;
                clxy[c] = (j < num) ? w.lxy[j] : make_short2(0, 0);
                crxy[c] = (j < num) ? w.rxy[j] : make_short2(0, 0);
            }
        }
        const int wbase = warp * CAPW;
        int used = 0;
        bool over = false;
        const unsigned lt_mask = (1u << lane) - 1u;
        unsigned char* rowj = w.rowj[warp];
        for (int i = warp; i < num; i += NW) {
            const short2 la = w.lxy[i], ra = w.rxy[i];
            // pass 1: columns that may connect to row i, ascending
            int cnt = 0;
            if constexpr (LOOKUP) {
                const float2 laf = w.lxyf[i], raf = w.rxyf[i];
                for (int j = lane; j - lane < num; j += 32) {
                    const bool pass = j < num && j != i && pair_may_connect<LOOKUP>(laf, w.lxyf[j], raf, w.rxyf[j]);
                    const unsigned m = __ballot_sync(0xffffffffu, pass);
                    if (pass) rowj[cnt + __popc(m & lt_mask)] = (unsigned char)j;
                    cnt += __popc(m);
                }
            } else {
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int j = lane + 32 * c;
                    const bool pass = j < num && j != i && pair_may_connect_px(la, clxy[c], ra, crxy[c]);
                    const unsigned m = __ballot_sync(0xffffffffu, pass);
                    if (pass) rowj[cnt + __popc(m & lt_mask)] = (unsigned char)j;
                    cnt += __popc(m);
                }
            }
            __syncwarp();
            // pass 2: exact entries of the survivors, compacted again (an exact entry can still be 0)
            int len = 0;
            for (int b0 = 0; b0 < cnt; b0 += 32) {
                const int idx = b0 + lane;
                float h = 0.0f;
                int j = 0;
                if (idx < cnt) {
                    j = rowj[idx];
                    h = pair_h<LOOKUP>(la, w.lxy[j], ra, w.rxy[j], table);
                }
                const unsigned m = __ballot_sync(0xffffffffu, h > 0.0f);
                const int pos = used + len + __popc(m & lt_mask);
                if (h > 0.0f && pos < CAPW) {
                    w.vals[wbase + pos] = h;
                    w.cols[wbase + pos] = (unsigned char)j;
                }
                len += __popc(m);
            }
            __syncwarp();
            if (used + len > CAPW) over = true;
            if (lane == 0) {
                w.row_start[i] = (unsigned short)(wbase + used);
                w.row_len[i] = (unsigned short)len;
            }
            used += len;
        }
        if (over && lane == 0) w.overflow = 1;
    }
    if (tid < num) w.b[tid] = w.v[tid];
    __syncthreads();
    if (w.overflow) return false;

    // ---- power iteration: c = H b (non-zeros, ascending k), b = c * (float)(1 / (sum c + 1e-5)) ----
    for (int it = 0; it < ITERS; ++it) {
        if (tid < num) {
            const int rs = w.row_start[tid], len = w.row_len[tid];
            float acc = 0.0f;
            for (int e = 0; e < len; ++e) acc = f_add(acc, f_mul(w.vals[rs + e], w.b[w.cols[rs + e]]));
            w.c[tid] = acc;
        }
        __syncthreads();
        if (tid == 0) {
            float sum = w.c[0];
#pragma unroll 8
            for (int i = 1; i < num; ++i) sum = f_add(sum, w.c[i]);
            w.f = (float)(1.0 / ((double)sum + 0.00001));
        }
        __syncthreads();
        if (tid < num) w.b[tid] = f_mul(w.c[tid], w.f);
        __syncthreads();
    }

    // ---- order by b descending with std::sort's permutation: (value desc, index asc) keys through the block's
    //      bitonic sort; equal values that the greedy pass can reach need the introsort replay ----
    {
        constexpr int P2 = G::MAXP <= 128 ? 128 : 256;
        static_assert(P2 <= NT && G::MAXN <= P2, "one key per thread");
        unsigned long long k = 0ull;  // padding keys sort last
        if (tid < num) {
            uint32_t u = __float_as_uint(w.b[tid]);  // b >= 0: the bit pattern orders like the value
            if (u == 0x80000000u) u = 0u;
            k = ((unsigned long long)u << 32) | (unsigned long long)(0xffffu - (unsigned)tid);
        }
        if (tid < P2) w.skeys[tid] = k;
        __syncthreads();
        block_bitonic_desc<NT, 1>(w.skeys, P2);
        if (tid < num) {
            const unsigned long long me = w.skeys[tid];
            w.y[tid] = (unsigned short)(0xffffu - (unsigned)(me & 0xffffu));
            if (tid + 1 < num && num > 16 && (me >> 32) == (w.skeys[tid + 1] >> 32) &&
                !((double)__uint_as_float((uint32_t)(me >> 32)) < 0.0001))
                w.tie = 1;
        }
    }
    __syncthreads();
    if (w.tie) {
        if (tid == 0) std_sort_desc_emulate<float, unsigned short>(w.b, w.y, num);
        __syncthreads();
    }
    if (warp != 0) return true;  // the rest runs in warp 0 only; thread 0 carries the result

    // ---- greedy selection (matcher.cpp:1305-1345): lanes own sorted positions p = lane + 32 c ----
    int n2 = 0;
    {
        unsigned short ind[CH];
        unsigned open = 0;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int p = lane + 32 * c;
            ind[c] = (p < num) ? w.y[p] : 0;
            if (p < num && !((double)w.b[ind[c]] < 0.0001)) open |= 1u << c;
        }
        for (;;) {
            int pos = -1;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const unsigned m = __ballot_sync(0xffffffffu, (open >> c) & 1u);
                if (m && pos < 0) pos = 32 * c + __ffs(m) - 1;
            }
            if (pos < 0) break;
            const int s = w.y[pos];
            if (lane == 0) w.sel[n2] = (unsigned short)s;
            ++n2;
            const unsigned short sli = w.li[s], srj = w.rj[s];
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (!((open >> c) & 1u)) continue;
                const int p = lane + 32 * c;
                bool keep = false;
                if (p != pos && w.li[ind[c]] != sli && w.rj[ind[c]] != srj) {
                    // H[ind][s] >= 1e-5 ?  (an absent entry is 0)
                    const int rs = w.row_start[ind[c]], len = w.row_len[ind[c]];
                    for (int e = 0; e < len; ++e)
                        if (w.cols[rs + e] == s) {
                            keep = !((double)w.vals[rs + e] < 0.00001);
                            break;
                        }
                }
                if (!keep) open &= ~(1u << c);
            }
        }
    }
    __syncwarp();
    if (n2 > 32) {
        if (lane == 0) w.overflow = 1;
        return false;
    }
    if (n2 == 0) return true;

    // ---- orientation-consistency graph on the n2 survivors, lane = survivor (acceptance order) ----
    const int me = (lane < n2) ? w.sel[lane] : w.sel[0];
    const short2 mlxy = w.lxy[me], mrxy = w.rxy[me];
    const float mlo = w.lo[me], mro = w.ro[me], mv = w.v[me];
    const unsigned short mli = w.li[me], mrj = w.rj[me];
    unsigned mask = 0;
    for (int j = 1; j < n2; ++j) {
        const short2 olxy = make_short2((short)__shfl_sync(0xffffffffu, (int)mlxy.x, j), (short)__shfl_sync(0xffffffffu, (int)mlxy.y, j));
        const short2 orxy = make_short2((short)__shfl_sync(0xffffffffu, (int)mrxy.x, j), (short)__shfl_sync(0xffffffffu, (int)mrxy.y, j));
        const float olo = __shfl_sync(0xffffffffu, mlo, j), oro = __shfl_sync(0xffffffffu, mro, j);
        if (lane < j && angle_compatible(mlxy, olxy, mrxy, orxy, mlo, olo, mro, oro)) mask |= 1u << j;
    }
    for (int j = 0; j < n2; ++j) {  // symmetric half
        const unsigned mj = __shfl_sync(0xffffffffu, mask, j);
        if (j < lane && ((mj >> lane) & 1u)) mask |= 1u << j;
    }
    if (lane >= n2) mask = 0;
    float S = (float)(1.0 / (double)n2);  // matcher.cpp:1558
    for (int it = 0; it < 5; ++it) {      // matcher.cpp:1563-1581
        float acc = 0.0f;
        for (int k = 0; k < n2; ++k) {
            const float sk = __shfl_sync(0xffffffffu, S, k);
            if ((mask >> k) & 1u) acc = f_add(acc, sk);
        }
        float sum = 0.0f;
        for (int j = 0; j < n2; ++j) sum = f_add(sum, __shfl_sync(0xffffffffu, acc, j));
        const float fac = (float)(1.0 / ((double)sum + 0.00001));
        S = f_mul(acc, fac);
    }
    // sorted position of every survivor
    int rank = 0;
    bool tie = false;
    for (int k = 0; k < n2; ++k) {
        const float ok = __shfl_sync(0xffffffffu, S, k);
        if (lane < n2) {
            rank += (ok > S) || (ok == S && k < lane);
            tie |= (ok == S && k != lane);
        }
    }
    const bool need_replay = __any_sync(0xffffffffu, lane < n2 && tie && n2 > 16 && !((double)S < 0.001));
    if (need_replay) {
        if (lane < n2) w.s2[lane] = S;
        __syncwarp();
        if (lane == 0) std_sort_desc_emulate<float, unsigned short>(w.s2, w.y2, n2);
        __syncwarp();
        if (lane < n2)
            for (int p = 0; p < n2; ++p)
                if (w.y2[p] == lane) rank = p;
    }
    // greedy over sorted positions (matcher.cpp:1596-1633); lane keeps its own survivor
    bool open = lane < n2 && !((double)S < 0.001);
    float score = 0.0f;
    for (;;) {
        // first open sorted position
        int best = open ? rank : 0x7fffffff;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, d));
        if (best == 0x7fffffff) break;
        const unsigned who = __ballot_sync(0xffffffffu, open && rank == best);
        const int sl = __ffs(who) - 1;  // lane of the accepted survivor
        score = f_add(score, __shfl_sync(0xffffffffu, mv, sl));
        const unsigned short sli = (unsigned short)__shfl_sync(0xffffffffu, (int)mli, sl);
        const unsigned short srj = (unsigned short)__shfl_sync(0xffffffffu, (int)mrj, sl);
        if (open && (lane == sl || mli == sli || mrj == srj || !((mask >> sl) & 1u))) open = false;
    }
    *score_out = score;
    return true;
}

// ------------------------------------------------------------------------------------------------
// minutiae components: grid = jobs = Q * n_chunk * 3, block = 128
// ------------------------------------------------------------------------------------------------
struct OverflowList {
    int* count;
    int* jobs;
};

__global__ void __launch_bounds__(SparseGeom<false>::NT) graph_minu_sparse_kernel(GraphMinuParams P, OverflowList ov) {
    extern __shared__ __align__(128) unsigned char smem[];
    SparseWork<false>& w = *reinterpret_cast<SparseWork<false>*>(smem);
    const int tid = threadIdx.x;
    const size_t oidx = blockIdx.x;  // (q * n_chunk + tl) * 3 + slot
    const int slot = (int)(oidx % 3);
    const size_t pair = oidx / 3;
    const int tl = (int)(pair % P.n_chunk), q = (int)(pair / P.n_chunk);
    const int num = P.corr_n[oidx];
    if (tid == 0) {
        w.overflow = 0;
        w.tie = 0;
    }
    int big = 0;
    if (tid < num) {
        const uint32_t ij = P.corr_ij[oidx * kTopCorrMinu + tid];
        const int i = (int)(ij >> 16), j = (int)(ij & 0xffffu);
        w.v[tid] = P.corr_v[oidx * kTopCorrMinu + tid];
        w.li[tid] = (unsigned short)i;
        w.rj[tid] = (unsigned short)j;
        const uint32_t lo = P.slot_off[q * 3 + slot] + i, go = P.minu_off[P.g0 + tl] + j;
        w.lxy[tid] = P.lat_xy[lo];
        w.rxy[tid] = P.gal_xy[go];
        w.lo[tid] = P.lat_ori[lo];
        w.ro[tid] = P.gal_ori[go];
        // the pre-test squares coordinate differences in fp32: exact only below 2048 px
        // (coordinates in [0, 2048): differences below 2048, squared distances below 2^23)
        big = ((unsigned)(int)w.lxy[tid].x | (unsigned)(int)w.lxy[tid].y | (unsigned)(int)w.rxy[tid].x |
               (unsigned)(int)w.rxy[tid].y) >= 2048u;
    }
    if (__syncthreads_or(big)) {  // larger images: the dense kernel evaluates every entry exactly
        if (tid == 0) ov.jobs[atomicAdd(ov.count, 1)] = (int)oidx;
        return;
    }
    float score;
    const bool ok = sparse_cascade<false>(w, num, nullptr, &score);
    if (tid == 0) {
        if (ok) P.comp[((size_t)q * P.G + P.g0 + tl) * 4 + slot] = score;
        else ov.jobs[atomicAdd(ov.count, 1)] = (int)oidx;
    }
}

// ------------------------------------------------------------------------------------------------
// texture component: grid = jobs = Q * n_chunk, block = 256
// ------------------------------------------------------------------------------------------------
// order-preserving key of (value desc, row asc) with the arg-max column in the low bits
__device__ __forceinline__ unsigned long long row_key(float v, int row, unsigned short col) {
    if (v == 0.0f) v = 0.0f;  // -0 and +0 compare equal in the reference's comparator
    uint32_t u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | ((unsigned long long)(0xffffu - (unsigned)row) << 16) | col;
}

__global__ void __launch_bounds__(SparseGeom<true>::NT) graph_tex_sparse_kernel(GraphTexParams P, OverflowList ov) {
    extern __shared__ __align__(128) unsigned char smem[];
    SparseWork<true>& w = *reinterpret_cast<SparseWork<true>*>(smem);
    constexpr int NT = SparseGeom<true>::NT;
    const int tid = threadIdx.x;
    const size_t pair = blockIdx.x;
    const int tl = (int)(pair % P.n_chunk), q = (int)(pair / P.n_chunk);
    const int g = P.g0 + tl;
    const int nLt = (P.lat_status[q] == 0) ? P.lat_nt[q] : 0;
    const uint32_t gbase = P.tex_off[g];
    const int nRt = (int)(P.tex_off[g + 1] - gbase);
    if (nLt <= 0 || nRt <= 0) {  // matcher.cpp:411: no texture template on one side, score stays 0
        if (tid == 0) P.comp[((size_t)q * P.G + g) * 4 + 3] = 0.0f;
        return;
    }
    if (tid == 0) {
        w.overflow = 0;
        w.tie = 0;
    }
    const size_t rbase = pair * (size_t)P.lt_stride;
    int num;
    if (nLt > kTopCorrTex) {
        // ---- K3b: the 200 best rows in std::sort order (matcher.cpp:736-749): bitonic sort of
        //      (value desc, row asc) keys in the (still unused) CSR value area ----
        unsigned long long* keys = reinterpret_cast<unsigned long long*>(w.vals);  // <= 1024 keys = 8 KB
        float* rv = reinterpret_cast<float*>(keys + 1024);                         // [1024]
        unsigned short* ry = reinterpret_cast<unsigned short*>(rv + 1024);         // [1024]
        int np2 = 256;
        while (np2 < nLt) np2 <<= 1;
        for (int i = tid; i < np2; i += NT) {
            unsigned long long k = 0ull;
            if (i < nLt) {
                const float val = P.rowmax_val[rbase + i];
                rv[i] = val;
                k = row_key(val, i, P.rowmax_j[rbase + i]);
            }
            keys[i] = k;
        }
        __syncthreads();
        block_bitonic_desc<NT, 1024 / NT>(keys, np2);
        // a tie that reaches into the first 200 positions makes the permutation introsort-specific
        if (tid < kTopCorrTex && (keys[tid] >> 32) == (keys[tid + 1] >> 32)) w.tie = 1;
        __syncthreads();
        num = kTopCorrTex;
        const bool replay = w.tie != 0;
        __syncthreads();
        if (replay) {
            if (tid == 0) {
                std_sort_desc_prefix(DenseKey<float>{rv}, ry, nLt, kTopCorrTex);
                atomicAdd(P.slow_path_count, 1ull);
                w.tie = 0;
            }
            __syncthreads();
            if (tid < num) {
                const int i = ry[tid];
                w.v[tid] = rv[i];
                w.li[tid] = (unsigned short)i;
                w.rj[tid] = P.rowmax_j[rbase + i];
            }
        } else if (tid < num) {
            const unsigned long long k = keys[tid];
            const int i = 0xffff - (int)((k >> 16) & 0xffffu);
            w.v[tid] = rv[i];
            w.li[tid] = (unsigned short)i;
            w.rj[tid] = (unsigned short)(k & 0xffffu);
        }
    } else {
        num = nLt;
        if (tid < num) {
            w.v[tid] = P.rowmax_val[rbase + tid];
            w.li[tid] = (unsigned short)tid;
            w.rj[tid] = P.rowmax_j[rbase + tid];
        }
    }
    __syncthreads();  // the key area becomes the CSR value area from here on
    if (tid < num) {
        const int i = w.li[tid], j = w.rj[tid];
        w.lxy[tid] = P.lat_xy[(size_t)q * P.lt_stride + i];
        w.rxy[tid] = P.gal_xy[gbase + j];
        w.lxyf[tid] = make_float2((float)w.lxy[tid].x, (float)w.lxy[tid].y);
        w.rxyf[tid] = make_float2((float)w.rxy[tid].x, (float)w.rxy[tid].y);
        w.lo[tid] = P.lat_ori[(size_t)q * P.lt_stride + i];
        w.ro[tid] = P.gal_ori[gbase + j];
    }
    __syncthreads();
    float score;
    const bool ok = sparse_cascade<true>(w, num, P.table, &score);
    if (tid == 0) {
        if (ok) P.comp[((size_t)q * P.G + g) * 4 + 3] = score;
        else ov.jobs[atomicAdd(ov.count, 1)] = (int)pair;
    }
}

}  // namespace lafis
