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
// and the graph fits in ~20 KB of shared memory instead of 57-160 KB: 10 (minutiae) / 5 (texture) CTAs
// per SM instead of 3 / 1.  The graph is built symmetrically from a bit matrix of the pairs a < b that pass a
// conservative pre-test; the exact entries are computed once per pair and stored at both CSR positions, which
// follow from popcounts of the bit rows (ascending column order, no sorting).
//
// Rankings (candidate order, the texture graph's top-200 rows) come from one bucket pass (block_rank_desc,
// device_common.cuh); std::sort's permutation inside groups of equal values is replayed only when it can change
// the outcome (see the greedy pass).
//
// The orientation graph runs on the survivors of the first stage (typically < 10) inside ONE warp with
// the graph as 32-bit row masks; no block barrier after the distance stage.
//
// Anything that does not fit - a CSR overflowing (clustered minutiae: second chance in graph_minu_mid_kernel;
// mated pairs have dense graphs) or more than 32 survivors - is appended to an overflow list and recomputed from
// scratch by the dense kernel.
#pragma once
#include "graph_prune.cuh"

namespace lafis {

template <bool LOOKUP, int TIER = 0>
struct SparseGeom;
template <>
struct SparseGeom<false, 0> {  // minutiae: <= 120 candidates
    static constexpr int MAXN = kTopCorrMinu, MAXP = 128, NT = 128, CAP = 2560, NCH = 4;
};
template <>
struct SparseGeom<false, 1> {  // minutiae, second chance for denser graphs (clustered minutiae): 46 % of the full matrix, 5 CTAs per SM
    static constexpr int MAXN = kTopCorrMinu, MAXP = 128, NT = 128, CAP = 6656, NCH = 4;
};
template <>
struct SparseGeom<true, 0> {  // texture: <= 200 candidates
    static constexpr int MAXN = kTopCorrTex, MAXP = 224, NT = 256, CAP = 4352, NCH = 7;
};

// Working set of one job.  The graph is built symmetrically: only the pairs a < b are tested, a bit matrix M records
// which of them may be connected (row a gets bit b from the test's ballot, row b gets bit a from a per-lane
// accumulator flushed with one shared-memory atomic per row chunk); the exact entries are then computed ONCE per
// surviving pair, the pairs spread evenly over all threads, and stored at both CSR positions, which follow from
// popcounts of the two bit rows.
template <bool LOOKUP, int TIER = 0>
struct SparseWork {
    using G = SparseGeom<LOOKUP, TIER>;
    static constexpr int P2 = G::MAXP <= 128 ? 128 : 256;
    float vals[G::CAP];            // CSR values; before the graph is built the texture kernel sorts row maxima here
    float4 cf[G::MAXP];            // candidate coordinates as floats (exact): latent x, rolled x, latent y, rolled y
    float v[G::MAXP];
    unsigned short li[G::MAXP], rj[G::MAXP];
    unsigned short row_start[P2], row_len[G::MAXP];
    unsigned char cols[G::CAP];
    unsigned short up_pos[P2];     // first CSR slot right of the diagonal of every row
    unsigned short up_start[P2];   // prefix sums of the rows' counts of pairs a < b
    union alignas(16) {            // the bit matrix is dead once the CSR is built: its storage becomes the iteration vectors
        uint32_t M[G::MAXP][G::NCH];       // bit matrix of possibly connected pairs
        struct {
            float b[G::MAXP], c[G::MAXP];
            unsigned short y[G::MAXP];     // candidates in std::sort order
            unsigned short sel[G::MAXP];   // accepted candidates, acceptance order
            uint2 buck[P2];                // candidate ranking (block_rank_desc): keys grouped by bucket
            int start[P2 + 1];             //   and the buckets' first ranks
        } it;
    } u;
    unsigned char mark[P2];        // greedy pass: mark[k] == n2 when H[s][k] >= 1e-5 for the candidate s accepted as number n2
    int wsum[2 * (G::NT / 32)];
    float f;
    int overflow, tie, nsel, npairs;
    // orientation stage (<= 32 survivors), only touched when the introsort replay is needed
    float s2[32];
    unsigned short y2[32];
    uint32_t omask[32];            // orientation graph: row masks of the <= 32 survivors
};

// Cheap, conservative test of "H[a][b] may be non-zero" for all pairs a < b; the exact entry is computed only for the
// survivors (~9 % of the pairs of a non-mated print).  Coordinates are small integers held exactly in fp32, laid out
// so that the latent and the rolled difference form one packed pair: two FFMA2, one FMUL2 and one FFMA2 give both
// squared distances (sm_100 packed fp32: half the issue slots of the scalar form in an issue-bound kernel).
//   minutiae (matcher.cpp:1372-1384): |d1 - d2| <= 30.04  <=>  (s1 + s2 - 902.4) / 2 <= sqrt(s1 s2), s = d^2 an exact
//     integer below 2^23 (coordinates below 2048: jobs with larger ones go to the dense kernel); rounding moves the
//     comparison by < 0.2 squared pixels, the margin over 30^2 is 2.4: nothing with |d1 - d2| <= 30 is rejected.
//   texture (matcher.cpp:1246-1266, table[dx][dy] = fl(16 sqrt(dx^2 + dy^2))): |d1 - d2| <= 30 needs
//     (sqrt(s1) - sqrt(s2))^2 <= (30/16)^2 = 3.515625; tested with 3.625 (table entries are rounded to fp32, relative
//     6e-8 on values <= 1110): never rejects a connected pair.  The |dx|, |dy| < 50 condition of :1257 is left to the
//     exact entry (pair_h returns 0 then; such a pair's test result is irrelevant).
template <bool LOOKUP>
__device__ __forceinline__ bool pair_may_connect(float4 a, float4 b) {
    const float2 m1 = make_float2(-1.0f, -1.0f);
    const float2 p = __ffma2_rn(make_float2(b.x, b.y), m1, make_float2(a.x, a.y));  // (dx1, dx2), exact
    const float2 q = __ffma2_rn(make_float2(b.z, b.w), m1, make_float2(a.z, a.w));  // (dy1, dy2), exact
    const float2 s = __ffma2_rn(q, q, __fmul2_rn(p, p));                            // (s1, s2), exact integers
    const float u = fmaxf(fmaf(s.x + s.y, 0.5f, LOOKUP ? -1.8125f : -451.2f), 0.0f);
    return u * u <= s.x * s.y;
}

// H entry of the distance-consistency graph for candidates a, b (symmetric in a, b), exact.
template <bool LOOKUP>
__device__ __forceinline__ float pair_h(float4 a, float4 b, const float* __restrict__ table) {
    const float dx1 = a.x - b.x, dx2 = a.y - b.y, dy1 = a.z - b.z, dy2 = a.w - b.w;  // exact (small integers)
    float d1, d2;
    if (LOOKUP) {  // matcher.cpp:1246-1262
        const int ix1 = (int)fabsf(dx1), ix2 = (int)fabsf(dx2), iy1 = (int)fabsf(dy1), iy2 = (int)fabsf(dy2);
        if ((ix1 >= kTableN) | (ix2 >= kTableN) | (iy1 >= kTableN) | (iy2 >= kTableN)) return 0.0f;
        d1 = __ldg(table + ix1 * kTableN + iy1);
        d2 = __ldg(table + ix2 * kTableN + iy2);
    } else {  // matcher.cpp:1372-1384
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

// number of set bits of a bit row strictly below column x
template <int NCH>
__device__ __forceinline__ int bits_below(const uint32_t* __restrict__ row, int x) {
    const int wx = x >> 5;
    const uint32_t low = (1u << (x & 31)) - 1u;
    int n = 0;
#pragma unroll
    for (int w = 0; w < NCH; ++w) {
        const uint32_t word = row[w];
        n += __popc(w < wx ? word : (w == wx ? (word & low) : 0u));
    }
    return n;
}

// orientation compatibility of survivors i < j (matcher.cpp:1495-1549); "1" is i, "2" is j; coordinates as the
// exact floats of SparseWork::cf (latent x, rolled x, latent y, rolled y)
__device__ __forceinline__ bool angle_compatible(float4 c1, float4 c2, float lo1, float lo2, float ro1, float ro2) {
    float a1 = adjust_angle_ref(f_sub(lo1, lo2));
    float a2 = adjust_angle_ref(f_sub(ro1, ro2));
    if ((double)angle_gap_ref(a1, a2) > LAFIS_PI_D / 4.) return false;
    const float dx1 = c1.x - c2.x, dy1 = c1.z - c2.z;  // exact: the reference's int difference converted to float
    const float line1 = -atan2f_fdlibm(dy1, dx1);
    const float dx2 = c1.y - c2.y, dy2 = c1.w - c2.w;
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
template <bool LOOKUP, int TIER>
__device__ bool sparse_cascade(SparseWork<LOOKUP, TIER>& w, int num, const float* __restrict__ table, const float* __restrict__ lat_ori,
                               const float* __restrict__ gal_ori, float* score_out) {
    using G = SparseGeom<LOOKUP, TIER>;
    constexpr int NT = G::NT, NW = NT / 32, CH = G::MAXP / 32, P2 = SparseWork<LOOKUP, TIER>::P2;
    constexpr int ITERS = LOOKUP ? 3 : 5;  // matcher.cpp:1284 / :1406
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    *score_out = 0.0f;
    if (num <= 0) return true;

    // ---- bit matrix of possibly connected pairs: only a < b is tested (rows a = warp, warp + NW, ...) ----
    {
        constexpr int NCH = G::NCH;
        // column coordinates: registers for the minutiae graph (4 chunks), shared memory for the texture graph
        // (7 chunks would cost 28 registers and an occupancy step)
        float4 colreg[LOOKUP ? 1 : NCH];
        if (!LOOKUP) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) colreg[c] = w.cf[lane + 32 * c];
        }
        bool valid[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) valid[c] = lane + 32 * c < num;
        // One phase per row chunk r, so that the column chunks it meets are known at compile time.  A pair is tested from
        // the row of its smaller index - except the pairs with a candidate of the LAST chunk, which are tested from that
        // candidate's row: the last chunk is the partial one (120 = 3 x 32 + 24 minutiae candidates, 200 = 6 x 32 + 8
        // texture candidates), and its few rows walking all chunks cost less than every other row walking a column chunk
        // that is three quarters empty (texture: 728 instead of 872 row x chunk steps).
        constexpr int L = NCH - 1;
#pragma unroll
        for (int r = 0; r < L; ++r) {  // column chunks [r, L)
            if (32 * r >= num) break;  // uniform
            uint32_t tacc[NCH];  // bit k of tacc[c]: row 32 r + k may connect to my column of chunk c (the transposed bits)
#pragma unroll
            for (int c = 0; c < NCH; ++c) tacc[c] = 0u;
            const int i_end = min(num, 32 * r + 32);
            for (int i = 32 * r + warp; i < i_end; i += NW) {
                const uint32_t kbit = 1u << (i & 31);
                const float4 row = w.cf[i];
                uint32_t keep = 0u;
#pragma unroll
                for (int c = r; c < L; ++c) {
                    const float4 col = LOOKUP ? w.cf[lane + 32 * c] : colreg[LOOKUP ? 0 : c];
                    const bool pass = pair_may_connect<LOOKUP>(row, col) & valid[c];
                    const unsigned m = __ballot_sync(0xffffffffu, pass);
                    if (c > r && pass) tacc[c] |= kbit;
                    if (lane == c) keep = m;
                }
                if (lane == r) keep &= ~kbit;                      // not with itself
                if (lane >= r && lane < L) w.u.M[i][lane] = keep;  // the row's own chunk and everything right of it but the last
            }
            // the transposed bits go to words no row stores directly (zeroed before): word r of the rows right of chunk r
#pragma unroll
            for (int c = r + 1; c < L; ++c)
                if (tacc[c]) atomicOr(&w.u.M[32 * c + lane][r], tacc[c]);
        }
        if (32 * L < num) {  // the last chunk's rows: every column chunk; transposed bits into word L of the other rows
            uint32_t tacc[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c) tacc[c] = 0u;
            for (int i = 32 * L + warp; i < num; i += NW) {
                const uint32_t kbit = 1u << (i & 31);
                const float4 row = w.cf[i];
                uint32_t keep = 0u;
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const float4 col = LOOKUP ? w.cf[lane + 32 * c] : colreg[LOOKUP ? 0 : c];
                    const bool pass = pair_may_connect<LOOKUP>(row, col) & valid[c];
                    const unsigned m = __ballot_sync(0xffffffffu, pass);
                    if (c < L && pass) tacc[c] |= kbit;
                    if (lane == c) keep = m;
                }
                if (lane == L) keep &= ~kbit;
                if (lane < NCH) w.u.M[i][lane] = keep;
            }
#pragma unroll
            for (int c = 0; c < L; ++c)
                if (tacc[c]) atomicOr(&w.u.M[32 * c + lane][L], tacc[c]);
        }
    }
    __syncthreads();
    // ---- CSR row extents from the bit rows; `up` counts the entries right of the diagonal (pairs a < b) ----
    {
        int len = 0, up = 0;
        if (tid < num) {
            const int r = tid >> 5, k = tid & 31;
            const uint32_t above = (k == 31) ? 0u : ~((2u << k) - 1u);
#pragma unroll
            for (int c = 0; c < G::NCH; ++c) {
                const uint32_t word = w.u.M[tid][c];
                len += __popc(word);
                up += __popc(c > r ? word : (c == r ? (word & above) : 0u));
            }
        }
        const int mine = len | (up << 16);
        int inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) w.wsum[warp] = inc;
        __syncthreads();
        int before = 0;
#pragma unroll
        for (int k = 0; k < NW; ++k) before += (k < warp) ? w.wsum[k] : 0;
        const int excl = before + inc - mine;
        if (tid < P2) {
            w.row_start[tid] = (unsigned short)(excl & 0xffff);
            w.up_pos[tid] = (unsigned short)((excl & 0xffff) + len - up);  // first CSR slot right of the diagonal
            if (tid < num) w.row_len[tid] = (unsigned short)len;
            w.up_start[tid] = tid < num ? (unsigned short)(excl >> 16) : (unsigned short)0xffffu;
        }
        if (tid == NT - 1) {
            w.npairs = (before + inc) >> 16;
            if (((before + inc) & 0xffff) > G::CAP) w.overflow = 1;
        }
    }
    __syncthreads();
    if (w.overflow) return false;
    // ---- exact entries, once per pair a < b, stored at both positions.  Pair e of the flat enumeration is found by a
    //      binary search over the rows' `up` prefix sums and a bit select inside the row; its two CSR slots follow from
    //      popcounts.  (An exact entry can still be 0: a zero entry adds +0 to the fp32 accumulators below and fails the
    //      >= 1e-5 test of the greedy pass, like an absent one.) ----
    {
        const int np = w.npairs;
        for (int e = tid; e < np; e += NT) {
            int a = 0;
#pragma unroll
            for (int step = P2 / 2; step >= 1; step >>= 1)
                if (w.up_start[a + step] <= e) a += step;
            const int k0 = e - w.up_start[a];
            const int r = a >> 5, ka = a & 31;
            const uint32_t above = (ka == 31) ? 0u : ~((2u << ka) - 1u);
            int k = k0, cw = 0;
            uint32_t word = 0u;
            bool found = false;
#pragma unroll
            for (int c = 0; c < G::NCH; ++c) {
                const uint32_t x0 = w.u.M[a][c];
                const uint32_t x = c > r ? x0 : (c == r ? (x0 & above) : 0u);
                const int cnt = __popc(x);
                const bool here = !found & (k < cnt);
                if (here) {
                    word = x;
                    cw = c;
                }
                found |= here;
                if (!found) k -= cnt;
            }
            int pos = 0;  // position of the k-th set bit of `word`
#pragma unroll
            for (int sh = 16; sh >= 1; sh >>= 1) {
                const int cnt = __popc((word >> pos) & ((1u << sh) - 1u));
                if (k >= cnt) {
                    k -= cnt;
                    pos += sh;
                }
            }
            const int b = 32 * cw + pos;
            const float h = pair_h<LOOKUP>(w.cf[a], w.cf[b], table);
            const int pa = w.up_pos[a] + k0;
            const int pb = w.row_start[b] + bits_below<G::NCH>(w.u.M[b], a);
            w.vals[pa] = h;
            w.cols[pa] = (unsigned char)b;
            w.vals[pb] = h;
            w.cols[pb] = (unsigned char)a;
        }
    }
    __syncthreads();  // the bit matrix is dead: its storage becomes the iteration vectors
    if (tid < num) w.u.it.b[tid] = w.v[tid];
    __syncthreads();

    // ---- power iteration: c = H b (non-zeros, ascending k), b = c * (float)(1 / (sum c + 1e-5)) ----
    for (int it = 0; it < ITERS; ++it) {
        if (tid < num) {
            const int rs = w.row_start[tid], len = w.row_len[tid];
            float acc = 0.0f;
#pragma unroll 2
            for (int e = 0; e < len; ++e) acc = f_add(acc, f_mul(w.vals[rs + e], w.u.it.b[w.cols[rs + e]]));
            w.u.it.c[tid] = acc;
        }
        __syncthreads();
        if (tid == 0) {  // c.sum() in index order
            const float4* c4 = reinterpret_cast<const float4*>(w.u.it.c);
            float4 v = c4[0];
            float sum = v.x;
            if (num >= 4) {
                sum = f_add(f_add(f_add(sum, v.y), v.z), v.w);
                int i = 4;
#pragma unroll 4
                for (; i + 4 <= num; i += 4) {
                    v = c4[i >> 2];
                    sum = f_add(f_add(f_add(f_add(sum, v.x), v.y), v.z), v.w);
                }
#pragma unroll 1
                for (; i < num; ++i) sum = f_add(sum, w.u.it.c[i]);
            } else {
                for (int i = 1; i < num; ++i) sum = f_add(sum, w.u.it.c[i]);
            }
            w.f = (float)(1.0 / ((double)sum + 0.00001));
        }
        __syncthreads();
        if (tid < num) w.u.it.b[tid] = f_mul(w.u.it.c[tid], w.f);
        __syncthreads();
    }

    // ---- order by b descending: rank in the order (value desc, index asc) by one bucket pass.  std::sort's own
    //      permutation differs from it only inside groups of equal values; the greedy pass below asks for it when -
    //      and only when - such a group can change its outcome ----
    {
        static_assert(P2 <= NT && G::MAXN <= P2, "one key per thread");
        uint32_t key[1] = {0u};
        if (tid < num) {
            key[0] = __float_as_uint(w.u.it.b[tid]);  // b >= 0: the bit pattern orders like the value
            if (key[0] == 0x80000000u) key[0] = 0u;
        }
        int rank[1], first[1];
        bool tied[1];
        block_rank_desc<NT, 1, P2>(key, num, w.u.it.start, w.u.it.buck, reinterpret_cast<uint32_t*>(w.wsum), rank, first, tied);
        if (tid < num) w.u.it.y[rank[0]] = (unsigned short)tid;
    }
    __syncthreads();
    if (warp != 0) return true;  // the rest runs in warp 0 only; thread 0 carries the result

    // ---- greedy selection (matcher.cpp:1305-1345): lanes own sorted positions p = lane + 32 c.  A candidate is
    //      accepted when it is the first one in sorted order that is still compatible with everything accepted so far
    //      ("open").  With the open set given, that is the open candidate of largest value whatever the order inside
    //      groups of equal values - unless another open candidate has the same value.  Only then does std::sort's
    //      permutation matter (which of the two is accepted first, and in which order they enter the next stage): lane 0
    //      replays the introsort and the pass starts over on the exact order.  Up to 16 elements std::sort is a stable
    //      insertion sort and the total order is exact. ----
    int n2 = 0;
    {
        unsigned short ind[CH];
        float bv[CH];
        bool exact = num <= 16;
        for (;;) {
            unsigned open = 0;
            n2 = 0;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const int p = lane + 32 * c;
                ind[c] = (p < num) ? w.u.it.y[p] : 0;
                bv[c] = (p < num) ? w.u.it.b[ind[c]] : 0.0f;
                if (p < num && !((double)bv[c] < 0.0001)) open |= 1u << c;
            }
#pragma unroll 1
            for (int c = lane; c < P2; c += 32) w.mark[c] = 0;
            __syncwarp();
            bool tie_hit = false;
            for (;;) {
                int pos = -1;
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const unsigned m = __ballot_sync(0xffffffffu, (open >> c) & 1u);
                    if (m && pos < 0) pos = 32 * c + __ffs(m) - 1;
                }
                if (pos < 0) break;
                const int s = w.u.it.y[pos];
                if (!exact) {
                    const float vs = w.u.it.b[s];
                    bool t = false;
#pragma unroll
                    for (int c = 0; c < CH; ++c) t |= ((open >> c) & 1u) && lane + 32 * c != pos && bv[c] == vs;
                    if (__any_sync(0xffffffffu, t)) {
                        tie_hit = true;
                        break;
                    }
                }
                if (lane == 0) w.u.it.sel[n2] = (unsigned short)s;
                ++n2;
                if (n2 > 32) break;  // more survivors than the orientation stage holds: dense kernel
                // H is symmetric: the candidates k with H[k][s] >= 1e-5 (an absent entry is 0) are the entries of row s
                {
                    const int rs = w.row_start[s], len = w.row_len[s];
                    for (int e = lane; e < len; e += 32)
                        if (!((double)w.vals[rs + e] < 0.00001)) w.mark[w.cols[rs + e]] = (unsigned char)n2;
                }
                __syncwarp();
                const unsigned short sli = w.li[s], srj = w.rj[s];
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    if (!((open >> c) & 1u)) continue;
                    const int p = lane + 32 * c;
                    const bool keep = p != pos && w.li[ind[c]] != sli && w.rj[ind[c]] != srj && w.mark[ind[c]] == (unsigned char)n2;
                    if (!keep) open &= ~(1u << c);
                }
                __syncwarp();
            }
            if (!tie_hit) break;
            __syncwarp();  // every lane is done with the provisional order
            if (lane == 0) std_sort_desc_emulate<float, unsigned short>(w.u.it.b, w.u.it.y, num);
            __syncwarp();
            exact = true;
        }
    }
    __syncwarp();
    if (n2 > 32) {
        if (lane == 0) w.overflow = 1;
        return false;
    }
    if (n2 == 0) return true;

    // ---- orientation-consistency graph on the n2 survivors, lane = survivor (acceptance order) ----
    const int me = (lane < n2) ? w.u.it.sel[lane] : w.u.it.sel[0];
    const float mv = w.v[me];
    const unsigned short mli = w.li[me], mrj = w.rj[me];
    // the n2 (n2 - 1) / 2 pairs i < j are spread over the lanes (a pair costs two fdlibm atan2f and six angle
    // reductions: one round of 32 pairs instead of one round per survivor)
    w.omask[lane] = 0u;
    __syncwarp();
    {
        const int npair = n2 * (n2 - 1) / 2;
        for (int e0 = 0; e0 < npair; e0 += 32) {
            const int e = e0 + lane;
            if (e < npair) {
                int j = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)e)) * 0.5f);  // pairs (., j) start at e = j (j - 1) / 2
                while (j * (j - 1) / 2 > e) --j;
                while ((j + 1) * j / 2 <= e) ++j;
                const int i = e - j * (j - 1) / 2;
                const int ci = w.u.it.sel[i], cj = w.u.it.sel[j];
                // orientations of the two correspondences' minutiae / texture points (lat_ori, gal_ori: this job's templates)
                if (angle_compatible(w.cf[ci], w.cf[cj], __ldg(lat_ori + w.li[ci]), __ldg(lat_ori + w.li[cj]), __ldg(gal_ori + w.rj[ci]),
                                     __ldg(gal_ori + w.rj[cj]))) {
                    atomicOr(&w.omask[i], 1u << j);
                    atomicOr(&w.omask[j], 1u << i);
                }
            }
        }
    }
    __syncwarp();
    const unsigned mask = (lane < n2) ? w.omask[lane] : 0u;
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
        const int best = (int)__reduce_min_sync(0xffffffffu, open ? (unsigned)rank : 0x7fffffffu);
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

// one (latent, template, slot) job; a job the working set cannot hold is appended to `ov`
template <int TIER>
__device__ __forceinline__ void graph_minu_job(SparseWork<false, TIER>& w, const GraphMinuParams& P, size_t oidx, int q, int tl, int slot,
                                               OverflowList ov) {
    using G = SparseGeom<false, TIER>;
    const int tid = threadIdx.x;
    const int num = P.corr_n[oidx];
    if (tid == 0) {
        w.overflow = 0;
        w.tie = 0;
        w.npairs = 0;
    }
#pragma unroll 1
    for (int e = tid; e < G::MAXP * G::NCH; e += G::NT) (&w.u.M[0][0])[e] = 0u;
    int big = 0;
    if (tid < num) {
        const uint32_t ij = P.corr_ij[oidx * kTopCorrMinu + tid];
        const int i = (int)(ij >> 16), j = (int)(ij & 0xffffu);
        w.v[tid] = P.corr_v[oidx * kTopCorrMinu + tid];
        w.li[tid] = (unsigned short)i;
        w.rj[tid] = (unsigned short)j;
        const uint32_t lo = P.slot_off[q * 3 + slot] + i, go = P.minu_off[P.g0 + tl] + j;
        const short2 lxy = P.lat_xy[lo], rxy = P.gal_xy[go];
        w.cf[tid] = make_float4((float)lxy.x, (float)rxy.x, (float)lxy.y, (float)rxy.y);
        // the pre-test squares coordinate differences in fp32: exact only below 2048 px
        // (coordinates in [0, 2048): differences below 2048, squared distances below 2^23)
        big = ((unsigned)(int)lxy.x | (unsigned)(int)lxy.y | (unsigned)(int)rxy.x | (unsigned)(int)rxy.y) >= 2048u;
    } else if (tid < G::MAXP) {
        w.cf[tid] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    if (__syncthreads_or(big)) {  // larger images: the dense kernel evaluates every entry exactly
        if (tid == 0) ov.jobs[atomicAdd(ov.count, 1)] = (int)oidx;
        return;
    }
    float score;
    const bool ok = sparse_cascade<false>(w, num, nullptr, P.lat_ori + P.slot_off[q * 3 + slot], P.gal_ori + P.minu_off[P.g0 + tl], &score);
    if (tid == 0) {
        if (ok) P.comp[((size_t)q * P.G + P.g0 + tl) * 4 + slot] = score;
        else ov.jobs[atomicAdd(ov.count, 1)] = (int)oidx;
    }
}

__global__ void __launch_bounds__(SparseGeom<false>::NT) graph_minu_sparse_kernel(GraphMinuParams P, OverflowList ov) {
    extern __shared__ __align__(128) unsigned char smem[];
    SparseWork<false>& w = *reinterpret_cast<SparseWork<false>*>(smem);
    // grid = (3 * n_chunk, latents)
    const int q = job_latent();
    if (q >= P.Q) return;
    const int tl = (int)(blockIdx.x / 3u), slot = (int)(blockIdx.x - 3u * (unsigned)tl);
    const size_t oidx = (size_t)q * gridDim.x + blockIdx.x;  // (q * n_chunk + tl) * 3 + slot
    graph_minu_job<0>(w, P, oidx, q, tl, slot, ov);
}

// Second chance for the jobs whose graph overflowed the 2,560 non-zeros of the kernel above (prints with clustered
// minutiae: SURVEY 8d's i.i.d. gallery has ~1,300): the same cascade over a CSR of 6,656 non-zeros at 5 CTAs per SM,
// ~4 x cheaper than the dense kernel, which remains for what is left (mated pairs, large images, > 32 survivors).
__global__ void __launch_bounds__(SparseGeom<false, 1>::NT) graph_minu_mid_kernel(GraphMinuParams P, const int* in_count,
                                                                                  const int* in_jobs, OverflowList ov,
                                                                                  unsigned long long* mid_jobs_total) {
    extern __shared__ __align__(128) unsigned char smem[];
    SparseWork<false, 1>& w = *reinterpret_cast<SparseWork<false, 1>*>(smem);
    const int n_jobs = *in_count;
    if (blockIdx.x == 0 && threadIdx.x == 0 && mid_jobs_total) atomicAdd(mid_jobs_total, (unsigned long long)n_jobs);
    for (int jb = blockIdx.x; jb < n_jobs; jb += gridDim.x) {
        const size_t oidx = (size_t)in_jobs[jb];  // (q * n_chunk + tl) * 3 + slot
        const int slot = (int)(oidx % 3);
        const size_t pair = oidx / 3;
        const int tl = (int)(pair % P.n_chunk), q = (int)(pair / P.n_chunk);
        __syncthreads();  // the previous job's working set is no longer in use
        graph_minu_job<1>(w, P, oidx, q, tl, slot, ov);
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
    // grid = (n_chunk, latents)
    const int q = job_latent();
    if (q >= P.Q) return;
    const int tl = (int)blockIdx.x;
    const size_t pair = (size_t)q * gridDim.x + blockIdx.x;  // q * n_chunk + tl
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
        w.npairs = 0;
    }
    for (int e = tid; e < SparseGeom<true>::MAXP * SparseGeom<true>::NCH; e += NT) (&w.u.M[0][0])[e] = 0u;
    const size_t rbase = pair * (size_t)P.lt_stride;
    int num;
    if (nLt > kTopCorrTex) {
        // ---- K3b: the 200 best rows in std::sort order (matcher.cpp:736-749): rank of every row maximum in the order
        //      (value desc, row asc) by one bucket pass, in the (still unused) CSR value area ----
        constexpr int NB = 512, KPT = 1024 / NT;
        int* start = reinterpret_cast<int*>(w.vals);                        // [NB + 1] (+ padding to 16 bytes)
        uint2* buck = reinterpret_cast<uint2*>(start + NB + 4);             // [1024]
        float* rv = reinterpret_cast<float*>(buck + 1024);                  // [1024]
        unsigned short* ry = reinterpret_cast<unsigned short*>(start);      // [1024], replay only: the buckets are dead then
        static_assert(sizeof(int) * (NB + 4) + sizeof(uint2) * 1024 + sizeof(float) * 1024 <= sizeof(w.vals), "K3b scratch");
        uint32_t key[KPT];
        float val[KPT];
#pragma unroll
        for (int e = 0; e < KPT; ++e) {
            const int i = tid + e * NT;
            key[e] = 0u;
            val[e] = 0.0f;
            if (i < nLt) {
                float x = P.rowmax_val[rbase + i];
                val[e] = x;
                rv[i] = x;
                if (x == 0.0f) x = 0.0f;  // -0 and +0 compare equal in the reference's comparator
                const uint32_t u = __float_as_uint(x);
                key[e] = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
            }
        }
        int rank[KPT], first[KPT];
        bool tied[KPT];
        block_rank_desc<NT, KPT, NB>(key, nLt, start, buck, reinterpret_cast<uint32_t*>(w.wsum), rank, first, tied);
        num = kTopCorrTex;
#pragma unroll
        for (int e = 0; e < KPT; ++e) {
            const int i = tid + e * NT;
            if (i < nLt) {
                if (rank[e] < kTopCorrTex) {
                    w.v[rank[e]] = val[e];
                    w.li[rank[e]] = (unsigned short)i;
                    w.rj[rank[e]] = P.rowmax_j[rbase + i];
                }
                // a tie that reaches into the first 200 positions makes the permutation introsort-specific
                if (tied[e] && first[e] < kTopCorrTex) w.tie = 1;
            }
        }
        __syncthreads();
        const bool replay = w.tie != 0;
        __syncthreads();
        if (replay) {
            if (tid < 32) {  // warp-cooperative replay of the introsort (stdsort_emul.h)
                warp_std_sort_desc_prefix(DenseKey<float>{rv}, ry, nLt, kTopCorrTex);
                if (tid == 0) {
                    atomicAdd(P.slow_path_count, 1ull);
                    w.tie = 0;
                }
            }
            __syncthreads();
            if (tid < num) {
                const int i = ry[tid];
                w.v[tid] = rv[i];
                w.li[tid] = (unsigned short)i;
                w.rj[tid] = P.rowmax_j[rbase + i];
            }
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
        const short2 lxy = P.lat_xy[(size_t)q * P.lt_stride + i], rxy = P.gal_xy[gbase + j];
        w.cf[tid] = make_float4((float)lxy.x, (float)rxy.x, (float)lxy.y, (float)rxy.y);
    } else if (tid < SparseGeom<true>::MAXP) {
        w.cf[tid] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    __syncthreads();
    float score;
    const bool ok = sparse_cascade<true>(w, num, P.table, P.lat_ori + (size_t)q * P.lt_stride, P.gal_ori + gbase, &score);
    if (tid == 0) {
        if (ok) P.comp[((size_t)q * P.G + g) * 4 + 3] = score;
        else ov.jobs[atomicAdd(ov.count, 1)] = (int)pair;
    }
}

}  // namespace lafis
