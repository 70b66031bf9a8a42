// Reproduces the permutation libstdc++'s std::sort produces for the reference's index sorts.
//
// The reference ranks candidates with `std::sort(y.begin(), y.end(), [&](int a,int b){return
// key[a] > key[b];})` (matching/matcher.cpp:476, :741, :1301, :1423, :1590).  std::sort is not
// stable and LSS_R_Fast2 (:1556-1581) creates exact ties by construction, so WHICH of two tied
// correspondences comes first decides which one survives the greedy selection.  The kernels sort in
// parallel with the total order (key desc, index asc); whenever a tie can influence the result they
// fall back to this routine, which walks through the introsort of GCC 13's bits/stl_algo.h
// (median-of-3 to first, unguarded Hoare partition, recursion on the right part, depth limit
// 2*floor(log2 n) with heap-sort fallback, final insertion sort with threshold 16) step for step.
// For n <= 16 std::sort is a plain insertion sort, i.e. stable, and the total order is already
// exact.
//
// Prefix replay.  Every caller only consumes the first `need` positions of the sorted sequence
// (top-120 / top-200 / the greedy pass).  After a partition step the right part [cut, last) holds
// keys <= pivot <= every key of the left part, and neither the later partition steps nor the final
// insertion pass (which moves an element left only past strictly smaller keys) ever carry an
// element across `cut`.  So partitions that start at or beyond `need` cannot influence positions
// [0, need) and are skipped; the replay costs O(n) instead of O(n log n).
//
// Host+device so that tests/test_host_math.py can pin it against std::sort on the CPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LAFIS_SORT_HD __host__ __device__
#else
#define LAFIS_SORT_HD
#endif

namespace lafis {

// Closed form of the unguarded Hoare partition (see BlockStdSortEmu below for the derivation).  Ls: positions
// that stop the up-scan, ascending; Rs: positions that stop the down-scan, ASCENDING (the scan meets them back to
// front).  Returns the cut; *m_out = number of swaps, swap k exchanges Ls[k] and Rs[NR-1-k].
template <typename IdxT>
LAFIS_SORT_HD inline int hoare_closed_form(const IdxT* Ls, int NL, const IdxT* Rs, int NR, int last, int* m_out) {
    int lo = 0, hi = NL < NR ? NL : NR;  // largest m with Ls[m-1] < Rs[NR-m]
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((int)Ls[mid - 1] < (int)Rs[NR - mid]) lo = mid;
        else hi = mid - 1;
    }
    int cut = lo < NL ? (int)Ls[lo] : last;
    if (lo > 0 && (int)Rs[NR - lo] < cut) cut = (int)Rs[NR - lo];
    *m_out = lo;
    return cut;
}

// KeyFn: key(index) -> comparable value.  IdxT: integer type of the index array.
template <typename KeyFn, typename IdxT>
struct StdSortEmu {
    KeyFn key;
    IdxT* y;

    LAFIS_SORT_HD bool before(IdxT a, IdxT b) const { return key((int)a) > key((int)b); }
    LAFIS_SORT_HD void swp(int a, int b) {
        IdxT t = y[a];
        y[a] = y[b];
        y[b] = t;
    }
    LAFIS_SORT_HD void unguarded_linear_insert(int last) {
        IdxT val = y[last];
        int next = last - 1;
        while (before(val, y[next])) {
            y[last] = y[next];
            last = next;
            --next;
        }
        y[last] = val;
    }
    LAFIS_SORT_HD void insertion_sort(int first, int last) {
        if (first == last) return;
        for (int i = first + 1; i != last; ++i) {
            if (before(y[i], y[first])) {
                IdxT val = y[i];
                for (int k = i; k > first; --k) y[k] = y[k - 1];
                y[first] = val;
            } else {
                unguarded_linear_insert(i);
            }
        }
    }
    LAFIS_SORT_HD void push_heap(int first, int hole, int top, IdxT value) {
        int parent = (hole - 1) / 2;
        while (hole > top && before(y[first + parent], value)) {
            y[first + hole] = y[first + parent];
            hole = parent;
            parent = (hole - 1) / 2;
        }
        y[first + hole] = value;
    }
    LAFIS_SORT_HD void adjust_heap(int first, int hole, int len, IdxT value) {
        const int top = hole;
        int child = hole;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (before(y[first + child], y[first + child - 1])) child--;
            y[first + hole] = y[first + child];
            hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            y[first + hole] = y[first + child - 1];
            hole = child - 1;
        }
        push_heap(first, hole, top, value);
    }
    LAFIS_SORT_HD void heap_sort(int first, int last) {
        const int len = last - first;
        if (len >= 2) {
            int parent = (len - 2) / 2;
            for (;;) {
                adjust_heap(first, parent, len, y[first + parent]);
                if (parent == 0) break;
                parent--;
            }
        }
        while (last - first > 1) {
            --last;
            IdxT value = y[last];
            y[last] = y[first];
            adjust_heap(first, 0, last - first, value);
        }
    }
    // one partition step on [first, last): returns the cut
    LAFIS_SORT_HD int partition_pivot(int first, int last) {
        const int mid = first + (last - first) / 2;
        const int a = first + 1, b = mid, c = last - 1;
        if (before(y[a], y[b])) {
            if (before(y[b], y[c])) swp(first, b);
            else if (before(y[a], y[c])) swp(first, c);
            else swp(first, a);
        } else if (before(y[a], y[c])) swp(first, a);
        else if (before(y[b], y[c])) swp(first, c);
        else swp(first, b);
        int lo = first + 1, hi = last;
        for (;;) {
            while (before(y[lo], y[first])) ++lo;
            --hi;
            while (before(y[first], y[hi])) --hi;
            if (!(lo < hi)) return lo;
            swp(lo, hi);
            ++lo;
        }
    }
    // the same step through the closed form (sequential list building; pins the formula on the CPU)
    LAFIS_SORT_HD int partition_pivot_closed(int first, int last, IdxT* Ls, IdxT* Rs) {
        const int mid = first + (last - first) / 2;
        const int a = first + 1, b = mid, c = last - 1;
        if (before(y[a], y[b])) {
            if (before(y[b], y[c])) swp(first, b);
            else if (before(y[a], y[c])) swp(first, c);
            else swp(first, a);
        } else if (before(y[a], y[c])) swp(first, a);
        else if (before(y[b], y[c])) swp(first, c);
        else swp(first, b);
        int NL = 0, NR = 0;
        for (int p = first; p < last; ++p) {
            if (p > first && !before(y[p], y[first])) Ls[NL++] = (IdxT)p;
            if (!before(y[first], y[p])) Rs[NR++] = (IdxT)p;
        }
        int m;
        const int cut = hoare_closed_form(Ls, NL, Rs, NR, last, &m);
        for (int k = 0; k < m; ++k) swp((int)Ls[k], (int)Rs[NR - 1 - k]);
        return cut;
    }
    // y must hold 0..n-1 on entry (std::iota).  On return positions [0, min(need, n)) are what
    // std::sort leaves there; later positions are unspecified.
    LAFIS_SORT_HD void sort_prefix(int n, int need, IdxT* Ls = nullptr, IdxT* Rs = nullptr) {
        if (n <= 0) return;
        if (need > n) need = n;
        int lg = 0;
        for (int m = n; m > 1; m >>= 1) ++lg;
        // __introsort_loop recurses into [cut,last) and continues on [first,cut); the two parts are
        // disjoint, so an explicit stack of deferred parts visits them with identical contents.
        struct Frame {
            int first, last, depth;
        };
        Frame stack[72];
        int sp = 0;
        int done_to = 0;  // positions [0, done_to) belong to fully partitioned (or heap-sorted) ranges
        stack[sp++] = Frame{0, n, 2 * lg};
        while (sp > 0) {
            Frame f = stack[--sp];
            if (f.first >= need) continue;
            while (f.last - f.first > 16) {
                if (f.depth == 0) {
                    heap_sort(f.first, f.last);
                    break;
                }
                --f.depth;
                const int cut = Ls ? partition_pivot_closed(f.first, f.last, Ls, Rs) : partition_pivot(f.first, f.last);
                if (cut < need) stack[sp++] = Frame{cut, f.last, f.depth};  // the recursive call
                f.last = cut;                                                // the loop's continuation
            }
            if (f.last > done_to) done_to = f.last;
        }
        if (n > 16) {
            // the leaf ranges with first < need tile [0, E) contiguously, so E >= need
            int E = done_to;
            if (E < 16) E = 16;
            if (E > n) E = n;
            insertion_sort(0, 16);
            for (int i = 16; i < E; ++i) unguarded_linear_insert(i);
        } else {
            insertion_sort(0, n);
        }
    }
};

template <typename KeyT>
struct DenseKey {
    const KeyT* k;
    LAFIS_SORT_HD KeyT operator()(int i) const { return k[i]; }
};

// full permutation (need = n)
template <typename KeyT, typename IdxT>
LAFIS_SORT_HD inline void std_sort_desc_emulate(const KeyT* key, IdxT* y, int n) {
    for (int i = 0; i < n; ++i) y[i] = (IdxT)i;
    StdSortEmu<DenseKey<KeyT>, IdxT> s{DenseKey<KeyT>{key}, y};
    s.sort_prefix(n, n);
}

// first `need` positions only
template <typename KeyFn, typename IdxT>
LAFIS_SORT_HD inline void std_sort_desc_prefix(KeyFn key, IdxT* y, int n, int need) {
    for (int i = 0; i < n; ++i) y[i] = (IdxT)i;
    StdSortEmu<KeyFn, IdxT> s{key, y};
    s.sort_prefix(n, need);
}


#if defined(__CUDACC__)
// Warp-cooperative replay: the same algorithm, executed by all 32 lanes of a warp in lock step (every
// lane sees the same shared-memory contents and takes the same branches).  The two scanning loops of the
// Hoare partition - where almost all of the O(n) work is - examine 32 positions per step with a ballot;
// everything that writes is done by lane 0 between __syncwarp()s.  `y` must be in shared memory.
template <typename KeyFn, typename IdxT>
struct WarpStdSortEmu {
    KeyFn key;
    IdxT* y;
    int lane;

    __device__ bool before(IdxT a, IdxT b) const { return key((int)a) > key((int)b); }
    __device__ void swp(int a, int b) {
        __syncwarp();  // every lane has read what it decided this swap on
        if (lane == 0) {
            IdxT t = y[a];
            y[a] = y[b];
            y[b] = t;
        }
        __syncwarp();
    }
    // first p >= lo with !before(y[p], piv); std::sort's pivot choice guarantees one below `limit`
    __device__ int scan_up(int lo, IdxT piv, int limit) {
        for (;;) {
            const int p = lo + lane;
            const bool stop = (p >= limit) || !before(y[p], piv);
            const unsigned m = __ballot_sync(0xffffffffu, stop);
            if (m) return lo + __ffs(m) - 1;
            lo += 32;
        }
    }
    // first p <= hi (descending) with !before(piv, y[p])
    __device__ int scan_down(int hi, IdxT piv, int limit) {
        for (;;) {
            const int p = hi - lane;
            const bool stop = (p < limit) || !before(piv, y[p]);
            const unsigned m = __ballot_sync(0xffffffffu, stop);
            if (m) return hi - (__ffs(m) - 1);
            hi -= 32;
        }
    }
    __device__ int partition_pivot(int first, int last) {
        const int mid = first + (last - first) / 2;
        const int a = first + 1, b = mid, c = last - 1;
        if (before(y[a], y[b])) {
            if (before(y[b], y[c])) swp(first, b);
            else if (before(y[a], y[c])) swp(first, c);
            else swp(first, a);
        } else if (before(y[a], y[c])) swp(first, a);
        else if (before(y[b], y[c])) swp(first, c);
        else swp(first, b);
        const IdxT piv = y[first];
        int lo = first + 1, hi = last;
        for (;;) {
            lo = scan_up(lo, piv, last);
            --hi;
            hi = scan_down(hi, piv, first);
            if (!(lo < hi)) return lo;
            swp(lo, hi);
            ++lo;
        }
    }
    __device__ void sort_prefix(int n, int need) {
        if (n <= 0) return;
        if (need > n) need = n;
        int lg = 0;
        for (int m = n; m > 1; m >>= 1) ++lg;
        StdSortEmu<KeyFn, IdxT> seq{key, y};  // heap sort / insertion sort: lane 0 only
        struct Frame {
            int first, last, depth;
        };
        Frame stack[72];
        int sp = 0, done_to = 0;
        stack[sp++] = Frame{0, n, 2 * lg};
        while (sp > 0) {
            Frame f = stack[--sp];
            if (f.first >= need) continue;
            while (f.last - f.first > 16) {
                if (f.depth == 0) {
                    __syncwarp();
                    if (lane == 0) seq.heap_sort(f.first, f.last);
                    __syncwarp();
                    break;
                }
                --f.depth;
                const int cut = partition_pivot(f.first, f.last);
                if (cut < need) stack[sp++] = Frame{cut, f.last, f.depth};
                f.last = cut;
            }
            if (f.last > done_to) done_to = f.last;
        }
        __syncwarp();
        if (lane == 0) {
            if (n > 16) {
                int E = done_to;
                if (E < 16) E = 16;
                if (E > n) E = n;
                seq.insertion_sort(0, 16);
                for (int i = 16; i < E; ++i) seq.unguarded_linear_insert(i);
            } else {
                seq.insertion_sort(0, n);
            }
        }
        __syncwarp();
    }
};

// first `need` positions, whole warp cooperating; y in shared memory, every lane passes the same arguments
template <typename KeyFn, typename IdxT>
__device__ inline void warp_std_sort_desc_prefix(KeyFn key, IdxT* y, int n, int need) {
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < n; i += 32) y[i] = (IdxT)i;
    __syncwarp();
    WarpStdSortEmu<KeyFn, IdxT> s{key, y, lane};
    s.sort_prefix(n, need);
}

// Block-cooperative replay for large index sets (the 9,600 normalised similarities of a minutiae match).
// With random keys each scanning loop of the Hoare partition stops after ~2 elements, so the warp version
// above spends its time on ~n/4 sequential swaps.  The partition's outcome has a closed form:
//   L = positions p in (first, last) where the up-scan stops,   key(y[p]) <= key(pivot), ascending;
//   R = positions p in [first, last) where the down-scan stops, key(y[p]) >= key(pivot), descending
//       (position `first` holds the pivot and always stops the down-scan).
// Until the scans cross they only ever read positions the earlier swaps have not touched, so swap k pairs
// L[k] with R[k] exactly while L[k] < R[k] (a prefix property: L ascends, R descends).  After the m swaps
// that happen, the up-scan continues over untouched positions until L[m] or until R[m-1], which now holds
// a value that stops it; the down-scan likewise ends at max(R[m], L[m-1]) <= that position, so the loop
// returns cut = min(L[m], R[m-1]).  Both lists come from one ballot/prefix-sum pass over the segment and
// the m swaps are independent: O(n / NT) per partition instead of O(n).
// `Ls` / `Rs`: scratch of n entries each; `sh`: 2 * NT/32 + 2 ints; all in shared memory.  Every thread of
// the block calls with the same arguments.
template <typename KeyFn, typename IdxT, int NT>
struct BlockStdSortEmu {
    static constexpr int NW = NT / 32;
    KeyFn key;
    IdxT* y;
    IdxT* Ls;
    IdxT* Rs;
    int* sh;

    __device__ int partition_pivot(int first, int last) {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        if (tid == 0) {
            StdSortEmu<KeyFn, IdxT> seq{key, y};
            const int mid = first + (last - first) / 2;
            const int a = first + 1, b = mid, c = last - 1;
            if (seq.before(y[a], y[b])) {
                if (seq.before(y[b], y[c])) seq.swp(first, b);
                else if (seq.before(y[a], y[c])) seq.swp(first, c);
                else seq.swp(first, a);
            } else if (seq.before(y[a], y[c])) seq.swp(first, a);
            else if (seq.before(y[b], y[c])) seq.swp(first, c);
            else seq.swp(first, b);
        }
        __syncthreads();
        const auto pk = key((int)y[first]);
        const unsigned lt_mask = (1u << lane) - 1u;
        int baseL = 0, baseR = 0;
        for (int p0 = first; p0 < last; p0 += NT) {
            const int p = p0 + tid;
            bool isL = false, isR = false;
            if (p < last) {
                const auto kp = key((int)y[p]);
                isL = p > first && !(kp > pk);
                isR = !(pk > kp);
            }
            const unsigned mL = __ballot_sync(0xffffffffu, isL), mR = __ballot_sync(0xffffffffu, isR);
            if (lane == 0) {
                sh[warp] = __popc(mL);
                sh[NW + warp] = __popc(mR);
            }
            __syncthreads();
            int offL = baseL, offR = baseR;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const int cL = sh[w], cR = sh[NW + w];
                if (w < warp) {
                    offL += cL;
                    offR += cR;
                }
                baseL += cL;
                baseR += cR;
            }
            if (isL) Ls[offL + __popc(mL & lt_mask)] = (IdxT)p;
            if (isR) Rs[offR + __popc(mR & lt_mask)] = (IdxT)p;  // ascending here, read back to front
            __syncthreads();
        }
        const int NL = baseL, NR = baseR;
        if (tid == 0) {
            int m0;
            sh[2 * NW + 1] = hoare_closed_form(Ls, NL, Rs, NR, last, &m0);
            sh[2 * NW] = m0;
        }
        __syncthreads();
        const int m = sh[2 * NW], cut = sh[2 * NW + 1];
        for (int k = tid; k < m; k += NT) {
            const int a = Ls[k], b = Rs[NR - 1 - k];
            const IdxT t = y[a];
            y[a] = y[b];
            y[b] = t;
        }
        __syncthreads();
        return cut;
    }

    __device__ void sort_prefix(int n, int need) {
        const int tid = threadIdx.x;
        for (int i = tid; i < n; i += NT) y[i] = (IdxT)i;
        __syncthreads();
        if (n <= 0) return;
        if (need > n) need = n;
        int lg = 0;
        for (int m = n; m > 1; m >>= 1) ++lg;
        StdSortEmu<KeyFn, IdxT> seq{key, y};  // small ranges, heap sort, insertion sort: thread 0 only
        struct Frame {
            int first, last, depth;
        };
        Frame stack[72];
        int sp = 0, done_to = 0;
        stack[sp++] = Frame{0, n, 2 * lg};
        while (sp > 0) {
            Frame f = stack[--sp];
            if (f.first >= need) continue;
            while (f.last - f.first > 16) {
                if (f.depth == 0) {
                    if (tid == 0) seq.heap_sort(f.first, f.last);
                    __syncthreads();
                    break;
                }
                --f.depth;
                int cut;
                if (f.last - f.first > 96) {
                    cut = partition_pivot(f.first, f.last);
                } else {
                    if (tid == 0) sh[2 * NW + 1] = seq.partition_pivot(f.first, f.last);
                    __syncthreads();
                    cut = sh[2 * NW + 1];
                    __syncthreads();
                }
                if (cut < need) stack[sp++] = Frame{cut, f.last, f.depth};
                f.last = cut;
            }
            if (f.last > done_to) done_to = f.last;
        }
        if (tid == 0) {
            if (n > 16) {
                int E = done_to;
                if (E < 16) E = 16;
                if (E > n) E = n;
                seq.insertion_sort(0, 16);
                for (int i = 16; i < E; ++i) seq.unguarded_linear_insert(i);
            } else {
                seq.insertion_sort(0, n);
            }
        }
        __syncthreads();
    }
};
#endif

}  // namespace lafis
