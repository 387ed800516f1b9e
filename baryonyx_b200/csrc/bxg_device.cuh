// bxg_device.cuh -- device-side building blocks of the sweep engine.
//
// Deterministic fp64 mode: every floating-point operation that feeds a result
// is written with the round-to-nearest intrinsics (__dadd_rn, __dmul_rn,
// __ddiv_rn), which the compiler never contracts into FMAs, and the file is
// compiled with -fmad=false as a second guard.  The reference side of the
// parity tests is built with -ffp-contract=off for the same reason
// (SURVEY.md 8c).
//
// The random stream and the orderings are those of the C++ standard library
// the reference runs on (libstdc++ 13): std::default_random_engine =
// minstd_rand0 (bits/random.h:1592-1641), bernoulli_distribution on
// generate_canonical<double,53> (random.tcc:3349-3381),
// uniform_int_distribution (uniform_int_dist.h:284-373), std::shuffle
// (stl_algo.h:3738-3803) and std::sort (stl_algo.h:1810-1957).
#ifndef BXG_DEVICE_CUH
#define BXG_DEVICE_CUH

#include <cstdint>
#include <cuda_runtime.h>

namespace bxg {

// ---------------------------------------------------------------------------
// minstd_rand0: s <- 16807 s mod (2^31 - 1)
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t
mod_m31(uint64_t x)
{
    // x < 2^62: two Mersenne folds, one conditional subtract
    x = (x & 0x7fffffffull) + (x >> 31);
    x = (x & 0x7fffffffull) + (x >> 31);
    return x >= 0x7fffffffull ? (uint32_t)(x - 0x7fffffffull) : (uint32_t)x;
}

__host__ __device__ __forceinline__ uint32_t
lcg_mul(uint32_t s, uint32_t a)
{
    return mod_m31((uint64_t)s * a);
}

__host__ __device__ __forceinline__ uint32_t
lcg_next(uint32_t& s)
{
    s = lcg_mul(s, 16807u);
    return s;
}

// 16807^k mod (2^31 - 1)
__host__ __device__ inline uint32_t
lcg_pow(uint64_t k)
{
    uint32_t r = 1u, b = 16807u;
    while (k) {
        if (k & 1)
            r = lcg_mul(r, b);
        b = lcg_mul(b, b);
        k >>= 1;
    }
    return r;
}

// generate_canonical<double,53> from two successive engine outputs
__device__ __forceinline__ double
canonical_from(uint32_t g1, uint32_t g2)
{
    const double r = 2147483646.0;
    const double rr = 2147483646.0 * 2147483646.0; // rounded once, like tmp*=r
    double sum = __dadd_rn((double)(g1 - 1u), __dmul_rn((double)(g2 - 1u), r));
    double ret = __ddiv_rn(sum, rr);
    if (ret >= 1.0)
        ret = 0.99999999999999988897769753748; // nextafter(1.0, 0.0)
    return ret;
}

__device__ __forceinline__ bool
bernoulli(uint32_t& s, double p)
{
    const uint32_t g1 = lcg_next(s);
    const uint32_t g2 = lcg_next(s);
    return canonical_from(g1, g2) < p;
}

// uniform_int_distribution<unsigned long>{0, b} for b < 2147483645
__device__ __forceinline__ uint32_t
uniform_below(uint32_t& s, uint32_t b)
{
    const uint32_t urngrange = 2147483645u;
    const uint32_t uerange = b + 1u;
    const uint32_t scaling = urngrange / uerange;
    const uint32_t past = uerange * scaling;
    uint32_t ret;
    do
        ret = lcg_next(s) - 1u;
    while (ret >= past);
    return ret / scaling;
}

// ---------------------------------------------------------------------------
// Array views for the single-thread algorithms below.  A view has an element
// type E and get / set; Plain<T> is a pointer, Pairs keeps a key array and an
// id array in step (the reduced-cost scratch of a row: sort key + position in
// the row).
// ---------------------------------------------------------------------------
template<typename T>
struct Plain
{
    using E = T;
    T* p;
    __device__ __forceinline__ T get(int i) const { return p[i]; }
    __device__ __forceinline__ void set(int i, const T& v) const { p[i] = v; }
};

struct KeyId
{
    unsigned long long k;
    int id;
};

struct Pairs
{
    using E = KeyId;
    unsigned long long* k;
    int* id;
    __device__ __forceinline__ KeyId get(int i) const { return KeyId{ k[i], id[i] }; }
    __device__ __forceinline__ void set(int i, const KeyId& v) const
    {
        k[i] = v.k;
        id[i] = v.id;
    }
};

template<typename A>
__device__ __forceinline__ void
seq_swap(const A& a, int i, int j)
{
    typename A::E t = a.get(i);
    a.set(i, a.get(j));
    a.set(j, t);
}

// ---------------------------------------------------------------------------
// std::shuffle (stl_algo.h:3738-3803) on a[base .. base+n), ONE thread.
// ---------------------------------------------------------------------------
template<typename A>
__device__ inline void
seq_shuffle(const A& a, int base, int n, uint32_t& s)
{
    if (n <= 1)
        return;

    if (2147483645u / (uint32_t)n >= (uint32_t)n) {
        int i = 1;
        if ((n % 2) == 0) {
            const uint32_t j = uniform_below(s, 1u);
            seq_swap(a, base + i, base + (int)j);
            ++i;
        }
        while (i != n) {
            const uint32_t swap_range = (uint32_t)i + 1u;
            const uint32_t b1 = swap_range + 1u;
            const uint32_t x = uniform_below(s, swap_range * b1 - 1u);
            seq_swap(a, base + i, base + (int)(x / b1));
            ++i;
            seq_swap(a, base + i, base + (int)(x % b1));
            ++i;
        }
        return;
    }

    for (int i = 1; i != n; ++i) {
        const uint32_t j = uniform_below(s, (uint32_t)i);
        seq_swap(a, base + i, base + (int)j);
    }
}

// The engine draws of std::shuffle on n elements without moving anything
// (the draws do not depend on the elements).
__device__ inline void
seq_shuffle_draws(int n, uint32_t& s)
{
    if (n <= 1)
        return;
    if (2147483645u / (uint32_t)n >= (uint32_t)n) {
        int i = 1;
        if ((n % 2) == 0) {
            uniform_below(s, 1u);
            ++i;
        }
        while (i != n) {
            const uint32_t swap_range = (uint32_t)i + 1u;
            uniform_below(s, swap_range * (swap_range + 1u) - 1u);
            i += 2;
        }
        return;
    }
    for (int i = 1; i != n; ++i)
        uniform_below(s, (uint32_t)i);
}

// ---------------------------------------------------------------------------
// std::sort (introsort + final insertion sort, stl_algo.h:1810-1957), ONE
// thread.  Less is a functor bool(const E&, const E&).
// ---------------------------------------------------------------------------
template<typename A, typename Less>
__device__ inline void
seq_unguarded_linear_insert(const A& a, int last, Less less)
{
    typename A::E val = a.get(last);
    int next = last - 1;
    for (;;) {
        typename A::E nv = a.get(next);
        if (!less(val, nv))
            break;
        a.set(last, nv);
        last = next;
        --next;
    }
    a.set(last, val);
}

template<typename A, typename Less>
__device__ inline void
seq_insertion_sort(const A& a, int first, int last, Less less)
{
    if (first == last)
        return;
    for (int i = first + 1; i != last; ++i) {
        typename A::E val = a.get(i);
        if (less(val, a.get(first))) {
            for (int j = i; j > first; --j)
                a.set(j, a.get(j - 1));
            a.set(first, val);
        } else
            seq_unguarded_linear_insert(a, i, less);
    }
}

template<typename A, typename Less>
__device__ inline void
seq_push_heap(const A& a, int first, int hole, int top, const typename A::E& value, Less less)
{
    int parent = (hole - 1) / 2;
    while (hole > top && less(a.get(first + parent), value)) {
        a.set(first + hole, a.get(first + parent));
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a.set(first + hole, value);
}

template<typename A, typename Less>
__device__ inline void
seq_adjust_heap(const A& a, int first, int hole, int len, const typename A::E& value, Less less)
{
    const int top = hole;
    int second = hole;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (less(a.get(first + second), a.get(first + (second - 1))))
            second--;
        a.set(first + hole, a.get(first + second));
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        a.set(first + hole, a.get(first + (second - 1)));
        hole = second - 1;
    }
    seq_push_heap(a, first, hole, top, value, less);
}

template<typename A, typename Less>
__device__ inline void
seq_heap_sort(const A& a, int first, int last, Less less)
{
    const int len = last - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        for (;;) {
            typename A::E value = a.get(first + parent);
            seq_adjust_heap(a, first, parent, len, value, less);
            if (parent == 0)
                break;
            parent--;
        }
    }
    while (last - first > 1) {
        --last;
        typename A::E value = a.get(last);
        a.set(last, a.get(first));
        seq_adjust_heap(a, first, 0, last - first, value, less);
    }
}

template<typename A, typename Less>
__device__ inline void
seq_move_median_to_first(const A& a, int result, int ia, int ib, int ic, Less less)
{
    const typename A::E va = a.get(ia), vb = a.get(ib), vc = a.get(ic);
    int pick;
    if (less(va, vb)) {
        if (less(vb, vc))
            pick = ib;
        else if (less(va, vc))
            pick = ic;
        else
            pick = ia;
    } else if (less(va, vc))
        pick = ia;
    else if (less(vb, vc))
        pick = ic;
    else
        pick = ib;
    seq_swap(a, result, pick);
}

template<typename A, typename Less>
__device__ inline int
seq_unguarded_partition(const A& a, int first, int last, int pivot, Less less)
{
    const typename A::E pv = a.get(pivot); // the pivot slot is outside [first, last)
    for (;;) {
        typename A::E fv = a.get(first);
        while (less(fv, pv)) {
            ++first;
            fv = a.get(first);
        }
        --last;
        typename A::E lv = a.get(last);
        while (less(pv, lv)) {
            --last;
            lv = a.get(last);
        }
        if (!(first < last))
            return first;
        a.set(first, lv);
        a.set(last, fv);
        ++first;
    }
}

// Explicit stack instead of recursion: the reference recurses on the right
// part and loops on the left; sub-ranges are disjoint, so order is free.
template<typename A, typename Less>
__device__ inline void
seq_std_sort(const A& a, int n, Less less)
{
    if (n <= 1)
        return;

    int lg = 0;
    for (int t = n; t > 1; t >>= 1)
        ++lg;

    // each pending entry: first, last, depth_limit
    int stack_first[64], stack_last[64], stack_depth[64];
    int sp = 0;
    stack_first[0] = 0;
    stack_last[0] = n;
    stack_depth[0] = lg * 2;
    sp = 1;

    while (sp) {
        --sp;
        int first = stack_first[sp];
        int last = stack_last[sp];
        int depth = stack_depth[sp];

        while (last - first > 16) {
            if (depth == 0) {
                seq_heap_sort(a, first, last, less);
                break;
            }
            --depth;
            const int mid = first + (last - first) / 2;
            seq_move_median_to_first(a, first, first + 1, mid, last - 1, less);
            const int cut = seq_unguarded_partition(a, first + 1, last, first, less);
            stack_first[sp] = cut;
            stack_last[sp] = last;
            stack_depth[sp] = depth;
            ++sp;
            last = cut;
        }
    }

    if (n > 16) {
        seq_insertion_sort(a, 0, 16, less);
        for (int i = 16; i != n; ++i)
            seq_unguarded_linear_insert(a, i, less);
    } else
        seq_insertion_sort(a, 0, n, less);
}

// ---------------------------------------------------------------------------
// The same std::sort by the whole CTA, same result element for element.
//
// What libstdc++'s introsort does to a range depends only on that range: the
// pivot (median of three moved to the front), the Hoare partition, the depth
// budget.  The two parts a partition leaves are disjoint, so they can be worked
// on at the same time; only the partition of one range is inherently a
// two-pointer walk.  One thread per warp takes ranges from a shared stack,
// partitions, pushes the right part and carries on with the left one, as the
// reference's loop does (stl_algo.h:1935-1952).  The final insertion sort
// (:1872-1885) never moves an element out of the <= 16-element range the
// partitioning left it in -- everything to the left is not greater -- and it is
// stable, so it is a stable insertion sort of every such range, done by the
// thread that produced the range.  A range that runs out of depth is
// heap-sorted (:1942-1945) by the thread that holds it.  Critical path: about
// 2 n partition steps instead of n log n.
// ---------------------------------------------------------------------------
struct SortQueue
{
    int lock;
    int top;
    int busy;
    int cap;    // tasks the stack holds
    int* tasks; // {first, last, depth} each
};

// the sequential statement on a sub-range with its depth budget (also the
// overflow path of the stack)
template<typename A, typename Less>
__device__ inline void
seq_sort_range(const A& a, int first0, int last0, int depth0, Less less)
{
    int stack_first[64], stack_last[64], stack_depth[64];
    int sp = 1;
    stack_first[0] = first0;
    stack_last[0] = last0;
    stack_depth[0] = depth0;
    while (sp) {
        --sp;
        int first = stack_first[sp], last = stack_last[sp], depth = stack_depth[sp];
        for (;;) {
            if (last - first <= 16) {
                seq_insertion_sort(a, first, last, less);
                break;
            }
            if (depth == 0) {
                seq_heap_sort(a, first, last, less);
                break;
            }
            --depth;
            const int mid = first + (last - first) / 2;
            seq_move_median_to_first(a, first, first + 1, mid, last - 1, less);
            const int cut = seq_unguarded_partition(a, first + 1, last, first, less);
            stack_first[sp] = cut;
            stack_last[sp] = last;
            stack_depth[sp] = depth;
            ++sp;
            last = cut;
        }
    }
}

// All threads of the CTA call; q lives in shared memory, q->tasks anywhere the
// CTA can read and write.  Ends with a CTA barrier.
template<typename A, typename Less>
__device__ inline void
cta_std_sort(const A& a, int n, Less less, SortQueue* q)
{
    const int tid = threadIdx.x;
    if (tid == 0) {
        int lg = 0;
        for (int t = n; t > 1; t >>= 1)
            ++lg;
        q->lock = 0;
        q->busy = 0;
        q->top = 0;
        if (n > 1) {
            q->tasks[0] = 0;
            q->tasks[1] = n;
            q->tasks[2] = 2 * lg;
            q->top = 1;
        }
    }
    __syncthreads();
    if ((tid & 31) == 0 && n > 1) {
        volatile int* vtop = &q->top;
        volatile int* vbusy = &q->busy;
        auto acquire = [&] {
            while (atomicCAS(&q->lock, 0, 1) != 0)
                ;
            __threadfence_block();
        };
        auto release = [&] {
            __threadfence_block();
            atomicExch(&q->lock, 0);
        };
        for (;;) {
            int first = -1, last = 0, depth = 0;
            bool done = false;
            acquire();
            if (*vtop > 0) {
                const int t = *vtop - 1;
                *vtop = t;
                first = q->tasks[3 * t];
                last = q->tasks[3 * t + 1];
                depth = q->tasks[3 * t + 2];
                *vbusy = *vbusy + 1;
            } else if (*vbusy == 0)
                done = true;
            release();
            if (done)
                break;
            if (first < 0) {
                __nanosleep(40);
                continue;
            }
            for (;;) {
                if (last - first <= 16) {
                    seq_insertion_sort(a, first, last, less);
                    break;
                }
                if (depth == 0) {
                    seq_heap_sort(a, first, last, less);
                    break;
                }
                --depth;
                const int mid = first + (last - first) / 2;
                seq_move_median_to_first(a, first, first + 1, mid, last - 1, less);
                const int cut = seq_unguarded_partition(a, first + 1, last, first, less);
                bool pushed = false;
                acquire();
                if (*vtop < q->cap) {
                    const int t = *vtop;
                    q->tasks[3 * t] = cut;
                    q->tasks[3 * t + 1] = last;
                    q->tasks[3 * t + 2] = depth;
                    *vtop = t + 1;
                    pushed = true;
                }
                release();
                if (!pushed)
                    seq_sort_range(a, cut, last, depth, less);
                last = cut;
            }
            acquire();
            *vbusy = *vbusy - 1;
            release();
        }
    }
    __syncthreads();
}

} // namespace bxg


#endif
