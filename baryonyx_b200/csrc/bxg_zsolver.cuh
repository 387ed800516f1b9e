// bxg_zsolver.cuh -- the sub-solvers of the Z-coefficient rows, ONE thread.
//
// Rows of the Z solver that carry a factor outside {-1, 0, 1} pick their
// variables with an exact sub-solver instead of the prefix rule
// (itm-solver-inequalities-Z.cpp:494-546):
//   * exhaustive_solver (lib/src/exhaustive-solver.hpp:40-253) for rows with
//     fewer than 10 elements: the feasible 0/1 patterns of the row are listed
//     once per problem (on the host, bxg_api.cu) and every update takes the
//     pattern with the best sum of reduced costs;
//   * branch_and_bound_solver (lib/src/branch-and-bound-solver.hpp:192-534)
//     for longer rows: a depth-first enumeration over the items in ratio
//     order, cut only by the factor sum.
// Both re-order the row's scratch so that the chosen elements come first and
// return the split index.  The enumeration order, the comparators and the
// std::sort calls decide ties, so they are restated exactly; what the
// reference computes but never reads (node bounds z, upper_bound) is left out.
//
// The subvectors of the reference's node stack (one int8 per item) are bit
// sets here, one slot per stack position: a node owns exactly one subvector
// and nodes are only ever removed from the top.
#ifndef BXG_ZSOLVER_CUH
#define BXG_ZSOLVER_CUH

#include "bxg_device.cuh"

#include <math_constants.h>

namespace bxg {

struct ZItem
{
    double r;
    double sum_z;
    int factor;
    int original_factor;
    int variable;
    int result;
};

struct ZNode
{
    double sumr;
    int sumfactor;
    int level;
};

__host__ __device__ inline int
z_words(int max_zrow)
{
    return (max_zrow + 63) / 64;
}

__host__ __device__ inline size_t
z_items_bytes(int max_zrow)
{
    return sizeof(ZItem) * (size_t)max_zrow;
}

__host__ __device__ inline size_t
z_nodes_bytes(int max_zrow)
{
    return sizeof(ZNode) * (size_t)(max_zrow + 2);
}

__host__ __device__ inline size_t
z_pool_bytes(int max_zrow)
{
    // one slot per stack position + the incumbent
    return sizeof(unsigned long long) * (size_t)(max_zrow + 3) * (size_t)z_words(max_zrow);
}

struct ZWork
{
    ZItem* items;
    ZNode* nodes;
    unsigned long long* pool;
    int words;     // 64-bit words per subvector
    int solution;  // slot of the incumbent

    __device__ ZWork(unsigned char* base, int max_zrow)
    {
        items = reinterpret_cast<ZItem*>(base);
        nodes = reinterpret_cast<ZNode*>(base + z_items_bytes(max_zrow));
        pool = reinterpret_cast<unsigned long long*>(base + z_items_bytes(max_zrow) +
                                                     z_nodes_bytes(max_zrow));
        words = z_words(max_zrow);
        solution = max_zrow + 2;
    }

    __device__ __forceinline__ unsigned long long* slot(int s) const
    {
        return pool + (size_t)s * words;
    }
    __device__ __forceinline__ void clear(int s) const
    {
        for (int w = 0; w < words; ++w)
            slot(s)[w] = 0ull;
    }
    __device__ __forceinline__ void copy(int src, int dst) const
    {
        for (int w = 0; w < words; ++w)
            slot(dst)[w] = slot(src)[w];
    }
    __device__ __forceinline__ void set(int s, int i) const
    {
        slot(s)[i >> 6] |= 1ull << (i & 63);
    }
    __device__ __forceinline__ bool test(int s, int i) const
    {
        return (slot(s)[i >> 6] >> (i & 63)) & 1ull;
    }
};

// operator< of the items, branch-and-bound-solver.hpp:212-220: both ratios
// use the LEFT operand's factor.
struct ZRatioLess
{
    bool maximize;
    __device__ bool operator()(const ZItem& l, const ZItem& r) const
    {
        const double f = (double)l.factor;
        const double a = __ddiv_rn(l.r, f), b = __ddiv_rn(r.r, f);
        return maximize ? b < a : a < b;
    }
};

// chosen items first, then by reduced cost (branch-and-bound-solver.hpp:
// 500-515, exhaustive-solver.hpp:222-237)
struct ZResultLess
{
    bool maximize;
    __device__ bool operator()(const ZItem& l, const ZItem& r) const
    {
        if (l.result == r.result)
            return maximize ? l.r > r.r : l.r < r.r;
        return l.result > r.result;
    }
};

// The row's scratch as the sub-solvers see it: value bits + position in the
// row, the factor looked up through the row's factors.
struct ZRow
{
    unsigned long long* sk;
    int* sid;
    const int* a; // factors of the row, by position
    __device__ __forceinline__ double value(int p) const
    {
        return __longlong_as_double((long long)sk[p]);
    }
    __device__ __forceinline__ int factor(int p) const { return a[sid[p]]; }
};

// Writes the items back in their final order and returns the split index
// (tail of both solve functions).
__device__ inline int
z_finish(const ZWork& w, const ZRow& R, int r_size, bool maximize)
{
    seq_std_sort(Plain<ZItem>{ w.items }, r_size, ZResultLess{ maximize });

    int middle = 0;
    while (middle != r_size && w.items[middle].result != 0)
        ++middle;

    for (int i = 0; i != r_size; ++i) {
        R.sk[i] = (unsigned long long)__double_as_longlong(w.items[i].r);
        R.sid[i] = w.items[i].variable;
    }

    if (middle == r_size)
        return w.items[0].result == 0 ? -1 : r_size;
    return middle - 1;
}

// exhaustive_solver::solve, exhaustive-solver.hpp:185-252.  masks[i] bit j =
// element j is taken by pattern i (positional, whatever order R is in).
__device__ inline int
z_ex_solve(const ZWork& w, const ZRow& R, int r_size, bool maximize, const int* masks, int solutions)
{
    for (int i = 0; i != r_size; ++i) {
        ZItem it;
        it.r = R.value(i);
        it.sum_z = 0.0;
        it.variable = R.sid[i];
        it.factor = it.original_factor = R.factor(i);
        it.result = 0;
        w.items[i] = it;
    }

    double z_best = 0.0;
    int best = 0;
    {
        const int mk = masks[0];
        for (int j = 0; j != r_size; ++j)
            if ((mk >> j) & 1)
                z_best = __dadd_rn(z_best, w.items[j].r);
    }
    for (int i = 1; i < solutions; ++i) {
        const int mk = masks[i];
        double z = 0.0;
        for (int j = 0; j != r_size; ++j)
            if ((mk >> j) & 1)
                z = __dadd_rn(z, w.items[j].r);
        if (maximize ? (z > z_best) : (z < z_best)) {
            z_best = z;
            best = i;
        }
    }

    const int mk = masks[best];
    for (int i = 0; i != r_size; ++i)
        w.items[i].result = (mk >> i) & 1;

    return z_finish(w, R, r_size, maximize);
}

// branch_and_bound_solver::solve, branch-and-bound-solver.hpp:450-533, with
// make_init_node_0/1 (:296-313), make_next_*_node (:315-394) and
// solve_equality / solve_inequality (:396-448).
__device__ inline int
z_bb_solve(const ZWork& w, const ZRow& R, int r_size, bool maximize, int b_min, int b_max)
{
    for (int i = 0; i != r_size; ++i) {
        ZItem it;
        it.r = R.value(i);
        it.sum_z = 0.0;
        it.variable = R.sid[i];
        it.original_factor = R.factor(i);
        it.factor = it.original_factor < 0 ? -it.original_factor : it.original_factor;
        it.result = 0;
        w.items[i] = it;
    }
    w.clear(w.solution);

    seq_std_sort(Plain<ZItem>{ w.items }, r_size, ZRatioLess{ maximize });

    for (int i = 0; i != r_size; ++i) {
        if (w.items[i].original_factor < 0) {
            b_min += w.items[i].factor;
            b_max += w.items[i].factor;
        }
    }

    const bool equality = b_max == b_min;
    double lower_bound = maximize ? -CUDART_INF : CUDART_INF;
    auto is_best = [&](double z) { return maximize ? z > lower_bound : z < lower_bound; };
    auto inside = [&](int s) { return equality ? s == b_min : (b_min <= s && s <= b_max); };

    int n = 0; // nodes on the stack
    bool done = false;

    // node 0: nothing taken
    w.nodes[0] = ZNode{ 0.0, 0, 0 };
    w.clear(0);
    n = 1;
    if (inside(0)) {
        w.copy(0, w.solution);
        if (equality)
            done = true; // solve_equality returns at once
    } else if (!equality && 0 > b_max)
        n = 0;

    if (!done) {
        // node 1: the first item taken
        w.nodes[n] = ZNode{ w.items[0].r, w.items[0].factor, 0 };
        w.clear(n);
        w.set(n, 0);
        ++n;
        const int sf = w.nodes[n - 1].sumfactor;
        if (inside(sf)) {
            w.copy(n - 1, w.solution);
            if (equality)
                --n;
        } else if (!equality && sf > b_max)
            --n;

        while (n) {
            const int top = n - 1;
            const int level = w.nodes[top].level + 1;
            if (level >= r_size) {
                --n;
                continue;
            }
            w.nodes[top].level = level;
            const double sumr = w.nodes[top].sumr;
            const int sumfactor = w.nodes[top].sumfactor;

            const double sumr1 = __dadd_rn(sumr, w.items[level].r);
            const int sf1 = sumfactor + w.items[level].factor;
            w.nodes[n] = ZNode{ sumr1, sf1, level };
            w.copy(top, n);
            w.set(n, level);
            ++n;

            if (inside(sf1)) {
                if (is_best(sumr1)) {
                    w.copy(n - 1, w.solution);
                    lower_bound = sumr1;
                }
                if (equality)
                    --n;
            } else if (!equality && sf1 > b_max)
                --n;
        }
    }

    for (int i = 0; i != r_size; ++i) {
        const int bit = w.test(w.solution, i) ? 1 : 0;
        w.items[i].result = w.items[i].original_factor > 0 ? bit : 1 - bit;
    }

    return z_finish(w, R, r_size, maximize);
}

} // namespace bxg

#endif
