// bxg_population.cuh -- the optimizer's population behind the C ABI
// (bxg_population_*, include/bxg.h).  Included at the end of bxg_api.cu.
//
// What the reference keeps in best_solution_recorder / storage / local_context
// (lib/src/itm-optimizer-common.hpp:52-594): `size` candidate vectors ordered by
// (remaining constraints, objective), and per worker an engine plus the
// distributions that pick parents, mutation rates and the member to replace.
// It is scalar, strictly sequential work on one random stream per worker, so
// it runs on the host -- with the standard library's engine and distributions,
// exactly like the reference, which is what makes the streams identical -- and
// only the bulk part runs on the device: scoring many vectors against the
// constraint matrix (k_score_vectors).  The lanes of a batch are the workers.
#ifndef BXG_POPULATION_CUH
#define BXG_POPULATION_CUH

#include "bxg_device.cuh"

namespace {

// remaining constraints and objective of `count` vectors; one CTA per vector.
// x: [count][words] 64-bit blocks (bit_array layout, bit-array.hpp:60-66).
// Raw merged-constraint bounds, as storage's constructor uses them
// (itm-optimizer-common.hpp:205-219).
__global__ void __launch_bounds__(256)
k_score_vectors(const __grid_constant__ DevProblem pb, const int* __restrict__ cmin,
                const int* __restrict__ cmax, const unsigned long long* __restrict__ x, int words,
                int* __restrict__ remaining, double* __restrict__ value)
{
    const unsigned long long* xv = x + (size_t)blockIdx.x * words;
    const int tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    __shared__ int s_bad[32];
    __shared__ double s_val[32];

    auto bit = [&](int i) { return (int)((xv[i >> 6] >> (i & 63)) & 1ull); };

    int bad = 0;
    for (int k = warp; k < pb.m; k += nwarps) {
        int v = 0;
        for (int e = pb.row_ptr[k] + lane; e < pb.row_ptr[k + 1]; e += 32)
            v += (pb.row_a ? pb.row_a[e] : 1) * bit(pb.row_col[e]);
#pragma unroll
        for (int o = 16; o; o >>= 1)
            v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && !(cmin[k] <= v && v <= cmax[k]))
            ++bad;
    }
    if (lane == 0)
        s_bad[warp] = bad;

    // default_cost_type::results (itm-common.hpp:1140-1147) adds in element
    // order; any order gives the same double when the sums are exact
    double part = 0.0;
    if (pb.obj_any_order) {
        for (int i = tid; i < pb.n_obj; i += T)
            if (bit(pb.obj_var[i]))
                part += pb.obj_factor[i];
#pragma unroll
        for (int o = 16; o; o >>= 1)
            part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0)
            s_val[warp] = part;
    }
    __syncthreads();
    if (tid == 0) {
        int total = 0;
        for (int w = 0; w < nwarps; ++w)
            total += s_bad[w];
        remaining[blockIdx.x] = total;
        double v = pb.cost_constant;
        if (pb.obj_any_order) {
            for (int w = 0; w < nwarps; ++w)
                v += s_val[w];
        } else {
            for (int i = 0; i < pb.n_obj; ++i)
                if (bit(pb.obj_var[i]))
                    v = __dadd_rn(v, pb.obj_factor[i]);
        }
        value[blockIdx.x] = v;
    }
}


// @region k_reinit
// best_solution_recorder::reinit's work over the variables
// (itm-optimizer-common.hpp:359-371,494-526) for every lane of a batch, from
// the plan the host drew (ReinitPlan): the crossover blocks
// x = ((x1 ^ x2) & mask) ^ x1 and the mutation loop.
//
// The mutation loop draws one Bernoulli(var_p) per variable and, after a hit,
// one Bernoulli(val_p) -- two engine steps each -- so where a variable's draws
// sit in the stream depends on the hits before it.  Seen as a sequence of
// PAIRS of engine steps it is a two-state automaton: pair q is a "value" pair
// iff pair q-1 was a "test" pair that hit.  Every thread takes kMutC
// consecutive pairs (LCG jump-ahead to its first), works out what its chunk
// does to either incoming state, and two scans over the CTA (the state maps,
// then the number of test pairs = variables) give every chunk its incoming
// state and its first variable: the whole stream is consumed in parallel and
// lands exactly where the sequential loop leaves it.
constexpr int kMutC = 16;

__device__ __forceinline__ unsigned
map_compose(unsigned g, unsigned f) // first g, then f; a map is bit0 = f(0), bit1 = f(1)
{
    return ((f >> (g & 1u)) & 1u) | (((f >> ((g >> 1) & 1u)) & 1u) << 1);
}

__global__ void __launch_bounds__(256)
k_reinit(const __grid_constant__ DevProblem pb, const __grid_constant__ DevBatch b)
{
    const int lane_id = blockIdx.x;
    const int tid = threadIdx.x, T = blockDim.x;
    const int ln = tid & 31, wp = tid >> 5, nw = T >> 5;
    const ReinitPlan pl = b.plan[lane_id];
    const int blocks = pb.xw32 >> 1;
    uint32_t* const x = b.x + (size_t)lane_id * pb.xw32;

    if (pl.mode == 1) {
        // storage::crossover(ctx, x, first, second), :359-371
        unsigned long long* x64 = reinterpret_cast<unsigned long long*>(x);
        const unsigned long long* x1 =
          reinterpret_cast<const unsigned long long*>(b.pop.x + (size_t)pl.first * pb.xw32);
        const unsigned long long* x2 =
          pl.second >= 0
            ? reinterpret_cast<const unsigned long long*>(b.pop.x + (size_t)pl.second * pb.xw32)
            : (pl.second == -1 ? reinterpret_cast<const unsigned long long*>(b.pop.bastert)
                               : b.aux + (size_t)lane_id * blocks);
        const unsigned long long* mask = b.masks + (size_t)lane_id * blocks;
        for (int i = tid; i < blocks; i += T) {
            const unsigned long long a = x1[i], c = x2[i];
            x64[i] = ((a ^ c) & mask[i]) ^ a;
        }
    }
    __syncthreads();

    uint32_t final_rng = pl.rng;
    if (pl.mutate) {
        __shared__ unsigned s_map[32];
        __shared__ int s_cnt[32];
        __shared__ long long s_end;
        if (tid == 0)
            s_end = 0;
        const int n = pb.n;
        const double var_p = pl.var_p, val_p = pl.val_p;
        long long q_base = 0;
        int i_base = 0;
        unsigned carry = 0; // the first pair of the tile is a value pair
        while (i_base < n) {
            const long long q0 = q_base + (long long)tid * kMutC;
            uint32_t st = lcg_mul(pl.rng, lcg_pow(2ull * (unsigned long long)q0));
            unsigned A = 0, V = 0; // pair j of the chunk: u < var_p, u < val_p (one pair past the chunk too)
#pragma unroll
            for (int j = 0; j <= kMutC; ++j) {
                const uint32_t g1 = lcg_next(st);
                const uint32_t g2 = lcg_next(st);
                const double u = canonical_from(g1, g2);
                A |= (unsigned)(u < var_p) << j;
                V |= (unsigned)(u < val_p) << j;
            }
            // what the chunk does to either incoming state, and its test pairs
            unsigned o[2];
            int tests[2];
#pragma unroll
            for (int in = 0; in != 2; ++in) {
                unsigned sv = in;
                int t = 0;
#pragma unroll
                for (int j = 0; j != kMutC; ++j) {
                    if (sv)
                        sv = 0;
                    else {
                        ++t;
                        sv = (A >> j) & 1u;
                    }
                }
                o[in] = sv;
                tests[in] = t;
            }
            unsigned f = o[0] | (o[1] << 1);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned g = __shfl_up_sync(0xffffffffu, f, d);
                if (ln >= d)
                    f = map_compose(g, f);
            }
            if (ln == 31)
                s_map[wp] = f;
            __syncthreads();
            unsigned before = 2u; // identity
            for (int w = 0; w < wp; ++w)
                before = map_compose(before, s_map[w]);
            unsigned total = before;
            for (int w = wp; w < nw; ++w)
                total = map_compose(total, s_map[w]);
            unsigned excl = __shfl_up_sync(0xffffffffu, f, 1);
            if (ln == 0)
                excl = 2u;
            excl = map_compose(before, excl);
            const unsigned in_state = (excl >> carry) & 1u;
            const int my_tests = tests[in_state];

            int incl = my_tests;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, d);
                if (ln >= d)
                    incl += v;
            }
            if (ln == 31)
                s_cnt[wp] = incl;
            __syncthreads();
            int cnt_before = 0, cnt_total = 0;
            for (int w = 0; w < nw; ++w) {
                const int v = s_cnt[w];
                if (w < wp)
                    cnt_before += v;
                cnt_total += v;
            }
            int i = i_base + cnt_before + incl - my_tests;

            unsigned sv = in_state;
#pragma unroll
            for (int j = 0; j != kMutC; ++j) {
                if (sv)
                    sv = 0;
                else {
                    const unsigned hit = (A >> j) & 1u;
                    if (i < n) {
                        if (hit) {
                            // x.set(i, value): bit_array::set(int, bool), bit-array.hpp:143-148,
                            // leaves the value in bit 0 of a block only and clears any other bit
                            const uint32_t bit = 1u << (i & 31);
                            if ((i & 63) == 0 && ((V >> (j + 1)) & 1u))
                                atomicOr(&x[i >> 5], bit);
                            else
                                atomicAnd(&x[i >> 5], ~bit);
                        }
                        if (i == n - 1)
                            s_end = q0 + j + 1 + hit;
                    }
                    ++i;
                    sv = hit;
                }
            }
            carry = (total >> carry) & 1u;
            i_base += cnt_total;
            q_base += (long long)T * kMutC;
            __syncthreads();
        }
        final_rng = lcg_mul(pl.rng, lcg_pow(2ull * (unsigned long long)s_end));
    }
    if (tid == 0)
        b.ctl[lane_id].rng = final_rng;
}

// storage::crossover's other branch (:389,402): member `first` takes the bits
// of the bastert / random vector and keeps its value and hash.  After k_reinit,
// so that no lane reads a member while it is rewritten.
__global__ void __launch_bounds__(256)
k_pop_replace(const __grid_constant__ DevProblem pb, const __grid_constant__ DevBatch b)
{
    const ReinitPlan pl = b.plan[blockIdx.x];
    if (pl.mode != 2 && pl.mode != 3)
        return;
    uint32_t* dst = b.pop.x + (size_t)pl.first * pb.xw32;
    const uint32_t* src = pl.mode == 2
                            ? b.pop.bastert
                            : reinterpret_cast<const uint32_t*>(b.aux + (size_t)blockIdx.x * (pb.xw32 >> 1));
    for (int w = threadIdx.x; w < pb.xw32; w += blockDim.x)
        dst[w] = src[w];
}

using engine_t = std::default_random_engine;

uint32_t
engine_state(const engine_t& e)
{
    // minstd_rand0 streams its state as a decimal number
    std::ostringstream os;
    os << e;
    return static_cast<uint32_t>(std::stoul(os.str()));
}

void
engine_set_state(engine_t& e, uint32_t state)
{
    std::istringstream is(std::to_string(state));
    is >> e;
}

} // namespace

struct bxg_population
{
    struct member
    {
        std::vector<uint64_t> x;
        double value = 0;
        uint64_t hash = 0;
        int64_t loop = 0;
        int remaining = 0;
    };

    // local_context, itm-optimizer-common.hpp:52-92
    struct worker
    {
        engine_t rng;
        std::normal_distribution<double> parent;
        std::normal_distribution<double> mutation_rate;
        std::normal_distribution<double> mutation_value;
        std::uniform_int_distribution<int> victim;
        std::uniform_int_distribution<std::size_t> mask;
        std::bernoulli_distribution special_parent;
        double improve = 0; // m_kappa_append[thread_id]
        bool seeded = false;
    };

    bxg_problem* pb = nullptr;
    bxg_params p{};
    int size = 0, words = 0;
    engine_t master;
    std::vector<int> order; // m_indices
    std::vector<member> data;
    std::vector<uint64_t> bastert, scratch;
    std::vector<worker> workers;
    DevBuf<int> d_cmin, d_cmax, d_remaining;
    DevBuf<unsigned long long> d_x;
    DevBuf<double> d_value;

    // The population as the lanes see it during bxg_optimize_round (DevPop).
    // The host vectors above and the device copy are brought in line on demand:
    // host_newer after the constructor or a host-side operation of this API,
    // device_newer after a round.
    DevBuf<uint32_t> d_px, d_bastert;
    DevBuf<double> d_pvalue, d_pduration;
    DevBuf<int> d_premaining, d_porder, d_plock;
    DevBuf<unsigned long long> d_phash;
    DevBuf<long long> d_ploop;
    DevBuf<unsigned long long> d_pfirst;
    DevBuf<ReinitPlan> d_plan;
    DevBuf<unsigned long long> d_masks, d_aux;
    std::vector<ReinitPlan> plans;
    bool host_newer = true, device_newer = false, reinit_done = false;
    unsigned long long t0_ns = 0;
    int64_t rounds = 0;

    DevPop dev_view() const
    {
        DevPop v{};
        v.size = size;
        v.victim_lo = p.init_population_size / 5;
        v.victim_hi = p.init_population_size - 1;
        v.x = d_px.p;
        v.value = d_pvalue.p;
        v.remaining = d_premaining.p;
        v.hash = d_phash.p;
        v.loop = d_ploop.p;
        v.duration = d_pduration.p;
        v.order = d_porder.p;
        v.lock = d_plock.p;
        v.bastert = d_bastert.p;
        v.first_ns = d_pfirst.p;
        v.t0_ns = t0_ns;
        return v;
    }

    int to_device(cudaStream_t st)
    {
        const size_t w32 = static_cast<size_t>(words) * 2;
        if (!d_px.p) {
            CUDA_TRY(d_px.alloc(static_cast<size_t>(size) * w32));
            CUDA_TRY(d_bastert.alloc(w32));
            CUDA_TRY(d_pvalue.alloc(size));
            CUDA_TRY(d_pduration.alloc(size));
            CUDA_TRY(d_premaining.alloc(size));
            CUDA_TRY(d_porder.alloc(size));
            CUDA_TRY(d_plock.alloc(1));
            CUDA_TRY(d_phash.alloc(size));
            CUDA_TRY(d_ploop.alloc(size));
            CUDA_TRY(d_pfirst.alloc(1));
            CUDA_TRY(cudaMemset(d_pfirst.p, 0xff, sizeof(unsigned long long)));
            CUDA_TRY(cudaMemset(d_plock.p, 0, sizeof(int)));
            CUDA_TRY(cudaMemset(d_pduration.p, 0, sizeof(double) * size));
            CUDA_TRY(cudaMemcpy(d_bastert.p, bastert.data(), sizeof(uint64_t) * words,
                                cudaMemcpyHostToDevice));
            CUDA_TRY(cudaStreamSynchronize(nullptr));
        }
        std::vector<uint64_t> flat(static_cast<size_t>(size) * words);
        std::vector<double> value(size);
        std::vector<int> remaining(size);
        std::vector<unsigned long long> hash(size);
        std::vector<long long> loop(size);
        for (int i = 0; i != size; ++i) {
            std::copy(data[i].x.begin(), data[i].x.end(), flat.begin() + static_cast<size_t>(i) * words);
            value[i] = data[i].value;
            remaining[i] = data[i].remaining;
            hash[i] = data[i].hash;
            loop[i] = data[i].loop;
        }
        CUDA_TRY(cudaMemcpyAsync(d_px.p, flat.data(), sizeof(uint64_t) * flat.size(),
                                 cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_pvalue.p, value.data(), sizeof(double) * size, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_premaining.p, remaining.data(), sizeof(int) * size,
                                 cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_phash.p, hash.data(), sizeof(unsigned long long) * size,
                                 cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_ploop.p, loop.data(), sizeof(long long) * size, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_porder.p, order.data(), sizeof(int) * size, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        host_newer = false;
        return BXG_OK;
    }

    // scores and order only (what reinit's draws and the callers' reports need)
    int fetch_scores(cudaStream_t st)
    {
        std::vector<double> value(size);
        std::vector<int> remaining(size);
        std::vector<unsigned long long> hash(size);
        std::vector<long long> loop(size);
        CUDA_TRY(cudaMemcpyAsync(value.data(), d_pvalue.p, sizeof(double) * size, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(remaining.data(), d_premaining.p, sizeof(int) * size,
                                 cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(hash.data(), d_phash.p, sizeof(unsigned long long) * size,
                                 cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(loop.data(), d_ploop.p, sizeof(long long) * size, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(order.data(), d_porder.p, sizeof(int) * size, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        for (int i = 0; i != size; ++i) {
            data[i].value = value[i];
            data[i].remaining = remaining[i];
            data[i].hash = hash[i];
            data[i].loop = loop[i];
        }
        return BXG_OK;
    }

    // everything, the vectors included (host-side operations, parity taps)
    int to_host()
    {
        if (!device_newer)
            return BXG_OK;
        CUDA_TRY(cudaSetDevice(pb->device));
        CUDA_TRY(cudaDeviceSynchronize());
        int rc = fetch_scores(nullptr);
        if (rc != BXG_OK)
            return rc;
        std::vector<uint64_t> flat(static_cast<size_t>(size) * words);
        CUDA_TRY(cudaMemcpy(flat.data(), d_px.p, sizeof(uint64_t) * flat.size(), cudaMemcpyDeviceToHost));
        for (int i = 0; i != size; ++i)
            data[i].x.assign(flat.begin() + static_cast<size_t>(i) * words,
                             flat.begin() + static_cast<size_t>(i + 1) * words);
        device_newer = false;
        return BXG_OK;
    }

    // The draws of best_solution_recorder::reinit (:528-554) that do not run
    // over the variables: which branch, the parents, the masks, the mutation
    // rates.  Leaves in `pl.rng` the engine state the device continues from.
    void plan_reinit(worker& w, double kappa_min, double kappa_max, ReinitPlan& pl,
                     unsigned long long* masks, unsigned long long* aux)
    {
        pl = ReinitPlan{};
        pl.kappa = kappa_min;
        pl.first = pl.second = 0;
        if (w.improve < p.init_kappa_improve_stop) {
            w.improve += p.init_kappa_improve_increase;
            pl.kappa = kappa_min + (kappa_max - kappa_min) * w.improve;
        } else {
            w.improve = p.init_kappa_improve_start;
            // storage::crossover(ctx, x), :373-421
            auto draw_masks = [&] {
                for (int i = 0; i != words; ++i)
                    masks[i] = w.mask(w.rng);
            };
            if (w.special_parent(w.rng)) {
                pl.first = order[pick_parent(w)];
                std::bernoulli_distribution coin(0.5);
                const bool use_bastert = coin(w.rng);
                if (!use_bastert) {
                    std::bernoulli_distribution half(0.5);
                    std::fill(aux, aux + words, 0ull);
                    for (int i = 0; i != pb->n; ++i)
                        if (half(w.rng))
                            aux[i >> 6] |= 1ull << (i & 63);
                }
                if (coin(w.rng)) {
                    pl.mode = 1;
                    pl.second = use_bastert ? -1 : -2;
                    draw_masks();
                } else
                    pl.mode = use_bastert ? 2 : 3;
            } else {
                pl.first = order[pick_parent(w)];
                pl.second = order[pick_parent(w)];
                while (pl.first == pl.second)
                    pl.second = order[pick_parent(w)];
                pl.mode = 1;
                draw_masks();
            }
        }
        // best_solution_recorder::mutation, :494-526
        if (!(w.mutation_value.mean() == 0.0 && w.mutation_value.stddev() == 0.0)) {
            double rate, ones;
            do
                rate = w.mutation_rate(w.rng);
            while (rate <= 0.0 || rate >= 1.0);
            do
                ones = w.mutation_value(w.rng);
            while (ones < 0.0 || ones > 1.0);
            pl.var_p = rate;
            pl.val_p = ones;
            pl.mutate = 1;
        }
        pl.rng = engine_state(w.rng);
    }

    bool maximize() const { return pb->maximize != 0; }

    bool better(double l, double r) const { return maximize() ? l > r : l < r; }

    static uint64_t hash_of(const uint64_t* x, int words)
    {
        // bit_array_hash, bit-array.hpp:410-423; std::hash<size_t> is the identity
        uint64_t h = 0;
        for (int i = 0; i != words; ++i)
            h ^= x[i] + 0x9e3779b9u + (h << 6) + (h >> 2);
        return h;
    }

    double objective(const uint64_t* x) const
    {
        double v = pb->cost_constant;
        for (size_t i = 0; i != pb->obj_var.size(); ++i)
            if ((x[pb->obj_var[i] >> 6] >> (pb->obj_var[i] & 63)) & 1u)
                v += pb->obj_factor[i];
        return v;
    }

    // storage::sort, itm-optimizer-common.hpp:423-441
    void sort()
    {
        std::sort(order.begin(), order.end(), [this](int a, int b) {
            const member &l = data[a], &r = data[b];
            if (l.remaining != r.remaining)
                return l.remaining < r.remaining;
            return better(l.value, r.value);
        });
    }

    // remaining constraints and objective of `count` vectors on the device
    int score(const uint64_t* x, int count, int* remaining, double* value)
    {
        if (count <= 0)
            return BXG_OK;
        CUDA_TRY(cudaSetDevice(pb->device));
        const size_t n = static_cast<size_t>(count) * words;
        if (d_x.n < n)
            CUDA_TRY(d_x.alloc(n));
        if (d_remaining.n < static_cast<size_t>(count)) {
            CUDA_TRY(d_remaining.alloc(count));
            CUDA_TRY(d_value.alloc(count));
        }
        CUDA_TRY(cudaMemcpy(d_x.p, x, sizeof(uint64_t) * n, cudaMemcpyHostToDevice));
        k_score_vectors<<<count, 256>>>(pb->dev, d_cmin.p, d_cmax.p, d_x.p, words, d_remaining.p,
                                        d_value.p);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpy(remaining, d_remaining.p, sizeof(int) * count, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(value, d_value.p, sizeof(double) * count, cudaMemcpyDeviceToHost));
        return BXG_OK;
    }

    // storage::choose_a_solution, :151-160
    int pick_parent(worker& w)
    {
        double v;
        do
            v = w.parent(w.rng);
        while (v < 0 || v > 1);
        return static_cast<int>(size * v);
    }

    // storage::crossover(ctx, x, first, second), :359-371
    void mix(worker& w, uint64_t* x, const uint64_t* a, const uint64_t* b)
    {
        for (int i = 0; i != words; ++i)
            x[i] = ((a[i] ^ b[i]) & w.mask(w.rng)) ^ a[i];
    }

    // storage::crossover(ctx, x), :373-421
    void crossover(worker& w, uint64_t* x)
    {
        if (w.special_parent(w.rng)) {
            member& first = data[order[pick_parent(w)]];
            std::bernoulli_distribution coin(0.5);
            const bool use_bastert = coin(w.rng);
            if (!use_bastert) {
                // init_with_random(m_random, rng, n, 0.5), itm-common.hpp:269-282
                std::bernoulli_distribution half(0.5);
                for (int i = 0; i != pb->n; ++i) {
                    const uint64_t bit = uint64_t{ 1 } << (i & 63);
                    if (half(w.rng))
                        scratch[i >> 6] |= bit;
                    else
                        scratch[i >> 6] &= ~bit;
                }
            }
            const std::vector<uint64_t>& other = use_bastert ? bastert : scratch;
            if (coin(w.rng))
                mix(w, x, first.x.data(), other.data());
            else
                first.x = other; // the member's value and hash stay as they were (:389,402)
        } else {
            const int first = order[pick_parent(w)];
            int second = order[pick_parent(w)];
            while (first == second)
                second = order[pick_parent(w)];
            mix(w, x, data[first].x.data(), data[second].x.data());
        }
    }

    // best_solution_recorder::mutation, :494-526
    void mutation(worker& w, uint64_t* x)
    {
        if (w.mutation_value.mean() == 0.0 && w.mutation_value.stddev() == 0.0)
            return;
        double rate, ones;
        do
            rate = w.mutation_rate(w.rng);
        while (rate <= 0.0 || rate >= 1.0);
        do
            ones = w.mutation_value(w.rng);
        while (ones < 0.0 || ones > 1.0);

        std::bernoulli_distribution hit(rate), one(ones);
        for (int i = 0; i != pb->n; ++i) {
            if (!hit(w.rng))
                continue;
            // bit_array::set(i, bool), bit-array.hpp:143-148: the value lands
            // in bit 0 of the block only; any other bit is cleared
            const uint64_t val = one(w.rng) ? 1u : 0u;
            const uint64_t bit = uint64_t{ 1 } << (i & 63);
            uint64_t& ref = x[i >> 6];
            ref ^= (val ^ ref) & bit;
        }
    }

    // best_solution_recorder::reinit, :528-554
    double reinit(worker& w, double kappa_min, double kappa_max, uint64_t* x)
    {
        double kappa = kappa_min;
        if (w.improve < p.init_kappa_improve_stop) {
            w.improve += p.init_kappa_improve_increase;
            kappa = kappa_min + (kappa_max - kappa_min) * w.improve;
        } else {
            w.improve = p.init_kappa_improve_start;
            crossover(w, x);
        }
        mutation(w, x);
        return kappa;
    }

    // storage::insert, :228-300: the victim comes from the worst four fifths
    void replace(worker& w, const uint64_t* x, double value, uint64_t hash, int64_t loop,
                 int remaining)
    {
        member& mb = data[order[w.victim(w.rng)]];
        mb.x.assign(x, x + words);
        mb.value = value;
        mb.hash = hash;
        mb.loop = loop;
        mb.remaining = remaining;
        sort();
    }

    // try_update, :571-584 / can_be_inserted(hash, value), :315-326
    bool offer_solution(worker& w, const uint64_t* x, double value, int64_t loop)
    {
        const uint64_t h = hash_of(x, words);
        for (const member& mb : data)
            if (mb.remaining == 0 && mb.value == value && mb.hash == h)
                return false;
        replace(w, x, value, h, loop, 0);
        return true;
    }

    // try_advance, :556-569 / can_be_inserted(hash, constraints), :302-313
    bool offer_advance(worker& w, const uint64_t* x, int remaining, int64_t loop)
    {
        const uint64_t h = hash_of(x, words);
        for (const member& mb : data)
            if (mb.remaining == remaining && mb.hash == h)
                return false;
        replace(w, x, objective(x), h, loop, remaining);
        return true;
    }
};

namespace {

// init_with_pre_solve with both vectors present (itm-common.hpp:284-374), as
// storage's constructor calls it.
void
host_pre_solve_pair(const bxg_problem& pb, engine_t& rng, double init_random,
                    std::vector<uint64_t>& xp, std::vector<uint64_t>& xo)
{
    struct rc
    {
        float value;
        int factor;
        int id;
    };
    std::vector<rc> R(pb.max_row);
    std::bernoulli_distribution take(init_random);
    const bool maximize = pb.maximize != 0;
    auto put = [](std::vector<uint64_t>& x, int i, bool v) {
        const uint64_t bit = uint64_t{ 1 } << (i & 63);
        if (v)
            x[i >> 6] |= bit;
        else
            x[i >> 6] &= ~bit;
    };

    for (int k = 0; k != pb.m; ++k) {
        if (!take(rng))
            continue;
        const int r_size = pb.row_ptr[k + 1] - pb.row_ptr[k];
        R.resize(r_size);
        for (int i = 0; i != r_size; ++i) {
            const int e = pb.row_ptr[k] + i;
            R[i] = { static_cast<float>(pb.cost[pb.row_col[e]]), pb.row_a[e], pb.row_col[e] };
        }
        std::shuffle(R.begin(), R.end(), rng);
        if (maximize)
            std::sort(R.begin(), R.end(), [](const rc& l, const rc& r) { return l.value > r.value; });
        else
            std::sort(R.begin(), R.end(), [](const rc& l, const rc& r) { return l.value < r.value; });

        const int lo = pb.cmin[k], hi = pb.cmax[k];
        int sum = 0, best = -2;
        for (int i = -1; i < r_size; ++i) { // pessimistic: the shortest prefix that fits
            if (lo <= sum && sum <= hi) {
                best = i;
                break;
            }
            if (i + 1 < r_size)
                sum += R[i + 1].factor;
        }
        for (int i = 0; i != r_size; ++i)
            put(xp, R[i].id, i <= best);

        sum = 0;
        best = -2;
        for (int i = -1; i < r_size; ++i) { // optimistic: the longest one before the sign change
            if (lo <= sum && sum <= hi)
                best = i;
            if (best != -2 && i + 1 < r_size && (maximize ? R[i + 1].value < 0 : R[i + 1].value > 0))
                break;
            if (i + 1 < r_size)
                sum += R[i + 1].factor;
        }
        for (int i = 0; i != r_size; ++i)
            put(xo, R[i].id, i <= best);
    }
}

} // namespace

extern "C" {

static int population_create_body(bxg_problem* pb, const bxg_params* p, uint64_t master_seed, int workers,
                      bxg_population** out);

int
bxg_population_create(bxg_problem* pb, const bxg_params* p, uint64_t master_seed, int workers,
                      bxg_population** out)
{
    return guarded("bxg_population_create", [&] { return population_create_body(pb, p, master_seed, workers, out); });
}

static int
population_create_body(bxg_problem* pb, const bxg_params* p, uint64_t master_seed, int workers,
                      bxg_population** out)
{
    if (!pb || !p || !out || workers < 1 || p->init_population_size < 1)
        return fail(BXG_E_INVALID, "bxg_population_create: bad argument");
    CUDA_TRY(cudaSetDevice(pb->device));
    std::unique_ptr<bxg_population> q(new (std::nothrow) bxg_population);
    if (!q)
        return fail(BXG_E_MEMORY, "out of host memory");
    q->pb = pb;
    q->p = *p;
    q->size = p->init_population_size;
    q->words = pb->n / 64 + 1;
    q->master.seed(static_cast<engine_t::result_type>(master_seed));
    q->workers.resize(workers);
    q->order.resize(q->size);
    q->data.resize(q->size);
    q->bastert.assign(q->words, 0);
    q->scratch.assign(q->words, 0);
    for (auto& mb : q->data)
        mb.x.assign(q->words, 0);
    CUDA_TRY(q->d_cmin.upload(pb->cmin));
    CUDA_TRY(q->d_cmax.upload(pb->cmax));

    // storage's constructor, itm-optimizer-common.hpp:162-226
    const int n = pb->n;
    for (int i = 0; i != n; ++i) {
        const double c = pb->cost[i];
        if (pb->maximize ? c > 0 : c < 0) // init_with_bastert, value_if_0 = 0
            q->bastert[i >> 6] |= uint64_t{ 1 } << (i & 63);
    }
    engine_t& rng = q->master;
    const int half = q->size / 2;
    for (int i = 0; i != half; ++i) {
        q->data[i].x = q->bastert;
        std::bernoulli_distribution flip(std::clamp(static_cast<double>(i) / (5 * half), 0.0, 1.0));
        for (int v = 0; v != n; ++v)
            if (flip(rng))
                q->data[i].x[v >> 6] ^= uint64_t{ 1 } << (v & 63);
    }
    for (int i = half; i + 1 < q->size; i += 2) {
        for (int side = 0; side != 2; ++side) {
            std::bernoulli_distribution one(side == 0 ? 0.2 : 0.8);
            std::vector<uint64_t>& x = q->data[i + side].x;
            for (int v = 0; v != n; ++v) {
                const uint64_t bit = uint64_t{ 1 } << (v & 63);
                if (one(rng))
                    x[v >> 6] |= bit;
                else
                    x[v >> 6] &= ~bit;
            }
        }
        host_pre_solve_pair(*pb, rng, std::clamp(static_cast<double>(i) / (5 * q->size), 0.0, 1.0),
                            q->data[i].x, q->data[i + 1].x);
    }

    // score every member on the device
    std::vector<uint64_t> flat(static_cast<size_t>(q->size) * q->words);
    for (int i = 0; i != q->size; ++i)
        std::copy(q->data[i].x.begin(), q->data[i].x.end(), flat.begin() + static_cast<size_t>(i) * q->words);
    std::vector<int> rem(q->size);
    std::vector<double> val(q->size);
    const int rc = q->score(flat.data(), q->size, rem.data(), val.data());
    if (rc != BXG_OK)
        return rc;
    for (int i = 0; i != q->size; ++i) {
        q->data[i].hash = bxg_population::hash_of(q->data[i].x.data(), q->words);
        q->data[i].value = val[i];
        q->data[i].remaining = rem[i];
        q->order[i] = i;
    }
    q->sort();
    CUDA_TRY(cudaStreamSynchronize(nullptr));
    *out = q.release();
    return BXG_OK;
}

void
bxg_population_destroy(bxg_population* pop)
{
    delete pop;
}

int
bxg_population_size(const bxg_population* pop)
{
    return pop ? pop->size : 0;
}

int
bxg_population_master_state(bxg_population* pop, uint32_t* state)
{
    if (!pop || !state)
        return fail(BXG_E_INVALID, "bxg_population_master_state: bad argument");
    *state = engine_state(pop->master);
    return BXG_OK;
}

int
bxg_population_generate_seeds(bxg_population* pop, int count, uint64_t* seeds)
{
    if (!pop || count < 1 || !seeds)
        return fail(BXG_E_INVALID, "bxg_population_generate_seeds: bad argument");
    // generate_seed, itm-common.hpp:1457-1484, on the master engine
    using type = engine_t::result_type;
    std::uniform_int_distribution<type> dst(std::numeric_limits<type>::min(),
                                            std::numeric_limits<type>::max());
    engine_t& rng = pop->master;
    seeds[0] = dst(rng);
    for (int i = 1; i != count; ++i) {
        seeds[i] = dst(rng);
        unsigned j = static_cast<unsigned>(i) - 1;
        while (j > 0) {
            if (seeds[j] == seeds[i]) {
                seeds[j] = dst(rng);
                j = static_cast<unsigned>(i);
            }
            --j;
        }
    }
    return BXG_OK;
}

int
bxg_population_member(bxg_population* pop, int i, uint64_t* x_words, double* value, int* remaining,
                      uint64_t* hash, int64_t* loop)
{
    if (!pop || i < 0 || i >= pop->size)
        return fail(BXG_E_INVALID, "bxg_population_member: bad argument");
    if (int rc_ = pop->to_host())
        return rc_;
    const bxg_population::member& mb = pop->data[pop->order[i]];
    if (x_words)
        std::copy(mb.x.begin(), mb.x.end(), x_words);
    if (value)
        *value = mb.value;
    if (remaining)
        *remaining = mb.remaining;
    if (hash)
        *hash = mb.hash;
    if (loop)
        *loop = mb.loop;
    return BXG_OK;
}

int
bxg_population_worker_seed(bxg_population* pop, int worker, uint64_t seed)
{
    if (!pop || worker < 0 || worker >= static_cast<int>(pop->workers.size()))
        return fail(BXG_E_INVALID, "bxg_population_worker_seed: bad argument");
    const bxg_params& p = pop->p;
    bxg_population::worker& w = pop->workers[worker];
    w = bxg_population::worker{};
    w.rng.seed(static_cast<engine_t::result_type>(seed));
    w.parent = std::normal_distribution<double>(p.init_crossover_solution_selection_mean,
                                                p.init_crossover_solution_selection_stddev);
    w.mutation_rate = std::normal_distribution<double>(p.init_mutation_variable_mean,
                                                       p.init_mutation_variable_stddev);
    w.mutation_value =
      std::normal_distribution<double>(p.init_mutation_value_mean, p.init_mutation_value_stddev);
    w.victim = std::uniform_int_distribution<int>(p.init_population_size / 5,
                                                  p.init_population_size - 1);
    w.mask = std::uniform_int_distribution<std::size_t>(0);
    w.special_parent = std::bernoulli_distribution(p.init_crossover_bastert_insertion);
    w.seeded = true;
    return BXG_OK;
}

#define BXG_WORKER(pop, worker, name)                                                            \
    if (!(pop) || (worker) < 0 || (worker) >= static_cast<int>((pop)->workers.size()) ||         \
        !(pop)->workers[worker].seeded)                                                          \
        return fail(BXG_E_INVALID, name ": bad population or worker (seed it first)");           \
    bxg_population::worker& w = (pop)->workers[worker]

int
bxg_population_worker_state(bxg_population* pop, int worker, uint32_t* get_state,
                            const uint32_t* set_state)
{
    BXG_WORKER(pop, worker, "bxg_population_worker_state");
    if (set_state)
        engine_set_state(w.rng, *set_state);
    if (get_state)
        *get_state = engine_state(w.rng);
    return BXG_OK;
}

int
bxg_population_reinit(bxg_population* pop, int worker, double kappa_min, double kappa_max,
                      uint64_t* x_words, double* kappa)
{
    BXG_WORKER(pop, worker, "bxg_population_reinit");
    if (int rc_ = pop->to_host())
        return rc_;
    pop->host_newer = true;
    if (!x_words || !kappa)
        return fail(BXG_E_INVALID, "bxg_population_reinit: bad argument");
    *kappa = pop->reinit(w, kappa_min, kappa_max, x_words);
    return BXG_OK;
}

int
bxg_population_crossover(bxg_population* pop, int worker, uint64_t* x_words)
{
    BXG_WORKER(pop, worker, "bxg_population_crossover");
    if (int rc_ = pop->to_host())
        return rc_;
    pop->host_newer = true;
    pop->crossover(w, x_words);
    return BXG_OK;
}

int
bxg_population_mutation(bxg_population* pop, int worker, uint64_t* x_words)
{
    BXG_WORKER(pop, worker, "bxg_population_mutation");
    if (int rc_ = pop->to_host())
        return rc_;
    pop->host_newer = true;
    pop->mutation(w, x_words);
    return BXG_OK;
}

int
bxg_population_try_update(bxg_population* pop, int worker, const uint64_t* x_words, double value,
                          int64_t loop)
{
    BXG_WORKER(pop, worker, "bxg_population_try_update");
    if (int rc_ = pop->to_host())
        return rc_;
    pop->host_newer = true;
    pop->offer_solution(w, x_words, value, loop);
    return BXG_OK;
}

int
bxg_population_try_advance(bxg_population* pop, int worker, const uint64_t* x_words, int remaining,
                           int64_t loop)
{
    BXG_WORKER(pop, worker, "bxg_population_try_advance");
    if (int rc_ = pop->to_host())
        return rc_;
    pop->host_newer = true;
    pop->offer_advance(w, x_words, remaining, loop);
    return BXG_OK;
}

int
bxg_population_score(bxg_population* pop, const uint64_t* x_words, int count, int32_t* remaining,
                     double* value)
{
    if (!pop || !x_words || !remaining || !value || count < 0)
        return fail(BXG_E_INVALID, "bxg_population_score: bad argument");
    return pop->score(x_words, count, remaining, value);
}

int
bxg_optimize_reinit(bxg_batch* b, bxg_population* pop, int first_round, double* kappa)
{
    if (!b || !pop || pop->pb != b->pb)
        return fail(BXG_E_INVALID, "bxg_optimize_round: batch and population of different problems");
    const int L = b->n_lanes;
    if (static_cast<int>(pop->workers.size()) < L)
        return fail(BXG_E_INVALID, "bxg_optimize_round: fewer workers than lanes");
    for (int l = 0; l != L; ++l)
        if (!pop->workers[l].seeded)
            return fail(BXG_E_INVALID, "bxg_optimize_round: seed every worker first");
    if (pop->size > b->cfg.stage_cap || pop->size < 5)
        return fail(BXG_E_UNSUPPORTED, "bxg_optimize_round: population size outside [5, staging buffer]");
    bxg_problem* pb = b->pb;
    CUDA_TRY(cudaSetDevice(pb->device));
    const int xw32 = pb->dev.xw32;
    const int words = pop->words;
    const size_t xn = static_cast<size_t>(L) * xw32;

    if (pop->host_newer) {
        int rc = pop->to_device(b->stream);
        if (rc != BXG_OK)
            return rc;
    }
    if (!pop->t0_ns) {
        // the device clock the durations count from (m_start, :467-480)
        int rc_ = device_clock(b->stream, &pop->t0_ns);
        if (rc_ != BXG_OK)
            return rc_;
    }

    if (first_round) {
        // bit_array x(variables) and a fresh Solver (:611,649): x = 0, P = pi = 0
        std::vector<uint32_t> rng0(L);
        for (int l = 0; l != L; ++l) {
            rng0[l] = engine_state(pop->workers[l].rng);
            pop->workers[l].improve = 0;
        }
        if (b->d_x0.n < xn) {
            CUDA_TRY(b->d_x0.alloc(xn));
            CUDA_TRY(cudaStreamSynchronize(nullptr)); // the batch's stream is non-blocking
        }
        CUDA_TRY(cudaMemsetAsync(b->d_x0.p, 0, sizeof(uint32_t) * xn, b->stream));
        CUDA_TRY(cudaMemcpyAsync(b->d_rng0.p, rng0.data(), sizeof(uint32_t) * L, cudaMemcpyHostToDevice,
                                 b->stream));
        int rc = run_begin(b, nullptr, b->d_x0.p, b->d_rng0.p, 1);
        if (rc != BXG_OK)
            return rc;
    } else if (!b->begun)
        return fail(BXG_E_INVALID, "bxg_optimize_round: the first call must have first_round = 1");

    // the engines where the lanes left them, the order of the population
    int rc = fetch_ctl(b);
    if (rc != BXG_OK)
        return rc;
    if (pop->device_newer) {
        rc = pop->fetch_scores(b->stream);
        if (rc != BXG_OK)
            return rc;
    }

    // reinit, host part: one plan per lane
    pop->plans.resize(L);
    if (pop->d_plan.n < static_cast<size_t>(L)) {
        CUDA_TRY(pop->d_plan.alloc(L));
        CUDA_TRY(cudaStreamSynchronize(nullptr));
    }
    std::vector<unsigned long long> masks(words), aux(words);
    bool any_replace = false;
    for (int l = 0; l != L; ++l) {
        bxg_population::worker& w = pop->workers[l];
        engine_set_state(w.rng, b->h_ctl[l].rng);
        ReinitPlan& pl = pop->plans[l];
        pop->plan_reinit(w, b->p.kappa_min, b->p.kappa_max, pl, masks.data(), aux.data());
        if (pl.mode == 1 || pl.mode == 3) {
            const size_t need = static_cast<size_t>(L) * words;
            if (pop->d_masks.n < need) {
                CUDA_TRY(pop->d_masks.alloc(need));
                CUDA_TRY(pop->d_aux.alloc(need));
                CUDA_TRY(cudaStreamSynchronize(nullptr));
            }
            // pageable source: the copy has left the host buffer when the call returns
            if (pl.mode == 1)
                CUDA_TRY(cudaMemcpyAsync(pop->d_masks.p + static_cast<size_t>(l) * words, masks.data(),
                                         sizeof(unsigned long long) * words, cudaMemcpyHostToDevice,
                                         b->stream));
            if (pl.second == -2 || pl.mode == 3)
                CUDA_TRY(cudaMemcpyAsync(pop->d_aux.p + static_cast<size_t>(l) * words, aux.data(),
                                         sizeof(unsigned long long) * words, cudaMemcpyHostToDevice,
                                         b->stream));
        }
        any_replace |= pl.mode == 2 || pl.mode == 3;
    }
    CUDA_TRY(cudaMemcpyAsync(pop->d_plan.p, pop->plans.data(), sizeof(ReinitPlan) * L,
                             cudaMemcpyHostToDevice, b->stream));

    refresh_params(b);
    b->dev.pop = pop->dev_view();
    b->dev.plan = pop->d_plan.p;
    b->dev.masks = pop->d_masks.p;
    b->dev.aux = pop->d_aux.p;

    CUDA_TRY(cudaEventRecord(b->ev0, b->stream));
    // reinit, device part
    k_reinit<<<L, 256, 0, b->stream>>>(pb->dev, b->dev);
    b->launches += 1;
    if (any_replace) {
        k_pop_replace<<<L, 256, 0, b->stream>>>(pb->dev, b->dev);
        b->launches += 1;
        pop->device_newer = true;
    }
    CUDA_TRY(cudaGetLastError());
    pop->reinit_done = true;
    if (kappa)
        for (int l = 0; l != L; ++l)
            kappa[l] = pop->plans[l].kappa;
    return BXG_OK;
}

int
bxg_optimize_run(bxg_batch* b, bxg_population* pop, int32_t* status, double* value,
                 int32_t* remaining)
{
    if (!b || !pop || pop->pb != b->pb || !pop->reinit_done)
        return fail(BXG_E_INVALID, "bxg_optimize_run: call bxg_optimize_reinit first");
    bxg_problem* pb = b->pb;
    CUDA_TRY(cudaSetDevice(pb->device));
    const int L = b->n_lanes;
    pop->reinit_done = false;
    CUDA_TRY(cudaEventRecord(b->ev0, b->stream)); // the dominant kernel alone (roofline)
    launch_optimize(pb->dev, b->prm, b->dev, b->cfg, b->stream);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(b->ev1, b->stream));
    b->launches += 1;
    pop->device_newer = true;
    pop->rounds += 1;
    int rc = fetch_ctl(b);
    if (rc != BXG_OK)
        return rc;
    CUDA_TRY(cudaEventElapsedTime(&b->last_ms, b->ev0, b->ev1));
    b->total_ms += b->last_ms;

    for (int l = 0; l != L; ++l) {
        const LaneCtl& c = b->h_ctl[l];
        if (status)
            status[l] = c.status;
        if (value)
            value[l] = c.best_has_x ? c.best_value : 0.0;
        if (remaining)
            remaining[l] = c.best_remaining;
    }
    return BXG_OK;
}

int
bxg_optimize_round(bxg_batch* b, bxg_population* pop, int first_round, int32_t* status,
                   double* value, int32_t* remaining)
{
    const int rc = bxg_optimize_reinit(b, pop, first_round, nullptr);
    if (rc != BXG_OK)
        return rc;
    return bxg_optimize_run(b, pop, status, value, remaining);
}

int
bxg_optimize(bxg_batch* b, bxg_population* pop, bxg_comm* comm, double time_limit,
             int64_t max_rounds, int exchange_every, bxg_update_fn update, void* user,
             int64_t* rounds_done)
{
    if (!b || !pop || pop->pb != b->pb)
        return fail(BXG_E_INVALID, "bxg_optimize: batch and population of different problems");
    if (!(time_limit > 0) && max_rounds <= 0)
        return fail(BXG_E_INVALID, "bxg_optimize: neither a time limit nor a number of rounds");
    CUDA_TRY(cudaSetDevice(b->pb->device));
    const auto start = std::chrono::steady_clock::now();
    auto last_report = start;
    auto elapsed = [&] {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    };

    b->deadline_ns = 0;
    if (time_limit > 0) {
        // stop_task for the kernels: the lanes leave their loops at the first
        // sweep boundary past the deadline (device clock)
        unsigned long long now = 0;
        int rc_ = device_clock(b->stream, &now);
        if (rc_ != BXG_OK)
            return rc_;
        b->deadline_ns = now + static_cast<unsigned long long>(time_limit * 1e9);
    }

    int64_t rounds = 0;
    int rc = BXG_OK;
    for (;;) {
        rc = bxg_optimize_round(b, pop, pop->rounds == 0 ? 1 : 0, nullptr, nullptr, nullptr);
        if (rc != BXG_OK)
            break;
        ++rounds;
        if (comm && exchange_every > 0 && rounds % exchange_every == 0) {
            rc = bxg_exchange_best(b, pop, comm, nullptr);
            if (rc != BXG_OK)
                break;
        }
        const auto now = std::chrono::steady_clock::now();
        if (update && std::chrono::duration<double>(now - last_report).count() >= 1.0) {
            // optimize_problem's once-a-second report, itm-optimizer-common.hpp:835-853
            last_report = now;
            rc = pop->fetch_scores(b->stream);
            if (rc != BXG_OK)
                break;
            const bxg_population::member& best = pop->data[pop->order[0]];
            update(best.remaining, best.value, best.loop, elapsed(), pop->rounds * b->n_lanes, user);
        }
        if (max_rounds > 0 && rounds >= max_rounds)
            break;
        if (time_limit > 0 && elapsed() >= time_limit)
            break;
    }
    b->deadline_ns = 0;
    if (rounds_done)
        *rounds_done = rounds;
    return rc;
}

int
bxg_population_stats(bxg_population* pop, int64_t* rounds, double* first_solution_seconds)
{
    if (!pop)
        return fail(BXG_E_INVALID, "null population");
    if (rounds)
        *rounds = pop->rounds;
    if (first_solution_seconds) {
        *first_solution_seconds = -1.0;
        if (pop->d_pfirst.p) {
            CUDA_TRY(cudaSetDevice(pop->pb->device));
            CUDA_TRY(cudaDeviceSynchronize());
            unsigned long long ns = 0;
            CUDA_TRY(cudaMemcpy(&ns, pop->d_pfirst.p, sizeof(ns), cudaMemcpyDeviceToHost));
            if (ns != ~0ull)
                *first_solution_seconds = static_cast<double>(ns) * 1e-9;
        }
    }
    return BXG_OK;
}

uint64_t
bxg_bit_array_hash(int n, const uint64_t* x_words)
{
    return bxg_population::hash_of(x_words, n / 64 + 1);
}


// ---------------------------------------------------------------------------
// multi-GPU: the incumbent exchange
// ---------------------------------------------------------------------------
} // extern "C"

#include <dlfcn.h>
#include <nccl.h>

namespace {

// NCCL is bound at run time: a process that never calls bxg_comm_* does not
// need the library, and inside a torch process the already loaded
// libnccl.so.2 is the one that answers.
struct nccl_api
{
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;

    bool load()
    {
        if (handle)
            return true;
        const char* names[] = { "libnccl.so.2", "libnccl.so" };
        for (const char* n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle)
                break;
        }
        if (!handle) {
            error = std::string("libnccl.so.2 not found: ") + dlerror();
            return false;
        }
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(dlsym(handle, "ncclGetUniqueId"));
        CommInitRank = reinterpret_cast<decltype(CommInitRank)>(dlsym(handle, "ncclCommInitRank"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(handle, "ncclCommDestroy"));
        AllGather = reinterpret_cast<decltype(AllGather)>(dlsym(handle, "ncclAllGather"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(handle, "ncclGetErrorString"));
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllGather || !GetErrorString) {
            error = "libnccl.so.2 lacks an entry point";
            handle = nullptr;
            return false;
        }
        return true;
    }
};

nccl_api g_nccl;

#define NCCL_TRY(expr)                                                                           \
    do {                                                                                         \
        ncclResult_t r_ = (expr);                                                                \
        if (r_ != ncclSuccess)                                                                   \
            return fail(BXG_E_CUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(r_));      \
    } while (0)

// What a rank sends: its best member.
struct IncumbentHdr
{
    int remaining;
    int rank;
    double value;
    unsigned long long hash;
    long long loop;
};

// storage::sort's comparator, itm-optimizer-common.hpp:427-441
struct MemberLess
{
    const int* rem;
    const double* val;
    bool maximize;
    __device__ bool operator()(const int& a, const int& b) const
    {
        if (rem[a] < rem[b])
            return true;
        if (rem[a] == rem[b])
            return maximize ? val[a] > val[b] : val[a] < val[b];
        return false;
    }
};

// best member of the local population -> send buffer
__global__ void __launch_bounds__(256)
k_pack_incumbent(const __grid_constant__ DevPop pop, int xw32, int rank, unsigned char* out)
{
    const int id = pop.order[0];
    if (threadIdx.x == 0) {
        IncumbentHdr h;
        h.remaining = pop.remaining[id];
        h.rank = rank;
        h.value = pop.value[id];
        h.hash = pop.hash[id];
        h.loop = pop.loop[id];
        *reinterpret_cast<IncumbentHdr*>(out) = h;
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + sizeof(IncumbentHdr));
    const uint32_t* src = pop.x + (size_t)id * xw32;
    for (int w = threadIdx.x; w < xw32; w += blockDim.x)
        dst[w] = src[w];
}

// The box-wide incumbent (fewest remaining constraints, then the better
// objective, then the lowest rank: raw_result's order,
// itm-optimizer-common.hpp:427-441) replaces the worst member of the local
// population unless the population already holds it; the order is rebuilt with
// storage::sort.  One CTA.
__global__ void __launch_bounds__(256)
k_adopt_incumbent(const __grid_constant__ DevPop pop, int xw32, int rank, int nranks, int maximize,
                  const unsigned char* gathered, size_t stride, int* winner_out)
{
    __shared__ int s_win, s_take, s_slot;
    __shared__ double s_val[256];
    __shared__ int s_rem[256];
    __shared__ int s_ord[256];
    const int tid = threadIdx.x, T = blockDim.x;
    if (tid == 0) {
        int win = 0;
        IncumbentHdr best = *reinterpret_cast<const IncumbentHdr*>(gathered);
        for (int r = 1; r < nranks; ++r) {
            const IncumbentHdr h = *reinterpret_cast<const IncumbentHdr*>(gathered + stride * r);
            bool better = false;
            if (h.remaining != best.remaining)
                better = h.remaining < best.remaining;
            else if (h.value != best.value)
                better = maximize ? h.value > best.value : h.value < best.value;
            if (better) {
                best = h;
                win = r;
            }
        }
        s_win = win;
        int take = win != rank;
        for (int i = 0; take && i < pop.size; ++i)
            if (pop.remaining[i] == best.remaining && pop.hash[i] == best.hash &&
                (best.remaining != 0 || pop.value[i] == best.value))
                take = 0;
        s_take = take;
        const int slot = pop.order[pop.size - 1];
        s_slot = slot;
        if (take) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            pop.remaining[slot] = best.remaining;
            pop.value[slot] = best.value;
            pop.hash[slot] = best.hash;
            pop.loop[slot] = best.loop;
            pop.duration[slot] = (double)(now - pop.t0_ns) * 1e-9;
        }
        if (winner_out)
            *winner_out = win;
    }
    __syncthreads();
    if (!s_take)
        return;
    const uint32_t* src =
      reinterpret_cast<const uint32_t*>(gathered + stride * s_win + sizeof(IncumbentHdr));
    uint32_t* dst = pop.x + (size_t)s_slot * xw32;
    for (int w = tid; w < xw32; w += T)
        dst[w] = src[w];
    for (int i = tid; i < pop.size; i += T) {
        s_val[i] = pop.value[i];
        s_rem[i] = pop.remaining[i];
        s_ord[i] = pop.order[i];
    }
    __syncthreads();
    if (tid == 0) {
        seq_std_sort(Plain<int>{ s_ord }, pop.size, MemberLess{ s_rem, s_val, maximize != 0 });
    }
    __syncthreads();
    for (int i = tid; i < pop.size; i += T)
        pop.order[i] = s_ord[i];
}

} // namespace

struct bxg_comm
{
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, device = 0;
    DevBuf<unsigned char> d_send, d_recv;
    DevBuf<int> d_winner;
    size_t stride = 0;
    int64_t exchanges = 0;
};

extern "C" {

int
bxg_comm_unique_id(void* id128)
{
    if (!id128)
        return fail(BXG_E_INVALID, "bxg_comm_unique_id: null buffer");
    if (!g_nccl.load())
        return fail(BXG_E_UNSUPPORTED, g_nccl.error);
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof(id));
    return BXG_OK;
}

int
bxg_comm_init(const void* id128, int rank, int nranks, int device, bxg_comm** out)
{
    if (!id128 || !out || nranks < 1 || rank < 0 || rank >= nranks)
        return fail(BXG_E_INVALID, "bxg_comm_init: bad argument");
    if (!g_nccl.load())
        return fail(BXG_E_UNSUPPORTED, g_nccl.error);
    CUDA_TRY(cudaSetDevice(device));
    std::unique_ptr<bxg_comm> c(new (std::nothrow) bxg_comm);
    if (!c)
        return fail(BXG_E_MEMORY, "out of host memory");
    c->rank = rank;
    c->nranks = nranks;
    c->device = device;
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    NCCL_TRY(g_nccl.CommInitRank(&c->comm, nranks, id, rank));
    CUDA_TRY(c->d_winner.alloc(1));
    *out = c.release();
    return BXG_OK;
}

void
bxg_comm_destroy(bxg_comm* c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->comm && g_nccl.handle)
        g_nccl.CommDestroy(c->comm);
    delete c;
}

int
bxg_comm_info(const bxg_comm* c, int* rank, int* nranks, int64_t* exchanges)
{
    if (!c)
        return fail(BXG_E_INVALID, "null communicator");
    if (rank)
        *rank = c->rank;
    if (nranks)
        *nranks = c->nranks;
    if (exchanges)
        *exchanges = c->exchanges;
    return BXG_OK;
}

int
bxg_exchange_best(bxg_batch* b, bxg_population* pop, bxg_comm* c, int* winner_rank)
{
    if (!b || !pop || !c || pop->pb != b->pb)
        return fail(BXG_E_INVALID, "bxg_exchange_best: bad argument");
    if (pop->size > 256)
        return fail(BXG_E_UNSUPPORTED, "bxg_exchange_best: population larger than 256");
    bxg_problem* pb = b->pb;
    CUDA_TRY(cudaSetDevice(pb->device));
    if (pop->host_newer) {
        int rc = pop->to_device(b->stream);
        if (rc != BXG_OK)
            return rc;
    }
    const int xw32 = pb->dev.xw32;
    const size_t stride = (sizeof(IncumbentHdr) + sizeof(uint32_t) * xw32 + 15) & ~size_t(15);
    if (c->stride != stride) {
        CUDA_TRY(c->d_send.alloc(stride));
        CUDA_TRY(c->d_recv.alloc(stride * c->nranks));
        CUDA_TRY(cudaStreamSynchronize(nullptr));
        c->stride = stride;
    }
    // everything on the batch's stream, no host synchronisation: the next
    // round's kernels queue behind the adoption
    const DevPop view = pop->dev_view();
    k_pack_incumbent<<<1, 256, 0, b->stream>>>(view, xw32, c->rank, c->d_send.p);
    NCCL_TRY(g_nccl.AllGather(c->d_send.p, c->d_recv.p, stride, ncclChar, c->comm, b->stream));
    k_adopt_incumbent<<<1, 256, 0, b->stream>>>(view, xw32, c->rank, c->nranks, pb->maximize,
                                                c->d_recv.p, stride, c->d_winner.p);
    CUDA_TRY(cudaGetLastError());
    b->launches += 2;
    c->exchanges += 1;
    pop->device_newer = true;
    if (winner_rank) {
        CUDA_TRY(cudaMemcpyAsync(winner_rank, c->d_winner.p, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
        CUDA_TRY(cudaStreamSynchronize(b->stream));
    }
    return BXG_OK;
}

} // extern "C"

#endif
