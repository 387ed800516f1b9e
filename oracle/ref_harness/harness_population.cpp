// harness_population.cpp -- the reference's optimizer population behind a C
// interface (see bxref_pop.h).  TEST INFRASTRUCTURE: instantiates
// best_solution_recorder / storage / local_context of
// /root/reference/lib/src/itm-optimizer-common.hpp unmodified.
#include "harness.hpp"

#include "itm-optimizer-common.hpp"

#include "bxref_pop.h"

using namespace baryonyx;

namespace {

using cost_t = itm::default_cost_type<double>;

struct pop_base
{
    virtual ~pop_base() = default;
    virtual int size() = 0;
    virtual void member(int i, uint64_t* x, double* value, int* remaining, uint64_t* hash,
                        int64_t* loop) = 0;
    virtual double reinit(itm::local_context& c, double kmin, double kmax, bit_array& x) = 0;
    virtual void mutation(itm::local_context& c, bit_array& x) = 0;
    virtual void crossover(itm::local_context& c, bit_array& x) = 0;
    virtual void try_update(itm::local_context& c, const bit_array& x, double value, long loop) = 0;
    virtual void try_advance(itm::local_context& c, const bit_array& x, int remaining, long loop) = 0;

    bxref::handle* h = nullptr;
};

template<typename Mode>
struct pop_impl : pop_base
{
    random_engine rng;
    cost_t cost;
    itm::best_solution_recorder<cost_t, double, Mode> rec;

    pop_impl(bxref::handle& hd, uint64_t master_seed, int threads)
      : rng(static_cast<random_engine::result_type>(master_seed))
      , cost(hd.pb.objective, hd.n)
      , rec(rng, static_cast<unsigned>(threads), cost, hd.pb.objective.value, hd.n, hd.csts,
            hd.ctx.parameters.init_population_size)
    {
        h = &hd;
    }

    int size() override { return h->ctx.parameters.init_population_size; }

    void member(int i, uint64_t* x, double* value, int* remaining, uint64_t* hash,
                int64_t* loop) override
    {
        const auto& r = rec.get_best(i);
        if (x)
            for (int b = 0, e = r.x.block_size(); b != e; ++b)
                x[b] = r.x.block(b);
        if (value)
            *value = r.value;
        if (remaining)
            *remaining = r.remaining_constraints;
        if (hash)
            *hash = r.hash;
        if (loop)
            *loop = r.loop;
    }

    double reinit(itm::local_context& c, double kmin, double kmax, bit_array& x) override
    {
        return rec.reinit(c, false, kmin, kmax, x);
    }
    void mutation(itm::local_context& c, bit_array& x) override { rec.mutation(c, x); }
    void crossover(itm::local_context& c, bit_array& x) override { rec.crossover(c, x); }
    void try_update(itm::local_context& c, const bit_array& x, double value, long loop) override
    {
        rec.try_update(c, x, value, loop);
    }
    void try_advance(itm::local_context& c, const bit_array& x, int remaining, long loop) override
    {
        rec.try_advance(c, x, remaining, loop);
    }
};

bit_array
to_bits(int n, const uint64_t* w)
{
    bit_array x(n);
    for (int b = 0, e = x.block_size(); b != e; ++b)
        x.set_block(b, w[b]);
    return x;
}

void
from_bits(const bit_array& x, uint64_t* w)
{
    for (int b = 0, e = x.block_size(); b != e; ++b)
        w[b] = x.block(b);
}

} // namespace

extern "C" {

void*
bxref_pop_create(void* hv, uint64_t master_seed, int threads, uint32_t* rng_state)
{
    auto* h = static_cast<bxref::handle*>(hv);
    pop_base* p;
    if (h->pb.type == objective_function_type::maximize) {
        auto* q = new pop_impl<itm::maximize_tag>(*h, master_seed, threads);
        if (rng_state)
            *rng_state = bxref::rng_state_of(q->rng);
        p = q;
    } else {
        auto* q = new pop_impl<itm::minimize_tag>(*h, master_seed, threads);
        if (rng_state)
            *rng_state = bxref::rng_state_of(q->rng);
        p = q;
    }
    return p;
}

void
bxref_pop_destroy(void* pop)
{
    delete static_cast<pop_base*>(pop);
}

int
bxref_pop_size(void* pop)
{
    return static_cast<pop_base*>(pop)->size();
}

void
bxref_pop_member(void* pop, int i, uint64_t* x_words, double* value, int* remaining,
                 uint64_t* hash, int64_t* loop)
{
    static_cast<pop_base*>(pop)->member(i, x_words, value, remaining, hash, loop);
}

void*
bxref_pop_ctx_create(void* pop, unsigned thread_id, uint64_t seed)
{
    auto* p = static_cast<pop_base*>(pop);
    return new itm::local_context(p->h->ctx, thread_id,
                                  static_cast<random_engine::result_type>(seed));
}

void
bxref_pop_ctx_destroy(void* ctx)
{
    delete static_cast<itm::local_context*>(ctx);
}

uint32_t
bxref_pop_ctx_rng(void* ctx)
{
    return bxref::rng_state_of(static_cast<itm::local_context*>(ctx)->rng);
}

double
bxref_pop_reinit(void* pop, void* ctx, double kappa_min, double kappa_max, uint64_t* x_words)
{
    auto* p = static_cast<pop_base*>(pop);
    auto x = to_bits(p->h->n, x_words);
    const double k =
      p->reinit(*static_cast<itm::local_context*>(ctx), kappa_min, kappa_max, x);
    from_bits(x, x_words);
    return k;
}

void
bxref_pop_mutation(void* pop, void* ctx, uint64_t* x_words)
{
    auto* p = static_cast<pop_base*>(pop);
    auto x = to_bits(p->h->n, x_words);
    p->mutation(*static_cast<itm::local_context*>(ctx), x);
    from_bits(x, x_words);
}

void
bxref_pop_crossover(void* pop, void* ctx, uint64_t* x_words)
{
    auto* p = static_cast<pop_base*>(pop);
    auto x = to_bits(p->h->n, x_words);
    p->crossover(*static_cast<itm::local_context*>(ctx), x);
    from_bits(x, x_words);
}

void
bxref_pop_try_update(void* pop, void* ctx, const uint64_t* x_words, double value, int64_t loop)
{
    auto* p = static_cast<pop_base*>(pop);
    p->try_update(*static_cast<itm::local_context*>(ctx), to_bits(p->h->n, x_words), value,
                  static_cast<long>(loop));
}

void
bxref_pop_try_advance(void* pop, void* ctx, const uint64_t* x_words, int remaining,
                      int64_t loop)
{
    auto* p = static_cast<pop_base*>(pop);
    p->try_advance(*static_cast<itm::local_context*>(ctx), to_bits(p->h->n, x_words), remaining,
                   static_cast<long>(loop));
}

uint64_t
bxref_bit_array_hash(int n, const uint64_t* x_words)
{
    return bit_array_hash()(to_bits(n, x_words));
}

} // extern "C"
