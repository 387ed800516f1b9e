// harness_b200.cpp -- TEST INFRASTRUCTURE: compiles the reference-side adapter
// of this repository (integration/itm_b200.cpp) against the reference's own
// headers and exposes it next to the reference's solve, so that a test on the
// GPU box can compare the two through the same result type.
#include "harness.hpp"

#include "../../integration/itm_b200.cpp"

#include "bxref_b200.h"

using namespace baryonyx;

extern "C" int
bxref_solve_b200(void* hv, uint32_t seed, int restarts, int device, int8_t* x, double* value,
                 int* remaining, int64_t* loop)
{
    auto* h = static_cast<bxref::handle*>(hv);
    context ctx(h->ctx);
    ctx.parameters.seed = seed;
    try {
        result r = itm::solve_b200(ctx, h->pb, restarts, device);
        if (remaining)
            *remaining = r.remaining_constraints;
        if (loop)
            *loop = r.loop;
        if (value)
            *value = r.solutions.empty() ? 0.0 : r.solutions.front().value;
        if (x) {
            std::fill_n(x, h->n, int8_t{ 0 });
            if (!r.solutions.empty())
                for (int i = 0, e = std::min<int>(h->n, length(r.solutions.front().variables)); i != e;
                     ++i)
                    x[i] = static_cast<int8_t>(r.solutions.front().variables[i]);
        }
        return static_cast<int>(r.status);
    } catch (const solver_failure& e) {
        return -100 - static_cast<int>(e.failure());
    }
}

// baryonyx::itm::optimize_b200 with the context's callbacks counted.
extern "C" int
bxref_optimize_b200(void* hv, uint32_t seed, int workers, int device, double time_limit,
                    int64_t max_rounds, int storage, int8_t* x, double* value, int* remaining,
                    int* n_solutions, int* calls /* start, update, finish */)
{
    auto* h = static_cast<bxref::handle*>(hv);
    context ctx(h->ctx);
    ctx.parameters.seed = seed;
    ctx.parameters.time_limit = time_limit;
    ctx.parameters.storage = static_cast<solver_parameters::storage_type>(storage);
    int started = 0, updated = 0, finished = 0;
    ctx.start = [&](const solver_parameters&) { ++started; };
    ctx.update = [&](int, double, long int, double, long int) { ++updated; };
    ctx.finish = [&](const result&) { ++finished; };
    try {
        result r = itm::optimize_b200(ctx, h->pb, workers, device, static_cast<long int>(max_rounds));
        if (remaining)
            *remaining = r.remaining_constraints;
        if (n_solutions)
            *n_solutions = length(r.solutions);
        if (value)
            *value = r.solutions.empty() ? 0.0 : r.solutions.front().value;
        if (x) {
            std::fill_n(x, h->n, int8_t{ 0 });
            if (!r.solutions.empty())
                for (int i = 0, e = std::min<int>(h->n, length(r.solutions.front().variables)); i != e;
                     ++i)
                    x[i] = static_cast<int8_t>(r.solutions.front().variables[i]);
        }
        if (calls) {
            calls[0] = started;
            calls[1] = updated;
            calls[2] = finished;
        }
        return static_cast<int>(r.status);
    } catch (const solver_failure& e) {
        return -100 - static_cast<int>(e.failure());
    }
}
