// harness.hpp -- shared declarations of the reference harness (test
// infrastructure; see bxref.h).  Compiled only together with the reference's
// own headers from /root/reference/lib/{include,src}.
#ifndef BXREF_HARNESS_HPP
#define BXREF_HARNESS_HPP

#include <baryonyx/core>

#include "itm-common.hpp"
#include "private.hpp"
#include "problem.hpp"

#include <atomic>
#include <functional>
#include <memory>
#include <sstream>
#include <vector>

#include "bxref.h"

namespace bxref {

struct handle
{
    baryonyx::context ctx;
    baryonyx::solver_parameters orig; // as given, before clamping
    baryonyx::raw_problem raw;
    baryonyx::problem pb;
    std::vector<baryonyx::itm::merged_constraint> csts;
    int m = 0;
    int n = 0;
    int nnz = 0;

    handle()
      : ctx(3)
    {}
};

// Abstract step-level run; one concrete type per (Solver, Mode).
struct run_base
{
    virtual ~run_base() = default;
    virtual int sweep(int push, double kappa, double obj_amp) = 0;
    virtual void state(uint64_t* x_words,
                       double* pi,
                       double* P,
                       uint32_t* rng_state,
                       int* order_list,
                       int* order_len,
                       double* norm_cost) = 0;
    virtual void set_x(const uint64_t* x_words) = 0;
    virtual void set_rng(uint32_t state) = 0;
    virtual double objective() = 0;
    virtual void info(double* delta, double* theta) = 0;
    virtual double tail_value() = 0;
};

struct solve_out
{
    int status = 0;
    std::vector<int8_t> x;
    double value = 0;
    int remaining = 0;
    int64_t updates = 0;
    int64_t sweeps = 0;
    double seconds = 0;
};

// trace of the optimizer at thread = 1 (filled by harness_solver.hpp's tracing<>)
struct trace_member
{
    std::vector<uint64_t> x;
    double value;
    int remaining;
    uint64_t hash;
    int64_t loop;
};

struct trace_round
{
    std::vector<uint64_t> x;
    uint32_t rng;
    double kappa;
    std::vector<double> pi, P;
    std::vector<trace_member> members;
};

struct trace_data
{
    std::vector<trace_round> rounds; // [rounds + 1]: state at the first sweep of each restart
    std::vector<int> sweep_round, sweep_push, sweep_len;
    std::vector<double> sweep_kappa;
    // live part
    const long int* call_number = nullptr;
    long int last_call = 0;
    int max_rounds = 0;
    std::atomic_bool* stop = nullptr;
    std::function<void(trace_round&)> population;
};


// last bxref_optimize: seconds until the population's best first had no violated
// row (1 s granularity: the reference's own report), restarts begun; harness_eq01.cpp
void optimize_report(double* first_feasible_seconds, int64_t* restarts);

// Per-solver factories, defined in harness_<solver>.cpp.
#define BXREF_DECLARE(tag)                                                   \
    run_base* make_run_##tag(handle& h, uint32_t seed);                       \
    void solve_##tag(handle& h, uint32_t seed, solve_out& out);               \
    void optimize_##tag(                                                      \
      handle& h, uint32_t seed, int threads, double tl, solve_out& out);     \
    void trace_##tag(handle& h, uint32_t seed, int rounds, trace_data& t);

BXREF_DECLARE(eq01)
BXREF_DECLARE(ineq01)
BXREF_DECLARE(eq101)
BXREF_DECLARE(ineq101)
BXREF_DECLARE(ineqZ)

inline uint32_t
rng_state_of(const baryonyx::random_engine& rng)
{
    std::ostringstream os;
    os << rng;
    return static_cast<uint32_t>(std::stoul(os.str()));
}

inline void
rng_set_state(baryonyx::random_engine& rng, uint32_t state)
{
    std::istringstream is(std::to_string(state));
    is >> rng;
}

} // namespace bxref

#endif
