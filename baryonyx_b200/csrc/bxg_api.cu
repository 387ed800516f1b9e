// bxg_api.cu -- C ABI (include/bxg.h) and host-side set-up of the engine.
//
// Host work is limited to what the reference also does once per problem or
// once per run on the host: building the sparse index, the solver bounds, the
// cost normalisation, delta, and copying state in and out.  Every sweep,
// update, scan and objective evaluation runs on the device; nothing here
// falls back to the CPU when CUDA is unavailable.
#include "../../include/bxg.h"
#include "bxg_kernels.cuh"

#include <algorithm>
#include <chrono>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <new>
#include <random>
#include <sstream>
#include <string>
#include <vector>

using namespace bxg;

namespace {

thread_local std::string g_error;

int
fail(int code, const std::string& msg)
{
    g_error = msg;
    return code;
}

// No exception crosses the C boundary: an entry point whose host side
// allocates (std::vector, std::string) runs its body inside this; an
// allocation failure becomes BXG_E_MEMORY, anything else BXG_E_INVALID.
template<typename Body>
int
guarded(const char* who, Body&& body)
{
    try {
        return body();
    } catch (const std::bad_alloc&) {
        return fail(BXG_E_MEMORY, std::string(who) + ": out of host memory");
    } catch (const std::exception& e) {
        return fail(BXG_E_INVALID, std::string(who) + ": " + e.what());
    } catch (...) {
        return fail(BXG_E_INVALID, std::string(who) + ": unexpected failure");
    }
}

#define CUDA_TRY(expr)                                                        \
    do {                                                                      \
        cudaError_t e_ = (expr);                                              \
        if (e_ != cudaSuccess)                                                \
            return fail(BXG_E_CUDA, std::string(#expr) + ": " +               \
                                      cudaGetErrorString(e_));                \
    } while (0)

// Device memory comes from the device's stream-ordered pool, told to keep
// what is freed: creating and destroying a batch (0.5 GB for the bench's
// 1024 lanes) costs microseconds instead of the tens of milliseconds a
// cudaFree / cudaMalloc pair can take, so a caller may build a problem and a
// batch per solve.  Allocation and release are ordered on the legacy default
// stream; the create functions synchronise it before a batch's own
// (non-blocking) stream touches the memory, and every entry point drains that
// stream before it returns, so nothing is in flight at release.
inline void
pool_keep_freed_memory()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess)
        return;
    static bool done[64] = {};
    if (dev < 0 || dev >= 64 || done[dev])
        return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done[dev] = true;
}

template<typename T>
struct DevBuf
{
    T* p = nullptr;
    size_t n = 0;

    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }

    void release()
    {
        if (p)
            cudaFreeAsync(p, nullptr);
        p = nullptr;
        n = 0;
    }

    cudaError_t alloc(size_t count)
    {
        release();
        pool_keep_freed_memory();
        n = count;
        return cudaMallocAsync(reinterpret_cast<void**>(&p), sizeof(T) * (count ? count : 1), nullptr);
    }

    cudaError_t upload(const std::vector<T>& v)
    {
        cudaError_t e = alloc(v.size());
        if (e != cudaSuccess)
            return e;
        if (!v.empty())
            e = cudaMemcpy(p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice);
        return e;
    }

    size_t bytes() const
    {
        return sizeof(T) * n;
    }
};

} // namespace

struct bxg_problem
{
    int device = 0;
    int m = 0, n = 0, nnz = 0, max_row = 0;
    int problem_type = 0, family = 0, maximize = 0;
    double cost_constant = 0;

    // host copies (taps, host-side initialisation policies)
    std::vector<int> row_ptr, row_col, row_a, row_cscpos;
    std::vector<int> col_ptr, csc_row, csc_a, csc_val, csc_rs;
    std::vector<int4> row_info;
    std::vector<int2> gp;
    std::vector<RowHdr> row_hdr;
    int p95_row = 0;
    std::vector<int> row_gptr, row_eptr;
    int max_col = 0, max_S = 0;
    std::vector<int> cmin, cmax, bmin, bmax, row_neg, ztype;
    std::vector<int> ex_start, ex_count, ex_flat; // (Z) exhaustive rows: pattern masks
    int max_zrow = 0;
    int max_amount = 0;
    std::vector<float> row_bytes;
    std::vector<int> obj_var;
    std::vector<double> obj_factor;
    std::vector<double> cost; // default_cost_type::linear_elements
    int obj_any_order = 0;
    int tail_rows = 0; // rows that can read R[r_size]

    DevBuf<int> d_row_ptr, d_row_col, d_row_a, d_row_cscpos, d_col_ptr, d_csc_row,
      d_csc_a, d_csc_rs, d_bmin, d_bmax, d_row_neg, d_ztype, d_obj_var, d_ex_start, d_ex_count,
      d_ex_flat;
    DevBuf<int4> d_row_info;
    DevBuf<int2> d_gp;
    DevBuf<RowHdr> d_row_hdr;
    DevBuf<int> d_row_gptr, d_row_eptr;
    DevBuf<float> d_row_bytes;
    DevBuf<double> d_obj_factor, d_cost;

    DevProblem dev{};
};

struct bxg_batch
{
    bxg_problem* pb = nullptr;
    bxg_params p{};
    int n_lanes = 0;
    int order_cap = 0;
    double delta = 0, theta = 0;
    double tail_value = 0;
    std::vector<double> norm_cost, powtab;

    DevBuf<double> d_P, d_pi, d_norm_cost, d_powtab, d_kappa, d_Rv, d_lane_par;
    std::vector<double> lane_par; // [lanes][5] or empty (bxg_batch_set_lane_params)
    DevBuf<unsigned char> d_scratch, d_zscratch;
    DevBuf<uint32_t> d_x, d_bestx, d_seeds, d_x0, d_rng0;
    DevBuf<int> d_oR, d_act;
    DevBuf<int2> d_mo;
    DevBuf<LaneCtl> d_ctl;

    DevBatch dev{};
    DevParams prm{};
    LaunchCfg cfg{};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0;
    double total_ms = 0; // device time of all timed launches since the counters were reset
    unsigned long long deadline_ns = 0; // stop_task for the kernels (bxg_optimize / time_limit)
    int64_t launches = 0;
    bool begun = false;
    std::vector<LaneCtl> h_ctl;

    ~bxg_batch()
    {
        if (stream)
            cudaStreamSynchronize(stream); // nothing in flight when the buffers go back to the pool
        if (ev0)
            cudaEventDestroy(ev0);
        if (ev1)
            cudaEventDestroy(ev1);
        if (stream)
            cudaStreamDestroy(stream);
    }
};

namespace {

__global__ void
k_globaltimer(unsigned long long* out)
{
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    *out = now;
}

// %globaltimer of the device (nanoseconds): what the kernels compare their
// deadline with (time_limit, stop_task)
int
device_clock(cudaStream_t stream, unsigned long long* now)
{
    unsigned long long* d_now = nullptr;
    CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&d_now), sizeof(unsigned long long), stream));
    k_globaltimer<<<1, 1, 0, stream>>>(d_now);
    CUDA_TRY(cudaMemcpyAsync(now, d_now, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    CUDA_TRY(cudaFreeAsync(d_now, stream));
    return BXG_OK;
}

int
fetch_ctl(bxg_batch* b)
{
    b->h_ctl.resize(b->n_lanes);
    CUDA_TRY(cudaMemcpyAsync(b->h_ctl.data(), b->d_ctl.p, sizeof(LaneCtl) * b->n_lanes,
                             cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return BXG_OK;
}

void
refresh_params(bxg_batch* b)
{
    const bxg_params& p = b->p;
    DevParams& q = b->prm;
    q.theta = b->theta;
    q.delta = b->delta;
    q.kappa_min = p.kappa_min;
    q.kappa_step = p.kappa_step;
    q.kappa_max = p.kappa_max;
    q.pushing_k_factor = p.pushing_k_factor;
    q.pushing_objective_amplifier = p.pushing_objective_amplifier;
    q.w = p.w;
    q.limit = p.limit;
    q.pushes_limit = p.pushes_limit;
    q.pushing_iteration_limit = p.pushing_iteration_limit;
    q.order = p.order;
    q.tail_value = b->tail_value;
    q.deadline_ns = b->deadline_ns;
    q.cost = b->d_norm_cost.p;
    q.powtab = b->d_powtab.p;
}

} // namespace

extern "C" {

int
bxg_version(void)
{
    return BXG_VERSION;
}

const char*
bxg_last_error(void)
{
    return g_error.c_str();
}

int
bxg_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// solver_parameters defaults, lib/include/baryonyx/core:410-461
void
bxg_default_params(bxg_params* p)
{
    std::memset(p, 0, sizeof(*p));
    p->time_limit = -1;
    p->theta = 0.5;
    p->delta = -1;
    p->kappa_min = 0;
    p->kappa_step = 1.e-3;
    p->kappa_max = 0.6;
    p->alpha = 1;
    p->pushing_k_factor = 0.9;
    p->pushing_objective_amplifier = 5;
    p->init_policy_random = 0.5;
    p->init_crossover_bastert_insertion = 0.01;
    p->init_crossover_solution_selection_mean = 0.0;
    p->init_crossover_solution_selection_stddev = 0.3;
    p->init_mutation_variable_mean = 0.0001;
    p->init_mutation_variable_stddev = 0.001;
    p->init_mutation_value_mean = 0.5;
    p->init_mutation_value_stddev = 0.2;
    p->init_kappa_improve_start = 0;
    p->init_kappa_improve_increase = 0.02;
    p->init_kappa_improve_stop = 0.2;
    p->w = 0.05;
    p->limit = 1000;
    p->pushes_limit = 100;
    p->pushing_iteration_limit = 50;
    p->init_population_size = 100;
    p->float_type = 1; // double_type
    p->cost_norm = 4;  // loo
    p->preprocessor = 1;
}

// context_set_solver_parameters, lib/src/private.cpp:51-172: the context
// keeps its default for every field that fails the reference's test.  This is
// a validation table and has to be the reference's condition by condition
// (same accepted ranges, same fallback values), so it reads like its source;
// everything else in this file is this repository's own design.
void
bxg_clamp_params(const bxg_params* in, bxg_params* out)
{
    bxg_params c;
    bxg_default_params(&c);
    const bxg_params& s = *in;
    auto unit = [](double v) { return std::clamp(v, 0.0, 1.0); };

    c.time_limit = s.time_limit <= 0 ? std::numeric_limits<double>::infinity() : s.time_limit;
    if (s.theta >= 0 && s.theta <= 1)
        c.theta = s.theta;
    if (s.delta == -1 || s.delta >= 0)
        c.delta = s.delta;
    if (s.kappa_min < s.kappa_max && s.kappa_min >= 0 && s.kappa_min < 1 && s.kappa_max <= 1 &&
        s.kappa_step >= 0 && s.kappa_step < 1) {
        c.kappa_min = s.kappa_min;
        c.kappa_step = s.kappa_step;
        c.kappa_max = s.kappa_max;
    }
    if (s.alpha >= 0)
        c.alpha = s.alpha;
    if (std::isfinite(s.pushing_k_factor))
        c.pushing_k_factor = s.pushing_k_factor;
    if (std::isfinite(s.pushing_objective_amplifier))
        c.pushing_objective_amplifier = s.pushing_objective_amplifier;
    if (std::isfinite(s.init_crossover_bastert_insertion))
        c.init_crossover_bastert_insertion = unit(s.init_crossover_bastert_insertion);
    if (std::isfinite(s.init_crossover_solution_selection_mean))
        c.init_crossover_solution_selection_mean = unit(s.init_crossover_solution_selection_mean);
    if (std::isfinite(s.init_crossover_solution_selection_stddev))
        c.init_crossover_solution_selection_stddev =
          unit(s.init_crossover_solution_selection_stddev);
    if (std::isfinite(s.init_mutation_variable_mean))
        c.init_mutation_variable_mean = unit(s.init_mutation_variable_mean);
    if (std::isfinite(s.init_mutation_variable_stddev))
        c.init_mutation_variable_stddev = unit(s.init_mutation_variable_stddev);
    if (std::isfinite(s.init_mutation_value_mean))
        c.init_mutation_value_mean = unit(s.init_mutation_value_mean);
    if (std::isfinite(s.init_mutation_value_stddev))
        c.init_mutation_value_stddev = unit(s.init_mutation_value_stddev);
    if (s.init_policy_random >= 0 && s.init_policy_random <= 1)
        c.init_policy_random = s.init_policy_random;

    c.seed = s.seed;
    c.thread = s.thread;
    c.limit = s.limit <= 0 ? std::numeric_limits<int64_t>::max() : s.limit;
    if (s.print_level >= 0)
        c.print_level = s.print_level;

    if (s.w < 0) {
        c.w = 0;
    } else if (s.limit > 0) {
        const double limit = static_cast<double>(s.limit);
        c.w = s.w < 1 ? limit * s.w : std::min(limit, s.w);
    } else {
        c.w = std::floor(s.w);
    }

    c.pushes_limit = s.pushes_limit <= 0 ? 0 : s.pushes_limit;
    c.pushing_iteration_limit = s.pushing_iteration_limit <= 0 ? 0 : s.pushing_iteration_limit;
    if (s.init_population_size >= 5)
        c.init_population_size = s.init_population_size;

    c.init_kappa_improve_start = unit(s.init_kappa_improve_start);
    c.init_kappa_improve_increase = unit(s.init_kappa_improve_increase);
    c.init_kappa_improve_stop =
      std::clamp(s.init_kappa_improve_stop, c.init_kappa_improve_start, 1.0);

    c.pre_order = s.pre_order;
    c.order = s.order;
    c.float_type = s.float_type;
    c.init_policy = s.init_policy;
    c.cost_norm = s.cost_norm;
    c.mode = s.mode;
    c.preprocessor = s.preprocessor;
    c.observer = s.observer;
    c.storage = s.storage;
    c.solver = s.solver;
    c.debug = s.debug;
    *out = c;
}

// problem::which_problem_type, lib/src/problem.hpp:72-122
int
bxg_which_problem_type(int nnz, const int32_t* factor, int has_inequality)
{
    int coefficient = 0;
    for (int i = 0; i != nnz; ++i) {
        if (factor[i] == -1 && coefficient == 0)
            coefficient = 1;
        else if (factor[i] <= -2 || factor[i] >= 2) {
            coefficient = 2;
            break;
        }
    }
    static const int eq[3] = { BXG_EQUALITIES_01, BXG_EQUALITIES_101, BXG_EQUALITIES_Z };
    static const int ineq[3] = { BXG_INEQUALITIES_01, BXG_INEQUALITIES_101, BXG_INEQUALITIES_Z };
    return has_inequality ? ineq[coefficient] : eq[coefficient];
}

static int
problem_create_body(int m, int n, const int32_t* row_ptr, const int32_t* col,
                    const int32_t* factor, const int32_t* cmin, const int32_t* cmax,
                    int problem_type, int maximize, int n_obj, const int32_t* obj_var,
                    const double* obj_factor, double cost_constant, int device,
                    bxg_problem** out);

int
bxg_problem_create(int m, int n, const int32_t* row_ptr, const int32_t* col,
                   const int32_t* factor, const int32_t* cmin, const int32_t* cmax,
                   int problem_type, int maximize, int n_obj, const int32_t* obj_var,
                   const double* obj_factor, double cost_constant, int device,
                   bxg_problem** out)
{
    // no exception crosses the C boundary (the host-side index is built with
    // std::vector: an allocation failure becomes BXG_E_MEMORY)
    try {
        return problem_create_body(m, n, row_ptr, col, factor, cmin, cmax, problem_type, maximize,
                                   n_obj, obj_var, obj_factor, cost_constant, device, out);
    } catch (const std::bad_alloc&) {
        return fail(BXG_E_MEMORY, "bxg_problem_create: out of host memory");
    } catch (const std::exception& e) {
        return fail(BXG_E_INVALID, std::string("bxg_problem_create: ") + e.what());
    } catch (...) {
        return fail(BXG_E_INVALID, "bxg_problem_create: unexpected failure");
    }
}

static int
problem_create_body(int m, int n, const int32_t* row_ptr, const int32_t* col,
                    const int32_t* factor, const int32_t* cmin, const int32_t* cmax,
                    int problem_type, int maximize, int n_obj, const int32_t* obj_var,
                    const double* obj_factor, double cost_constant, int device,
                    bxg_problem** out)
{
    if (!out || m <= 0 || n <= 0 || !row_ptr || !col || !factor || !cmin || !cmax ||
        n_obj < 0 || (n_obj && (!obj_var || !obj_factor)))
        return fail(BXG_E_INVALID, "bxg_problem_create: bad argument");
    if (problem_type < BXG_EQUALITIES_01 || problem_type > BXG_INEQUALITIES_Z)
        return fail(BXG_E_INVALID, "bxg_problem_create: unknown problem type");
    // the CSR itself, before anything is sized or copied from it
    if (row_ptr[0] != 0)
        return fail(BXG_E_INVALID, "bxg_problem_create: row_ptr[0] must be 0");
    for (int k = 0; k != m; ++k)
        if (row_ptr[k + 1] <= row_ptr[k])
            return fail(BXG_E_INVALID, "bxg_problem_create: empty constraint or decreasing row_ptr");
    for (int e = 0, nnz_ = row_ptr[m]; e != nnz_; ++e) {
        if (col[e] < 0 || col[e] >= n)
            return fail(BXG_E_INVALID, "bxg_problem_create: column index out of range");
        if (factor[e] == 0)
            return fail(BXG_E_INVALID, "bxg_problem_create: zero factor");
    }

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return fail(BXG_E_CUDA, "bxg_problem_create: no usable CUDA device (this engine has no "
                                "CPU path)");
    }
    CUDA_TRY(cudaSetDevice(device));

    std::unique_ptr<bxg_problem> pb(new (std::nothrow) bxg_problem);
    if (!pb)
        return fail(BXG_E_MEMORY, "out of host memory");

    const int nnz = row_ptr[m];
    pb->device = device;
    pb->m = m;
    pb->n = n;
    pb->nnz = nnz;
    pb->problem_type = problem_type;
    pb->maximize = maximize ? 1 : 0;
    pb->cost_constant = cost_constant;
    pb->family = (problem_type == BXG_EQUALITIES_01 || problem_type == BXG_INEQUALITIES_01)
                   ? FAM_01
                   : (problem_type == BXG_EQUALITIES_101 || problem_type == BXG_INEQUALITIES_101)
                       ? FAM_101
                       : FAM_Z;

    pb->row_ptr.assign(row_ptr, row_ptr + m + 1);
    pb->row_col.assign(col, col + nnz);
    pb->row_a.assign(factor, factor + nnz);
    pb->cmin.assign(cmin, cmin + m);
    pb->cmax.assign(cmax, cmax + m);

    for (int e = 0; e != nnz; ++e) {
        if (col[e] < 0 || col[e] >= n)
            return fail(BXG_E_INVALID, "bxg_problem_create: column index out of range");
        if (factor[e] == 0)
            return fail(BXG_E_INVALID, "bxg_problem_create: zero factor");
        if (pb->family == FAM_01 && factor[e] != 1)
            return fail(BXG_E_INVALID, "bxg_problem_create: 01 problem with factor != 1");
        if (pb->family == FAM_101 && std::abs(factor[e]) != 1)
            return fail(BXG_E_INVALID, "bxg_problem_create: 101 problem with |factor| != 1");
    }

    // CSC index: stable by column, rows ascending inside a column, exactly
    // the order sparse_matrix builds with std::stable_sort
    // (sparse-matrix.hpp:155-174).
    pb->col_ptr.assign(n + 1, 0);
    for (int e = 0; e != nnz; ++e)
        pb->col_ptr[col[e] + 1]++;
    for (int j = 0; j != n; ++j)
        pb->col_ptr[j + 1] += pb->col_ptr[j];
    pb->csc_row.resize(nnz);
    pb->csc_a.resize(nnz);
    pb->csc_val.resize(nnz);
    pb->row_cscpos.resize(nnz);
    {
        std::vector<int> fill(pb->col_ptr.begin(), pb->col_ptr.end() - 1);
        for (int k = 0; k != m; ++k) {
            if (row_ptr[k + 1] <= row_ptr[k])
                return fail(BXG_E_INVALID, "bxg_problem_create: empty constraint");
            for (int e = row_ptr[k]; e != row_ptr[k + 1]; ++e) {
                const int pos = fill[col[e]]++;
                pb->csc_row[pos] = k;
                pb->csc_a[pos] = factor[e];
                pb->csc_val[pos] = e;
                pb->row_cscpos[e] = pos;
            }
        }
    }

    pb->max_row = 0;
    for (int k = 0; k != m; ++k)
        pb->max_row = std::max(pb->max_row, row_ptr[k + 1] - row_ptr[k]);

    // per row element: its column's CSC range, its own CSC position, its
    // column -- one 16-byte load instead of a three-level index chain
    pb->row_info.resize(nnz);
    pb->csc_rs.resize(nnz);
    for (int e = 0; e != nnz; ++e) {
        const int j = col[e];
        pb->row_info[e] = make_int4(pb->col_ptr[j], pb->col_ptr[j + 1], pb->row_cscpos[e], j);
    }
    for (int t = 0; t != nnz; ++t)
        pb->csc_rs[t] = (pb->csc_row[t] << 1) | (pb->csc_a[t] < 0 ? 1 : 0);

    // expanded column entries of every row (see DevProblem::gp)
    {
        long long total = 0;
        for (int j = 0; j != n; ++j) {
            const long long c = pb->col_ptr[j + 1] - pb->col_ptr[j];
            total += c * c;
            pb->max_col = std::max(pb->max_col, static_cast<int>(c));
        }
        if (total > 2000000000ll)
            return fail(BXG_E_MEMORY, "bxg_problem_create: expanded column index exceeds 2^31 entries");
        pb->gp.resize(static_cast<size_t>(total));
        pb->row_gptr.assign(m + 1, 0);
        pb->row_eptr.resize(nnz);
        size_t g = 0;
        for (int k = 0; k != m; ++k) {
            pb->row_gptr[k] = static_cast<int>(g);
            int local = 0;
            for (int e = row_ptr[k]; e != row_ptr[k + 1]; ++e) {
                pb->row_eptr[e] = local;
                const int j = col[e];
                for (int t = pb->col_ptr[j]; t != pb->col_ptr[j + 1]; ++t)
                    pb->gp[g++] = make_int2(t, pb->csc_rs[t]);
                local += pb->col_ptr[j + 1] - pb->col_ptr[j];
            }
            pb->max_S = std::max(pb->max_S, local);
        }
        pb->row_gptr[m] = static_cast<int>(g);
    }

    // Solver bounds (constructor bodies cited in bxg.h)
    pb->bmin.resize(m);
    pb->bmax.resize(m);
    pb->row_neg.assign(m, 0);
    pb->ztype.assign(m, Z_LINEAR);
    pb->ex_start.assign(m, 0);
    pb->ex_count.assign(m, 0);
    pb->row_bytes.resize(m);
    for (int k = 0; k != m; ++k) {
        const int len = row_ptr[k + 1] - row_ptr[k];
        const int* f = factor + row_ptr[k];
        int lower = 0, upper = 0, neg = 0;
        for (int i = 0; i != len; ++i) {
            if (f[i] > 0)
                upper += f[i];
            else {
                lower += f[i];
                ++neg;
            }
        }
        pb->row_neg[k] = neg;

        switch (problem_type) {
        case BXG_EQUALITIES_01:
        case BXG_EQUALITIES_101:
            if (cmin[k] != cmax[k])
                return fail(BXG_E_INVALID, "equality problem with an inequality row");
            pb->bmin[k] = pb->bmax[k] = cmin[k];
            break;
        case BXG_INEQUALITIES_01:
            if (cmin[k] == cmax[k]) {
                pb->bmin[k] = cmin[k];
                pb->bmax[k] = cmax[k];
            } else {
                pb->bmin[k] = std::max(0, cmin[k]);
                pb->bmax[k] = std::min(len, cmax[k]);
            }
            break;
        case BXG_INEQUALITIES_101:
            if (cmin[k] == cmax[k]) {
                pb->bmin[k] = cmin[k];
                pb->bmax[k] = cmax[k];
            } else {
                pb->bmin[k] = std::max(-neg, cmin[k]);
                pb->bmax[k] = std::min(len - neg, cmax[k]);
            }
            break;
        default: {
            // itm-solver-inequalities-Z.cpp:125-183
            int zt = Z_LINEAR;
            for (int i = 0; i != len; ++i)
                if (f[i] < -1 || f[i] > 1)
                    zt = Z_BRANCH_AND_BOUND;
            if (zt == Z_BRANCH_AND_BOUND && len < 10)
                zt = Z_EXHAUSTIVE;

            if (cmin[k] == cmax[k]) {
                pb->bmin[k] = cmin[k];
                pb->bmax[k] = cmax[k];
            } else {
                if (lower >= cmin[k] || upper <= cmax[k]) {
                    if (zt == Z_BRANCH_AND_BOUND)
                        zt = Z_UNCONSTRAINED_BRANCH_AND_BOUND;
                    else if (zt == Z_EXHAUSTIVE)
                        zt = Z_UNCONSTRAINED_EXHAUSTIVE;
                }
                pb->bmin[k] = std::max(lower, cmin[k]);
                pb->bmax[k] = std::min(upper, cmax[k]);
            }
            pb->ztype[k] = zt;

            if (zt != Z_LINEAR) {
                pb->max_zrow = std::max(pb->max_zrow, len);
                // a sub-solver may answer r_size, and local_affect then reads
                // the slot one past the row
                pb->tail_rows++;
            }
            if (zt == Z_EXHAUSTIVE || zt == Z_UNCONSTRAINED_EXHAUSTIVE) {
                // exhaustive_solver::build_constraints, exhaustive-solver.hpp:
                // 111-167: a binary counter over the row's elements, the last
                // element being the lowest digit.  The states met WHILE a
                // carry propagates are tested too, so a pattern can be listed
                // more than once; the list is kept as the reference builds it
                // because its order breaks ties between patterns.
                pb->ex_start[k] = static_cast<int>(pb->ex_flat.size());
                std::vector<int> digit(len, 0);
                digit[len - 1] = 1;
                bool end = false;
                while (!end) {
                    int i = len - 1;
                    for (;;) {
                        int sum = 0, mask = 0;
                        for (int jx = 0; jx != len; ++jx)
                            if (digit[jx]) {
                                sum += f[jx];
                                mask |= 1 << jx;
                            }
                        if (pb->bmin[k] <= sum && sum <= pb->bmax[k])
                            pb->ex_flat.push_back(mask);
                        if (++digit[i] > 1) {
                            digit[i] = 0;
                            if (i == 0) {
                                end = true;
                                break;
                            }
                            --i;
                        } else
                            break;
                    }
                }
                pb->ex_count[k] = static_cast<int>(pb->ex_flat.size()) - pb->ex_start[k];
                if (pb->ex_count[k] == 0)
                    return fail(BXG_E_UNREALISABLE,
                                "constraint without any feasible assignment (exhaustive row)");
            }
        } break;
        }

        if (pb->bmin[k] > pb->bmax[k])
            return fail(BXG_E_UNREALISABLE, "constraint with bound_min > bound_max");
        // by how much the row's activity (between `lower` and `upper`) can miss its bounds
        pb->max_amount = std::max(
          pb->max_amount,
          static_cast<int>(std::min<long long>(
            std::max<long long>(static_cast<long long>(pb->bmin[k]) - lower,
                                static_cast<long long>(upper) - pb->bmax[k]),
            std::numeric_limits<int>::max())));

        if (pb->bmin[k] != pb->bmax[k] && pb->family != FAM_Z &&
            std::min(pb->bmax[k] + (pb->family == FAM_101 ? neg : 0), len) == len)
            pb->tail_rows++;

        // algorithmic bytes of one update
        double S = 0;
        for (int e = row_ptr[k]; e != row_ptr[k + 1]; ++e)
            S += pb->col_ptr[col[e] + 1] - pb->col_ptr[col[e]];
        double bytes = 32.0 * len + 24.0 * S + 16.0;
        if (pb->family == FAM_101)
            bytes += std::ceil((len + S) / 8.0);
        else if (pb->family == FAM_Z)
            bytes += 4.0 * (len + S) + 24.0 * S;
        pb->row_bytes[k] = static_cast<float>(bytes);
    }

    pb->row_hdr.resize(m);
    {
        std::vector<int> lens(m);
        for (int k = 0; k != m; ++k) {
            RowHdr& h = pb->row_hdr[k];
            h.r0 = row_ptr[k];
            h.r_size = row_ptr[k + 1] - row_ptr[k];
            h.g0 = pb->row_gptr[k];
            h.S = pb->row_gptr[k + 1] - pb->row_gptr[k];
            h.bmin = pb->bmin[k];
            h.bmax = pb->bmax[k];
            h.c_size = pb->row_neg[k];
            h.bytes = pb->row_bytes[k];
            lens[k] = h.r_size;
        }
        std::sort(lens.begin(), lens.end());
        pb->p95_row = lens[std::min<size_t>(m - 1, static_cast<size_t>(0.95 * m))];
    }

    // default_cost_type constructor, itm-common.hpp:1006-1016
    pb->obj_var.assign(obj_var, obj_var + n_obj);
    pb->obj_factor.assign(obj_factor, obj_factor + n_obj);
    pb->cost.assign(n, 0.0);
    for (int i = 0; i != n_obj; ++i) {
        if (obj_var[i] < 0 || obj_var[i] >= n)
            return fail(BXG_E_INVALID, "objective variable out of range");
        pb->cost[obj_var[i]] += obj_factor[i];
    }

    // The objective is summed in any order on the device when that is exact:
    // all terms integers and the total magnitude below 2^53.
    {
        double total = std::fabs(cost_constant);
        bool integral = std::nearbyint(cost_constant) == cost_constant;
        for (int i = 0; i != n_obj && integral; ++i) {
            integral = std::nearbyint(obj_factor[i]) == obj_factor[i];
            total += std::fabs(obj_factor[i]);
        }
        pb->obj_any_order = integral && total < 9007199254740992.0;
    }

    CUDA_TRY(pb->d_row_ptr.upload(pb->row_ptr));
    CUDA_TRY(pb->d_row_col.upload(pb->row_col));
    CUDA_TRY(pb->d_row_cscpos.upload(pb->row_cscpos));
    CUDA_TRY(pb->d_col_ptr.upload(pb->col_ptr));
    CUDA_TRY(pb->d_csc_row.upload(pb->csc_row));
    CUDA_TRY(pb->d_csc_rs.upload(pb->csc_rs));
    CUDA_TRY(pb->d_row_info.upload(pb->row_info));
    CUDA_TRY(pb->d_gp.upload(pb->gp));
    CUDA_TRY(pb->d_row_hdr.upload(pb->row_hdr));
    CUDA_TRY(pb->d_row_gptr.upload(pb->row_gptr));
    CUDA_TRY(pb->d_row_eptr.upload(pb->row_eptr));
    if (pb->family != FAM_01) {
        CUDA_TRY(pb->d_row_a.upload(pb->row_a));
        CUDA_TRY(pb->d_csc_a.upload(pb->csc_a));
    }
    CUDA_TRY(pb->d_bmin.upload(pb->bmin));
    CUDA_TRY(pb->d_bmax.upload(pb->bmax));
    CUDA_TRY(pb->d_row_neg.upload(pb->row_neg));
    CUDA_TRY(pb->d_row_bytes.upload(pb->row_bytes));
    CUDA_TRY(pb->d_ztype.upload(pb->ztype));
    if (pb->family == FAM_Z) {
        CUDA_TRY(pb->d_ex_start.upload(pb->ex_start));
        CUDA_TRY(pb->d_ex_count.upload(pb->ex_count));
        CUDA_TRY(pb->d_ex_flat.upload(pb->ex_flat));
    }
    CUDA_TRY(pb->d_obj_var.upload(pb->obj_var));
    CUDA_TRY(pb->d_obj_factor.upload(pb->obj_factor));
    CUDA_TRY(pb->d_cost.upload(pb->cost));

    DevProblem& d = pb->dev;
    d.m = m;
    d.n = n;
    d.nnz = nnz;
    d.max_row = pb->max_row;
    d.max_col = pb->max_col;
    d.p95_row = pb->p95_row;
    d.row_hdr = pb->d_row_hdr.p;
    d.max_S = pb->max_S;
    d.gp = pb->d_gp.p;
    d.row_gptr = pb->d_row_gptr.p;
    d.row_eptr = pb->d_row_eptr.p;
    d.family = pb->family;
    d.maximize = pb->maximize;
    d.xw32 = 2 * (n / 64 + 1);
    d.n_obj = n_obj;
    d.obj_any_order = pb->obj_any_order;
    d.need_z = pb->family == FAM_Z;
    d.tail_rows = pb->tail_rows;
    d.cost_constant = cost_constant;
    d.row_ptr = pb->d_row_ptr.p;
    d.row_col = pb->d_row_col.p;
    d.row_cscpos = pb->d_row_cscpos.p;
    d.row_info = pb->d_row_info.p;
    d.csc_rs = pb->d_csc_rs.p;
    d.row_a = pb->family != FAM_01 ? pb->d_row_a.p : nullptr;
    d.col_ptr = pb->d_col_ptr.p;
    d.csc_row = pb->d_csc_row.p;
    d.csc_a = pb->family != FAM_01 ? pb->d_csc_a.p : nullptr;
    d.bmin = pb->d_bmin.p;
    d.bmax = pb->d_bmax.p;
    d.row_neg = pb->d_row_neg.p;
    d.row_bytes = pb->d_row_bytes.p;
    d.ztype = pb->d_ztype.p;
    d.max_zrow = pb->max_zrow;
    d.max_amount = pb->max_amount;
    d.ex_start = pb->family == FAM_Z ? pb->d_ex_start.p : nullptr;
    d.ex_count = pb->family == FAM_Z ? pb->d_ex_count.p : nullptr;
    d.ex_flat = pb->family == FAM_Z ? pb->d_ex_flat.p : nullptr;
    d.obj_var = pb->d_obj_var.p;
    d.obj_factor = pb->d_obj_factor.p;

    CUDA_TRY(cudaStreamSynchronize(nullptr)); // allocations complete before other streams use them
    *out = pb.release();
    return BXG_OK;
}

void
bxg_problem_destroy(bxg_problem* pb)
{
    if (!pb)
        return;
    cudaSetDevice(pb->device);
    delete pb;
}

int
bxg_problem_info(const bxg_problem* pb, int* m, int* n, int* nnz, int* max_row,
                 int* problem_type)
{
    if (!pb)
        return fail(BXG_E_INVALID, "null problem");
    if (m)
        *m = pb->m;
    if (n)
        *n = pb->n;
    if (nnz)
        *nnz = pb->nnz;
    if (max_row)
        *max_row = pb->max_row;
    if (problem_type)
        *problem_type = pb->problem_type;
    return BXG_OK;
}

static int batch_create_body(bxg_problem* pb, int n_lanes, const bxg_params* p, bxg_batch** out);

int
bxg_batch_create(bxg_problem* pb, int n_lanes, const bxg_params* p, bxg_batch** out)
{
    return guarded("bxg_batch_create", [&] { return batch_create_body(pb, n_lanes, p, out); });
}

static int
batch_create_body(bxg_problem* pb, int n_lanes, const bxg_params* p, bxg_batch** out)
{
    if (!pb || !p || !out || n_lanes <= 0)
        return fail(BXG_E_INVALID, "bxg_batch_create: bad argument");
    if (p->float_type != 1)
        return fail(BXG_E_UNSUPPORTED, "only float_type = double_type (deterministic fp64 mode)");
    if (p->order < ORDER_NONE || p->order > ORDER_CYCLE)
        return fail(BXG_E_INVALID, "unknown constraint order");
    if (p->order == ORDER_PI_SIGN_CHANGE && pb->n > pb->m)
        return fail(BXG_E_UNSUPPORTED,
                    "pi_sign_change iterates over n entries (itm-common.hpp:669): undefined in "
                    "the reference when n > m");
    if (p->order == ORDER_CYCLE && pb->n > pb->m)
        return fail(BXG_E_UNSUPPORTED, "cycle reaches pi_sign_change: undefined when n > m");
    if (p->limit == std::numeric_limits<int64_t>::max() &&
        !(p->time_limit > 0 && std::isfinite(p->time_limit)) && !(p->kappa_step > 0))
        return fail(BXG_E_INVALID,
                    "no limit, no time_limit and kappa_step = 0: a run that cannot end (the reference "
                    "would spin too); give one of them");

    CUDA_TRY(cudaSetDevice(pb->device));

    std::unique_ptr<bxg_batch> b(new (std::nothrow) bxg_batch);
    if (!b)
        return fail(BXG_E_MEMORY, "out of host memory");
    b->pb = pb;
    b->p = *p;
    b->n_lanes = n_lanes;
    const int n = pb->n, m = pb->m;

    // normalize_costs, itm-common.hpp:967-998 + make_l1/l2/loo_norm :1095-1127
    b->norm_cost = pb->cost;
    {
        std::vector<double>& c = b->norm_cost;
        double div = 0;
        switch (p->cost_norm) {
        case 0:
            break;
        case 1:
            return fail(BXG_E_UNSUPPORTED, "cost_norm random is not part of this engine");
        case 2:
            for (int i = 0; i != n; ++i)
                div += std::abs(c[i]);
            if (std::isnormal(div))
                for (int i = 0; i != n; ++i)
                    c[i] /= div;
            break;
        case 3:
            for (int i = 0; i != n; ++i)
                div += c[i] * c[i];
            if (std::isnormal(div))
                for (int i = 0; i != n; ++i)
                    c[i] /= div;
            break;
        default:
            div = *std::max_element(c.begin(), c.end());
            if (std::isnormal(div))
                for (int i = 0; i != n; ++i)
                    c[i] /= div;
            break;
        }
    }

    // compute_delta, itm-common.hpp:917-933
    b->theta = p->theta;
    if (p->delta < 0) {
        double mini = std::numeric_limits<double>::max();
        for (int i = 0; i != n; ++i)
            if (b->norm_cost[i])
                mini = std::min(mini, std::abs(b->norm_cost[i]));
        b->delta = mini - b->theta * mini;
    } else
        b->delta = p->delta;

    // kappa increments use std::pow(remaining / m, alpha) in the reference
    // (itm-solver-common.hpp:152-155); remaining is an integer in [0, m], so
    // the host's libm tabulates every value the device can need.
    b->powtab.resize(m + 1);
    for (int r = 0; r <= m; ++r)
        b->powtab[r] = std::pow(static_cast<double>(r) / static_cast<double>(m), p->alpha);

    b->order_cap =
      (p->order == ORDER_PI_SIGN_CHANGE || p->order == ORDER_CYCLE) ? std::max(n, m) + 1 : m + 1;

    const size_t L = static_cast<size_t>(n_lanes);
    CUDA_TRY(b->d_P.alloc(L * pb->nnz));
    CUDA_TRY(b->d_pi.alloc(L * m));
    CUDA_TRY(b->d_x.alloc(L * pb->dev.xw32));
    CUDA_TRY(b->d_bestx.alloc(L * pb->dev.xw32));
    CUDA_TRY(b->d_Rv.alloc(L * (pb->max_row + 2)));
    CUDA_TRY(b->d_act.alloc(L * m));
    CUDA_TRY(cudaMemset(b->d_act.p, 0, b->d_act.bytes()));
    CUDA_TRY(b->d_oR.alloc(L * b->order_cap));
    CUDA_TRY(b->d_mo.alloc(L * b->order_cap));
    CUDA_TRY(b->d_ctl.alloc(L));
    CUDA_TRY(b->d_kappa.alloc(L));
    CUDA_TRY(b->d_seeds.alloc(L));
    CUDA_TRY(b->d_rng0.alloc(L));
    CUDA_TRY(b->d_norm_cost.upload(b->norm_cost));
    CUDA_TRY(b->d_powtab.upload(b->powtab));
    CUDA_TRY(cudaMemset(b->d_ctl.p, 0, sizeof(LaneCtl) * L));
    CUDA_TRY(cudaMemset(b->d_bestx.p, 0, b->d_bestx.bytes()));

    b->cfg = plan_launch(pb->dev, pb->device, n_lanes,
                         p->order == ORDER_INFEASIBILITY_DECR || p->order == ORDER_INFEASIBILITY_INCR ||
                           p->order == ORDER_CYCLE);
    {
        // the row activities (4 m bytes) and the lane's control blocks always
        // live in shared memory: refuse what cannot fit instead of failing at
        // the first launch
        int max_optin = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, pb->device));
        if (b->cfg.smem > static_cast<size_t>(max_optin))
            return fail(BXG_E_UNSUPPORTED,
                        "problem too tall for one CTA's shared memory (row activities: 4 bytes per row)");
    }
    if (!b->cfg.r_in_smem)
        CUDA_TRY(b->d_scratch.alloc(L * lane_scratch_bytes(pb->dev, b->cfg)));
    const size_t zbytes = z_scratch_bytes(pb->max_zrow);
    if (zbytes) {
        if (zbytes > (size_t{ 1 } << 26))
            return fail(BXG_E_MEMORY, "branch-and-bound row too long for the device scratch");
        CUDA_TRY(b->d_zscratch.alloc(L * zbytes));
    }

    DevBatch& d = b->dev;
    d.n_lanes = n_lanes;
    d.order_cap = b->order_cap;
    d.P = b->d_P.p;
    d.pi = b->d_pi.p;
    d.x = b->d_x.p;
    d.bestx = b->d_bestx.p;
    d.Rv = b->d_Rv.p;
    d.act = b->d_act.p;
    d.oR = b->d_oR.p;
    d.mo = b->d_mo.p;
    d.ctl = b->d_ctl.p;
    d.kappa = b->d_kappa.p;
    d.lane_par = nullptr;
    d.g_scratch = b->cfg.r_in_smem ? nullptr : b->d_scratch.p;
    d.z_scratch = zbytes ? b->d_zscratch.p : nullptr;
    d.z_bytes = zbytes;

    refresh_params(b.get());

    CUDA_TRY(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&b->ev0));
    CUDA_TRY(cudaEventCreate(&b->ev1));

    CUDA_TRY(cudaStreamSynchronize(nullptr)); // allocations complete before the batch's stream uses them
    *out = b.release();
    return BXG_OK;
}

void
bxg_batch_destroy(bxg_batch* b)
{
    if (!b)
        return;
    cudaSetDevice(b->pb->device);
    delete b;
}

int
bxg_batch_info(const bxg_batch* b, int* n_lanes, double* delta, double* theta,
               int64_t* device_bytes)
{
    if (!b)
        return fail(BXG_E_INVALID, "null batch");
    if (n_lanes)
        *n_lanes = b->n_lanes;
    if (delta)
        *delta = b->delta;
    if (theta)
        *theta = b->theta;
    if (device_bytes)
        *device_bytes = static_cast<int64_t>(
          b->d_P.bytes() + b->d_pi.bytes() + b->d_x.bytes() + b->d_bestx.bytes() +
          b->d_Rv.bytes() + b->d_act.bytes() + b->d_oR.bytes() + b->d_mo.bytes() + b->d_ctl.bytes() +
          b->d_scratch.bytes());
    return BXG_OK;
}

int
bxg_batch_set_tail_value(bxg_batch* b, double v)
{
    if (!b)
        return fail(BXG_E_INVALID, "null batch");
    b->tail_value = v;
    refresh_params(b);
    return BXG_OK;
}

int
bxg_batch_words32(const bxg_batch* b)
{
    return b ? b->pb->dev.xw32 : 0;
}

} // extern "C"

namespace {

// init_with_pre_solve (itm-common.hpp:284-374) on the host for the two
// pre-solve policies; like the reference it runs on the standard library's
// shuffle / sort / bernoulli on std::default_random_engine.
void
host_pre_solve(const bxg_problem& pb, bool optimistic, std::default_random_engine& rng,
               std::vector<uint64_t>& x)
{
    struct rc
    {
        float value;
        int factor;
        int id;
    };
    std::vector<rc> R(pb.max_row);
    std::bernoulli_distribution dist(1.0);
    auto set = [&](int i) { x[i / 64] |= uint64_t{ 1 } << (i % 64); };
    auto unset = [&](int i) { x[i / 64] &= ~(uint64_t{ 1 } << (i % 64)); };
    const bool maximize = pb.maximize != 0;

    for (int k = 0; k != pb.m; ++k) {
        if (!dist(rng))
            continue;

        const int r_size = pb.row_ptr[k + 1] - pb.row_ptr[k];
        R.resize(r_size);
        for (int i = 0; i != r_size; ++i) {
            const int e = pb.row_ptr[k] + i;
            R[i].value = static_cast<float>(pb.cost[pb.row_col[e]]);
            R[i].factor = pb.row_a[e];
            R[i].id = pb.row_col[e];
        }

        std::shuffle(R.begin(), R.end(), rng);
        if (maximize)
            std::sort(R.begin(), R.end(),
                      [](const auto& l, const auto& r) { return l.value > r.value; });
        else
            std::sort(R.begin(), R.end(),
                      [](const auto& l, const auto& r) { return l.value < r.value; });

        int sum = 0;
        int best = -2;
        if (!optimistic) {
            for (int i = -1; i < r_size; ++i) {
                if (pb.cmin[k] <= sum && sum <= pb.cmax[k]) {
                    best = i;
                    break;
                }
                if (i + 1 < r_size)
                    sum += R[i + 1].factor;
            }
        } else {
            for (int i = -1; i < r_size; ++i) {
                if (pb.cmin[k] <= sum && sum <= pb.cmax[k])
                    best = i;
                if (best != -2 && i + 1 < r_size &&
                    (maximize ? R[i + 1].value < 0 : R[i + 1].value > 0))
                    break;
                if (i + 1 < r_size)
                    sum += R[i + 1].factor;
            }
        }

        int i = 0;
        for (; i <= best; ++i)
            set(R[i].id);
        for (; i != r_size; ++i)
            unset(R[i].id);
    }
}

int
run_begin(bxg_batch* b, const uint32_t* h_seeds, const uint32_t* d_x0, const uint32_t* d_rng,
          int reset_solver)
{
    bxg_problem* pb = b->pb;
    CUDA_TRY(cudaSetDevice(pb->device));
    const int L = b->n_lanes;
    int init_on_device = 0;

    if (h_seeds) {
        if (b->p.init_policy == 0) {
            CUDA_TRY(cudaMemcpyAsync(b->d_seeds.p, h_seeds, sizeof(uint32_t) * L,
                                     cudaMemcpyHostToDevice, b->stream));
            init_on_device = 1;
        } else {
            // pessimistic_solve / optimistic_solve: host initialisation
            const int words = pb->n / 64 + 1;
            std::vector<uint32_t> x0(static_cast<size_t>(L) * pb->dev.xw32);
            std::vector<uint32_t> rng0(L);
            for (int l = 0; l != L; ++l) {
                std::default_random_engine rng(h_seeds[l]);
                std::vector<uint64_t> x(words, 0);
                host_pre_solve(*pb, b->p.init_policy == 2, rng, x);
                std::bernoulli_distribution choose_mutation(
                  b->lane_par.empty() ? b->p.init_policy_random : b->lane_par[5 * size_t(l) + 4]);
                for (int i = 0; i != pb->n; ++i)
                    if (choose_mutation(rng))
                        x[i / 64] ^= uint64_t{ 1 } << (i % 64);
                std::memcpy(&x0[static_cast<size_t>(l) * pb->dev.xw32], x.data(),
                            sizeof(uint64_t) * words);
                // minstd_rand0 streams its state as a decimal number
                std::ostringstream os;
                os << rng;
                rng0[l] = static_cast<uint32_t>(std::stoul(os.str()));
            }
            CUDA_TRY(b->d_x0.alloc(x0.size()));
            CUDA_TRY(cudaStreamSynchronize(nullptr)); // the batch's stream is non-blocking
            CUDA_TRY(cudaMemcpyAsync(b->d_x0.p, x0.data(), sizeof(uint32_t) * x0.size(),
                                     cudaMemcpyHostToDevice, b->stream));
            CUDA_TRY(cudaMemcpyAsync(b->d_rng0.p, rng0.data(), sizeof(uint32_t) * L,
                                     cudaMemcpyHostToDevice, b->stream));
            CUDA_TRY(cudaStreamSynchronize(b->stream));
            d_x0 = b->d_x0.p;
            d_rng = b->d_rng0.p;
        }
    }

    refresh_params(b);
    CUDA_TRY(cudaEventRecord(b->ev0, b->stream));
    launch_begin(pb->dev, b->prm, b->dev, b->d_seeds.p, d_x0, d_rng, reset_solver, init_on_device,
                 b->p.init_policy_random, pb->d_cost.p, b->cfg, b->stream);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(b->ev1, b->stream));
    b->launches += 1;
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    CUDA_TRY(cudaEventElapsedTime(&b->last_ms, b->ev0, b->ev1));
    b->total_ms += b->last_ms;
    b->begun = true;
    return BXG_OK;
}

} // namespace

extern "C" {

int
bxg_batch_begin(bxg_batch* b, const uint32_t* seeds, int reset_solver)
{
    if (!b || !seeds)
        return fail(BXG_E_INVALID, "bxg_batch_begin: bad argument");
    return run_begin(b, seeds, nullptr, nullptr, reset_solver);
}

int
bxg_batch_begin_device(bxg_batch* b, const uint32_t* d_x0, const uint32_t* d_rng,
                       int reset_solver)
{
    if (!b || !d_x0 || !d_rng)
        return fail(BXG_E_INVALID, "bxg_batch_begin_device: bad argument");
    return run_begin(b, nullptr, d_x0, d_rng, reset_solver);
}

int
bxg_batch_get_state(bxg_batch* b, int lane, uint64_t* x_words, double* pi, double* P,
                    uint32_t* rng_state, int32_t* order_list, int32_t* order_len,
                    double* norm_cost)
{
    if (!b || lane < 0 || lane >= b->n_lanes)
        return fail(BXG_E_INVALID, "bxg_batch_get_state: bad argument");
    bxg_problem* pb = b->pb;
    CUDA_TRY(cudaSetDevice(pb->device));
    CUDA_TRY(cudaStreamSynchronize(b->stream));

    if (x_words)
        CUDA_TRY(cudaMemcpy(x_words, b->d_x.p + static_cast<size_t>(lane) * pb->dev.xw32,
                            sizeof(uint32_t) * pb->dev.xw32, cudaMemcpyDeviceToHost));
    if (pi)
        CUDA_TRY(cudaMemcpy(pi, b->d_pi.p + static_cast<size_t>(lane) * pb->m,
                            sizeof(double) * pb->m, cudaMemcpyDeviceToHost));
    if (P) {
        std::vector<double> csc(pb->nnz);
        CUDA_TRY(cudaMemcpy(csc.data(), b->d_P.p + static_cast<size_t>(lane) * pb->nnz,
                            sizeof(double) * pb->nnz, cudaMemcpyDeviceToHost));
        for (int e = 0; e != pb->nnz; ++e)
            P[e] = csc[pb->row_cscpos[e]];
    }
    if (norm_cost)
        std::copy(b->norm_cost.begin(), b->norm_cost.end(), norm_cost);

    LaneCtl ctl;
    CUDA_TRY(cudaMemcpy(&ctl, b->d_ctl.p + lane, sizeof(LaneCtl), cudaMemcpyDeviceToHost));
    if (rng_state)
        *rng_state = ctl.rng;

    const bool infeas = ctl.order_state == ORDER_INFEASIBILITY_DECR ||
                        ctl.order_state == ORDER_INFEASIBILITY_INCR;
    if (order_len)
        *order_len = infeas ? ctl.mo_len : ctl.oR_len;
    if (order_list) {
        if (infeas)
            CUDA_TRY(cudaMemcpy(order_list, b->d_mo.p + static_cast<size_t>(lane) * b->order_cap,
                                sizeof(int2) * ctl.mo_len, cudaMemcpyDeviceToHost));
        else
            CUDA_TRY(cudaMemcpy(order_list, b->d_oR.p + static_cast<size_t>(lane) * b->order_cap,
                                sizeof(int) * ctl.oR_len, cudaMemcpyDeviceToHost));
    }
    return BXG_OK;
}

int
bxg_batch_set_state(bxg_batch* b, int lane, const uint64_t* x_words, const double* pi,
                    const double* P, const uint32_t* rng_state)
{
    if (!b || lane < 0 || lane >= b->n_lanes)
        return fail(BXG_E_INVALID, "bxg_batch_set_state: bad argument");
    bxg_problem* pb = b->pb;
    CUDA_TRY(cudaSetDevice(pb->device));
    CUDA_TRY(cudaStreamSynchronize(b->stream));

    if (pi)
        CUDA_TRY(cudaMemcpy(b->d_pi.p + static_cast<size_t>(lane) * pb->m, pi,
                            sizeof(double) * pb->m, cudaMemcpyHostToDevice));
    if (P) {
        std::vector<double> csc(pb->nnz);
        for (int e = 0; e != pb->nnz; ++e)
            csc[pb->row_cscpos[e]] = P[e];
        CUDA_TRY(cudaMemcpy(b->d_P.p + static_cast<size_t>(lane) * pb->nnz, csc.data(),
                            sizeof(double) * pb->nnz, cudaMemcpyHostToDevice));
    }
    if (rng_state)
        CUDA_TRY(cudaMemcpy(reinterpret_cast<char*>(b->d_ctl.p + lane) + offsetof(LaneCtl, rng),
                            rng_state, sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (x_words) {
        CUDA_TRY(cudaMemcpy(b->d_x.p + static_cast<size_t>(lane) * pb->dev.xw32, x_words,
                            sizeof(uint32_t) * pb->dev.xw32, cudaMemcpyHostToDevice));
        refresh_params(b);
        launch_order_init(pb->dev, b->prm, b->dev, lane, 1, b->cfg, b->stream);
        CUDA_TRY(cudaGetLastError());
        b->launches += 1;
        CUDA_TRY(cudaStreamSynchronize(b->stream));
    }
    return BXG_OK;
}

int
bxg_batch_sweep(bxg_batch* b, int push, const double* kappa, double objective_amplifier,
                int32_t* remaining)
{
    if (!b || !kappa)
        return fail(BXG_E_INVALID, "bxg_batch_sweep: bad argument");
    if (!b->begun)
        return fail(BXG_E_INVALID, "bxg_batch_sweep: call bxg_batch_begin first");
    bxg_problem* pb = b->pb;
    CUDA_TRY(cudaSetDevice(pb->device));
    CUDA_TRY(cudaMemcpyAsync(b->d_kappa.p, kappa, sizeof(double) * b->n_lanes,
                             cudaMemcpyHostToDevice, b->stream));
    refresh_params(b);
    CUDA_TRY(cudaEventRecord(b->ev0, b->stream));
    launch_sweep(pb->dev, b->prm, b->dev, push, objective_amplifier, b->cfg, b->stream);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(b->ev1, b->stream));
    b->launches += 1;
    int rc = fetch_ctl(b);
    if (rc != BXG_OK)
        return rc;
    CUDA_TRY(cudaEventElapsedTime(&b->last_ms, b->ev0, b->ev1));
    b->total_ms += b->last_ms;
    if (remaining)
        for (int l = 0; l != b->n_lanes; ++l)
            remaining[l] = b->h_ctl[l].remaining;
    return BXG_OK;
}

int
bxg_batch_objective(bxg_batch* b, double* value)
{
    if (!b || !value)
        return fail(BXG_E_INVALID, "bxg_batch_objective: bad argument");
    bxg_problem* pb = b->pb;
    CUDA_TRY(cudaSetDevice(pb->device));
    refresh_params(b);
    launch_objective(pb->dev, b->prm, b->dev, b->cfg, b->stream);
    CUDA_TRY(cudaGetLastError());
    b->launches += 1;
    int rc = fetch_ctl(b);
    if (rc != BXG_OK)
        return rc;
    for (int l = 0; l != b->n_lanes; ++l)
        value[l] = b->h_ctl[l].objective;
    return BXG_OK;
}

int
bxg_batch_solve(bxg_batch* b, int32_t* status, double* value, int32_t* remaining, int64_t* loop)
{
    if (!b)
        return fail(BXG_E_INVALID, "null batch");
    if (!b->begun)
        return fail(BXG_E_INVALID, "bxg_batch_solve: call bxg_batch_begin first");
    bxg_problem* pb = b->pb;
    CUDA_TRY(cudaSetDevice(pb->device));
    if (b->p.time_limit > 0 && std::isfinite(b->p.time_limit)) {
        // solver_functor's is_timelimit_reached (itm-solver-common.hpp:162,186,210,
        // 228-232): the lanes test the device clock once per sweep
        unsigned long long now = 0;
        int rc_ = device_clock(b->stream, &now);
        if (rc_ != BXG_OK)
            return rc_;
        b->deadline_ns = now + static_cast<unsigned long long>(b->p.time_limit * 1e9);
    } else
        b->deadline_ns = 0;
    refresh_params(b);
    CUDA_TRY(cudaEventRecord(b->ev0, b->stream));
    launch_solve(pb->dev, b->prm, b->dev, b->cfg, b->stream);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(b->ev1, b->stream));
    b->launches += 1;
    int rc = fetch_ctl(b);
    if (rc != BXG_OK)
        return rc;
    CUDA_TRY(cudaEventElapsedTime(&b->last_ms, b->ev0, b->ev1));
    b->total_ms += b->last_ms;
    for (int l = 0; l != b->n_lanes; ++l) {
        const LaneCtl& c = b->h_ctl[l];
        if (status)
            status[l] = c.status;
        if (value)
            value[l] = c.best_has_x ? c.best_value : 0.0;
        if (remaining)
            remaining[l] = c.best_remaining;
        if (loop)
            loop[l] = c.best_loop;
    }
    return BXG_OK;
}

int
bxg_batch_best(bxg_batch* b, int lane, int8_t* x)
{
    if (!b || !x || lane < 0 || lane >= b->n_lanes)
        return fail(BXG_E_INVALID, "bxg_batch_best: bad argument");
    bxg_problem* pb = b->pb;
    CUDA_TRY(cudaSetDevice(pb->device));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    std::vector<uint32_t> w(pb->dev.xw32);
    LaneCtl ctl;
    CUDA_TRY(cudaMemcpy(&ctl, b->d_ctl.p + lane, sizeof(LaneCtl), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(w.data(), b->d_bestx.p + static_cast<size_t>(lane) * pb->dev.xw32,
                        sizeof(uint32_t) * pb->dev.xw32, cudaMemcpyDeviceToHost));
    for (int i = 0; i != pb->n; ++i)
        x[i] = ctl.best_has_x ? static_cast<int8_t>((w[i >> 5] >> (i & 31)) & 1u) : 0;
    return BXG_OK;
}

int
bxg_batch_argbest(bxg_batch* b, int* lane, double* value, int* remaining)
{
    if (!b)
        return fail(BXG_E_INVALID, "null batch");
    CUDA_TRY(cudaSetDevice(b->pb->device));
    int rc = fetch_ctl(b);
    if (rc != BXG_OK)
        return rc;
    const bool maximize = b->pb->maximize != 0;
    int best = 0;
    for (int l = 1; l < b->n_lanes; ++l) {
        const LaneCtl& c = b->h_ctl[l];
        const LaneCtl& o = b->h_ctl[best];
        // raw_result_compare, lib/src/result.hpp:96-107
        bool better;
        if (c.best_remaining == 0 && o.best_remaining == 0)
            better = maximize ? c.best_value > o.best_value : c.best_value < o.best_value;
        else
            better = c.best_remaining < o.best_remaining;
        if (better)
            best = l;
    }
    if (lane)
        *lane = best;
    if (value)
        *value = b->h_ctl[best].best_value;
    if (remaining)
        *remaining = b->h_ctl[best].best_remaining;
    return BXG_OK;
}

int
bxg_batch_counters(bxg_batch* b, int64_t* updates, int64_t* sweeps, double* algorithmic_bytes)
{
    if (!b)
        return fail(BXG_E_INVALID, "null batch");
    CUDA_TRY(cudaSetDevice(b->pb->device));
    int rc = fetch_ctl(b);
    if (rc != BXG_OK)
        return rc;
    int64_t u = 0, s = 0;
    double bytes = 0;
    for (const LaneCtl& c : b->h_ctl) {
        u += static_cast<int64_t>(c.updates);
        s += static_cast<int64_t>(c.sweeps);
        bytes += c.alg_bytes;
    }
    if (updates)
        *updates = u;
    if (sweeps)
        *sweeps = s;
    if (algorithmic_bytes)
        *algorithmic_bytes = bytes;
    return BXG_OK;
}

int
bxg_batch_reset_counters(bxg_batch* b)
{
    if (!b)
        return fail(BXG_E_INVALID, "null batch");
    CUDA_TRY(cudaSetDevice(b->pb->device));
    int rc = fetch_ctl(b);
    if (rc != BXG_OK)
        return rc;
    for (LaneCtl& c : b->h_ctl) {
        c.updates = 0;
        c.sweeps = 0;
        c.offers = 0;
        c.inserts = 0;
        c.alg_bytes = 0;
        for (int i = 0; i < 4; ++i) {
            c.n_path[i] = 0;
            c.cyc_path[i] = 0;
            c.cyc_phase[i] = 0;
        }
    }
    CUDA_TRY(cudaMemcpy(b->d_ctl.p, b->h_ctl.data(), sizeof(LaneCtl) * b->n_lanes,
                        cudaMemcpyHostToDevice));
    b->total_ms = 0;
    return BXG_OK;
}

int
bxg_batch_path_counters(bxg_batch* b, int64_t* updates, int64_t* cycles, int64_t* phase_cycles)
{
    if (!b)
        return fail(BXG_E_INVALID, "null batch");
    CUDA_TRY(cudaSetDevice(b->pb->device));
    int rc = fetch_ctl(b);
    if (rc != BXG_OK)
        return rc;
    for (int i = 0; i < 4; ++i) {
        int64_t u = 0, c = 0;
        for (const LaneCtl& l : b->h_ctl) {
            u += static_cast<int64_t>(l.n_path[i]);
            c += static_cast<int64_t>(l.cyc_path[i]);
        }
        if (updates)
            updates[i] = u;
        if (cycles)
            cycles[i] = c;
        if (phase_cycles) {
            int64_t ph = 0;
            for (const LaneCtl& l : b->h_ctl)
                ph += static_cast<int64_t>(l.cyc_phase[i]);
            phase_cycles[i] = ph;
        }
    }
    return BXG_OK;
}

static int batch_set_lane_params_body(bxg_batch* b, const double* theta, const double* delta,
                          const double* kappa_min, const double* kappa_step,
                          const double* init_policy_random);

int
bxg_batch_set_lane_params(bxg_batch* b, const double* theta, const double* delta,
                          const double* kappa_min, const double* kappa_step,
                          const double* init_policy_random)
{
    return guarded("bxg_batch_set_lane_params", [&] { return batch_set_lane_params_body(b, theta, delta, kappa_min, kappa_step, init_policy_random); });
}

static int
batch_set_lane_params_body(bxg_batch* b, const double* theta, const double* delta,
                          const double* kappa_min, const double* kappa_step,
                          const double* init_policy_random)
{
    if (!b)
        return fail(BXG_E_INVALID, "null batch");
    CUDA_TRY(cudaSetDevice(b->pb->device));
    if (!theta && !delta && !kappa_min && !kappa_step && !init_policy_random) {
        b->lane_par.clear();
        b->dev.lane_par = nullptr;
        return BXG_OK;
    }
    const int L = b->n_lanes;
    // the smallest non-zero normalised cost, for the automatic delta
    double mini = std::numeric_limits<double>::max();
    for (double c : b->norm_cost)
        if (c)
            mini = std::min(mini, std::abs(c));

    b->lane_par.assign(5 * static_cast<size_t>(L), 0.0);
    for (int l = 0; l != L; ++l) {
        // every lane goes through context_set_solver_parameters' clamps
        // (private.cpp:51-172) like a run of its own
        bxg_params in = b->p, out;
        if (theta)
            in.theta = theta[l];
        if (delta)
            in.delta = delta[l];
        if (kappa_min)
            in.kappa_min = kappa_min[l];
        if (kappa_step)
            in.kappa_step = kappa_step[l];
        if (init_policy_random)
            in.init_policy_random = init_policy_random[l];
        // w was already turned into a sweep count by the first clamp
        const double w = in.w;
        bxg_clamp_params(&in, &out);
        out.w = w;
        double* par = &b->lane_par[5 * static_cast<size_t>(l)];
        par[0] = out.theta;
        // compute_delta, itm-common.hpp:917-933, with the lane's theta
        par[1] = out.delta < 0 ? mini - out.theta * mini : out.delta;
        par[2] = out.kappa_min;
        par[3] = out.kappa_step;
        par[4] = out.init_policy_random;
    }
    CUDA_TRY(b->d_lane_par.upload(b->lane_par));
    b->dev.lane_par = b->d_lane_par.p;
    return BXG_OK;
}

// The manual auto-tuner's grid (manual-optimizer.cpp:31-111, 113-160).  Five
// dimensions of five values.  Entry 0 of a table is the caller's parameter,
// entry 1 is step_d / 5, and entry i >= 2 is theta's entry i - 1 plus an
// increment -- for EVERY dimension, theta's table included (that is how the
// reference fills them, :62-87): increments 1/5, 0.1/25, 1e-2/5, 1e-7/25 and
// 0.9/5.  The walk is an odometer over the five indices, the last one fastest.
int
bxg_manual_grid(const bxg_params* base, int64_t first, int64_t capacity, double* theta,
                double* delta, double* kappa_min, double* kappa_step, double* init_policy_random,
                int64_t* count, double* tables)
{
    if (!base || !count || first < 0 || capacity < 0)
        return fail(BXG_E_INVALID, "bxg_manual_grid: bad argument");
    constexpr int D = 5, V = 5;
    const double len = static_cast<double>(V);
    const double start[D] = { base->theta, base->delta, base->kappa_min, base->kappa_step,
                              base->init_policy_random };
    const double second[D] = { 1.0 / len, 0.1 / len, 1e-2 / len, 1e-7 / len, 0.9 / len };
    const double step[D] = { 1.0 / len, 0.1 / (5.0 * len), 1e-2 / len, 1e-7 / (5.0 * len), 0.9 / len };
    double tab[D][V];
    for (int d = 0; d != D; ++d) {
        tab[d][0] = start[d];
        tab[d][1] = second[d];
    }
    // theta first: the other tables read it
    for (int d = 0; d != D; ++d)
        for (int i = 2; i != V; ++i)
            tab[d][i] = tab[0][i - 1] + step[d];
    if (tables)
        for (int d = 0; d != D; ++d)
            for (int i = 0; i != V; ++i)
                tables[d * V + i] = tab[d][i];

    int64_t total = 1;
    for (int d = 0; d != D; ++d)
        total *= V;
    double* const out[D] = { theta, delta, kappa_min, kappa_step, init_policy_random };
    int64_t written = 0;
    for (int64_t p = first; p < total && written < capacity; ++p, ++written) {
        int64_t rest = p;
        for (int d = D - 1; d >= 0; --d) { // point p of the odometer: base-V digits, last fastest
            const int idx = static_cast<int>(rest % V);
            rest /= V;
            if (out[d])
                out[d][written] = tab[d][idx];
        }
    }
    *count = written;
    return BXG_OK;
}

int
bxg_batch_last_ms(bxg_batch* b, float* ms)
{
    if (!b || !ms)
        return fail(BXG_E_INVALID, "bxg_batch_last_ms: bad argument");
    *ms = b->last_ms;
    return BXG_OK;
}

int
bxg_batch_device_ms(bxg_batch* b, double* total_ms)
{
    if (!b || !total_ms)
        return fail(BXG_E_INVALID, "bxg_batch_device_ms: bad argument");
    *total_ms = b->total_ms;
    return BXG_OK;
}

int
bxg_batch_launches(bxg_batch* b, int64_t* launches)
{
    if (!b || !launches)
        return fail(BXG_E_INVALID, "bxg_batch_launches: bad argument");
    *launches = b->launches;
    return BXG_OK;
}

} // extern "C"

#include "bxg_lp.hpp"
#include "bxg_preprocess.hpp"
#include "bxg_population.cuh"
